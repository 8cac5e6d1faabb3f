import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from qsi_b200 import bfp_ops as ours, _lib
from _refload import load_reference
ref = load_reference()
kw = dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=7, weight_mant_bits=15,
          block_size=64, w_sparsity=True, N=2, M=4, first="s", sparsity_mode="structured", sparsity_frac=0.5, device="cuda")
torch.manual_seed(0)
lin = ours.BFPLinear(768, 3072, bias=True, **dict(kw)).cuda()
x = torch.randn(8, 512, 768, device="cuda")
with torch.no_grad():
    print("kind:", ours._tensor_core_kind(x, lin.weight, lin.bfp_args))
    n0 = _lib.launch_count()
    y = lin(x); torch.cuda.synchronize()
    print("launches first call:", _lib.launch_count() - n0)
    n0 = _lib.launch_count(); t0 = time.perf_counter()
    for _ in range(20): y = lin(x)
    torch.cuda.synchronize(); print("launches/call:", (_lib.launch_count() - n0) / 20, "ms/call:", (time.perf_counter() - t0) / 20 * 1e3)
    for kind in ("bf16", "i8", "sp"):
        os.environ["BFP_GEMM_KIND"] = kind
        y2 = lin(x); torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(20): y2 = lin(x)
        torch.cuda.synchronize(); print(kind, "ms/call:", (time.perf_counter() - t0) / 20 * 1e3, "rel vs default", float((y2 - y).norm() / y.norm()))
    os.environ.pop("BFP_GEMM_KIND")
    if ref is not None:
        rl = ref.BFPLinear(768, 3072, bias=True, **dict(kw)).cuda(); rl.weight, rl.bias = lin.weight, lin.bias
        yr = rl(x); torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5): yr = rl(x)
        torch.cuda.synchronize(); print("reference ms/call:", (time.perf_counter() - t0) / 5 * 1e3, "rel err ours vs ref:", float((y - yr).norm() / yr.norm()), "max abs", float((y - yr).abs().max()))
