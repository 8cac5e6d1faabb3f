"""BFPLinear with sparsity_num_format='int' (row f2) at LLaMA-7B shapes, fp32 modules, T = 4096 tokens: tensor-core path
(three-plane activations x integer weight grid) vs this repo's fake-quant + F.linear and the reference on the same GPU."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["BFP_INT_LINEAR"] = "tc"        # the path under test is opt-in (see bfp_ops._int_tc_eligible)
import torch
from qsi_b200 import bfp_ops as ours
from _refload import load_reference
ref = load_reference()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
res = []
for mode in ("structured", "unstructured"):
    kw = dict(num_format="bfp", sparsity_num_format="int", rounding_mode="determ", epsilon=1e-8, mant_bits=8, block_size=64, w_sparsity=True,
              N=2, M=4, first="s", sparsity_mode=mode, sparsity_frac=0.5, device="cuda")
    for N, K in [(4096, 4096), (11008, 4096), (4096, 11008)]:
        torch.manual_seed(0)
        lin = ours.BFPLinear(K, N, bias=True, **dict(kw)).cuda().eval()
        x = torch.randn(4096, K, device="cuda")
        with torch.no_grad():
            ms = t(lambda: lin(x)); y = lin(x)
            os.environ["BFP_LINEAR_PATH"] = "fakequant"; ms_fq = t(lambda: lin(x), 3); y_fq = lin(x); os.environ["BFP_LINEAR_PATH"] = "tc"
            row = {"sparsity": mode, "N": N, "K": K, "T": 4096, "tc_ms": ms, "tflops_dense_equiv": 2 * 4096 * N * K / ms / 1e9,
                   "fakequant_sgemm_ms": ms_fq, "rel_diff_vs_fakequant": float((y - y_fq).norm() / y_fq.norm())}
            if ref is not None:
                rl = ref.BFPLinear(K, N, bias=True, **dict(kw)).cuda(); rl.weight, rl.bias = lin.weight, lin.bias
                row["reference_ms"] = t(lambda: rl(x), 3); yr = rl(x)
                row["rel_diff_vs_reference"] = float((y - yr).norm() / yr.norm()); row["speedup_vs_reference"] = row["reference_ms"] / ms
        print(json.dumps(row), flush=True); res.append(row)
if len(sys.argv) > 1: json.dump(res, open(sys.argv[1], "w"), indent=1)
