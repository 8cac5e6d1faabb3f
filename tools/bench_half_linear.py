"""fp16 BFPLinear forward (how the reference runs LLaMA): tensor-core path with the dtype conversion fused into the epilogue vs the
fake-quant + library HGEMM structure, LLaMA-7B shapes, 4096 tokens."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qsi_b200 import bfp_ops as ops
kw = dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=7, block_size=64,
          w_sparsity=True, N=2, M=4, first="s", sparsity_mode="structured", device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
for dt in (torch.float16, torch.bfloat16):
    for (N, K) in ((4096, 4096), (11008, 4096), (4096, 11008)):
        lin = ops.BFPLinear(K, N, bias=False, **dict(kw)).cuda().to(dt).eval()
        xs = [torch.randn(4096, K, device="cuda").to(dt) for _ in range(3)]
        i = [0]
        def f():
            i[0] += 1
            return lin(xs[i[0] % 3])
        with torch.no_grad():
            os.environ["BFP_LINEAR_PATH"] = "tc"; ms_tc = t(f)
            os.environ["BFP_LINEAR_PATH"] = "fakequant"; ms_fq = t(f)
            os.environ["BFP_LINEAR_PATH"] = "tc"
        fl = 2.0 * 4096 * N * K
        print(f"{str(dt)[6:]} T=4096 N={N} K={K}: tensor-core path {ms_tc:.3f} ms = {fl/ms_tc/1e9:.0f} TFLOP/s | fake-quant + HGEMM {ms_fq:.3f} ms = {fl/ms_fq/1e9:.0f} | x{ms_fq/ms_tc:.2f}", flush=True)
