"""One configuration of bfp_gemm_mx for ncu.  usage: python tools/prof_gemm_mx.py T N K tile_n"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from qsi_b200 import bfp_ops as ops
T, N, K, tbn = (int(v) for v in sys.argv[1:5])
a = ops.unpack_bfp_args(dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=3, block_size=64,
                             w_sparsity=False, N=2, M=4, first="s", sparsity_mode="structured", device="cuda"))
x, w = torch.randn(T, K, device="cuda"), torch.randn(N, K, device="cuda") * 0.05
xp, wp = ops.pack_bfp_mx(x, 128, identifier="in", **a), ops.pack_bfp_mx(w, tbn, identifier="w", **a)
for _ in range(6):
    y = ops.bfp_linear_mx(xp, wp)
torch.cuda.synchronize()
print("ok", float(y[0, 0]))
