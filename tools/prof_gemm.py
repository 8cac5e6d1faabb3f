"""A few launches of one GEMM kind for ncu.  python tools/prof_gemm.py [i8|bf16|sp|sp1] [B] [T N K]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qsi_b200 import _lib, bfp_ops as ops
kind = sys.argv[1] if len(sys.argv) > 1 else "i8"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
T, N, K = (int(v) for v in sys.argv[3:6]) if len(sys.argv) > 5 else (4096, 4096, 4096)
x = torch.randn(T, K, device="cuda"); w = torch.randn(N, K, device="cuda") * 0.02
kw = ops.unpack_bfp_args(dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", mant_bits=7, block_size=B,
                              w_sparsity=True, N=2, M=4, first="s", sparsity_mode="structured", device="cuda"))
if kind == "i8":
    xp, wp = ops.pack_bfp(x, identifier="in", **kw), ops.pack_bfp(w, identifier="w", **kw)
    for _ in range(4): y = ops.bfp_linear_packed(xp, wp)
elif kind in ("sp", "sp1"):
    _lib.set_option("gemm_sp_cta_group", 1 if kind == "sp1" else 0)
    xb, wb = ops.pack_bfp_bf16(x, identifier="in", **kw), ops.pack_bfp_bf16(w, identifier="w", **kw)
    ws = ops.compress_2to4_bf16(wb)
    for _ in range(4): y = ops.bfp_linear_bf16_sp(xb, ws)
else:
    xb, wb = ops.pack_bfp_bf16(x, identifier="in", **kw), ops.pack_bfp_bf16(w, identifier="w", **kw)
    for _ in range(4): y = ops.bfp_linear_bf16(xb, wb)
torch.cuda.synchronize(); print("ok", float(y.abs().sum()))
