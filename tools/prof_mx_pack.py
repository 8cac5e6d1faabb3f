"""A handful of launches of the operand-form packs (OCP MX and BFP block-scaled) and the MX fake-quantiser, for ncu.
    python tools/prof_mx_pack.py [rows K]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qsi_b200  # noqa: F401
from qsi_b200 import bfp_ops, mx_layers as mx
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 11008
xs = [torch.randn(rows, K, device="cuda") for _ in range(4)]
sp = mx.finalize_mx_specs(mx.apply_mx_specs(dict(block_size=32, bfloat=16, scale_bits=8, w_elem_format="fp8_e4m3", a_elem_format="fp8_e4m3")))
args = bfp_ops.unpack_bfp_args(dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=3, block_size=64, device="cuda"))
for i in range(6):
    mx._pack_block_scaled(xs[i % 4], mx.ELEM_FORMATS["fp8_e4m3"], 128, sp, 16)
    mx._mx_quantize_last(xs[i % 4], mx.ELEM_FORMATS["fp8_e4m3"], 32, 8, 16, False)
    bfp_ops.pack_activation_mx(xs[i % 4], args)
torch.cuda.synchronize()
print("ok")
