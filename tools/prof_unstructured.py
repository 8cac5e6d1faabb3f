import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qsi_b200 import bfp_ops as ours
a = ours.unpack_bfp_args(dict(num_format="bfp", sparsity_num_format="fp32", rounding_mode="determ", epsilon=1e-8, mant_bits=7, block_size=64, w_sparsity=True,
                              N=2, M=4, first="s", sparsity_mode="unstructured", sparsity_frac=0.5, device="cuda"))
w = torch.randn(4096, 11008, device="cuda") * 0.02
for _ in range(3): y = ours.float_to_bfp_blocked(w, **a, identifier="w")
torch.cuda.synchronize(); print("ok", float(y.abs().sum()))
