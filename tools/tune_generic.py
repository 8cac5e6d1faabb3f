"""Throughput of the ragged-shape (generic) quantiser path: K not a multiple of the block size, e.g. the ViT patch-embedding input."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qsi_b200 import _lib, bfp_ops as ops
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1e3
for shape, B, ident, sp in (((256, 3, 224, 224), 64, "in", False), ((768, 3, 16, 16), 64, "w", True), ((64 * 197, 768), 64, "in", False), ((4096, 4100), 64, "w", True), ((4096, 4096), 48, "w", True)):
    a = ops.unpack_bfp_args(dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=5, block_size=B,
                                 w_sparsity=sp, N=2, M=4, first="s", sparsity_mode="structured", device="cuda"))
    x = torch.randn(*shape, device="cuda")
    us = t(lambda: ops.float_to_bfp_blocked(x, **a, identifier=ident))
    usp = t(lambda: ops.pack_bfp_bf16(x, identifier=ident, **a))
    print(f"{shape} B={B} {ident}: fake-quant {us:.1f} us = {x.numel()*8/us/1e3:.0f} GB/s | pack bf16 {usp:.1f} us = {x.numel()*6/usp/1e3:.0f} GB/s", flush=True)
