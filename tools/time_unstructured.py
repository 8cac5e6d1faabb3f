import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qsi_b200 import bfp_ops as ours, _lib
L = _lib.lib(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1e3
for shape in [(4096, 4096), (4096, 11008)]:
    for scale, name in ((0.02, "randn*0.02"), (1.0, "randn")):
        w = torch.randn(*shape, device="cuda") * scale; out = torch.empty_like(w); n = w.numel()
        ws = torch.empty(L.bfp_unstructured_workspace_bytes() // 8 + 1, dtype=torch.int64, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        us = t(lambda: _lib.check(L.bfp_unstructured_sparsify(w.data_ptr(), out.data_ptr(), n, 0, n // 2, ws.data_ptr(), st)))
        a = ours.unpack_bfp_args(dict(num_format="bfp", sparsity_num_format="fp32", rounding_mode="determ", epsilon=1e-8, mant_bits=7, block_size=64, w_sparsity=True,
                                      N=2, M=4, first="s", sparsity_mode="unstructured", sparsity_frac=0.5, device="cuda"))
        us2 = t(lambda: ours.float_to_bfp_blocked(w, **a, identifier="w"))
        print(f"{shape} {name}: C ABI {us:.1f} us ({n*8/us/1e3:.0f} GB/s), python API {us2:.1f} us", flush=True)
