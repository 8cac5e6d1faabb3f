"""Timing experiment: is the sparse GEMM bound by L2->SM traffic?  debug bit 0/1 make every tile load the same X / W tile."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qsi_b200 import _lib, bfp_ops as ops
kw = ops.unpack_bfp_args(dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", mant_bits=7, block_size=64,
                              w_sparsity=True, N=2, M=4, first="s", sparsity_mode="structured", device="cuda"))
L = _lib.lib(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for (T, N, K) in [(4096, 8192, 8192), (4096, 4096, 4096)]:
    x = torch.randn(T, K, device="cuda"); w = torch.randn(N, K, device="cuda") * 0.02
    xb, wb = ops.pack_bfp_bf16(x, identifier="in", **kw), ops.pack_bfp_bf16(w, identifier="w", **kw)
    ws = ops.compress_2to4_bf16(wb); out = torch.empty(T, N, device="cuda"); st = torch.cuda.current_stream().cuda_stream
    for cg in (2, 1):
        for dbg in (0, 3, 4):
            _lib.set_option("gemm_sp_cta_group", cg); _lib.set_option("gemm_sp_debug", dbg)
            f = lambda: _lib.check(L.bfp_gemm_bf16_sp(xb.data_ptr(), ws.comp.data_ptr(), ws.meta.data_ptr(), None, out.data_ptr(), T, N, K, st))
            for _ in range(3): f()
            torch.cuda.synchronize(); e0.record()
            for _ in range(20): f()
            e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 20
            print(f"T={T} N={N} K={K} cg={cg} debug={dbg}: {ms:.3f} ms = {2.0*T*N*K/ms/1e9:.0f} TOPS", flush=True)
_lib.set_option("gemm_sp_debug", 0)
