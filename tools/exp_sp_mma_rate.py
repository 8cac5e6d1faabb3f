"""MMA-only rate of the sparse / dense kernels on a FEW SMs (no power limit): T=256, N=2048, K=131072, sparse debug bit 4 = no loads."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qsi_b200 import _lib, bfp_ops as ops
L = _lib.lib(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
T, N, K = 256, 2048, 131072
xb = torch.randn(T, K, device="cuda").to(torch.bfloat16)
w = torch.randn(N, K, device="cuda"); w.view(N, -1, 4)[:, :, 2:] = 0
wb = w.to(torch.bfloat16); del w
ws = ops.compress_2to4_bf16(wb); out = torch.empty(T, N, device="cuda"); st = torch.cuda.current_stream().cuda_stream
def t(fn):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / 5
for cg in (2, 1):
    for dbg in (4, 0):
        _lib.set_option("gemm_sp_cta_group", cg); _lib.set_option("gemm_sp_debug", dbg)
        ms = t(lambda: _lib.check(L.bfp_gemm_bf16_sp(xb.data_ptr(), ws.comp.data_ptr(), ws.meta.data_ptr(), None, out.data_ptr(), T, N, K, st)))
        ctas = (N // 256) * 2 if cg == 2 else N // 128
        mmas = K // 32                                         # per CTA (pair) tile
        print(f"sparse cg={cg} debug={dbg}: {ms:.3f} ms, {ctas} CTAs, {ms*1e-3*1.965e9/mmas:.1f} clk per MMA @1.965 GHz, {2.0*T*N*K/ms/1e9/ctas:.2f} TOPS per SM", flush=True)
_lib.set_option("gemm_sp_debug", 0)
for cg in (2, 1):
    _lib.set_option("gemm_bf16_cta_group", cg)
    ms = t(lambda: _lib.check(L.bfp_gemm_bf16(xb.data_ptr(), wb.data_ptr(), None, out.data_ptr(), T, N, K, st)))
    ctas = (N // 256) * 2 if cg == 2 else (T // 128) * (N // 256)
    print(f"dense cg={cg}: {ms:.3f} ms, {ctas} CTAs, {ms*1e-3*1.965e9/(K//16):.1f} clk per MMA @1.965 GHz, {2.0*T*N*K/ms/1e9/ctas:.2f} TOPS per SM", flush=True)
