"""Per-tile fixed cost of the sparse kernels: N = 256 * 74 * W waves (W full waves on 74 CTA pairs), T = one tile;
time(K) = W * (f + slabs * c).  Fit f (fixed clk per tile) and c (clk per 128-k slab) from K = 1024 .. 16384."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qsi_b200 import _lib, bfp_ops as ops
L = _lib.lib(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
WAVES = 4
N = 256 * 74                      # one wave of W row-tiles; the WAVES token tiles re-read W from L2
for tile, T, tma in ((256, 256 * WAVES, 1), (256, 256 * WAVES, 0), (480, 480 * WAVES, 1), (480, 480 * WAVES, 0)):
    _lib.set_option("gemm_sp_tile", tile); _lib.set_option("gemm_out_tma", tma); print("tma stores", tma)
    pts = []
    for K in (512, 1024, 2048, 4096):
        xb = torch.randn(T, K, device="cuda").to(torch.bfloat16)
        w = torch.randn(N, K, device="cuda"); w.view(N, -1, 4)[:, :, 2:] = 0
        wb = w.to(torch.bfloat16); del w
        ws = ops.compress_2to4_bf16(wb, check=False); del wb
        out = torch.empty(T, N, device="cuda"); st = torch.cuda.current_stream().cuda_stream
        f = lambda: _lib.check(L.bfp_gemm_bf16_sp(xb.data_ptr(), ws.comp.data_ptr(), ws.meta.data_ptr(), None, out.data_ptr(), T, N, K, st))
        for _ in range(3): f()
        torch.cuda.synchronize(); e0.record()
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize(); us = e0.elapsed_time(e1) / 10 * 1e3
        pts.append((K // 128, us / WAVES))
        print(f"tile {tile} K={K}: {us:.1f} us total, {us/WAVES:.2f} us per tile, {2.0*T*N*K/us/1e6:.0f} TOPS", flush=True)
        del ws, out, xb
    # least squares fit us = f + slabs * c
    n = len(pts); sx = sum(p[0] for p in pts); sy = sum(p[1] for p in pts); sxx = sum(p[0] ** 2 for p in pts); sxy = sum(p[0] * p[1] for p in pts)
    c = (n * sxy - sx * sy) / (n * sxx - sx * sx); f0 = (sy - c * sx) / n
    print(f"tile {tile}: per-slab {c*1e3:.0f} ns, fixed per tile {f0:.2f} us (includes launch / n_waves)", flush=True)
_lib.set_option("gemm_sp_tile", 0); _lib.set_option("gemm_out_tma", 1)
