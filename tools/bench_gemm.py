"""TOPS of the tcgen05 int8 BFP GEMM at the LLaMA shapes (device-resident packed operands, CUDA events)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qsi_b200 import _lib, bfp_ops as ops
ap = argparse.ArgumentParser(); ap.add_argument("--iters", type=int, default=20); ap.add_argument("--out", default=""); ap.add_argument("--blocks", default="64,32,128")
ap.add_argument("--shapes", default="7b"); a = ap.parse_args()
SH = {"7b": [(4096, 4096, 4096), (4096, 11008, 4096), (4096, 4096, 11008)], "13b": [(4096, 5120, 5120), (4096, 13824, 5120), (4096, 5120, 13824)],
      "65b": [(4096, 8192, 8192), (4096, 22016, 8192), (4096, 8192, 22016)]}
res = []
for (T, N, K) in sum((SH[s] for s in a.shapes.split(",")), []):
    x = torch.randn(T, K, device="cuda"); w = torch.randn(N, K, device="cuda") * 0.02
    for B in [int(b) for b in a.blocks.split(",")]:
        kw = ops.unpack_bfp_args(dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", mant_bits=7, block_size=B,
                                      w_sparsity=True, N=2, M=4, first="s", sparsity_mode="structured", device="cuda"))
        xp, wp = ops.pack_bfp(x, identifier="in", **kw), ops.pack_bfp(w, identifier="w", **kw)
        out = torch.empty(T, N, device="cuda"); L = _lib.lib(); st = torch.cuda.current_stream().cuda_stream
        run = lambda: _lib.check(L.bfp_gemm_i8(xp.mant.data_ptr(), xp.scale_t.data_ptr(), wp.mant.data_ptr(), wp.scale_t.data_ptr(), None, out.data_ptr(), T, N, K, B, st))
        for _ in range(3): run()
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); e0.record()
        for _ in range(a.iters): run()
        e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / a.iters
        # activation pack on the fly
        e0.record()
        for _ in range(a.iters): ops.pack_bfp(x, identifier="in", **kw)
        e1.record(); torch.cuda.synchronize(); ms_pack = e0.elapsed_time(e1) / a.iters
        tops = 2.0 * T * N * K / ms / 1e9
        res.append(dict(T=T, N=N, K=K, B=B, gemm_ms=ms, tops=tops, pack_x_ms=ms_pack))
        if B == 64:
            xb, wb = ops.pack_bfp_bf16(x, identifier="in", **kw), ops.pack_bfp_bf16(w, identifier="w", **kw)
            runb = lambda: _lib.check(L.bfp_gemm_bf16(xb.data_ptr(), wb.data_ptr(), None, out.data_ptr(), T, N, K, st))
            for _ in range(3): runb()
            torch.cuda.synchronize(); e0.record()
            for _ in range(a.iters): runb()
            e1.record(); torch.cuda.synchronize(); msb = e0.elapsed_time(e1) / a.iters
            e0.record()
            for _ in range(a.iters): ops.pack_bfp_bf16(x, identifier="in", **kw)
            e1.record(); torch.cuda.synchronize(); msbp = e0.elapsed_time(e1) / a.iters
            print(f"T={T} N={N} K={K} exact-bf16 kind: gemm {msb:.3f} ms = {2.0*T*N*K/msb/1e9:.0f} TOPS; pack x {msbp*1e3:.1f} us", flush=True)
            for tn in (128, 256):
                _lib.set_option("gemm_bf16_tile_n", tn)
                for _ in range(3): runb()
                torch.cuda.synchronize(); e0.record()
                for _ in range(a.iters): runb()
                e1.record(); torch.cuda.synchronize(); mst = e0.elapsed_time(e1) / a.iters
                print(f"    tile 128x{tn}: {mst:.3f} ms = {2.0*T*N*K/mst/1e9:.0f} TOPS", flush=True)
                res.append(dict(T=T, N=N, K=K, kind="bf16", tile_n=tn, gemm_ms=mst, tops=2.0 * T * N * K / mst / 1e9))
            _lib.set_option("gemm_bf16_tile_n", 0)
            res.append(dict(T=T, N=N, K=K, kind="bf16", gemm_ms=msb, tops=2.0 * T * N * K / msb / 1e9, pack_x_ms=msbp))
        print(f"T={T} N={N} K={K} B={B}: gemm {ms:.3f} ms = {tops:.0f} TOPS ({100*tops/4500:.1f}% of 4500 nominal int8); pack x {ms_pack*1e3:.1f} us", flush=True)
    # yardsticks
    xq = ops.unpack_bfp(xp); wq = ops.unpack_bfp(wp)
    for name, fn in (("torch fp32 F.linear (reference GEMM)", lambda: torch.nn.functional.linear(xq, wq)),
                     ("torch bf16 F.linear", lambda xb=xq.bfloat16(), wb=wq.bfloat16(): torch.nn.functional.linear(xb, wb)),
                     ("torch._int_mm", lambda xa=xp.mant[:, :K].contiguous(), wa=wp.mant[:, :K].t(): torch._int_mm(xa, wa))):
        try:
            for _ in range(2): fn()
            torch.cuda.synchronize(); e0.record()
            for _ in range(5): fn()
            e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 5
            print(f"    {name}: {ms:.3f} ms = {2.0*T*N*K/ms/1e9:.0f} TOPS", flush=True)
            res.append(dict(T=T, N=N, K=K, yardstick=name, ms=ms, tops=2.0 * T * N * K / ms / 1e9))
        except Exception as e:
            print("    ", name, "failed", repr(e)[:100])
if a.out: json.dump(res, open(a.out, "w"), indent=1)
