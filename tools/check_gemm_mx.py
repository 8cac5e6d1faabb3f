"""bfp_gemm_mx (tcgen05.mma.kind::mxf8f6f4.block_scale) against the fp64 product of the fake-quantised operands, with the
hypotheses that would explain a wrong layout evaluated beside it.  usage: python tools/check_gemm_mx.py [bench]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from qsi_b200 import bfp_ops as ops, _lib
L = _lib.lib()
torch.manual_seed(0)

def args(m, B):
    return ops.unpack_bfp_args(dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=m, block_size=B,
                                    w_sparsity=False, N=2, M=4, first="s", sparsity_mode="structured", device="cuda"))

def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))

def one(T, N, K, m, B, tbn, variant, fold=False, scale_spread=True):
    _lib.check(L.bfp_set_option(b"gemm_mx_variant", variant))
    a = args(m, B)
    x = torch.randn(T, K, device="cuda")
    w = torch.randn(N, K, device="cuda") * 0.05
    if scale_spread:   # block exponents that differ along K and across rows, so a mis-indexed scale shows
        x = x * torch.exp2(torch.randint(-6, 7, (T, K // B), device="cuda").repeat_interleave(B, 1).float())
        w = w * torch.exp2(torch.randint(-3 if fold else -6, 4 if fold else 7, (N, K // B), device="cuda").repeat_interleave(B, 1).float())
    xq = ops.float_to_bfp_blocked(x, **a, identifier="in")
    wq = ops.float_to_bfp_blocked(w, **a, identifier="w")
    ref = xq.double() @ wq.double().t()
    xp = ops.pack_bfp_mx(x, 128, identifier="in", **a)
    wp = ops.pack_bfp_mx(w, tbn, fold=fold, identifier="w", **a)
    assert wp is not None, "the weight did not fit the requested form"
    y = ops.bfp_linear_mx(xp, wp)
    torch.cuda.synchronize()
    r = rel(y, ref)
    msg = f"T={T} N={N} K={K} m={m} B={B} tile_n={tbn} variant={variant} fold={int(fold)}: rel err {r:.3e}"
    if r > 1e-5:
        # hypotheses: scales ignored; every MMA uses the slab's first scale; A / B scales swapped between 32-groups
        pk_x, pk_w = ops.pack_bfp(x, identifier="in", **a), ops.pack_bfp(w, identifier="w", **a)
        qx, qw = pk_x.mant[:, :K].double(), pk_w.mant[:, :K].double()
        sx = pk_x.scale_t[: K // B, :T].t().double().repeat_interleave(B, 1)      # [T, K]
        sw = pk_w.scale_t[: K // B, :N].t().double().repeat_interleave(B, 1)
        hyp = {"no scales": qx @ qw.t(), "A scales only": (qx * sx) @ qw.t(), "B scales only": qx @ (qw * sw).t()}
        g0x = sx.view(T, K // 128, 128)[:, :, :1].expand(T, K // 128, 128).reshape(T, K)
        g0w = sw.view(N, K // 128, 128)[:, :, :1].expand(N, K // 128, 128).reshape(N, K)
        hyp["first scale of the slab for all four MMAs"] = (qx * g0x) @ (qw * g0w).t()
        best = min(hyp.items(), key=lambda kv: rel(y, kv[1]))
        msg += f"  | closest hypothesis: {best[0]} (rel {rel(y, best[1]):.2e}); finite {bool(torch.isfinite(y).all())} |y| {float(y.abs().mean()):.3e} |ref| {float(ref.abs().mean()):.3e}"
        # per-row / per-column error pattern
        err = (y.double() - ref).abs() / ref.abs().clamp_min(1e-30)
        msg += f"  rows ok {int((err.max(1).values < 1e-4).sum())}/{T} cols ok {int((err.max(0).values < 1e-4).sum())}/{N}"
    print(msg, flush=True)
    return r

if "bench" in sys.argv:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for (T, N, K) in ((4096, 4096, 4096), (4096, 11008, 4096), (4096, 4096, 11008)):
        a = args(3, 64)
        x, w = torch.randn(T, K, device="cuda"), torch.randn(N, K, device="cuda") * 0.05
        xp = ops.pack_bfp_mx(x, 128, identifier="in", **a)
        for tbn, variant, fold in ((240, 0, True), (240, 12, True), (240, 12 + 16, True), (240, 12 + 32, True), (240, 6, True)):
            _lib.check(L.bfp_set_option(b"gemm_mx_variant", variant))
            wp = ops.pack_bfp_mx(w, tbn, fold=fold, identifier="w", **a)
            for _ in range(3): ops.bfp_linear_mx(xp, wp)
            torch.cuda.synchronize(); e0.record()
            for _ in range(20): ops.bfp_linear_mx(xp, wp)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print(f"T={T} N={N} K={K} tile_n={tbn} variant {variant} fold {int(fold)}: {ms * 1e3:.1f} us = {2 * T * N * K / ms / 1e9:.0f} TFLOP/s", flush=True)
        _lib.check(L.bfp_set_option(b"gemm_mx_variant", 0))
else:
    ok = True
    cases = ((128, 128, 128, 3, 32, 128), (128, 256, 128, 3, 32, 256), (256, 256, 128, 3, 32, 256), (128, 128, 512, 3, 32, 128), (256, 512, 1024, 3, 64, 256),
             (200, 260, 640, 4, 32, 128), (300, 520, 640, 4, 32, 256), (300, 520, 640, 4, 32, 240), (256, 240, 256, 3, 32, 240), (4096, 4096, 4096, 3, 64, 256),
             (4096, 4096, 4096, 3, 64, 240))
    for variant, fold in ((0, False), (0, True), (1, False)):        # variant 0: CTA pairs where they apply, 1: single CTAs
        for (T, N, K, m, B, tbn) in cases:
            ok = (one(T, N, K, m, B, tbn, variant, fold) <= 1e-5) and ok
    # folded weights take any block size (the activation side needs 32-multiples): HBFP4 weights in blocks of 16 against blocks of 32
    print("ALL OK" if ok else "MISMATCH")
