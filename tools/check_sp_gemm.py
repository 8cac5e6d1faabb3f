"""Bring-up check of the 2:4 structured-sparse GEMM (bfp_gemm_bf16_sp) against the dense exact-bf16 kernel and fp64.
    python tools/check_sp_gemm.py [--diag]
"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from qsi_b200 import bfp_ops as ops

ARGS = dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=7, weight_mant_bits=15,
            block_size=64, in_sparsity=False, w_sparsity=True, grad_sparsity=False, N=2, M=4, first="s", sparsity_mode="structured",
            sparsity_frac=0.5, device="cuda")


def diag():
    # x = identity  =>  y[t, n] = W[n, t]: any metadata / descriptor mistake shows up as a readable permutation
    T = K = 256; N = 128
    g = torch.Generator().manual_seed(0)
    W = torch.zeros(N, K)
    for n in range(N):
        for grp in range(K // 4):
            i, j = sorted(torch.randperm(4, generator=g)[:2].tolist())
            W[n, 4 * grp + i] = float(1 + (n * 7 + grp * 3) % 50)
            W[n, 4 * grp + j] = -float(1 + (n * 5 + grp) % 50)
    wb = W.to(torch.bfloat16).cuda()
    xb = torch.eye(T, K).to(torch.bfloat16).cuda()
    ws = ops.compress_2to4_bf16(wb)
    y = ops.bfp_linear_bf16_sp(xb, ws).cpu()           # [T, N]
    ref = W.t()
    ok = torch.equal(y, ref)
    print("identity test equal:", ok)
    if not ok:
        bad = (y != ref)
        print("mismatching entries:", int(bad.sum()), "of", bad.numel())
        tt, nn = torch.nonzero(bad, as_tuple=True)
        for t, n in list(zip(tt.tolist(), nn.tolist()))[:24]:
            # where does the value we got live in W?
            got = float(y[t, n]); src = torch.nonzero(W == got)
            print(f"  y[t={t}, n={n}] = {got}, expected {float(ref[t, n])}; W has that value at {src[:4].tolist()}")
    return ok


def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--diag", action="store_true"); a = ap.parse_args()
    from qsi_b200 import _lib
    ok, worst = True, 0.0
    for cg in (1, 2):
      _lib.set_option("gemm_sp_cta_group", cg)
      print("cta_group", cg)
      ok = diag() and ok
      for (T, N, K) in [(256, 128, 128), (512, 256, 448), (300, 200, 264), (77, 300, 136), (4096, 4096, 4096), (1000, 11008, 4096)]:
        torch.manual_seed(T + N + K)
        x = torch.randn(T, K, device="cuda"); w = torch.randn(N, K, device="cuda") * 0.02
        bias = torch.randn(N, device="cuda")
        xb = ops.pack_bfp_bf16(x, identifier="in", **ARGS)
        wb = ops.pack_bfp_bf16(w, identifier="w", **ARGS)
        ws = ops.compress_2to4_bf16(wb)
        y_sp = ops.bfp_linear_bf16_sp(xb, ws, bias)
        y_d = ops.bfp_linear_bf16(xb, wb, bias)
        ref = (xb.double() @ wb.double().t() + bias.double())
        rel_sp = float((y_sp.double() - ref).norm() / ref.norm()); rel_d = float((y_d.double() - ref).norm() / ref.norm())
        mx = float((y_sp.double() - ref).abs().max() / ref.abs().max())
        worst = max(worst, rel_sp)
        print(f"  T={T} N={N} K={K}: sparse rel {rel_sp:.3e} (max {mx:.3e}), dense rel {rel_d:.3e}, finite {bool(torch.isfinite(y_sp).all())}")
    print("RESULT", "PASS" if ok and worst <= 1e-5 else "FAIL", worst)


if __name__ == "__main__":
    main()
