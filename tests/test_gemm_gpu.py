"""GPU parity suite for the packed BFP operands and the tcgen05 int8 BFP GEMM (C ABI: bfp_quantize_pack, bfp_unpack,
bfp_gemm_i8).  Tolerance (north_star): relative error <= 1e-5 against the reference's dequantise-then-matmul; the
integer part of the contraction is exact, so with unit scales the result must be bit-exact."""
import itertools

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available()
    from qsi_b200 import bfp_ops, _lib
    _lib.lib()
    return bfp_ops


def _make_packed(ops, q, scale, B, m=7):
    """q: int8 [rows, K] tensor, scale: fp32 [rows, nkb] -> PackedBFP in the documented layout."""
    from qsi_b200 import _lib
    rows, K = q.shape
    Kp, rows_pad, nkb_pad = _lib.packed_layout(rows, K, B)
    mant = torch.zeros(rows, Kp, dtype=torch.int8, device="cuda")
    mant[:, :K] = q
    st = torch.zeros(nkb_pad, rows_pad, dtype=torch.float32, device="cuda")
    st[: scale.shape[1], :rows] = scale.t()
    return ops.PackedBFP(mant, st, (rows, K), B, m)


def _ref(qa, sa, qb, sb, B, bias=None):
    a = qa.double() * sa.double().repeat_interleave(B, dim=1)[:, : qa.shape[1]]
    b = qb.double() * sb.double().repeat_interleave(B, dim=1)[:, : qb.shape[1]]
    y = a @ b.t()
    return y if bias is None else y + bias.double()


@pytest.mark.parametrize("B", [64, 32, 128])
@pytest.mark.parametrize("shape", [(128, 256, 128), (256, 512, 512), (200, 300, 384), (1, 8, 64 * 5), (4096, 768, 1024)])
def test_gemm_unit_scales_is_exact_integer_matmul(ops, B, shape):
    T, N, K = shape
    g = torch.Generator(device="cuda").manual_seed(T + N + K + B)
    qa = torch.randint(-127, 128, (T, K), generator=g, device="cuda", dtype=torch.int32).to(torch.int8)
    qb = torch.randint(-127, 128, (N, K), generator=g, device="cuda", dtype=torch.int32).to(torch.int8)
    nkb = -(-K // B)
    sa = torch.ones(T, nkb, device="cuda")
    sb = torch.ones(N, nkb, device="cuda")
    y = ops.bfp_linear_packed(_make_packed(ops, qa, sa, B), _make_packed(ops, qb, sb, B))
    ref = _ref(qa, sa, qb, sb, B)
    if K * 127 * 127 < 2 ** 24:
        assert torch.equal(y.double(), ref)
    else:
        assert (y.double() - ref).abs().max() <= ref.abs().max() * 2 ** -22


@pytest.mark.parametrize("B", [64, 32, 128])
@pytest.mark.parametrize("shape", [(256, 512, 512), (200, 300, 384), (130, 260, 200), (512, 1024, 2048)])
def test_gemm_power_of_two_scales_and_bias(ops, B, shape):
    T, N, K = shape
    g = torch.Generator(device="cuda").manual_seed(7 * T + N + K + B)
    qa = torch.randint(-127, 128, (T, K), generator=g, device="cuda", dtype=torch.int32).to(torch.int8)
    qb = torch.randint(-127, 128, (N, K), generator=g, device="cuda", dtype=torch.int32).to(torch.int8)
    nkb = -(-K // B)
    sa = torch.exp2(torch.randint(-12, 4, (T, nkb), generator=g, device="cuda").float())
    sb = torch.exp2(torch.randint(-14, -6, (N, nkb), generator=g, device="cuda").float())
    bias = torch.randn(N, generator=g, device="cuda")
    y = ops.bfp_linear_packed(_make_packed(ops, qa, sa, B), _make_packed(ops, qb, sb, B), bias)
    ref = _ref(qa, sa, qb, sb, B, bias)
    rel = ((y.double() - ref).norm() / ref.norm()).item()
    assert rel <= 1e-6, rel
    assert ((y.double() - ref).abs() <= 1e-5 * ref.abs() + 1e-6 * ref.abs().mean()).all()


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16, torch.float16])
def test_pack_unpack_round_trip_equals_fake_quant(ops, dt):
    """Contract of the packed layout: unpack(pack(x)) == float_to_bfp_blocked(x) (torch.equal: -0.0 unpacks as +0.0)."""
    from qsi_b200 import _lib
    for shape, (m, B), first, sparse in itertools.product([(256, 1024), (37, 96), (3, 5, 200), (130, 70)], [(7, 64), (3, 32), (5, 128), (7, 16)],
                                                           ["s", "q"], [True, False]):
        g = torch.Generator().manual_seed(len(shape) * 100 + m + B)
        x = (torch.randn(*shape, generator=g) * 0.05).to(dt).cuda()
        x.view(-1)[::37] = 0
        args = ops.unpack_bfp_args(dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=m,
                                        block_size=B, w_sparsity=sparse, N=2, M=4, first=first, sparsity_mode="structured", device="cuda"))
        fq = ops.float_to_bfp_blocked(x, **args, identifier="w").float()
        for generic in (0, 1):
            _lib.set_option("force_generic", generic)
            try:
                p = ops.pack_bfp(x, identifier="w", **args)
            finally:
                _lib.set_option("force_generic", 0)
            assert p.mant.abs().max().item() <= 2 ** m - 1
            assert torch.equal(ops.unpack_bfp(p), fq), (shape, m, B, first, sparse, generic)
    # stochastic packing draws the same Philox stream as the fake-quant kernel
    x = torch.randn(64, 512, device="cuda")
    a = ops.unpack_bfp_args(dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="stoc", mant_bits=7, block_size=64))
    p = ops.pack_bfp(x, identifier="in", philox=(5, 3), **a)
    fq = ops._fused(x, _lib.ORDER_QUANT_ONLY, block_size=64, mant_bits=7, epsilon=1e-8, rounding_mode="stoc", philox=(5, 3))
    assert torch.equal(ops.unpack_bfp(p), fq)


def test_unrepresentable_blocks_become_nan_rows(ops):
    x = torch.randn(4, 128, device="cuda")
    x[1, 5] = float("inf")
    x[2, 70] = float("nan")
    a = ops.unpack_bfp_args(dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", mant_bits=7, block_size=64))
    p = ops.pack_bfp(x, identifier="in", **a)
    w = ops.pack_bfp(torch.randn(8, 128, device="cuda"), identifier="w", **a)
    y = ops.bfp_linear_packed(p, w)
    ref = ops.float_to_bfp_blocked(x, **a, identifier="in") @ ops.float_to_bfp_blocked(torch.zeros(8, 128, device="cuda"), **a, identifier="w").t()
    assert torch.isnan(y[1]).all() and torch.isnan(y[2]).all() and torch.isfinite(y[0]).all() and torch.isfinite(y[3]).all()
    assert torch.isnan(ref[1]).all() and torch.isnan(ref[2]).all()          # the reference arithmetic NaNs the same rows


@pytest.mark.parametrize("shape", [(128, 256, 64), (256, 512, 512), (200, 300, 384), (1, 8, 72), (4096, 768, 1024), (130, 260, 200)])
def test_gemm_bf16_exact_products(ops, shape):
    """kind::f16 path: small-integer bf16 operands -> the fp32 result is the exact integer matmul."""
    T, N, K = shape
    g = torch.Generator(device="cuda").manual_seed(T * 3 + N + K)
    a = torch.randint(-15, 16, (T, K), generator=g, device="cuda").float()
    b = torch.randint(-15, 16, (N, K), generator=g, device="cuda").float()
    Kp = -(-K // 8) * 8
    ab = torch.zeros(T, Kp, dtype=torch.bfloat16, device="cuda"); ab[:, :K] = a
    bb = torch.zeros(N, Kp, dtype=torch.bfloat16, device="cuda"); bb[:, :K] = b
    bias = torch.randint(-5, 6, (N,), generator=g, device="cuda").float()
    from qsi_b200 import _lib
    ref = a.double() @ b.double().t() + bias.double()
    for tile_n, cg, tma in ((0, 0, 1), (128, 1, 1), (256, 1, 1), (0, 2, 1), (0, 2, 0), (256, 1, 0)):   # auto; tile widths; CTA pairs; plain-store epilogue
        _lib.set_option("gemm_bf16_tile_n", tile_n)
        _lib.set_option("gemm_bf16_cta_group", cg)
        _lib.set_option("gemm_out_tma", tma)
        try:
            y = ops.bfp_linear_bf16(ab, bb, bias)
        finally:
            _lib.set_option("gemm_bf16_tile_n", 0)
            _lib.set_option("gemm_bf16_cta_group", 0)
            _lib.set_option("gemm_out_tma", 1)
        assert torch.equal(y.double(), ref), (tile_n, cg, tma)
        if N % 8 == 0 or tma == 0:                     # half outputs = the same accumulators rounded once in the epilogue
            _lib.set_option("gemm_bf16_tile_n", tile_n); _lib.set_option("gemm_bf16_cta_group", cg); _lib.set_option("gemm_out_tma", tma)
            try:
                for dt in (torch.float16, torch.bfloat16):
                    yh = ops.bfp_linear_bf16(ab, bb, bias, out_dtype=dt)
                    assert yh.dtype == dt and torch.equal(yh, y.to(dt)), (tile_n, cg, tma, dt)
            finally:
                _lib.set_option("gemm_bf16_tile_n", 0); _lib.set_option("gemm_bf16_cta_group", 0); _lib.set_option("gemm_out_tma", 1)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16, torch.float16])
def test_pack_bf16_equals_fake_quant(ops, dt):
    """bf16 operand = float_to_bfp_blocked(x) exactly (m <= 8), for every block size including 16."""
    from qsi_b200 import _lib
    for shape, (m, B), first, sparse in itertools.product([(256, 1024), (37, 96), (3, 5, 200), (130, 72)], [(7, 64), (3, 16), (5, 32), (8, 128)],
                                                           ["s", "q"], [True, False]):
        g = torch.Generator().manual_seed(len(shape) * 10 + m + B)
        x = (torch.randn(*shape, generator=g) * 0.05).to(dt).cuda()
        args = ops.unpack_bfp_args(dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=m,
                                        block_size=B, w_sparsity=sparse, N=2, M=4, first=first, sparsity_mode="structured", device="cuda"))
        fq = ops.float_to_bfp_blocked(x, **args, identifier="w").float().reshape(-1, shape[-1])
        for generic in (0, 1):
            _lib.set_option("force_generic", generic)
            try:
                pb = ops.pack_bfp_bf16(x, identifier="w", **args)
            finally:
                _lib.set_option("force_generic", 0)
            assert torch.equal(pb[:, : shape[-1]].float(), fq), (shape, m, B, first, sparse, generic)
            assert (pb[:, shape[-1]:] == 0).all()


@pytest.mark.parametrize("kind", ["sp", "bf16", "i8"])
@pytest.mark.parametrize("first", ["s", "q"])
@pytest.mark.parametrize("mB", [(7, 64), (5, 32), (3, 128), (3, 16)])
def test_bfplinear_tensor_core_path_matches_oracle(ops, oracle, first, mB, kind, monkeypatch):
    m, B = mB
    monkeypatch.setenv("BFP_GEMM_KIND", kind)
    kw = dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=m, block_size=B,
              w_sparsity=True, N=2, M=4, first=first, sparsity_mode="structured", device="cuda")
    torch.manual_seed(1)
    lin = ops.BFPLinear(512, 384, bias=True, **dict(kw)).cuda()
    x = torch.randn(3, 70, 512, device="cuda")
    x[0, 0, 3] = 25.0
    with torch.no_grad():
        y_tc = lin(x)
        monkeypatch.setenv("BFP_LINEAR_PATH", "fakequant")
        y_fq = lin(x)
        monkeypatch.setenv("BFP_LINEAR_PATH", "tc")
    xq, _ = oracle.bfp_quantize(x.cpu().numpy(), B, m)
    wq, _ = oracle.float_to_bfp_blocked(lin.weight.detach().cpu().numpy(), m, B, "sq" if first == "s" else "qs", tie_rule="cuda")
    ref = oracle.linear(xq, wq, lin.bias.detach().cpu().numpy())
    for y in (y_tc, y_fq):
        rel = np.linalg.norm(y.cpu().numpy() - ref) / np.linalg.norm(ref)
        assert rel <= 1e-5, rel
    if B == 16 and kind == "i8":
        return                                   # one int8 MMA is 32 deep: B = 16 is served by the bf16 kind
    # weight pack is cached until the weight changes
    k0 = lin._packed_w[0]
    assert k0[0] == kind
    with torch.no_grad():
        lin(x)
        assert lin._packed_w[0] == k0
        lin.weight.mul_(2.0)
        y2 = lin(x)
    assert lin._packed_w[0] != k0
    assert torch.allclose(y2 - lin.bias, 2 * (y_tc - lin.bias), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("NK", [(4096, 4096), (11008, 4096), (4096, 11008)])
def test_gemm_llama7b_shapes_full_size(ops, NK):
    """BASELINE LLaMA-7B shapes, T = 4096 tokens: parity against fp64 matmul of the dequantised operands."""
    N, K = NK
    T = 4096
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(T, K, device="cuda", generator=g)
    x[::97, ::53] *= 20.0
    w = torch.randn(N, K, device="cuda", generator=g) * 0.02
    a = ops.unpack_bfp_args(dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", mant_bits=7, block_size=64,
                                 w_sparsity=True, N=2, M=4, first="s", sparsity_mode="structured", device="cuda"))
    xp, wp = ops.pack_bfp(x, identifier="in", **a), ops.pack_bfp(w, identifier="w", **a)
    y = ops.bfp_linear_packed(xp, wp)
    rows = torch.arange(0, T, 61, device="cuda")
    ref = ops.unpack_bfp(xp)[rows].double() @ ops.unpack_bfp(wp).double().t()
    rel = ((y[rows].double() - ref).norm() / ref.norm()).item()
    assert rel <= 1e-5, rel
    # linearity in the activations' block scales: doubling x doubles y exactly (power-of-two scaling commutes with BFP)
    y2 = ops.bfp_linear_packed(ops.pack_bfp(2 * x, identifier="in", **a), wp)
    assert torch.equal(y2, 2 * y)
    # the exact-bf16 kind computes the same function (only the fp32 summation order differs)
    yb = ops.bfp_linear_bf16(ops.pack_bfp_bf16(x, identifier="in", **a), ops.pack_bfp_bf16(w, identifier="w", **a))
    assert ((yb[rows].double() - ref).norm() / ref.norm()).item() <= 1e-5
    assert ((yb - y).norm() / y.norm()).item() <= 2e-6


# ---------------------------------------------------------------------------------------------------------------------
# 2:4 structured-sparse kind (bfp_compress_2to4_bf16 + bfp_gemm_bf16_sp, tcgen05.mma.sp)
# ---------------------------------------------------------------------------------------------------------------------
def _random_2to4(rows, K, g, scale=1.0):
    """bf16 [rows, K] with at most two non-zeros per aligned group of four, including empty / single-entry groups and -0.0."""
    w = torch.randn(rows, K, generator=g) * scale
    grp = w.view(rows, -1, 4)
    keep = torch.zeros_like(grp, dtype=torch.bool)
    order = torch.rand(grp.shape, generator=g).argsort(dim=-1)
    nkeep = torch.randint(0, 3, grp.shape[:2], generator=g)              # 0, 1 or 2 survivors
    for j in range(2):
        keep.scatter_(2, order[..., j:j + 1], (nkeep > j).unsqueeze(-1))
    out = torch.where(keep, grp, torch.zeros_like(grp))
    out[0, 0, 0] = -0.0
    return out.view(rows, K).to(torch.bfloat16)


def _decompress(ws, rows, K):
    """numpy restatement of the documented layout (csrc/bfp_gemm_sp.cu header) -> dense fp32 [rows, K]."""
    comp = ws.comp.float().cpu().numpy()
    meta = ws.meta.cpu().numpy()
    atoms = comp.shape[1] // 64
    dense = np.zeros((rows, atoms * 128), dtype=np.float32)
    for r in range(rows):
        m = r & 127
        m0, m1, m2 = m & 7, (m >> 3) & 1, m >> 4
        for h in range(atoms * 8):
            atom, k1, k2 = h >> 3, h & 1, (h >> 1) & 3
            lane = m0 + 8 * k1 + 16 * m2
            off = ((r >> 7) * atoms + atom) * 2048 + lane * 16 + k2 * 4 + m1 * 2
            word = int(meta[off]) | (int(meta[off + 1]) << 8)
            for gi in range(4):
                nib = (word >> (4 * gi)) & 15
                i0, i1 = nib & 3, nib >> 2
                assert i0 < i1, (r, h, gi, nib)
                dense[r, h * 16 + gi * 4 + i0] += comp[r, h * 8 + gi * 2]
                dense[r, h * 16 + gi * 4 + i1] += comp[r, h * 8 + gi * 2 + 1]
    return dense[:, :K]


@pytest.mark.parametrize("shape", [(128, 128), (200, 264), (5, 8), (300, 1000)])
def test_compress_2to4_layout_round_trip(ops, shape):
    rows, K = shape
    g = torch.Generator().manual_seed(rows * 7 + K)
    wb = _random_2to4(rows, K, g).cuda()
    ws = ops.compress_2to4_bf16(wb)
    from qsi_b200 import _lib
    Kc, mb = _lib.sp_layout(rows, K)
    assert tuple(ws.comp.shape) == (rows, Kc) and ws.meta.numel() == mb and Kc == -(-K // 128) * 64
    assert np.array_equal(_decompress(ws, rows, K), wb.float().cpu().numpy())


def test_compress_rejects_dense_groups(ops):
    wb = torch.zeros(16, 64, dtype=torch.bfloat16, device="cuda")
    wb[3, 8:11] = 1.0                                  # three non-zeros in one group of four
    with pytest.raises(ValueError):
        ops.compress_2to4_bf16(wb)
    assert ops.compress_2to4_bf16(wb, check=False).comp.shape == (16, 64)


@pytest.mark.parametrize("cta_group,tile", [(1, 0), (2, 256), (2, 480), (2, 240), (2, -256), (2, -480), (2, -240)])   # negative: plain-store epilogue
@pytest.mark.parametrize("shape", [(256, 128, 128), (512, 256, 448), (300, 200, 264), (77, 300, 136), (1, 8, 72), (1000, 1536, 2048), (963, 520, 392)])
def test_gemm_sp_exact_products(ops, shape, cta_group, tile):
    """Small-integer operands: the sparse kernel must return the exact integer matmul, for both CTA-group modes and both
    pair tiles (256 tokens / one accumulator, 480 tokens / two accumulators of one tile, 240 tokens / two accumulators ping-ponged)."""
    from qsi_b200 import _lib
    T, N, K = shape
    g = torch.Generator().manual_seed(T + 3 * N + K)
    wb = (_random_2to4(N, -(-K // 8) * 8, g) * 4).round().clamp(-15, 15).to(torch.bfloat16)
    wb[:, K:] = 0
    xb = torch.zeros(T, wb.shape[1], dtype=torch.bfloat16)
    xb[:, :K] = torch.randint(-15, 16, (T, K), generator=g).to(torch.bfloat16)
    bias = torch.randint(-5, 6, (N,), generator=g).float()
    ref = xb.double() @ wb.double().t() + bias.double()
    _lib.set_option("gemm_sp_cta_group", cta_group)
    _lib.set_option("gemm_sp_tile", abs(tile))
    _lib.set_option("gemm_out_tma", 0 if tile < 0 else 1)
    try:
        y = ops.bfp_linear_bf16_sp(xb.cuda(), ops.compress_2to4_bf16(wb.cuda()), bias.cuda())
    finally:
        _lib.set_option("gemm_sp_cta_group", 0)
        _lib.set_option("gemm_sp_tile", 0)
        _lib.set_option("gemm_out_tma", 1)
    assert torch.equal(y.double().cpu(), ref)
    # half-precision outputs are the same fp32 accumulators rounded once in the epilogue
    if wb.shape[0] % 8 == 0:
        _lib.set_option("gemm_sp_cta_group", cta_group)
        _lib.set_option("gemm_sp_tile", abs(tile))
        _lib.set_option("gemm_out_tma", 0 if tile < 0 else 1)
        try:
            ws = ops.compress_2to4_bf16(wb.cuda())
            for dt in (torch.float16, torch.bfloat16):
                yh = ops.bfp_linear_bf16_sp(xb.cuda(), ws, bias.cuda(), out_dtype=dt)
                assert yh.dtype == dt and torch.equal(yh, y.to(dt)), dt
        finally:
            _lib.set_option("gemm_sp_cta_group", 0)
            _lib.set_option("gemm_sp_tile", 0)
            _lib.set_option("gemm_out_tma", 1)


@pytest.mark.parametrize("NK", [(4096, 4096), (11008, 4096), (4096, 11008)])
def test_gemm_sp_llama7b_shapes_full_size(ops, NK):
    """BASELINE LLaMA-7B shapes, T = 4096, HBFP8 block 64, 2:4 s->q weights: sparse kernel == dense exact-bf16 kernel on the
    uncompressed operand (same products, fp32 accumulation), and within 1e-5 of the fp64 matmul."""
    N, K = NK
    T = 4096
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(T, K, device="cuda", generator=g)
    x[::97, ::53] *= 20.0
    w = torch.randn(N, K, device="cuda", generator=g) * 0.02
    a = ops.unpack_bfp_args(dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", mant_bits=7, block_size=64,
                                 w_sparsity=True, N=2, M=4, first="s", sparsity_mode="structured", device="cuda"))
    xb, wb = ops.pack_bfp_bf16(x, identifier="in", **a), ops.pack_bfp_bf16(w, identifier="w", **a)
    y = ops.bfp_linear_bf16_sp(xb, ops.compress_2to4_bf16(wb))
    yd = ops.bfp_linear_bf16(xb, wb)
    assert ((y - yd).norm() / yd.norm()).item() <= 1e-6
    rows = torch.arange(0, T, 61, device="cuda")
    ref = xb[rows].double() @ wb.double().t()
    assert ((y[rows].double() - ref).norm() / ref.norm()).item() <= 1e-5


# ---------------------------------------------------------------------------------------------------------------------
# training on the tensor cores (row f3): forward, dgrad and wgrad of BFPLinear vs the reference's structure
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("w_sparse", [True, False])
@pytest.mark.parametrize("shape", [(3, 70, 512, 384), (5, 33, 200, 136), (1, 256, 1024, 768)])
def test_bfplinear_training_on_tensor_cores(ops, shape, w_sparse, monkeypatch):
    """grad_x = Q_g(gy) Q_w(w), grad_w = Q_g(gy)^T Q_in(x), grad_b = sum Q_g(gy) (bfp_ops.py:160-192) -- the tensor-core
    autograd path against (1) the fake-quant path through torch autograd and (2) an fp64 restatement; tolerance 1e-5."""
    b0, b1, K, N = shape
    kw = dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=7, block_size=64,
              w_sparsity=w_sparse, N=2, M=4, first="s", sparsity_mode="structured", device="cuda")
    torch.manual_seed(b0 + K)
    lin = ops.BFPLinear(K, N, bias=True, **dict(kw)).cuda()
    x0 = torch.randn(b0, b1, K, device="cuda")
    gy = torch.randn(b0, b1, N, device="cuda") * 0.1
    res = {}
    for path in ("tc", "fakequant"):
        monkeypatch.setenv("BFP_TRAIN_PATH", path)
        monkeypatch.setenv("BFP_LINEAR_PATH", "tc" if path == "tc" else "fakequant")
        lin.zero_grad()
        x = x0.clone().requires_grad_(True)
        y = lin(x)
        y.backward(gy)
        res[path] = (y.detach(), x.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone())
        if path == "tc":
            assert y.grad_fn.name().startswith("_BFPLinearTC"), y.grad_fn.name()
    a = ops.unpack_bfp_args(dict(kw))
    xq = ops.float_to_bfp_blocked(x0, **a, identifier="in").double().view(-1, K)
    wq = ops.float_to_bfp_blocked(lin.weight.detach(), **a, identifier="w").double()
    gq = ops.float_to_bfp_blocked(gy, **a, identifier="grad").double().view(-1, N)
    ref = ((xq @ wq.t() + lin.bias.detach().double()).view(b0, b1, N), (gq @ wq).view(b0, b1, K), gq.t() @ xq, gq.sum(0))
    for got in res.values():
        for g, r in zip(got, ref):
            assert ((g.double() - r).norm() / r.norm()).item() <= 1e-5
    if w_sparse:
        assert (res["tc"][2].view(N, -1, 4) != 0).sum(-1).float().mean() > 3.5      # STE: the weight gradient is dense


def test_bfplinear_training_stochastic_rounding_runs_on_tensor_cores(ops):
    """rounding_mode='stoc' (the reference's default): weights are re-quantised on every forward, the backward uses the
    forward's draw; the result is an unbiased estimate of the fp32 linear, so a mean over draws converges."""
    kw = dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="stoc", epsilon=1e-8, mant_bits=5, block_size=32, device="cuda")
    torch.manual_seed(3)
    lin = ops.BFPLinear(256, 128, bias=False, **dict(kw)).cuda()
    x = torch.randn(64, 256, device="cuda", requires_grad=True)
    ys = []
    for _ in range(64):
        y = lin(x)
        assert y.grad_fn.name().startswith("_BFPLinearTC")
        ys.append(y.detach())
    y.sum().backward()
    assert x.grad is not None and torch.isfinite(x.grad).all() and lin.weight.grad.shape == lin.weight.shape
    assert not torch.equal(ys[0], ys[1])                              # fresh uniforms per call
    exact = x.detach() @ lin.weight.detach().t()
    err1 = (ys[0] - exact).norm() / exact.norm()
    errm = (torch.stack(ys).mean(0) - exact).norm() / exact.norm()
    assert errm < 0.35 * err1, (float(err1), float(errm))            # averaging 64 draws shrinks the error ~8x


# ---------------------------------------------------------------------------------------------------------------------
# row f4: BFPConv2d as im2col + BFP GEMM, F_matmul_bfp / F_linear_bfp on the tensor cores
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", [
    dict(B=4, C=3, H=64, W=64, O=96, k=16, stride=16, padding=0, dilation=1),       # ViT patch embedding (kernel = stride)
    dict(B=2, C=8, H=30, W=34, O=40, k=3, stride=1, padding=1, dilation=1),         # overlapping windows + zero padding
    dict(B=3, C=5, H=33, W=40, O=24, k=(3, 5), stride=(2, 1), padding=(0, 2), dilation=(2, 1)),
])
@pytest.mark.parametrize("w_sparse", [False, True])
def test_bfpconv2d_im2col_tensor_core_path(ops, cfg, w_sparse, monkeypatch):
    kw = dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=5, block_size=64,
              w_sparsity=w_sparse, N=2, M=4, first="s", sparsity_mode="structured", device="cuda")
    torch.manual_seed(cfg["O"])
    conv = ops.BFPConv2d(cfg["C"], cfg["O"], cfg["k"], cfg["stride"], cfg["padding"], cfg["dilation"], 1, True, **dict(kw)).cuda()
    x = torch.randn(cfg["B"], cfg["C"], cfg["H"], cfg["W"], device="cuda")
    a = ops.unpack_bfp_args(dict(kw))
    patch = cfg["k"] == cfg["stride"]
    if not patch:
        monkeypatch.setenv("BFP_CONV_IM2COL_MAX_EXPANSION", "1e9")                  # overlapping windows: im2col is opt-in
    from qsi_b200 import _lib
    with torch.no_grad():
        n0 = _lib.lib().bfp_launch_count()
        y = conv(x)
        assert _lib.lib().bfp_launch_count() - n0 == 3                              # two packs + one tcgen05 GEMM
        monkeypatch.setenv("BFP_LINEAR_PATH", "fakequant")
        y_fq = conv(x)                                                              # the reference's structure (cuDNN on fake-quant)
        monkeypatch.setenv("BFP_LINEAR_PATH", "tc")
    xq = ops.float_to_bfp_blocked(x, **a, identifier="in").double()
    wq = ops.float_to_bfp_blocked(conv.weight.detach(), **a, identifier="w").double()
    ref = torch.nn.functional.conv2d(xq, wq, conv.bias.detach().double(), conv.stride, conv.padding, conv.dilation, 1)
    assert y.shape == ref.shape == y_fq.shape
    assert ((y.double() - ref).norm() / ref.norm()).item() <= 1e-5
    assert ((y_fq.double() - ref).norm() / ref.norm()).item() <= 2e-3               # cuDNN may use TF32


def test_f_matmul_and_f_linear_bfp_on_tensor_cores(ops):
    kw = dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=7, block_size=32, device="cuda")
    a = ops.unpack_bfp_args(dict(kw))
    mm, lin = ops.F_matmul_bfp(**dict(kw)), ops.F_linear_bfp(**dict(kw))
    g = torch.Generator(device="cuda").manual_seed(9)

    def q(t, ident):
        return ops.float_to_bfp_blocked(t, **a, identifier=ident).double()

    for xs, ws in (((2, 4, 48, 64), (2, 4, 64, 40)), ((5, 100, 96), (96, 72)), ((3, 1, 20, 128), (1, 6, 128, 24)), ((130, 200), (200, 264))):
        x = torch.randn(*xs, device="cuda", generator=g)
        w = torch.randn(*ws, device="cuda", generator=g)
        with torch.no_grad():
            y = mm(x, w)
        ref = torch.matmul(q(x, "in"), q(w.transpose(-1, -2).contiguous(), "w").transpose(-1, -2))
        assert y.shape == ref.shape and ((y.double() - ref).norm() / ref.norm()).item() <= 1e-5, (xs, ws)
    x = torch.randn(7, 33, 160, device="cuda", generator=g)
    w = torch.randn(88, 160, device="cuda", generator=g)
    b = torch.randn(88, device="cuda", generator=g)
    with torch.no_grad():
        y = lin(x, w, b)
    ref = q(x, "in") @ q(w, "w").t() + b.double()
    assert ((y.double() - ref).norm() / ref.norm()).item() <= 1e-5


def test_bfplinear_forward_is_cuda_graph_capturable(ops):
    """Everything BFPLinear.forward launches (activation pack + tensor-core GEMM; cached weight) goes to the current stream
    with no host synchronisation, so it can be captured once and replayed (the launch-bound regime of small layers)."""
    kw = dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=7, block_size=64,
              w_sparsity=True, N=2, M=4, first="s", sparsity_mode="structured", device="cuda")
    torch.manual_seed(11)
    lins = [ops.BFPLinear(512, 768, bias=True, **dict(kw)).cuda().eval(), ops.BFPLinear(768, 256, bias=False, **dict(dict(kw), w_sparsity=False)).cuda().eval()]
    x = torch.randn(300, 512, device="cuda")
    with torch.no_grad():
        ref = lins[1](lins[0](x))                       # also packs + caches the weights outside the capture
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            lins[1](lins[0](x))
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            y = lins[1](lins[0](x))
        for trial in range(3):
            x.copy_(torch.randn(300, 512, device="cuda"))
            g.replay()
            torch.cuda.synchronize()
            assert torch.equal(y, lins[1](lins[0](x))), trial
    assert ref.shape == y.shape


def test_packed_activation_cache_is_safe(ops):
    """Sibling projections reuse the packed form of the same input tensor object; an in-place write or a different tensor
    (even at the same address) must miss."""
    kw = dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=7, block_size=64, device="cuda")
    from qsi_b200 import _lib
    torch.manual_seed(2)
    q, k = ops.BFPLinear(256, 128, bias=False, **dict(kw)).cuda().eval(), ops.BFPLinear(256, 64, bias=False, **dict(kw)).cuda().eval()
    x = torch.randn(96, 256, device="cuda")
    with torch.no_grad():
        q(x); k(x)                                           # weights packed, x's packed form cached
        n0 = _lib.launch_count(); yq = q(x); yk = k(x); same = _lib.launch_count() - n0
        assert same == 2                                     # two GEMMs; the pack of x was made by the calls above
        x.mul_(2.0)                                          # in-place: version bump -> repack
        n0 = _lib.launch_count(); yq2 = q(x); assert _lib.launch_count() - n0 == 2
        assert torch.equal(yq2, 2 * yq)
        ptr = x.data_ptr(); del x
        z = torch.randn(96, 256, device="cuda")              # very likely the same address, a different tensor
        yz = q(z)
        assert torch.equal(yz, ops.bfp_linear_bf16(ops.pack_bfp_bf16(z, identifier="in", **q.bfp_args), q._packed_weight("bf16")))
        assert ptr == ptr and yk.shape == (96, 64)


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("w_sparse", [True, False])
def test_bfplinear_half_precision_inference_on_tensor_cores(ops, dt, w_sparse, monkeypatch):
    """fp16 / bf16 modules (how the reference runs LLaMA): the tensor-core path accumulates in fp32 and rounds to the dtype once,
    like the library HGEMM the reference calls on the fake-quantised operands.  Against the fp64 contraction of the SAME
    quantised operands the result must be the correctly rounded value up to one unit in the last place."""
    kw = dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=7, block_size=64,
              w_sparsity=w_sparse, N=2, M=4, first="s", sparsity_mode="structured", device="cuda")
    torch.manual_seed(5)
    lin = ops.BFPLinear(1024, 768, bias=True, **dict(kw)).cuda().to(dt)
    x = torch.randn(4, 130, 1024, device="cuda").to(dt)
    a = ops.unpack_bfp_args(dict(kw))
    with torch.no_grad():
        y = lin(x)
        assert y.dtype == dt and lin._packed_w is not None and lin._packed_w[0][0] == ("sp" if w_sparse else "bf16")
        monkeypatch.setenv("BFP_LINEAR_PATH", "fakequant")
        y_ref = lin(x)                                       # the reference's structure: fused fake-quant + torch HGEMM
        monkeypatch.setenv("BFP_LINEAR_PATH", "tc")
    xq = ops.float_to_bfp_blocked(x, **a, identifier="in").double()
    wq = ops.float_to_bfp_blocked(lin.weight.detach(), **a, identifier="w").double()
    exact = xq @ wq.t() + lin.bias.detach().double()
    for got in (y, y_ref):
        err = (got.double() - exact).abs()
        ulp = torch.maximum(exact.abs(), torch.tensor(1e-3, device="cuda", dtype=torch.float64)) * (2.0 ** (-10 if dt == torch.float16 else -7))
        assert (err <= ulp).all()
    assert ((y.double() - y_ref.double()).abs() <= 2.0 ** (-9 if dt == torch.float16 else -6) * exact.abs().clamp_min(1e-3)).all()


@pytest.mark.parametrize("dt", [torch.float32, torch.float16])
@pytest.mark.parametrize("first", ["s", "q"])
def test_bfplinear_unstructured_sparsity_runs_on_tensor_cores(ops, dt, first, monkeypatch):
    """sparsity_mode='unstructured' (4 of the reference's 7 LM scripts): the weight goes through the radix-select kernels and
    the quantiser once, is cached as an exact-bf16 operand, and the contraction runs on the dense tcgen05 kernel -- the
    same function as fake-quant + F.linear."""
    kw = dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=5, block_size=64,
              w_sparsity=True, N=2, M=4, first=first, sparsity_mode="unstructured", sparsity_frac=0.5, device="cuda")
    torch.manual_seed(7)
    lin = ops.BFPLinear(1000, 520, bias=True, **dict(kw)).cuda().to(dt)          # K = 1000: ragged last block, padded pack
    x = torch.randn(2, 77, 1000, device="cuda").to(dt)
    a = ops.unpack_bfp_args(dict(kw))
    with torch.no_grad():
        y = lin(x)
        assert y.dtype == dt and lin._packed_w is not None and lin._packed_w[0][0] == "bf16"
        wb = lin._packed_weight("bf16")
        wq = ops.float_to_bfp_blocked(lin.weight.detach(), **a, identifier="w")
        assert torch.equal(wb[:, :1000].float(), wq.float()) and (wb[:, 1000:] == 0).all()
        assert abs(float((wq == 0).float().mean()) - 0.5) < 0.02
        monkeypatch.setenv("BFP_LINEAR_PATH", "fakequant")
        y_fq = lin(x)
        monkeypatch.setenv("BFP_LINEAR_PATH", "tc")
    xq = ops.float_to_bfp_blocked(x, **a, identifier="in").double()
    exact = xq @ wq.double().t() + lin.bias.detach().double()
    tol = 1e-5 if dt == torch.float32 else 2.0 ** -10
    for got in (y, y_fq):
        assert float((got.double() - exact).norm() / exact.norm()) <= tol
    # F_matmul_bfp with a broadcast weight: the global threshold is taken over the weight's own shape
    mm = ops.F_matmul_bfp(**dict(kw))
    xb, w2 = torch.randn(3, 40, 256, device="cuda").to(dt), (torch.randn(256, 72, device="cuda") * 0.1).to(dt)
    w3 = w2.unsqueeze(0)                                                        # [1, K, N] -> broadcast over the batch of 3
    with torch.no_grad():
        y2, y3 = mm(xb, w2), mm(xb, w3)
        wq2 = ops.float_to_bfp_blocked(w2.t().contiguous(), **a, identifier="w").t().double()
        e2 = ops.float_to_bfp_blocked(xb, **a, identifier="in").double() @ wq2
    for got in (y2, y3):
        assert got.shape == (3, 40, 72) and float((got.double() - e2).norm() / e2.norm()) <= tol


@pytest.mark.parametrize("tile", [0, 256, 480, 240])
@pytest.mark.parametrize("shape", [(300, 200, 264), (1000, 1536, 1024), (77, 300, 72), (963, 520, 392)])
def test_gemm_accumulating_variants_add_exactly(ops, shape, tile):
    """bfp_gemm_bf16_acc / bfp_gemm_bf16_sp_acc: out += A . B^T through the TMA reduce-add epilogue (the K-chunked
    contraction).  Small-integer operands: two chunks must give the exact integer matmul of the concatenated K."""
    from qsi_b200 import _lib
    T, N, K = shape
    g = torch.Generator().manual_seed(T + N + K)
    Kp = -(-K // 8) * 8
    L, st = _lib.lib(), torch.cuda.current_stream().cuda_stream
    xs = [torch.randint(-15, 16, (T, Kp), generator=g).to(torch.bfloat16).cuda() for _ in range(3)]
    wd = [torch.randint(-15, 16, (N, Kp), generator=g).to(torch.bfloat16).cuda() for _ in range(3)]
    wsp = [(_random_2to4(N, Kp, g) * 4).round().clamp(-15, 15).to(torch.bfloat16).cuda() for _ in range(3)]
    out = torch.empty(T, N, device="cuda")
    _lib.check(L.bfp_gemm_bf16(xs[0].data_ptr(), wd[0].data_ptr(), None, out.data_ptr(), T, N, Kp, st))
    for i in (1, 2):
        _lib.check(L.bfp_gemm_bf16_acc(xs[i].data_ptr(), wd[i].data_ptr(), out.data_ptr(), T, N, Kp, st))
    ref = sum(x.double() @ w.double().t() for x, w in zip(xs, wd))
    assert torch.equal(out.double(), ref)
    _lib.set_option("gemm_sp_tile", tile)
    try:
        comp = [ops.compress_2to4_bf16(w) for w in wsp]
        _lib.check(L.bfp_gemm_bf16_sp(xs[0].data_ptr(), comp[0].comp.data_ptr(), comp[0].meta.data_ptr(), None, out.data_ptr(), T, N, Kp, st))
        for i in (1, 2):
            _lib.check(L.bfp_gemm_bf16_sp_acc(xs[i].data_ptr(), comp[i].comp.data_ptr(), comp[i].meta.data_ptr(), out.data_ptr(), T, N, Kp, st))
    finally:
        _lib.set_option("gemm_sp_tile", 0)
    ref = sum(x.double() @ w.double().t() for x, w in zip(xs, wsp))
    assert torch.equal(out.double(), ref)
    # no TMA path (N % 4 != 0): the accumulating form refuses instead of silently overwriting
    if N % 4 == 0:
        o2 = torch.empty(T, N - 1, device="cuda")
        assert L.bfp_gemm_bf16_acc(xs[0].data_ptr(), wd[0].data_ptr(), o2.data_ptr(), T, N - 1, Kp, st) == _lib.E_UNSUPPORTED


@pytest.mark.parametrize("mode", ["structured", "unstructured"])
def test_bfplinear_input_sparsity_on_tensor_cores(ops, mode, monkeypatch):
    """in_sparsity (activations pruned too): the pack kernel applies the N:M mask to x as well (or composes the global
    pruning), and the contraction still runs on the tensor cores -- same function as fake-quant + F.linear."""
    kw = dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=7, block_size=32,
              in_sparsity=True, w_sparsity=True, N=2, M=4, first="s", sparsity_mode=mode, sparsity_frac=0.4, device="cuda")
    torch.manual_seed(11)
    lin = ops.BFPLinear(512, 264, bias=True, **dict(kw)).cuda()
    x = torch.randn(5, 40, 512, device="cuda")
    a = ops.unpack_bfp_args(dict(kw))
    with torch.no_grad():
        y = lin(x)
        assert lin._packed_w is not None and lin._packed_w[0][0] == ("sp" if mode == "structured" else "bf16")
        monkeypatch.setenv("BFP_LINEAR_PATH", "fakequant")
        y_fq = lin(x)
        monkeypatch.setenv("BFP_LINEAR_PATH", "tc")
    xq = ops.float_to_bfp_blocked(x, **a, identifier="in")
    assert abs(float((xq == 0).float().mean()) - (0.5 if mode == "structured" else 0.4)) < 0.03
    exact = xq.double() @ ops.float_to_bfp_blocked(lin.weight.detach(), **a, identifier="w").double().t() + lin.bias.detach().double()
    for got in (y, y_fq):
        assert float((got.double() - exact).norm() / exact.norm()) <= 1e-5


def test_bfplinear_training_with_grad_sparsity_on_tensor_cores(ops, monkeypatch):
    """grad_sparsity (bfp_ops.py:136-137): the N:M mask of the output gradient is part of the same pack as its quantisation, so
    the tensor-core autograd Function covers it; gradients equal the reference's structure (fake-quant + torch autograd)."""
    kw = dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=7, block_size=64,
              w_sparsity=True, grad_sparsity=True, N=2, M=4, first="s", sparsity_mode="structured", device="cuda")
    torch.manual_seed(2)
    lin = ops.BFPLinear(256, 128, bias=True, **dict(kw)).cuda()
    x = torch.randn(64, 256, device="cuda", requires_grad=True)
    y = lin(x)
    assert y.requires_grad and y.grad_fn.name().startswith("_BFPLinearTC")
    y.square().sum().backward()
    gx, gw, gb = x.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone()
    x.grad = None; lin.weight.grad = None; lin.bias.grad = None
    monkeypatch.setenv("BFP_LINEAR_PATH", "fakequant")
    monkeypatch.setenv("BFP_TRAIN_PATH", "fakequant")
    y2 = lin(x)
    assert not y2.grad_fn.name().startswith("_BFPLinearTC")
    y2.square().sum().backward()
    for a, b in ((gx, x.grad), (gw, lin.weight.grad), (gb, lin.bias.grad)):
        assert ((a.double() - b.double()).norm() / b.double().norm()).item() <= 1e-5
    # the gradient really was sparsified: half of every group of four output-gradient entries contributes nothing to grad_w's rows
    assert (gx != 0).any()


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("w_sparse", [False, True])
def test_bfplinear_half_precision_training_on_tensor_cores(ops, dt, w_sparse, monkeypatch):
    """fp16 / bf16 modules (what run_llama.py loads) train on the tensor cores too: fp32 accumulation in TMEM, one rounding to the
    dtype -- the semantics of the library HGEMM the reference's structure calls on the fake-quantised half tensors -- and the output
    gradient quantised in its own dtype.  Against the fake-quant path through torch autograd, within a few ulps of the dtype."""
    kw = dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=7, block_size=64,
              w_sparsity=w_sparse, N=2, M=4, first="s", sparsity_mode="structured", device="cuda")
    torch.manual_seed(11)
    lin = ops.BFPLinear(512, 256, bias=True, **dict(kw)).cuda().to(dt)
    x0 = (torch.randn(4, 96, 512, device="cuda") * 0.5).to(dt)
    gy = (torch.randn(4, 96, 256, device="cuda") * 0.1).to(dt)
    res = {}
    for path in ("tc", "fakequant"):
        monkeypatch.setenv("BFP_TRAIN_PATH", path)
        monkeypatch.setenv("BFP_LINEAR_PATH", "tc" if path == "tc" else "fakequant")
        lin.zero_grad()
        x = x0.clone().requires_grad_(True)
        y = lin(x)
        assert y.dtype == dt
        y.backward(gy)
        res[path] = (y.detach(), x.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone())
        assert y.grad_fn.name().startswith("_BFPLinearTC") == (path == "tc")
    tol = 4e-3 if dt == torch.float16 else 2e-2                          # a few ulps of the dtype (different accumulation orders round differently)
    for g, r in zip(res["tc"], res["fakequant"]):
        assert g.dtype == r.dtype
        assert ((g.double() - r.double()).norm() / r.double().norm()).item() <= tol


@pytest.mark.parametrize("shape", [(5, 200, 136, 72), (12, 512, 512, 64), (7, 300, 64, 512), (3, 1, 8, 8), (2, 1000, 520, 264)])
@pytest.mark.parametrize("dt", [torch.float32, torch.float16])
def test_gemm_bf16_batched_exact_products(ops, shape, dt):
    """bfp_gemm_bf16_batched: every entry of the batch is the exact integer matmul of its own operands -- also when tiles run
    past an entry's rows (T, N not tile multiples): the 3-D output map clips them, nothing leaks into the next entry."""
    from qsi_b200 import _lib
    b, T, N, K = shape
    g = torch.Generator().manual_seed(sum(shape))
    xs = torch.randint(-15, 16, (b, T, K), generator=g).to(torch.bfloat16).cuda()
    ws = torch.randint(-15, 16, (b, N, K), generator=g).to(torch.bfloat16).cuda()
    out = torch.full((b, T, N), float("nan"), dtype=dt, device="cuda")
    guard = torch.full((1 << 16,), 7.0, dtype=dt, device="cuda")              # likely right behind `out` in the allocator's block
    _lib.check(_lib.lib().bfp_gemm_bf16_batched(xs.data_ptr(), ws.data_ptr(), out.data_ptr(), _lib.DT_F32 if dt == torch.float32 else _lib.DT_F16,
                                                b, T, N, K, torch.cuda.current_stream().cuda_stream))
    ref = torch.bmm(xs.double(), ws.double().transpose(1, 2))
    assert torch.equal(out.double(), ref.to(dt).double())
    assert (guard == 7.0).all()


def test_f_matmul_bfp_attention_shapes_one_launch(ops):
    """GPT-2 style attention products (modeling_gpt2.py:205-207, :295): [B, H, S, D] x [B, H, D, S] and [B, H, S, S] x [B, H, S, D]
    through F_matmul_bfp -- one batched launch each, equal to the per-entry contraction of the same quantised operands."""
    from qsi_b200 import _lib
    kw = dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=7, block_size=64, device="cuda")
    mm = ops.F_matmul_bfp(**dict(kw))
    a = ops.unpack_bfp_args(dict(kw))
    torch.manual_seed(4)
    q, k, v = (torch.randn(2, 12, 200, 64, device="cuda") for _ in range(3))
    with torch.no_grad():
        n0 = _lib.lib().bfp_launch_count()
        s = mm(q, k.transpose(-1, -2))
        n1 = _lib.lib().bfp_launch_count()
        assert n1 - n0 == 3                                                   # two packs + ONE GEMM for 24 heads
        p = torch.softmax(s / 8.0, dim=-1)
        o = mm(p, v)
    qq = ops.float_to_bfp_blocked(q, **a, identifier="in").double()
    kq = ops.float_to_bfp_blocked(k, **a, identifier="w").double()            # blocked along D, the contraction
    assert s.shape == (2, 12, 200, 200) and float((s.double() - qq @ kq.transpose(-1, -2)).abs().max()) <= 1e-4
    pq = ops.float_to_bfp_blocked(p, **a, identifier="in").double()
    vq = ops.float_to_bfp_blocked(v.transpose(-1, -2).contiguous(), **a, identifier="w").double().transpose(-1, -2)
    e = pq @ vq
    assert o.shape == (2, 12, 200, 64) and float((o.double() - e).norm() / e.norm()) <= 1e-5


def test_bfpconv1d_repair_equals_bfplinear_on_the_transposed_weight(ops):
    """BFPConv1D (imported by the reference's GPT-2, never defined there): weight [nx, nf], quantised along the contraction --
    the same function as BFPLinear holding weight^T; inference on the tensor cores, training through autograd."""
    kw = dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=7, block_size=64,
              w_sparsity=True, N=2, M=4, first="s", sparsity_mode="structured", device="cuda")
    torch.manual_seed(8)
    c1 = ops.BFPConv1D(384, 256, **dict(kw)).cuda()
    lin = ops.BFPLinear(256, 384, bias=True, **dict(kw)).cuda()
    with torch.no_grad():
        c1.bias.normal_()
        lin.weight.copy_(c1.weight.t()); lin.bias.copy_(c1.bias)
    x = torch.randn(4, 50, 256, device="cuda")
    with torch.no_grad():
        y, yl = c1(x), lin(x)
    assert y.shape == (4, 50, 384) and torch.allclose(y, yl, rtol=1e-5, atol=1e-5)
    xg = x.clone().requires_grad_(True)
    c1(xg).square().sum().backward()
    assert xg.grad is not None and c1.weight.grad is not None and c1.weight.grad.shape == (256, 384)
    plain = ops.BFPConv1D(384, 256).cuda()                                    # default num_format 'fp32': plain Conv1D
    with torch.no_grad():
        assert torch.allclose(plain(x), x @ plain.weight + plain.bias, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("shape", [(4096, 4096, 4096), (100, 37, 40), (1, 8, 8), (65, 129, 136), (1000, 520, 520), (63, 64, 70)])
def test_transpose_pad_16(ops, shape):
    """bfp_transpose_pad_16: out[c][r] = in[r][c] with zero padding, for strided inputs and ragged edges."""
    R, C, ld_in = shape
    ld_out = -(-R // 8) * 8
    g = torch.Generator().manual_seed(R + C)
    src = torch.randn(R, ld_in, generator=g).to(torch.bfloat16).cuda()
    out = ops._transpose_pad(src, R, C, ld_out)
    assert out.shape == (C, ld_out) and torch.equal(out[:, :R], src[:, :C].t()) and (out[:, R:] == 0).all()


@pytest.mark.parametrize("w_sparse", [True, False])
def test_bfplinear_stochastic_inference_on_tensor_cores(ops, w_sparse, monkeypatch):
    """rounding_mode='stoc' at inference (what the reference's scripts set): both operands are re-quantised per call with fresh
    uniforms and contracted on the tensor cores.  Same distribution as the fake-quant path: fresh draws per call, the 2:4 mask
    exact, the mean over draws converging to the same limit with the same per-draw error."""
    from qsi_b200 import _lib
    kw = dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="stoc", epsilon=1e-8, mant_bits=5, block_size=64,
              w_sparsity=w_sparse, N=2, M=4, first="s", sparsity_mode="structured", device="cuda")
    torch.manual_seed(6)
    lin = ops.BFPLinear(512, 256, bias=True, **dict(kw)).cuda().eval()
    x = torch.randn(3, 40, 512, device="cuda")
    with torch.no_grad():
        y0 = lin(x)                                                            # (first call builds the static 2:4 structure when w_sparse)
        n0 = _lib.lib().bfp_launch_count()
        lin(x)
        assert _lib.lib().bfp_launch_count() - n0 == 3                         # pack x, pack w (compressed form when 2:4), one tcgen05 GEMM
        assert lin._packed_w is None                                           # no quantised weight is cached: every call re-quantises
        ys = torch.stack([lin(x) for _ in range(64)])
        monkeypatch.setenv("BFP_LINEAR_PATH", "fakequant")
        fs = torch.stack([lin(x) for _ in range(64)])
        monkeypatch.setenv("BFP_LINEAR_PATH", "tc")
    assert y0.shape == (3, 40, 256) and not torch.equal(ys[0], ys[1])
    a = ops.unpack_bfp_args(dict(kw, rounding_mode="determ"))
    w_lim = ops._structured_N_M_sparsity(lin.weight.detach(), "cuda", 2, 4) if w_sparse else lin.weight.detach()
    limit = x @ w_lim.t() + lin.bias.detach()                                  # stochastic rounding is unbiased around this
    e_tc, e_fq = (ys[0] - limit).norm() / limit.norm(), (fs[0] - limit).norm() / limit.norm()
    m_tc, m_fq = (ys.mean(0) - limit).norm() / limit.norm(), (fs.mean(0) - limit).norm() / limit.norm()
    assert 0.7 < float(e_tc / e_fq) < 1.4 and float(m_tc) < 0.35 * float(e_tc) and float(m_fq) < 0.35 * float(e_fq)


@pytest.mark.parametrize("dt", [torch.float32, torch.float16, torch.bfloat16])
def test_pack_bf16_stochastic_equals_fake_quant_with_the_same_draws(ops, dt):
    """With the same Philox (seed, offset) the packed bf16 operand is the stochastic fake-quant result value for value (fp32
    for every input dtype, like the reference's torch.rand promotion) -- so the tensor-core path samples the same
    distribution as the fused quantiser that the statistical parity tests pin."""
    from qsi_b200 import _lib
    for shape, (m, B), first, sparse in itertools.product([(64, 1024), (37, 200)], [(7, 64), (5, 32), (3, 16)], ["s", "q"], [True, False]):
        g = torch.Generator().manual_seed(m + B)
        x = (torch.randn(*shape, generator=g) * 0.05).to(dt).cuda()
        args = ops.unpack_bfp_args(dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="stoc", epsilon=1e-8, mant_bits=m,
                                        block_size=B, w_sparsity=sparse, N=2, M=4, first=first, sparsity_mode="structured", device="cuda"))
        order = _lib.ORDER_QUANT_ONLY if not sparse else (_lib.ORDER_SPARSIFY_QUANT if first == "s" else _lib.ORDER_QUANT_SPARSIFY)
        ph = (1234, 77)
        fq = ops._fused(x, order, block_size=B, mant_bits=m, epsilon=1e-8, rounding_mode="stoc", N=2, M=4, philox=ph)
        pb = ops.pack_bfp_bf16(x, identifier="w", philox=ph, **args)
        assert fq.dtype == torch.float32 and torch.equal(pb[:, : shape[-1]].float(), fq), (shape, m, B, first, sparse)


def test_bfplinear_half_precision_stochastic_inference(ops, monkeypatch):
    """fp16 module, rounding_mode='stoc', no bias (the reference's LLaMA scripts): fp32 result like the reference (its stochastic
    quantiser promotes to fp32), tensor cores, same distribution as the fake-quant path."""
    kw = dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="stoc", epsilon=1e-8, mant_bits=5, block_size=64,
              w_sparsity=True, N=2, M=4, first="s", sparsity_mode="structured", device="cuda")
    torch.manual_seed(9)
    lin = ops.BFPLinear(512, 256, bias=False, **dict(kw)).cuda().half().eval()
    x = torch.randn(3, 40, 512, device="cuda").half()
    from qsi_b200 import _lib
    with torch.no_grad():
        y = lin(x)
        n0 = _lib.lib().bfp_launch_count()
        lin(x)
        assert _lib.lib().bfp_launch_count() - n0 == 3 and y.dtype == torch.float32
        ys = torch.stack([lin(x) for _ in range(32)])
        monkeypatch.setenv("BFP_LINEAR_PATH", "fakequant")
        f0 = lin(x)
        fs = torch.stack([lin(x) for _ in range(32)])
    assert f0.dtype == torch.float32
    limit = x.float() @ ops._structured_N_M_sparsity(lin.weight.detach(), "cuda", 2, 4).float().t()
    e_tc, e_fq = (ys[0] - limit).norm() / limit.norm(), (fs[0] - limit).norm() / limit.norm()
    assert 0.7 < float(e_tc / e_fq) < 1.4 and float((ys.mean(0) - limit).norm() / limit.norm()) < 0.4 * float(e_tc)


@pytest.mark.parametrize("dt", [torch.float32, torch.float16])
@pytest.mark.parametrize("nm", [(2, 4), (1, 4)])
def test_static_mask_compressed_weight_equals_regular_sparse_path(ops, dt, nm):
    """Stochastic inference with first == 's' quantises the cached COMPRESSED weight (block size B/2, static metadata).  With
    nearest rounding the same machinery must reproduce the regular path -- quantise the full weight, then compress -- exactly."""
    N_, M_ = nm
    kw = dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=5, block_size=64,
              w_sparsity=True, N=N_, M=M_, first="s", sparsity_mode="structured", device="cuda")
    torch.manual_seed(13)
    lin = ops.BFPLinear(1000, 264, bias=False, **dict(kw)).cuda().to(dt)      # K = 1000: ragged last block, K padded to 1024
    with torch.no_grad():
        lin.weight[5, 1::2] = 0                                                # half the row zero: groups with fewer than two survivors
        lin.weight[6, ::4] = 0.25                                              # exact powers of two: the exponent's edge case
    x = torch.randn(70, 1000, device="cuda").to(dt)
    a = ops.unpack_bfp_args(dict(kw))
    with torch.no_grad():
        y = lin(x)
        assert lin._packed_w[0][0] == "sp"
        comp, meta = lin._static_sparse_weight()
        assert comp.dtype == dt and comp.shape == (264, 512)
        wc = ops.pack_bfp_bf16(comp, identifier="w", **dict(a, w_sparsity=False, block_size=32))
        xb = ops.pack_bfp_bf16(x, identifier="in", **a)
        y2 = ops.bfp_linear_bf16_sp(xb, ops.SparseBF16(wc, meta, 264, xb.shape[1]), None, out_dtype=dt)
    assert torch.equal(y, y2)


_KW = dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=7, block_size=64,
           w_sparsity=True, N=2, M=4, first="s", sparsity_mode="structured", device="cuda")


def test_weight_written_through_data_is_never_served_stale(ops, monkeypatch):
    """Writes through `.data` do not bump torch's version counter (the reference's BFPOptim, DeepSpeed / apex master-weight
    copies and pruning scripts all write that way).  Training-mode modules with trainable weights therefore never cache the
    packed weight; eval-mode modules cache it and are refreshed by invalidate_packed() / load_state_dict / .to(); the
    BFP_WEIGHT_CACHE=verify mode catches the write by checksum."""
    torch.manual_seed(0)
    x = torch.randn(96, 256, device="cuda")
    lin = ops.BFPLinear(256, 128, bias=False, **dict(_KW)).cuda()           # nn.Module default: training mode
    with torch.no_grad():
        y0 = lin(x)
        v = lin.weight._version
        lin.weight.data.mul_(0.5)
        assert lin.weight._version == v                                      # the write is invisible to the version counter
        y1 = lin(x)
    assert torch.allclose(y1, 0.5 * y0, rtol=1e-6, atol=1e-7) and not torch.equal(y1, y0)
    lin.eval()
    with torch.no_grad():
        y2 = lin(x)
        assert torch.equal(y2, y1)
        lin.weight.data.mul_(2.0)
        lin.invalidate_packed()
        assert torch.equal(lin(x), y0)
        sd = {k: t.clone() for k, t in lin.state_dict().items()}
        sd["weight"] = sd["weight"] * 0.5
        lin(x)                                                               # packed form of the current weight is cached ...
        lin.load_state_dict(sd)                                              # ... and dropped by load_state_dict (copy_ also bumps the version)
        assert torch.equal(lin(x), y1)
        monkeypatch.setenv("BFP_WEIGHT_CACHE", "verify")
        lin(x)
        lin.weight.data.mul_(2.0)
        assert torch.equal(lin(x), y0)                                       # caught by the checksum, no invalidate call
    # the key also covers the quantiser arguments
    monkeypatch.delenv("BFP_WEIGHT_CACHE")
    with torch.no_grad():
        ya = lin(x)
        lin.bfp_args['mant_bits'] = 3
        yb = lin(x)
    assert not torch.equal(ya, yb)


def test_trainable_bias_alone_gets_its_gradient(ops):
    """BitFit / frozen backbones: only the bias requires grad.  The reference's F.linear(xq, wq, bias) back-propagates to it."""
    torch.manual_seed(1)
    for dt in (torch.float32, torch.float16):
        lin = ops.BFPLinear(256, 128, bias=True, **dict(_KW)).cuda().to(dt)
        lin.weight.requires_grad_(False)
        x = torch.randn(40, 256, device="cuda", dtype=dt)
        y = lin(x)
        assert y.requires_grad and y.dtype == dt
        y.float().sum().backward()
        assert lin.bias.grad is not None and lin.weight.grad is None
        # d(sum y)/d bias = sum over rows of Q_grad(1.0); a block of ones has e = 0 and 1.0 saturates to (2^7 - 1) / 2^7
        assert torch.allclose(lin.bias.grad.float(), torch.full((128,), 40.0 * 127 / 128, device="cuda"))
        op = ops.F_linear_bfp(**dict(_KW))
        b = torch.zeros(128, device="cuda", dtype=dt, requires_grad=True)
        y2 = op(x, lin.weight, b)
        assert y2.requires_grad
        y2.float().sum().backward()
        assert torch.allclose(b.grad.float(), torch.full((128,), 40.0 * 127 / 128, device="cuda"))


# ---------------------------------------------------------------------------------------------------------------
# Block-scaled FP8-class kind (csrc/bfp_gemm_mx.cu): HBFP4 / HBFP5 on tcgen05.mma.kind::mxf8f6f4.block_scale
# ---------------------------------------------------------------------------------------------------------------
def _args(mant_bits=3, block_size=64, w_sparsity=False, first="s"):
    from qsi_b200 import bfp_ops
    return bfp_ops.unpack_bfp_args(dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=mant_bits,
                                        block_size=block_size, w_sparsity=w_sparsity, N=2, M=4, first=first, sparsity_mode="structured", device="cuda"))


def _mx_decode(p):
    """PackedMX -> fp32 [rows, K]: E4M3 byte x 2^(scale byte - 127), restating the operand layout of include/bfp_b200.h."""
    v = p.vals.cpu().numpy().astype(np.int64)
    sign = np.where(v & 0x80, -1.0, 1.0)
    e, mm = (v >> 3) & 15, v & 7
    mag = np.where(e > 0, (1.0 + mm / 8.0) * np.exp2(e - 7.0), mm / 8.0 * 2.0 ** -6)
    vals = (sign * mag)[:, : p.K]
    sf = p.sf.cpu().numpy()
    rows, K, tr = p.rows, p.K, p.tile_rows
    atoms, n_tiles = (tr + 127) // 128, (rows + tr - 1) // tr
    r = np.arange(rows)[:, None]
    k = np.arange(K)[None, :]
    tile, in_tile = r // tr, r % tr
    atom, ra = in_tile // 128, in_tile % 128
    slab, g = (0 * k if p.folded else k // 128), (k % 128) // 32
    off = ((slab * n_tiles + tile) * atoms + atom) * 512 + 16 * (ra % 32) + 4 * (ra // 32) + g
    scale = np.exp2(sf[off].astype(np.float64) - 127.0)
    return (vals * scale).astype(np.float32)


@pytest.mark.parametrize("m,B", [(3, 32), (3, 64), (4, 128)])
def test_mx_pack_contract(ops, m, B):
    """decode(pack_mx(x)) == float_to_bfp_blocked(x), for the general form, the folded (weight) form, the fused activation pack."""
    g = torch.Generator(device="cuda").manual_seed(31)
    a = _args(mant_bits=m, block_size=B, w_sparsity=False)
    for rows, K in ((128, 256), (200, 640), (1, 128), (300, 1024)):
        x = torch.randn(rows, K, device="cuda", generator=g) * torch.exp2(torch.randint(-4, 5, (rows, K // B), device="cuda", generator=g).repeat_interleave(B, 1).float())
        ref = ops.float_to_bfp_blocked(x, **a, identifier="in").cpu().numpy()
        for tile_rows, fold in ((128, False), (240, True), (256, True), (128, True)):
            p = ops.pack_bfp_mx(x, tile_rows, fold=fold, identifier="in", **a)
            assert p is not None
            assert np.array_equal(_mx_decode(p), ref), (rows, K, tile_rows, fold)
        if K % 128 == 0:
            pf = ops.pack_activation_mx(x, a)
            p2 = ops.pack_bfp_mx(x, 128, identifier="in", **a)
            assert torch.equal(pf.vals, p2.vals)
            assert np.array_equal(_mx_decode(pf), ref)
    # half-precision activations through the fused pack
    for dt in (torch.float16, torch.bfloat16):
        x = (torch.randn(130, 512, device="cuda", generator=g) * 0.1).to(dt)
        ref = ops.float_to_bfp_blocked(x, **a, identifier="in").float().cpu().numpy()
        assert np.array_equal(_mx_decode(ops.pack_activation_mx(x, a)), ref), dt
    # a weight whose block exponents span more than the folded form holds is refused (the general form still takes it)
    w = torch.randn(64, 256, device="cuda", generator=g)
    w[:, :B] *= 2.0 ** 20
    assert ops.pack_bfp_mx(w, 128, fold=True, identifier="w", **a) is None
    assert np.array_equal(_mx_decode(ops.pack_bfp_mx(w, 128, identifier="w", **a)), ops.float_to_bfp_blocked(w, **a, identifier="w").cpu().numpy())


@pytest.mark.parametrize("T,N,K,m,B", [(128, 128, 128, 3, 32), (200, 260, 640, 4, 32), (300, 520, 1024, 3, 64), (1000, 1004, 512, 3, 128), (4096, 4096, 4096, 3, 64)])
def test_mx_gemm_matches_oracle(ops, oracle, T, N, K, m, B):
    g = torch.Generator(device="cuda").manual_seed(33)
    a = _args(mant_bits=m, block_size=B, w_sparsity=False)
    x = torch.randn(T, K, device="cuda", generator=g) * torch.exp2(torch.randint(-5, 6, (T, K // B), device="cuda", generator=g).repeat_interleave(B, 1).float())
    w = torch.randn(N, K, device="cuda", generator=g) * 0.05 * torch.exp2(torch.randint(-3, 4, (N, K // B), device="cuda", generator=g).repeat_interleave(B, 1).float())
    bias = torch.randn(N, device="cuda", generator=g)
    xq, wq = ops.float_to_bfp_blocked(x, **a, identifier="in"), ops.float_to_bfp_blocked(w, **a, identifier="w")
    ref = (xq.double() @ wq.double().t() + bias.double())
    xp = ops.pack_activation_mx(x, a)
    for tile_rows, fold in ((240, True), (256, True), (128, True), (240, False), (256, False), (128, False)):
        wp = ops.pack_bfp_mx(w, tile_rows, fold=fold, identifier="w", **a)
        y = ops.bfp_linear_mx(xp, wp, bias)
        rel = float((y.double() - ref).norm() / ref.norm())
        assert rel <= 1e-5, (tile_rows, fold, rel)                       # north_star tolerance; measured ~1e-7
    if T <= 300:
        o = oracle.linear(xq.cpu().numpy(), wq.cpu().numpy(), bias.cpu().numpy())
        assert float(np.linalg.norm(y.cpu().numpy() - o) / np.linalg.norm(o)) <= 1e-5


@pytest.mark.parametrize("w_sparse", [False, True])
def test_bfplinear_hbfp4_takes_the_block_scaled_path(ops, oracle, w_sparse):
    """HBFP4 modules run on the block-scaled MMA (folded weight cached, activation packed by the fused kernel) and give the oracle's
    result; the bf16 kinds stay available through BFP_GEMM_KIND."""
    import os
    g = torch.Generator().manual_seed(35)
    a = _args(mant_bits=3, block_size=64, w_sparsity=w_sparse)
    lin = ops.BFPLinear(512, 384, bias=True, **a).cuda()
    x = torch.randn(3, 50, 512, generator=g).cuda()
    with torch.no_grad():
        y = lin(x)
    assert lin._packed_w[0][0] == "mx" and y.shape == (3, 50, 384)
    xq, _ = oracle.bfp_quantize(x.cpu().numpy(), 64, 3)
    wq, _ = oracle.float_to_bfp_blocked(lin.weight.detach().cpu().numpy(), 3, 64, "sq" if w_sparse else "q")
    ref = oracle.linear(xq.reshape(-1, 512), wq, lin.bias.detach().cpu().numpy()).reshape(3, 50, 384)
    assert float(np.linalg.norm(y.cpu().numpy() - ref) / np.linalg.norm(ref)) <= 1e-5
    os.environ["BFP_GEMM_KIND"] = "bf16"
    try:
        with torch.no_grad():
            y2 = lin(x)
        assert lin._packed_w[0][0] in ("bf16", "sp")
    finally:
        os.environ.pop("BFP_GEMM_KIND")
    assert float((y - y2).norm() / y2.norm()) <= 1e-6
