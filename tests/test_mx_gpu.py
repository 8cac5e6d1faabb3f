"""GPU parity tests of the MX (OCP Microscaling) path (csrc/bfp_ocp_mx.cu, mx_layers.py) against oracle/mx_oracle.py, through the C ABI.
Bar: the quantiser is bit-exact on fp32 tensors (NaN == NaN), every operand form decodes to the fake-quantised tensor bit for bit; the
layers agree with the oracle's float64-accumulated product to the fp32 summation error, i.e. identical except where that error
straddles a bfloat16 rounding boundary (<= 1 bf16 ulp = 2^-7 relative, on a small fraction of the outputs)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

FORMATS = ["int8", "int4", "fp8_e5m2", "fp8_e4m3", "fp6_e3m2", "fp6_e2m3", "fp4_e2m1"]
SPEC = dict(block_size=32, bfloat=16, scale_bits=8)


@pytest.fixture(scope="module")
def mx():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import qsi_b200  # noqa: F401
    from qsi_b200 import mx_layers
    return mx_layers


@pytest.fixture(scope="module")
def O():
    from oracle import mx_oracle
    return mx_oracle


def _data(seed, shape, scale=1.0):
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal(shape) * scale).astype(np.float32)
    x.flat[::97] *= 20.0
    x.flat[5::211] = 0.0
    x.flat[7::223] = -0.0
    return x


def _same(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return int(((a.view(np.uint32) != b.view(np.uint32)) & ~(np.isnan(a) & np.isnan(b))).sum())


def _q(mx, x, fmt, block, bfloat, out_kind=0, flush=False):
    t = x if isinstance(x, torch.Tensor) else torch.from_numpy(x).cuda()
    return mx._mx_quantize_last(t.contiguous(), mx.ELEM_FORMATS[fmt], block, 8, bfloat, flush, out_kind)


@pytest.mark.parametrize("fmt", FORMATS)
@pytest.mark.parametrize("block", [16, 32, 64, 128])
@pytest.mark.parametrize("bfloat", [0, 16])
def test_mx_quantiser_bit_exact_vs_oracle(mx, O, fmt, block, bfloat):
    for seed, scale in ((0, 1.0), (1, 1e-3), (2, 300.0)):
        x = _data(seed, (96, 512), scale)
        want = O.quantize_mx(O.quantize_bfloat(x, bfloat), fmt, block)
        got = _q(mx, x, fmt, block, bfloat).cpu().numpy()
        assert _same(got, want) == 0, (fmt, block, bfloat, seed)


@pytest.mark.parametrize("fmt", ["fp8_e4m3", "int8", "fp4_e2m1"])
def test_mx_quantiser_special_values(mx, O, fmt):
    x = _data(3, (64, 256))
    x[0, :32] = 0.0                                   # all-zero block
    x[1, 3] = np.inf; x[2, 40] = -np.inf; x[3, 70] = np.nan       # noqa: E702  NaN-marked blocks
    x[4, :64] *= 1e-41                                # subnormal blocks: the scale clamps at 2^-127
    x[5, :64] *= 1e-36                                # around the smallest normal numbers
    x[6, :32] = 3e38; x[6, 5] = 1e38                  # noqa: E702  top of the range
    x[9, :32] = 1.0; x[9, 1] = -1.0                   # noqa: E702
    for bfloat in (0, 16):
        want = O.quantize_mx(O.quantize_bfloat(x, bfloat), fmt, 32)
        got = _q(mx, x, fmt, 32, bfloat).cpu().numpy()
        assert _same(got, want) == 0, (fmt, bfloat)
    want = O.quantize_mx(x, fmt, 32, flush_fp32_subnorms=True)
    assert _same(_q(mx, x, fmt, 32, 0, flush=True).cpu().numpy(), want) == 0


def _torch_quantize_mx(A, fmt, block, O):
    """The library's _quantize_mx written with the torch ops it uses, evaluated ON THE GPU (torch-CUDA is the backend the reference
    runs on: same log2f, same pow) -- checks the kernel's bit tricks against the literal op sequence, near powers of two included."""
    ebits, mbits, emax, max_norm = O.FORMATS[fmt]
    rows, K = A.shape
    B = A.view(rows, K // block, block)
    m = B.abs().amax(dim=-1, keepdim=True)
    se = torch.floor(torch.log2(m + float(O.FP32_MIN_NORMAL) * (m == 0).float())) - emax
    se = torch.where(se > 127, torch.full_like(se, float("nan")), se).clamp(min=-127)
    a = B / (2 ** se)
    if ebits:
        pe = torch.floor(torch.log2(a.abs() + (a == 0).float())).clamp(min=2 - 2 ** (ebits - 1))
        out = a / (2 ** pe) * (2 ** (mbits - 2))
    else:
        pe, out = None, a * (2 ** (mbits - 2))
    out = torch.sign(out) * torch.floor(out.abs() + 0.5)
    out = out / (2 ** (mbits - 2)) * (2 ** pe) if ebits else out / (2 ** (mbits - 2))
    out = torch.clamp(out, -max_norm, max_norm)
    return (out * (2 ** se)).view(rows, K)


@pytest.mark.parametrize("fmt", FORMATS)
def test_mx_quantiser_equals_the_literal_torch_cuda_op_sequence(mx, O, fmt):
    x = _data(20, (128, 512))
    k = np.arange(128) - 64
    x[:, 0] = (np.float32(2.0) ** k * (1 - np.float32(2.0 ** -24))).astype(np.float32)       # block maxima one ulp below a power of two:
    x[:, 0] *= 64                                                                              # torch's fp32 log2 may round up to the integer
    x[:, 1:32] *= (np.abs(x[:, :1]) / 64)
    x[:, 33] = np.float32(4.0) * (1 - np.float32(2.0 ** -23))
    t = torch.from_numpy(x).cuda()
    want = _torch_quantize_mx(t, fmt, 32, O)
    got = _q(mx, t, fmt, 32, 0)
    assert _same(got.cpu().numpy(), want.cpu().numpy()) == 0


@pytest.mark.parametrize("shape,block", [((7, 3), 32), ((33, 100), 32), ((5, 70), 64), ((4, 96), 0), ((9, 250), 24), ((2, 3, 5, 36), 32)])
def test_mx_quantiser_generic_shapes(mx, O, shape, block):
    x = _data(4, shape)
    for fmt in ("fp6_e2m3", "int8"):
        want = O.quantize_mx(O.quantize_bfloat(x, 16), fmt, block)
        got = _q(mx, x, fmt, block, 16).cpu().numpy()
        assert got.shape == x.shape and _same(got, want) == 0, (shape, block, fmt)


def test_mx_generic_kernel_equals_stream_kernel(mx, O):
    from qsi_b200 import _lib
    x = torch.from_numpy(_data(5, (64, 512))).cuda()
    a = _q(mx, x, "fp8_e4m3", 32, 16)
    _lib.check(_lib.lib().bfp_set_option(b"force_generic", 1))
    try:
        b = _q(mx, x, "fp8_e4m3", 32, 16)
    finally:
        _lib.check(_lib.lib().bfp_set_option(b"force_generic", 0))
    assert torch.equal(a.view(torch.int32), b.view(torch.int32))


def test_bfloat_round_bit_exact(mx, O):
    x = _data(6, (257, 129))
    x.view(np.uint32)[::3] &= 0xFFFF8000
    x.view(np.uint32)[::3] |= 0x00008000              # exact ties of the bfloat16 rounding
    x[0, :8] = [0.0, -0.0, 1e-40, -1e-40, 3.4e38, -3.4e38, np.inf, np.nan]
    for bfloat in (16, 12, 24):
        sp = mx.finalize_mx_specs(mx.apply_mx_specs(dict(bfloat=bfloat)))
        got = mx.quantize_elemwise_op(torch.from_numpy(x).cuda(), sp).cpu().numpy()
        assert _same(got, O.quantize_bfloat(x, bfloat)) == 0, bfloat
    assert mx.quantize_elemwise_op(torch.ones(3), None) is not None           # mx_specs None: identity, no device needed


def test_half_precision_inputs_compute_in_fp32_and_round_once(mx, O):
    for dt in (torch.bfloat16, torch.float16):
        x = torch.from_numpy(_data(8, (64, 512))).to(dt).cuda()
        got = _q(mx, x, "fp8_e4m3", 32, 16)
        assert got.dtype == dt
        want = torch.from_numpy(O.quantize_mx(O.quantize_bfloat(x.float().cpu().numpy(), 16), "fp8_e4m3", 32)).to(dt)
        assert torch.equal(got.cpu().view(torch.int16), want.view(torch.int16))


@pytest.mark.parametrize("fmt", FORMATS)
def test_bf16_operand_is_exact(mx, O, fmt):
    x = _data(9, (40, 200))                           # K not a multiple of the block: generic kernel; Kp = 200
    want = O.quantize_mx(O.quantize_bfloat(x, 16), fmt, 32)
    got = _q(mx, x, fmt, 32, 16, out_kind=1)
    assert got.dtype == torch.bfloat16 and tuple(got.shape) == (40, 200)
    assert _same(got.float().cpu().numpy(), want) == 0
    x = _data(10, (64, 260))                          # Kp = 264: padded columns are zero
    got = _q(mx, x, fmt, 0, 16, out_kind=1)
    assert tuple(got.shape) == (64, 264) and not got[:, 260:].any()
    assert _same(got[:, :260].float().cpu().numpy(), O.quantize_mx(O.quantize_bfloat(x, 16), fmt, 0)) == 0


def _decode_block_scaled(vals, sf, rows, K, tile_rows):
    """numpy restatement of the operand form of include/bfp_b200.h bfp_ocp_mx_pack: E4M3 bytes + UE8M0 scale atoms."""
    v = vals.astype(np.int32)
    s, e, m = (v >> 7) & 1, (v >> 3) & 15, v & 7
    mag = np.where(e == 0, m * 2.0 ** -9, (1 + m / 8.0) * 2.0 ** (e - 7.0))
    elem = np.where(s == 1, -mag, mag)
    atoms = (tile_rows + 127) // 128
    tiles = (rows + tile_rows - 1) // tile_rows
    r = np.arange(rows)
    rt, rr = r // tile_rows, r % tile_rows
    out = np.empty((rows, K), np.float64)
    for slab in range(K // 128):
        for g in range(4):
            idx = ((slab * tiles + rt) * atoms + (rr >> 7)) * 512 + 16 * ((rr & 127) & 31) + 4 * ((rr & 127) >> 5) + g
            sb = sf[idx].astype(np.int32)
            scale = np.where(sb == 255, np.nan, 2.0 ** (sb - 127.0))
            out[:, slab * 128 + g * 32: slab * 128 + (g + 1) * 32] = elem[:, slab * 128 + g * 32: slab * 128 + (g + 1) * 32] * scale[:, None]
    return out.astype(np.float32)


@pytest.mark.parametrize("fmt", ["fp8_e4m3", "fp6_e3m2", "fp6_e2m3", "fp4_e2m1", "int4"])
@pytest.mark.parametrize("block,tile_rows,rows", [(32, 128, 256), (64, 240, 500), (128, 128, 130), (32, 256, 300)])
def test_block_scaled_operand_decodes_to_fake_quant(mx, O, fmt, block, tile_rows, rows):
    x = _data(11, (rows, 384))
    x[3, 64:96] = 0.0
    x[5, 7] = np.inf
    sp = mx.finalize_mx_specs(mx.apply_mx_specs(dict(SPEC, block_size=block, w_elem_format=fmt, a_elem_format=fmt)))
    p = mx._pack_block_scaled(torch.from_numpy(x).cuda(), mx.ELEM_FORMATS[fmt], tile_rows, sp, 16)
    got = _decode_block_scaled(p.vals.cpu().numpy(), p.sf.cpu().numpy(), rows, 384, tile_rows)
    want = O.quantize_mx(O.quantize_bfloat(x, 16), fmt, block)
    bad = ~((got == want) | (np.isnan(got) & np.isnan(want)))
    assert int(bad.sum()) == 0


def _layer_close(got, want, frac=0.02):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    diff = np.abs(got - want)
    tol = np.maximum(np.abs(want), 1e-30) * 2.0 ** -7 * 1.01
    assert (diff <= tol).all(), float((diff / np.maximum(np.abs(want), 1e-30)).max())
    assert (diff > 0).mean() <= frac, float((diff > 0).mean())


@pytest.mark.parametrize("fmt", FORMATS)
@pytest.mark.parametrize("bias", [False, True])
def test_mxlinear_vs_oracle(mx, O, fmt, bias):
    torch.manual_seed(0)
    lin = mx.MXLinear(512, 384, bias=bias, mx_specs=dict(SPEC, w_elem_format=fmt, a_elem_format=fmt)).cuda().eval()
    x = torch.from_numpy(_data(12, (3, 50, 512))).cuda()
    with torch.no_grad():
        y = lin(x)
    assert y.shape == (3, 50, 384) and y.dtype == torch.float32
    want, _, _ = O.mx_linear(x.cpu().numpy().reshape(-1, 512), lin.weight.detach().cpu().numpy(), lin.bias.detach().cpu().numpy() if bias else None,
                             fmt, fmt, 32, 16, 8)
    _layer_close(y.cpu().numpy().reshape(-1, 384), want)


@pytest.mark.parametrize("fmt,block,K,N", [("fp8_e4m3", 64, 512, 96), ("fp4_e2m1", 128, 256, 1000), ("int8", 16, 72, 30), ("fp8_e4m3", 32, 200, 64),
                                           ("fp6_e2m3", 32, 4096, 4096)])
def test_mxlinear_shapes_and_mixed_formats(mx, O, fmt, block, K, N):
    torch.manual_seed(1)
    wf = "fp4_e2m1" if fmt == "fp8_e4m3" else fmt
    lin = mx.MXLinear(K, N, bias=True, mx_specs=dict(SPEC, block_size=block, w_elem_format=wf, a_elem_format=fmt)).cuda().eval()
    T = 256 if K * N > 1 << 22 else 37
    x = torch.from_numpy(_data(13, (T, K))).cuda()
    with torch.no_grad():
        y = lin(x)
    want, _, _ = O.mx_linear(x.cpu().numpy(), lin.weight.detach().cpu().numpy(), lin.bias.detach().cpu().numpy(), wf, fmt, block, 16, 8)
    _layer_close(y.cpu().numpy(), want)


@pytest.mark.parametrize("fmt", ["fp8_e4m3", "int8"])
@pytest.mark.parametrize("mode", ["structured", "unstructured"])
def test_mxlinear_prunes_its_weight_once_like_the_reference(mx, O, fmt, mode):
    torch.manual_seed(2)
    lin = mx.MXLinear(256, 128, bias=False, mx_specs=dict(SPEC, w_elem_format=fmt, a_elem_format=fmt), sparsity=True, device="cuda", sparsity_mode=mode,
                      sparsity_frac=0.5, N=2, M=4).cuda().eval()
    w0 = lin.weight.detach().clone()
    x = torch.from_numpy(_data(14, (64, 256))).cuda()
    with torch.no_grad():
        y = lin(x)
        y2 = lin(x)
    assert lin.sparsity_init and torch.equal(y, y2)
    w1 = lin.weight.detach()
    assert float((w1 == 0).float().mean()) == pytest.approx(0.5, abs=0.01)
    if mode == "structured":
        g = w0.abs().view(-1, 4)
        keep = torch.zeros_like(g, dtype=torch.bool).scatter_(1, g.topk(2, dim=1).indices, True)
        assert torch.equal(w1.view(-1, 4) != 0, keep)
    want, _, _ = O.mx_linear(x.cpu().numpy(), w1.cpu().numpy(), None, fmt, fmt, 32, 16, 8)
    _layer_close(y.cpu().numpy(), want)


def test_mxlinear_packed_weight_cache_follows_the_weight(mx, O):
    torch.manual_seed(5)
    fmt = "fp8_e4m3"
    lin = mx.MXLinear(256, 128, bias=False, mx_specs=dict(SPEC, w_elem_format=fmt, a_elem_format=fmt)).cuda().eval()
    x = torch.from_numpy(_data(21, (32, 256))).cuda()

    def check():
        with torch.no_grad():
            y = lin(x)
        want, _, _ = O.mx_linear(x.cpu().numpy(), lin.weight.detach().cpu().numpy(), None, fmt, fmt, 32, 16, 8)
        _layer_close(y.cpu().numpy(), want)
    check()
    assert len(lin._packed.entries) == 1
    with torch.no_grad():
        lin.weight.mul_(0.37)                       # in-place through autograd's view of the tensor: the version moves
    check()
    lin.weight.data.mul_(1.7)                       # through .data: the version does NOT move -- the documented protocol is invalidate_packed()
    lin.invalidate_packed()
    check()
    lin.load_state_dict({"weight": torch.randn(128, 256)})          # hooks drop the cache
    check()
    lin.train()                                     # a training module with a trainable weight never serves a cached form
    lin.weight.data.mul_(0.5)
    check()
    lin.eval().half().float()                       # _apply (dtype / device moves) drops it too
    check()


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("fmt", ["fp8_e4m3", "int8"])
def test_mxlinear_half_precision_modules(mx, O, dt, fmt):
    """fp16 / bf16 modules (what run_llama.py loads): computed in fp32, rounded once to the module's dtype."""
    torch.manual_seed(6)
    lin = mx.MXLinear(512, 256, bias=True, mx_specs=dict(SPEC, w_elem_format=fmt, a_elem_format=fmt)).cuda().to(dt).eval()
    x = torch.from_numpy(_data(22, (40, 512))).cuda().to(dt)
    with torch.no_grad():
        y = lin(x)
    assert y.dtype == dt and y.shape == (40, 256)
    want, _, _ = O.mx_linear(x.float().cpu().numpy(), lin.weight.detach().float().cpu().numpy(), lin.bias.detach().float().cpu().numpy(), fmt, fmt, 32, 16, 8)
    _layer_close(y.float().cpu().numpy(), torch.from_numpy(want).to(dt).float().numpy())


def test_mxlinear_without_specs_is_a_plain_linear(mx):
    lin = mx.MXLinear(64, 32, mx_specs=None).cuda()
    x = torch.randn(5, 64, device="cuda")
    assert lin.mx_none and torch.equal(lin(x), torch.nn.functional.linear(x, lin.weight, lin.bias))


def test_mxmatmul_vs_oracle(mx, O):
    a = torch.from_numpy(_data(15, (2, 3, 48, 64))).cuda()
    b = torch.from_numpy(_data(16, (2, 3, 64, 40))).cuda()
    y = mx.MXMatmul(a, b, mx_specs=dict(SPEC, a_elem_format="fp8_e4m3", w_elem_format="fp8_e4m3"))
    want, _, _ = O.mx_matmul(a.cpu().numpy(), b.cpu().numpy(), "fp8_e4m3", 32, 16, 8)
    assert y.shape == (2, 3, 48, 40)
    _layer_close(y.cpu().numpy(), want)
    # the second operand pruned 2:4 along its contraction dim first (mx_layers.py:103-108)
    y = mx.MXMatmul(a, b, mx_specs=dict(SPEC, a_elem_format="int8"), sparsity=True, sparsity_mode="structured", device="cuda", N=2, M=4)
    bt = b.transpose(-1, -2).contiguous()
    g = bt.abs().view(-1, 4)
    keep = torch.zeros_like(g, dtype=torch.bool).scatter_(1, g.topk(2, dim=1).indices, True)
    bs = (bt.view(-1, 4) * keep).view_as(bt).transpose(-1, -2)
    want, _, _ = O.mx_matmul(a.cpu().numpy(), bs.cpu().numpy(), "int8", 32, 16, 8)
    _layer_close(y.cpu().numpy(), want)
    assert torch.equal(mx.MXMatmul(a, b), torch.matmul(a, b))


def test_mxconv2d_vs_oracle(mx, O):
    torch.manual_seed(3)
    conv = mx.MXConv2d(3, 32, kernel_size=4, stride=4, mx_specs=dict(SPEC, a_elem_format="fp8_e4m3", w_elem_format="fp8_e4m3")).cuda().eval()
    x = torch.from_numpy(_data(17, (2, 3, 16, 16))).cuda()
    with torch.no_grad():
        y = conv(x)
    xq = O.quantize_mx(O.quantize_bfloat(x.cpu().numpy(), 16), "fp8_e4m3", 32, axis=1)
    wq = O.quantize_mx(O.quantize_bfloat(conv.weight.detach().cpu().numpy(), 16), "fp8_e4m3", 32, axis=1)
    ref = torch.nn.functional.conv2d(torch.from_numpy(xq).double(), torch.from_numpy(wq).double(), None, stride=4).float().numpy()
    ref = O.quantize_bfloat(ref, 16)
    ref = O.quantize_bfloat(ref + O.quantize_bfloat(conv.bias.detach().cpu().numpy(), 16).reshape(1, -1, 1, 1), 16)
    _layer_close(y.cpu().numpy(), ref)


def test_mxlinear_training_follows_the_library_backward(mx, O):
    torch.manual_seed(4)
    fmt = "fp8_e4m3"
    lin = mx.MXLinear(64, 48, bias=True, mx_specs=dict(SPEC, w_elem_format=fmt, a_elem_format=fmt)).cuda()
    x = torch.from_numpy(_data(18, (40, 64))).cuda().requires_grad_(True)
    y = lin(x)
    gy = torch.from_numpy(_data(19, (40, 48))).cuda()
    y.backward(gy)
    xn, wn, gn = x.detach().cpu().numpy(), lin.weight.detach().cpu().numpy(), gy.cpu().numpy()
    want, _, _ = O.mx_linear(xn, wn, lin.bias.detach().cpu().numpy(), fmt, fmt, 32, 16, 8)
    _layer_close(y.detach().cpu().numpy(), want)
    rb = lambda t: O.quantize_bfloat(t, 16)                                                    # noqa: E731
    g = rb(gn)
    bx, bw = rb(xn), rb(wn)
    gw = rb((O.quantize_mx(g, fmt, 32, axis=0).astype(np.float64).T @ O.quantize_mx(bx, fmt, 32, axis=0).astype(np.float64)).astype(np.float32))
    gx = rb((O.quantize_mx(g, fmt, 32, axis=-1).astype(np.float64) @ O.quantize_mx(bw, fmt, 32, axis=0).astype(np.float64)).astype(np.float32))
    _layer_close(lin.weight.grad.cpu().numpy(), gw)
    _layer_close(x.grad.cpu().numpy(), gx)
    _layer_close(lin.bias.grad.cpu().numpy(), rb(g.sum(0).astype(np.float32)), frac=0.2)


def test_unsupported_options_raise(mx):
    with pytest.raises(NotImplementedError):
        mx.quantize_mx_op(torch.ones(4, 32, device="cuda"), mx.finalize_mx_specs(mx.apply_mx_specs(dict(SPEC))), "fp8_e4m3", axes=[-1], round="floor")
    with pytest.raises(Exception):
        mx._format_id("fp9")
    with pytest.raises(ValueError):
        mx.quantize_elemwise_op(torch.ones(4, device="cuda"), mx.finalize_mx_specs(mx.apply_mx_specs(dict(bfloat=8))))
    with pytest.raises(ValueError):
        mx.quantize_mx_op(torch.ones(4, 32), mx.finalize_mx_specs(mx.apply_mx_specs(dict(SPEC))), "fp8_e4m3", axes=[-1])       # CPU tensor: no fallback


class _EmulatedMXLinear(torch.nn.Linear):
    """mx.Linear written with the torch ops the library's emulation uses (the reference's MXLinear minus the import it cannot satisfy
    here): bfloat16 rounding half away from zero, MX along the contraction dim, fp32 F.linear, rounding, bias, rounding."""

    def __init__(self, src, fmt, block, O):
        super().__init__(src.in_features, src.out_features, bias=src.bias is not None)
        self.weight, self.bias, self.fmt, self.block, self.O = src.weight, src.bias, fmt, block, O

    @staticmethod
    def rb(t):
        return ((t.contiguous().view(torch.int32) + 0x8000) & ~0xFFFF).view(torch.float32)

    def forward(self, x):
        K = x.shape[-1]
        qx = _torch_quantize_mx(self.rb(x).view(-1, K), self.fmt, self.block, self.O).view(x.shape)
        qw = _torch_quantize_mx(self.rb(self.weight.detach()), self.fmt, self.block, self.O)
        y = self.rb(torch.nn.functional.linear(qx, qw))
        return self.rb(y + self.rb(self.bias.detach())) if self.bias is not None else y


def test_opt_model_with_mxlinear_matches_the_emulated_library(mx, O):
    """Model level: stock OPT (OPTConfig() widths, 2 layers, 4 x 256 tokens) with the six block linears of every layer swapped for
    MXLinear (fp8_e4m3, block 32, bfloat16, 2:4-pruned weights) -- the substitution modeling_opt.py:165-169,328-330 makes -- against
    the same model with the emulation written in torch ops.  Individual outputs differ only where the fp32 summation order crosses a
    bfloat16 rounding boundary, so the logits agree to a small multiple of the bfloat16 step."""
    import transformers
    fmt, block = "fp8_e4m3", 32
    targets = ("q_proj", "k_proj", "v_proj", "out_proj", "fc1", "fc2")

    def build(make):
        torch.manual_seed(0)
        cfg = transformers.OPTConfig()
        cfg.num_hidden_layers = 2
        model = transformers.OPTForCausalLM(cfg).eval()
        n = 0
        for parent in list(model.modules()):
            for name, ch in list(parent.named_children()):
                if isinstance(ch, torch.nn.Linear) and name in targets:
                    setattr(parent, name, make(ch))
                    n += 1
        return model.cuda(), cfg, n

    def ours(ch):
        new = mx.MXLinear(ch.in_features, ch.out_features, bias=ch.bias is not None, mx_specs=dict(SPEC, w_elem_format=fmt, a_elem_format=fmt),
                          sparsity=True, device="cuda", sparsity_mode="structured", N=2, M=4)
        new.weight, new.bias = ch.weight, ch.bias
        return new.eval()

    m1, cfg, n1 = build(ours)
    ids = torch.randint(0, cfg.vocab_size, (4, 256), generator=torch.Generator().manual_seed(1)).cuda()
    with torch.no_grad():
        y1 = m1(input_ids=ids).logits.float()
    pruned = {k: v.detach().clone() for k, v in m1.state_dict().items()}          # the weights after MXLinear's one-time 2:4 pruning
    del m1
    m2, _, n2 = build(lambda ch: _EmulatedMXLinear(ch, fmt, block, O))
    m2.load_state_dict(pruned)
    with torch.no_grad():
        y2 = m2(input_ids=ids).logits.float()
    assert n1 == n2 == 12
    rel = float((y1 - y2).norm() / y2.norm())
    print(f"MX OPT drop-in: logits rel. difference vs the emulated library {rel:.3e}")
    assert rel < 1e-3                     # measured 8.9e-5 on B200
