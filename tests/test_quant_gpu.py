"""GPU parity suite for the fused quantise + N:M kernels.  Everything goes through the reference-shaped Python API
(qsi_b200.bfp_ops) -> ctypes -> C ABI (libbfp_b200.so); the oracle (oracle/) is only the checker.

Bar: bit-exact for round-to-nearest values and for the sparsity masks (including the sign of zero); stochastic
rounding is checked bit-exactly given the kernel's own Philox uniforms and statistically for unbiasedness.
"""
import itertools
import os

import numpy as np
import pytest
import torch

import _golden
from _refload import load_reference, ref_args

pytestmark = pytest.mark.gpu

TORCH_DT = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from qsi_b200 import bfp_ops, _lib
    _lib.lib()
    return bfp_ops


def _to_dev(a, dt):
    from oracle import bfp_oracle as O
    return O.to_torch(a, dt).cuda()


def _np(t):
    from oracle import bfp_oracle as O
    return O.from_torch(t)


def _run(ops, x, kind, m, B, N, M, rounding="determ"):
    if kind == "q":
        return ops._no_sparsity_float_to_bfp(x, B, m, 1e-8, rounding, "cuda")
    if kind == "s":
        return ops._structured_N_M_sparsity(x, "cuda", N, M)
    args = ops.unpack_bfp_args(dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode=rounding, epsilon=1e-8,
                                    mant_bits=m, block_size=B, w_sparsity=True, N=N, M=M,
                                    first="s" if kind == "sq" else "q", sparsity_mode="structured", device="cuda"))
    return ops.float_to_bfp_blocked(x, **args, identifier="w")


# ---------------------------------------------------------------------------------------------------------------
# 1. golden fixtures recorded from the reference (torch-CUDA on a B200, and torch-CPU)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("device", ["cuda", "cpu"])
@pytest.mark.parametrize("generic", [0, 1])
def test_kernel_matches_reference_golden(ops, device, generic, monkeypatch):
    from qsi_b200 import _lib
    z = _golden.load(device)
    if z is None:
        pytest.skip(f"tests/golden/ref_{device}.npz not recorded")
    monkeypatch.setenv("BFP_TIE_RULE", device)
    _lib.set_option("force_generic", generic)
    try:
        n = 0
        for c in _golden.quant_cases(z):
            if device == "cpu" and c["kind"] != "q" and (c["N"], c["M"]) != (2, 4):
                continue                       # torch-CPU's tie order is reproduced for 2:4 only
            x = _to_dev(c["x"], c["dt"])
            y = _run(ops, x, c["kind"], c["m"], c["B"], c["N"], c["M"])
            out, odt = _np(y)
            assert odt == c["odt"] and out.shape == c["y"].shape, c["key"]
            assert _golden.mismatches(out, c["y"], odt) == 0, (c["key"], generic)
            n += 1
        assert n > 400
    finally:
        _lib.set_option("force_generic", 0)


@pytest.mark.parametrize("device", ["cuda", "cpu"])
def test_kernel_matches_reference_golden_exponent_and_ties(ops, device, monkeypatch):
    z = _golden.load(device)
    if z is None:
        pytest.skip("not recorded")
    monkeypatch.setenv("BFP_TIE_RULE", device)
    for dt in ("f32", "f16", "bf16"):
        x, _ = _golden.get(z, f"expb_in_{dt}")
        e = ops.get_exponent(_to_dev(x, dt), 1e-8).float().cpu().numpy()
        ref = z[f"expb_out_{dt}"]
        assert ((e == ref) | (np.isnan(e) & np.isnan(ref))).all(), dt
    x, _ = _golden.get(z, "tie_in")
    for (N, M) in ((2, 4), (1, 4), (3, 4)):
        if device == "cpu" and (N, M) != (2, 4):
            continue
        y, _ = _golden.get(z, f"tie_out_{N}:{M}")
        out = ops._structured_N_M_sparsity(_to_dev(x, "f32"), "cuda", N, M).cpu().numpy()
        assert _golden.mismatches(out, y, "f32") == 0, (N, M)


def test_cpu_tie_table_on_device(ops):
    from qsi_b200 import _lib
    import ctypes, re
    buf = (ctypes.c_uint8 * 256)()
    _lib.check(_lib.lib().bfp_debug_cpu_tie_lut(ctypes.addressof(buf)))
    txt = "\n".join(l for l in open(os.path.join(_lib.CSRC, "nm_cpu_tie_lut.inc")).read().splitlines() if not l.startswith("//"))
    assert list(buf) == [int(v, 16) for v in re.findall(r"0x([0-9A-F]{2})", txt)]


# ---------------------------------------------------------------------------------------------------------------
# 2. seeded random inputs vs the oracle: dtype x mant x block x order x N:M, aligned and ragged shapes
# ---------------------------------------------------------------------------------------------------------------
def _inputs(seed, shape, dtype, scale):
    g = torch.Generator().manual_seed(seed)
    t = torch.randn(*shape, generator=g) * scale
    flat = t.view(-1)
    flat[torch.randint(0, flat.numel(), (max(1, flat.numel() // 40),), generator=g)] = 0.0
    flat[torch.randint(0, flat.numel(), (max(1, flat.numel() // 200),), generator=g)] *= 30.0
    return t.to(dtype)


@pytest.mark.parametrize("dt", ["f32", "bf16", "f16"])
@pytest.mark.parametrize("shape", [(256, 1024), (3, 5, 200), (37, 96), (6, 3, 16, 16), (5, 7), (1, 64)])
def test_kernel_matches_oracle_sweep(ops, oracle, dt, shape):
    for si, scale in enumerate((1.0, 0.02, 200.0)):
        x = _inputs(11 * si + len(shape), shape, TORCH_DT[dt], scale)
        xa, _ = _np(x)
        xd = x.cuda()
        for (m, B) in ((3, 16), (5, 32), (7, 64), (7, 16), (15, 64), (3, 128), (7, 48)):
            y, _ = _np(ops._no_sparsity_float_to_bfp(xd, B, m, 1e-8, "determ", "cuda"))
            o, _ = oracle.bfp_quantize(xa, B, m, dt=dt)
            assert _golden.mismatches(y, o, dt) == 0, ("q", scale, m, B)
            for (N, M), kind in itertools.product(((2, 4), (1, 4), (3, 4), (1, 2), (4, 8), (2, 8), (3, 5)), ("sq", "qs")):
                if (m, B) not in ((3, 16), (7, 64), (5, 32)):
                    continue
                y, _ = _np(_run(ops, xd, kind, m, B, N, M))
                o, _ = oracle.float_to_bfp_blocked(xa, m, B, kind, N=N, M=M, tie_rule="cuda", dt=dt)
                assert _golden.mismatches(y, o, dt) == 0, (kind, scale, m, B, N, M)
        for (N, M) in ((2, 4), (1, 4), (3, 4), (1, 2), (4, 8), (2, 8), (8, 16), (3, 5), (1, 1)):
            y, _ = _np(ops._structured_N_M_sparsity(xd, "cuda", N, M))
            o = oracle.nm_sparsify(xa, N, M, tie_rule="cuda", dt=dt)
            assert _golden.mismatches(y, o, dt) == 0, ("s", scale, N, M)


def test_special_values_match_oracle(ops, oracle):
    """zeros, -0.0, denormals, huge values, Inf / NaN blocks (NaN compares equal to NaN)."""
    x = torch.zeros(8, 128)
    x[1] = torch.randn(128) * 1e-40
    x[2] = torch.randn(128) * 3e38
    x[3, 7] = float("inf")
    x[4, 70] = float("nan")
    x[5, ::2] = -0.0
    x[6] = torch.randn(128)
    x[6, 3] = -float("inf")
    x[7] = torch.randn(128) * 1e-9
    for dt in ("f32", "bf16", "f16"):
        xt = x.to(TORCH_DT[dt])
        xa, _ = _np(xt)
        for kind, (m, B) in itertools.product(("q", "sq", "qs", "s"), ((7, 64), (3, 16))):
            y, _ = _np(_run(ops, xt.cuda(), kind, m, B, 2, 4))
            if kind == "s":
                o = oracle.nm_sparsify(xa, 2, 4, dt=dt)
            else:
                o, _ = oracle.float_to_bfp_blocked(xa, m, B, kind, dt=dt)
            assert _golden.mismatches(y, o, dt) == 0, (dt, kind, m, B)


def test_known_answers_on_device(ops):
    q = lambda v, m, B: ops._no_sparsity_float_to_bfp(torch.tensor([v], device="cuda"), B, m, 1e-8, "determ", "cuda")[0].cpu()
    assert torch.equal(q([1.0, 0.99, -1.0, 0.5], 7, 4), torch.tensor([127 / 128, 127 / 128, -127 / 128, 0.5]))
    y = q([-0.3, 0.3, 100.0, -100.0], 3, 4)
    assert torch.equal(y, torch.tensor([-0.0, 0.0, 96.0, -96.0])) and torch.signbit(y).tolist() == [True, False, False, True]
    s = ops._structured_N_M_sparsity(torch.tensor([[-1.0, 2.0, -3.0, 4.0]], device="cuda"), "cuda", 2, 4).cpu()
    assert torch.equal(s, torch.tensor([[0.0, 0.0, -3.0, 4.0]]))
    assert torch.isnan(ops._no_sparsity_float_to_bfp(torch.zeros(1, 8, device="cuda", dtype=torch.float16), 8, 7, 1e-8, "determ", "cuda")).all()


# ---------------------------------------------------------------------------------------------------------------
# 3. live against the reference on torch-CUDA (baseline/_ref travels to the GPU box; skipped when absent)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", ["f32", "bf16", "f16"])
def test_kernel_matches_live_reference_on_cuda(ops, dt):
    ref = load_reference()
    if ref is None:
        pytest.skip("reference sources not present on this box")
    for seed, (shape, scale) in enumerate(itertools.product([(512, 1024), (3, 5, 200), (6, 3, 16, 16)], [0.02, 1.0])):
        x = _inputs(500 + seed, shape, TORCH_DT[dt], scale).cuda()
        for (m, B), (N, M), first in itertools.product(((3, 16), (5, 32), (7, 64)), ((2, 4), (1, 4), (4, 8)), ("s", "q")):
            args = ref_args(ref, mant_bits=m, block_size=B, first=first, N=N, M=M, device="cuda")
            r = ref.float_to_bfp_blocked(x, **args, identifier="w")
            y = _run(ops, x, "sq" if first == "s" else "qs", m, B, N, M)
            assert y.dtype == r.dtype and y.shape == r.shape
            assert _golden.mismatches(_np(y)[0], _np(r)[0], dt) == 0, (shape, scale, m, B, N, M, first)
        r = ref._no_sparsity_float_to_bfp(x, 64, 7, 1e-8, "determ", "cuda")
        assert _golden.mismatches(_np(ops._no_sparsity_float_to_bfp(x, 64, 7, 1e-8, "determ", "cuda"))[0], _np(r)[0], dt) == 0


# ---------------------------------------------------------------------------------------------------------------
# 4. stochastic rounding: bit-exact given the kernel's Philox stream, unbiased, fp32 output for half inputs
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", ["f32", "bf16", "f16"])
@pytest.mark.parametrize("shape", [(64, 512), (9, 100)])
def test_stochastic_rounding_bit_exact_given_philox(ops, oracle, dt, shape):
    from qsi_b200 import _lib
    x = _inputs(77, shape, TORCH_DT[dt], 1.0)
    xa, _ = _np(x)
    xd = x.cuda()
    for (m, B), (seed, off) in itertools.product(((3, 16), (7, 64)), ((1234, 0), (2 ** 40 + 17, 5))):
        u = oracle.philox_uniforms(x.numel(), seed, off).reshape(x.shape)
        for kind, order in (("q", _lib.ORDER_QUANT_ONLY), ("sq", _lib.ORDER_SPARSIFY_QUANT), ("qs", _lib.ORDER_QUANT_SPARSIFY)):
            y = ops._fused(xd, order, block_size=B, mant_bits=m, epsilon=1e-8, rounding_mode="stoc", N=2, M=4, philox=(seed, off))
            assert y.dtype == torch.float32                    # type promotion of the reference (bfp_ops.py:22-23)
            o, odt = oracle.float_to_bfp_blocked(xa, m, B, kind, rounding_mode="stoc", rand_u=u, dt=dt)
            assert odt == "f32" and _golden.mismatches(y.cpu().numpy(), o, "f32") == 0, (kind, m, B, seed)


def test_stochastic_rounding_statistics(ops):
    torch.manual_seed(3)
    n = 1 << 18
    x = torch.full((n, 8), 0.3, device="cuda")
    x[:, 0] = 1.0
    y = ops._no_sparsity_float_to_bfp(x, 8, 3, 1e-8, "stoc", "cuda")
    vals = torch.unique(y[:, 1:]).cpu().tolist()
    assert set(vals) <= {0.25, 0.375}
    assert abs(y[:, 1:].double().mean().item() - 0.3) < 5 * 0.0625 / np.sqrt(7 * n)
    assert (y[:, 0] == 0.875).all()
    y2 = ops._no_sparsity_float_to_bfp(x, 8, 3, 1e-8, "stoc", "cuda")
    assert not torch.equal(y, y2)                              # successive calls advance the Philox offset
    # error distribution: for uniform inputs the rounding error is ~U(-delta, delta) triangular with mean 0
    w = torch.rand(1 << 20, device="cuda") * 0.9
    w[::64] = 1.0
    yq = ops._no_sparsity_float_to_bfp(w.view(-1, 64), 64, 7, 1e-8, "stoc", "cuda").view(-1)
    err = (yq - w)[w < 0.95].double()
    delta = 2.0 ** -7
    assert abs(err.mean().item()) < 5 * delta / np.sqrt(err.numel())
    assert err.abs().max().item() < delta


# ---------------------------------------------------------------------------------------------------------------
# 5. host-buffer entry point and full-size properties
# ---------------------------------------------------------------------------------------------------------------
def test_host_buffer_path_equals_device_path(ops):
    from qsi_b200 import _lib
    x = _inputs(5, (1000, 1024), torch.float32, 0.02)
    _lib.set_option("host_chunk_bytes", 1 << 20)             # force several pipelined chunks
    try:
        for kind, rounding in itertools.product(("q", "sq", "qs", "s"), ("determ",)):
            y_host = _run(ops, x.pin_memory(), kind, 7, 64, 2, 4)
            y_page = _run(ops, x, kind, 7, 64, 2, 4)
            assert not y_host.is_cuda
            monkey = os.environ.get("BFP_TIE_RULE")
            os.environ["BFP_TIE_RULE"] = "cpu"               # CPU tensors default to the torch-CPU 2:4 tie order
            try:
                y_dev = _run(ops, x.cuda(), kind, 7, 64, 2, 4).cpu()
            finally:
                if monkey is None:
                    del os.environ["BFP_TIE_RULE"]
                else:
                    os.environ["BFP_TIE_RULE"] = monkey
            assert _golden.mismatches(y_host.numpy(), y_dev.numpy(), "f32") == 0, kind
            assert _golden.mismatches(y_page.numpy(), y_dev.numpy(), "f32") == 0, kind
        # stochastic: chunking must not change the Philox stream
        a = ops._fused(x.pin_memory(), _lib.ORDER_QUANT_ONLY, block_size=64, mant_bits=7, epsilon=1e-8, rounding_mode="stoc", philox=(9, 1))
        b = ops._fused(x.cuda(), _lib.ORDER_QUANT_ONLY, block_size=64, mant_bits=7, epsilon=1e-8, rounding_mode="stoc", philox=(9, 1)).cpu()
        assert torch.equal(a, b)
    finally:
        _lib.set_option("host_chunk_bytes", 8 << 20)
        _lib.lib().bfp_host_staging_release()


@pytest.mark.parametrize("shape", [(4096, 4096), (4096, 11008)])
def test_full_size_properties(ops, oracle, shape):
    """BASELINE config 2 shapes: oracle on a row sample, chunk invariance, N:M structure, idempotence of the mask."""
    g = torch.Generator(device="cuda").manual_seed(0)
    w = torch.randn(*shape, device="cuda", generator=g) * 0.02
    rows = torch.randint(0, shape[0], (48,), generator=torch.Generator().manual_seed(1)).tolist()
    for (m, B), kind in itertools.product(((3, 16), (5, 32), (7, 64)), ("sq", "qs")):
        y = _run(ops, w, kind, m, B, 2, 4)
        # (a) bit-exact against the oracle on sampled rows
        xa = w[rows].cpu().numpy()
        o, _ = oracle.float_to_bfp_blocked(xa, m, B, kind)
        assert _golden.mismatches(y[rows].cpu().numpy(), o, "f32") == 0, (kind, m, B)
        # (b) row-chunk invariance
        parts = torch.cat([_run(ops, c, kind, m, B, 2, 4) for c in w.chunk(8, dim=0)], dim=0)
        assert torch.equal(parts.view(torch.int32), y.view(torch.int32))
        # (c) at least 2 zeros in every group of 4; kept values lie on the block grid
        assert int((y.view(-1, 4) == 0).sum(dim=1).min()) >= 2
        # (d) the mask is idempotent
        assert torch.equal(ops._structured_N_M_sparsity(y, "cuda", 2, 4), y)
    # bf16 / fp16 at full size against the oracle on the row sample
    for dt in ("bf16", "f16"):
        wt = w.to(TORCH_DT[dt])
        y = _run(ops, wt, "sq", 7, 64, 2, 4)
        o, _ = oracle.float_to_bfp_blocked(_np(wt[rows])[0], 7, 64, "sq", dt=dt)
        assert _golden.mismatches(_np(y[rows])[0], o, dt) == 0


def test_api_conventions_on_device(ops):
    x = torch.randn(4, 6, 64, device="cuda")
    xt = x.transpose(0, 1)                                     # non-contiguous input
    args = ops.unpack_bfp_args(dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", mant_bits=7,
                                    block_size=64, w_sparsity=True, N=2, M=4, first="s", sparsity_mode="structured"))
    y = ops.float_to_bfp_blocked(xt, **args, identifier="w")
    assert y.shape == xt.shape and y.data_ptr() != xt.data_ptr()
    assert torch.equal(y, ops.float_to_bfp_blocked(xt.contiguous(), **args, identifier="w"))
    y_in = ops.float_to_bfp_blocked(x, **args, identifier="in")      # activations are not sparsified
    assert torch.equal(y_in, ops._no_sparsity_float_to_bfp(x, 64, 7, 1e-8, "determ", "cuda"))
    assert ops.float_to_bfp_blocked(torch.empty(0, 64, device="cuda"), **args, identifier="w").shape == (0, 64)
    with pytest.raises(AssertionError):                       # bfp_ops.py:130: only 'bfp' / 'fp32' / 'int' are accepted
        ops.float_to_bfp_blocked(x, **dict(args, sparsity_num_format="fp8"), identifier="w")
    assert ops.float_to_bfp_blocked(x, **dict(args, sparsity_num_format="int"), identifier="w").dtype == torch.float32


def test_bfp_linear_module_matches_oracle(ops, oracle):
    """BFPLinear / F_matmul_bfp forward + STE backward (bfp_ops.py:160-192, 270-287)."""
    torch.manual_seed(0)
    kw = dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=7,
              block_size=64, w_sparsity=True, N=2, M=4, first="s", sparsity_mode="structured", device="cuda")
    lin = ops.BFPLinear(256, 96, bias=True, **dict(kw)).cuda()
    x = torch.randn(3, 20, 256, device="cuda", requires_grad=True)
    y = lin(x)
    xq, _ = oracle.bfp_quantize(x.detach().cpu().numpy(), 64, 7)
    wq, _ = oracle.float_to_bfp_blocked(lin.weight.detach().cpu().numpy(), 7, 64, "sq")
    ref = oracle.linear(xq, wq, lin.bias.detach().cpu().numpy())
    rel = np.linalg.norm(y.detach().cpu().numpy() - ref) / np.linalg.norm(ref)
    assert rel <= 1e-5, rel                                      # north_star GEMM tolerance
    y.sum().backward()
    # STE: dL/dx = Q_grad(1) @ W_q ; the all-ones output gradient quantises to 127/128 (block max 1.0 saturates to
    # vmax = 1 - 2^-7, bfp_ops.py:39,44), identifier='grad' is not sparsified
    gx = (0.9921875 * torch.from_numpy(wq).cuda().sum(dim=0)).expand_as(x)
    assert torch.allclose(x.grad, gx, rtol=1e-5, atol=1e-6)
    assert lin.weight.grad is not None and lin.weight.grad.shape == lin.weight.shape
    mm = ops.F_matmul_bfp(**dict(kw))
    a = torch.randn(2, 4, 16, 64, device="cuda")
    b = torch.randn(2, 4, 64, 32, device="cuda")
    out = mm(a, b)
    aq, _ = oracle.bfp_quantize(a.cpu().numpy(), 64, 7)
    bq, _ = oracle.float_to_bfp_blocked(b.transpose(-1, -2).contiguous().cpu().numpy(), 7, 64, "sq")
    ref = np.matmul(aq.astype(np.float64), np.swapaxes(bq, -1, -2).astype(np.float64))
    assert np.linalg.norm(out.cpu().numpy() - ref) / np.linalg.norm(ref) <= 1e-5


def test_pdl_launches_keep_stream_order():
    """The streaming kernels are launched with programmatic stream serialization (their CTAs may be scheduled while the
    previous kernel drains).  A chain of dependent calls -- each reads what the previous one wrote, in place included, with
    torch kernels in between -- must give exactly the result of the same chain with PDL off and a sync after every step."""
    from qsi_b200 import _lib, bfp_ops
    g = torch.Generator(device="cuda").manual_seed(5)
    x0 = torch.randn(2048, 4096, device="cuda", generator=g)
    args = bfp_ops.unpack_bfp_args(dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, block_size=64,
                                        w_sparsity=True, N=2, M=4, sparsity_mode="structured", device="cuda"))

    def chain(sync):
        y = x0.clone()
        for step in range(12):
            m = 7 - (step % 5)
            y = bfp_ops.float_to_bfp_blocked(y * 1.25 + 0.01, **dict(args, mant_bits=m, first="s" if step % 2 else "q"), identifier="w")
            if sync:
                torch.cuda.synchronize()
            p = bfp_ops.pack_bfp_bf16(y, identifier="in", **dict(args, mant_bits=m))
            y = p.float() + y
            if sync:
                torch.cuda.synchronize()
        return y

    _lib.set_option("pdl", 0)
    try:
        ref = chain(True)
    finally:
        _lib.set_option("pdl", 1)
    for _ in range(3):
        assert torch.equal(chain(False), ref)


@pytest.mark.parametrize("dtn", ["bf16", "f16"])
def test_half_precision_exponent_table_matches_torch_log2(dtn):
    """The kernels take ceil(log2(s)) for fp16 / bf16 blocks from a table of step positions (csrc/bfp_common.cuh).  Every
    (exponent, mantissa) pair of the dtype is checked against torch's own evaluation on the GPU: s.log2().ceil() (bfp_ops.py:33)."""
    import ctypes
    from qsi_b200 import _lib
    dt, code, mb = (torch.bfloat16, _lib.DT_BF16, 7) if dtn == "bf16" else (torch.float16, _lib.DT_F16, 10)
    tab = (ctypes.c_uint16 * 256)()
    _lib.check(_lib.lib().bfp_debug_exp_table(code, tab))
    tab = np.array(tab, dtype=np.int64)
    krange = range(-100, 127) if dtn == "bf16" else range(-24, 16)
    assert all(tab[k + 128] != 0 for k in krange), "fast-path exponents must be tabulated"
    ks = torch.tensor(list(krange), dtype=torch.float64)
    f = torch.arange(1 << mb, dtype=torch.float64)
    s = (torch.exp2(ks)[:, None] * (1.0 + f[None, :] / (1 << mb))).to(dt).cuda()        # exact: every value is representable
    e_torch = s.log2().ceil().float().cpu().numpy()
    step = tab[[k + 128 for k in krange]] - 1
    e_table = np.array(list(krange))[:, None] + (np.arange(1 << mb)[None, :] >= step[:, None])
    # fp16 subnormals (k < -14) carry fewer mantissa bits: only the f that survive the conversion exactly are real inputs
    exact = (s.double().cpu() == torch.exp2(ks)[:, None] * (1.0 + f[None, :] / (1 << mb))).numpy()
    assert exact[[i for i, k in enumerate(krange) if k >= (-14 if dtn == "f16" else -126)]].all()
    assert np.array_equal(e_table[exact], e_torch[exact])


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape,B,sparse", [((16, 3, 224, 224), 64, False), ((768, 3, 16, 16), 64, True), ((515, 4100), 64, True), ((33, 40), 32, True),
                                            ((7, 1048), 128, False)])
def test_padded_row_mode_equals_generic_path(dt, shape, B, sparse):
    """Rows that are vector-aligned but not block-aligned (the ViT patch-embedding input, conv weights, odd widths) run on the
    streaming kernel with zero-padded slots; the result must equal the gather-style generic kernel bit for bit, for the
    fake-quant output (nearest and stochastic) and for the bf16 operand pack."""
    from qsi_b200 import _lib, bfp_ops
    V = 4 if dt == torch.float32 else 8
    if shape[-1] % V:
        pytest.skip("row length not vector-aligned for this dtype: generic path either way")
    g = torch.Generator().manual_seed(sum(shape) + B)
    x = (torch.randn(*shape, generator=g) * 0.3).to(dt).cuda()
    x.view(-1)[::97] *= 40.0
    for first, rmode in (("s", "determ"), ("q", "determ"), ("s", "stoc")):
        a = bfp_ops.unpack_bfp_args(dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode=rmode, epsilon=1e-8, mant_bits=5, block_size=B,
                                         w_sparsity=sparse, N=2, M=4, first=first, sparsity_mode="structured", device="cuda"))
        outs, packs = [], []
        for generic in (0, 1):
            _lib.set_option("force_generic", generic)
            try:
                outs.append(bfp_ops._fused(x, bfp_ops._order_for(a, "w"), block_size=B, mant_bits=5, epsilon=1e-8, rounding_mode=rmode, N=2, M=4,
                                           philox=(1234, 5)))
                if rmode == "determ":
                    packs.append(bfp_ops.pack_bfp_bf16(x, identifier="w", **a))
            finally:
                _lib.set_option("force_generic", 0)
        assert outs[0].dtype == outs[1].dtype and torch.equal(outs[0].view(torch.int16 if outs[0].element_size() == 2 else torch.int32),
                                                             outs[1].view(torch.int16 if outs[1].element_size() == 2 else torch.int32)), (first, rmode)
        if packs:
            assert torch.equal(packs[0].view(torch.int16), packs[1].view(torch.int16)), (first, "pack")


# ---------------------------------------------------------------------------------------------------------------
# Row f3: the optimiser wrappers (bfp_optim.py:8-64, bfp_optim_lstm.py:12-95) -- sgd_update selects the wide mantissa
# ---------------------------------------------------------------------------------------------------------------
def _optim_args(**kw):
    base = dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=7, weight_mant_bits=15,
                block_size=64, device="cuda")
    base.update(kw)
    return base


def test_bfp_optim_wrapper_keeps_wide_and_narrow_weights(oracle):
    import qsi_b200  # noqa: F401
    from qsi_b200 import bfp_optim
    torch.manual_seed(0)
    p = torch.nn.Parameter(torch.randn(32, 256, device="cuda") * 0.1)
    p0 = p.detach().cpu().numpy().copy()
    opt = bfp_optim.get_bfp_optim(torch.optim.SGD, "SGD")([p], lr=0.1, **_optim_args())
    assert type(opt).__name__ == "BFPSGD" and bfp_optim.get_bfp_optim(torch.optim.SGD, "SGD") is type(opt)
    g1 = torch.randn_like(p)
    p.grad = g1.clone()
    opt.step()
    # first step: the weight is constrained to the wide format, updated in fp32, then stored wide (shadow) and narrow (p)
    def plain_sgd(start, g):                                  # the wrapped optimiser's own fp32 update from `start`
        r = torch.nn.Parameter(torch.from_numpy(start).cuda())
        r.grad = g.clone()
        torch.optim.SGD([r], lr=0.1).step()
        return r.detach().cpu().numpy()
    w0, _ = oracle.bfp_quantize(p0, 64, 15)
    upd = plain_sgd(w0, g1)
    wide, _ = oracle.bfp_quantize(upd, 64, 15)
    narrow, _ = oracle.bfp_quantize(upd, 64, 7)
    assert np.array_equal(opt.state[p]["shadow_p"].cpu().numpy().view(np.uint32), wide.view(np.uint32))
    assert np.array_equal(p.detach().cpu().numpy().view(np.uint32), narrow.view(np.uint32))
    # second step starts from the wide copy, not from the narrow weights the model sees
    g2 = torch.randn_like(p)
    p.grad = g2.clone()
    opt.step()
    upd2 = plain_sgd(wide, g2)
    wide2, _ = oracle.bfp_quantize(upd2, 64, 15)
    narrow2, _ = oracle.bfp_quantize(upd2, 64, 7)
    assert np.array_equal(opt.state[p]["shadow_p"].cpu().numpy().view(np.uint32), wide2.view(np.uint32))
    assert np.array_equal(p.detach().cpu().numpy().view(np.uint32), narrow2.view(np.uint32))
    # num_format fp32: the wrapped optimiser, untouched
    q = torch.nn.Parameter(torch.ones(4, 64, device="cuda"))
    o2 = bfp_optim.get_bfp_optim(torch.optim.SGD, "SGD")([q], lr=0.5, num_format="fp32")
    q.grad = torch.ones_like(q)
    o2.step()
    assert torch.equal(q.detach(), torch.full_like(q, 0.5)) and "shadow_p" not in o2.state[q]


def test_bfp_adam_constrains_the_update_to_the_wide_format(oracle):
    import qsi_b200  # noqa: F401
    from qsi_b200 import bfp_optim
    torch.manual_seed(1)
    p = torch.nn.Parameter(torch.randn(16, 128, device="cuda"))
    r = torch.nn.Parameter(p.detach().clone())
    opt = bfp_optim.BFPAdam([p], lr=1e-2, bfp_args=_optim_args())
    ref = torch.optim.Adam([r], lr=1e-2)
    for _ in range(3):
        g = torch.randn_like(p)
        p.grad, r.grad = g.clone(), g.clone()
        before = p.detach().clone()
        opt.step()
        # one plain Adam step from the same point, then the wide-mantissa constraint
        r.data.copy_(before)
        ref.step()
        want, _ = oracle.bfp_quantize(r.detach().cpu().numpy(), 64, 15)
        got = p.detach().cpu().numpy()
        assert np.abs(got - want).max() <= 2.0 ** -14 * np.abs(want).max()      # same grid; the two Adam formulations differ by an fp32 ulp
        q, _ = oracle.bfp_quantize(got, 64, 15)
        assert np.array_equal(q.view(np.uint32), got.view(np.uint32))           # the result is on the wide BFP grid
