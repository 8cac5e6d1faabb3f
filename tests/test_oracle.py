"""CPU suite: pins the oracle (oracle/bfp_oracle.c) against the reference's recorded outputs (tests/golden/) and,
when the reference sources are present (build container), live against the reference itself."""
import itertools

import numpy as np
import pytest
import torch

import _golden
from _refload import load_reference, ref_args


def _tie_for(device):
    return "cuda" if device == "cuda" else "cpu"


@pytest.mark.parametrize("device", ["cpu", "cuda"])
def test_oracle_matches_golden_quant_and_sparsify(oracle, device):
    z = _golden.load(device)
    if z is None:
        pytest.skip(f"tests/golden/ref_{device}.npz not recorded yet")
    n = 0
    for c in _golden.quant_cases(z):
        if c["kind"] == "s":
            out = oracle.nm_sparsify(c["x"], c["N"], c["M"], tie_rule=_tie_for(device), dt=c["dt"])
            odt = c["dt"]
        else:
            out, odt = oracle.float_to_bfp_blocked(c["x"], c["m"], c["B"], c["kind"], N=c["N"], M=c["M"],
                                                   tie_rule=_tie_for(device), dt=c["dt"])
        assert odt == c["odt"], c["key"]
        assert out.shape == c["y"].shape, c["key"]
        assert _golden.mismatches(out, c["y"], odt) == 0, c["key"]
        n += 1
    assert n > 500


@pytest.mark.parametrize("device", ["cpu", "cuda"])
def test_oracle_matches_golden_exponent_boundaries(oracle, device):
    z = _golden.load(device)
    if z is None:
        pytest.skip("not recorded")
    for dt in ("f32", "f16", "bf16"):
        x, _ = _golden.get(z, f"expb_in_{dt}")
        e = oracle.bfp_exponent(x, 1, dt=dt)
        ref = z[f"expb_out_{dt}"]
        same = (e == ref) | (np.isnan(e) & np.isnan(ref))
        assert same.all(), (dt, np.argwhere(~same)[:5])


@pytest.mark.parametrize("device", ["cpu", "cuda"])
def test_oracle_matches_golden_tie_table(oracle, device):
    z = _golden.load(device)
    if z is None:
        pytest.skip("not recorded")
    x, _ = _golden.get(z, "tie_in")
    for (N, M) in ((2, 4), (1, 4), (3, 4)):
        y, _ = _golden.get(z, f"tie_out_{N}:{M}")
        out = oracle.nm_sparsify(x, N, M, tie_rule=_tie_for(device))
        assert _golden.mismatches(out, y, "f32") == 0


@pytest.mark.parametrize("device", ["cpu", "cuda"])
def test_oracle_matches_golden_stochastic_with_recorded_uniforms(oracle, device):
    z = _golden.load(device)
    if z is None:
        pytest.skip("not recorded")
    for dt in ("f32", "f16", "bf16"):
        for (m, B) in ((3, 16), (7, 64)):
            x, _ = _golden.get(z, f"stoc_in_{dt}_m{m}_b{B}")
            u = z[f"stoc_u_{dt}_m{m}_b{B}"]
            y, ydt = _golden.get(z, f"stoc_out_{dt}_m{m}_b{B}")
            assert ydt == "f32"                      # fp32 output for half inputs (type promotion, bfp_ops.py:22-23)
            out, odt = oracle.bfp_quantize(x, B, m, rounding_mode="stoc", rand_u=u, dt=dt)
            assert odt == "f32"
            assert _golden.mismatches(out, y, "f32") == 0, (dt, m, B)


@pytest.mark.parametrize("device", ["cpu", "cuda"])
def test_oracle_linear_matches_golden(oracle, device):
    z = _golden.load(device)
    if z is None:
        pytest.skip("not recorded")
    for first, order in (("s", "sq"), ("q", "qs")):
        x, _ = _golden.get(z, f"lin_{first}_x")
        w, _ = _golden.get(z, f"lin_{first}_w")
        b, _ = _golden.get(z, f"lin_{first}_b")
        y, _ = _golden.get(z, f"lin_{first}_y")
        xq, _ = oracle.bfp_quantize(x, 64, 7)
        wq, _ = oracle.float_to_bfp_blocked(w, 7, 64, order, tie_rule=_tie_for(device))
        out = oracle.linear(xq, wq, b)
        rel = np.linalg.norm(out - y) / np.linalg.norm(y)
        assert rel <= 1e-5, rel                     # north_star GEMM tolerance


def test_known_answers(oracle):
    """SURVEY.md appendix A.4/A.5/A.7 hand-checked values."""
    q = lambda x, m, B: oracle.bfp_quantize(np.array([x], np.float32), B, m)[0][0]
    np.testing.assert_array_equal(q([1.0, 0.99, -1.0, 0.5], 7, 4), np.float32([127 / 128, 127 / 128, -127 / 128, 0.5]))
    y = q([1.0, 0.0625, 0.1875, 0.3125, 0.4375, -0.0625, -0.1875, 0.99], 3, 8)
    np.testing.assert_array_equal(y, np.float32([0.875, 0.0, 0.25, 0.25, 0.5, -0.0, -0.25, 0.875]))
    assert np.signbit(y[5]) and not np.signbit(y[1])
    np.testing.assert_array_equal(q([128, .5, 1.5, 2.5, -.5, -1.5], 8, 6), np.float32([127.5, .5, 1.5, 2.5, -.5, -1.5]))
    y = q([-0.3, 0.3, 100.0, -100.0], 3, 4)
    np.testing.assert_array_equal(y, np.float32([-0.0, 0.0, 96.0, -96.0]))
    assert list(np.signbit(y)) == [True, False, False, True]
    s = oracle.nm_sparsify(np.array([[-1, 2, -3, 4]], np.float32), 2, 4)
    np.testing.assert_array_equal(s, np.float32([[0, 0, -3, 4]]))
    e = lambda v: oracle.bfp_exponent(np.array([[v]], np.float32), 1)[0, 0]
    assert e(0.0) == -26 and e(2.0 ** -28) == -26 and e(2.0 ** -27) == -25
    assert e(1.0) == 0 and e(np.nextafter(np.float32(1), np.float32(2))) == 1
    assert e(2.0 ** -3) == -2 and e(2.0 ** -10) == -9 and e(32.0) == 5
    assert e(np.float32(32.0) + np.float32(2.0 ** -18)) == 5           # 32 + 1 ulp: log2 rounds back to 5
    # fp16 all-zero block -> NaN; bf16 -> 0
    assert np.isnan(oracle.bfp_quantize(np.zeros((1, 8), np.float16), 8, 7, dt="f16")[0]).all()
    assert (oracle.bfp_quantize(np.zeros((1, 8), np.uint16), 8, 7, dt="bf16")[0] == 0).all()


def test_stochastic_rounding_statistics(oracle):
    """constant 0.3 with block max 1.0, m=3 -> support {0.25, 0.375}, mean 0.3; max stays 0.875."""
    n = 1 << 16
    x = np.full((n, 8), 0.3, np.float32)
    x[:, 0] = 1.0
    u = oracle.philox_uniforms(x.size, seed=123, offset=0)
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 2e-3
    y, _ = oracle.bfp_quantize(x, 8, 3, rounding_mode="stoc", rand_u=u)
    assert set(np.unique(y[:, 1:])) <= {np.float32(0.25), np.float32(0.375)}
    assert abs(y[:, 1:].mean() - 0.3) < 5 * 0.0625 / np.sqrt(y[:, 1:].size)
    assert (y[:, 0] == 0.875).all()


def test_philox_known_answer(oracle):
    """Philox4x32-10 KAT from the Random123 distribution (kat_vectors): counter=key=0 -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8."""
    import ctypes
    L = oracle.lib()
    L.oracle_philox_word.argtypes = [ctypes.c_uint64] * 3
    L.oracle_philox_word.restype = ctypes.c_uint32
    kat = [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    # element i uses word (i & 3) of the block for counter (i>>2, 0, offset, 0), key = seed
    assert [L.oracle_philox_word(0, 0, i) for i in range(4)] == kat
    # the uniform is that word truncated to 24 significant bits, times 2^-32 (the kernel's cvt.rz.f32.u32 and an exact scale)
    for i, w in enumerate(kat):
        drop = max(0, w.bit_length() - 24)
        assert L.oracle_philox_uniform(0, 0, i) == np.float32((w >> drop << drop) * 2.0 ** -32)
    assert L.oracle_philox_uniform(0, 0, 1) == np.float32((0xe169c58d >> 8) * 2.0 ** -24)       # words >= 2^31: (w >> 8) * 2^-24


# ---------------------------------------------------------------------------------------------------------------
# live differential tests against the reference (only where its sources exist: the build container)
# ---------------------------------------------------------------------------------------------------------------
ref = load_reference()
needs_ref = pytest.mark.skipif(ref is None, reason="reference sources not present")


def _rand_input(seed, shape, dtype, scale):
    g = torch.Generator().manual_seed(seed)
    t = torch.randn(*shape, generator=g) * scale
    flat = t.view(-1)
    idx = torch.randint(0, flat.numel(), (max(1, flat.numel() // 50),), generator=g)
    flat[idx] = 0.0
    return t.to(dtype)


@needs_ref
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_live_quantize_vs_reference(oracle, dtype):
    for seed, (shape, scale) in enumerate(itertools.product([(33, 192), (2, 3, 100), (5, 7)], [1.0, 0.02, 1e-3, 250.0])):
        t = _rand_input(seed, shape, dtype, scale)
        a, dt = oracle.from_torch(t)
        for m, B in ((3, 16), (5, 32), (7, 64), (15, 64), (7, 48)):
            r = ref._no_sparsity_float_to_bfp(t, B, m, 1e-8, "determ", "cpu")
            o, odt = oracle.bfp_quantize(a, B, m, dt=dt)
            assert _golden.mismatches(o, oracle.from_torch(r)[0], odt) == 0, (shape, scale, m, B)


@needs_ref
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_live_full_path_vs_reference_cpu_tie_rule(oracle, dtype):
    for seed, shape in enumerate([(16, 256), (3, 5, 200), (4, 10)]):
        t = _rand_input(100 + seed, shape, dtype, 0.05)
        a, dt = oracle.from_torch(t)
        for (m, B), (N, M), first in itertools.product(((3, 16), (7, 64)), ((2, 4), (1, 4), (4, 8)), ("s", "q")):
            args = ref_args(ref, mant_bits=m, block_size=B, first=first, N=N, M=M)
            r = ref.float_to_bfp_blocked(t, **args, identifier="w")
            o, odt = oracle.float_to_bfp_blocked(a, m, B, "sq" if first == "s" else "qs", N=N, M=M, tie_rule="cpu", dt=dt)
            assert _golden.mismatches(o, oracle.from_torch(r)[0], odt) == 0, (shape, m, B, N, M, first)


@needs_ref
def test_live_exponent_sweep_vs_reference(oracle):
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(1 << 20, 1, generator=g).abs() * torch.exp2(torch.randint(-40, 40, (1 << 20, 1), generator=g).float()))
    e_ref = ref.get_exponent(x, 1e-8).numpy()
    e = oracle.bfp_exponent(x.numpy(), 1)
    assert (e == e_ref).all()
    ks = torch.arange(-120, 121).float()
    x = torch.exp2(ks).reshape(-1, 1)
    for _ in range(40):
        assert (oracle.bfp_exponent(x.numpy(), 1) == ref.get_exponent(x, 1e-8).numpy()).all()
        x = torch.nextafter(x, torch.tensor(float("inf")))


@pytest.mark.parametrize("device", ["cpu", "cuda"])
def test_oracle_matches_golden_int_format_and_unstructured(oracle, device):
    z = _golden.load(device)
    if z is None or not any(k.startswith("int_in_") for k in z.files):
        pytest.skip("not recorded")
    n = 0
    for k in z.files:
        if k.startswith("int_out_"):
            name, dtn, ident, bits = k[len("int_out_"):].rsplit("__", 1)[0].rsplit("_", 3)
            x, dt = _golden.get(z, f"int_in_{name}_{dtn}")
            out = oracle.int_quantize(x, int(bits[1:]), ident == "w", dt=dt)
            assert k.endswith("__f32") and _golden.mismatches(out, z[k], "f32") == 0, k       # fp32 output for every dtype
            n += 1
        elif k.startswith("un_out_"):
            name, dtn, frac = k[len("un_out_"):].rsplit("__", 1)[0].rsplit("_", 2)
            x, dt = _golden.get(z, f"un_in_{name}_{dtn}")
            out = oracle.unstructured_sparsify(x, float(frac), dt=dt)
            assert _golden.mismatches(out, z[k], dt) == 0, k
            n += 1
    assert n >= 36
