"""CPU suite for the MX (OCP Microscaling) oracle, oracle/mx_oracle.py.  The arithmetic of the reference's mx_layers.py lives in
microsoft/microxcaling, which the reference neither vendors nor pins (parity unpinned); what CAN be checked is checked here:
  * the element-format parameters against the reference's own formats.py (when /root/reference is present) and against the
    OCP MX v1.0 tables restated below;
  * the element rounding against an independent implementation -- torch's float8 / bfloat16 casts agree with 'nearest' everywhere
    except at exact ties (half away from zero vs half to even), which are checked by hand-derived known answers;
  * hand-derived block examples and the size-independent properties of the format (scale invariance by powers of two, idempotence,
    every output on the format's grid, |error| <= half a step).
"""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import mx_oracle as M

REF = "/root/reference/src/transformers/bfp"
FORMATS = ["int8", "int4", "fp8_e5m2", "fp8_e4m3", "fp6_e3m2", "fp6_e2m3", "fp4_e2m1"]


def _load_ref(name):
    path = os.path.join(REF, name + ".py")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("ref_" + name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _data(seed, shape, scale=1.0):
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal(shape) * scale).astype(np.float32)
    x.flat[:: 97] *= 20.0
    return x


def test_format_table_matches_reference_formats_py():
    ref = _load_ref("formats")
    if ref is None:
        pytest.skip("reference sources not present")
    for name, (ebits, mbits, emax, max_norm) in M.FORMATS.items():
        r = ref._get_format_params(name)
        assert (r[0], r[1], r[2], float(r[3])) == (ebits, mbits, emax, float(max_norm)), name
        assert ref.ElemFormat.from_str(name).value == M.FORMAT_IDS[name], name


def test_format_table_matches_ocp_mx_v1_tables():
    # OCP MX v1.0, tables 1-3: max normal and the exponent range of every element type
    assert M.FORMATS["fp8_e4m3"][3] == 448.0 and M.FORMATS["fp8_e5m2"][3] == 57344.0
    assert M.FORMATS["fp6_e3m2"][3] == 28.0 and M.FORMATS["fp6_e2m3"][3] == 7.5 and M.FORMATS["fp4_e2m1"][3] == 6.0
    assert M.FORMATS["int8"][3] == 127.0 / 64.0                              # 1 + 63/64: MXINT8 is a fixed-point number with 6 fraction bits


@pytest.mark.parametrize("fmt,tdt", [("fp8_e4m3", torch.float8_e4m3fn), ("fp8_e5m2", torch.float8_e5m2)])
def test_element_rounding_agrees_with_torch_float8_casts_away_from_ties(fmt, tdt):
    ebits, mbits, emax, max_norm = M.FORMATS[fmt]
    for seed, scale in ((0, 1.0), (1, 30.0), (2, 1e-3)):
        x = _data(seed, (64, 128), scale)
        x = np.clip(x, -max_norm, max_norm)
        q = M.quantize_elemwise_core(x, mbits, ebits, max_norm, saturate_normals=True)
        t = torch.from_numpy(x).to(tdt).float().numpy()
        assert np.array_equal(q, t)


def test_nearest_is_half_away_from_zero_known_answers():
    # e4m3 (mbits 5 = sign + implicit + 3): the grid around 1 is 1.0, 1.125, 1.25; around 16: 16, 18, 20
    q = M.quantize_elemwise_core(np.array([1.0625, -1.0625, 1.1875, 17.0, -19.0, 0.0009765625], np.float32), 5, 4, 448.0, saturate_normals=True)
    assert q.tolist() == [1.125, -1.125, 1.25, 18.0, -20.0, 0.001953125]       # ties go away from zero; 2^-10 is half the subnormal step 2^-9
    # bfloat16: 1 + 2^-8 is a tie between 1.0 and 1 + 2^-7
    b = M.quantize_bfloat(np.array([1.00390625, -1.00390625, 1.0 + 2.0 ** -9], np.float32), 16)
    assert b.tolist() == [1.0078125, -1.0078125, 1.0]
    # zeros come out +0.0, a negative value that rounds to zero keeps its sign (sign(x) * floor(|x| + 0.5))
    z = M.quantize_mx(np.array([[-0.0, 0.0, -1e-3, 4.0] + [0.0] * 28], np.float32), "fp4_e2m1", 32)
    assert np.signbit(z[0, :4]).tolist() == [False, False, True, False] and z[0, 3] == 4.0


def test_bfloat16_agrees_with_torch_cast_away_from_ties():
    x = _data(3, (128, 128))
    x.view(np.uint32)[:] |= 1                                                 # odd low bit: never a tie
    assert np.array_equal(M.quantize_bfloat(x, 16), torch.from_numpy(x).bfloat16().float().numpy())
    big = np.array([3.4e38, -3.4e38], np.float32)                             # beyond the bf16 maximum: Inf (saturate_normals False)
    assert np.isinf(M.quantize_bfloat(big, 16)).all()


def test_mx_block_known_answers():
    # one block of 32: max 6.5 -> floor(log2) = 2
    x = np.zeros((1, 32), np.float32)
    x[0, :6] = [6.5, 1.0, -0.75, 0.3, 0.1, -3.2]
    # fp4_e2m1 (emax 2): scale 2^0; grid {0, .5, 1, 1.5, 2, 3, 4, 6}
    assert M.quantize_mx(x, "fp4_e2m1", 32)[0, :6].tolist() == [6.0, 1.0, -1.0, 0.5, 0.0, -3.0]
    # int8 (emax 0): scale 2^2, step 2^2 / 64 = 0.0625
    assert M.quantize_mx(x, "int8", 32)[0, :6].tolist() == [6.5, 1.0, -0.75, 0.3125, 0.125, -3.1875]
    # fp8_e4m3 (emax 8): scale 2^-6; 6.5 * 64 = 416 is on the grid (step 32 in [256, 512))
    q = M.quantize_mx(x, "fp8_e4m3", 32)[0, :6]
    assert q.tolist() == [6.5, 1.0, -0.75, 0.3125, 0.1015625, -3.25]


@pytest.mark.parametrize("fmt", FORMATS)
@pytest.mark.parametrize("block", [16, 32, 64])
def test_mx_properties(fmt, block):
    ebits, mbits, emax, max_norm = M.FORMATS[fmt]
    x = _data(7, (16, 256))
    q = M.quantize_mx(x, fmt, block)
    # scale invariance: a power-of-two factor moves the shared exponents and nothing else
    assert np.array_equal(M.quantize_mx(x * np.float32(2.0 ** 9), fmt, block), q * np.float32(2.0 ** 9))
    assert np.array_equal(M.quantize_mx(x * np.float32(2.0 ** -20), fmt, block), q * np.float32(2.0 ** -20))
    # idempotence
    assert np.array_equal(M.quantize_mx(q, fmt, block), q)
    # every output is (integer / 2^(mbits-2)) * 2^pe * 2^se within its block, and the error is at most half a step (or the clamp)
    xb, qb = x.reshape(-1, block), q.reshape(-1, block)
    se = np.floor(np.log2(np.abs(xb).max(axis=1, keepdims=True))) - emax
    a, aq = xb / 2.0 ** se, qb / 2.0 ** se
    assert np.abs(aq).max() <= max_norm
    min_exp = 2 - 2 ** (ebits - 1) if ebits else 0
    pe = np.maximum(np.floor(np.log2(np.maximum(np.abs(a), 1e-30))), min_exp) if ebits else np.zeros_like(a)
    step = 2.0 ** (pe - (mbits - 2))
    on_grid = np.abs(aq / step - np.round(aq / step)) < 1e-6
    assert on_grid.all()
    err_ok = (np.abs(aq - a) <= step / 2 + 1e-9) | (np.abs(a) > max_norm)
    assert err_ok.all()


def test_mx_ragged_padding_and_axis():
    x = _data(11, (5, 70))
    q = M.quantize_mx(x, "fp6_e2m3", 32)
    # the last block of each row holds 6 real values + zero padding: same as quantising it alone
    assert np.array_equal(q[:, 64:], M.quantize_mx(x[:, 64:], "fp6_e2m3", 32))
    # another axis = the same quantiser on the moved axis
    y = _data(12, (4, 40, 6))
    assert np.array_equal(M.quantize_mx(y, "int8", 32, axis=1), np.moveaxis(M.quantize_mx(np.moveaxis(y, 1, -1), "int8", 32), -1, 1))
    # block_size 0: the whole axis is one block
    assert np.array_equal(M.quantize_mx(x, "fp8_e5m2", 0), M.quantize_mx(x, "fp8_e5m2", 70))


def test_mx_special_blocks():
    z = np.zeros((1, 32), np.float32)
    assert not M.quantize_mx(z, "fp8_e4m3", 32).any()
    for bad in (np.inf, -np.inf, np.nan):
        z[0, 5] = bad
        assert np.isnan(M.quantize_mx(z, "fp8_e4m3", 32)).all()            # 2^se is NaN: the whole block
    z[0, 5] = 1e-42                                                          # subnormal maximum: scale clamps at 2^-127, everything flushes
    assert not M.quantize_mx(z, "fp6_e3m2", 32).any()
    z[0, 5], z[0, 6] = 3e38, 1e38                                            # no overflow at the top of the range
    q = M.quantize_mx(z, "int8", 32)
    assert np.isfinite(q).all() and abs(q[0, 5] / 3e38 - 1) < 0.02 and abs(q[0, 6] / 1e38 - 1) < 0.02


def test_specs_helpers_match_reference_specs_py():
    ref = _load_ref("specs")
    if ref is None:
        pytest.skip("reference sources not present")
    import qsi_b200  # noqa: F401
    from qsi_b200 import mx_layers as L
    for given in (None, {}, {"bfloat": 16}, {"w_elem_format": "fp8_e4m3", "a_elem_format": "fp6_e2m3", "block_size": 32, "bfloat": 16, "scale_bits": 8},
                  {"a_elem_format": "int8", "scale_bits": 8, "block_size": 64, "round": "floor"}):
        r = ref.finalize_mx_specs(ref.apply_mx_specs(dict(given) if given is not None else None))
        o = L.finalize_mx_specs(L.apply_mx_specs(dict(given) if given is not None else None))
        assert (r is None) == (o is None), given
        if r is not None:
            assert dict(r) == dict(o), given
    with pytest.raises(KeyError):
        L.apply_mx_specs({"not_a_spec": 1})
    with pytest.raises(KeyError):
        ref.apply_mx_specs({"not_a_spec": 1})


def test_mx_layers_host_logic_without_a_gpu():
    """The host mirror on a CPU-only machine: modules without MX specs are plain torch layers; anything that would need the CUDA
    library refuses loudly (no CPU fallback, no route through the oracle)."""
    import qsi_b200  # noqa: F401
    from qsi_b200 import mx_layers as L
    lin = L.MXLinear(16, 8, mx_specs=None, sparsity=False)
    x = torch.randn(3, 16)
    assert lin.mx_none and torch.equal(lin(x), torch.nn.functional.linear(x, lin.weight, lin.bias))
    conv = L.MXConv2d(3, 4, kernel_size=2, mx_specs={})
    assert conv.mx_none and conv(torch.randn(1, 3, 4, 4)).shape == (1, 4, 3, 3)
    assert torch.equal(L.MXMatmul(x, x.t()), x @ x.t())
    spec = dict(w_elem_format="fp8_e4m3", a_elem_format="fp8_e4m3", block_size=32, bfloat=16, scale_bits=8)
    q = L.MXLinear(64, 8, mx_specs=spec)
    assert not q.mx_none and q.mx_specs["w_elem_format_bp"] == "fp8_e4m3"
    if not torch.cuda.is_available():
        with pytest.raises(ValueError, match="CUDA"):
            q(torch.randn(2, 64))
        with pytest.raises(ValueError, match="CUDA"):
            L.quantize_mx_op(torch.randn(2, 64), q.mx_specs, "fp8_e4m3", axes=[-1])
    assert L._format_id("fp4") == L._format_id("FP4_E2M1") == 8 and L._format_id(None) is None
    assert L._block_scaled_ok(5, 8, q.mx_specs, 4096, torch.float32, torch.float32, 4096)
    assert not L._block_scaled_ok(5, 1, q.mx_specs, 4096, torch.float32, torch.float32, 4096)          # int8 is not an E4M3 subset
    assert not L._block_scaled_ok(5, 5, dict(q.mx_specs, block_size=16), 4096, torch.float32, torch.float32, 4096)
    assert not L._block_scaled_ok(5, 5, q.mx_specs, 4224, torch.bfloat16, torch.bfloat16, 4096)        # 16-bit inputs need K % 256 == 0
    with pytest.raises(NotImplementedError):
        L._need_nearest("floor")
    with pytest.raises(ValueError):
        L._bfloat_of(dict(q.mx_specs, bfloat=8))


def test_packed_weight_cache_protocol_on_cpu(monkeypatch):
    """The cache of the MX modules (the BFPLinear protocol): keyed on storage / version, bypassed while a module trains with a
    trainable weight, dropped by invalidate(), switched off or checksummed by BFP_WEIGHT_CACHE."""
    import qsi_b200  # noqa: F401
    from qsi_b200 import mx_layers as L
    lin = torch.nn.Linear(8, 4).eval()
    cache, builds = L._PackedWeightCache(), []

    def build():
        builds.append(1)
        return len(builds)
    assert cache.get(lin, lin.weight, "bs", build) == 1 and cache.get(lin, lin.weight, "bs", build) == 1          # hit
    assert cache.get(lin, lin.weight, "bf16", build) == 2 and len(cache.entries) == 2                             # second kind, same weight
    with torch.no_grad():
        lin.weight.mul_(2.0)                                                                                       # version moves: every form goes
    assert cache.get(lin, lin.weight, "bs", build) == 3 and len(cache.entries) == 1
    lin.weight.data.mul_(2.0)                                                                                      # .data write: version does not move ...
    assert cache.get(lin, lin.weight, "bs", build) == 3
    cache.invalidate()                                                                                             # ... the documented protocol
    assert cache.get(lin, lin.weight, "bs", build) == 4
    lin.train()                                                                                                    # training + trainable weight: never cached
    assert cache.get(lin, lin.weight, "bs", build) == 5 and cache.get(lin, lin.weight, "bs", build) == 6
    lin.eval()
    monkeypatch.setenv("BFP_WEIGHT_CACHE", "verify")
    assert cache.get(lin, lin.weight, "bs", build) == 7 and cache.get(lin, lin.weight, "bs", build) == 7
    lin.weight.data.mul_(2.0)                                                                                      # the checksum sees a .data write
    assert cache.get(lin, lin.weight, "bs", build) == 8
    monkeypatch.setenv("BFP_WEIGHT_CACHE", "0")
    assert cache.get(lin, lin.weight, "bs", build) == 9 and cache.get(lin, lin.weight, "bs", build) == 10


def test_philox_offsets_are_disjoint_across_ranks(monkeypatch):
    import qsi_b200  # noqa: F401
    from qsi_b200 import bfp_ops as B
    B._PhiloxState.seed = None
    monkeypatch.setenv("RANK", "0")
    s0, o0 = B._PhiloxState.next()
    _, o0b = B._PhiloxState.next()
    monkeypatch.setenv("RANK", "3")
    s3, o3 = B._PhiloxState.next()
    assert s0 == s3 == (torch.initial_seed() & 0xFFFFFFFFFFFFFFFF)
    assert o0b == o0 + 1 and (o3 >> 48) == 3 and (o0 >> 48) == 0 and (o3 & ((1 << 48) - 1)) == 2
    B._PhiloxState.seed = None


@pytest.mark.parametrize("fmt", FORMATS)
def test_numpy_oracle_equals_the_torch_op_sequence(fmt):
    """Two independent restatements of the library's _quantize_mx -- numpy (the oracle) and the torch op sequence the library itself
    runs (tests/test_mx_gpu._torch_quantize_mx, evaluated here by torch-CPU; the GPU suite evaluates it with torch-CUDA against the
    kernel) -- agree bit for bit, signed zeros included."""
    import test_mx_gpu as G
    for seed, scale in ((0, 1.0), (1, 1e-3), (2, 300.0)):
        x = _data(30 + seed, (48, 256), scale)
        x.flat[3::101] = 0.0
        x.flat[5::103] = -0.0
        x[0, :32] = 0.0
        want = M.quantize_mx(x, fmt, 32)
        got = G._torch_quantize_mx(torch.from_numpy(x.copy()), fmt, 32, M).numpy()
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (fmt, seed)


def test_tables_match_the_recorded_reference_outputs():
    """tests/golden/mx_reference_tables.json (made by tests/golden/make_mx_golden.py from the reference's formats.py / specs.py):
    the oracle's format table and this package's spec helpers reproduce what the reference tree itself pins of the MX path."""
    import json
    import qsi_b200  # noqa: F401
    from qsi_b200 import mx_layers as L
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mx_reference_tables.json")))
    for name, r in g["formats"].items():
        assert M.FORMATS[name] == (r["ebits"], r["mbits"], r["emax"], r["max_norm"]), name
        assert M.FORMAT_IDS[name] == r["id"] == L.ELEM_FORMATS[name], name
    for case in g["specs"]:
        out = L.finalize_mx_specs(L.apply_mx_specs(dict(case["given"]) if case["given"] is not None else None))
        assert (out is None) == (case["finalized"] is None), case["given"]
        if out is not None:
            assert dict(out) == case["finalized"], case["given"]
