"""CPU suite: the C-ABI library loads, exports every symbol include/bfp_b200.h declares, validates arguments like the
reference does, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import itertools
import os
import re

import numpy as np
import pytest
import torch

import qsi_b200
from qsi_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    out = []
    for fn in sorted(os.listdir(os.path.join(ROOT, "include"))):
        src = open(os.path.join(ROOT, "include", fn)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        out += re.findall(r"\b(bfp_[a-z0-9_]+)\s*\(", src)
    return sorted(set(out))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = _header_functions()
    assert len(names) >= 9
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in _lib.py"
    assert L.bfp_version() == 100


def test_argument_validation_mirrors_reference_asserts():
    L = _lib.lib()
    buf = (ctypes.c_float * 64)()
    p = ctypes.addressof(buf)
    q = lambda **kw: L.bfp_quantize(p, p + 128, kw.get("rows", 1), kw.get("K", 16), kw.get("dt", 0), kw.get("odt", 0),
                                    kw.get("B", 16), kw.get("m", 7), 1e-8, kw.get("rnd", 0), 0, 0, kw.get("N", 2),
                                    kw.get("M", 4), kw.get("order", 1), kw.get("tie", 0), None)
    assert q(B=0) == _lib.E_ARG and b"block_size" in L.bfp_last_error()          # bfp_ops.py:130
    assert q(N=0) == _lib.E_ARG and q(N=5, M=4) == _lib.E_ARG                     # bfp_ops.py:74
    assert q(rnd=7) == _lib.E_ARG and b"Rounding mode" in L.bfp_last_error()      # bfp_ops.py:27
    assert q(odt=2) == _lib.E_ARG                                                 # nearest keeps the dtype
    assert q(rnd=1, odt=0, dt=2) == _lib.E_CUDA or q(rnd=1, odt=0, dt=2) == _lib.OK   # stoc on bf16 -> fp32 out is valid
    assert q(rnd=1, odt=2, dt=2) == _lib.E_ARG
    assert q(M=65, N=1) == _lib.E_UNSUPPORTED
    assert q(order=9) == _lib.E_ARG
    assert L.bfp_quantize(p + 2, p + 128, 1, 16, 0, 0, 16, 7, 1e-8, 0, 0, 0, 2, 4, 1, 0, None) == _lib.E_ALIGN


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_compute_fails_loudly_without_gpu():
    from qsi_b200 import bfp_ops
    t = torch.randn(4, 64)
    with pytest.raises(_lib.BFPLibraryError) as ei:
        bfp_ops._no_sparsity_float_to_bfp(t, 64, 7, 1e-8, "determ", "cpu")
    assert ei.value.code == _lib.E_CUDA and "no CPU fallback" in str(ei.value)


def test_cpu_tie_table_matches_oracle_and_torch(oracle):
    """The committed 2:4 torch-CPU tie table (csrc/nm_cpu_tie_lut.inc) == the oracle's std::nth_element == torch.topk."""
    txt = open(os.path.join(_lib.CSRC, "nm_cpu_tie_lut.inc")).read()
    txt = "\n".join(l for l in txt.splitlines() if not l.startswith("//"))
    lut = [int(x, 16) for x in re.findall(r"0x([0-9A-F]{2})", txt)]
    assert len(lut) == 256
    seen = 0
    for p in itertools.product(range(4), repeat=4):
        c = [sum(p[j] < p[i] for j in range(4)) for i in range(4)]
        idx = c[0] + 4 * c[1] + 16 * c[2] + 64 * c[3]
        x = np.array([[float(v + 1) for v in p]], np.float32)
        y = oracle.nm_sparsify(x, 2, 4, tie_rule="cpu")[0]
        assert lut[idx] == sum(1 << i for i in range(4) if y[i] == 0)
        _, drop = torch.topk(torch.from_numpy(x).abs(), k=2, dim=1, largest=False)
        assert lut[idx] == sum(1 << int(i) for i in drop[0])
        seen += 1
    assert seen == 256 and sum(v != 0xFF for v in lut) == 75


def test_python_surface_matches_reference_names():
    from qsi_b200 import bfp_ops
    names = ["rounding_modes", "round_tensor", "get_exponent", "_convert_blocked_float_to_bfp",
             "_no_sparsity_float_to_bfp", "_unstructured_sparsity", "_structured_N_M_sparsity", "_sparsify", "_quantize",
             "float_to_bfp_blocked", "MxM_pre_processing", "_get_op_name", "_gen_bfp_op", "_get_bfp_op",
             "unpack_bfp_args", "F_linear_bfp", "F_matmul_bfp", "BFPConv2d", "BFPLinear", "float_to_bfp_tiled", "BFPConv1D"]
    for n in names:
        assert hasattr(bfp_ops, n), n
    # signature parity with the reference where it is present
    from _refload import load_reference
    ref = load_reference()
    if ref is not None:
        import inspect
        for n in names:
            if hasattr(ref, n) and inspect.isfunction(getattr(ref, n)):
                assert str(inspect.signature(getattr(ref, n))) == str(inspect.signature(getattr(bfp_ops, n))), n
    kw = dict(num_format="bfp", mant_bits=7, bogus=1)
    a = bfp_ops.unpack_bfp_args(kw)
    assert kw == {"bogus": 1} and a["rounding_mode"] == "stoc" and a["device"] == "cpu" and a["epsilon"] == 1e-8
    assert len(a) == 20
    assert bfp_ops.F_linear_bfp(num_format="fp32") is torch.nn.functional.linear
    assert bfp_ops.F_matmul_bfp() is torch.matmul
    assert bfp_ops._get_op_name("linear", **bfp_ops.unpack_bfp_args(dict(mant_bits=7, rounding_mode="determ"))) == "linear_BFP_determ_7"
    m = qsi_b200.install_as_reference_module()
    import importlib
    assert importlib.import_module("transformers.bfp.bfp_ops") is m


def test_entry_point_exceptions_without_touching_the_gpu():
    from qsi_b200 import bfp_ops
    t = torch.randn(2, 8)
    base = bfp_ops.unpack_bfp_args(dict(num_format="bfp", sparsity_num_format="fp32", rounding_mode="determ"))
    assert bfp_ops.float_to_bfp_blocked(t, **base, identifier="w") is t              # identity returns the same object
    with pytest.raises(AssertionError):
        bfp_ops.float_to_bfp_blocked(t, **dict(base, num_format="fp32"))
    with pytest.raises(AssertionError):
        bfp_ops.float_to_bfp_blocked(t, **dict(base, sparsity_num_format="bfp", block_size=0))
    with pytest.raises(ValueError):
        bfp_ops._sparsify(t, True, "banana", "cpu", 2, 4, 0.5)
    with pytest.raises(ValueError):
        bfp_ops._quantize(t, "fp8", 64, 7, 15, False, 1e-8, "determ", "cpu", "w")
    with pytest.raises(AssertionError):
        bfp_ops._structured_N_M_sparsity(t, "cpu", 0, 4)
    with pytest.raises(NotImplementedError):
        bfp_ops._fused(t, _lib.ORDER_QUANT_ONLY, block_size=8, mant_bits=3, rounding_mode="nearest")
    lin = bfp_ops.BFPLinear(8, 4, num_format="fp32")
    assert lin(t).shape == (2, 4) and set(dict(lin.named_parameters())) == {"weight", "bias"}
    lin.num_format = "weird"
    with pytest.raises(NotImplementedError):
        lin(t)


def test_nm_patterns_that_fit_the_2to4_tensor_core_format():
    """_nm_fits_2to4 against brute force: every choice of N survivors per M-group, worst aligned group of four."""
    import itertools, math
    from qsi_b200 import bfp_ops
    for M in (1, 2, 3, 4, 6, 8, 12, 16):
        for N in range(1, M):
            period = M * 4 // math.gcd(M, 4)
            worst = 0
            for keeps in itertools.product(itertools.combinations(range(M), N), repeat=period // M):
                mask = [(i % M) in keeps[i // M] for i in range(period)]
                worst = max(worst, max(sum(mask[i:i + 4]) for i in range(0, period, 4)))
            assert bfp_ops._nm_fits_2to4(N, M) == (worst <= 2), (N, M, worst)
    assert bfp_ops._nm_fits_2to4(2, 4) and bfp_ops._nm_fits_2to4(1, 4) and bfp_ops._nm_fits_2to4(1, 2) and bfp_ops._nm_fits_2to4(2, 8)
    assert not bfp_ops._nm_fits_2to4(4, 4) and not bfp_ops._nm_fits_2to4(0, 4) and not bfp_ops._nm_fits_2to4(3, 4)
