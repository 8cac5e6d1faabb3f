"""GPU parity for global (unstructured) magnitude sparsity, bfp_ops.py:61-71:
bit-exact against the oracle (torch-CUDA tie order) and against the live reference on torch-CUDA when present."""
import itertools

import numpy as np
import pytest
import torch

import _golden
from _refload import load_reference, ref_args

pytestmark = pytest.mark.gpu
TORCH_DT = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available()
    from qsi_b200 import bfp_ops, _lib
    _lib.lib()
    return bfp_ops


def _np(t):
    from oracle import bfp_oracle as O
    return O.from_torch(t)


def _cases():
    g = torch.Generator().manual_seed(3)
    yield "randn", torch.randn(257, 1000, generator=g)
    yield "ties", torch.randint(-4, 5, (100, 513), generator=g).float()
    z = torch.randn(64, 256, generator=g)
    z[::2] = 0.0
    z[1, 3] = -0.0
    yield "zeros", z
    s = torch.randn(33, 77, generator=g)
    s[0, 0] = float("nan"); s[5, 5] = float("inf"); s[6, 6] = -float("inf")
    yield "specials", s
    yield "tiny", torch.randn(3, generator=g)
    yield "const", torch.full((10, 100), 0.5)


@pytest.mark.parametrize("dt", ["f32", "bf16", "f16"])
def test_unstructured_matches_oracle(ops, oracle, dt):
    for (name, x32), frac in itertools.product(_cases(), (0.5, 0.1, 0.9, 0.999, 1e-4)):
        x = x32.to(TORCH_DT[dt])
        y = ops._unstructured_sparsity(x.cuda(), "cuda", frac)
        assert y.shape == x.shape and y.dtype == x.dtype
        o = oracle.unstructured_sparsify(_np(x)[0], frac, dt=dt)
        assert _golden.mismatches(_np(y)[0], o, dt) == 0, (name, frac)
        k = int(x.numel() * frac)
        assert int((y == 0).sum()) >= min(k, x.numel()) - int(torch.isnan(x).sum())


@pytest.mark.parametrize("dt", ["f32", "bf16"])
def test_unstructured_matches_live_reference_on_cuda(ops, dt):
    ref = load_reference()
    if ref is None:
        pytest.skip("reference sources not present on this box")
    for (name, x32), frac in itertools.product(_cases(), (0.5, 0.25, 0.9)):
        x = x32.to(TORCH_DT[dt]).cuda()
        k = int(x.numel() * frac)
        if k == 0:
            continue
        r = ref._unstructured_sparsity(x, "cuda", frac)
        y = ops._unstructured_sparsity(x, "cuda", frac)
        assert _golden.mismatches(_np(y)[0], _np(r)[0], dt) == 0, (name, frac)
    # a large slice (torch switches its top-k algorithm with the slice size)
    g = torch.Generator(device="cuda").manual_seed(0)
    w = (torch.randn(2048, 4096, device="cuda", generator=g) * 0.02).to(TORCH_DT[dt])
    w.view(-1)[::5] = w.view(-1)[1::5]                     # plenty of exact ties
    assert _golden.mismatches(_np(ops._unstructured_sparsity(w, "cuda", 0.5))[0], _np(ref._unstructured_sparsity(w, "cuda", 0.5))[0], dt) == 0


@pytest.mark.parametrize("first", ["s", "q"])
def test_entry_point_with_unstructured_sparsity(ops, oracle, first):
    """float_to_bfp_blocked with sparsity_mode='unstructured' composes the two kernels in the reference's order."""
    g = torch.Generator().manual_seed(9)
    w = torch.randn(128, 512, generator=g) * 0.02
    args = ops.unpack_bfp_args(dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=5,
                                    block_size=32, w_sparsity=True, sparsity_frac=0.6, first=first, sparsity_mode="unstructured",
                                    device="cuda"))
    y = ops.float_to_bfp_blocked(w.cuda(), **args, identifier="w").cpu().numpy()
    if first == "s":
        o, _ = oracle.bfp_quantize(oracle.unstructured_sparsify(w.numpy(), 0.6), 32, 5)
    else:
        o = oracle.unstructured_sparsify(oracle.bfp_quantize(w.numpy(), 32, 5)[0], 0.6)
    assert _golden.mismatches(y, o, "f32") == 0
    assert int((y == 0).sum()) >= int(w.numel() * 0.6)
    ys = ops.float_to_bfp_blocked(w.cuda(), **dict(args, sparsity_num_format="fp32"), identifier="w").cpu().numpy()
    assert _golden.mismatches(ys, oracle.unstructured_sparsify(w.numpy(), 0.6), "f32") == 0


def test_full_size_unstructured_properties(ops, oracle):
    g = torch.Generator(device="cuda").manual_seed(1)
    w = torch.randn(4096, 4096, device="cuda", generator=g) * 0.02
    y = ops._unstructured_sparsity(w, "cuda", 0.5)
    k = w.numel() // 2
    assert int((y == 0).sum()) == k                                    # continuous data: exactly k zeros
    kept = y != 0
    assert torch.equal(y[kept], w[kept])
    assert w[~kept].abs().max() <= w[kept].abs().min()                 # a magnitude threshold separates them
    assert torch.equal(ops._unstructured_sparsity(y, "cuda", 0.5), y)  # idempotent
    rows = w[:8].cpu().numpy()                                         # the oracle on a slice with its own k
    assert _golden.mismatches(ops._unstructured_sparsity(w[:8].contiguous(), "cuda", 0.5).cpu().numpy(), oracle.unstructured_sparsify(rows, 0.5), "f32") == 0


# ---------------------------------------------------------------------------------------------------------------
# The fused two-read pipeline (csrc/bfp_unstructured_fused.cu): sampled bracket -> counting pass -> refine -> apply.
# ---------------------------------------------------------------------------------------------------------------
def _fused_cases():
    """Shapes the fused path takes (numel a multiple of the vector, K a multiple of the block)."""
    g = torch.Generator().manual_seed(11)
    yield "randn", torch.randn(257, 1024, generator=g) * 0.02
    yield "ties", torch.randint(-4, 5, (100, 512), generator=g).float()
    z = torch.randn(64, 256, generator=g)
    z[::2] = 0.0
    z[1, 3] = -0.0
    yield "zeros", z
    s = torch.randn(33, 128, generator=g)
    s[0, 0] = float("nan"); s[5, 5] = float("inf"); s[6, 6] = -float("inf")
    yield "specials", s
    yield "tiny", torch.randn(1, 64, generator=g)
    yield "const_round", torch.full((16, 128), 0.5)
    yield "const_odd", torch.full((300, 256), 0.3)          # one non-round value everywhere: the candidate list overflows
    o = torch.randn(64, 512, generator=g)
    o[:, 7] *= 1000.0                                        # an outlier channel
    yield "outliers", o
    yield "two_values", torch.where(torch.rand(128, 256, generator=g) < 0.5, torch.tensor(0.3), torch.tensor(0.7))


def _oracle_compose(oracle, x_np, frac, order, B, m, dt):
    if order == "s":
        return oracle.unstructured_sparsify(x_np, frac, dt=dt)
    if order == "sq":
        return oracle.bfp_quantize(oracle.unstructured_sparsify(x_np, frac, dt=dt), B, m, dt=dt)[0]
    return oracle.unstructured_sparsify(oracle.bfp_quantize(x_np, B, m, dt=dt)[0], frac, dt=dt)


def _run_fused(ops, x, frac, order, B, m, rounding="determ", philox=None):
    from qsi_b200 import _lib
    code = {"s": _lib.ORDER_SPARSIFY_ONLY, "sq": _lib.ORDER_SPARSIFY_QUANT, "qs": _lib.ORDER_QUANT_SPARSIFY}[order]
    y = ops._unstructured_fused(x, frac, code, block_size=B, mant_bits=m, epsilon=1e-8, rounding_mode=rounding, philox=philox)
    assert y is not None, "the fused path refused a shape it should take"
    return y


@pytest.mark.parametrize("dt", ["f32", "bf16", "f16"])
@pytest.mark.parametrize("order", ["s", "sq", "qs"])
def test_fused_unstructured_matches_oracle(ops, oracle, dt, order):
    for (name, x32), frac, (B, m) in itertools.product(_fused_cases(), (0.5, 0.1, 0.9, 0.999), ((64, 7), (16, 3), (32, 5))):
        if order == "s" and (B, m) != (64, 7):
            continue
        x = x32.to(TORCH_DT[dt])
        if int(x.numel() * frac) in (0, x.numel()):
            continue
        y = _run_fused(ops, x.cuda(), frac, order, B, m)
        o = _oracle_compose(oracle, _np(x)[0], frac, order, B, m, dt)
        assert _golden.mismatches(_np(y)[0], o, dt) == 0, (name, frac, B, m)


@pytest.mark.parametrize("dt", ["f32", "bf16"])
@pytest.mark.parametrize("order", ["s", "sq", "qs"])
def test_fused_unstructured_whole_tensor_select_path(ops, oracle, dt, order):
    """The rare path of the refine kernel (missed bracket / overflowing candidate list), forced: same results."""
    from qsi_b200 import _lib
    _lib.check(_lib.lib().bfp_set_option(b"unstructured_force_fallback", 1))
    try:
        for (name, x32), frac in itertools.product(_fused_cases(), (0.5, 0.25)):
            x = x32.to(TORCH_DT[dt])
            if int(x.numel() * frac) in (0, x.numel()):
                continue
            y = _run_fused(ops, x.cuda(), frac, order, 32, 5)
            assert _golden.mismatches(_np(y)[0], _oracle_compose(oracle, _np(x)[0], frac, order, 32, 5, dt), dt) == 0, (name, frac)
    finally:
        _lib.check(_lib.lib().bfp_set_option(b"unstructured_force_fallback", 0))


def _compose_kernels(ops, x, frac, order, B, m, rounding="determ", philox=None):
    """The same result from the stand-alone kernels (multi-pass radix select + streaming quantiser)."""
    import os
    from qsi_b200 import _lib
    os.environ["BFP_UNSTRUCTURED_FUSED"] = "0"
    try:
        q = lambda t: ops._fused(t, _lib.ORDER_QUANT_ONLY, block_size=B, mant_bits=m, epsilon=1e-8, rounding_mode=rounding, philox=philox)
        if order == "s":
            return ops._unstructured_sparsity(x, "cuda", frac)
        if order == "sq":
            return q(ops._unstructured_sparsity(x, "cuda", frac))
        return ops._unstructured_sparsity(q(x), "cuda", frac)
    finally:
        os.environ.pop("BFP_UNSTRUCTURED_FUSED", None)


@pytest.mark.parametrize("dt", ["f32", "bf16", "f16"])
def test_fused_unstructured_full_size_equals_composition(ops, dt):
    """LLaMA-7B sized tensors: the two-read pipeline against the stand-alone kernels, every order; inputs with continuous
    values, with massive ties (pre-quantised, half zeros) and ReLU-like zeros -- the ranked-tie apply path over thousands of tiles."""
    g = torch.Generator(device="cuda").manual_seed(5)
    w = (torch.randn(4096, 4096, device="cuda", generator=g) * 0.02).to(TORCH_DT[dt])
    relu = torch.relu(w)
    coarse = (torch.randn(2048, 4096, device="cuda", generator=g) * 3).round().to(TORCH_DT[dt])      # small integers: huge tie groups
    for x, name in ((w, "randn"), (relu, "relu"), (coarse, "coarse")):
        for order, frac in (("s", 0.5), ("sq", 0.5), ("qs", 0.5), ("qs", 0.3), ("sq", 0.7)):
            y = _run_fused(ops, x, frac, order, 64, 7)
            r = _compose_kernels(ops, x, frac, order, 64, 7)
            assert y.dtype == r.dtype and torch.equal(y.view(torch.int16 if dt != "f32" else torch.int32), r.view(torch.int16 if dt != "f32" else torch.int32)), (name, order, frac)
            k = int(x.numel() * frac)
            assert int((y == 0).sum()) >= k
    # HBFP4 after quantisation: a handful of distinct magnitudes
    y = _run_fused(ops, w, 0.5, "qs", 64, 3)
    r = _compose_kernels(ops, w, 0.5, "qs", 64, 3)
    assert torch.equal(y, r)


@pytest.mark.parametrize("dt", ["f32", "bf16"])
@pytest.mark.parametrize("order", ["sq", "qs"])
def test_fused_unstructured_stochastic_equals_composition(ops, dt, order):
    """Stochastic rounding uses the Philox counters of the streaming quantiser, so the fused result equals the composition draw
    for draw; with first == 'q' the keys of the quantised values are recomputed identically in every pass."""
    g = torch.Generator(device="cuda").manual_seed(6)
    w = (torch.randn(1024, 2048, device="cuda", generator=g) * 0.02).to(TORCH_DT[dt])
    for m, B in ((7, 64), (3, 16)):
        y = _run_fused(ops, w, 0.5, order, B, m, rounding="stoc", philox=(1234, 7))
        r = _compose_kernels(ops, w, 0.5, order, B, m, rounding="stoc", philox=(1234, 7))
        assert y.dtype == torch.float32 and torch.equal(y, r), (m, B)
        assert int((y == 0).sum()) >= w.numel() // 2


def test_fused_unstructured_entry_point_and_live_reference(ops):
    """float_to_bfp_blocked with sparsity_mode='unstructured' takes the fused path and agrees with the live reference."""
    ref = load_reference()
    g = torch.Generator(device="cuda").manual_seed(8)
    w = torch.randn(512, 1024, device="cuda", generator=g) * 0.02
    n0 = None
    from qsi_b200 import _lib
    for first in ("s", "q"):
        cfg = dict(num_format="bfp", sparsity_num_format="bfp", rounding_mode="determ", epsilon=1e-8, mant_bits=7, block_size=64,
                   w_sparsity=True, sparsity_frac=0.5, first=first, sparsity_mode="unstructured", device="cuda")
        n0 = _lib.launch_count()
        y = ops.float_to_bfp_blocked(w, **ops.unpack_bfp_args(dict(cfg)), identifier="w")
        assert _lib.launch_count() - n0 == 4, "expected the four kernels of the fused pipeline"
        if ref is not None:
            r = ref.float_to_bfp_blocked(w, **ref_args(ref, **cfg), identifier="w")
            assert torch.equal(y, r), first


@pytest.mark.parametrize("dt", ["f32", "bf16", "f16"])
def test_fused_unstructured_midsize_brackets_match_oracle(ops, oracle, dt):
    """Tensors larger than the sample (so the bracket is a real estimate): window mode, list mode (a bracket straddling the zeros
    of a ReLU output), duplicated values whose ties are cut by index, a cut that falls between two ranges."""
    g = torch.Generator().manual_seed(21)
    w = torch.randn(512, 1024, generator=g) * 0.02
    dup = w.clone(); dup.view(-1)[::3] = dup.view(-1)[1::3][: dup.view(-1)[::3].numel()]          # plenty of exact ties
    cases = [("randn", w), ("relu", torch.relu(w)), ("coarse", (torch.randn(512, 1024, generator=g) * 3).round()), ("dup", dup),
             ("sparse90", torch.where(torch.rand(512, 1024, generator=g) < 0.9, torch.zeros(()), w))]
    for (name, x32), (order, frac) in itertools.product(cases, (("s", 0.5), ("sq", 0.5), ("qs", 0.5), ("s", 0.93), ("sq", 0.25), ("qs", 0.75))):
        x = x32.to(TORCH_DT[dt])
        y = _run_fused(ops, x.cuda(), frac, order, 64, 7)
        o = _oracle_compose(oracle, _np(x)[0], frac, order, 64, 7, dt)
        assert _golden.mismatches(_np(y)[0], o, dt) == 0, (name, order, frac)
