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
