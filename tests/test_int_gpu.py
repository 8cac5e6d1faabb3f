"""GPU parity for the 'int' number format (sparsity_num_format='int', bfp_ops.py:111-120 -> int_ops.Quantizer;
SURVEY.md section 8 row f2): bit-exact against the oracle, the golden fixtures and the live reference."""
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


def _args(ops, bits, **kw):
    return ops.unpack_bfp_args(dict(num_format="bfp", sparsity_num_format="int", rounding_mode="determ", epsilon=1e-8, mant_bits=bits,
                                    weight_mant_bits=15, block_size=64, device="cuda", **kw))


@pytest.mark.parametrize("dt", ["f32", "bf16", "f16"])
def test_int_format_matches_oracle(ops, oracle, dt):
    g = torch.Generator().manual_seed(2)
    for shape, ident, bits in itertools.product([(64, 1000), (3, 50, 96), (2, 7, 12, 20), (2, 7, 4, 4), (5, 8), (4096, 64)], ("w", "in"), (8, 4, 2)):
        x = (torch.randn(*shape, generator=g) * 0.3).to(TORCH_DT[dt])
        x.view(-1)[::11] = 0
        if len(shape) == 2:
            x[1] = x[1].abs()
            x[2] = 0
        y = ops.float_to_bfp_blocked(x.cuda(), **_args(ops, bits), identifier=ident)
        assert y.dtype == torch.float32 and y.shape == x.shape
        from oracle import bfp_oracle as O
        o = oracle.int_quantize(O.from_torch(x)[0], bits, ident == "w", dt=dt)
        assert _golden.mismatches(y.cpu().numpy(), o, "f32") == 0, (shape, ident, bits)
        levels = torch.unique(y[0]).numel() if ident == "w" else 0
        assert levels <= 2 ** bits


@pytest.mark.parametrize("device", ["cuda", "cpu"])
def test_int_format_and_unstructured_match_golden(ops, device):
    from oracle import bfp_oracle as O
    z = _golden.load(device)
    if z is None or not any(k.startswith("int_in_") for k in z.files):
        pytest.skip("not recorded")
    n = 0
    for k in z.files:
        if k.startswith("int_out_"):
            name, dtn, ident, bits = k[len("int_out_"):].rsplit("__", 1)[0].rsplit("_", 3)
            x, dt = _golden.get(z, f"int_in_{name}_{dtn}")
            y = ops.float_to_bfp_blocked(O.to_torch(x, dt).cuda(), **_args(ops, int(bits[1:])), identifier=ident)
            assert _golden.mismatches(y.cpu().numpy(), z[k], "f32") == 0, k
            n += 1
        elif k.startswith("un_out_"):
            name, dtn, frac = k[len("un_out_"):].rsplit("__", 1)[0].rsplit("_", 2)
            x, dt = _golden.get(z, f"un_in_{name}_{dtn}")
            y = ops._unstructured_sparsity(O.to_torch(x, dt).cuda(), "cuda", float(frac))
            assert _golden.mismatches(O.from_torch(y)[0], z[k], dt) == 0, k
            n += 1
    assert n >= 36


def test_int_format_with_sparsity_and_live_reference(ops, oracle):
    """int format composed with 2:4 / unstructured sparsity in both orders, against the live reference when present."""
    ref = load_reference()
    g = torch.Generator().manual_seed(4)
    w = torch.randn(96, 256, generator=g) * 0.05
    for first, mode in itertools.product(("s", "q"), ("structured", "unstructured")):
        a = _args(ops, 4, w_sparsity=True, N=2, M=4, first=first, sparsity_mode=mode, sparsity_frac=0.5)
        y = ops.float_to_bfp_blocked(w.cuda(), **a, identifier="w").cpu().numpy()
        sp = (lambda t: oracle.nm_sparsify(t, 2, 4)) if mode == "structured" else (lambda t: oracle.unstructured_sparsify(t, 0.5))
        o = oracle.int_quantize(sp(w.numpy()), 4, True) if first == "s" else sp(oracle.int_quantize(w.numpy(), 4, True))
        assert _golden.mismatches(y, o, "f32") == 0, (first, mode)
        if ref is not None:
            ra = ref_args(ref, sparsity_num_format="int", mant_bits=4, first=first, sparsity_mode=mode, device="cuda")
            r = ref.float_to_bfp_blocked(w.cuda(), **ra, identifier="w").cpu().numpy()
            assert _golden.mismatches(y, r, "f32") == 0, (first, mode, "reference")
    # sgd_update switches to weight_mant_bits (bfp_ops.py:113-114)
    y = ops.float_to_bfp_blocked(w.cuda(), **_args(ops, 4), identifier="w", sgd_update=True).cpu().numpy()
    assert _golden.mismatches(y, oracle.int_quantize(w.numpy(), 15, True), "f32") == 0


@pytest.mark.parametrize("dt", ["f32", "bf16", "f16"])
def test_int_with_n4_sparsity_single_pass_equals_composition(ops, oracle, dt):
    """2-D weights take one fused kernel (bfp_int_quantize_nm); it must equal the two stand-alone kernels composed the way
    bfp_ops.py:143-149 composes them, and the oracle, bit for bit -- including the heavy ties of 2- and 4-bit grids."""
    from oracle import bfp_oracle as O
    g = torch.Generator().manual_seed(9)
    for (rows, K), bits, N, first in itertools.product([(40, 64), (33, 4096), (9, 11008), (5, 16384)], (8, 4, 2), (1, 2, 3), ("s", "q")):
        w = (torch.randn(rows, K, generator=g) * 0.05).to(TORCH_DT[dt])
        w.view(-1)[::7] = 0
        w[1] = w[1].abs()
        w[2] = 0
        a = _args(ops, bits, w_sparsity=True, N=N, M=4, first=first, sparsity_mode="structured")
        assert ops._int_nm_fusable(w, N, 4)
        y = ops.float_to_bfp_blocked(w.cuda(), **a, identifier="w")
        assert y.dtype == torch.float32 and y.shape == w.shape
        if first == "s":
            c = ops._int_quantize(ops._structured_N_M_sparsity(w.cuda(), "cuda", N, 4), bits, weight=True)
        else:
            c = ops._structured_N_M_sparsity(ops._int_quantize(w.cuda(), bits, weight=True), "cuda", N, 4)
        assert torch.equal(y.view(torch.int32), c.view(torch.int32)), (rows, K, bits, N, first)
        if K <= 4096:
            x = O.from_torch(w)[0]
            o = (oracle.int_quantize(oracle.nm_sparsify(x, N, 4, dt=dt), bits, True, dt=dt) if first == "s"
                 else oracle.nm_sparsify(oracle.int_quantize(x, bits, True, dt=dt), N, 4))
            assert _golden.mismatches(y.cpu().numpy(), o, "f32") == 0, (rows, K, bits, N, first, "oracle")
    # shapes outside the fused kernel's reach keep composing
    assert not ops._int_nm_fusable(torch.zeros(4, 6), 2, 4) and not ops._int_nm_fusable(torch.zeros(2, 3, 8), 2, 4)
    assert not ops._int_nm_fusable(torch.zeros(4, 8), 2, 8)


def test_int_activation_three_plane_split(ops):
    """bfp_int_quantize_split3: hi + mid + lo reproduces the fp32 fake-quantised activation (exactly, up to a last-place
    remainder the third bf16 plane cannot hold), in the segment-major layout of include/bfp_b200.h."""
    from qsi_b200 import _lib
    g = torch.Generator().manual_seed(12)
    A, C, kseg = 150, 264, 64
    x = (torch.randn(3, 50, C, generator=g) * torch.rand(C, generator=g) * 4).cuda()
    x[..., 5] = 0
    for bits in (8, 4):
        xq = ops._int_quantize(x, bits, weight=False).view(A, C)
        x3 = torch.empty(A * 3 * C, dtype=torch.bfloat16, device="cuda")
        L = _lib.lib()
        ws = torch.empty(L.bfp_int_workspace_bytes(C) // 8 + 2, dtype=torch.int64, device="cuda")
        _lib.check(L.bfp_int_quantize_split3(x.data_ptr(), x3.data_ptr(), A, C, kseg, _lib.DT_F32, bits, ws.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream))
        hi = torch.cat([x3[k0 * A: k0 * A + A * min(kseg, C - k0)].view(A, -1) for k0 in range(0, C, kseg)], dim=1)
        ml = x3[A * C:].view(A, 2 * C)
        s = hi.double() + ml[:, :C].double() + ml[:, C:].double()
        err = (s - xq.double()).abs()
        assert (err <= xq.double().abs() * 2.0 ** -24).all()
        assert float((err == 0).double().mean()) > 0.99
        assert torch.equal(hi, xq.to(torch.bfloat16))


@pytest.mark.parametrize("cfg", [("structured", "s"), ("structured", "q"), ("unstructured", "s"), ("none", "s")])
@pytest.mark.parametrize("bits", [8, 4])
def test_bfplinear_int_format_on_tensor_cores(ops, cfg, bits, monkeypatch):
    """BFPLinear with sparsity_num_format='int' (inference): three-plane activations x integer weight grid on the bf16 tensor
    cores (2:4-compressed when the weight is pruned 2:4) against the fp64 contraction of the same fake-quantised operands,
    and against the reference's structure (fake-quant + F.linear)."""
    mode, first = cfg
    monkeypatch.setenv("BFP_INT_LINEAR", "tc")                    # opt-in: the default keeps the library GEMM (bit-identical logits)
    kw = dict(num_format="bfp", sparsity_num_format="int", rounding_mode="determ", epsilon=1e-8, mant_bits=bits, block_size=64,
              w_sparsity=mode != "none", N=2, M=4, first=first, sparsity_mode="structured" if mode == "none" else mode,
              sparsity_frac=0.5, device="cuda")
    torch.manual_seed(3)
    lin = ops.BFPLinear(1096, 520, bias=True, **dict(kw)).cuda()
    with torch.no_grad():
        lin.weight[3] = lin.weight[3].abs()                   # a non-negative row: xmin = 0, asymmetric grid
        lin.weight[4] = 0                                     # a dead row: scale 2 / maxq
    x = torch.randn(2, 77, 1096, device="cuda") * (torch.rand(1096, device="cuda") * 3 + 0.1)
    a = ops.unpack_bfp_args(dict(kw))
    with torch.no_grad():
        y = lin(x)
        assert lin._packed_w is not None and lin._packed_w[0][0] == "int" and lin._packed_w[1] is not None
        assert lin._packed_w[1].sparse == (mode == "structured")
        monkeypatch.setenv("BFP_LINEAR_PATH", "fakequant")
        y_fq = lin(x)
        monkeypatch.setenv("BFP_LINEAR_PATH", "tc")
    xq = ops.float_to_bfp_blocked(x, **a, identifier="in").double()
    wq = ops.float_to_bfp_blocked(lin.weight.detach(), **a, identifier="w").double()
    exact = xq @ wq.t() + lin.bias.detach().double()
    for got in (y, y_fq):
        assert got.dtype == torch.float32 and got.shape == (2, 77, 520)
        assert float((got.double() - exact).norm() / exact.norm()) <= 1e-5
    assert float((y.double() - exact).norm() / exact.norm()) <= 2e-6


def test_bfplinear_int_format_default_is_the_library_gemm(ops):
    """Without BFP_INT_LINEAR=tc the INT-format module keeps the reference's structure (fused quantiser + F.linear): its output is
    bit-identical to F.linear on the fake-quantised operands, which is what keeps whole-model logits equal to the reference's."""
    kw = dict(num_format="bfp", sparsity_num_format="int", rounding_mode="determ", epsilon=1e-8, mant_bits=8, block_size=64,
              w_sparsity=True, N=2, M=4, first="s", sparsity_mode="structured", device="cuda")
    torch.manual_seed(3)
    lin = ops.BFPLinear(512, 256, bias=True, **dict(kw)).cuda()
    x = torch.randn(4, 33, 512, device="cuda")
    a = ops.unpack_bfp_args(dict(kw))
    with torch.no_grad():
        y = lin(x)
        ref = torch.nn.functional.linear(ops.float_to_bfp_blocked(x, **a, identifier="in"),
                                         ops.float_to_bfp_blocked(lin.weight.detach(), **a, identifier="w"), lin.bias)
    assert lin._packed_w is None and torch.equal(y, ref)
