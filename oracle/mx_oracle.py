"""CPU oracle of the MX (OCP Microscaling) path behind the reference's `mx_layers.py`.  TEST INFRASTRUCTURE ONLY: imported by tests/,
never by the product.

**Parity unpinned.**  The arithmetic of `MXLinear` / `MXConv2d` / `MXMatmul` (/root/reference/src/transformers/bfp/mx_layers.py:18-21
imports `mx.Linear`, `mx.Conv2d`, `mx.matmul`) lives in microsoft/microxcaling, an un-vendored, un-pinned git dependency
(/root/reference/requirements_pip.txt:65) that is not installable here.  What IS in the reference tree pins part of it:
`formats.py:52-128` (element-format parameters: ebits, mbits, emax, max_norm -- restated in FORMATS below and compared with the
reference file by tests/test_mx_oracle.py when it is present) and `specs.py:30-66` (defaults: round 'nearest', shared_exp_method 'max',
bfloat_subnorms True).  The rest restates the library's published PyTorch emulation (mx/elemwise_ops.py `_round_mantissa`,
`_quantize_elemwise_core`, `_quantize_bfloat`; mx/mx_ops.py `_shared_exponents`, `_reshape_to_blocks`, `_quantize_mx`; mx/linear.py
`LinearFunction.forward`; mx/matmul.py `MatMulFunction.forward`; mx/convolution.py `ConvFunction.forward`) and the OCP Microscaling
Formats (MX) v1.0 specification, operation by operation in float32, as the library evaluates it on fp32 tensors.

Call sites that fix the configuration (bfp_util.py:29-36): w_elem_format / a_elem_format from the yaml, block_size = the BFP block
size, bfloat = 16, scale_bits = 8; everything else default -> every rounding is 'nearest' = round half AWAY from zero.
"""
import numpy as np

F32 = np.float32
FP32_EXPONENT_BIAS = 127
FP32_MIN_NORMAL = F32(2.0 ** (-FP32_EXPONENT_BIAS + 1))

# name -> (ebits, mbits, emax, max_norm): formats.py:86-123 (mbits counts the sign and the implicit bit)
FORMATS = {
    "int8": (0, 8, 0, 1.984375),
    "int4": (0, 4, 0, 1.75),
    "int2": (0, 2, 0, 1.0),
    "fp8_e5m2": (5, 4, 15, 57344.0),
    "fp8_e4m3": (4, 5, 8, 448.0),
    "fp6_e3m2": (3, 4, 4, 28.0),
    "fp6_e2m3": (2, 5, 2, 7.5),
    "fp4_e2m1": (2, 3, 2, 6.0),
    "fp4": (2, 3, 2, 6.0),
}
FORMAT_IDS = {"int8": 1, "int4": 2, "int2": 3, "fp8_e5m2": 4, "fp8_e4m3": 5, "fp6_e3m2": 6, "fp6_e2m3": 7, "fp4_e2m1": 8, "fp4": 8}   # formats.py:24-33


def _pow2(e):
    """2 ** e for a float32 array of integer-valued exponents (torch.pow(2, t) on fp32: exact, denormal results kept)."""
    e = np.asarray(e, dtype=F32)
    ok = np.isfinite(e)
    with np.errstate(over="ignore", under="ignore", invalid="ignore"):
        p = np.ldexp(np.ones_like(e), np.where(ok, e, 0).astype(np.int32)).astype(F32)
    return np.where(ok, p, np.where(np.isnan(e), F32(np.nan), np.where(e > 0, F32(np.inf), F32(0.0)))).astype(F32)


def round_mantissa(a, rnd):
    """mx/elemwise_ops.py _round_mantissa (no clamp)."""
    a = a.astype(F32)
    if rnd == "nearest":
        return (np.sign(a) * np.floor(np.abs(a) + F32(0.5))).astype(F32)
    if rnd == "floor":
        return (np.sign(a) * np.floor(np.abs(a))).astype(F32)
    if rnd == "even":
        absa = np.abs(a)
        mask = (np.mod(absa - F32(0.5), F32(2.0)) == 0).astype(F32)
        return (np.sign(a) * (np.floor(absa + F32(0.5)) - mask)).astype(F32)
    raise ValueError(f"unknown rounding {rnd!r}")


def quantize_elemwise_core(a, bits, exp_bits, max_norm, rnd="nearest", saturate_normals=False, allow_denorm=True):
    """mx/elemwise_ops.py _quantize_elemwise_core on an fp32 tensor."""
    a = np.asarray(a, dtype=F32)
    out = a.copy()
    with np.errstate(all="ignore"):
        if not allow_denorm and exp_bits > 0:
            min_norm = F32(2.0 ** (2 - 2 ** (exp_bits - 1)))
            out = (np.abs(out) >= min_norm).astype(F32) * out
        if exp_bits != 0:
            private_exp = np.floor(np.log2(np.abs(a) + (a == 0).astype(F32)).astype(F32))
            min_exp = -(2 ** (exp_bits - 1)) + 2
            private_exp = np.maximum(private_exp, F32(min_exp))
            out = (out / _pow2(private_exp)).astype(F32) * F32(2.0 ** (bits - 2))
        else:
            private_exp = None
            out = out * F32(2.0 ** (bits - 2))
        out = round_mantissa(out.astype(F32), rnd)
        if private_exp is not None:
            out = (out / F32(2.0 ** (bits - 2))).astype(F32) * _pow2(private_exp)
        else:
            out = out / F32(2.0 ** (bits - 2))
        out = out.astype(F32)
        if saturate_normals or exp_bits == 0:
            out = np.where(np.isnan(out), out, np.clip(out, F32(-max_norm), F32(max_norm)))
        else:
            out = np.where(np.abs(out) > F32(max_norm), np.sign(out) * F32(np.inf), out)
        out = np.where(a == F32(np.inf), F32(np.inf), out)
        out = np.where(a == F32(-np.inf), F32(-np.inf), out)
    return out.astype(F32)


def quantize_bfloat(a, bfloat=16, rnd="nearest", allow_denorm=True):
    """mx/elemwise_ops.py _quantize_bfloat: bfloatX = 8 exponent bits + sign + (X - 9) explicit mantissa bits; overflow -> Inf."""
    if bfloat == 0 or bfloat == 32:
        return np.asarray(a, dtype=F32)
    mbits = bfloat - 7
    max_norm = 2.0 ** 127 * float(2 ** (mbits - 1) - 1) / 2 ** (mbits - 2)          # formats.py:57-61 _get_max_norm(8, mbits)
    return quantize_elemwise_core(a, mbits, 8, max_norm, rnd, saturate_normals=False, allow_denorm=allow_denorm)


def shared_exponents(blocks):
    """mx/mx_ops.py _shared_exponents(method='max', ebits=0) over the last axis (kept)."""
    with np.errstate(all="ignore"):
        m = np.max(np.abs(blocks), axis=-1, keepdims=True)                     # NaN propagates like torch.max
        m = np.where(np.isnan(blocks).any(axis=-1, keepdims=True), F32(np.nan), m).astype(F32)
        return np.floor(np.log2(m + FP32_MIN_NORMAL * (m == 0).astype(F32)).astype(F32)).astype(F32)


def quantize_mx(a, elem_format, block_size=32, scale_bits=8, axis=-1, rnd="nearest", flush_fp32_subnorms=False):
    """mx/mx_ops.py _quantize_mx along one axis: zero-pad the axis to a multiple of block_size (block_size 0: the whole axis is one
    block), one shared power-of-two scale per block, elements in `elem_format`, result in float32 with the input's shape."""
    if elem_format is None:
        return np.asarray(a, dtype=F32)
    ebits, mbits, emax, max_norm = FORMATS[elem_format]
    assert scale_bits > 0
    a = np.moveaxis(np.asarray(a, dtype=F32), axis, -1)
    shape = a.shape
    k = shape[-1]
    bs = block_size if block_size > 0 else k
    kp = -(-k // bs) * bs
    if kp != k:
        a = np.concatenate([a, np.zeros(shape[:-1] + (kp - k,), F32)], axis=-1)
    blocks = a.reshape(shape[:-1] + (kp // bs, bs))
    with np.errstate(all="ignore"):
        se = shared_exponents(blocks)
        if flush_fp32_subnorms:
            blocks = blocks * (se > -FP32_EXPONENT_BIAS).astype(F32)
        se = se - F32(emax)
        scale_emax = 2 ** (scale_bits - 1) - 1
        se = np.where(se > scale_emax, F32(np.nan), se)
        se = np.where(se < -scale_emax, F32(-scale_emax), se).astype(F32)
        scale = _pow2(se)
        q = (blocks / scale).astype(F32)
        q = quantize_elemwise_core(q, mbits, ebits, max_norm, rnd, saturate_normals=True, allow_denorm=True)
        q = (q * scale).astype(F32)
    q = q.reshape(shape[:-1] + (kp,))[..., :k]
    return np.ascontiguousarray(np.moveaxis(q, -1, axis))


def mx_linear(x, w, bias, w_fmt, a_fmt, block_size=32, bfloat=16, scale_bits=8):
    """mx/linear.py LinearFunction.forward: bfloat round of x, w, bias -> MX along the contraction dim -> GEMM -> bfloat round ->
    (+ bias -> bfloat round).  The contraction is accumulated in float64 here (the library runs an fp32 GEMM: results agree to the
    fp32 summation error, then collapse onto the bfloat grid)."""
    bx, bw = quantize_bfloat(x, bfloat), quantize_bfloat(w, bfloat)
    qx = quantize_mx(bx, a_fmt, block_size, scale_bits, -1)
    qw = quantize_mx(bw, w_fmt, block_size, scale_bits, -1)
    y = (qx.astype(np.float64) @ qw.astype(np.float64).T).astype(F32)
    y = quantize_bfloat(y, bfloat)
    if bias is not None:
        y = quantize_bfloat((y + quantize_bfloat(bias, bfloat)).astype(F32), bfloat)
    return y, qx, qw


def mx_matmul(in1, in2, a_fmt, block_size=32, bfloat=16, scale_bits=8):
    """mx/matmul.py MatMulFunction.forward (mode 'aa'): both operands in the activation format, in1 blocked along its last axis,
    in2 along its second-to-last (the contraction dim of each)."""
    b1, b2 = quantize_bfloat(in1, bfloat), quantize_bfloat(in2, bfloat)
    q1 = quantize_mx(b1, a_fmt, block_size, scale_bits, -1)
    q2 = quantize_mx(b2, a_fmt, block_size, scale_bits, -2)
    y = np.matmul(q1.astype(np.float64), q2.astype(np.float64)).astype(F32)
    return quantize_bfloat(y, bfloat), q1, q2
