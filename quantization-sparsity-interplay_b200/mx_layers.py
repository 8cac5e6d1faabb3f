"""Host-side mirror of the reference's `transformers.bfp.mx_layers` (/root/reference/src/transformers/bfp/mx_layers.py:23-109):
`MXLinear`, `MXConv2d`, `MXMatmul` with the reference's constructor / call signatures, plus the two functions of microsoft/microxcaling
they rest on (`quantize_elemwise_op`, `quantize_mx_op`) and the spec helpers the reference vendors in specs.py (`apply_mx_specs`,
`finalize_mx_specs`).  microxcaling itself is not vendored by the reference and not installable here: the arithmetic is restated in
csrc/bfp_ocp_mx.cu (one fused CUDA pass) and, independently, in oracle/mx_oracle.py -- **parity unpinned** (DESIGN.md section 7.2).

B200 mapping: an MX block of 32 with an E8M0 scale is exactly the operand of `tcgen05.mma.kind::mxf8f6f4.block_scale`, so for the
element formats whose values are E4M3 numbers (fp8_e4m3, fp6_e3m2, fp6_e2m3, fp4_e2m1, int4, int2) the linear runs on the block-scaled
tensor-core kind with the scales applied in hardware (bfp_ocp_mx_pack + bfp_gemm_mx_round, the bfloat output rounding fused into the
epilogue); int8 / fp8_e5m2 (and block sizes the hardware scale granularity does not cover) run on the exact-bf16 kinds -- every MX
value has at most 8 significant bits -- dense or 2:4-sparse.  There is no CPU fallback.
"""
import ctypes
import math

import torch
import torch.nn.functional as F

from . import _lib
from .bfp_ops import (_DT, _on, _stream, _structured_N_M_sparsity, _unstructured_sparsity, SparseBF16, PackedMX, bfp_linear_bf16,
                      bfp_linear_bf16_sp, compress_2to4_bf16, _is_patch_embedding)

# formats.py:24-33 (ElemFormat values) -- the ids of include/bfp_b200.h BFP_MX_*
ELEM_FORMATS = {"int8": 1, "int4": 2, "int2": 3, "fp8_e5m2": 4, "fp8_e4m3": 5, "fp6_e3m2": 6, "fp6_e2m3": 7, "fp4": 8, "fp4_e2m1": 8}
_E4M3_SUBSET = {2, 3, 5, 6, 7, 8}          # formats whose every value is an E4M3 number (the block-scaled tensor-core operand type)

# specs.py:30-66: every MX option and its default
_SPEC_DEFAULTS = {
    "scale_bits": 0, "w_elem_format": None, "a_elem_format": None, "w_elem_format_bp": None, "a_elem_format_bp_ex": None,
    "a_elem_format_bp_os": None, "mx_flush_fp32_subnorms": False, "shared_exp_method": "max", "block_size": 0, "bfloat": 0, "fp": 0,
    "bfloat_subnorms": True, "quantize_backprop": True, "round": "nearest", "round_m": "nearest", "round_weight": "nearest",
    "round_output": "nearest", "round_grad_weight": "nearest", "round_grad_input": "nearest", "round_mx_output": "nearest",
    "round_mx_input_grad_input": "nearest", "round_mx_weight_grad_input": "nearest", "round_mx_grad_output_grad_input": "nearest",
    "round_mx_input_grad_weight": "nearest", "round_mx_grad_output_grad_weight": "nearest", "softmax_exp2": False, "vec_use_exp2": False,
    "vec_use_recip": False, "custom_cuda": False,
}


def apply_mx_specs(mx_specs, default_mx_specs=None):
    """specs.py:172-192: the defaults overlaid with the non-None entries of `mx_specs`; unknown keys raise KeyError."""
    out = dict(_SPEC_DEFAULTS) if not default_mx_specs else default_mx_specs
    if not mx_specs:
        return out
    for k, v in mx_specs.items():
        if v is not None:
            if k not in out:
                raise KeyError(f"Unknown key '{k}' passed to mx specs")
            out[k] = v
    return out


def finalize_mx_specs(specs, early_exit=True):
    """specs.py:236-279: None when nothing quantises (no element format, no bfloat / fp), else the back-propagation formats default
    to the forward ones."""
    if early_exit and not any(specs.get(k, 0) for k in ("w_elem_format", "a_elem_format", "w_elem_format_bp", "a_elem_format_bp_os",
                                                        "a_elem_format_bp_ex", "bfloat", "fp")):
        return None
    for dst, src in (("w_elem_format_bp", "w_elem_format"), ("a_elem_format_bp_os", "a_elem_format"), ("a_elem_format_bp_ex", "a_elem_format")):
        if specs.get(dst) is None and src in specs:
            specs[dst] = specs[src]
    return apply_mx_specs(specs, dict(_SPEC_DEFAULTS))


def _format_id(elem_format):
    if elem_format is None:
        return None
    if isinstance(elem_format, int):
        return elem_format
    name = getattr(elem_format, "name", elem_format)
    try:
        return ELEM_FORMATS[str(name).lower()]
    except KeyError:
        raise Exception("Undefined elem format", elem_format)          # formats.py:49


def _need_nearest(rnd):
    if rnd not in (None, "nearest"):
        raise NotImplementedError(f"MX rounding mode {rnd!r}: only 'nearest' (every default of specs.py:54-65) is built")


def _bfloat_of(mx_specs):
    if mx_specs is None:
        return 0
    if mx_specs["bfloat"] > 0 and mx_specs["fp"] > 0:
        raise ValueError("Cannot set both [bfloat] and [fp] in mx_specs.")
    if mx_specs["fp"] > 0:
        raise NotImplementedError("fpX elementwise format (specs 'fp'): the reference only sets 'bfloat' (bfp_util.py:34)")
    b = int(mx_specs["bfloat"])
    if 0 < b <= 9:
        raise ValueError("Cannot set [bfloat] <= 9")
    if b and not mx_specs.get("bfloat_subnorms", True):
        raise NotImplementedError("bfloat_subnorms=False")
    return 0 if b == 32 else b


def _check_tensor(t):
    if t.dtype not in _DT:
        raise TypeError(f"bfp_b200 supports float32 / float16 / bfloat16 tensors, got {t.dtype}")
    if not t.is_cuda:
        raise ValueError("the MX path needs a CUDA tensor (no CPU fallback)")


def quantize_elemwise_op(A, mx_specs, round=None):
    """mx/elemwise_ops.py quantize_elemwise_op: bfloatX rounding (half away from zero) of every element; identity when mx_specs is
    None or sets no elementwise format."""
    if mx_specs is None:
        return A
    _need_nearest(round if round is not None else mx_specs["round"])
    b = _bfloat_of(mx_specs)
    if b == 0:
        return A
    _check_tensor(A)
    src = A.detach().contiguous()
    out = torch.empty_like(src)
    if src.numel():
        with _on(src.device):
            _lib.check(_lib.lib().bfp_bfloat_round(src.data_ptr(), out.data_ptr(), None, src.numel(), 0, _DT[src.dtype], b, _stream()))
    return out


def _mx_quantize_last(src, fmt, block_size, scale_bits, bfloat, flush, out_kind=0):
    """bfloat rounding + quantize_mx along the last dim of a contiguous tensor in ONE kernel.  out_kind 0: same dtype / shape;
    1: exact-bf16 GEMM operand [rows, Kp] (Kp = K rounded up to 8)."""
    K = src.shape[-1] if src.dim() else 1
    rows = src.numel() // K if K else 0
    if out_kind == 0:
        out, ld = torch.empty_like(src), K
    else:
        ld = -(-K // 8) * 8
        out = (torch.empty if ld == K else torch.zeros)((rows, ld), dtype=torch.bfloat16, device=src.device)
    if rows and K:
        with _on(src.device):
            _lib.check(_lib.lib().bfp_ocp_mx_quantize(src.data_ptr(), out.data_ptr(), rows, K, _DT[src.dtype], out_kind, ld, int(block_size), fmt,
                                                      int(scale_bits), int(bfloat), int(bool(flush)), _stream()))
    return out


def quantize_mx_op(A, mx_specs, elem_format=None, block_size=32, axes=None, round="nearest", expand_and_reshape=False):
    """mx/mx_ops.py quantize_mx_op: MX fake-quantisation of `A` along ONE axis (the library's layers only ever pass one); the block
    size and scale bits come from mx_specs like the library's.  Straight-through in autograd."""
    fmt = _format_id(elem_format)
    if fmt is None:
        return A
    _need_nearest(round)
    if mx_specs["shared_exp_method"] != "max":
        raise NotImplementedError("shared_exp_method other than 'max'")
    axes = [axes] if isinstance(axes, int) else list(axes)
    if len(axes) != 1:
        raise NotImplementedError("MX quantisation along several axes at once")
    _check_tensor(A)
    ax = axes[0] % A.dim()
    moved = A.detach().movedim(ax, -1).contiguous()
    q = _mx_quantize_last(moved, fmt, mx_specs["block_size"], mx_specs["scale_bits"], 0, mx_specs["mx_flush_fp32_subnorms"])
    q = q.movedim(-1, ax)
    if torch.is_grad_enabled() and A.requires_grad:
        return A + (q - A).detach()
    return q


# ---------------------------------------------------------------------------------------------------------------
# the linear: mx/linear.py LinearFunction
# ---------------------------------------------------------------------------------------------------------------
def _block_scaled_ok(fmt_a, fmt_w, mx_specs, K, x_dtype, w_dtype, N):
    vec = 4 if (x_dtype == torch.float32 and w_dtype == torch.float32) else 8
    return (fmt_a in _E4M3_SUBSET and fmt_w in _E4M3_SUBSET and mx_specs["block_size"] in (32, 64, 128) and mx_specs["scale_bits"] == 8
            and K % (32 * vec) == 0 and K % 128 == 0 and N % 4 == 0)


def _pack_block_scaled(src2d, fmt, tile_rows, mx_specs, bfloat):
    rows, K = src2d.shape
    L = _lib.lib()
    sfb = ctypes.c_int64()
    _lib.check(L.bfp_mx_layout(rows, K, tile_rows, 0, None, ctypes.byref(sfb)))
    vals = torch.empty((rows, K), dtype=torch.uint8, device=src2d.device)
    sf = (torch.zeros if rows % tile_rows else torch.empty)(max(sfb.value, 16), dtype=torch.uint8, device=src2d.device)
    with _on(src2d.device):
        _lib.check(L.bfp_ocp_mx_pack(src2d.data_ptr(), vals.data_ptr(), sf.data_ptr(), rows, K, _DT[src2d.dtype], tile_rows, int(mx_specs["block_size"]), fmt,
                                     int(mx_specs["scale_bits"]), bfloat, int(bool(mx_specs["mx_flush_fp32_subnorms"])), _stream()))
    return PackedMX(vals, sf, rows, K, tile_rows, int(mx_specs["block_size"]), 0, False)


def _weight_tile(N):
    return 240 if N >= 240 else 128


class _PackedWeightCache:
    """Packed forms of a module's weight, with the protocol of BFPLinear's cache (bfp_ops.BFPLinear._packed_weight): keyed on the
    parameter's storage, version, shape and dtype; never served while the module trains with a trainable weight (in-place optimiser
    writes through `.data` do not bump the version); dropped by `invalidate()` -- the modules call it from `_apply` / `load_state_dict`
    and expose it as `invalidate_packed()` for callers that rewrite the weight through `.data` in eval mode.  BFP_WEIGHT_CACHE=0
    disables caching; =verify adds a checksum of the whole weight to the key (one extra read and a host sync per forward)."""

    def __init__(self):
        self.entries = {}

    def invalidate(self):
        self.entries = {}

    def get(self, module, w, kind, build):
        import os
        mode = os.environ.get("BFP_WEIGHT_CACHE", "1")
        if mode == "0" or (module is not None and module.training and w.requires_grad):
            return build()
        key = (w.data_ptr(), w._version, tuple(w.shape), w.dtype, w.device)
        if mode == "verify":
            key = key + (int(w.detach().view(torch.int16 if w.element_size() == 2 else torch.int32).sum(dtype=torch.int64).item()),)
        hit = self.entries.get(kind)
        if hit is None or hit[0] != key:
            self.entries = {k: v for k, v in self.entries.items() if v[0] == key}          # a changed weight drops every stale form
            hit = (key, build())
            self.entries[kind] = hit
        return hit[1]


def mx_linear_forward(x, w, bias, mx_specs, cache=None, module=None, prefer_sparse=False):
    """Inference forward of mx/linear.py LinearFunction on the tensor cores:
        y = rb(rb(Q_a(rb(x)) . Q_w(rb(w))^T) + rb(bias)),  rb = bfloat rounding, Q = MX along the contraction dim."""
    fmt_a, fmt_w = _format_id(mx_specs["a_elem_format"]), _format_id(mx_specs["w_elem_format"])
    bfloat = _bfloat_of(mx_specs)
    for r in ("round_output", "round_weight", "round_mx_output"):
        _need_nearest(mx_specs[r])
    _check_tensor(x)
    _check_tensor(w)
    K, N = w.shape[1], w.shape[0]
    src = x.detach().contiguous().view(-1, K)
    out_shape = tuple(x.shape[:-1]) + (N,)
    wd = w.detach().contiguous()
    b32 = bias.detach().to(torch.float32).contiguous() if bias is not None else None
    flush = mx_specs["mx_flush_fp32_subnorms"]
    get = (lambda kind, build: cache.get(module, w, kind, build)) if cache is not None else (lambda kind, build: build())
    if fmt_a is not None and fmt_w is not None and _block_scaled_ok(fmt_a, fmt_w, mx_specs, K, src.dtype, wd.dtype, N) and src.shape[0]:
        tile = _weight_tile(N)
        wp = get(("bs", fmt_w, tile, bfloat), lambda: _pack_block_scaled(wd, fmt_w, tile, mx_specs, bfloat))
        xp = _pack_block_scaled(src, fmt_a, 128, mx_specs, bfloat)
        out = torch.empty((src.shape[0], N), dtype=torch.float32, device=src.device)
        with _on(src.device):
            _lib.check(_lib.lib().bfp_gemm_mx_round(xp.vals.data_ptr(), xp.sf.data_ptr(), wp.vals.data_ptr(), wp.sf.data_ptr(), wp.tile_rows, 0,
                                                    b32.data_ptr() if b32 is not None else None, out.data_ptr(), src.shape[0], N, K, bfloat or 0, _stream()))
        return out.view(out_shape).to(x.dtype)
    # exact-bf16 operands: the MX values (or, with no element format, the bfloat-rounded ones when bfloat <= 16) are bf16 numbers
    if (fmt_a is None or fmt_w is None) and not (0 < bfloat <= 16):
        raise NotImplementedError("MX linear without element formats needs bfloat <= 16 to run on the bf16 tensor cores")

    def operand(t2d, fmt):
        if fmt is None:
            return _pad8(quantize_elemwise_op(t2d, mx_specs).to(torch.bfloat16))
        return _mx_quantize_last(t2d, fmt, mx_specs["block_size"], mx_specs["scale_bits"], bfloat, flush, out_kind=1)

    xb = operand(src, fmt_a)
    if prefer_sparse and N % 4 == 0:
        ws = get(("sp", fmt_w, bfloat), lambda: _try_compress(operand(wd, fmt_w)))
        if isinstance(ws, SparseBF16):
            y = bfp_linear_bf16_sp(xb, ws, None)
        else:
            y = bfp_linear_bf16(xb, ws, None)
    else:
        wb = get(("bf16", fmt_w, bfloat), lambda: operand(wd, fmt_w))
        y = bfp_linear_bf16(xb, wb, None)
    if bfloat or b32 is not None:
        with _on(y.device):
            _lib.check(_lib.lib().bfp_bfloat_round(y.data_ptr(), y.data_ptr(), b32.data_ptr() if b32 is not None else None, y.numel(), N, _lib.DT_F32,
                                                   bfloat, _stream()))
    return y.view(out_shape).to(x.dtype)


def _pad8(t2d):
    K = t2d.shape[1]
    Kp = -(-K // 8) * 8
    if Kp == K:
        return t2d.contiguous()
    out = torch.zeros((t2d.shape[0], Kp), dtype=t2d.dtype, device=t2d.device)
    out[:, :K] = t2d
    return out


def _try_compress(wb):
    try:
        return compress_2to4_bf16(wb, check=True)
    except ValueError:
        return wb


def _mx_operand(t2d, axis, fmt, sp):
    """Exact-bf16 GEMM operand of the MX quantisation of a 2-D tensor along `axis`: [other dim, Kp] with the quantised axis last
    (K-major, zero padded to a multiple of 8) -- for axis 0 the transposed contiguous copy the quantiser needs anyway IS the operand."""
    src = (t2d.detach().t() if axis == 0 else t2d.detach()).contiguous()
    return _mx_quantize_last(src, fmt, sp["block_size"], sp["scale_bits"], 0, sp["mx_flush_fp32_subnorms"], out_kind=1)


def _round_inplace(y, sp):
    b = _bfloat_of(sp)
    if b and y.numel():
        with _on(y.device):
            _lib.check(_lib.lib().bfp_bfloat_round(y.data_ptr(), y.data_ptr(), None, y.numel(), 0, _DT[y.dtype], b, _stream()))
    return y


class _MXLinearFunction(torch.autograd.Function):
    """Training path (mx/linear.py LinearFunction forward + backward with quantize_backprop).  The forward is the inference forward;
    the backward quantises along the axes the library uses -- weight gradient: both factors along the token dim; input gradient: the
    weight along out_features, the output gradient along its last dim -- straight into exact-bf16 operands, and both contractions
    run on the tcgen05 bf16 kind (fp32 accumulation), followed by the library's bfloat rounding of each gradient."""

    @staticmethod
    def forward(ctx, x, w, bias, mx_specs):
        ctx.has_bias, ctx.mx_specs = bias is not None, mx_specs
        if mx_specs["quantize_backprop"]:
            ctx.save_for_backward(quantize_elemwise_op(x, mx_specs), quantize_elemwise_op(w, mx_specs))
        else:
            ctx.save_for_backward(x, w)
        return mx_linear_forward(x, w, bias, mx_specs)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        sp = ctx.mx_specs if ctx.mx_specs["quantize_backprop"] else None
        out_dim, in_dim = w.shape
        gy = quantize_elemwise_op(gy.contiguous(), sp)
        g2, x2 = gy.reshape(-1, out_dim), x.reshape(-1, in_dim)
        f_ex = _format_id(sp["a_elem_format_bp_ex"]) if sp is not None else None
        f_w, f_os = (_format_id(sp["w_elem_format_bp"]), _format_id(sp["a_elem_format_bp_os"])) if sp is not None else (None, None)
        if sp is None or f_ex is None or f_w is None or f_os is None or not (g2.dtype == x2.dtype == w.dtype == torch.float32):
            # no element format for some operand (or half-precision modules): the library's own structure on fake-quantised tensors
            qx = quantize_mx_op(x2, sp, sp["a_elem_format_bp_ex"], axes=[0]) if sp is not None else x2
            qg = quantize_mx_op(g2, sp, sp["a_elem_format_bp_ex"], axes=[0]) if sp is not None else g2
            gw = quantize_elemwise_op(qg.t() @ qx, sp)
            qw = quantize_mx_op(w, sp, sp["w_elem_format_bp"], axes=[0]) if sp is not None else w
            qo = quantize_mx_op(gy, sp, sp["a_elem_format_bp_os"], axes=[-1]) if sp is not None else gy
            gx = quantize_elemwise_op(qo @ qw, sp)
        else:
            gw = _round_inplace(bfp_linear_bf16(_mx_operand(g2, 0, f_ex, sp), _mx_operand(x2, 0, f_ex, sp)), sp)                 # [out, in]
            gx = _round_inplace(bfp_linear_bf16(_mx_operand(g2, 1, f_os, sp), _mx_operand(w, 0, f_w, sp)), sp).view(gy.shape[:-1] + (in_dim,))
        gb = None
        if ctx.has_bias:
            gb = quantize_elemwise_op(g2.sum(0), sp)
        return gx, gw, gb, None


def _sparsify_weight_once(module):
    """mx_layers.py:49-56 / :91-98: on the first forward the weight Parameter is REPLACED by its pruned copy."""
    if module.sparsity and not module.sparsity_init:
        if module.sparsity_mode == "structured":
            module.weight = torch.nn.Parameter(_structured_N_M_sparsity(module.weight, module.device, module.N, module.M))
        else:
            module.weight = torch.nn.Parameter(_unstructured_sparsity(module.weight, module.device, module.sparsity_frac))
        module.sparsity_init = True


class MXLinear(torch.nn.Linear):
    """mx_layers.py:23-57"""

    def __init__(self, in_features, out_features, bias=True, mx_specs=None, name=None, sparsity=False, device=None, sparsity_mode="structured",
                 sparsity_frac=0.0, N=0, M=0):
        mx_specs = finalize_mx_specs(apply_mx_specs(mx_specs))
        super().__init__(in_features, out_features, bias)
        assert (sparsity_mode in ["structured", "unstructured"])
        self.mx_none = mx_specs is None
        self.mx_specs, self.name = mx_specs, name
        self.sparsity, self.device, self.sparsity_mode, self.sparsity_frac, self.N, self.M = sparsity, device, sparsity_mode, sparsity_frac, N, M
        self.sparsity_init = False
        self._packed = _PackedWeightCache()

    def invalidate_packed(self):
        """Drops the cached packed weight (call after modifying the weight through `.data` outside training)."""
        self._packed.invalidate()

    def _apply(self, fn, *args, **kwargs):
        if "_packed" in self.__dict__:
            self._packed.invalidate()
        return super()._apply(fn, *args, **kwargs)

    def _load_from_state_dict(self, *args, **kwargs):
        self._packed.invalidate()
        return super()._load_from_state_dict(*args, **kwargs)

    def forward(self, inputs):
        _sparsify_weight_once(self)
        if self.mx_none:
            return F.linear(inputs, self.weight, self.bias)
        if torch.is_grad_enabled() and (inputs.requires_grad or self.weight.requires_grad or (self.bias is not None and self.bias.requires_grad)):
            return _MXLinearFunction.apply(inputs, self.weight, self.bias, self.mx_specs)
        two_four = self.sparsity and self.sparsity_mode == "structured" and (self.N, self.M) == (2, 4)
        return mx_linear_forward(inputs, self.weight, self.bias, self.mx_specs, self._packed, self, prefer_sparse=two_four)


class MXConv2d(torch.nn.Conv2d):
    """mx_layers.py:59-99 over mx/convolution.py ConvFunction.forward: bfloat rounding, MX along the CHANNEL axis of input and weight
    (axes=[1]), the convolution on the quantised tensors, bfloat rounding, bias, bfloat rounding.  The quantiser is one fused CUDA pass
    per tensor over a channels-last view; patch embeddings (kernel == stride: ViT) run as im2col-by-permutation + a tcgen05 GEMM, any
    other convolution is the library's on the quantised tensors (exact: MX values fit TF32)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True, mx_specs=None, name=None,
                 sparsity=False, device=None, sparsity_mode="structured", sparsity_frac=0.0, N=0, M=0):
        mx_specs = finalize_mx_specs(apply_mx_specs(mx_specs))
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias)
        assert (sparsity_mode in ["structured", "unstructured"])
        self.mx_none = mx_specs is None
        self.mx_specs, self.name = mx_specs, name
        self.sparsity, self.device, self.sparsity_mode, self.sparsity_frac, self.N, self.M = sparsity, device, sparsity_mode, sparsity_frac, N, M
        self.sparsity_init = False

    def forward(self, inputs):
        _sparsify_weight_once(self)
        if self.mx_none:
            return super().forward(inputs)
        sp = self.mx_specs
        bfloat = _bfloat_of(sp)

        def q(t, fmt):
            fmt = _format_id(fmt)
            cl = t.detach().permute(0, 2, 3, 1).contiguous()                      # channels last: the MX axis becomes the last dim
            if fmt is None:
                return quantize_elemwise_op(cl, sp).permute(0, 3, 1, 2)
            return _mx_quantize_last(cl, fmt, sp["block_size"], sp["scale_bits"], bfloat, sp["mx_flush_fp32_subnorms"]).permute(0, 3, 1, 2)

        xq, wq = q(inputs, sp["a_elem_format"]), q(self.weight, sp["w_elem_format"])
        O, C, kh, kw = self.weight.shape
        if (self.groups == 1 and _is_patch_embedding(inputs, self.weight, self.stride, self.padding, self.dilation) and (C * kh * kw) % 8 == 0
                and inputs.dtype == torch.float32 and self.weight.dtype == torch.float32 and not (torch.is_grad_enabled() and (inputs.requires_grad or self.weight.requires_grad))):
            # patch embedding (kernel == stride, the ViT case): the windows tile the image, so im2col is ONE permuting copy of the quantised
            # input (exact in bf16: <= 8 significant bits) and the convolution is a tcgen05 GEMM; the result is a channels-last view
            B, _, H, W = inputs.shape
            Ho, Wo = H // kh, W // kw
            a = xq.reshape(B, C, Ho, kh, Wo, kw).permute(0, 2, 4, 1, 3, 5).reshape(B * Ho * Wo, C * kh * kw).to(torch.bfloat16)
            y = bfp_linear_bf16(a, wq.reshape(O, C * kh * kw).to(torch.bfloat16), None)                      # [B * L, O] fp32
            b32 = self.bias.detach().to(torch.float32).contiguous() if self.bias is not None else None
            with _on(y.device):
                _lib.check(_lib.lib().bfp_bfloat_round(y.data_ptr(), y.data_ptr(), b32.data_ptr() if b32 is not None else None, y.numel(), O, _lib.DT_F32,
                                                       bfloat, _stream()))
            return y.view(B, Ho, Wo, O).permute(0, 3, 1, 2)
        y = F.conv2d(xq, wq, None, self.stride, self.padding, self.dilation, self.groups)
        y = quantize_elemwise_op(y, sp)
        if self.bias is not None:
            y = quantize_elemwise_op(y + quantize_elemwise_op(self.bias, sp).view(1, -1, 1, 1), sp)
        return y


def MXMatmul(in1, in2, mx_specs=None, sparsity=False, sparsity_mode="unstructured", device=None, N=0, M=0, sparsity_frac=0):
    """mx_layers.py:101-109 over mx/matmul.py MatMulFunction.forward (mode 'aa'): the second operand is pruned along its contraction
    dim, both operands are bfloat-rounded and MX-quantised in the activation format along the contraction dim, the product is
    bfloat-rounded."""
    assert (sparsity_mode in ["structured", "unstructured"])
    if sparsity:
        if sparsity_mode == "structured":
            in2 = torch.transpose(_structured_N_M_sparsity(torch.transpose(in2, -1, -2), device, N, M), -1, -2)
        else:
            in2 = torch.transpose(_unstructured_sparsity(torch.transpose(in2, -1, -2), device, sparsity_frac), -1, -2)
    sp = finalize_mx_specs(apply_mx_specs(mx_specs)) if mx_specs is not None else None
    if sp is None:
        return torch.matmul(in1, in2)
    fmt = _format_id(sp["a_elem_format"])
    bfloat = _bfloat_of(sp)
    _check_tensor(in1)
    _check_tensor(in2)
    K, Nn, Mm = in1.shape[-1], in2.shape[-1], in1.shape[-2]
    batch = torch.broadcast_shapes(in1.shape[:-2], in2.shape[:-2])
    nb = math.prod(batch)
    flush = sp["mx_flush_fp32_subnorms"]
    if fmt is None and not (0 < bfloat <= 16):
        raise NotImplementedError("MX matmul without an element format needs bfloat <= 16")

    def operand(t):                       # [..., rows, K] -> exact bf16 [prod(...) * rows, Kp]
        t = t.detach().contiguous()
        if fmt is None:
            return _pad8(quantize_elemwise_op(t.view(-1, K), sp).to(torch.bfloat16))
        return _mx_quantize_last(t.view(-1, K), fmt, sp["block_size"], sp["scale_bits"], bfloat, flush, out_kind=1)

    a = operand(in1)
    b = operand(in2.transpose(-1, -2))
    Kp = a.shape[-1]
    a = a.view(tuple(in1.shape[:-2]) + (Mm, Kp)).expand(batch + (Mm, Kp)).reshape(nb, Mm, Kp).contiguous()
    b = b.view(tuple(in2.shape[:-2]) + (Nn, Kp)).expand(batch + (Nn, Kp)).reshape(nb, Nn, Kp).contiguous()
    out = torch.empty((nb, Mm, Nn), dtype=torch.float32, device=in1.device)
    L = _lib.lib()
    with _on(in1.device):
        st = _stream()
        if Nn % 4 == 0 and nb > 1:
            _lib.check(L.bfp_gemm_bf16_batched(a.data_ptr(), b.data_ptr(), out.data_ptr(), _lib.DT_F32, nb, Mm, Nn, Kp, st))
        else:
            for i in range(nb):
                _lib.check(L.bfp_gemm_bf16(a[i].data_ptr(), b[i].data_ptr(), None, out[i].data_ptr(), Mm, Nn, Kp, st))
        if bfloat:
            _lib.check(L.bfp_bfloat_round(out.data_ptr(), out.data_ptr(), None, out.numel(), 0, _lib.DT_F32, bfloat, st))
    return out.view(batch + (Mm, Nn)).to(in1.dtype)
