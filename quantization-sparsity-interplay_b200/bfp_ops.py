"""Host-side mirror of the reference operator module `transformers.bfp.bfp_ops`
(/root/reference/src/transformers/bfp/bfp_ops.py): the same public names, argument lists, defaults and error
behaviour, so the reference's patched OPT / LLaMA / ViT models (modeling_opt.py:42,162-176, modeling_llama.py:65,
225-237,305-319, modeling_vit.py:41,168-215) run on it unchanged -- see `qsi_b200.install_as_reference_module()`.

What differs is where the work happens: every quantise / sparsify call is ONE hand-written sm_100a kernel behind the
C ABI (include/bfp_b200.h) instead of ~25 eager torch kernels.  torch is used here for device memory, streams and
autograd plumbing only.  There is no CPU fallback: CPU tensors are processed on the GPU through the host-buffer entry
point (bfp_quantize_host), and a missing library or GPU raises.

Semantics kept from the reference (line numbers refer to the reference file):
  * returns a new tensor of the input's shape; the fp32 format without sparsity returns the input object itself (:105-106,
    :101-102); dtype = input dtype for 'determ', fp32 for 'stoc' on half inputs (:22-23)
  * `first == 's'` sparsifies then quantises, ANY other value quantises then sparsifies (:141-149)
  * `sparsity_num_format` selects the number format, `num_format` must be 'bfp' (:129-130)
  * AssertionError / ValueError / NotImplementedError in the same places (:27, :62, :74, :100, :122, :129-130, :268, :287)
  * `device` is advisory: the tensor's own device is used
  * N:M ties follow the torch.topk backend the reference would have used for that tensor: torch-CUDA's rule (smallest
    by (|v|, index)) for CUDA tensors, torch-CPU's rule for CPU tensors (2:4 only); override with BFP_TIE_RULE=cuda|cpu.
"""
import ctypes
import math
import os
import weakref

import torch
import torch.nn.functional as F

from . import _lib


# ---------------------------------------------------------------------------------------------------------------
# launch plumbing: the raw stream handle and a device guard that costs nothing when the tensor already lives on the
# current device (torch.cuda.current_stream() / torch.cuda.device() are ~13 us and ~5 us of Python per call).
# ---------------------------------------------------------------------------------------------------------------
def _stream(device=None):
    idx = torch.cuda.current_device() if device is None or device.index is None else device.index
    return torch._C._cuda_getCurrentRawStream(idx)


class _NoGuard:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def _on(device):
    if device.index is None or device.index == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(device)


_DT = {torch.float32: _lib.DT_F32, torch.float16: _lib.DT_F16, torch.bfloat16: _lib.DT_BF16}


class rounding_modes:
    """bfp_ops.py:16-18"""
    STOC, DETERM = 'stoc', 'determ'
    modes = [STOC, DETERM]


# ---------------------------------------------------------------------------------------------------------------
# Philox stream for stochastic rounding: key = torch's seed (so torch.manual_seed makes runs reproducible),
# offset = number of stochastic calls made since that seed was first seen.
# ---------------------------------------------------------------------------------------------------------------
class _PhiloxState:
    seed = None
    calls = 0

    @classmethod
    def next(cls):
        if torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
            # the (seed, offset) pair is baked into the kernel arguments: a replayed graph would reuse the same uniforms every step
            raise RuntimeError("bfp_b200: stochastic rounding cannot be captured into a CUDA graph (its Philox offset advances on the host); "
                               "capture with rounding_mode='determ' or run the stochastic calls eagerly")
        s = torch.initial_seed() & 0xFFFFFFFFFFFFFFFF
        if s != cls.seed:
            cls.seed, cls.calls = s, 0
        # ranks of one job draw from disjoint streams: the rank rides in the top bits of the 64-bit offset (sharded weights and
        # replicated activations then never share uniforms across GPUs)
        off = cls.calls | (int(os.environ.get("RANK", "0")) & 0xFFFF) << 48
        cls.calls += 1
        return s, off


def _tie_rule(t, N, M):
    env = os.environ.get("BFP_TIE_RULE", "")
    if env == "cuda":
        return _lib.TIE_TORCH_CUDA
    if env == "cpu":
        return _lib.TIE_TORCH_CPU
    if (not t.is_cuda) and N == 2 and M == 4:
        return _lib.TIE_TORCH_CPU
    return _lib.TIE_TORCH_CUDA


def _rounding_code(rounding_mode):
    if rounding_mode == rounding_modes.DETERM:
        return _lib.ROUND_NEAREST
    if rounding_mode == rounding_modes.STOC:
        return _lib.ROUND_STOCHASTIC
    raise NotImplementedError("Rounding mode %s is not implemented", rounding_mode)      # bfp_ops.py:27


def _fused(t, order, block_size=0, mant_bits=0, epsilon=0.0, rounding_mode=rounding_modes.DETERM, N=0, M=0,
           philox=None):
    """One launch of the fused kernel on `t` viewed as [rows, K] (K = last dim).  Returns a new tensor."""
    if t.dim() == 0:
        raise IndexError("tuple index out of range")            # the reference indexes t.shape[-1]
    if t.dtype not in _DT:
        raise TypeError(f"bfp_b200 supports float32 / float16 / bfloat16 tensors, got {t.dtype}")
    rounding = _rounding_code(rounding_mode) if order != _lib.ORDER_SPARSIFY_ONLY else _lib.ROUND_NEAREST
    stoc = rounding == _lib.ROUND_STOCHASTIC
    src = t.detach().contiguous()
    out_dtype = torch.float32 if stoc else src.dtype
    out = torch.empty(src.shape, dtype=out_dtype, device=src.device,
                      pin_memory=(not src.is_cuda) and src.is_pinned())   # pinned in -> pinned out (full-speed D2H)
    K = src.shape[-1]
    rows = src.numel() // K if K else 0
    if src.numel() == 0:
        return out
    seed, offset = (philox if philox is not None else _PhiloxState.next()) if stoc else (0, 0)
    tie = _tie_rule(src, N, M)
    L = _lib.lib()
    if src.is_cuda:
        with _on(src.device):
            stream = _stream()
            rc = L.bfp_quantize(src.data_ptr(), out.data_ptr(), rows, K, _DT[src.dtype], _DT[out_dtype],
                                int(block_size), int(mant_bits), float(epsilon), rounding, seed, offset, int(N), int(M),
                                order, tie, stream)
    else:
        rc = L.bfp_quantize_host(src.data_ptr(), out.data_ptr(), rows, K, _DT[src.dtype], _DT[out_dtype],
                                 int(block_size), int(mant_bits), float(epsilon), rounding, seed, offset, int(N), int(M),
                                 order, tie)
    _lib.check(rc)
    return out


# ---------------------------------------------------------------------------------------------------------------
# bfp_ops.py:20-59: rounding, exponent, block conversion
# ---------------------------------------------------------------------------------------------------------------
def round_tensor(t, mode, device):
    """bfp_ops.py:20-27.  Stand-alone helper (the fused kernel rounds in registers and never calls this)."""
    if mode == rounding_modes.STOC:
        sampled = torch.rand(t.shape, device=t.device) - 0.5
        return sampled.add_(t).round()
    elif mode == rounding_modes.DETERM:
        return t.round()
    raise NotImplementedError("Rounding mode %s is not implemented", mode)


def get_exponent(t, epsilon):
    """bfp_ops.py:29-33: per row of the 2-D [nblk, B] view, ceil(log2(max|t| + eps)) in t.dtype; shape [nblk, 1]."""
    if t.dtype not in _DT:
        raise TypeError(f"unsupported dtype {t.dtype}")
    src = t.detach().contiguous()
    nblk, B = src.shape
    dev_src = src if src.is_cuda else src.cuda()
    e = torch.empty((nblk, 1), dtype=torch.float32, device=dev_src.device)
    if nblk and B:
        with _on(dev_src.device):
            _lib.check(_lib.lib().bfp_block_exponent(dev_src.data_ptr(), e.data_ptr(), nblk, B, _DT[src.dtype], B,
                                                    float(epsilon), _stream()))
    return e.to(device=t.device, dtype=t.dtype)


def _convert_blocked_float_to_bfp(t, mant_bits, epsilon, rounding_mode, device):
    """bfp_ops.py:35-44: t is the [nblk, B] view; every row is one block."""
    return _fused(t, _lib.ORDER_QUANT_ONLY, block_size=t.shape[-1], mant_bits=mant_bits, epsilon=epsilon,
                  rounding_mode=rounding_mode)


def _no_sparsity_float_to_bfp(t, block_size, mant_bits, epsilon, rounding_mode, device):
    """bfp_ops.py:46-59: blocks of block_size along the last dim (zero-padded tail, never straddling rows)."""
    return _fused(t, _lib.ORDER_QUANT_ONLY, block_size=block_size, mant_bits=mant_bits, epsilon=epsilon,
                  rounding_mode=rounding_mode)


# ---------------------------------------------------------------------------------------------------------------
# bfp_ops.py:61-102: sparsity
# ---------------------------------------------------------------------------------------------------------------
def _unstructured_fused(t, sparsity_frac, order, block_size=0, mant_bits=0, epsilon=0.0, rounding_mode=rounding_modes.DETERM,
                        philox=None):
    """Global magnitude pruning fused with the BFP quantiser (csrc/bfp_unstructured_fused.cu): a sampled bracket of the k-th
    magnitude, one counting read, one masking + quantising read/write.  Returns None when the call does not fit that path
    (CPU tensor, ragged block structure, k == 0 or k == numel): the caller then composes the stand-alone kernels."""
    if os.environ.get("BFP_UNSTRUCTURED_FUSED", "1") == "0" or not t.is_cuda or t.dtype not in _DT or t.dim() == 0:
        return None
    src = t.detach().contiguous()
    n = src.numel()
    k = int(n * sparsity_frac)                                   # bfp_ops.py:66
    vec = 4 if src.dtype == torch.float32 else 8
    if n == 0 or k <= 0 or k >= n or n % vec or src.data_ptr() % 16:
        return None
    K = src.shape[-1]
    quant = order != _lib.ORDER_SPARSIFY_ONLY
    if quant:
        B = int(block_size)
        if B < vec or B & (B - 1) or B // vec > 32 or K % B:
            return None
    rounding = _rounding_code(rounding_mode) if quant else _lib.ROUND_NEAREST
    stoc = rounding == _lib.ROUND_STOCHASTIC
    out = torch.empty(src.shape, dtype=torch.float32 if stoc else src.dtype, device=src.device)
    seed, offset = (philox if philox is not None else _PhiloxState.next()) if stoc else (0, 0)
    L = _lib.lib()
    with _on(src.device):
        nbytes = L.bfp_unstructured_quantize_workspace_bytes(n, _DT[src.dtype])
        ws = torch.empty(nbytes // 8 + 1, dtype=torch.int64, device=src.device)
        _lib.check(L.bfp_unstructured_quantize(src.data_ptr(), out.data_ptr(), n // K, K, _DT[src.dtype], _DT[out.dtype], k, order,
                                               int(block_size), int(mant_bits), float(epsilon), rounding, seed, offset,
                                               ws.data_ptr(), nbytes, _stream()))
    return out


def _unstructured_sparsity(t, device, sparsity_frac=0):
    """bfp_ops.py:61-71: global magnitude pruning -- the k = int(numel * frac) smallest |t| of the whole tensor become
    +0.0 (ties at the threshold in index order, like torch-CUDA's topk).  Two reads + one write through the sampled-bracket
    pipeline (csrc/bfp_unstructured_fused.cu); shapes it does not take go through the multi-pass radix select
    (csrc/bfp_unstructured.cu); CPU tensors are staged through the GPU."""
    assert (sparsity_frac > 0)
    if t.dtype not in _DT:
        raise TypeError(f"bfp_b200 supports float32 / float16 / bfloat16 tensors, got {t.dtype}")
    fused = _unstructured_fused(t, sparsity_frac, _lib.ORDER_SPARSIFY_ONLY)
    if fused is not None:
        return fused
    src = t.detach().contiguous()
    n = src.numel()
    k = int(n * sparsity_frac)                                   # bfp_ops.py:66
    dev_src = src if src.is_cuda else src.cuda()
    out = torch.empty_like(dev_src)
    if n:
        L = _lib.lib()
        with _on(dev_src.device):
            ws = torch.empty(L.bfp_unstructured_workspace_bytes() // 8 + 1, dtype=torch.int64, device=dev_src.device)
            _lib.check(L.bfp_unstructured_sparsify(dev_src.data_ptr(), out.data_ptr(), n, _DT[src.dtype], k, ws.data_ptr(),
                                                   _stream()))
    return out if src.is_cuda else out.cpu()


def _structured_N_M_sparsity(t, device, N=0, M=0):
    """bfp_ops.py:73-91: zero the M-N smallest-magnitude entries of every group of M along the last dim."""
    assert ((N > 0) and (M > 0) and (N <= M))
    return _fused(t, _lib.ORDER_SPARSIFY_ONLY, N=N, M=M)


def _sparsify(t, sparsity, sparsity_mode, device, N, M, sparsity_frac):
    """bfp_ops.py:93-102"""
    if sparsity == True:  # noqa: E712  (the reference compares with ==, so 1 counts as True)
        if sparsity_mode == 'structured':
            return _structured_N_M_sparsity(t, device, N, M)
        elif sparsity_mode == 'unstructured':
            return _unstructured_sparsity(t, device, sparsity_frac)
        else:
            raise ValueError(f'Unknown sparsity mode: {sparsity_mode} given as argument')
    return t


def _quantize(t, num_format, block_size, mant_bits, weight_mant_bits, sgd_update, epsilon, rounding_mode, device,
              identifier):
    """bfp_ops.py:104-122"""
    if num_format == 'fp32':
        return t
    elif num_format == 'bfp':
        if sgd_update:
            mant_bits = weight_mant_bits
        return _no_sparsity_float_to_bfp(t, block_size, mant_bits, epsilon, rounding_mode, device)
    elif num_format == 'int':
        if sgd_update:
            mant_bits = weight_mant_bits
        return _int_quantize(t, mant_bits, weight=(identifier == 'w'))
    raise ValueError(f'Unknown quantization format: {num_format} given as argument')


def _int_quantize(t, bits, weight):
    """The 'int' format: int_ops.Quantizer (perchannel, sym) as used by bfp_ops.py:111-120 -- configure(bits),
    find_params(t, weight), quantize(t) -- in four small kernels (csrc/bfp_int.cu).  fp32 output for every input dtype,
    like the reference.  Channel = dim 0 for weights; last dim (2-D / 3-D) or dim 1 (4-D) for activations."""
    if t.dtype not in _DT:
        raise TypeError(f"bfp_b200 supports float32 / float16 / bfloat16 tensors, got {t.dtype}")
    shape = tuple(t.shape)
    if weight:
        if len(shape) < 2:
            raise IndexError("Dimension out of range (the reference flattens weights from dim 1)")
        A, C, inner = 1, shape[0], (t.numel() // shape[0] if shape[0] else 0)
    elif len(shape) == 4:
        A, C, inner = shape[0], shape[1], shape[2] * shape[3]
    elif len(shape) in (2, 3):
        C = shape[-1]
        A, inner = (t.numel() // C if C else 0), 1
    else:
        raise IndexError("int format: activations must be 2-D, 3-D or 4-D (int_ops.py:44-50)")
    src = t.detach().contiguous()
    dev_src = src if src.is_cuda else src.cuda()
    out = torch.empty(shape, dtype=torch.float32, device=dev_src.device)
    if src.numel():
        L = _lib.lib()
        with _on(dev_src.device):
            ws = torch.empty(L.bfp_int_workspace_bytes(C) // 4 + 1, dtype=torch.int32, device=dev_src.device)
            _lib.check(L.bfp_int_quantize(dev_src.data_ptr(), out.data_ptr(), A, C, inner, _DT[src.dtype], int(bits), ws.data_ptr(),
                                          _stream()))
    return out if src.is_cuda else out.cpu()


def _int_nm_fusable(t, N, M):
    """2-D weights whose rows fit one CTA's registers: N:4 mask and INT per-channel quantiser in a single pass."""
    if M != 4 or not (0 < N < 4) or t.dim() != 2 or t.dtype not in _DT:
        return False
    vec = 4 if t.dtype == torch.float32 else 8
    K = t.shape[1]
    return t.numel() > 0 and K % vec == 0 and K // vec <= 4096


def _int_quantize_nm(t, bits, N, M, order):
    """_quantize(_sparsify(t)) / _sparsify(_quantize(t)) for the 'int' format with N:4 sparsity (bfp_ops.py:143-149 over
    :73-91 and int_ops.py), one kernel: each CTA holds a weight row in registers."""
    src = t.detach().contiguous()
    dev_src = src if src.is_cuda else src.cuda()
    out = torch.empty(tuple(t.shape), dtype=torch.float32, device=dev_src.device)
    with _on(dev_src.device):
        _lib.check(_lib.lib().bfp_int_quantize_nm(dev_src.data_ptr(), out.data_ptr(), src.shape[0], src.shape[1], _DT[src.dtype],
                                                  int(bits), int(N), int(M), order, _stream()))
    return out if src.is_cuda else out.cpu()


def float_to_bfp_blocked(t, mant_bits, epsilon, rounding_mode, device, block_size,
                         num_format, weight_mant_bits, in_sparsity, w_sparsity, grad_sparsity,
                         sparsity_frac, N, M, sparsity_num_format, first, sparsity_mode, identifier='', sgd_update=False,
                         mx_w_elem_format='', mx_a_elem_format='', scale_bits=0, bfloat=0):
    """bfp_ops.py:124-149, the quantiser entry point.  Structured sparsity with the 'bfp' or 'fp32' format is ONE fused
    kernel in either order; the remaining combinations compose the stand-alone functions exactly like the reference."""
    assert (num_format == 'bfp')
    assert (((sparsity_num_format == 'bfp') and (block_size > 0)) or (sparsity_num_format == 'fp32')
            or (sparsity_num_format == 'int'))

    sparsity = ((in_sparsity == True and identifier == 'in') or (w_sparsity == True and identifier == 'w')  # noqa: E712
                or (grad_sparsity == True and identifier == 'grad'))                                        # noqa: E712

    fusable = sparsity_num_format in ('bfp', 'fp32') and (not sparsity or sparsity_mode == 'structured')
    if fusable:
        if sparsity:
            assert ((N > 0) and (M > 0) and (N <= M))                                   # bfp_ops.py:74
        if sparsity_num_format == 'fp32':
            return _fused(t, _lib.ORDER_SPARSIFY_ONLY, N=N, M=M) if sparsity else t
        m = weight_mant_bits if sgd_update else mant_bits                               # bfp_ops.py:108-109
        if not sparsity:
            order = _lib.ORDER_QUANT_ONLY
        else:
            order = _lib.ORDER_SPARSIFY_QUANT if first == 's' else _lib.ORDER_QUANT_SPARSIFY
        return _fused(t, order, block_size=block_size, mant_bits=m, epsilon=epsilon, rounding_mode=rounding_mode,
                      N=N, M=M)

    if (sparsity and sparsity_mode == 'structured' and sparsity_num_format == 'int' and identifier == 'w'
            and _int_nm_fusable(t, N, M)):
        bits = weight_mant_bits if sgd_update else mant_bits                            # bfp_ops.py:113-114
        return _int_quantize_nm(t, bits, N, M, _lib.ORDER_SPARSIFY_QUANT if first == 's' else _lib.ORDER_QUANT_SPARSIFY)

    if sparsity and sparsity_mode == 'unstructured' and sparsity_num_format == 'bfp':
        # global magnitude pruning and the BFP quantiser in two reads + one write (either order)
        assert (sparsity_frac > 0)                                                      # bfp_ops.py:62
        m = weight_mant_bits if sgd_update else mant_bits                               # bfp_ops.py:108-109
        y = _unstructured_fused(t, sparsity_frac, _lib.ORDER_SPARSIFY_QUANT if first == 's' else _lib.ORDER_QUANT_SPARSIFY,
                                block_size=block_size, mant_bits=m, epsilon=epsilon, rounding_mode=rounding_mode)
        if y is not None:
            return y

    if first == 's':
        sparse_t = _sparsify(t, sparsity, sparsity_mode, device, N, M, sparsity_frac)
        return _quantize(sparse_t, sparsity_num_format, block_size, mant_bits, weight_mant_bits, sgd_update, epsilon,
                         rounding_mode, device, identifier)
    quant_t = _quantize(t, sparsity_num_format, block_size, mant_bits, weight_mant_bits, sgd_update, epsilon,
                        rounding_mode, device, identifier)
    return _sparsify(quant_t, sparsity, sparsity_mode, device, N, M, sparsity_frac)


def float_to_bfp_tiled(t, **bfp_args):
    """Name imported by the reference's bfp_optim.py:6 but missing from its bfp_ops.py;
    called as float_to_bfp_tiled(p.data, sgd_update=True, **bfp_args)."""
    return float_to_bfp_blocked(t, **bfp_args)


# ---------------------------------------------------------------------------------------------------------------
# Packed operands + tensor-core BFP linear (new: the reference only ever holds dequantised floats).
# ---------------------------------------------------------------------------------------------------------------
class PackedBFP:
    """A [rows, K] tensor in the packed BFP operand format of include/bfp_b200.h: int8 mantissas [rows, Kp] and the
    block-major fp32 scale table [nkb_pad, rows_pad] (scale = 2^(e - m))."""
    __slots__ = ("mant", "scale_t", "shape", "rows", "K", "block_size", "mant_bits")

    def __init__(self, mant, scale_t, shape, block_size, mant_bits):
        self.mant, self.scale_t, self.shape = mant, scale_t, tuple(shape)
        self.K = self.shape[-1]
        self.rows = mant.shape[0]
        self.block_size, self.mant_bits = block_size, mant_bits


def _order_for(bfp_args, identifier):
    sparsity = ((bfp_args['in_sparsity'] == True and identifier == 'in') or (bfp_args['w_sparsity'] == True and identifier == 'w')  # noqa: E712
                or (bfp_args['grad_sparsity'] == True and identifier == 'grad'))                                                  # noqa: E712
    if not sparsity:
        return _lib.ORDER_QUANT_ONLY
    return _lib.ORDER_SPARSIFY_QUANT if bfp_args['first'] == 's' else _lib.ORDER_QUANT_SPARSIFY


def pack_bfp(t, identifier='', philox=None, **bfp_args):
    """float_to_bfp_blocked (bfp_ops.py:124-149) straight into the packed form (one fused kernel).  Needs the 'bfp' format,
    structured (or no) sparsity and mant_bits <= 7.  unpack_bfp(pack_bfp(x)) == float_to_bfp_blocked(x) up to the sign of
    zero; N:M ties follow the torch-CUDA rule."""
    assert (bfp_args['num_format'] == 'bfp') and (bfp_args['sparsity_num_format'] == 'bfp') and (bfp_args['block_size'] > 0)
    order = _order_for(bfp_args, identifier)
    if order != _lib.ORDER_QUANT_ONLY and bfp_args['sparsity_mode'] != 'structured':
        raise NotImplementedError("packed operands support structured N:M sparsity only")
    if not t.is_cuda:
        raise ValueError("pack_bfp needs a CUDA tensor")
    if t.dtype not in _DT:
        raise TypeError(f"unsupported dtype {t.dtype}")
    src = t.detach().contiguous()
    K = src.shape[-1]
    rows = src.numel() // K if K else 0
    B, m = int(bfp_args['block_size']), int(bfp_args['mant_bits'])
    Kp, rows_pad, nkb_pad = _lib.packed_layout(rows, K, B)
    nkb = -(-K // B)
    alloc_m = torch.empty if Kp == K else torch.zeros
    alloc_s = torch.empty if (rows_pad == rows and nkb_pad == nkb) else torch.zeros
    mant = alloc_m((rows, Kp), dtype=torch.int8, device=src.device)
    scale_t = alloc_s((nkb_pad, rows_pad), dtype=torch.float32, device=src.device)
    rounding = _rounding_code(bfp_args['rounding_mode'])
    seed, offset = (philox if philox is not None else _PhiloxState.next()) if rounding == _lib.ROUND_STOCHASTIC else (0, 0)
    if rows and K:
        with _on(src.device):
            _lib.check(_lib.lib().bfp_quantize_pack(src.data_ptr(), mant.data_ptr(), scale_t.data_ptr(), rows, K, _DT[src.dtype], B, m,
                                                    float(bfp_args['epsilon']), rounding, seed, offset, int(bfp_args['N']),
                                                    int(bfp_args['M']), order, _stream()))
    return PackedBFP(mant, scale_t, src.shape, B, m)


def unpack_bfp(p):
    """packed -> fp32 tensor of the original shape."""
    out = torch.empty(p.shape, dtype=torch.float32, device=p.mant.device)
    if out.numel():
        with _on(out.device):
            _lib.check(_lib.lib().bfp_unpack(p.mant.data_ptr(), p.scale_t.data_ptr(), out.data_ptr(), p.rows, p.K, p.block_size,
                                             _stream()))
    return out


def bfp_linear_packed(xp, wp, bias=None):
    """y[..., n] = sum_k x[..., k] w[n, k] + bias[n] on packed operands: tcgen05 int8 MMA per BFP block, fp32 rescale
    (include/bfp_b200.h bfp_gemm_i8).  Returns fp32 of shape xp.shape[:-1] + (N,)."""
    assert xp.K == wp.K and xp.block_size == wp.block_size
    N = wp.rows
    out = torch.empty(xp.shape[:-1] + (N,), dtype=torch.float32, device=xp.mant.device)
    b = None
    if bias is not None:
        b = bias.detach().to(dtype=torch.float32).contiguous()
    if out.numel():
        with _on(out.device):
            _lib.check(_lib.lib().bfp_gemm_i8(xp.mant.data_ptr(), xp.scale_t.data_ptr(), wp.mant.data_ptr(), wp.scale_t.data_ptr(),
                                              b.data_ptr() if b is not None else None, out.data_ptr(), xp.rows, N, xp.K,
                                              xp.block_size, _stream()))
    return out


class PackedMX:
    """Block-scaled FP8-class operand (include/bfp_b200.h bfp_mx_from_packed): E4M3 bytes of the mantissas + UE8M0 scale atoms."""
    __slots__ = ("vals", "sf", "rows", "K", "tile_rows", "block_size", "mant_bits", "folded")

    def __init__(self, vals, sf, rows, K, tile_rows, block_size, mant_bits, folded):
        self.vals, self.sf, self.rows, self.K, self.tile_rows = vals, sf, rows, K, tile_rows
        self.block_size, self.mant_bits, self.folded = block_size, mant_bits, folded


def pack_bfp_mx(t, tile_rows=128, fold=False, identifier='', philox=None, check=True, **bfp_args):
    """float_to_bfp_blocked into the block-scaled operand form of bfp_gemm_mx: needs mant_bits <= 4 (the integer mantissa is exact
    in E4M3).  General form (activations, tile_rows 128): block_size a multiple of 32 (one hardware scale per 32 elements of K).
    fold=True (weights; tile_rows = the GEMM's N tile, 128 / 240 / 256): the block exponents ride in the E4M3 values and the row has
    one scale -- any block size, exact while a row's block exponents span at most ten octaves.  Returns None (check=True) when the
    tensor does not fit the requested form."""
    m, B = int(bfp_args['mant_bits']), int(bfp_args['block_size'])
    if not (1 <= m <= 4) or (not fold and B % 32):
        raise ValueError("the block-scaled form needs mant_bits <= 4 and (general form) block_size % 32 == 0")
    pk = pack_bfp(t, identifier=identifier, philox=philox, **bfp_args)
    L = _lib.lib()
    kp, sfb = ctypes.c_int64(), ctypes.c_int64()
    _lib.check(L.bfp_mx_layout(pk.rows, pk.K, int(tile_rows), int(bool(fold)), ctypes.byref(kp), ctypes.byref(sfb)))
    dev = pk.mant.device
    vals = torch.empty((pk.rows, kp.value), dtype=torch.uint8, device=dev)
    sf = torch.empty(max(sfb.value, 16), dtype=torch.uint8, device=dev)
    viol = torch.zeros(1, dtype=torch.int32, device=dev)
    ref = torch.empty(max(pk.rows, 1), dtype=torch.int32, device=dev) if fold else None
    if pk.rows and pk.K:
        with _on(dev):
            _lib.check(L.bfp_mx_from_packed(pk.mant.data_ptr(), pk.scale_t.data_ptr(), pk.rows, pk.K, B, int(tile_rows), int(bool(fold)),
                                            vals.data_ptr(), sf.data_ptr(), ref.data_ptr() if fold else None, viol.data_ptr(), _stream()))
        if check and int(viol.item()):
            return None
    return PackedMX(vals, sf, pk.rows, pk.K, int(tile_rows), B, m, bool(fold))


def bfp_linear_mx(xp, wp, bias=None, out_shape=None):
    """y = x w^T + bias on block-scaled operands (include/bfp_b200.h bfp_gemm_mx): tcgen05.mma.kind::mxf8f6f4.block_scale."""
    assert xp.K == wp.K and xp.tile_rows == 128 and not xp.folded
    N = wp.rows
    out = torch.empty((xp.rows, N), dtype=torch.float32, device=xp.vals.device)
    b = bias.detach().to(dtype=torch.float32).contiguous() if bias is not None else None
    if out.numel():
        with _on(out.device):
            _lib.check(_lib.lib().bfp_gemm_mx(xp.vals.data_ptr(), xp.sf.data_ptr(), wp.vals.data_ptr(), wp.sf.data_ptr(), wp.tile_rows, int(wp.folded),
                                              b.data_ptr() if b is not None else None, out.data_ptr(), xp.rows, N, xp.K, _stream()))
    return out if out_shape is None else out.view(out_shape)


def pack_activation_mx(x, bfp_args):
    """An activation tensor [..., K] in the general block-scaled form (tile_rows 128) for bfp_gemm_mx: one fused kernel
    (bfp_quantize_pack_mx) when the configuration is quantise-only with nearest rounding and K is a multiple of 128 (fp32) / 256
    (half); otherwise bfp_quantize_pack + bfp_mx_from_packed."""
    B, m = int(bfp_args['block_size']), int(bfp_args['mant_bits'])
    src = x.detach().contiguous()
    K = src.shape[-1]
    rows = src.numel() // K if K else 0
    vec = 4 if src.dtype == torch.float32 else 8
    fused = (bfp_args['in_sparsity'] != True and bfp_args['rounding_mode'] == rounding_modes.DETERM and B in (32, 64, 128)      # noqa: E712
             and K % (32 * vec) == 0 and rows > 0 and src.data_ptr() % 16 == 0)
    if not fused:
        return pack_bfp_mx(src.view(rows, K), 128, identifier='in', check=False, **bfp_args)
    L = _lib.lib()
    sfb = ctypes.c_int64()
    _lib.check(L.bfp_mx_layout(rows, K, 128, 0, None, ctypes.byref(sfb)))
    vals = torch.empty((rows, K), dtype=torch.uint8, device=src.device)
    sf = (torch.zeros if rows % 128 else torch.empty)(max(sfb.value, 16), dtype=torch.uint8, device=src.device)
    with _on(src.device):
        _lib.check(L.bfp_quantize_pack_mx(src.data_ptr(), vals.data_ptr(), sf.data_ptr(), rows, K, _DT[src.dtype], B, m, float(bfp_args['epsilon']),
                                          _stream()))
    return PackedMX(vals, sf, rows, K, 128, B, m, False)


def _mx_eligible(x, w, bfp_args):
    """HBFP4 / HBFP5 inference on the block-scaled FP8-class tensor-core path (csrc/bfp_gemm_mx.cu): mant_bits <= 4, activation
    blocks that are multiples of 32, an output width the TMA store can take.  BFP_GEMM_KIND=sp|bf16|i8 opts out, =mx insists."""
    want = os.environ.get("BFP_GEMM_KIND", "")
    if want not in ("", "mx"):
        return False
    return (bfp_args['rounding_mode'] == rounding_modes.DETERM and 1 <= bfp_args['mant_bits'] <= 4 and bfp_args['block_size'] in (32, 64, 128)
            and w.shape[0] % 4 == 0 and x.dim() >= 1 and x.shape[-1] == w.shape[1]
            and not (bfp_args['in_sparsity'] == True and bfp_args['sparsity_mode'] == 'unstructured'))      # noqa: E712


def _packed_activation_mx(x, bfp_args):
    if os.environ.get("BFP_ACT_CACHE", "1") != "1" or torch.cuda.is_current_stream_capturing():
        return pack_activation_mx(x, bfp_args)
    key = ("mx", bfp_args['block_size'], bfp_args['mant_bits'], float(bfp_args['epsilon']), bfp_args['in_sparsity'] == True,   # noqa: E712
           bfp_args['N'], bfp_args['M'], bfp_args['first'], bfp_args['sparsity_mode'], float(bfp_args['sparsity_frac']), _stream(x.device))
    hit = _ACT_CACHE.get(x.device)
    if hit is not None and hit[0]() is x and hit[1] == x._version and hit[2] == key:
        return hit[3]
    xp = pack_activation_mx(x, bfp_args)
    _act_cache_put(x, key, xp)
    return xp


def pack_bfp_bf16(t, identifier='', philox=None, **bfp_args):
    """float_to_bfp_blocked straight to a dequantised bf16 [rows, Kp] operand (Kp = K rounded up to 8).  Exact for
    mant_bits <= 8: q * 2^(e-m) has at most 8 significant bits.  Any block size."""
    assert (bfp_args['num_format'] == 'bfp') and (bfp_args['sparsity_num_format'] == 'bfp') and (bfp_args['block_size'] > 0)
    order = _order_for(bfp_args, identifier)
    if not t.is_cuda:
        raise ValueError("pack_bfp_bf16 needs a CUDA tensor")
    src = t.detach().contiguous()
    K = src.shape[-1]
    rows = src.numel() // K if K else 0
    Kp = -(-K // 8) * 8
    out = (torch.empty if Kp == K else torch.zeros)((rows, Kp), dtype=torch.bfloat16, device=src.device)
    if order != _lib.ORDER_QUANT_ONLY and bfp_args['sparsity_mode'] != 'structured':
        # global magnitude pruning is a whole-tensor selection, not a per-vector one: compose the radix-select kernels with
        # the quantiser exactly like bfp_ops.py:143-149, then narrow to bf16 -- exact, the values are BFP values or zeros
        y = float_to_bfp_blocked(src, identifier=identifier, **bfp_args)
        if rows and K:
            out[:, :K].copy_(y.reshape(rows, K))
        return out
    rounding = _rounding_code(bfp_args['rounding_mode'])
    seed, offset = (philox if philox is not None else _PhiloxState.next()) if rounding == _lib.ROUND_STOCHASTIC else (0, 0)
    if rows and K:
        with _on(src.device):
            _lib.check(_lib.lib().bfp_quantize_pack_bf16(src.data_ptr(), out.data_ptr(), rows, K, _DT[src.dtype],
                                                         int(bfp_args['block_size']), int(bfp_args['mant_bits']), float(bfp_args['epsilon']),
                                                         rounding, seed, offset, int(bfp_args['N']), int(bfp_args['M']), order,
                                                         _stream()))
    return out


def bfp_linear_bf16(xb, wb, bias=None, out_shape=None, out_dtype=torch.float32):
    """y = x w^T + bias on exact-bf16 BFP operands (include/bfp_b200.h bfp_gemm_bf16_ex): tcgen05.mma.kind::f16, fp32 TMEM
    accumulation, no per-block rescale; the epilogue writes fp32 or rounds once to fp16 / bf16.  xb [T, Kp], wb [N, Kp] bf16."""
    T, Kp = xb.shape
    N = wb.shape[0]
    assert wb.shape[1] == Kp and xb.dtype == torch.bfloat16 and wb.dtype == torch.bfloat16
    out = torch.empty((T, N), dtype=out_dtype, device=xb.device)
    b = bias.detach().to(dtype=torch.float32).contiguous() if bias is not None else None
    if out.numel():
        with _on(out.device):
            _lib.check(_lib.lib().bfp_gemm_bf16_ex(xb.data_ptr(), wb.data_ptr(), b.data_ptr() if b is not None else None, out.data_ptr(),
                                                   _DT[out_dtype], T, N, Kp, _stream()))
    return out.view(out_shape) if out_shape is not None else out


class SparseBF16:
    """A 2:4-pruned exact-bf16 operand in the compressed form of include/bfp_b200.h (bfp_compress_2to4_bf16): the two kept
    values of every group of four along K (`comp` bf16 [rows, Kc]) and the index nibbles (`meta` uint8)."""

    def __init__(self, comp, meta, rows, K):
        self.comp, self.meta, self.rows, self.K = comp, meta, rows, K


def compress_2to4_bf16(wb, check=True):
    """bf16 [rows, Kp] operand whose groups of four along K hold at most two non-zeros -> SparseBF16.  With check=True the
    violation counter is read back (one sync; weights are compressed once and cached) and a non-2:4 input raises."""
    assert wb.dtype == torch.bfloat16 and wb.dim() == 2 and wb.is_cuda and wb.is_contiguous()
    rows, Kp = wb.shape
    Kc, meta_bytes = _lib.sp_layout(rows, Kp)
    comp = torch.empty((rows, Kc), dtype=torch.bfloat16, device=wb.device)
    meta = torch.empty((meta_bytes,), dtype=torch.uint8, device=wb.device)
    viol = torch.zeros((1,), dtype=torch.int32, device=wb.device)
    if rows and Kp:
        with _on(wb.device):
            _lib.check(_lib.lib().bfp_compress_2to4_bf16(wb.data_ptr(), rows, Kp, comp.data_ptr(), meta.data_ptr(), viol.data_ptr(),
                                                         _stream()))
    if check and int(viol.item()) != 0:
        raise ValueError(f"operand is not 2:4 sparse along K ({int(viol.item())} groups of 16 with a dense group of four)")
    return SparseBF16(comp, meta, rows, Kp)


def bfp_linear_bf16_sp(xb, ws, bias=None, out_shape=None, out_dtype=torch.float32):
    """y = x w^T + bias with the 2:4-compressed weight `ws` (include/bfp_b200.h bfp_gemm_bf16_sp_ex): tcgen05.mma.sp.kind::f16,
    fp32 TMEM accumulation; the epilogue writes fp32, or rounds once to fp16 / bf16 (`out_dtype`).  xb bf16 [T, Kp]."""
    T, Kp = xb.shape
    N = ws.rows
    assert ws.K == Kp and xb.dtype == torch.bfloat16
    out = torch.empty((T, N), dtype=out_dtype, device=xb.device)
    b = bias.detach().to(dtype=torch.float32).contiguous() if bias is not None else None
    if out.numel():
        with _on(out.device):
            _lib.check(_lib.lib().bfp_gemm_bf16_sp_ex(xb.data_ptr(), ws.comp.data_ptr(), ws.meta.data_ptr(),
                                                      b.data_ptr() if b is not None else None, out.data_ptr(), _DT[out_dtype], N, T, N, Kp,
                                                      _stream()))
    return out.view(out_shape) if out_shape is not None else out


def _nm_fits_2to4(N, M):
    """True when keeping any N of every M consecutive values along K (groups start at the row start) leaves at most two
    non-zeros in every aligned group of four -- what tcgen05.mma.sp needs.  Worst case per window: each overlapping
    M-group contributes min(N, overlap)."""
    if N <= 0 or M <= 0 or N >= M:
        return False
    period = M * 4 // math.gcd(M, 4)
    for w0 in range(0, period, 4):
        overlap = {}
        for i in range(w0, w0 + 4):
            overlap[i // M] = overlap.get(i // M, 0) + 1
        if sum(min(N, c) for c in overlap.values()) > 2:
            return False
    return True


def _tensor_core_kind(x, w, bfp_args):
    """Which tensor-core contraction serves this configuration:
      'sp'   exact-bf16 operands, weight 2:4-compressed, tcgen05.mma.sp (w_sparsity with an N:M that fits 2:4);
      'bf16' exact-bf16 operands, dense MMA (any block size, mant_bits <= 8);
      'i8'   int8 mantissas + per-block rescale (block 32/64/128, mant_bits <= 7);
      None   fake-quant + library GEMM, the reference's own structure.
    BFP_GEMM_KIND=sp|bf16|i8 picks among the eligible ones (default: sp when the weight is 2:4, else bf16 -- the faster
    of the dense two for every block size <= 64)."""
    if not _tensor_core_eligible(x, w, dict(bfp_args, block_size=64 if bfp_args['block_size'] > 0 else 0, mant_bits=min(bfp_args['mant_bits'], 7))):
        return None
    B, m = bfp_args['block_size'], bfp_args['mant_bits']
    want = os.environ.get("BFP_GEMM_KIND", "")
    unstructured = bfp_args['w_sparsity'] == True and bfp_args['sparsity_mode'] == 'unstructured'            # noqa: E712
    unstructured = unstructured or (bfp_args['in_sparsity'] == True and bfp_args['sparsity_mode'] == 'unstructured')    # noqa: E712
    i8_ok = B in (32, 64, 128) and 1 <= m <= 7 and not unstructured
    bf16_ok = 1 <= m <= 8 and B >= 4 and (B & (B - 1)) == 0
    sp_ok = (bf16_ok and bfp_args['w_sparsity'] == True and bfp_args['sparsity_mode'] == 'structured'     # noqa: E712
             and _nm_fits_2to4(bfp_args['N'], bfp_args['M']))
    if want == "i8" and i8_ok:
        return 'i8'
    if want in ("", "sp") and sp_ok:
        return 'sp'
    if bf16_ok:
        return 'bf16'
    return 'i8' if i8_ok else None


def _tensor_core_eligible(x, w, bfp_args):
    """The packed tensor-core path computes the same function as quantise + F.linear (fp32 accumulate order aside); it is
    taken on CUDA tensors when the configuration has a packed form: fp32 (inference and training), fp16 / bf16 (inference:
    the contraction accumulates in fp32 like the library HGEMM the reference calls and is rounded to the dtype once)."""
    if os.environ.get("BFP_LINEAR_PATH", "tc") != "tc":
        return False
    training = torch.is_grad_enabled() and (x.requires_grad or w.requires_grad)
    if training and os.environ.get("BFP_TRAIN_PATH", "tc") != "tc":
        return False
    dtype_ok = x.dtype == w.dtype and x.dtype in _DT          # half precision: fp32 accumulation, one rounding to the dtype (library HGEMM semantics)
    if bfp_args['rounding_mode'] != rounding_modes.DETERM and not training:
        # the stochastic quantiser promotes every operand to fp32: any mix of float dtypes is one case
        dtype_ok = x.dtype in _DT and w.dtype in _DT
    return (x.is_cuda and w.is_cuda and dtype_ok
            and bfp_args['num_format'] == 'bfp' and bfp_args['sparsity_num_format'] == 'bfp'
            and 1 <= bfp_args['mant_bits'] <= 7 and bfp_args['block_size'] in (32, 64, 128)
            and ((bfp_args['w_sparsity'] != True and bfp_args['in_sparsity'] != True)             # noqa: E712
                 or (bfp_args['sparsity_mode'] == 'unstructured' and 0 < bfp_args['sparsity_frac'] and 1 <= bfp_args['mant_bits'] <= 8)
                 or (bfp_args['sparsity_mode'] == 'structured' and 0 < bfp_args['N'] <= bfp_args['M'] <= 64
                     and (bfp_args['first'] == 's' or bfp_args['block_size'] % bfp_args['M'] == 0))))


# ---------------------------------------------------------------------------------------------------------------
# bfp_ops.py:151-200: operand pre-processing and the autograd wrappers
# ---------------------------------------------------------------------------------------------------------------
# ---------------------------------------------------------------------------------------------------------------
# Row f4: the wrapped matmul / conv2d on the tensor cores (inference).  Same functions as new_op(...) below with
# op = torch.matmul / F.conv2d: both operands quantised along the contraction (bfp_ops.py:151-155), exact-bf16 operands,
# fp32 accumulation in TMEM.
# ---------------------------------------------------------------------------------------------------------------
def _tc_inference_ok(x, w, bfp_args):
    return ((bfp_args['rounding_mode'] == rounding_modes.DETERM or (x.dtype == torch.float32 and w.dtype == torch.float32))
            and not (torch.is_grad_enabled() and (x.requires_grad or w.requires_grad))
            and _tensor_core_kind(x, w, bfp_args) is not None and 1 <= bfp_args['mant_bits'] <= 8)


def _tc_matmul(x, w, bfp_args):
    """torch.matmul(Q_in(x), Q_w(w^T)^T) for x [..., M, K], w [..., K, N] (F_matmul_bfp, bfp_ops.py:240-245 with transpose=True):
    one GEMM when w is a matrix, ONE batched launch over the broadcast batch otherwise (bfp_gemm_bf16_batched)."""
    K, N = w.shape[-2], w.shape[-1]
    if w.dim() == 2:
        wb = pack_bfp_bf16(w.t(), identifier='w', **bfp_args)                          # [N, Kp], blocked along K
        return bfp_linear_bf16(pack_bfp_bf16(x, identifier='in', **bfp_args), wb, None, out_shape=tuple(x.shape[:-1]) + (N,))
    batch = torch.broadcast_shapes(x.shape[:-2], w.shape[:-2])
    nb = math.prod(batch)
    M = x.shape[-2]
    # both operands are quantised in their OWN shape (a global magnitude threshold must not see broadcast copies), then
    # broadcast and materialised as contiguous [b, rows, Kp] stacks (no copy when nothing is broadcast)
    xb = pack_bfp_bf16(x, identifier='in', **bfp_args)
    xb = xb.view(tuple(x.shape[:-2]) + (M, xb.shape[-1])).expand(batch + (M, xb.shape[-1])).reshape(nb, M, -1).contiguous()
    wb = pack_bfp_bf16(w.transpose(-1, -2), identifier='w', **bfp_args)
    wb = wb.view(tuple(w.shape[:-2]) + (N, wb.shape[-1])).expand(batch + (N, wb.shape[-1])).reshape(nb, N, -1).contiguous()
    out = torch.empty((nb, M, N), dtype=torch.float32, device=x.device)
    L = _lib.lib()
    with _on(x.device):
        stream = _stream()
        if N % 4 == 0:                 # one launch for the whole batch (3-D output map: needs 16-byte aligned output rows)
            _lib.check(L.bfp_gemm_bf16_batched(xb.data_ptr(), wb.data_ptr(), out.data_ptr(), _lib.DT_F32, nb, M, N, xb.shape[-1], stream))
        else:
            for b in range(nb):
                _lib.check(L.bfp_gemm_bf16(xb[b].data_ptr(), wb[b].data_ptr(), None, out[b].data_ptr(), M, N, xb.shape[-1], stream))
    return out.view(batch + (M, N))


def _tc_conv2d(x, w, bias, stride, padding, dilation, groups, bfp_args):
    """F.conv2d(Q_in(x), Q_w(w), ...) as im2col + BFP GEMM (BFPConv2d, bfp_ops.py:247-268): the input is blocked along W and the
    weight along kw exactly like the reference (quantise FIRST, then unfold -- zero padding is applied to the quantised tensor
    as F.conv2d does).  The result is returned in channels-last memory format (a view of the GEMM's [B*L, O] output: no
    transposing copy; ViT flattens it straight back to [B, L, O])."""
    B, C, H, W = x.shape
    O, _, kh, kw = w.shape
    Kc = C * kh * kw
    Kp = -(-Kc // 8) * 8
    Ho = (H + 2 * padding[0] - dilation[0] * (kh - 1) - 1) // stride[0] + 1
    Wo = (W + 2 * padding[1] - dilation[1] * (kw - 1) - 1) // stride[1] + 1
    xq = pack_bfp_bf16(x, identifier='in', **bfp_args)[:, :W].reshape(B, C, H, W)       # exact bf16 values of Q_in(x)
    if _is_patch_embedding(x, w, stride, padding, dilation) and Kp == Kc:
        # kernel == stride: the windows tile the image, im2col is ONE permuting copy
        a = xq.view(B, C, Ho, kh, Wo, kw).permute(0, 2, 4, 1, 3, 5).reshape(B * Ho * Wo, Kc)
    else:
        cols = F.unfold(xq, (kh, kw), dilation=dilation, padding=padding, stride=stride)    # [B, C*kh*kw, L]
        Lout = cols.shape[-1]
        a = torch.zeros((B * Lout, Kp), dtype=torch.bfloat16, device=x.device) if Kp != Kc else torch.empty((B * Lout, Kp), dtype=torch.bfloat16, device=x.device)
        a.view(B, Lout, Kp)[:, :, :Kc] = cols.transpose(1, 2)
    wq = _pad_cols(pack_bfp_bf16(w, identifier='w', **bfp_args)[:, :kw].reshape(O, Kc), Kp)
    y = bfp_linear_bf16(a, wq, bias)                                                    # [B*L, O]
    return y.view(B, Ho, Wo, O).permute(0, 3, 1, 2)


def _is_patch_embedding(x, w, stride, padding, dilation):
    kh, kw = w.shape[-2], w.shape[-1]
    return (tuple(stride) == (kh, kw) and tuple(padding) == (0, 0) and tuple(dilation) == (1, 1)
            and x.shape[-2] % kh == 0 and x.shape[-1] % kw == 0)


def _im2col_pays(x, w, stride, padding, dilation):
    """Measured (tools/bench_conv.py, profiles/r01_conv_paths.json): with torch's unfold / transposing copies as the im2col,
    the tensor-core path only wins where im2col is a single permuting copy -- patch embeddings (kernel == stride, ViT).  For
    overlapping windows (3x3, 7x7) and 1x1 convolutions over NCHW the fused quantiser + the library convolution -- the
    reference's own structure, exact here too because BFP values fit TF32 -- is 5-10x faster, so that is the default.
    BFP_CONV_IM2COL_MAX_EXPANSION=<kh*kw / (sh*sw) limit> forces the im2col path for other shapes (tests, experiments)."""
    limit = os.environ.get("BFP_CONV_IM2COL_MAX_EXPANSION")
    if limit is not None:
        return w.shape[-2] * w.shape[-1] <= float(limit) * stride[0] * stride[1]
    return (_is_patch_embedding(x, w, stride, padding, dilation) and w.shape[-2] * w.shape[-1] > 1      # 1x1: the library wins (0.32 vs 0.65 ms)
            and (w.shape[1] * w.shape[2] * w.shape[3]) % 8 == 0)


def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def MxM_pre_processing(x, w, transpose, **bfp_args):
    """bfp_ops.py:151-155: both operands are blocked along the contraction dim."""
    xq = float_to_bfp_blocked(x, **bfp_args, identifier='in')
    if transpose == True:  # noqa: E712
        wq = torch.transpose(float_to_bfp_blocked(torch.transpose(w, -1, -2), **bfp_args, identifier='w'), -1, -2)
    else:
        wq = float_to_bfp_blocked(w, **bfp_args, identifier='w')
    return (xq, wq)


def _get_op_name(name, epsilon, mant_bits, rounding_mode, **kwargs):
    """bfp_ops.py:157-158"""
    return '%s_BFP_%s_%d' % (name, rounding_mode, mant_bits)


def _gen_bfp_op(op, name, bfp_args, transpose=False):
    """bfp_ops.py:160-192: new_op(x, w, ...) = OutGrad(op(*QuantIn(x, w), ...)).
    QuantIn: forward quantises both operands, backward is the straight-through identity (:168-170).
    OutGrad: forward identity, backward quantises the output gradient with identifier='grad' (:180-182)."""
    name = _get_op_name(name, **bfp_args)

    class NewOpIn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, w):
            return MxM_pre_processing(x, w, transpose, **bfp_args)

        @staticmethod
        def backward(ctx, grad_x, grad_w):
            return (grad_x, grad_w)

    NewOpIn.__name__ = name + '_In'

    class NewOpOut(torch.autograd.Function):
        @staticmethod
        def forward(ctx, op_out):
            return op_out.view_as(op_out)

        @staticmethod
        def backward(ctx, op_out_grad):
            return float_to_bfp_blocked(op_out_grad, **bfp_args, identifier='grad')

    NewOpOut.__name__ = name + '_Out'

    def new_op(x, w, *args, **kwargs):
        # inference fast paths on the tensor cores (row f4); anything else runs the reference's structure below
        aux_grad = torch.is_grad_enabled() and any(torch.is_tensor(t) and t.requires_grad for t in args)     # e.g. a trainable bias
        if torch.is_tensor(x) and torch.is_tensor(w) and x.dim() >= 2 and w.dim() >= 2 and not aux_grad and _tc_inference_ok(x, w, bfp_args):
            if op is torch.matmul and transpose and not args and not kwargs and x.shape[-1] == w.shape[-2]:
                return _tc_matmul(x, w, bfp_args).to(x.dtype)
            if op is F.linear and not transpose and w.dim() == 2 and len(args) <= 1 and not kwargs:
                return bfp_linear_bf16(pack_bfp_bf16(x, identifier='in', **bfp_args), pack_bfp_bf16(w, identifier='w', **bfp_args),
                                       args[0] if args else None, out_shape=tuple(x.shape[:-1]) + (w.shape[0],)).to(x.dtype)
            if op is F.conv2d and x.dim() == 4 and w.dim() == 4 and not kwargs and len(args) == 5 and args[4] == 1 \
                    and not isinstance(args[2], str) and _im2col_pays(x, w, _pair(args[1]), _pair(args[2]), _pair(args[3])):
                return _tc_conv2d(x, w, args[0], _pair(args[1]), _pair(args[2]), _pair(args[3]), 1, bfp_args).to(x.dtype)
        x, w = NewOpIn.apply(x, w)
        out = op(x, w, *args, **kwargs)
        return NewOpOut.apply(out)

    new_op.__name__ = name
    return new_op


def _get_bfp_op(op, name, bfp_args, transpose=False):
    """bfp_ops.py:194-200 (the reference's cache is a local dict, so every call generates a fresh op; same here)."""
    return _gen_bfp_op(op, name, bfp_args, transpose)


_BFP_ARG_DEFAULTS = (
    ('num_format', 'fp32'), ('sparsity_num_format', 'fp32'), ('rounding_mode', 'stoc'), ('epsilon', 1e-8),
    ('mant_bits', 0), ('block_size', 0), ('weight_mant_bits', 0), ('in_sparsity', False), ('w_sparsity', False),
    ('grad_sparsity', False), ('N', 0), ('M', 0), ('first', 's'), ('sparsity_mode', 'unstructured'),
    ('sparsity_frac', 0), ('mx_w_elem_format', ''), ('mx_a_elem_format', ''), ('bfloat', 16), ('scale_bits', 8),
    ('device', 'cpu'),
)


def unpack_bfp_args(kwargs):
    """bfp_ops.py:202-231: pops the 20 known keys (with the reference's defaults) out of `kwargs`, which is mutated;
    unknown keys stay behind and are ignored by the callers."""
    bfp_args = {}
    for arg, default in _BFP_ARG_DEFAULTS:
        bfp_args[arg] = kwargs.pop(arg) if arg in kwargs else default
    return bfp_args


def F_linear_bfp(**kwargs):
    """bfp_ops.py:233-238"""
    bfp_args = unpack_bfp_args(kwargs)
    if bfp_args['num_format'] == 'bfp':
        return _get_bfp_op(F.linear, 'linear', bfp_args)
    return F.linear


def F_matmul_bfp(**kwargs):
    """bfp_ops.py:240-245"""
    bfp_args = unpack_bfp_args(kwargs)
    if bfp_args['num_format'] == 'bfp':
        return _get_bfp_op(torch.matmul, 'matmul', bfp_args, True)
    return torch.matmul


def _pad_cols(t, cols):
    """bf16 [r, c] -> contiguous [r, cols] (zero-padded); the GEMM operands need a contraction length that is a multiple of 8."""
    if t.shape[1] == cols and t.is_contiguous():
        return t
    out = torch.zeros((t.shape[0], cols), dtype=t.dtype, device=t.device)
    out[:, :t.shape[1]] = t
    return out


def _transpose_pad(t, rows, cols, ld_out):
    """bf16 / fp16 t [>= rows, >= cols] (row-major, any row stride) -> contiguous [cols, ld_out] = t[:rows, :cols]^T, zero-padded
    (csrc/bfp_pack.cu transpose16_kernel: torch's transposing copy runs at about a quarter of this kernel's bandwidth)."""
    assert t.dim() == 2 and t.stride(1) == 1 and t.element_size() == 2 and ld_out % 2 == 0 and ld_out >= rows
    out = torch.empty((cols, ld_out), dtype=t.dtype, device=t.device)
    if rows and cols:
        with _on(t.device):
            _lib.check(_lib.lib().bfp_transpose_pad_16(t.data_ptr(), out.data_ptr(), rows, cols, t.stride(0), ld_out, _stream()))
    elif out.numel():
        out.zero_()
    return out


class _BFPLinearTC(torch.autograd.Function):
    """BFPLinear forward AND backward on the tensor cores.  Same function as the reference's
    new_op (bfp_ops.py:160-192): forward F.linear(Q_in(x), Q_w(w), bias); backward = F.linear's backward applied to the
    output gradient quantised with identifier='grad' (:180-182), straight-through for x and w (:168-170):
        grad_x = Q_g(gy) . Q_w(w)        grad_w = Q_g(gy)^T . Q_in(x)        grad_b = sum_t Q_g(gy)
    All three contractions run on exact-bf16 BFP operands (bfp_gemm_bf16 / bfp_gemm_bf16_sp); the operands saved for
    backward are the packed bf16 tensors (half the bytes of the reference's fp32 copies)."""

    @staticmethod
    def forward(ctx, x, w, bias, bfp_args, dense_w, cached_w):
        K, N = x.shape[-1], w.shape[0]
        xb = pack_bfp_bf16(x, identifier='in', **bfp_args)                      # [T, Kp]
        wb = cached_w if cached_w is not None else pack_bfp_bf16(w, identifier='w', **bfp_args)   # dense [N, Kp] or SparseBF16
        out_shape = tuple(x.shape[:-1]) + (N,)
        if isinstance(wb, SparseBF16):
            y = bfp_linear_bf16_sp(xb, wb, bias, out_shape=out_shape, out_dtype=x.dtype)
        else:
            y = bfp_linear_bf16(xb, wb, bias, out_shape=out_shape, out_dtype=x.dtype)
        ctx.bfp_args, ctx.K, ctx.N, ctx.x_shape, ctx.has_bias = bfp_args, K, N, tuple(x.shape), bias is not None
        ctx.x_dtype, ctx.w_dtype, ctx.b_dtype = x.dtype, w.dtype, (bias.dtype if bias is not None else None)
        ctx.wb_dense = wb if not isinstance(wb, SparseBF16) else dense_w      # tensor, or a callable that packs it on demand
        ctx.save_for_backward(xb)
        return y

    @staticmethod
    def backward(ctx, gy):
        (xb,) = ctx.saved_tensors
        a, K, N = ctx.bfp_args, ctx.K, ctx.N
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        # quantised in the gradient's own dtype (the half-precision block exponent differs from the fp32 one);
        # with grad_sparsity the N:M mask of identifier='grad' is part of the same pack
        gq = pack_bfp_bf16(gy.reshape(-1, N).contiguous(), identifier='grad', **a)  # [T, Np], blocked along N (bfp_ops.py:181)
        T, Np = gq.shape
        grad_x = grad_w = grad_b = None
        if need_x:
            wb = ctx.wb_dense() if callable(ctx.wb_dense) else ctx.wb_dense     # dgrad contracts over N: the dense form
            wbT = _transpose_pad(wb, N, K, Np)                                  # [K, Np]
            grad_x = bfp_linear_bf16(gq, wbT, out_dtype=ctx.x_dtype).view(ctx.x_shape)
        if need_w:
            Tp = -(-T // 8) * 8
            gqT = _transpose_pad(gq, T, N, Tp)                                  # [N, Tp]
            xbT = _transpose_pad(xb, T, K, Tp)                                  # [K, Tp]
            grad_w = bfp_linear_bf16(gqT, xbT, out_dtype=ctx.w_dtype)           # [N, K]
        if ctx.has_bias and ctx.needs_input_grad[2]:
            grad_b = gq[:, :N].sum(0, dtype=torch.float32).to(ctx.b_dtype)
        return grad_x, grad_w, grad_b, None, None, None


# ---------------------------------------------------------------------------------------------------------------
# Row f2 on the tensor cores: BFPLinear with sparsity_num_format == 'int' (inference, fp32 modules).
#   w_q[n, k] = s_n * j_w[n, k]       per-row scale, |j_w| <= 2^(bits-1): the integer grid is exact in bf16
#   x_q[t, k] = s_k * j_x[t, k]       per-COLUMN scale (int_ops.py:47-50 reduces activations over rows): the scale runs along
#                                     the contraction, so this is not an integer GEMM.  x_q (an fp32 value) is written as
#                                     three bf16 planes that sum to it, and
#   y[t, n]   = s_n * sum_k (hi + mid + lo)[t, k] * j_w[n, k] + bias[n]
# is a bf16 contraction over K' = 3K (dense, or 2:4 when the weight is pruned that way), then a per-column scale: the
# reference's F.linear(x_q, w_q, bias) up to the rounding of s_n * j_w (relative 2^-24).  Unlike BFP operands these sums are
# not exactly representable, and the tensor cores truncate at every accumulation step, so the contraction is chunked along
# K (include/bfp_b200.h bfp_gemm_bf16_acc): one launch for the small mid|lo planes, one accumulating launch per 1024 columns
# of the hi plane.
# ---------------------------------------------------------------------------------------------------------------
def _int_tc_eligible(x, w, bfp_args):
    # Opt-in (BFP_INT_LINEAR=tc).  Each linear agrees with the reference's fp32 GEMM to ~1e-6, but unlike BFP operands the
    # sums are not exact, and the NEXT layer's INT quantiser turns a 1e-6 difference into whole quantisation steps wherever a
    # value sits on a rounding boundary: over OPT-125M's 12 layers the logits drift to the quantisation-noise level (2.5e-2
    # for INT8, 0.2 for INT4 -- the same drift any other GEMM summation order, GPU or library version gives the reference
    # itself).  The default keeps the library GEMM, whose logits are bit-identical to the reference's on the same GPU.
    if os.environ.get("BFP_LINEAR_PATH", "tc") != "tc" or os.environ.get("BFP_INT_LINEAR", "library") != "tc":
        return False
    training = torch.is_grad_enabled() and (x.requires_grad or w.requires_grad)
    return (not training and x.is_cuda and w.is_cuda and x.dtype == torch.float32 and w.dtype == torch.float32
            and bfp_args['num_format'] == 'bfp' and bfp_args['sparsity_num_format'] == 'int' and 1 <= bfp_args['mant_bits'] <= 8
            and not (bfp_args['in_sparsity'] == True)                                                  # noqa: E712
            and w.dim() == 2 and x.dim() in (2, 3) and w.shape[1] % 8 == 0 and w.shape[0] % 4 == 0 and x.numel() > 0 and w.numel() > 0)


_INT_KSEG = 1024     # hi-plane contraction chunk: 64 dense MMA steps, i.e. ~1e-6 of truncation drift per chunk


class _IntPackedWeight:
    """Integer weight grid in the chunks of the K-chunked contraction: `hi[s]` = columns of hi segment s, `midlo` = the grid
    twice ([N, 2K]); dense bf16 tensors, or SparseBF16 when the weight is pruned 2:4.  `scale` = per-row s_n."""

    def __init__(self, hi, midlo, scale, sparse):
        self.hi, self.midlo, self.scale, self.sparse = hi, midlo, scale, sparse


def _int_pack_weight(w, bfp_args):
    """Fake-quantised weight (the reference composition, either order, any sparsity mode) -> integer grid j_w + per-row scale.
    Returns None -- the caller then keeps the reference's structure -- if s_n * j_w does not reproduce the fake-quantised
    weight bit for bit."""
    wq = float_to_bfp_blocked(w, identifier='w', **bfp_args)
    sparse = bfp_args['w_sparsity'] == True                                                            # noqa: E712
    seen = _sparsify(w, sparse, bfp_args['sparsity_mode'], w.device, bfp_args['N'], bfp_args['M'], bfp_args['sparsity_frac']) \
        if (sparse and bfp_args['first'] == 's') else w                    # the tensor Quantizer.find_params saw
    maxq = torch.tensor(float(2 ** int(bfp_args['mant_bits']) - 1), device=w.device)   # a tensor: IEEE division, not a reciprocal multiply
    xmin = seen.min(dim=1).values.clamp(max=0.0)                           # int_ops.py:55-56
    xmax = seen.max(dim=1).values.clamp(min=0.0)
    xmax = torch.maximum(xmin.abs(), xmax)                                 # :59
    xmin = torch.where(xmin < 0, -xmax, xmin)                              # :60-62
    dead = (xmin == 0) & (xmax == 0)                                       # :63-65
    xmin, xmax = torch.where(dead, -torch.ones_like(xmin), xmin), torch.where(dead, torch.ones_like(xmax), xmax)
    scale = (xmax - xmin) / maxq                                           # :67
    grid = torch.round(wq / scale[:, None])
    if not torch.equal(grid * scale[:, None], wq) or float(grid.abs().max()) > 256:
        return None
    grid = grid.to(torch.bfloat16)
    K = grid.shape[1]
    fits = sparse and bfp_args['sparsity_mode'] == 'structured' and _nm_fits_2to4(bfp_args['N'], bfp_args['M']) \
        and os.environ.get("BFP_GEMM_KIND", "") in ("", "sp")
    form = compress_2to4_bf16 if fits else (lambda t: t)
    hi = [form(grid[:, k0:k0 + _INT_KSEG].contiguous()) for k0 in range(0, K, _INT_KSEG)]
    return _IntPackedWeight(hi, form(grid.repeat(1, 2).contiguous()), scale.contiguous(), fits)


def _int_tc_linear(x, packed, bias, bfp_args):
    K = x.shape[-1]
    T = x.numel() // K
    N = packed.scale.shape[0]
    src = x.detach().contiguous()
    x3 = torch.empty(T * 3 * K, dtype=torch.bfloat16, device=x.device)
    acc = torch.empty((T, N), dtype=torch.float32, device=x.device)
    L = _lib.lib()
    with _on(x.device):
        stream = _stream()
        ws = torch.empty(L.bfp_int_workspace_bytes(K) // 8 + 2, dtype=torch.int64, device=x.device)
        _lib.check(L.bfp_int_quantize_split3(src.data_ptr(), x3.data_ptr(), T, K, _INT_KSEG, _DT[src.dtype], int(bfp_args['mant_bits']),
                                             ws.data_ptr(), stream))
        # smallest terms first: mid|lo planes in one launch (store), then one accumulating launch per hi segment
        chunks = [(x3[T * K:], packed.midlo, 2 * K)] + [(x3[k0 * T:], wseg, min(_INT_KSEG, K - k0)) for k0, wseg in zip(range(0, K, _INT_KSEG), packed.hi)]
        for i, (xc, wc, kc) in enumerate(chunks):
            if packed.sparse:
                if i == 0:
                    _lib.check(L.bfp_gemm_bf16_sp(xc.data_ptr(), wc.comp.data_ptr(), wc.meta.data_ptr(), None, acc.data_ptr(), T, N, kc, stream))
                else:
                    _lib.check(L.bfp_gemm_bf16_sp_acc(xc.data_ptr(), wc.comp.data_ptr(), wc.meta.data_ptr(), acc.data_ptr(), T, N, kc, stream))
            elif i == 0:
                _lib.check(L.bfp_gemm_bf16(xc.data_ptr(), wc.data_ptr(), None, acc.data_ptr(), T, N, kc, stream))
            else:
                _lib.check(L.bfp_gemm_bf16_acc(xc.data_ptr(), wc.data_ptr(), acc.data_ptr(), T, N, kc, stream))
    acc = acc.view(tuple(x.shape[:-1]) + (N,))
    return torch.addcmul(bias.detach(), acc, packed.scale) if bias is not None else acc.mul_(packed.scale)


class BFPConv2d(torch.nn.Conv2d):
    """bfp_ops.py:247-268: input blocked along W, weight along kw; the convolution itself runs on the dequantised
    operands (conv as im2col + BFP GEMM is "next")."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1,
                 padding=0, dilation=1, groups=1, bias=True, **kwargs):
        self.bfp_args = unpack_bfp_args(kwargs)
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias)
        self.num_format = self.bfp_args['num_format']
        self.conv_op = _get_bfp_op(F.conv2d, 'Conv2d', self.bfp_args)

    def forward(self, input):
        if self.num_format == 'fp32':
            return F.conv2d(input, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
        elif self.num_format == 'bfp':
            return self.conv_op(input, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
        raise NotImplementedError('NumFormat not implemented')


# One-entry cache of the last packed activation: sibling projections (q/k/v, gate/up) are called back to back with the SAME
# tensor object, and its packed form depends only on (values, block_size, mant_bits, epsilon).  The entry holds a reference to
# the tensor itself (so its storage cannot be recycled under the key) and its version counter (in-place writes invalidate).
# Not used while a CUDA graph is being captured (the capture must contain the pack it replays).
_ACT_CACHE = {}


def _act_cache_put(x, key, packed):
    """One entry per device: the packed form of the LAST activation, so sibling projections (q / k / v, gate / up) quantise their
    shared input once.  The activation itself is held weakly -- the cache never keeps a tensor alive -- and the entry (with its
    packed copy) goes when the activation does."""
    dev = x.device

    def gone(ref, dev=dev):
        hit = _ACT_CACHE.get(dev)
        if hit is not None and hit[0] is ref:
            del _ACT_CACHE[dev]
    _ACT_CACHE[dev] = (weakref.ref(x, gone), x._version, key, packed)


def _packed_activation(x, bfp_args):
    if os.environ.get("BFP_ACT_CACHE", "1") != "1" or torch.cuda.is_current_stream_capturing():
        return pack_bfp_bf16(x, identifier='in', **bfp_args)
    key = (bfp_args['block_size'], bfp_args['mant_bits'], float(bfp_args['epsilon']), bfp_args['in_sparsity'] == True,   # noqa: E712
           bfp_args['N'], bfp_args['M'], bfp_args['first'], bfp_args['sparsity_mode'], float(bfp_args['sparsity_frac']),
           _stream(x.device))                                                         # same stream: the entry is ordered before its reuse
    hit = _ACT_CACHE.get(x.device)
    if hit is not None and hit[0]() is x and hit[1] == x._version and hit[2] == key:
        return hit[3]
    xb = pack_bfp_bf16(x, identifier='in', **bfp_args)
    _act_cache_put(x, key, xb)
    return xb


class BFPLinear(torch.nn.Linear):
    """bfp_ops.py:270-287: nn.Linear whose operands are BFP-quantised (activations id 'in', weights id 'w') on every
    forward; parameters and state-dict are nn.Linear's.  The bias is never quantised."""

    def __init__(self, in_features, out_features, bias=True, **kwargs):
        self.bfp_args = unpack_bfp_args(kwargs)
        super().__init__(in_features, out_features, bias)
        self.num_format = self.bfp_args['num_format']
        self.linear_op = _get_bfp_op(F.linear, 'linear', self.bfp_args)
        self._packed_w = None          # (key, packed weight) of the last kind used: re-packed only when the weight changes
        self._packed_by_kind = {}

    # ---- packed-weight cache -----------------------------------------------------------------------------------------
    # The packed / compressed weight is a pure function of (weight values, bfp_args).  torch's version counter sees in-place
    # ops on the parameter but NOT writes through `.data` (p.data.copy_(), p.data.mul_(mask): the reference's own BFPOptim,
    # DeepSpeed / apex master-weight copies and pruning scripts all do that), so:
    #   * nothing is cached while the module is in training mode and the weight requires grad -- every forward re-packs, which
    #     is one fused kernel and is what the reference does anyway (it re-quantises the weight on every call);
    #   * in eval mode / for frozen weights the cache is keyed on (storage, version, shape, device, bfp_args) and dropped by
    #     load_state_dict, .to() / .half() / .cuda() and `invalidate_packed()` -- call that after writing through `.data`;
    #   * BFP_WEIGHT_CACHE=0 disables caching, BFP_WEIGHT_CACHE=verify re-checks a checksum of the whole weight on every forward
    #     (one extra read of the weight and a host sync: a debugging aid, not a fast path).
    def invalidate_packed(self):
        """Drops every cached packed form of the weight (call after modifying the weight through `.data`)."""
        self._packed_w = None
        self._packed_by_kind = {}

    def _apply(self, fn, *args, **kwargs):
        self.invalidate_packed()
        return super()._apply(fn, *args, **kwargs)

    def _load_from_state_dict(self, *args, **kwargs):
        self.invalidate_packed()
        return super()._load_from_state_dict(*args, **kwargs)

    def _args_key(self):
        a = self.bfp_args
        return (a['num_format'], a['sparsity_num_format'], a['rounding_mode'], float(a['epsilon']), a['mant_bits'], a['block_size'],
                a['w_sparsity'] == True, a['N'], a['M'], a['first'], a['sparsity_mode'], float(a['sparsity_frac']))     # noqa: E712

    def _cacheable(self):
        mode = os.environ.get("BFP_WEIGHT_CACHE", "1")
        if mode == "0":
            return False
        return not (self.training and self.weight.requires_grad)

    def _packed_weight(self, kind, key_prefix=None):
        w = self.weight
        key = (key_prefix or kind, w.data_ptr(), w._version, tuple(w.shape), w.device, w.dtype, self._args_key())
        if os.environ.get("BFP_WEIGHT_CACHE", "1") == "verify":
            key = key + (int(w.detach().view(torch.int16 if w.element_size() == 2 else torch.int32).sum(dtype=torch.int64).item()),)
        hit = self._packed_by_kind.get(kind) if self._cacheable() else None
        if hit is None or hit[0] != key:
            if kind == 'int':
                packed = _int_pack_weight(w.detach(), self.bfp_args)
            elif kind == 'sp':
                # raises if the pruned weight is not 2:4 (cannot happen for sp_ok configs); the dense form is not kept
                # (the violation count is read back -- one sync -- only when the result is going to be cached)
                packed = compress_2to4_bf16(pack_bfp_bf16(w.detach(), identifier='w', **self.bfp_args), check=self._cacheable())
            elif kind == 'sp_static':
                packed = self._build_static_sparse_weight()
            elif kind == 'mx':
                # folded block-scaled form (None when a row's block exponents span more than the form holds: the caller moves on)
                packed = pack_bfp_mx(w.detach(), 240 if w.shape[0] >= 240 else 128, fold=True, identifier='w', **self.bfp_args)
            else:
                packed = (pack_bfp if kind == 'i8' else pack_bfp_bf16)(w.detach(), identifier='w', **self.bfp_args)
            hit = (key, packed)
            if self._cacheable():
                self._packed_by_kind = {k: v for k, v in self._packed_by_kind.items() if v[0][1:] == key[1:]}   # drop stale kinds
                self._packed_by_kind[kind] = hit
        if kind != 'sp_static' and hit[1] is not None:
            self._packed_w = hit
        return hit[1]

    def _static_sparse_weight(self):
        """first == 's': the N:M mask is taken on the unquantised weight, so it does not depend on the rounding draw.  Cached like
        the other packed forms: the kept values in compressed order [N, Kc] (weight dtype) and the (static) tcgen05 metadata.  A
        block of B weights is B/2 consecutive compressed values holding the block's maximum, so the per-forward stochastic
        quantisation runs on the compressed tensor with block size B/2: half the elements, no mask, no re-compression."""
        return self._packed_weight('sp_static')

    def _build_static_sparse_weight(self):
        w = self.weight
        a = self.bfp_args
        ws = _structured_N_M_sparsity(w.detach(), w.device, a['N'], a['M'])       # stays in the weight's dtype: the block
        # exponent is computed in that dtype's arithmetic
        n_out, K = ws.shape
        K128 = -(-K // 128) * 128
        g = F.pad(ws, (0, K128 - K)).view(n_out, K128 // 4, 4)
        nz = g != 0
        cnt = nz.sum(-1)
        idx = torch.arange(4, device=w.device).expand_as(g)
        i0 = torch.where(nz, idx, 4).min(-1).values
        i1 = torch.where(nz & (idx > i0.unsqueeze(-1)), idx, 4).min(-1).values
        # the compress kernel's padding rule (csrc/bfp_gemm_sp.cu): none -> (0, 1); one -> (i0, 3), or (0, 3) when i0 == 3
        i0f = torch.where((cnt == 0) | ((cnt == 1) & (i0 == 3)), torch.zeros_like(i0), i0)
        i1f = torch.where(cnt == 0, torch.ones_like(i1), torch.where(cnt == 1, torch.full_like(i1, 3), i1))
        kept = torch.stack([g.gather(-1, i0f.unsqueeze(-1)).squeeze(-1), g.gather(-1, i1f.unsqueeze(-1)).squeeze(-1)], -1)
        comp32 = kept.reshape(n_out, K128 // 2).contiguous()
        pattern = nz.view(n_out, K128)[:, :-(-K // 8) * 8].to(torch.bfloat16).contiguous()
        meta = compress_2to4_bf16(pattern).meta                       # raises if the mask is not 2:4 (cannot happen for sp_ok)
        return (comp32, meta)

    def _stochastic_forward(self, input):
        """Inference with rounding_mode='stoc' -- what every script of the reference sets (bfp_config.yaml:4) -- on the tensor
        cores: like the reference, BOTH operands are re-quantised with fresh uniforms on every call (no weight cache, no
        activation cache), then contracted as exact-bf16 operands; a 2:4-pruned weight is re-compressed per call (the mask
        is applied before rounding in either order, so the pattern holds whatever the draw).  Half-precision modules (the
        reference's LLaMA scripts): the reference's stochastic quantiser returns fp32 tensors, so
        its F.linear is an fp32 SGEMM with an fp32 result -- same here, fp32 out; only bias-free modules (a half bias
        would not type-check against fp32 operands in the reference either)."""
        xb = pack_bfp_bf16(input, identifier='in', **self.bfp_args)
        out_shape = tuple(input.shape[:-1]) + (self.out_features,)
        sp = _tensor_core_kind(input, self.weight, self.bfp_args) == 'sp'
        if sp and self.bfp_args['first'] == 's' and self.bfp_args['block_size'] >= 8 and os.environ.get("BFP_STOC_STATIC_MASK", "1") == "1":
            comp32, meta = self._static_sparse_weight()
            wc = pack_bfp_bf16(comp32, identifier='w', **dict(self.bfp_args, w_sparsity=False, block_size=self.bfp_args['block_size'] // 2))
            return bfp_linear_bf16_sp(xb, SparseBF16(wc, meta, self.out_features, xb.shape[1]), self.bias, out_shape=out_shape)
        wb = pack_bfp_bf16(self.weight.detach(), identifier='w', **self.bfp_args)
        if sp:
            return bfp_linear_bf16_sp(xb, compress_2to4_bf16(wb, check=False), self.bias, out_shape=out_shape)
        return bfp_linear_bf16(xb, wb, self.bias, out_shape=out_shape)

    def forward(self, input):
        if self.num_format == 'fp32':
            return F.linear(input, self.weight, self.bias)
        elif self.num_format == 'bfp':
            determ = self.bfp_args['rounding_mode'] == rounding_modes.DETERM
            # a trainable bias alone (BitFit, frozen backbones) also needs the autograd path: the inference kernels take bias.detach()
            training = torch.is_grad_enabled() and (input.requires_grad or self.weight.requires_grad
                                                    or (self.bias is not None and self.bias.requires_grad))
            if not training and _int_tc_eligible(input, self.weight, self.bfp_args):
                packed = self._packed_weight('int')
                if packed is not None:
                    return _int_tc_linear(input, packed, self.bias, self.bfp_args)
            kind = _tensor_core_kind(input, self.weight, self.bfp_args) if (determ or training) else None
            if (training and kind is not None and self.bfp_args['mant_bits'] <= 8
                    and (self.bfp_args['grad_sparsity'] != True or self.bfp_args['sparsity_mode'] == 'structured')   # noqa: E712
                    and input.dtype == self.weight.dtype and input.dtype in _DT
                    and (determ or input.dtype == torch.float32)):
                # training: forward + dgrad + wgrad on the tensor cores.  Stochastic rounding re-quantises the weight on
                # every forward like the reference (no cache), and keeps it dense for the backward contraction over N.
                tkind = kind if kind in ('sp', 'bf16') else 'bf16'
                cached = self._packed_weight(tkind) if determ else None
                return _BFPLinearTC.apply(input, self.weight, self.bias, self.bfp_args, (lambda: self._packed_weight('bf16')), cached)
            if (not determ and not training and self.bfp_args['mant_bits'] <= 8
                    and (self.bias is None or (input.dtype == torch.float32 and self.weight.dtype == torch.float32))
                    # ^ stoc promotes both operands to fp32 and returns fp32: a half-precision bias would not type-check
                    and _tensor_core_kind(input, self.weight, self.bfp_args) is not None):
                return self._stochastic_forward(input)
            if not determ or training:
                kind = None             # training configurations the autograd Function does not cover keep the reference's structure
            y = None
            if kind is not None and _mx_eligible(input, self.weight, self.bfp_args):
                # HBFP4 / HBFP5: block-scaled FP8-class MMA (twice the bf16 rate); zeros of a pruned weight are just zeros here
                wp = self._packed_weight('mx')
                if wp is not None:
                    y = bfp_linear_mx(_packed_activation_mx(input, self.bfp_args), wp, self.bias,
                                      out_shape=tuple(input.shape[:-1]) + (self.out_features,))
                    kind = 'mx'
            if y is not None:
                pass
            elif kind == 'i8':
                # inference fast path: pack activations on the fly, cached packed weight, tcgen05 int8 BFP GEMM
                y = bfp_linear_packed(pack_bfp(input, identifier='in', **self.bfp_args), self._packed_weight(kind), self.bias)
            elif kind == 'sp':
                # 2:4-pruned weight: compressed once, tcgen05.mma.sp skips the zeros
                y = bfp_linear_bf16_sp(_packed_activation(input, self.bfp_args), self._packed_weight(kind), self.bias,
                                       out_shape=tuple(input.shape[:-1]) + (self.out_features,), out_dtype=input.dtype)
            elif kind == 'bf16':
                y = bfp_linear_bf16(_packed_activation(input, self.bfp_args), self._packed_weight(kind), self.bias,
                                    out_shape=tuple(input.shape[:-1]) + (self.out_features,), out_dtype=input.dtype)
            if y is not None:
                return y if input.dtype == torch.float32 else y.to(input.dtype)     # fp32 accumulation, one rounding to the dtype
            return self.linear_op(input, self.weight, self.bias)
        raise NotImplementedError('NumFormat not implemented')


class BFPConv1D(torch.nn.Module):
    """A name the reference's callers import but bfp_ops.py never defines (modeling_gpt2.py:58 imports it; :173-181 and
    :580-581 construct `BFPConv1D(nf, nx, **bfp_args)`, so the reference's GPT-2 cannot be imported as published).  Provided
    as the repair SURVEY.md section 8(b) lists: Hugging Face's Conv1D -- weight [nx, nf], y = x @ weight + bias -- with the
    product taken by F_matmul_bfp (bfp_ops.py:240-245), i.e. both operands BFP-quantised along the contraction exactly like
    BFPLinear does for the transposed weight; 'fp32' format = plain Conv1D.  The bias is never quantised."""

    def __init__(self, nf, nx, **kwargs):
        super().__init__()
        self.bfp_args = unpack_bfp_args(kwargs)
        self.num_format = self.bfp_args['num_format']
        self.nf = nf
        self.weight = torch.nn.Parameter(torch.empty(nx, nf))
        self.bias = torch.nn.Parameter(torch.zeros(nf))
        torch.nn.init.normal_(self.weight, std=0.02)
        self.matmul_op = _get_bfp_op(torch.matmul, 'matmul', self.bfp_args, transpose=True)

    def forward(self, x):
        size_out = x.size()[:-1] + (self.nf,)
        x2 = x.reshape(-1, x.size(-1))
        if self.num_format == 'fp32':
            y = torch.addmm(self.bias, x2, self.weight)
        elif self.num_format == 'bfp':
            y = self.matmul_op(x2, self.weight) + self.bias
        else:
            raise NotImplementedError('NumFormat not implemented')
        return y.view(size_out)
