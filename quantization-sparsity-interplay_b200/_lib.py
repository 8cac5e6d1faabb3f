"""ctypes binding of libbfp_b200.so (C ABI: include/bfp_b200.h).  Plain pointers and sizes only."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BFP_B200_LIB") or os.path.join(_HERE, "libbfp_b200.so")     # the override is for A/B runs of kernel variants (tools/)
CSRC = os.path.join(_HERE, "csrc")

DT_F32, DT_F16, DT_BF16 = 0, 1, 2
ROUND_NEAREST, ROUND_STOCHASTIC = 0, 1
ORDER_QUANT_ONLY, ORDER_SPARSIFY_QUANT, ORDER_QUANT_SPARSIFY, ORDER_SPARSIFY_ONLY = 0, 1, 2, 3
TIE_TORCH_CUDA, TIE_TORCH_CPU = 0, 1
OK, E_ARG, E_UNSUPPORTED, E_CUDA, E_ALIGN = 0, 1, 2, 3, 4

_lib = None


class BFPLibraryError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libbfp_b200 error {code}: {msg}")
        self.code = code


def build(verbose=False, jobs=8):
    """Compiles csrc/*.cu for sm_100a with nvcc (cross-compiles without a GPU) into libbfp_b200.so."""
    r = subprocess.run(["make", "-C", CSRC, f"-j{jobs}"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("building libbfp_b200.so failed")
    return LIB_PATH


_i64, _i32, _f32, _vp, _u64 = ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_uint64

# name -> (restype, argtypes); must list every symbol include/bfp_b200.h declares (tests/test_abi.py checks)
SIGNATURES = {
    "bfp_version": (_i32, []),
    "bfp_last_error": (ctypes.c_char_p, []),
    "bfp_launch_count": (_u64, []),
    "bfp_set_option": (_i32, [ctypes.c_char_p, _i64]),
    "bfp_device_info": (_i32, [ctypes.POINTER(_i32)] * 3 + [ctypes.POINTER(ctypes.c_size_t)]),
    "bfp_quantize": (_i32, [_vp, _vp, _i64, _i64, _i32, _i32, _i32, _i32, _f32, _i32, _u64, _u64, _i32, _i32, _i32, _i32, _vp]),
    "bfp_nm_sparsify": (_i32, [_vp, _vp, _i64, _i64, _i32, _i32, _i32, _i32, _vp]),
    "bfp_int_workspace_bytes": (ctypes.c_size_t, [_i64]),
    "bfp_int_quantize": (_i32, [_vp, _vp, _i64, _i64, _i64, _i32, _i32, _vp, _vp]),
    "bfp_int_quantize_split3": (_i32, [_vp, _vp, _i64, _i64, _i64, _i32, _i32, _vp, _vp]),
    "bfp_int_quantize_nm": (_i32, [_vp, _vp, _i64, _i64, _i32, _i32, _i32, _i32, _i32, _vp]),
    "bfp_unstructured_workspace_bytes": (ctypes.c_size_t, []),
    "bfp_unstructured_sparsify": (_i32, [_vp, _vp, _i64, _i32, _u64, _vp, _vp]),
    "bfp_mx_layout": (_i32, [_i64, _i64, _i32, _i32, ctypes.POINTER(_i64), ctypes.POINTER(_i64)]),
    "bfp_mx_from_packed": (_i32, [_vp, _vp, _i64, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "bfp_quantize_pack_mx": (_i32, [_vp, _vp, _vp, _i64, _i64, _i32, _i32, _i32, _f32, _vp]),
    "bfp_gemm_mx": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _i64, _i64, _i64, _vp]),
    "bfp_bfloat_round": (_i32, [_vp, _vp, _vp, _i64, _i64, _i32, _i32, _vp]),
    "bfp_ocp_mx_quantize": (_i32, [_vp, _vp, _i64, _i64, _i32, _i32, _i64, _i32, _i32, _i32, _i32, _i32, _vp]),
    "bfp_ocp_mx_pack": (_i32, [_vp, _vp, _vp, _i64, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "bfp_gemm_mx_round": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _i64, _i64, _i64, _i32, _vp]),
    "bfp_unstructured_quantize_workspace_bytes": (ctypes.c_size_t, [_i64, _i32]),
    "bfp_unstructured_quantize": (_i32, [_vp, _vp, _i64, _i64, _i32, _i32, _u64, _i32, _i32, _i32, _f32, _i32, _u64, _u64, _vp,
                                         ctypes.c_size_t, _vp]),
    "bfp_block_exponent": (_i32, [_vp, _vp, _i64, _i64, _i32, _i32, _f32, _vp]),
    "bfp_quantize_host": (_i32, [_vp, _vp, _i64, _i64, _i32, _i32, _i32, _i32, _f32, _i32, _u64, _u64, _i32, _i32, _i32, _i32]),
    "bfp_host_staging_release": (_i32, []),
    "bfp_debug_cpu_tie_lut": (_i32, [_vp]),
    "bfp_debug_exp_table": (_i32, [_i32, _vp]),
    "bfp_packed_layout": (_i32, [_i64, _i64, _i32] + [ctypes.POINTER(_i64)] * 3),
    "bfp_quantize_pack": (_i32, [_vp, _vp, _vp, _i64, _i64, _i32, _i32, _i32, _f32, _i32, _u64, _u64, _i32, _i32, _i32, _vp]),
    "bfp_unpack": (_i32, [_vp, _vp, _vp, _i64, _i64, _i32, _vp]),
    "bfp_quantize_pack_bf16": (_i32, [_vp, _vp, _i64, _i64, _i32, _i32, _i32, _f32, _i32, _u64, _u64, _i32, _i32, _i32, _vp]),
    "bfp_gemm_bf16": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp]),
    "bfp_gemm_bf16_ex": (_i32, [_vp, _vp, _vp, _vp, _i32, _i64, _i64, _i64, _vp]),
    "bfp_transpose_pad_16": (_i32, [_vp, _vp, _i64, _i64, _i64, _i64, _vp]),
    "bfp_gemm_bf16_batched": (_i32, [_vp, _vp, _vp, _i32, _i64, _i64, _i64, _i64, _vp]),
    "bfp_gemm_bf16_acc": (_i32, [_vp, _vp, _vp, _i64, _i64, _i64, _vp]),
    "bfp_gemm_bf16_sp_acc": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp]),
    "bfp_sp_layout": (_i32, [_i64, _i64] + [ctypes.POINTER(_i64)] * 2),
    "bfp_compress_2to4_bf16": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp, _vp]),
    "bfp_gemm_bf16_sp": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp]),
    "bfp_gemm_bf16_sp_gather": (_i32, [_vp, _vp, _vp, _vp, ctypes.POINTER(_vp), _i32, _i32, _i64, _i64, _i64, _i64, _vp]),
    "bfp_gemm_bf16_sp_ex": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _i64, _i64, _i64, _i64, _vp]),
    "bfp_gemm_i8": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i32, _vp]),
}


def lib():
    """The loaded library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(nvcc, sm_100a).  There is no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise BFPLibraryError(rc, lib().bfp_last_error().decode())


def launch_count():
    return int(lib().bfp_launch_count())


def set_option(name, value):
    check(lib().bfp_set_option(name.encode(), int(value)))


def packed_layout(rows, K, block_size):
    """(Kp, rows_pad, nkb_pad) of the packed form of a [rows, K] tensor (include/bfp_b200.h)."""
    kp, rp, nk = _i64(), _i64(), _i64()
    check(lib().bfp_packed_layout(int(rows), int(K), int(block_size), ctypes.byref(kp), ctypes.byref(rp), ctypes.byref(nk)))
    return kp.value, rp.value, nk.value


def sp_layout(rows, K):
    """(Kc, meta_bytes) of the 2:4-compressed form of a [rows, K] bf16 operand (include/bfp_b200.h)."""
    kc, mb = _i64(), _i64()
    check(lib().bfp_sp_layout(int(rows), int(K), ctypes.byref(kc), ctypes.byref(mb)))
    return kc.value, mb.value
