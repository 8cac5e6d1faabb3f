// bfp_abi.cu -- the extern "C" surface declared in include/bfp_b200.h: argument validation (mirroring the
// reference's asserts / exceptions), error reporting, device bookkeeping.  No torch types, no allocation.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "bfp_internal.h"

namespace bfp {

static thread_local char t_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

int set_error(int code, const char* msg) {
    snprintf(t_err, sizeof(t_err), "%s", msg ? msg : "");
    return code;
}
int set_errorf(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
    return code;
}
int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_errorf(BFP_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return BFP_OK;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

Tuning& tuning() {
    static Tuning t = [] {
        Tuning x;
        if (const char* s = getenv("BFP_STREAM_CTAS_PER_SM")) x.stream_ctas_per_sm = atoi(s) > 0 ? atoi(s) : x.stream_ctas_per_sm;
        if (const char* s = getenv("BFP_FORCE_GENERIC")) x.force_generic = atoi(s);
        if (const char* s = getenv("BFP_QUANT_TMA")) x.quant_tma = atoi(s) ? 1 : 0;
        if (const char* s = getenv("BFP_HOST_CHUNK_MB")) x.host_chunk_bytes = (int64_t)(atoi(s) > 0 ? atoi(s) : 16) << 20;
        return x;
    }();
    return t;
}

const DeviceInfo& device_info() {
    static DeviceInfo cache[64];
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
        static DeviceInfo none;
        return none;
    }
    std::lock_guard<std::mutex> lk(mu);
    DeviceInfo& d = cache[dev];
    if (d.device != dev) {
        int v = 0;
        cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev);
        cudaDeviceGetAttribute(&d.cc_minor, cudaDevAttrComputeCapabilityMinor, dev);
        cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, dev);
        d.l2_bytes = (size_t)v;
        d.device = dev;
    }
    return d;
}

static int require_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return set_error(BFP_E_CUDA, "no CUDA device: libbfp_b200 has no CPU fallback");
    }
    const DeviceInfo& d = device_info();
    if (d.cc_major != 10)
        return set_errorf(BFP_E_CUDA, "device is sm_%d%d; libbfp_b200 is built for sm_100a only", d.cc_major, d.cc_minor);
    return BFP_OK;
}

int validate_quant_args(const QuantArgs& a, bool /*device_pointers*/) {
    if (a.rows < 0 || a.K < 0) return set_error(BFP_E_ARG, "negative shape");
    if (a.in_dtype < 0 || a.in_dtype > 2 || a.out_dtype < 0 || a.out_dtype > 2) return set_error(BFP_E_ARG, "bad dtype");
    if (a.order < 0 || a.order > 3) return set_error(BFP_E_ARG, "bad order");
    if (a.rounding != BFP_ROUND_NEAREST && a.rounding != BFP_ROUND_STOCHASTIC)
        return set_error(BFP_E_ARG, "Rounding mode is not implemented");                         // bfp_ops.py:27
    const bool quant = a.order != BFP_ORDER_SPARSIFY_ONLY, sparse = a.order != BFP_ORDER_QUANT_ONLY;
    if (quant) {
        if (a.B <= 0) return set_error(BFP_E_ARG, "block_size must be > 0 for the bfp format");  // bfp_ops.py:130
        if (a.m < 0 || a.m > 23) return set_error(BFP_E_ARG, "mant_bits must be in [0, 23]");
    }
    if (sparse) {
        if (!(a.N > 0 && a.M > 0 && a.N <= a.M)) return set_error(BFP_E_ARG, "need 0 < N <= M");  // bfp_ops.py:74
        if (a.M > 64) return set_error(BFP_E_UNSUPPORTED, "N:M groups larger than 64 are not supported");
        if (a.tie != BFP_TIE_TORCH_CUDA && a.tie != BFP_TIE_TORCH_CPU) return set_error(BFP_E_ARG, "bad tie_rule");
    }
    const bool stoc = quant && a.rounding == BFP_ROUND_STOCHASTIC;
    const int want_out = stoc ? BFP_DT_F32 : a.in_dtype;
    if (a.out_dtype != want_out)
        return set_error(BFP_E_ARG, "out_dtype must equal in_dtype for nearest rounding and be fp32 for stochastic rounding");
    if (a.rows * a.K > 0) {
        if (!a.in || !a.out) return set_error(BFP_E_ARG, "null pointer");
        if (reinterpret_cast<uintptr_t>(a.in) % dtype_size(a.in_dtype) || reinterpret_cast<uintptr_t>(a.out) % dtype_size(a.out_dtype))
            return set_error(BFP_E_ALIGN, "pointer not aligned to its element size");
        if (a.in == a.out && !(a.order == BFP_ORDER_QUANT_ONLY && a.in_dtype == a.out_dtype))
            return set_error(BFP_E_ARG, "in-place is only supported for quantise-only with equal dtypes");
    }
    return BFP_OK;
}

}  // namespace bfp

using namespace bfp;

extern "C" {

int bfp_version(void) { return BFP_B200_VERSION; }
const char* bfp_last_error(void) { return t_err; }
uint64_t bfp_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int bfp_set_option(const char* name, int64_t value) {
    if (!name) return set_error(BFP_E_ARG, "null option name");
    Tuning& t = tuning();
    if (!strcmp(name, "stream_ctas_per_sm") && value > 0 && value <= 32) t.stream_ctas_per_sm = (int)value;
    else if (!strcmp(name, "force_generic")) t.force_generic = value != 0;
    else if (!strcmp(name, "host_chunk_bytes") && value >= 4096) t.host_chunk_bytes = value;
    else if (!strcmp(name, "host_chunk_min_bytes") && value >= 4096) t.host_chunk_min_bytes = value;
    else if (!strcmp(name, "quant_tma") && (value == 0 || value == 1)) t.quant_tma = (int)value;
    else if (!strcmp(name, "pdl") && (value == 0 || value == 1)) t.pdl = (int)value;
    else if (!strcmp(name, "gemm_sp_debug") && value >= 0 && value <= 31) t.gemm_sp_debug = (int)value;
    else if (!strcmp(name, "gemm_bf16_cta_group") && value >= 0 && value <= 2) t.gemm_bf16_cta_group = (int)value;
    else if (!strcmp(name, "gemm_out_tma") && (value == 0 || value == 1)) t.gemm_out_tma = (int)value;
    else if (!strcmp(name, "gemm_sp_tile") && (value == 0 || value == 240 || value == 256 || value == 480)) t.gemm_sp_tile = (int)value;
    else if (!strcmp(name, "gemm_sp_cta_group") && value >= 0 && value <= 2) t.gemm_sp_cta_group = (int)value;
    else if (!strcmp(name, "gemm_bf16_tile_n") && (value == 0 || value == 128 || value == 256)) t.gemm_bf16_tile_n = (int)value;
    else if (!strcmp(name, "gemm_mx_variant") && value >= 0 && value <= 255) t.gemm_mx_variant = (int)value;
    else if (!strcmp(name, "unstructured_force_fallback") && (value == 0 || value == 1)) t.unstructured_force_fallback = (int)value;
    else return set_errorf(BFP_E_ARG, "unknown option or bad value: %s", name);
    return BFP_OK;
}

int bfp_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* l2_bytes) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return set_error(BFP_E_CUDA, "no CUDA device");
    }
    const DeviceInfo& d = device_info();
    if (sm_count) *sm_count = d.sm_count;
    if (cc_major) *cc_major = d.cc_major;
    if (cc_minor) *cc_minor = d.cc_minor;
    if (l2_bytes) *l2_bytes = d.l2_bytes;
    return BFP_OK;
}

int bfp_quantize(const void* in, void* out, int64_t rows, int64_t K, int in_dtype, int out_dtype, int block_size,
                 int mant_bits, float eps, int rounding, uint64_t seed, uint64_t offset, int N, int M, int order,
                 int tie_rule, void* stream) {
    QuantArgs a{in, out, rows, K, in_dtype, out_dtype, block_size, mant_bits, eps, rounding, seed, offset, N, M, order, tie_rule};
    if (int rc = validate_quant_args(a, true)) return rc;
    if (int rc = require_device()) return rc;
    return quantize_device(a, static_cast<cudaStream_t>(stream));
}

int bfp_nm_sparsify(const void* in, void* out, int64_t rows, int64_t K, int dtype, int N, int M, int tie_rule, void* stream) {
    return bfp_quantize(in, out, rows, K, dtype, dtype, 0, 0, 0.0f, BFP_ROUND_NEAREST, 0, 0, N, M, BFP_ORDER_SPARSIFY_ONLY,
                        tie_rule, stream);
}

size_t bfp_int_workspace_bytes(int64_t C) { return int_workspace_bytes(C < 0 ? 0 : C); }

int bfp_int_quantize(const void* in, float* out, int64_t A, int64_t C, int64_t inner, int dtype, int bits, void* workspace, void* stream) {
    if (A < 0 || C < 0 || inner < 0 || dtype < 0 || dtype > 2) return set_error(BFP_E_ARG, "bad argument");
    if (bits < 1 || bits > 23) return set_error(BFP_E_ARG, "bits must be in [1, 23]");
    if (A * C * inner > 0 && (!in || !out || !workspace)) return set_error(BFP_E_ARG, "null pointer");
    if (reinterpret_cast<uintptr_t>(workspace) % 4) return set_error(BFP_E_ALIGN, "workspace must be 4-byte aligned");
    if (int rc = require_device()) return rc;
    return int_quantize_device(in, out, A, C, inner, dtype, bits, workspace, static_cast<cudaStream_t>(stream));
}

int bfp_int_quantize_split3(const void* in, void* out_bf16, int64_t A, int64_t C, int64_t kseg, int dtype, int bits, void* workspace, void* stream) {
    if (A < 0 || C < 0 || dtype < 0 || dtype > 2 || kseg <= 0 || kseg % 8) return set_error(BFP_E_ARG, "bad argument (kseg: a positive multiple of 8)");
    if (bits < 1 || bits > 23) return set_error(BFP_E_ARG, "bits must be in [1, 23]");
    if (C % 8) return set_error(BFP_E_UNSUPPORTED, "the three-plane form needs C to be a multiple of 8");
    if (A * C > 0 && (!in || !out_bf16 || !workspace)) return set_error(BFP_E_ARG, "null pointer");
    if (reinterpret_cast<uintptr_t>(workspace) % 16 || reinterpret_cast<uintptr_t>(out_bf16) % 16)
        return set_error(BFP_E_ALIGN, "workspace and out_bf16 must be 16-byte aligned");
    if (int rc = require_device()) return rc;
    return int_quantize_split3_device(in, out_bf16, A, C, kseg, dtype, bits, workspace, static_cast<cudaStream_t>(stream));
}

int bfp_int_quantize_nm(const void* in, float* out, int64_t C, int64_t K, int dtype, int bits, int N, int M, int order, void* stream) {
    if (C < 0 || K < 0 || dtype < 0 || dtype > 2) return set_error(BFP_E_ARG, "bad argument");
    if (bits < 1 || bits > 23) return set_error(BFP_E_ARG, "bits must be in [1, 23]");
    if (order != BFP_ORDER_SPARSIFY_QUANT && order != BFP_ORDER_QUANT_SPARSIFY) return set_error(BFP_E_ARG, "order must be s->q or q->s");
    if (M != 4 || N < 1 || N > 3) return set_error(BFP_E_UNSUPPORTED, "fused INT + N:M supports M == 4 with 0 < N < 4; compose bfp_nm_sparsify and bfp_int_quantize otherwise");
    if (C * K > 0 && (!in || !out)) return set_error(BFP_E_ARG, "null pointer");
    if (int rc = require_device()) return rc;
    return int_quantize_nm_device(in, out, C, K, dtype, bits, N, order, static_cast<cudaStream_t>(stream));
}

size_t bfp_unstructured_workspace_bytes(void) { return unstructured_workspace_bytes(); }

int bfp_unstructured_sparsify(const void* in, void* out, int64_t numel, int dtype, uint64_t k, void* workspace, void* stream) {
    if (numel < 0 || dtype < 0 || dtype > 2) return set_error(BFP_E_ARG, "bad argument");
    if (numel > 0 && (!in || !out || !workspace)) return set_error(BFP_E_ARG, "null pointer");
    if (reinterpret_cast<uintptr_t>(workspace) % 8) return set_error(BFP_E_ALIGN, "workspace must be 8-byte aligned");
    if (int rc = require_device()) return rc;
    return unstructured_device(in, out, numel, dtype, k, workspace, static_cast<cudaStream_t>(stream));
}

size_t bfp_unstructured_quantize_workspace_bytes(int64_t numel, int dtype) { return unstructured_fused_workspace_bytes(numel, dtype); }

int bfp_unstructured_quantize(const void* in, void* out, int64_t rows, int64_t K, int in_dtype, int out_dtype, uint64_t k, int order,
                              int block_size, int mant_bits, float eps, int rounding, uint64_t seed, uint64_t offset, void* workspace,
                              size_t workspace_bytes, void* stream) {
    if (rows < 0 || K < 0 || in_dtype < 0 || in_dtype > 2 || out_dtype < 0 || out_dtype > 2) return set_error(BFP_E_ARG, "bad argument");
    if (order != BFP_ORDER_SPARSIFY_ONLY && order != BFP_ORDER_SPARSIFY_QUANT && order != BFP_ORDER_QUANT_SPARSIFY)
        return set_error(BFP_E_ARG, "order must be SPARSIFY_ONLY, SPARSIFY_QUANT or QUANT_SPARSIFY");
    const bool quant = order != BFP_ORDER_SPARSIFY_ONLY;
    if (rounding != BFP_ROUND_NEAREST && rounding != BFP_ROUND_STOCHASTIC) return set_error(BFP_E_ARG, "Rounding mode is not implemented");   // bfp_ops.py:27
    if (quant && block_size <= 0) return set_error(BFP_E_ARG, "block_size must be > 0 for the bfp format");                         // bfp_ops.py:130
    if (quant && (mant_bits < 0 || mant_bits > 23)) return set_error(BFP_E_ARG, "mant_bits must be in [0, 23]");
    const int want_out = (quant && rounding == BFP_ROUND_STOCHASTIC) ? BFP_DT_F32 : in_dtype;
    if (out_dtype != want_out) return set_error(BFP_E_ARG, "out_dtype must equal in_dtype for nearest rounding and be fp32 for stochastic rounding");
    const int64_t n = rows * K;
    if (n == 0) return BFP_OK;
    if (!in || !out || !workspace) return set_error(BFP_E_ARG, "null pointer");
    if (in == out) return set_error(BFP_E_ARG, "out must not alias in");
    if (k == 0 || k >= (uint64_t)n) return set_error(BFP_E_ARG, "k must be in (0, numel): nothing or everything dropped needs no selection");
    if (workspace_bytes < unstructured_fused_workspace_bytes(n, in_dtype)) return set_error(BFP_E_ARG, "workspace smaller than bfp_unstructured_quantize_workspace_bytes()");
    UnstructuredArgs a{in, out, workspace, n, K, in_dtype, (unsigned long long)k, order, block_size, mant_bits, eps,
                       quant ? rounding : BFP_ROUND_NEAREST, seed, offset};
    if (!unstructured_fused_supported(a))
        return set_error(BFP_E_UNSUPPORTED, "fused unstructured pruning needs 16-byte aligned buffers, numel a multiple of the 128-bit vector and K a "
                                            "multiple of a power-of-two block_size; compose bfp_unstructured_sparsify and bfp_quantize otherwise");
    if (int rc = require_device()) return rc;
    return unstructured_fused_device(a, static_cast<cudaStream_t>(stream));
}

int bfp_block_exponent(const void* in, float* exp_out, int64_t rows, int64_t K, int dtype, int block_size, float eps, void* stream) {
    if (rows < 0 || K < 0 || block_size <= 0 || dtype < 0 || dtype > 2) return set_error(BFP_E_ARG, "bad argument");
    if (rows * K > 0 && (!in || !exp_out)) return set_error(BFP_E_ARG, "null pointer");
    if (int rc = require_device()) return rc;
    return block_exponent_device(in, exp_out, rows, K, dtype, block_size, eps, static_cast<cudaStream_t>(stream));
}

int bfp_quantize_host(const void* host_in, void* host_out, int64_t rows, int64_t K, int in_dtype, int out_dtype,
                      int block_size, int mant_bits, float eps, int rounding, uint64_t seed, uint64_t offset, int N,
                      int M, int order, int tie_rule) {
    QuantArgs a{host_in, host_out, rows, K, in_dtype, out_dtype, block_size, mant_bits, eps, rounding, seed, offset, N, M, order, tie_rule};
    if (int rc = validate_quant_args(a, false)) return rc;
    if (int rc = require_device()) return rc;
    return quantize_host(a);
}

int bfp_host_staging_release(void) { return host_staging_release(); }

int bfp_packed_layout(int64_t rows, int64_t K, int block_size, int64_t* Kp, int64_t* rows_pad, int64_t* nkb_pad) {
    if (rows < 0 || K < 0 || block_size <= 0) return set_error(BFP_E_ARG, "bad argument");
    if (Kp) *Kp = packed_kp(K);
    if (rows_pad) *rows_pad = packed_rows_pad(rows);
    if (nkb_pad) *nkb_pad = std::max<int64_t>(packed_nkb_pad(K, block_size), (K + block_size - 1) / block_size);
    return BFP_OK;
}

int bfp_quantize_pack(const void* in, int8_t* mant, float* scale_t, int64_t rows, int64_t K, int in_dtype, int block_size,
                      int mant_bits, float eps, int rounding, uint64_t seed, uint64_t offset, int N, int M, int order, void* stream) {
    const int out_dt = rounding == BFP_ROUND_STOCHASTIC ? BFP_DT_F32 : in_dtype;
    QuantArgs a{in, mant, rows, K, in_dtype, out_dt, block_size, mant_bits, eps, rounding, seed, offset, N, M, order, BFP_TIE_TORCH_CUDA};
    if (order == BFP_ORDER_SPARSIFY_ONLY) return set_error(BFP_E_ARG, "packing needs a quantising order");
    if (int rc = validate_quant_args(a, true)) return rc;
    if (mant_bits < 1 || mant_bits > 7) return set_error(BFP_E_UNSUPPORTED, "packed mantissas are int8: mant_bits must be in [1, 7]");
    if (rows * K > 0 && !scale_t) return set_error(BFP_E_ARG, "null pointer");
    if (reinterpret_cast<uintptr_t>(scale_t) % 4) return set_error(BFP_E_ALIGN, "scale_t not aligned");
    if (int rc = require_device()) return rc;
    return pack_device(a, mant, scale_t, packed_kp(K), packed_rows_pad(rows), static_cast<cudaStream_t>(stream));
}

int bfp_quantize_pack_bf16(const void* in, void* out_bf16, int64_t rows, int64_t K, int in_dtype, int block_size, int mant_bits,
                           float eps, int rounding, uint64_t seed, uint64_t offset, int N, int M, int order, void* stream) {
    const int out_dt = rounding == BFP_ROUND_STOCHASTIC ? BFP_DT_F32 : in_dtype;
    QuantArgs a{in, out_bf16, rows, K, in_dtype, out_dt, block_size, mant_bits, eps, rounding, seed, offset, N, M, order, BFP_TIE_TORCH_CUDA};
    if (order == BFP_ORDER_SPARSIFY_ONLY) return set_error(BFP_E_ARG, "packing needs a quantising order");
    if (int rc = validate_quant_args(a, true)) return rc;
    if (mant_bits < 1 || mant_bits > 8) return set_error(BFP_E_UNSUPPORTED, "bf16 operands are exact for mant_bits in [1, 8] only");
    if (reinterpret_cast<uintptr_t>(out_bf16) % 16) return set_error(BFP_E_ALIGN, "out_bf16 must be 16-byte aligned");
    if (int rc = require_device()) return rc;
    return pack_bf16_device(a, out_bf16, round_up(K, 8), static_cast<cudaStream_t>(stream));
}

int bfp_gemm_bf16(const void* a_bf16, const void* b_bf16, const float* bias, float* out, int64_t T, int64_t N, int64_t K, void* stream) {
    if (T < 0 || N < 0 || K <= 0) return set_error(BFP_E_ARG, "bad argument");
    if (T * N > 0 && (!a_bf16 || !b_bf16 || !out)) return set_error(BFP_E_ARG, "null pointer");
    if (int rc = require_device()) return rc;
    return gemm_bf16_device(a_bf16, b_bf16, bias, out, T, N, round_up(K, 8), static_cast<cudaStream_t>(stream));
}

int bfp_gemm_bf16_ex(const void* a_bf16, const void* b_bf16, const float* bias, void* out, int out_dtype, int64_t T, int64_t N, int64_t K,
                     void* stream) {
    if (T < 0 || N < 0 || K <= 0) return set_error(BFP_E_ARG, "bad argument");
    if (T * N > 0 && (!a_bf16 || !b_bf16 || !out)) return set_error(BFP_E_ARG, "null pointer");
    if (int rc = require_device()) return rc;
    return gemm_bf16_ex_device(a_bf16, b_bf16, bias, out, out_dtype, T, N, round_up(K, 8), static_cast<cudaStream_t>(stream));
}

int bfp_transpose_pad_16(const void* in, void* out, int64_t rows, int64_t cols, int64_t ld_in, int64_t ld_out, void* stream) {
    if (rows < 0 || cols < 0) return set_error(BFP_E_ARG, "bad argument");
    if (rows * cols > 0 && (!in || !out)) return set_error(BFP_E_ARG, "null pointer");
    if (int rc = require_device()) return rc;
    return transpose16_device(in, out, rows, cols, ld_in, ld_out, static_cast<cudaStream_t>(stream));
}

int bfp_gemm_bf16_batched(const void* a_bf16, const void* b_bf16, void* out, int out_dtype, int64_t batch, int64_t T, int64_t N, int64_t K,
                          void* stream) {
    if (batch < 0 || T < 0 || N < 0 || K <= 0) return set_error(BFP_E_ARG, "bad argument");
    if (batch * T * N == 0) return BFP_OK;
    if (!a_bf16 || !b_bf16 || !out) return set_error(BFP_E_ARG, "null pointer");
    if (K % 8) return set_error(BFP_E_ARG, "batched operands are contiguous [batch, rows, K]: K must be a multiple of 8");
    if (int rc = require_device()) return rc;
    return gemm_bf16_ex_device(a_bf16, b_bf16, nullptr, out, out_dtype, T, N, K, static_cast<cudaStream_t>(stream), 0, batch);
}

int bfp_gemm_bf16_acc(const void* a_bf16, const void* b_bf16, float* out, int64_t T, int64_t N, int64_t K, void* stream) {
    if (T < 0 || N < 0 || K <= 0) return set_error(BFP_E_ARG, "bad argument");
    if (T * N > 0 && (!a_bf16 || !b_bf16 || !out)) return set_error(BFP_E_ARG, "null pointer");
    if (int rc = require_device()) return rc;
    return gemm_bf16_ex_device(a_bf16, b_bf16, nullptr, out, BFP_DT_F32, T, N, round_up(K, 8), static_cast<cudaStream_t>(stream), 1);
}

int bfp_gemm_bf16_sp_acc(const void* x_bf16, const void* w_comp, const void* w_meta, float* out, int64_t T, int64_t N, int64_t K, void* stream) {
    if (T < 0 || N < 0 || K <= 0) return set_error(BFP_E_ARG, "bad argument");
    if (T * N > 0 && (!x_bf16 || !w_comp || !w_meta || !out)) return set_error(BFP_E_ARG, "null pointer");
    if (int rc = require_device()) return rc;
    void* outs[1] = {out};
    return gemm_bf16_sp_multi_device(x_bf16, w_comp, w_meta, nullptr, outs, 1, BFP_DT_F32, N, T, N, round_up(K, 8), static_cast<cudaStream_t>(stream), 1);
}

int bfp_mx_layout(int64_t rows, int64_t K, int tile_rows, int fold, int64_t* Kp, int64_t* sf_bytes) { return mx_layout(rows, K, tile_rows, fold, Kp, sf_bytes); }

int bfp_mx_from_packed(const int8_t* mant, const float* scale_t, int64_t rows, int64_t K, int block_size, int tile_rows, int fold, void* vals, void* sf,
                       int32_t* row_ref, uint32_t* violations, void* stream) {
    if (rows < 0 || K < 0 || block_size <= 0 || tile_rows < 1) return set_error(BFP_E_ARG, "bad argument");
    if (rows * K > 0 && (!mant || !scale_t || !vals || !sf || !violations)) return set_error(BFP_E_ARG, "null pointer");
    if (reinterpret_cast<uintptr_t>(mant) % 16 || reinterpret_cast<uintptr_t>(vals) % 16 || reinterpret_cast<uintptr_t>(sf) % 16)
        return set_error(BFP_E_ALIGN, "mant, vals and sf must be 16-byte aligned");
    if (int rc = require_device()) return rc;
    return mx_from_packed_device(mant, scale_t, packed_rows_pad(rows), rows, K, block_size, tile_rows, fold, static_cast<uint8_t*>(vals),
                                 static_cast<uint8_t*>(sf), row_ref, violations, static_cast<cudaStream_t>(stream));
}

int bfp_quantize_pack_mx(const void* in, void* vals, void* sf, int64_t rows, int64_t K, int in_dtype, int block_size, int mant_bits, float eps, void* stream) {
    if (rows < 0 || K < 0 || in_dtype < 0 || in_dtype > 2 || block_size <= 0) return set_error(BFP_E_ARG, "bad argument");
    if (rows * K > 0 && (!in || !vals || !sf)) return set_error(BFP_E_ARG, "null pointer");
    if (int rc = require_device()) return rc;
    return mx_pack_device(in, in_dtype, rows, K, block_size, mant_bits, eps, static_cast<uint8_t*>(vals), static_cast<uint8_t*>(sf), static_cast<cudaStream_t>(stream));
}

int bfp_gemm_mx(const void* a_vals, const void* a_sf, const void* b_vals, const void* b_sf, int b_tile_rows, int b_folded, const float* bias, float* out,
                int64_t T, int64_t N, int64_t K, void* stream) {
    if (T < 0 || N < 0 || K < 0) return set_error(BFP_E_ARG, "negative shape");
    if (T * N > 0 && (!a_vals || !a_sf || !b_vals || !b_sf || !out)) return set_error(BFP_E_ARG, "null pointer");
    if (int rc = require_device()) return rc;
    return gemm_mx_device(static_cast<const uint8_t*>(a_vals), static_cast<const uint8_t*>(a_sf), static_cast<const uint8_t*>(b_vals),
                          static_cast<const uint8_t*>(b_sf), b_tile_rows, b_folded, bias, out, T, N, round_up(K, 128), static_cast<cudaStream_t>(stream));
}

int bfp_bfloat_round(const void* in, void* out, const float* bias, int64_t n, int64_t ncols, int dtype, int bfloat, void* stream) {
    if (n < 0 || dtype < 0 || dtype > 2) return set_error(BFP_E_ARG, "bad argument");
    if (n > 0 && (!in || !out)) return set_error(BFP_E_ARG, "null pointer");
    if (bias && (ncols <= 0 || n % ncols)) return set_error(BFP_E_ARG, "bias: n must be a multiple of ncols");
    if (int rc = require_device()) return rc;
    return bfloat_round_device(in, out, bias, n, ncols, dtype, bfloat, static_cast<cudaStream_t>(stream));
}

int bfp_ocp_mx_quantize(const void* in, void* out, int64_t rows, int64_t K, int in_dtype, int out_kind, int64_t ld_out, int block_size, int elem_format,
                        int scale_bits, int bfloat, int flush_fp32_subnorms, void* stream) {
    if (rows < 0 || K < 0 || in_dtype < 0 || in_dtype > 2 || block_size < 0 || out_kind < 0 || out_kind > 1) return set_error(BFP_E_ARG, "bad argument");
    if (rows * K > 0 && (!in || !out)) return set_error(BFP_E_ARG, "null pointer");
    if (int rc = require_device()) return rc;
    return ocp_mx_quantize_device(in, out, rows, K, in_dtype, out_kind, ld_out, block_size, elem_format, scale_bits, bfloat, flush_fp32_subnorms,
                                  static_cast<cudaStream_t>(stream));
}

int bfp_ocp_mx_pack(const void* in, void* vals, void* sf, int64_t rows, int64_t K, int in_dtype, int tile_rows, int block_size, int elem_format, int scale_bits,
                    int bfloat, int flush_fp32_subnorms, void* stream) {
    if (rows < 0 || K < 0 || in_dtype < 0 || in_dtype > 2 || block_size <= 0) return set_error(BFP_E_ARG, "bad argument");
    if (rows * K > 0 && (!in || !vals || !sf)) return set_error(BFP_E_ARG, "null pointer");
    if (int rc = require_device()) return rc;
    return ocp_mx_pack_device(in, static_cast<uint8_t*>(vals), static_cast<uint8_t*>(sf), rows, K, in_dtype, tile_rows, block_size, elem_format, scale_bits,
                              bfloat, flush_fp32_subnorms, static_cast<cudaStream_t>(stream));
}

int bfp_gemm_mx_round(const void* a_vals, const void* a_sf, const void* b_vals, const void* b_sf, int b_tile_rows, int b_folded, const float* bias, float* out,
                      int64_t T, int64_t N, int64_t K, int bfloat, void* stream) {
    if (T < 0 || N < 0 || K < 0) return set_error(BFP_E_ARG, "negative shape");
    if (bfloat != 0 && (bfloat < 10 || bfloat > 32)) return set_error(BFP_E_ARG, "bfloat must be 0 or in [10, 32]");
    if (T * N > 0 && (!a_vals || !a_sf || !b_vals || !b_sf || !out)) return set_error(BFP_E_ARG, "null pointer");
    if (int rc = require_device()) return rc;
    return gemm_mx_device(static_cast<const uint8_t*>(a_vals), static_cast<const uint8_t*>(a_sf), static_cast<const uint8_t*>(b_vals),
                          static_cast<const uint8_t*>(b_sf), b_tile_rows, b_folded, bias, out, T, N, round_up(K, 128), static_cast<cudaStream_t>(stream),
                          bfloat == 32 ? 0 : bfloat);
}

int bfp_sp_layout(int64_t rows, int64_t K, int64_t* Kc, int64_t* meta_bytes) { return sp_layout(rows, round_up(K, 8), Kc, meta_bytes); }

int bfp_compress_2to4_bf16(const void* w_bf16, int64_t rows, int64_t K, void* w_comp, void* w_meta, uint32_t* violations, void* stream) {
    if (rows < 0 || K < 0) return set_error(BFP_E_ARG, "bad argument");
    if (rows * K > 0 && (!w_bf16 || !w_comp || !w_meta || !violations)) return set_error(BFP_E_ARG, "null pointer");
    if (int rc = require_device()) return rc;
    const int64_t Kp = round_up(K, 8);
    return compress_2to4_bf16_device(w_bf16, rows, Kp, Kp, w_comp, w_meta, violations, static_cast<cudaStream_t>(stream));
}

int bfp_gemm_bf16_sp(const void* x_bf16, const void* w_comp, const void* w_meta, const float* bias, float* out, int64_t T, int64_t N,
                     int64_t K, void* stream) {
    if (T < 0 || N < 0 || K <= 0) return set_error(BFP_E_ARG, "bad argument");
    if (T * N > 0 && (!x_bf16 || !w_comp || !w_meta || !out)) return set_error(BFP_E_ARG, "null pointer");
    if (int rc = require_device()) return rc;
    return gemm_bf16_sp_device(x_bf16, w_comp, w_meta, bias, out, T, N, round_up(K, 8), static_cast<cudaStream_t>(stream));
}

int bfp_gemm_bf16_sp_gather(const void* x_bf16, const void* w_comp, const void* w_meta, const float* bias, void* const* out_slices, int n_out,
                            int out_dtype, int64_t ld_out, int64_t T, int64_t N, int64_t K, void* stream) {
    if (T < 0 || N < 0 || K <= 0 || n_out < 1 || !out_slices) return set_error(BFP_E_ARG, "bad argument");
    if (T * N > 0 && (!x_bf16 || !w_comp || !w_meta)) return set_error(BFP_E_ARG, "null pointer");
    if (int rc = require_device()) return rc;
    return gemm_bf16_sp_multi_device(x_bf16, w_comp, w_meta, bias, out_slices, n_out, out_dtype, ld_out, T, N, round_up(K, 8),
                                     static_cast<cudaStream_t>(stream));
}

int bfp_gemm_bf16_sp_ex(const void* x_bf16, const void* w_comp, const void* w_meta, const float* bias, void* out, int out_dtype, int64_t ld_out,
                        int64_t T, int64_t N, int64_t K, void* stream) {
    if (T < 0 || N < 0 || K <= 0) return set_error(BFP_E_ARG, "bad argument");
    if (T * N > 0 && (!x_bf16 || !w_comp || !w_meta || !out)) return set_error(BFP_E_ARG, "null pointer");
    if (int rc = require_device()) return rc;
    void* outs[1] = {out};
    return gemm_bf16_sp_multi_device(x_bf16, w_comp, w_meta, bias, outs, 1, out_dtype, ld_out, T, N, round_up(K, 8), static_cast<cudaStream_t>(stream));
}

int bfp_unpack(const int8_t* mant, const float* scale_t, float* out, int64_t rows, int64_t K, int block_size, void* stream) {
    if (rows < 0 || K < 0 || block_size <= 0) return set_error(BFP_E_ARG, "bad argument");
    if (rows * K > 0 && (!mant || !scale_t || !out)) return set_error(BFP_E_ARG, "null pointer");
    if (int rc = require_device()) return rc;
    return unpack_device(mant, scale_t, out, rows, K, packed_kp(K), packed_rows_pad(rows), block_size, static_cast<cudaStream_t>(stream));
}

int bfp_gemm_i8(const int8_t* a_mant, const float* a_scale_t, const int8_t* b_mant, const float* b_scale_t, const float* bias,
                float* out, int64_t T, int64_t N, int64_t K, int block_size, void* stream) {
    if (T < 0 || N < 0 || K <= 0 || block_size <= 0) return set_error(BFP_E_ARG, "bad argument");
    if (T * N > 0 && (!a_mant || !a_scale_t || !b_mant || !b_scale_t || !out)) return set_error(BFP_E_ARG, "null pointer");
    if (int rc = require_device()) return rc;
    return gemm_i8_device(a_mant, a_scale_t, packed_rows_pad(T), b_mant, b_scale_t, packed_rows_pad(N), bias, out, T, N, packed_kp(K),
                          block_size, static_cast<cudaStream_t>(stream));
}

int bfp_debug_exp_table(int dtype, uint16_t out[256]) {
    if (int rc = require_device()) return rc;
    return debug_exp_table(dtype, out);
}

int bfp_debug_cpu_tie_lut(uint8_t out[256]) {
    if (int rc = require_device()) return rc;
    return debug_cpu_tie_lut(out);
}

}  // extern "C"
