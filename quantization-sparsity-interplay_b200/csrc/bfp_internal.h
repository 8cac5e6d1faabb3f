// bfp_internal.h -- host-side glue shared by the translation units of libbfp_b200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "../../include/bfp_b200.h"

namespace bfp {

struct DeviceInfo {
    int device = -1;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    size_t l2_bytes = 0;
};
const DeviceInfo& device_info();          // of the current device (cached per device)

struct Tuning {
    int stream_ctas_per_sm = 8;           // resident CTAs of the stream kernel per SM (grid = sm_count * this)
    int force_generic = 0;                // tests: route everything through the generic kernel
    int64_t host_chunk_bytes = 16 << 20;      // bfp_quantize_host: largest pipelined chunk (input bytes)
    int64_t host_chunk_min_bytes = 1 << 20;   // ... and the smallest (first / last chunks of the tapered schedule)
    int quant_tma = 0;                    // 1 = TMA-staged (cp.async.bulk -> smem ring) variant of the streaming quantiser
    int pdl = 1;                          // programmatic dependent launch for the streaming kernels (bfp_stream.cuh)
    int gemm_sp_debug = 0;                // timing experiments (wrong results): see bfp_gemm_sp.cu Params::debug
    int gemm_bf16_cta_group = 0;          // dense bf16 kind: 0 = CTA pairs when T > 128 and N > 128; 1 or 2 forces the mode
    int gemm_out_tma = 1;                 // GEMM epilogues write the output tile with TMA stores (0 = plain st.global)
    int gemm_sp_tile = 0;                 // sparse kind on CTA pairs: 0 = pick 256- or 480-token tiles by cost; 256 / 480 forces one
    int gemm_sp_cta_group = 0;            // 0 = CTA pairs (cta_group::2) when N > 128; 1 or 2 forces the mode
    int gemm_bf16_tile_n = 0;             // 0 = default tile (128x256); 128 or 256 forces the width
    int gemm_mx_variant = 0;              // bfp_gemm_mx debug knobs (layout experiments)
    int unstructured_force_fallback = 0;  // tests: the fused unstructured pipeline selects over the whole tensor (its rare path)
};
Tuning& tuning();

int set_error(int code, const char* msg);         // records the thread's last error, returns code
int set_errorf(int code, const char* fmt, ...);
int check_launch(const char* what);               // cudaGetLastError() -> BFP_E_CUDA
void count_launch();

struct QuantArgs {
    const void* in;
    void* out;
    int64_t rows, K;
    int in_dtype, out_dtype;
    int B, m;
    float eps;
    int rounding;
    uint64_t seed, offset;
    int N, M, order, tie;
    int64_t index_base = 0;   // flat element index of in[0] within the whole tensor (Philox counter base; multiple of 8)
};
int validate_quant_args(const QuantArgs& a, bool device_pointers);
int quantize_device(const QuantArgs& a, cudaStream_t st);
int block_exponent_device(const void* in, float* e_out, int64_t rows, int64_t K, int dtype, int B, float eps, cudaStream_t st);
int quantize_host(const QuantArgs& a);
int host_staging_release();
int debug_cpu_tie_lut(uint8_t out[256]);
int debug_exp_table(int dtype, uint16_t out[256]);
int pack_device(const QuantArgs& a, int8_t* mant, float* scale_t, int64_t Kp, int64_t rows_pad, cudaStream_t st);
int unpack_device(const int8_t* mant, const float* scale_t, float* out, int64_t rows, int64_t K, int64_t Kp, int64_t rows_pad, int B,
                  cudaStream_t st);
int gemm_i8_device(const int8_t* a_mant, const float* a_scale_t, int64_t lda_s, const int8_t* b_mant, const float* b_scale_t,
                   int64_t ldb_s, const float* bias, float* out, int64_t T, int64_t N, int64_t Kp, int block_size, cudaStream_t st);

int pack_bf16_device(const QuantArgs& a, void* out_bf16, int64_t Kp, cudaStream_t st);
int gemm_bf16_device(const void* a_bf16, const void* b_bf16, const float* bias, float* out, int64_t T, int64_t N, int64_t Kp,
                     cudaStream_t st);

int gemm_bf16_ex_device(const void* a_bf16, const void* b_bf16, const float* bias, void* out, int out_dtype, int64_t T, int64_t N, int64_t Kp,
                        cudaStream_t st, int accumulate = 0, int64_t batch = 1);

int sp_layout(int64_t rows, int64_t Kp, int64_t* Kc, int64_t* meta_bytes);
int compress_2to4_bf16_device(const void* w_bf16, int64_t rows, int64_t Kp, int64_t ld_w, void* comp, void* meta, unsigned int* violations,
                              cudaStream_t st);
int gemm_bf16_sp_device(const void* x_bf16, const void* w_comp, const void* w_meta, const float* bias, float* out, int64_t T, int64_t N,
                        int64_t Kp, cudaStream_t st);
int gemm_bf16_sp_multi_device(const void* x_bf16, const void* w_comp, const void* w_meta, const float* bias, void* const* out_ptrs, int n_out,
                              int out_dtype, int64_t ld_out, int64_t T, int64_t N, int64_t Kp, cudaStream_t st, int accumulate = 0);

int mx_layout(int64_t rows, int64_t K, int tile_rows, int fold, int64_t* Kp, int64_t* sf_bytes);
int mx_from_packed_device(const int8_t* mant, const float* scale_t, int64_t ld_s, int64_t rows, int64_t K, int block_size, int tile_rows, int fold,
                          uint8_t* vals, uint8_t* sf, int* row_ref, unsigned int* violations, cudaStream_t st);
int mx_pack_device(const void* in, int dtype, int64_t rows, int64_t K, int block_size, int mant_bits, float eps, uint8_t* vals, uint8_t* sf, cudaStream_t st);
int gemm_mx_device(const uint8_t* a_vals, const uint8_t* a_sf, const uint8_t* b_vals, const uint8_t* b_sf, int b_tile_rows, int b_folded, const float* bias,
                   float* out, int64_t T, int64_t N, int64_t Kp, cudaStream_t st, int out_bfloat = 0);
// MX (OCP Microscaling) formats, bfp_ocp_mx.cu
int ocp_mx_quantize_device(const void* in, void* out, int64_t rows, int64_t K, int dtype, int out_kind, int64_t ld_out, int block_size, int elem_format,
                           int scale_bits, int bfloat, int flush, cudaStream_t st);
int ocp_mx_pack_device(const void* in, uint8_t* vals, uint8_t* sf, int64_t rows, int64_t K, int dtype, int tile_rows, int block_size, int elem_format,
                       int scale_bits, int bfloat, int flush, cudaStream_t st);
int bfloat_round_device(const void* in, void* out, const float* bias, int64_t n, int64_t ncols, int dtype, int bfloat, cudaStream_t st);

size_t int_workspace_bytes(int64_t C);
int int_quantize_device(const void* in, float* out, int64_t A, int64_t C, int64_t inner, int dtype, int bits, void* workspace, cudaStream_t s);
int int_quantize_split3_device(const void* in, void* out_bf16, int64_t A, int64_t C, int64_t kseg, int dtype, int bits, void* workspace, cudaStream_t s);
int transpose16_device(const void* in, void* out, int64_t R, int64_t C, int64_t ld_in, int64_t ld_out, cudaStream_t s);
int int_quantize_nm_device(const void* in, float* out, int64_t C, int64_t K, int dtype, int bits, int N, int order, cudaStream_t s);
size_t unstructured_workspace_bytes();
int unstructured_device(const void* in, void* out, int64_t n, int dtype, unsigned long long k, void* workspace, cudaStream_t s);
struct UnstructuredArgs {           // global magnitude pruning fused with the BFP quantiser (bfp_unstructured_fused.cu)
    const void* in;
    void* out;
    void* workspace;
    int64_t n, K;                   // numel, last dim
    int dtype;
    unsigned long long k;           // entries to drop
    int order;                      // BFP_ORDER_SPARSIFY_ONLY / SPARSIFY_QUANT / QUANT_SPARSIFY
    int B, m;
    float eps;
    int rounding;
    uint64_t seed, offset;
};
size_t unstructured_fused_workspace_bytes(int64_t n, int dtype);
bool unstructured_fused_supported(const UnstructuredArgs& a);
int unstructured_fused_device(const UnstructuredArgs& a, cudaStream_t s);

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
inline int64_t packed_kp(int64_t K) { return round_up(K, 16); }
inline int64_t packed_rows_pad(int64_t rows) { return round_up(rows, 256); }
inline int64_t packed_nkb_pad(int64_t K, int B) { return round_up(packed_kp(K), 128) / 128 * (B <= 128 ? 128 / B : 1) ; }

// Persistent grid of the streaming kernels: exactly the number of CTAs that are resident at once (SMs x occupancy of
// that kernel instantiation, capped by the stream_ctas_per_sm knob), tiles dealt round-robin, so the tail is at most one
// 16 KB tile per CTA instead of a partial second wave of whole CTAs.
template <class Kernel>
inline int kernel_occupancy(Kernel kernel, int threads) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, 0) != cudaSuccess || occ < 1) occ = 4;
    return occ;
}
inline int stream_grid(int occ, int64_t n_tiles) {
    return (int)std::min<int64_t>(n_tiles, (int64_t)device_info().sm_count * std::min(occ, tuning().stream_ctas_per_sm));
}

inline size_t dtype_size(int dt) { return dt == BFP_DT_F32 ? 4 : 2; }

}  // namespace bfp
