// bfp_stream.cuh -- helpers shared by the streaming kernels (fake-quant in bfp_quant.cu, packed in bfp_pack.cu).
#pragma once
#include "bfp_common.cuh"
#include "bfp_internal.h"

namespace bfp {

constexpr int kStreamThreads = 256;
constexpr int kStreamUnroll = 4;

// ---------------------------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  A kernel launched with launch_pdl may have its CTAs scheduled while the previous
// kernel on the stream is still draining; pdl_wait() blocks until that kernel has completed and its writes are visible, so
// stream-order semantics are unchanged -- only launch latency and the CTA prologue overlap the predecessor's tail.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <class Kernel, class Params>
inline int launch_pdl(Kernel kernel, int grid, int threads, cudaStream_t st, const Params& p, size_t smem_bytes = 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)threads); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = tuning().pdl ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, p);
    if (e != cudaSuccess) return set_errorf(BFP_E_CUDA, "cudaLaunchKernelEx: %s", cudaGetErrorString(e));
    return BFP_OK;
}

// Block maximum of unsigned keys over the 2^j adjacent lanes that share a BFP / MX block.  Default: xor butterfly (j shuffle + max
// pairs).  -DBFP_REDUX_MAX: one redux.sync.max.u32 per value with the block's lanes as member mask (every lane of the warp executes
// it; the groups of a warp reduce independently).
__device__ __forceinline__ uint32_t block_lane_mask(int lanes_per_block) {
    const uint32_t lane = threadIdx.x & 31u;
    return lanes_per_block >= 32 ? 0xffffffffu : (((1u << lanes_per_block) - 1u) << (lane & ~(uint32_t)(lanes_per_block - 1)));
}
__device__ __forceinline__ uint32_t block_max_u32(uint32_t v, int lanes_per_block, uint32_t mask) {
#ifdef BFP_REDUX_MAX
    (void)lanes_per_block;
    return __reduce_max_sync(mask, v);
#else
    (void)mask;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1)
        if (off < lanes_per_block) v = max(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
#endif
}

// n / d for a launch-invariant divisor (Granlund & Montgomery 1994, figure 4.1): q = (t + ((n - t) >> sh1)) >> sh2, t = mulhi(m, n)
struct FastDiv { uint32_t m, sh1, sh2; };
inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    uint32_t l = 0;
    while ((1ull << l) < d) ++l;                                        // ceil(log2 d)
    f.m = (uint32_t)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
    f.sh1 = l < 1 ? l : 1; f.sh2 = l > 0 ? l - 1 : 0;
    return f;
}
__device__ __forceinline__ uint32_t fastdiv(uint32_t n, const FastDiv& f) {
    const uint32_t t = __umulhi(f.m, n);
    return (t + ((n - t) >> f.sh1)) >> f.sh2;
}


// ---------------------------------------------------------------------------------------------------------------
// 128-bit streaming loads / stores
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream(uint4* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int DT> __device__ __forceinline__ void unpack_vec(const uint4& raw, float* v);
template <> __device__ __forceinline__ void unpack_vec<BFP_DT_F32>(const uint4& raw, float* v) {
    v[0] = __uint_as_float(raw.x); v[1] = __uint_as_float(raw.y); v[2] = __uint_as_float(raw.z); v[3] = __uint_as_float(raw.w);
}
template <> __device__ __forceinline__ void unpack_vec<BFP_DT_BF16>(const uint4& raw, float* v) {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
template <> __device__ __forceinline__ void unpack_vec<BFP_DT_F16>(const uint4& raw, float* v) {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
}
template <int DT> __device__ __forceinline__ uint4 pack_vec(const float* v);   // 8 (half) or 4 (fp32) values -> 16 B
template <> __device__ __forceinline__ uint4 pack_vec<BFP_DT_F32>(const float* v) {
    return make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
}
template <> __device__ __forceinline__ uint4 pack_vec<BFP_DT_BF16>(const float* v) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]); w[i] = *reinterpret_cast<uint32_t*>(&h); }
    return make_uint4(w[0], w[1], w[2], w[3]);
}
template <> __device__ __forceinline__ uint4 pack_vec<BFP_DT_F16>(const float* v) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]); w[i] = *reinterpret_cast<uint32_t*>(&h); }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// N:M mask of every group in a vector.  M: group size held in-lane (0 = none).  KD: compile-time M-N for M == 4
// (0 = use the runtime kdrop).
template <int M, int KD, int TIE, int V>
__device__ __forceinline__ void mask_vec(float* v, int kdrop) {
    if (M == 0) return;
    constexpr int MM = M > 0 ? M : 1;
#pragma unroll
    for (int g = 0; g < V / MM; ++g) {
        if (M == 4 && KD > 0 && TIE == BFP_TIE_TORCH_CUDA) nm_mask4<(KD > 0 ? KD : 1)>(v + g * MM);
        else nm_mask_group<MM, TIE>(v + g * MM, kdrop);
    }
}

template <int DT>
__device__ __forceinline__ float ld_pad(const void* in, int64_t row, int64_t col, int64_t K) {
    return (col < K) ? DType<DT>::load(in, row * K + col) : 0.0f;     // zero padding of F.pad (bfp_ops.py:52, :81)
}

// drop flag of element `col` under the N:M mask of `src(col)` (torch-CUDA rule, or CPU table for 2:4)
template <class Src>
__device__ __forceinline__ bool nm_dropped(const Src& src, int64_t col, int N, int M, int tie) {
    const int64_t g0 = (col / M) * M;
    const int i = (int)(col - g0);
    const int kdrop = M - N;
    if (tie == BFP_TIE_TORCH_CPU && M == 4 && kdrop == 2) {
        uint32_t key[4];
        for (int j = 0; j < 4; ++j) key[j] = abs_bits(src(g0 + j));
        int idx = 0;
        for (int a = 0; a < 4; ++a) { int c = 0; for (int b = 0; b < 4; ++b) c += (b != a) && (key[b] < key[a]); idx += c << (2 * a); }
        return (c_cpu_tie_lut[idx] >> i) & 1u;
    }
    const uint32_t ki = abs_bits(src(col));
    int rank = 0;
    for (int j = 0; j < M; ++j) {
        const uint32_t kj = abs_bits(src(g0 + j));
        rank += (j < i) ? (kj <= ki) : ((j > i) ? (kj < ki) : 0);
    }
    return rank < kdrop;
}

}  // namespace bfp
