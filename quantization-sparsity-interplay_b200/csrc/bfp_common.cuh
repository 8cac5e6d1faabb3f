// bfp_common.cuh -- device-side arithmetic of the BFP quantiser and the N:M mask, shared by every kernel.
//
// Numerical contract (DESIGN.md "Arithmetic"): bit-exact with the reference's torch-CUDA evaluation of
// src/transformers/bfp/bfp_ops.py:20-59,73-91.  torch evaluates each elementwise op of a fp16/bf16 tensor in fp32 and
// rounds the result to the tensor dtype, so the slow path below rounds through the dtype after every op; the fast path
// is taken only when every one of those roundings is provably the identity (power-of-two scale inside the dtype's
// normal range), which is the case for all blocks except all-zero / denormal-scale / overflowing ones.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bfp_b200.h"

namespace bfp {

constexpr int kMaxGroup = 64;   // largest N:M group size M supported by the generic path

// ---------------------------------------------------------------------------------------------------------------
// dtype plumbing
// ---------------------------------------------------------------------------------------------------------------
template <int DT> struct DType;
template <> struct DType<BFP_DT_F32> {
    using T = float;
    static constexpr int kVec = 4;                        // elements per 128-bit vector
    static constexpr int kMaxMant = 23;                   // fast path needs (2^m - 1) * 2^p exact in the dtype
    static constexpr int kMinExp = -100, kMaxExp = 126;   // fast-path range of the block exponent e
    static constexpr int kMinScaleExp = -126;             // smallest p = e - m with 2^p exact (normal or subnormal)
    __device__ static __forceinline__ float rnd(float x) { return x; }
    __device__ static __forceinline__ float load(const void* p, int64_t i) { return static_cast<const float*>(p)[i]; }
    __device__ static __forceinline__ void store(void* p, int64_t i, float v) { static_cast<float*>(p)[i] = v; }
};
template <> struct DType<BFP_DT_F16> {
    using T = __half;
    static constexpr int kVec = 8;
    static constexpr int kMaxMant = 11;
    static constexpr int kMinExp = -100, kMaxExp = 15;
    static constexpr int kMinScaleExp = -24;
    __device__ static __forceinline__ float rnd(float x) { return __half2float(__float2half_rn(x)); }
    __device__ static __forceinline__ float load(const void* p, int64_t i) { return __half2float(static_cast<const __half*>(p)[i]); }
    __device__ static __forceinline__ void store(void* p, int64_t i, float v) { static_cast<__half*>(p)[i] = __float2half_rn(v); }
};
template <> struct DType<BFP_DT_BF16> {
    using T = __nv_bfloat16;
    static constexpr int kVec = 8;
    static constexpr int kMaxMant = 8;
    static constexpr int kMinExp = -100, kMaxExp = 126;
    static constexpr int kMinScaleExp = -126;
    __device__ static __forceinline__ float rnd(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
    __device__ static __forceinline__ float load(const void* p, int64_t i) { return __bfloat162float(static_cast<const __nv_bfloat16*>(p)[i]); }
    __device__ static __forceinline__ void store(void* p, int64_t i, float v) { static_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v); }
};

__device__ __forceinline__ uint32_t abs_bits(float x) { return __float_as_uint(x) & 0x7fffffffu; }

// torch.maximum / torch.minimum: NaN in either operand propagates.
__device__ __forceinline__ float t_max(float a, float b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }
__device__ __forceinline__ float t_min(float a, float b) { return (a != a) ? a : ((b != b) ? b : (a < b ? a : b)); }

// ---------------------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter = (elt/4 lo, elt/4 hi, offset lo, offset hi), key = seed.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint64_t ctr, uint64_t offset, uint64_t seed) {
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = (uint32_t)offset, c3 = (uint32_t)(offset >> 32);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ float u01(uint32_t w) { return (float)(w >> 8) * 5.9604644775390625e-08f; }  // [0,1), 24 bit

// ---------------------------------------------------------------------------------------------------------------
// block scale: everything derived from the block's max |t|   (bfp_ops.py:29-33, :38-39)
// ---------------------------------------------------------------------------------------------------------------
struct BlockScale {
    float inv;     // 2^-(e-m)            fast path
    float delta;   // 2^(e-m)  (interval, bfp_ops.py:38)
    float vmax;    // 2^e - interval (max_v, bfp_ops.py:39); fast path: = qmax * delta
    float e;       // block exponent as torch holds it (float in the tensor dtype)
    bool fast;
};

template <int DT>
__device__ __forceinline__ BlockScale make_scale(uint32_t amax_bits, int m, float eps) {
    using D = DType<DT>;
    BlockScale sc;
    const float a = __uint_as_float(amax_bits);          // max |t| (NaN if any NaN: integer max on |bits|)
    const float s = D::rnd(a + eps);                      // bfp_ops.py:33  max_v + epsilon
    const uint32_t sb = __float_as_uint(s);
    const int k = (int)(sb >> 23) - 127;                  // floor(log2 s) for normal s >= 0; 128 for inf/NaN
    bool ok = (k >= D::kMinExp) && (k < D::kMaxExp) && (m >= 1) && (m <= D::kMaxMant);
    int e = 0;
    if (ok) {
        if (DT == BFP_DT_F32 && (sb & 0x7fffffu) > 128u) {
            // log2f(s) lies at least 128*1.44*2^-23 above k: no 1-ulp log2f can round it down to k
            // (fp32 spacing near |k| <= 127 is at most 2^-17), so ceil(log2f(s)) = k + 1 without evaluating it.
            e = k + 1;
        } else {
            e = (int)ceilf(D::rnd(log2f(s)));             // bfp_ops.py:33  .log2().ceil()
        }
        ok = (e - m >= D::kMinScaleExp) && (e <= D::kMaxExp) && (e >= D::kMinExp);
    }
    sc.fast = ok;
    if (ok) {
        const int p = e - m;                              // in [-126, 125]: both 2^p and 2^-p are normal floats
        sc.delta = __uint_as_float((uint32_t)(p + 127) << 23);
        sc.inv = __uint_as_float((uint32_t)(127 - p) << 23);
        sc.vmax = (float)((1 << m) - 1);                  // fast path clamps on the integer grid: |q| <= 2^m - 1
        sc.e = (float)e;
    } else {
        // literal evaluation, rounding through the dtype after every op (all-zero fp16 block -> NaN, inf -> NaN, ...)
        const float ef = ceilf(D::rnd(log2f(s)));
        const float p = D::rnd(ef - (float)m);
        sc.delta = D::rnd(powf(2.0f, p));
        sc.vmax = D::rnd(D::rnd(powf(2.0f, ef)) - sc.delta);
        sc.inv = 0.0f;
        sc.e = ef;
    }
    return sc;
}

// one element: bfp_ops.py:40-44.  STOC: u is the element's uniform in [0,1).
template <int DT, bool STOC>
__device__ __forceinline__ float quant_elt(float t, const BlockScale& sc, float u) {
    using D = DType<DT>;
    if (sc.fast) {
        const float x = t * sc.inv;                                   // exact (power-of-two scale)
        const float r = STOC ? rintf((u - 0.5f) + x) : rintf(x);      // bfp_ops.py:22-25
        return fminf(fmaxf(r, -sc.vmax), sc.vmax) * sc.delta;         // clamp on the grid, exact product
    }
    const float x = D::rnd(t / sc.delta);
    float y;
    if (STOC) y = rintf((u - 0.5f) + x) * sc.delta;                   // fp32 from here on (type promotion)
    else y = D::rnd(rintf(x) * sc.delta);
    return t_min(t_max(y, -sc.vmax), sc.vmax);
}

// ---------------------------------------------------------------------------------------------------------------
// N:M mask inside one group held in registers (bfp_ops.py:84-87).  Drops (-> +0.0) the KDROP = M-N entries that
// torch.topk(|t|, k, largest=False) returns.  TORCH_CUDA rule: smallest by (|v|, index); NaN is largest.
// ---------------------------------------------------------------------------------------------------------------
static __constant__ uint8_t c_cpu_tie_lut[256] = {
#include "nm_cpu_tie_lut.inc"
};

template <int M, int TIE>
__device__ __forceinline__ void nm_mask_group(float* v, int kdrop) {
    uint32_t key[M];
#pragma unroll
    for (int i = 0; i < M; ++i) key[i] = abs_bits(v[i]);
    if (M == 4 && TIE == BFP_TIE_TORCH_CPU) {
        // kdrop == 2 guaranteed by the host.  c_i = #{j : key_j < key_i}
        int idx = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int c = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) c += (j != i) && (key[j] < key[i]);
            idx += c << (2 * i);
        }
        const uint32_t mask = c_cpu_tie_lut[idx];
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (mask & (1u << i)) v[i] = 0.0f;
        return;
    }
#pragma unroll
    for (int i = 0; i < M; ++i) {
        int rank = 0;
#pragma unroll
        for (int j = 0; j < M; ++j) {
            if (j < i) rank += (key[j] <= key[i]);
            if (j > i) rank += (key[j] < key[i]);
        }
        if (rank < kdrop) v[i] = 0.0f;
    }
}

}  // namespace bfp
