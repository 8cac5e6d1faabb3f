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
// Order key of torch.topk over |x| on CUDA (bfp_ops.py:66): the bit pattern of |x|, with every NaN mapped to ONE key above
// +inf -- torch's radix select converts all NaNs to the same all-ones pattern, so NaNs tie with each other (index order).
__device__ __forceinline__ uint32_t topk_key(float x) { return min(abs_bits(x), 0x7f800001u); }

// torch.maximum / torch.minimum: NaN in either operand propagates.
__device__ __forceinline__ float t_max(float a, float b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }
__device__ __forceinline__ float t_min(float a, float b) { return (a != a) ? a : ((b != b) ? b : (a < b ? a : b)); }

// bfloatX rounding of one fp32 value (mx/elemwise_ops.py _quantize_bfloat, round 'nearest', subnormals kept, overflow -> Inf): on the
// bit pattern of |x| the library's floor(|x| 2^(bits-2-pe) + 0.5) is "add half of the dropped field, clear it" -- the carry into the
// exponent is the rounding up to the next binade, 0x7f800000 is the overflow to Inf, and the subnormal range is linear in the bits.
__device__ __forceinline__ float round_bfloat(float x, int bfloat) {
    if (bfloat <= 0 || bfloat >= 32) return x;
    const uint32_t b = __float_as_uint(x), a = b & 0x7fffffffu;
    if (a >= 0x7f800000u) return x;                                   // Inf / NaN pass through
    const int drop = 32 - bfloat;
    const uint32_t r = (a + (1u << (drop - 1))) & ~((1u << drop) - 1u);
    if (r == 0u) return a == 0u ? 0.0f : __uint_as_float(b & 0x80000000u);   // sign(0) = 0 -> +0.0; a rounded-away negative keeps -0.0
    return __uint_as_float(r | (b & 0x80000000u));
}

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
// uniform in [0,1): the 32-bit word truncated to 24 significant bits (round toward zero), times 2^-32 -- one conversion, no shift.
// Words >= 2^31 give exactly (w >> 8) * 2^-24; smaller words keep proportionally finer steps.  Never 1.0.
__device__ __forceinline__ float u32_rz(uint32_t w) { return __uint2float_rz(w); }
__device__ __forceinline__ float u01(uint32_t w) { return u32_rz(w) * 2.3283064365386963e-10f; }
// fl(u - 0.5): the sample the reference adds to t / interval (bfp_ops.py:22), in one fused operation (u itself is exact)
__device__ __forceinline__ float u01_centered(uint32_t w) { return __fmaf_rn(u32_rz(w), 2.3283064365386963e-10f, -0.5f); }

// ---------------------------------------------------------------------------------------------------------------
// block scale: everything derived from the block's max |t|   (bfp_ops.py:29-33, :38-39)
// ---------------------------------------------------------------------------------------------------------------
struct BlockScale {
    float inv;     // 2^-(e-m)                       fast path only
    float delta;   // 2^(e-m)  (interval, bfp_ops.py:38)
    float vmax;    // fast path: 2^m - 1 (clamp on the integer grid); slow path: 2^e - interval (max_v, bfp_ops.py:39)
    float e;       // block exponent as torch holds it (float in the tensor dtype)
    int p;         // e - m (fast path only)
    bool fast;
};

// Literal evaluation of bfp_ops.py:33,38-39, rounding through the dtype after every op.  Rare blocks only (all-zero,
// denormal scale, overflow, Inf/NaN, mant_bits outside the exact range), so it is kept out of line.
// returns (delta, vmax, e)
template <int DT>
__device__ __noinline__ float3 make_scale_slow(float s, int m) {
    using D = DType<DT>;
    const float ef = ceilf(D::rnd(log2f(s)));
    const float p = D::rnd(ef - (float)m);
    const float delta = D::rnd(powf(2.0f, p));
    const float vmax = D::rnd(D::rnd(powf(2.0f, ef)) - delta);
    return make_float3(delta, vmax, ef);
}

// ceil(log2f(s)) for the few fp32 values whose mantissa is within 128 ulp of a power of two (out of line: log2f is
// ~20 instructions and this branch is taken for ~1.5e-5 of the blocks).
static __device__ __noinline__ int exponent_near_pow2(float s) { return (int)ceilf(log2f(s)); }

// ---------------------------------------------------------------------------------------------------------------
// Half-precision block exponent without log2f.  For fp16 / bf16 tensors torch evaluates ceil(log2(s)) with log2(s) ROUNDED TO
// THE DTYPE first, so e is k or k+1 depending on where the mantissa f of s = 2^k (1 + f 2^-mb)
// sits relative to a threshold that depends on k (the spacing of the dtype at magnitude |k|).  The thresholds are not
// derived: a one-time kernel evaluates the literal formula (the same libdevice log2f as before) for every (k, f), checks
// that it is a step function of f, and stores the step position; the hot path does one table load and one compare.
// Entry 0 = "not tabulated" (before initialisation, or a k whose step check failed): evaluate the formula as before.
// ---------------------------------------------------------------------------------------------------------------
static __device__ uint16_t g_exp_step[2][256];           // [0] fp16, [1] bf16; index k + 128; value = step position + 1

template <int DT> struct HalfBits;
template <> struct HalfBits<BFP_DT_F16> { static constexpr int kMant = 10, kTable = 0; };
template <> struct HalfBits<BFP_DT_BF16> { static constexpr int kMant = 7, kTable = 1; };

template <int DT>
static __global__ void exp_table_init_kernel() {
    using D = DType<DT>;
    constexpr int MB = HalfBits<DT>::kMant;
    const int k = (int)threadIdx.x - 128;
    uint32_t value = 0;
    if (k >= -126 && k <= 127) {
        int step = 1 << MB;                               // first f with e = k + 1 (1 << MB: none)
        bool valid = true;
        for (int f = 0; f < (1 << MB); ++f) {
            const float s = __uint_as_float(((uint32_t)(k + 127) << 23) | ((uint32_t)f << (23 - MB)));
            const int e = (int)ceilf(D::rnd(log2f(s)));
            if (e == k + 1) { if (step == (1 << MB)) step = f; }
            else if (e == k) { if (step != (1 << MB)) valid = false; }      // back down after the step: not monotone
            else valid = false;
        }
        if (valid) value = (uint32_t)step + 1u;
    }
    g_exp_step[HalfBits<DT>::kTable][threadIdx.x] = (uint16_t)value;
}

// Per translation unit and device, once: fills this TU's copy of the table.  Skipped (formula path stays in use) while the
// caller's stream is being captured into a CUDA graph.
static inline void ensure_exp_tables(cudaStream_t user_stream) {
    static bool done[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || done[dev]) return;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(user_stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { cudaGetLastError(); return; }
    cudaStream_t st;
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return; }
    exp_table_init_kernel<BFP_DT_F16><<<1, 256, 0, st>>>();
    exp_table_init_kernel<BFP_DT_BF16><<<1, 256, 0, st>>>();
    if (cudaStreamSynchronize(st) == cudaSuccess) done[dev] = true; else cudaGetLastError();
    cudaStreamDestroy(st);
}

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
        if (DT == BFP_DT_F32) {
            // mantissa > 128 ulp above 2^k: log2f(s) >= k + 128*1.44*2^-23, which no 1-ulp log2f can round down to k
            // (fp32 spacing near |k| <= 127 is at most 2^-17), so ceil(log2f(s)) = k + 1 without evaluating it.
            e = ((sb & 0x7fffffu) > 128u) ? k + 1 : exponent_near_pow2(s);
        } else {
            constexpr int MB = HalfBits<DT == BFP_DT_F32 ? BFP_DT_F16 : DT>::kMant;
            const uint32_t step1 = g_exp_step[HalfBits<DT == BFP_DT_F32 ? BFP_DT_F16 : DT>::kTable][k + 128];
            const uint32_t f = (sb >> (23 - MB)) & ((1u << MB) - 1u);
            if (step1 != 0u) e = k + (int)(f + 1u >= step1);          // tabulated step of ceil(rnd(log2 s)) in f
            else e = (int)ceilf(D::rnd(log2f(s)));        // bfp_ops.py:33  .log2().ceil(), log2 rounded to the dtype
        }
        ok = (e - m >= D::kMinScaleExp) && (e <= D::kMaxExp) && (e >= D::kMinExp);
    }
    if (ok) {
        const int p = e - m;                              // in [-126, 125]: both 2^p and 2^-p are normal floats
        sc.delta = __uint_as_float((uint32_t)(p + 127) << 23);
        sc.inv = __uint_as_float((uint32_t)(127 - p) << 23);
        sc.vmax = (float)((1 << m) - 1);
        sc.e = (float)e;
        sc.p = p;
        sc.fast = true;
    } else {
        const float3 r = make_scale_slow<DT>(s, m);
        sc.delta = r.x; sc.vmax = r.y; sc.e = r.z; sc.inv = 0.0f; sc.p = 0; sc.fast = false;
    }
    return sc;
}

// one element, fast path: bfp_ops.py:40-44 with every product exact.  STOC: smp = fl(u - 0.5), u the element's uniform in [0,1).
template <bool STOC>
__device__ __forceinline__ float quant_elt_fast(float t, const BlockScale& sc, float smp) {
    const float x = t * sc.inv;                                   // exact (power-of-two scale)
    const float r = STOC ? rintf(smp + x) : rintf(x);             // bfp_ops.py:22-25
    return fminf(fmaxf(r, -sc.vmax), sc.vmax) * sc.delta;         // clamp on the grid, exact product
}

// one element, literal evaluation (out of line; see make_scale_slow)
template <int DT, bool STOC>
__device__ __noinline__ float quant_elt_slow(float t, float delta, float vmax, float smp) {
    using D = DType<DT>;
    const float x = D::rnd(t / delta);
    float y;
    if (STOC) y = rintf(smp + x) * delta;                         // fp32 from here on (type promotion)
    else y = D::rnd(rintf(x) * delta);
    return t_min(t_max(y, -vmax), vmax);
}

template <int DT, bool STOC>
__device__ __forceinline__ float quant_elt(float t, const BlockScale& sc, float smp) {
    return sc.fast ? quant_elt_fast<STOC>(t, sc, smp) : quant_elt_slow<DT, STOC>(t, sc.delta, sc.vmax, smp);
}

// ---------------------------------------------------------------------------------------------------------------
// N:M mask inside one group held in registers (bfp_ops.py:84-87).  Drops (-> +0.0) the KDROP = M-N entries that
// torch.topk(|t|, k, largest=False) returns.  TORCH_CUDA rule: smallest by (|v|, index); NaN is largest.
// ---------------------------------------------------------------------------------------------------------------
static __constant__ uint8_t c_cpu_tie_lut[256] = {
#include "nm_cpu_tie_lut.inc"
};

// rank < K for a rank that is the number of true predicates among (a, b, c)
template <int K> __device__ __forceinline__ bool fewer_than(bool a, bool b, bool c) {
    if (K == 1) return !(a | b | c);
    if (K == 2) return !((a & b) | (a & c) | (b & c));
    return !(a & b & c);
}

// M = 4 with compile-time KDROP, torch-CUDA rule: element i is dropped iff fewer than KDROP elements precede it in
// (|v|, index) order.  6 compares + 4 predicate LUTs + 4 selects.
template <int KDROP>
__device__ __forceinline__ void nm_mask4(float* v) {
    const uint32_t k0 = abs_bits(v[0]), k1 = abs_bits(v[1]), k2 = abs_bits(v[2]), k3 = abs_bits(v[3]);
    if (KDROP == 2) {
        // c_ij = (k_i <= k_j), i < j.  j > i precedes i iff !c_ij; j < i precedes i iff c_ji.  With K = 2,
        // "fewer than 2 of (a,b,c)" = !maj(a,b,c), and !maj(!a,!b,!c) = maj(a,b,c):
        //   drop0 = maj(c01, c02, c03)   drop1 = maj(!c01, c12, c13)   drop2 = maj(!c02, !c12, c23)   drop3 = maj(!c03, !c13, !c23)
        // written on predicate registers so ptxas emits one PLOP3 per majority.
        asm("{\n\t"
            ".reg .pred c01, c02, c03, c12, c13, c23, n01, n02, n03, n12, n13, n23, t0, t1, t2, d;\n\t"
            "setp.le.u32 c01, %4, %5;\n\t setp.le.u32 c02, %4, %6;\n\t setp.le.u32 c03, %4, %7;\n\t"
            "setp.le.u32 c12, %5, %6;\n\t setp.le.u32 c13, %5, %7;\n\t setp.le.u32 c23, %6, %7;\n\t"
            "not.pred n01, c01;\n\t not.pred n02, c02;\n\t not.pred n03, c03;\n\t"
            "not.pred n12, c12;\n\t not.pred n13, c13;\n\t not.pred n23, c23;\n\t"
            "and.pred t0, c01, c02;\n\t and.pred t1, c01, c03;\n\t and.pred t2, c02, c03;\n\t"
            "or.pred d, t0, t1;\n\t or.pred d, d, t2;\n\t selp.f32 %0, 0f00000000, %0, d;\n\t"
            "and.pred t0, n01, c12;\n\t and.pred t1, n01, c13;\n\t and.pred t2, c12, c13;\n\t"
            "or.pred d, t0, t1;\n\t or.pred d, d, t2;\n\t selp.f32 %1, 0f00000000, %1, d;\n\t"
            "and.pred t0, n02, n12;\n\t and.pred t1, n02, c23;\n\t and.pred t2, n12, c23;\n\t"
            "or.pred d, t0, t1;\n\t or.pred d, d, t2;\n\t selp.f32 %2, 0f00000000, %2, d;\n\t"
            "and.pred t0, n03, n13;\n\t and.pred t1, n03, n23;\n\t and.pred t2, n13, n23;\n\t"
            "or.pred d, t0, t1;\n\t or.pred d, d, t2;\n\t selp.f32 %3, 0f00000000, %3, d;\n\t"
            "}"
            : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3])
            : "r"(k0), "r"(k1), "r"(k2), "r"(k3));
        return;
    }
    const bool c01 = k0 <= k1, c02 = k0 <= k2, c03 = k0 <= k3, c12 = k1 <= k2, c13 = k1 <= k3, c23 = k2 <= k3;
    const bool d0 = fewer_than<KDROP>(!c01, !c02, !c03);   // j > 0 precedes 0 iff key_j <  key_0
    const bool d1 = fewer_than<KDROP>(c01, !c12, !c13);    // 0 precedes 1     iff key_0 <= key_1
    const bool d2 = fewer_than<KDROP>(c02, c12, !c23);
    const bool d3 = fewer_than<KDROP>(c03, c13, c23);
    v[0] = d0 ? 0.0f : v[0]; v[1] = d1 ? 0.0f : v[1]; v[2] = d2 ? 0.0f : v[2]; v[3] = d3 ? 0.0f : v[3];
}

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


// ---------------------------------------------------------------------------------------------------------------
// 2:4-style masks (M = 4, torch-CUDA rule) on sign bits instead of predicates.
// For i < j let s_ij = sign(k_j - k_i) = (k_j < k_i) = "j precedes i" (j > i needs a strictly smaller key); "i precedes j" is
// its negation.  Element i is dropped iff fewer than KD of the other three precede it: one LOP3 per element on the three
// difference words (only bit 31 matters), with the negations folded into the look-up table.
// ---------------------------------------------------------------------------------------------------------------
__host__ __device__ constexpr uint32_t drop_lut(int kdrop, int na, int nb, int nc) {
    uint32_t lut = 0;
    for (int i = 0; i < 8; ++i) {
        const int a = (i >> 2) & 1, b = (i >> 1) & 1, c = i & 1;
        if ((a ^ na) + (b ^ nb) + (c ^ nc) < kdrop) lut |= 1u << i;
    }
    return lut;
}
template <uint32_t LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return d;
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// four fp32 values (any bit patterns): dropped entries become +0.0, kept entries keep their bits
template <int KD>
__device__ __forceinline__ void nm_mask4_bits(uint32_t* w) {
    const uint32_t k0 = w[0] & 0x7fffffffu, k1 = w[1] & 0x7fffffffu, k2 = w[2] & 0x7fffffffu, k3 = w[3] & 0x7fffffffu;
    const uint32_t s01 = k1 - k0, s02 = k2 - k0, s03 = k3 - k0, s12 = k2 - k1, s13 = k3 - k1, s23 = k3 - k2;   // keys < 2^31: no wrap
    const uint32_t d0 = lop3<drop_lut(KD, 0, 0, 0)>(s01, s02, s03);
    const uint32_t d1 = lop3<drop_lut(KD, 1, 0, 0)>(s01, s12, s13);
    const uint32_t d2 = lop3<drop_lut(KD, 1, 1, 0)>(s02, s12, s23);
    const uint32_t d3 = lop3<drop_lut(KD, 1, 1, 1)>(s03, s13, s23);
    w[0] &= ~(uint32_t)((int32_t)d0 >> 31); w[1] &= ~(uint32_t)((int32_t)d1 >> 31);
    w[2] &= ~(uint32_t)((int32_t)d2 >> 31); w[3] &= ~(uint32_t)((int32_t)d3 >> 31);
}

// four 16-bit floats (fp16 or bf16 bit patterns) packed as w0 = (e0 | e1 << 16), w1 = (e2 | e3 << 16).  Per 16-bit lane,
// bit 15 of (k_j | 0x8000) - k_i is (k_j >= k_i); lanes never borrow from each other because the minuend has bit 15 set.
template <int KD>
__device__ __forceinline__ void nm_mask4_packed16(uint32_t& w0, uint32_t& w1) {
    constexpr uint32_t H = 0x80008000u;
    const uint32_t a0 = w0 & ~H, a1 = w1 & ~H;                    // keys
    const uint32_t x0 = prmt(w0 | H, 0u, 0x1032u), x1 = prmt(w1 | H, 0u, 0x1032u);   // halves swapped, bit 15 forced
    const uint32_t X = (w1 | H) - a0;                             // lo: k2 >= k0      hi: k3 >= k1
    const uint32_t Y = x1 - a0;                                   // lo: k3 >= k0      hi: k2 >= k1
    const uint32_t R0 = x0 - a0 - 0x10000u;                       // lo: k1 >= k0      hi: k0 >  k1
    const uint32_t R1 = x1 - a1 - 0x10000u;                       // lo: k3 >= k2      hi: k2 >  k3
    const uint32_t Ys = prmt(Y, 0u, 0x1032u);                     // lo: k2 >= k1      hi: k3 >= k0
    // element 0 (lo) is preceded by j iff k_j < k_0: !R0, !X, !Y.  element 1 (hi): by 0 iff k0 <= k1 = !R0.hi, by 3 iff !X.hi, by 2 iff !Y.hi
    const uint32_t d01 = lop3<drop_lut(KD, 1, 1, 1)>(R0, X, Y);
    // element 2 (lo): by 0 iff k0 <= k2 = X.lo, by 1 iff k1 <= k2 = Ys.lo, by 3 iff k3 < k2 = !R1.lo.  element 3 (hi): X.hi, Ys.hi, !R1.hi
    const uint32_t d23 = lop3<drop_lut(KD, 0, 0, 1)>(X, Ys, R1);
    w0 &= ~prmt(d01, 0u, 0xBB99u);                                // bit 15 / bit 31 replicated over their 16-bit lane
    w1 &= ~prmt(d23, 0u, 0xBB99u);
}

}  // namespace bfp
