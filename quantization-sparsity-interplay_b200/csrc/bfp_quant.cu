// bfp_quant.cu -- the fused BFP quantise + N:M sparsify kernels (fake-quant output) for sm_100a.
//
// Replaces, in one launch and one pass over HBM (read once, write once), what the reference does with ~25 eager torch
// kernels: float_to_bfp_blocked (bfp_ops.py:124-149) = _structured_N_M_sparsity (:73-91) and
// _no_sparsity_float_to_bfp (:46-59) in either order.
//
//  * stream kernel  -- the hot path.  Applies when K % block_size == 0, block_size is a power-of-two multiple of the
//    128-bit vector (4 fp32 / 8 half elements), and the N:M group fits in one vector.  Then rows are irrelevant: the
//    tensor is a flat sequence of vectors; lane l of a warp owns vector 32*w + l, a block is 2^j adjacent lanes,
//    block max = butterfly of __shfl_xor, the N:M mask is lane-local.  HBM-bound: 8 B/element fp32, 4 B/element half.
//  * generic kernel -- everything else (ragged K, odd block sizes, groups straddling vectors).  One thread per block
//    (or per group), gather-style, no temporaries.  Correctness path for the ViT conv shapes; not tuned.
#include <algorithm>

#include "bfp_stream.cuh"
#include "bfp_internal.h"

namespace bfp {

// ---------------------------------------------------------------------------------------------------------------
// stream kernel
// ---------------------------------------------------------------------------------------------------------------
struct StreamParams {
    const uint4* in;
    uint4* out;
    int64_t n_vec;          // number of 128-bit input vectors
    int lanes_per_block;    // block_size / kVec, power of two in [1, 32]
    int m;
    float eps;
    int kdrop;              // M - N
    uint64_t seed, offset;
    int64_t ctr_base;       // Philox counter of vector 0 (= flat element index / 4)
    // Padded-row mode (K a multiple of the vector width but not of the block size): every row is seen as slots_per_row
    // vector slots (a whole number of blocks), of which the first vec_per_row exist; the rest read as zeros -- exactly the
    // F.pad of bfp_ops.py:52 -- and are never stored.  n_vec then counts SLOTS.  0 = flat mode (rows do not matter).
    uint32_t vec_per_row, slots_per_row;
};

// slot index g -> index of the real vector (or -1 for a padding slot / beyond the tensor)
__device__ __forceinline__ int64_t real_vec(const StreamParams& p, int64_t g) {
    if (p.slots_per_row == 0u) return g < p.n_vec ? g : -1;
    if (g >= p.n_vec) return -1;
    const uint32_t row = (uint32_t)((uint64_t)g / p.slots_per_row), slot = (uint32_t)((uint64_t)g - (uint64_t)row * p.slots_per_row);
    return slot < p.vec_per_row ? (int64_t)row * p.vec_per_row + slot : -1;
}


// One 128-bit vector (4 fp32 / 8 half elements of one block's lane) through mask -> block max (butterfly over the lanes that
// share the block) -> scale -> round -> mask, in the order ORDER.  vec_index = flat index of the vector (Philox counter).
// Every lane of the warp must call this (the butterfly shuffles are warp-wide).
template <int DT, int ORDER, int M, int KD, int TIE, bool STOC>
__device__ __forceinline__ void process_vec(const uint4& raw, const StreamParams& p, int64_t vec_index, uint4* out) {
    using D = DType<DT>;
    constexpr int V = D::kVec;
    constexpr bool kQuant = ORDER != BFP_ORDER_SPARSIFY_ONLY;
    constexpr bool kSparseFirst = ORDER == BFP_ORDER_SPARSIFY_QUANT || ORDER == BFP_ORDER_SPARSIFY_ONLY;
    constexpr bool kSparseLast = ORDER == BFP_ORDER_QUANT_SPARSIFY;
    constexpr int kOutVecs = (STOC && V == 8) ? 2 : 1;     // fp32 output of 8 half inputs = two 16-B stores
    float v[V];
    unpack_vec<DT>(raw, v);
    uint32_t amax = 0u;
    if (kQuant && kSparseFirst) {
        // the block max always survives an N:M mask with N >= 1, so max over the unmasked keys is the
        // masked block's max (SURVEY.md appendix A.4 i)
#pragma unroll
        for (int i = 0; i < V; ++i) amax = max(amax, abs_bits(v[i]));
    }
    if (kSparseFirst) mask_vec<M, KD, TIE, V>(v, p.kdrop);
    if (kQuant) {
        if (!kSparseFirst) {
#pragma unroll
            for (int i = 0; i < V; ++i) amax = max(amax, abs_bits(v[i]));
        }
        // maximum over the lanes that share this block; every lane of the warp takes part
        amax = block_max_u32(amax, p.lanes_per_block, block_lane_mask(p.lanes_per_block));
        const BlockScale sc = make_scale<DT>(amax, p.m, p.eps);
        float un[STOC ? V : 1];
        if (STOC) {
#pragma unroll
            for (int q = 0; q < V / 4; ++q) {
                const uint4 r = philox4x32_10((uint64_t)(p.ctr_base + vec_index * (V / 4) + q), p.offset, p.seed);
                un[4 * q] = u01_centered(r.x); un[4 * q + 1] = u01_centered(r.y); un[4 * q + 2] = u01_centered(r.z); un[4 * q + 3] = u01_centered(r.w);
            }
        }
        if (sc.fast) {                                 // one branch per vector, uniform across the block's lanes
#pragma unroll
            for (int i = 0; i < V; ++i) v[i] = quant_elt_fast<STOC>(v[i], sc, STOC ? un[i] : 0.0f);
        } else {
#pragma unroll
            for (int i = 0; i < V; ++i) v[i] = quant_elt_slow<DT, STOC>(v[i], sc.delta, sc.vmax, STOC ? un[i] : 0.0f);
        }
    }
    if (kSparseLast) mask_vec<M, KD, TIE, V>(v, p.kdrop);
    if (kOutVecs == 1) {
        out[0] = STOC ? pack_vec<BFP_DT_F32>(v) : pack_vec<DT>(v);
    } else {
        out[0] = pack_vec<BFP_DT_F32>(v);
        out[kOutVecs - 1] = pack_vec<BFP_DT_F32>(v + 4);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Tile path (the common configurations: no mask or an N:4 mask with the torch-CUDA tie rule).  A thread's kStreamUnroll
// vectors go through the phases TOGETHER -- mask / block max of all, one butterfly (the shuffle latencies of the vectors
// overlap and the lanes_per_block test is taken once per step, not once per vector), scales, Philox for all, rounding --
// and the arithmetic is laid out for the pipes the kernel is short of: it is bound by the ALU pipe (logic, compares,
// min/max: 16 lanes/clk/SMSP), so work is moved to the FMA pipe wherever that is exact.
//   * nearest rounding never divides or multiplies: with C = 2^(p + MB) (p = e - m, MB = explicit mantissa bits of the
//     arithmetic type) fl(|t| + C) - C is |t| rounded half-to-even to a multiple of 2^p, exactly, because |t| <= 2^e <= C
//     keeps the sum inside [C, 2C] where the spacing is 2^p; then min(., (2^m - 1) 2^p) and the sign of t.  For fp16 / bf16
//     tensors this runs on packed pairs (HADD2 / HMNMX2), no unpacking: 3 FMA-pipe + 1 logic instruction per TWO elements.
//   * stochastic rounding keeps the reference's order of operations, rint(fl(fl(u - 0.5) + t / 2^p)) (bfp_ops.py:22-23); the
//     clamp to +-(2^m - 1) -- only r = +-2^m can exceed it -- is r - ((r K1 + 1.5 2^23) - 1.5 2^23) with
//     K1 = 2^-m (1/2 + 2^-(m+2)): the bracket rounds to +-1 exactly when |r| = 2^m and to 0 otherwise (three FMA-pipe
//     instructions instead of two FMNMX), and keeps -0.0.
//   * the N:4 mask works on sign bits of key differences (nm_mask4_bits / nm_mask4_packed16 in bfp_common.cuh).
// Blocks outside the exact range (all-zero fp16, Inf / NaN, denormal scale, huge values) fall back, per vector, to the literal
// evaluation in quant_elt_slow.  Results are bit-identical to process_vec (tests run both against the oracle and golden files).
// ---------------------------------------------------------------------------------------------------------------
template <int DT> struct PackedLimits;          // exponent range of p = e - m for which C = 2^(p + MB) is a normal number of the type
template <> struct PackedLimits<BFP_DT_F32> { static constexpr int kMB = 23, kPMax = 100, kMMax = 23; };
template <> struct PackedLimits<BFP_DT_F16> { static constexpr int kMB = 10, kPMax = 5, kMMax = 10; };
template <> struct PackedLimits<BFP_DT_BF16> { static constexpr int kMB = 7, kPMax = 119, kMMax = 7; };

__device__ __forceinline__ uint32_t hadd2_bits(uint32_t a, uint32_t b, bool bf16) {
    uint32_t d;
    if (bf16) asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    else asm("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t hsub2_bits(uint32_t a, uint32_t b, bool bf16) {
    uint32_t d;
    if (bf16) asm("sub.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    else asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t hmin2_bits(uint32_t a, uint32_t b, bool bf16) {
    uint32_t d;
    if (bf16) asm("min.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    else asm("min.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

template <int DT, int ORDER, int M, int KD, int TIE, bool STOC>
struct TilePath {
    static constexpr bool kMaskOk = (ORDER == BFP_ORDER_QUANT_ONLY) || (M == 4 && KD > 0 && TIE == BFP_TIE_TORCH_CUDA);
    // fp32 with nearest rounding keeps the per-vector code: it is already at the copy roofline with 32 registers and 8 CTAs per SM,
    // and measured 9 % faster back to back than the tile path (40 registers, 6 CTAs): profiles/r02_tune_quant_variants.log
    static constexpr bool kEnabled = kMaskOk && (STOC || DT != BFP_DT_F32);
};

// General (rare) blocks of the tile path: exponent needs the real log2 / the literal half-precision formulas, all-zero or
// non-finite blocks, scales outside the exact range.  One vector (already masked when the mask comes first), out of line.
struct VecOut { uint4 a, b; };
template <int DT, bool STOC>
__device__ __noinline__ VecOut quant_vec_general(uint4 wv, uint32_t abits, int m, float eps, uint4 smp_lo, uint4 smp_hi) {
    using D = DType<DT>;
    constexpr int V = D::kVec;
    float v[V];
    unpack_vec<DT>(wv, v);
    const BlockScale sc = make_scale<DT>(abits, m, eps);
    float smp[8] = {__uint_as_float(smp_lo.x), __uint_as_float(smp_lo.y), __uint_as_float(smp_lo.z), __uint_as_float(smp_lo.w),
                    __uint_as_float(smp_hi.x), __uint_as_float(smp_hi.y), __uint_as_float(smp_hi.z), __uint_as_float(smp_hi.w)};
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = quant_elt<DT, STOC>(v[i], sc, STOC ? smp[i] : 0.0f);
    VecOut o;
    if (STOC) { o.a = pack_vec<BFP_DT_F32>(v); o.b = pack_vec<BFP_DT_F32>(v + (V == 8 ? 4 : 0)); }
    else { o.a = pack_vec<DT>(v); o.b = o.a; }
    return o;
}

// Per-launch constants of the simple path (uniform: functions of mant_bits only).
//   A block is "simple" when its exponent is e = k + 1 with k = floor(log2 s), s = max|t| + eps, WITHOUT evaluating a
//   logarithm -- fp32: the mantissa of s is more than 128 ulp above 2^k (bfp_common.cuh make_scale); fp16 / bf16: the mantissa
//   is at or above the tabulated step of k -- and k lies in the range where every constant below is an exact normal number.
//   Then p = k + 1 - m and max|t| < 2^e, so |t / 2^p| < 2^m for every element of the block.
template <int DT, bool STOC>
struct SimpleConsts {
    uint32_t lo_bits, span_bits;   // simple iff bits(s) - lo_bits < span_bits  (s > 0: its bit pattern is monotone in s)
    __device__ __forceinline__ SimpleConsts(int m) {
        using PL = PackedLimits<DT>;
        using D = DType<DT>;
        int klo = max(D::kMinExp, D::kMinScaleExp + m - 1);                       // p = k + 1 - m >= kMinScaleExp
        int khi = min(D::kMaxExp - 1, PL::kPMax + m - 1);                         // e = k + 1 <= kMaxExp, p <= kPMax
        const bool m_ok = m >= 1 && m <= (STOC ? min(20, D::kMaxMant) : PL::kMMax);
        if (!m_ok || khi < klo) { klo = 0; khi = -1; }
        lo_bits = (uint32_t)(klo + 127) << 23;
        span_bits = (uint32_t)(khi + 1 - klo) << 23;                              // 0 = never simple
    }
};

template <int DT, int ORDER, int M, int KD, int TIE, bool STOC, class Store>
__device__ __forceinline__ void process_tile(const uint4* raw, const StreamParams& p, const int64_t* vidx, Store&& store) {
    using D = DType<DT>;
    using PL = PackedLimits<DT>;
    constexpr int V = D::kVec;
    constexpr int U = kStreamUnroll;
    constexpr bool kQuant = ORDER != BFP_ORDER_SPARSIFY_ONLY;
    constexpr bool kSparseFirst = ORDER == BFP_ORDER_SPARSIFY_QUANT || ORDER == BFP_ORDER_SPARSIFY_ONLY;
    constexpr bool kSparseLast = ORDER == BFP_ORDER_QUANT_SPARSIFY;
    constexpr bool kHalf = DT != BFP_DT_F32;
    constexpr bool kBf16 = DT == BFP_DT_BF16;
    constexpr int kOutVecs = (STOC && V == 8) ? 2 : 1;
    constexpr int KDD = KD > 0 ? KD : 1;
    uint32_t w[U][4];
    uint32_t amax[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        w[u][0] = raw[u].x; w[u][1] = raw[u].y; w[u][2] = raw[u].z; w[u][3] = raw[u].w;
        if (kQuant) {
            // block max over the UNMASKED keys: the maximum always survives an N:M mask with N >= 1 (SURVEY.md appendix A.4 i)
            if (kHalf) {
                const uint32_t a01 = __vmaxu2(w[u][0] & 0x7fff7fffu, w[u][1] & 0x7fff7fffu), a23 = __vmaxu2(w[u][2] & 0x7fff7fffu, w[u][3] & 0x7fff7fffu);
                amax[u] = __vmaxu2(a01, a23);                     // (max of the even elements | max of the odd elements << 16)
            } else {
                amax[u] = max(max(w[u][0] & 0x7fffffffu, w[u][1] & 0x7fffffffu), max(w[u][2] & 0x7fffffffu, w[u][3] & 0x7fffffffu));
            }
        }
        if (kSparseFirst && M == 4) {
            if (kHalf) { nm_mask4_packed16<KDD>(w[u][0], w[u][1]); nm_mask4_packed16<KDD>(w[u][2], w[u][3]); }
            else nm_mask4_bits<KDD>(w[u]);
        }
    }
    if (kQuant) {
        // maximum over the lanes that share a block; every lane of the warp takes part
#ifdef BFP_REDUX_MAX
        {
            const uint32_t bmask = block_lane_mask(p.lanes_per_block);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (kHalf) amax[u] = max(amax[u] & 0xffffu, amax[u] >> 16);       // one 16-bit key per lane before the reduction
                amax[u] = block_max_u32(amax[u], p.lanes_per_block, bmask);
            }
        }
#else
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            if (off < p.lanes_per_block) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint32_t o = __shfl_xor_sync(0xffffffffu, amax[u], off);
                    amax[u] = kHalf ? __vmaxu2(amax[u], o) : max(amax[u], o);
                }
            }
        }
#endif
        const SimpleConsts<DT, STOC> sk(p.m);
        const int m = p.m;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            // the vector's uniforms, fl(u - 0.5): generated here, not for the whole tile up front -- 8 (not 32) live registers for
            // 16-bit inputs; the scheduler still overlaps this vector's Philox rounds with the previous vector's rounding
            float un[STOC ? V : 1];
            if (STOC) {
#pragma unroll
                for (int q = 0; q < V / 4; ++q) {
                    const uint4 r = philox4x32_10((uint64_t)(p.ctr_base + vidx[u] * (V / 4) + q), p.offset, p.seed);
                    un[4 * q] = u01_centered(r.x); un[4 * q + 1] = u01_centered(r.y);
                    un[4 * q + 2] = u01_centered(r.z); un[4 * q + 3] = u01_centered(r.w);
                }
            }
            uint32_t abits;
            if (kHalf) {
                const uint32_t a16 = max(amax[u] & 0xffffu, amax[u] >> 16);
                abits = kBf16 ? (a16 << 16) : __float_as_uint(__half2float(__ushort_as_half((unsigned short)a16)));
            } else {
                abits = amax[u];
            }
            const uint32_t sb = __float_as_uint(D::rnd(__uint_as_float(abits) + p.eps));      // bfp_ops.py:33  max_v + epsilon
            bool simple = (sb - sk.lo_bits) < sk.span_bits;
            if (kHalf) {
                constexpr int MB = HalfBits<kHalf ? DT : BFP_DT_F16>::kMant;
                const uint32_t step1 = g_exp_step[HalfBits<kHalf ? DT : BFP_DT_F16>::kTable][((sb >> 23) + 1u) & 0xffu];   // index k + 128 = biased exponent + 1 (Inf / NaN wrap to entry 0 = not tabulated)
                const uint32_t f = (sb >> (23 - MB)) & ((1u << MB) - 1u);
                simple = simple && step1 != 0u && f + 1u >= step1;
            } else {
                simple = simple && (sb & 0x7fffffu) > 128u;
            }
            const uint32_t ebits = sb & 0x7f800000u;                                  // 2^k
            uint4 o0, o1;
            if (simple) {
                if (!STOC && kHalf) {
                    // packed pairs: C = 2^(p + MB) and (2^m - 1) 2^p in the tensor's own 16-bit format, both halves
                    uint32_t c16;
                    if (kBf16) c16 = (ebits >> 16) + (uint32_t)((1 - m + PL::kMB) << 7);
                    else c16 = (((ebits >> 23) - 127u + 15u + (uint32_t)(1 - m + PL::kMB)) << 10);
                    const uint32_t c2 = c16 * 0x00010001u;
                    // (2^m - 1) 2^p = C (2^m - 1) 2^-MB, the factor exact in the type (m <= MB significant bits)
                    const float vf = (float)((1 << m) - 1) * (kBf16 ? 0.0078125f : 0.0009765625f);
                    const uint32_t vk16 = kBf16 ? (__float_as_uint(vf) >> 16) : (uint32_t)__half_as_ushort(__float2half_rn(vf));
                    uint32_t v2;
                    {
                        const uint32_t vk2 = vk16 * 0x00010001u;
                        if (kBf16) asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(v2) : "r"(c2), "r"(vk2));
                        else asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(v2) : "r"(c2), "r"(vk2));
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint32_t a = w[u][i] & 0x7fff7fffu;
                        const uint32_t r = hmin2_bits(hsub2_bits(hadd2_bits(a, c2, kBf16), c2, kBf16), v2, kBf16);
                        w[u][i] = r | (w[u][i] & 0x80008000u);
                    }
                    o0 = make_uint4(w[u][0], w[u][1], w[u][2], w[u][3]); o1 = o0;
                } else if (!STOC) {
                    const float c = __uint_as_float(ebits + (uint32_t)((1 - m + PL::kMB) << 23));
                    const float vd = c * ((float)((1 << m) - 1) * 1.1920928955078125e-07f);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float t = __uint_as_float(w[u][i]);
                        const float r = fminf((fabsf(t) + c) - c, vd);
                        w[u][i] = __float_as_uint(r) | (w[u][i] & 0x80000000u);
                    }
                    o0 = make_uint4(w[u][0], w[u][1], w[u][2], w[u][3]); o1 = o0;
                } else {
                    // stochastic rounding: fp32 arithmetic and fp32 output for every input type (torch promotion, bfp_ops.py:22-23).
                    // |t / 2^p| < 2^m here, so |r| <= 2^m and the FMA-pipe clamp is exact.
                    float v[V];
                    unpack_vec<DT>(make_uint4(w[u][0], w[u][1], w[u][2], w[u][3]), v);
                    const uint32_t dbits = ebits + (uint32_t)((1 - m) << 23);       // 2^p
                    const float delta = __uint_as_float(dbits), inv = __uint_as_float(0x7f000000u - dbits);
                    const float k1 = __uint_as_float((uint32_t)(127 - m - 1) << 23) + __uint_as_float((uint32_t)(127 - 2 * m - 2) << 23);
#pragma unroll
                    for (int i = 0; i < V; ++i) {
                        const float r = rintf(__fmaf_rn(v[i], inv, un[i]));      // = fl(fl(u - 0.5) + t / 2^p): the product is exact
                        const float c = __fmaf_rn(r, k1, 12582912.0f) - 12582912.0f;
                        v[i] = (r - c) * delta;
                    }
                    o0 = pack_vec<BFP_DT_F32>(v); o1 = pack_vec<BFP_DT_F32>(v + (V == 8 ? 4 : 0));
                }
            } else {
                uint4 s0 = make_uint4(0u, 0u, 0u, 0u), s1 = s0;
                if (STOC) {
                    s0 = make_uint4(__float_as_uint(un[0]), __float_as_uint(un[1]), __float_as_uint(un[2]), __float_as_uint(un[3]));
                    if (V == 8) s1 = make_uint4(__float_as_uint(un[V - 4]), __float_as_uint(un[V - 3]), __float_as_uint(un[V - 2]), __float_as_uint(un[V - 1]));
                }
                const VecOut g = quant_vec_general<DT, STOC>(make_uint4(w[u][0], w[u][1], w[u][2], w[u][3]), abits, m, p.eps, s0, s1);
                o0 = g.a; o1 = g.b;
            }
            if (kSparseLast && M == 4) {
                if (STOC || !kHalf) {
                    uint32_t b[4] = {o0.x, o0.y, o0.z, o0.w};
                    nm_mask4_bits<KDD>(b);
                    o0 = make_uint4(b[0], b[1], b[2], b[3]);
                    if (kOutVecs == 2) {
                        uint32_t c[4] = {o1.x, o1.y, o1.z, o1.w};
                        nm_mask4_bits<KDD>(c);
                        o1 = make_uint4(c[0], c[1], c[2], c[3]);
                    }
                } else {
                    nm_mask4_packed16<KDD>(o0.x, o0.y); nm_mask4_packed16<KDD>(o0.z, o0.w);
                }
            }
            uint4 o[kOutVecs];
            o[0] = o0;
            if (kOutVecs == 2) o[kOutVecs - 1] = o1;
            store(u, o);
        }
    } else {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            uint4 o[kOutVecs];
            o[0] = make_uint4(w[u][0], w[u][1], w[u][2], w[u][3]);
            store(u, o);
        }
    }
}

// The per-vector streaming loop (round 1's kernel body, unchanged): every vector is loaded up front, then processed and stored one
// after the other with 32-bit in-tile indexing -- 32 registers, 8 CTAs per SM.  Used where the tile path is not (fp32 with nearest
// rounding, exotic group sizes, the torch-CPU tie rule).
template <int DT, int ORDER, int M, int KD, int TIE, bool STOC, bool PADDED>
__device__ __forceinline__ void stream_body_per_vector(const StreamParams& p) {
    using D = DType<DT>;
    constexpr int V = D::kVec;
    constexpr bool kQuant = ORDER != BFP_ORDER_SPARSIFY_ONLY;
    constexpr bool kSparseFirst = ORDER == BFP_ORDER_SPARSIFY_QUANT || ORDER == BFP_ORDER_SPARSIFY_ONLY;
    constexpr bool kSparseLast = ORDER == BFP_ORDER_QUANT_SPARSIFY;
    constexpr int kOutVecs = (STOC && V == 8) ? 2 : 1;     // fp32 output of 8 half inputs = two 16-B stores
    constexpr int kTileVecs = kStreamThreads * kStreamUnroll;

    const int64_t n_tiles = (p.n_vec + kTileVecs - 1) / kTileVecs;

    // Programmatic dependent launch: let the next kernel on the stream get its CTAs resident while this one drains, and
    // do not touch global memory before everything earlier on the stream has completed (stream order is preserved).
    pdl_launch_dependents();
    pdl_wait();

    if constexpr (STOC && !PADDED) {
        // Stochastic rounding is ~43 instructions per element, most of them the Philox rounds: a tile's loads are far apart in
        // time unless the NEXT tile's vectors are requested before this tile's arithmetic starts (software prefetch, one tile
        // ahead in registers).  Same vectors, same counters, same results as the plain loop below.
        uint4 nxt[kStreamUnroll];
        auto fetch = [&](int64_t t) {
            const int64_t base = t * kTileVecs;
            const int rem = (int)min((int64_t)kTileVecs, p.n_vec - base);
#pragma unroll
            for (int u = 0; u < kStreamUnroll; ++u) {
                const int li = (int)threadIdx.x + u * kStreamThreads;
                nxt[u] = (li < rem) ? ld_stream(p.in + base + li) : make_uint4(0u, 0u, 0u, 0u);
            }
        };
        if ((int64_t)blockIdx.x < n_tiles) fetch(blockIdx.x);
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t tile_base = tile * kTileVecs;
            const int rem = (int)min((int64_t)kTileVecs, p.n_vec - tile_base);
            uint4 raw[kStreamUnroll];
#pragma unroll
            for (int u = 0; u < kStreamUnroll; ++u) raw[u] = nxt[u];
            if (tile + gridDim.x < n_tiles) fetch(tile + gridDim.x);
#pragma unroll
            for (int u = 0; u < kStreamUnroll; ++u) {
                const int li = (int)threadIdx.x + u * kStreamThreads;
                uint4 o[kOutVecs];
                process_vec<DT, ORDER, M, KD, TIE, STOC>(raw[u], p, tile_base + li, o);
                if (li < rem) {
                    uint4* dst = p.out + (tile_base + li) * kOutVecs;
                    st_stream(dst, o[0]);
                    if (kOutVecs == 2) st_stream(dst + 1, o[kOutVecs - 1]);
                }
            }
        }
        return;
    }
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t tile_base = tile * kTileVecs;
        uint4 raw[kStreamUnroll];
        if constexpr (!PADDED) {
            // flat mode: 32-bit in-tile indexing, one bounds compare per vector
            const int rem = (int)min((int64_t)kTileVecs, p.n_vec - tile_base);   // vectors of this tile that exist
            const uint4* src = p.in + tile_base;
#pragma unroll
            for (int u = 0; u < kStreamUnroll; ++u) {
                const int li = (int)threadIdx.x + u * kStreamThreads;
                raw[u] = (li < rem) ? ld_stream(src + li) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < kStreamUnroll; ++u) {
                const int li = (int)threadIdx.x + u * kStreamThreads;
                uint4 o[kOutVecs];
                process_vec<DT, ORDER, M, KD, TIE, STOC>(raw[u], p, tile_base + li, o);
                if (li < rem) {
                    uint4* dst = p.out + (tile_base + li) * kOutVecs;
                    st_stream(dst, o[0]);
                    if (kOutVecs == 2) st_stream(dst + 1, o[kOutVecs - 1]);
                }
            }
        } else {
            int64_t rv[kStreamUnroll];                                        // real vector index, -1 = padding / out of range
#pragma unroll
            for (int u = 0; u < kStreamUnroll; ++u) {
                rv[u] = real_vec(p, tile_base + (int)threadIdx.x + u * kStreamThreads);
                raw[u] = rv[u] >= 0 ? ld_stream(p.in + rv[u]) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < kStreamUnroll; ++u) {
                uint4 o[kOutVecs];
                process_vec<DT, ORDER, M, KD, TIE, STOC>(raw[u], p, rv[u], o);
                if (rv[u] >= 0) {
                    uint4* dst = p.out + rv[u] * kOutVecs;
                    st_stream(dst, o[0]);
                    if (kOutVecs == 2) st_stream(dst + 1, o[kOutVecs - 1]);
                }
            }
        }
    }
}

// kStreamUnroll vectors of one thread: the tile path where it applies, else vector by vector.  store(u, o) receives vector u's
// kOutVecs output vectors as soon as they are final.
template <int DT, int ORDER, int M, int KD, int TIE, bool STOC, class Store>
__device__ __forceinline__ void process_vectors(const uint4* raw, const StreamParams& p, const int64_t* vidx, Store&& store) {
    constexpr int kOutVecs = (STOC && DType<DT>::kVec == 8) ? 2 : 1;
    if constexpr (TilePath<DT, ORDER, M, KD, TIE, STOC>::kEnabled) {
        process_tile<DT, ORDER, M, KD, TIE, STOC>(raw, p, vidx, store);
    } else {
#pragma unroll
        for (int u = 0; u < kStreamUnroll; ++u) {
            uint4 o[kOutVecs];
            process_vec<DT, ORDER, M, KD, TIE, STOC>(raw[u], p, vidx[u], o);
            store(u, o);
        }
    }
}

// ORDER: BFP_ORDER_*.  M / KD / TIE: see mask_vec.  STOC: stochastic rounding (fp32 output).
#ifndef BFP_STOC_MIN_CTAS
#define BFP_STOC_MIN_CTAS 3
#endif
#ifndef BFP_HALF_MIN_CTAS
#define BFP_HALF_MIN_CTAS 4
#endif
template <int DT, bool STOC> struct StreamOcc { static constexpr int kMinCtas = STOC ? BFP_STOC_MIN_CTAS : (DT == BFP_DT_F32 ? 1 : BFP_HALF_MIN_CTAS); };
template <int DT, int ORDER, int M, int KD, int TIE, bool STOC, bool PADDED>
__global__ void __launch_bounds__(kStreamThreads, StreamOcc<DT, STOC>::kMinCtas) quant_stream_kernel(const StreamParams p) {
    using D = DType<DT>;
    constexpr int V = D::kVec;
    constexpr int kOutVecs = (STOC && V == 8) ? 2 : 1;     // fp32 output of 8 half inputs = two 16-B stores
    constexpr int kTileVecs = kStreamThreads * kStreamUnroll;

    if constexpr (!TilePath<DT, ORDER, M, KD, TIE, STOC>::kEnabled) {
        stream_body_per_vector<DT, ORDER, M, KD, TIE, STOC, PADDED>(p);
        return;
    }
    const int64_t n_tiles = (p.n_vec + kTileVecs - 1) / kTileVecs;

    // Programmatic dependent launch: let the next kernel on the stream get its CTAs resident while this one drains, and
    // do not touch global memory before everything earlier on the stream has completed (stream order is preserved).
    pdl_launch_dependents();
    pdl_wait();

    if constexpr (STOC && !PADDED) {
        // Stochastic rounding is instruction-heavy (Philox4x32-10): a tile's loads are far apart in time unless the NEXT tile's
        // vectors are requested before this tile's arithmetic starts (software prefetch, one tile ahead in registers).
        uint4 nxt[kStreamUnroll];
        auto fetch = [&](int64_t t) {
            const int64_t base = t * kTileVecs;
            const int rem = (int)min((int64_t)kTileVecs, p.n_vec - base);
#pragma unroll
            for (int u = 0; u < kStreamUnroll; ++u) {
                const int li = (int)threadIdx.x + u * kStreamThreads;
                nxt[u] = (li < rem) ? ld_stream(p.in + base + li) : make_uint4(0u, 0u, 0u, 0u);
            }
        };
#ifdef BFP_STOC_PINGPONG
        // two named register buffers alternate between "being processed" and "being fetched", so the hand-over needs no copies
        // (the single-buffer form moves 16 registers per tile)
        uint4 alt[kStreamUnroll];
        auto fetch_to = [&](uint4* dst, int64_t t) {
            const int64_t base = t * kTileVecs;
            const int rem = (int)min((int64_t)kTileVecs, p.n_vec - base);
#pragma unroll
            for (int u = 0; u < kStreamUnroll; ++u) {
                const int li = (int)threadIdx.x + u * kStreamThreads;
                dst[u] = (li < rem) ? ld_stream(p.in + base + li) : make_uint4(0u, 0u, 0u, 0u);
            }
        };
        auto run = [&](const uint4* raw, int64_t tile) {
            const int64_t tile_base = tile * kTileVecs;
            const int rem = (int)min((int64_t)kTileVecs, p.n_vec - tile_base);
            int64_t vidx[kStreamUnroll];
#pragma unroll
            for (int u = 0; u < kStreamUnroll; ++u) vidx[u] = tile_base + (int)threadIdx.x + u * kStreamThreads;
            process_vectors<DT, ORDER, M, KD, TIE, STOC>(raw, p, vidx, [&](int u, const uint4* o) {
                const int li = (int)threadIdx.x + u * kStreamThreads;
                if (li < rem) {
                    uint4* dst = p.out + (tile_base + li) * kOutVecs;
                    st_stream(dst, o[0]);
                    if (kOutVecs == 2) st_stream(dst + 1, o[kOutVecs - 1]);
                }
            });
        };
        int64_t tile = blockIdx.x;
        if (tile < n_tiles) fetch_to(nxt, tile);
        while (tile < n_tiles) {
            const int64_t t1 = tile + gridDim.x;
            if (t1 < n_tiles) fetch_to(alt, t1);
            run(nxt, tile);
            if (t1 >= n_tiles) break;
            const int64_t t2 = t1 + gridDim.x;
            if (t2 < n_tiles) fetch_to(nxt, t2);
            run(alt, t1);
            tile = t2;
        }
#else
        if ((int64_t)blockIdx.x < n_tiles) fetch(blockIdx.x);
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t tile_base = tile * kTileVecs;
            const int rem = (int)min((int64_t)kTileVecs, p.n_vec - tile_base);
            uint4 raw[kStreamUnroll];
            int64_t vidx[kStreamUnroll];
#pragma unroll
            for (int u = 0; u < kStreamUnroll; ++u) { raw[u] = nxt[u]; vidx[u] = tile_base + (int)threadIdx.x + u * kStreamThreads; }
            if (tile + gridDim.x < n_tiles) fetch(tile + gridDim.x);
            process_vectors<DT, ORDER, M, KD, TIE, STOC>(raw, p, vidx, [&](int u, const uint4* o) {
                const int li = (int)threadIdx.x + u * kStreamThreads;
                if (li < rem) {
                    uint4* dst = p.out + (tile_base + li) * kOutVecs;
                    st_stream(dst, o[0]);
                    if (kOutVecs == 2) st_stream(dst + 1, o[kOutVecs - 1]);
                }
            });
        }
#endif
        return;
    }
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t tile_base = tile * kTileVecs;
        uint4 raw[kStreamUnroll];
        int64_t vidx[kStreamUnroll];
        if constexpr (!PADDED) {
            // flat mode: 32-bit in-tile indexing, one bounds compare per vector
            const int rem = (int)min((int64_t)kTileVecs, p.n_vec - tile_base);   // vectors of this tile that exist
            const uint4* src = p.in + tile_base;
#pragma unroll
            for (int u = 0; u < kStreamUnroll; ++u) {
                const int li = (int)threadIdx.x + u * kStreamThreads;
                raw[u] = (li < rem) ? ld_stream(src + li) : make_uint4(0u, 0u, 0u, 0u);
                vidx[u] = tile_base + li;
            }
            process_vectors<DT, ORDER, M, KD, TIE, STOC>(raw, p, vidx, [&](int u, const uint4* o) {
                const int li = (int)threadIdx.x + u * kStreamThreads;
                if (li < rem) {
                    uint4* dst = p.out + (tile_base + li) * kOutVecs;
                    st_stream(dst, o[0]);
                    if (kOutVecs == 2) st_stream(dst + 1, o[kOutVecs - 1]);
                }
            });
        } else {
#pragma unroll
            for (int u = 0; u < kStreamUnroll; ++u) {
                vidx[u] = real_vec(p, tile_base + (int)threadIdx.x + u * kStreamThreads);     // -1 = padding / out of range
                raw[u] = vidx[u] >= 0 ? ld_stream(p.in + vidx[u]) : make_uint4(0u, 0u, 0u, 0u);
            }
            process_vectors<DT, ORDER, M, KD, TIE, STOC>(raw, p, vidx, [&](int u, const uint4* o) {
                if (vidx[u] >= 0) {
                    uint4* dst = p.out + vidx[u] * kOutVecs;
                    st_stream(dst, o[0]);
                    if (kOutVecs == 2) st_stream(dst + 1, o[kOutVecs - 1]);
                }
            });
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// TMA-staged variant of the stream kernel: a producer warp moves 16 KB tiles global -> shared with cp.async.bulk (the copy
// engine, completion on an mbarrier) into a ring of kTmaStages tiles; eight consumer warps read their vectors from shared
// memory (conflict-free 128-bit loads), hand the slot back at once, and run the same per-vector code.  Loads are issued
// kTmaStages tiles ahead of the arithmetic by one thread instead of by every thread just before use.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kTmaStages = 6;
constexpr int kTmaUnroll = 1;                      // vectors per consumer thread per tile: 4 KB tiles keep 7 CTAs (1792 consumers) per SM
constexpr int kTmaThreads = kStreamThreads + 32;

__device__ __forceinline__ uint32_t q_smem_u32(const void* ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ void q_mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = q_smem_u32(bar);
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity), "r"(100000u) : "memory");
}

template <int DT, int ORDER, int M, int KD, int TIE, bool STOC>
__global__ void __launch_bounds__(kTmaThreads) quant_tma_kernel(const StreamParams p) {
    using D = DType<DT>;
    constexpr int V = D::kVec;
    constexpr int kOutVecs = (STOC && V == 8) ? 2 : 1;
    constexpr int kTileVecs = kStreamThreads * kTmaUnroll;
    extern __shared__ __align__(128) uint8_t q_dyn_smem[];
    uint4 (*ring)[kTileVecs] = reinterpret_cast<uint4 (*)[kTileVecs]>(q_dyn_smem);
    uint64_t* full = reinterpret_cast<uint64_t*>(q_dyn_smem + kTmaStages * kTileVecs * 16);
    uint64_t* empty = full + kTmaStages;

    const int64_t n_tiles = (p.n_vec + kTileVecs - 1) / kTileVecs;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kTmaStages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(q_smem_u32(&full[s])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(q_smem_u32(&empty[s])), "r"(kStreamThreads / 32));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_launch_dependents();
    pdl_wait();

    if (warp == kStreamThreads / 32) {
        // ---- producer: one lane issues the bulk copies ----
        if (lane == 0) {
            int s = 0; uint32_t phase = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int64_t tile_base = tile * kTileVecs;
                const uint32_t bytes = (uint32_t)min((int64_t)kTileVecs, p.n_vec - tile_base) * 16u;
                q_mbar_wait(&empty[s], phase ^ 1);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(q_smem_u32(&full[s])), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(q_smem_u32(&ring[s][0])), "l"(p.in + tile_base), "r"(bytes), "r"(q_smem_u32(&full[s])) : "memory");
                if (++s == kTmaStages) { s = 0; phase ^= 1; }
            }
        }
        return;
    }
    // ---- consumers ----
    int s = 0; uint32_t phase = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t tile_base = tile * kTileVecs;
        const int rem = (int)min((int64_t)kTileVecs, p.n_vec - tile_base);
        q_mbar_wait(&full[s], phase);
        uint4 raw[kTmaUnroll];
#pragma unroll
        for (int u = 0; u < kTmaUnroll; ++u) {
            const int li = (int)threadIdx.x + u * kStreamThreads;
            raw[u] = (li < rem) ? ring[s][li] : make_uint4(0u, 0u, 0u, 0u);
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(q_smem_u32(&empty[s])) : "memory");   // slot read: refill it
        if (++s == kTmaStages) { s = 0; phase ^= 1; }
#pragma unroll
        for (int u = 0; u < kTmaUnroll; ++u) {
            const int li = (int)threadIdx.x + u * kStreamThreads;
            uint4 o[kOutVecs];
            process_vec<DT, ORDER, M, KD, TIE, STOC>(raw[u], p, tile_base + li, o);
            if (li < rem) {
                uint4* dst = p.out + (tile_base + li) * kOutVecs;
                st_stream(dst, o[0]);
                if (kOutVecs == 2) st_stream(dst + 1, o[kOutVecs - 1]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// generic kernel: gather-style, any K / block_size / N:M.  unit = one block (orders q, s->q) or one group (s, q->s).
// ---------------------------------------------------------------------------------------------------------------
struct GenericParams {
    const void* in;
    void* out;
    int64_t rows, K;
    int B, m;
    float eps;
    int N, M, tie;
    uint64_t seed, offset;
    int64_t index_base;
};

template <int DT, int ORDER, bool STOC>
__global__ void __launch_bounds__(128) quant_generic_kernel(const GenericParams p) {
    using D = DType<DT>;
    using DO = DType<STOC ? BFP_DT_F32 : DT>;
    const int64_t nblk = (p.K + p.B - 1) / p.B;
    const int64_t ngrp = (ORDER == BFP_ORDER_QUANT_ONLY) ? 0 : (p.K + p.M - 1) / p.M;
    const bool per_group = (ORDER == BFP_ORDER_SPARSIFY_ONLY || ORDER == BFP_ORDER_QUANT_SPARSIFY);
    const int64_t units = p.rows * (per_group ? ngrp : nblk);
    for (int64_t unit = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; unit < units; unit += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = unit / (per_group ? ngrp : nblk);
        const int64_t idx = unit % (per_group ? ngrp : nblk);
        auto raw = [&](int64_t c) { return ld_pad<DT>(p.in, row, c, p.K); };
        auto uni = [&](int64_t c) {
            const uint64_t flat = (uint64_t)(p.index_base + row * p.K + c);
            const uint4 r = philox4x32_10(flat >> 2, p.offset, p.seed);
            const uint32_t w = (flat & 3) == 0 ? r.x : ((flat & 3) == 1 ? r.y : ((flat & 3) == 2 ? r.z : r.w));
            return u01_centered(w);
        };
        if (ORDER == BFP_ORDER_SPARSIFY_ONLY) {
            for (int64_t c = idx * p.M; c < min(p.K, idx * p.M + p.M); ++c)
                D::store(p.out, row * p.K + c, nm_dropped(raw, c, p.N, p.M, p.tie) ? 0.0f : raw(c));
        } else if (ORDER == BFP_ORDER_QUANT_ONLY || ORDER == BFP_ORDER_SPARSIFY_QUANT) {
            auto src = [&](int64_t c) {
                const float t = raw(c);
                if (ORDER == BFP_ORDER_SPARSIFY_QUANT) return (c < p.K && nm_dropped(raw, c, p.N, p.M, p.tie)) ? 0.0f : t;
                return t;
            };
            const int64_t c0 = idx * p.B, c1 = min(p.K, c0 + p.B);
            uint32_t amax = 0u;
            for (int64_t c = c0; c < c1; ++c) amax = max(amax, abs_bits(src(c)));
            const BlockScale sc = make_scale<DT>(amax, p.m, p.eps);
            for (int64_t c = c0; c < c1; ++c)
                DO::store(p.out, row * p.K + c, quant_elt<DT, STOC>(src(c), sc, STOC ? uni(c) : 0.0f));
        } else {   // QUANT_SPARSIFY: quantise the group's elements (each with its own block's scale), then mask
            float q[kMaxGroup];
            const int64_t g0 = idx * p.M;
            int64_t cur_blk = -1;
            BlockScale sc = {};
            for (int j = 0; j < p.M; ++j) {
                const int64_t c = g0 + j;
                if (c >= p.K) { q[j] = 0.0f; continue; }          // re-padded after the narrow: plain zeros
                const int64_t b = c / p.B;
                if (b != cur_blk) {
                    cur_blk = b;
                    uint32_t amax = 0u;
                    for (int64_t cc = b * p.B; cc < min(p.K, b * p.B + p.B); ++cc) amax = max(amax, abs_bits(raw(cc)));
                    sc = make_scale<DT>(amax, p.m, p.eps);
                }
                q[j] = quant_elt<DT, STOC>(raw(c), sc, STOC ? uni(c) : 0.0f);
            }
            auto qsrc = [&](int64_t c) { return q[c - g0]; };
            for (int j = 0; j < p.M && g0 + j < p.K; ++j)
                DO::store(p.out, row * p.K + g0 + j, nm_dropped(qsrc, g0 + j, p.N, p.M, p.tie) ? 0.0f : q[j]);
        }
    }
}

// get_exponent per block (bfp_ops.py:29-33), generic layout
template <int DT>
__global__ void __launch_bounds__(128) block_exponent_kernel(const void* in, float* e_out, int64_t rows, int64_t K, int B, float eps) {
    const int64_t nblk = (K + B - 1) / B;
    for (int64_t unit = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; unit < rows * nblk; unit += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = unit / nblk, kb = unit % nblk;
        uint32_t amax = 0u;
        for (int64_t c = kb * B; c < min(K, kb * B + B); ++c) amax = max(amax, abs_bits(DType<DT>::load(in, row * K + c)));
        // m = 1 keeps make_scale on its fast path whenever possible; e does not depend on m
        e_out[unit] = make_scale<DT>(amax, 1, eps).e;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host-side dispatch
// ---------------------------------------------------------------------------------------------------------------
static inline bool is_pow2(int64_t x) { return x > 0 && (x & (x - 1)) == 0; }

template <int DT, int ORDER, int M, int KD, int TIE, bool STOC>
static int launch_stream_t(const StreamParams& p, cudaStream_t st) {
    const int64_t tile_vecs = (int64_t)kStreamThreads * kStreamUnroll;
    const int64_t n_tiles = (p.n_vec + tile_vecs - 1) / tile_vecs;
    if (n_tiles == 0) return BFP_OK;
    const DeviceInfo& di = device_info();
    if (p.slots_per_row != 0) {
        static const int occ_p = kernel_occupancy(quant_stream_kernel<DT, ORDER, M, KD, TIE, STOC, true>, kStreamThreads);
        if (int rc = launch_pdl(quant_stream_kernel<DT, ORDER, M, KD, TIE, STOC, true>, stream_grid(occ_p, n_tiles), kStreamThreads, st, p)) return rc;
        count_launch();
        return check_launch("quant_stream_kernel (padded rows)");
    }
    static const int occ = kernel_occupancy(quant_stream_kernel<DT, ORDER, M, KD, TIE, STOC, false>, kStreamThreads);
    const int grid = stream_grid(occ, n_tiles);
    (void)di;
    if (tuning().quant_tma && p.slots_per_row == 0) {
        constexpr int kTmaSmem = kTmaStages * kStreamThreads * kTmaUnroll * 16 + 2 * kTmaStages * 8;
        static const int occ_t = [] {
            cudaFuncSetAttribute(quant_tma_kernel<DT, ORDER, M, KD, TIE, STOC>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTmaSmem);
            int occ = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, quant_tma_kernel<DT, ORDER, M, KD, TIE, STOC>, kTmaThreads, kTmaSmem) != cudaSuccess || occ < 1) occ = 2;
            return occ;
        }();
        const int64_t n_tiles_t = (p.n_vec + kStreamThreads * kTmaUnroll - 1) / (kStreamThreads * kTmaUnroll);
        const int grid_t = (int)std::min<int64_t>(n_tiles_t, (int64_t)di.sm_count * occ_t);
        if (int rc = launch_pdl(quant_tma_kernel<DT, ORDER, M, KD, TIE, STOC>, grid_t, kTmaThreads, st, p, kTmaSmem)) return rc;
        count_launch();
        return check_launch("quant_tma_kernel");
    }
    if (int rc = launch_pdl(quant_stream_kernel<DT, ORDER, M, KD, TIE, STOC, false>, grid, kStreamThreads, st, p)) return rc;
    count_launch();
    return check_launch("quant_stream_kernel");
}

template <int DT, int ORDER, bool STOC>
static int launch_stream_m(const StreamParams& p, int M, int tie, cudaStream_t st) {
    constexpr int V = DType<DT>::kVec;
    if (ORDER == BFP_ORDER_QUANT_ONLY) return launch_stream_t<DT, ORDER, 0, 0, BFP_TIE_TORCH_CUDA, STOC>(p, st);
    if (tie == BFP_TIE_TORCH_CPU) {
        if (M == 4 && p.kdrop == 2) return launch_stream_t<DT, ORDER, 4, 0, BFP_TIE_TORCH_CPU, STOC>(p, st);
        return set_error(BFP_E_UNSUPPORTED, "BFP_TIE_TORCH_CPU is implemented for N:M = 2:4 only");
    }
    switch (M) {
    case 1: return launch_stream_t<DT, ORDER, 1, 0, BFP_TIE_TORCH_CUDA, STOC>(p, st);
    case 2: return launch_stream_t<DT, ORDER, 2, 0, BFP_TIE_TORCH_CUDA, STOC>(p, st);
    case 4:
        switch (p.kdrop) {
        case 0: return launch_stream_t<DT, ORDER, 1, 0, BFP_TIE_TORCH_CUDA, STOC>(p, st);   // 4:4 keeps everything
        case 1: return launch_stream_t<DT, ORDER, 4, 1, BFP_TIE_TORCH_CUDA, STOC>(p, st);
        case 2: return launch_stream_t<DT, ORDER, 4, 2, BFP_TIE_TORCH_CUDA, STOC>(p, st);
        case 3: return launch_stream_t<DT, ORDER, 4, 3, BFP_TIE_TORCH_CUDA, STOC>(p, st);
        }
        break;
    case 8: if (V == 8) return launch_stream_t<DT, ORDER, (V == 8 ? 8 : 4), 0, BFP_TIE_TORCH_CUDA, STOC>(p, st);
    }
    return set_error(BFP_E_UNSUPPORTED, "internal: stream path called with unsupported M");
}

template <int DT, bool STOC>
static int launch_stream_o(const StreamParams& p, int order, int M, int tie, cudaStream_t st) {
    switch (order) {
    case BFP_ORDER_QUANT_ONLY: return launch_stream_m<DT, BFP_ORDER_QUANT_ONLY, STOC>(p, M, tie, st);
    case BFP_ORDER_SPARSIFY_QUANT: return launch_stream_m<DT, BFP_ORDER_SPARSIFY_QUANT, STOC>(p, M, tie, st);
    case BFP_ORDER_QUANT_SPARSIFY: return launch_stream_m<DT, BFP_ORDER_QUANT_SPARSIFY, STOC>(p, M, tie, st);
    case BFP_ORDER_SPARSIFY_ONLY:
        if (STOC) break;
        return launch_stream_m<DT, BFP_ORDER_SPARSIFY_ONLY, false>(p, M, tie, st);
    }
    return set_error(BFP_E_ARG, "bad order");
}

template <int DT, int ORDER, bool STOC>
static int launch_generic_t(const GenericParams& p, cudaStream_t st) {
    const bool per_group = (ORDER == BFP_ORDER_SPARSIFY_ONLY || ORDER == BFP_ORDER_QUANT_SPARSIFY);
    const int64_t units = p.rows * (per_group ? (p.K + p.M - 1) / p.M : (p.K + p.B - 1) / p.B);
    if (units == 0) return BFP_OK;
    const int grid = (int)std::min<int64_t>((units + 127) / 128, (int64_t)device_info().sm_count * 16);
    quant_generic_kernel<DT, ORDER, STOC><<<grid, 128, 0, st>>>(p);
    count_launch();
    return check_launch("quant_generic_kernel");
}

template <int DT, bool STOC>
static int launch_generic_o(const GenericParams& p, int order, cudaStream_t st) {
    switch (order) {
    case BFP_ORDER_QUANT_ONLY: return launch_generic_t<DT, BFP_ORDER_QUANT_ONLY, STOC>(p, st);
    case BFP_ORDER_SPARSIFY_QUANT: return launch_generic_t<DT, BFP_ORDER_SPARSIFY_QUANT, STOC>(p, st);
    case BFP_ORDER_QUANT_SPARSIFY: return launch_generic_t<DT, BFP_ORDER_QUANT_SPARSIFY, STOC>(p, st);
    case BFP_ORDER_SPARSIFY_ONLY:
        if (STOC) break;
        return launch_generic_t<DT, BFP_ORDER_SPARSIFY_ONLY, false>(p, st);
    }
    return set_error(BFP_E_ARG, "bad order");
}

template <int DT>
static int quantize_dt(const QuantArgs& a, cudaStream_t st) {
    constexpr int V = DType<DT>::kVec;
    const bool stoc = a.rounding == BFP_ROUND_STOCHASTIC && a.order != BFP_ORDER_SPARSIFY_ONLY;
    const bool sparse = a.order != BFP_ORDER_QUANT_ONLY;
    const bool quant = a.order != BFP_ORDER_SPARSIFY_ONLY;
    const int64_t numel = a.rows * a.K;
    // stream-path eligibility
    bool fast = (numel % V == 0) && (a.index_base % 4 == 0) && (reinterpret_cast<uintptr_t>(a.in) % 16 == 0) && (reinterpret_cast<uintptr_t>(a.out) % 16 == 0);
    if (quant) fast = fast && is_pow2(a.B) && a.B >= V && a.B <= 32 * V && (a.K % a.B == 0);
    if (sparse) fast = fast && is_pow2(a.M) && a.M <= V && (a.K % a.M == 0) && !(a.tie == BFP_TIE_TORCH_CPU && !(a.M == 4 && a.N == 2));
    if (!quant) fast = fast && (a.K % V == 0 || true);      // groups never straddle vectors: M | V and M | K
    // padded-row mode: rows are vector-aligned (K % V == 0) but not block-aligned -- the ViT patch-embedding input (K = 224),
    // conv weights (K = kw < B), odd widths like 4100
    bool padded = false;
    if (!fast && quant && !tuning().force_generic) {
        padded = (a.K % V == 0) && (a.index_base % 4 == 0) && (reinterpret_cast<uintptr_t>(a.in) % 16 == 0) && (reinterpret_cast<uintptr_t>(a.out) % 16 == 0) &&
                 is_pow2(a.B) && a.B >= V && a.B <= 32 * V && (a.K % a.B != 0) && a.K / V < (int64_t)1 << 31 && a.rows < (int64_t)1 << 31;
        if (sparse) padded = padded && is_pow2(a.M) && a.M <= V && (a.K % a.M == 0) && !(a.tie == BFP_TIE_TORCH_CPU && !(a.M == 4 && a.N == 2));
    }
    if (tuning().force_generic) fast = false;
    if (fast || padded) {
        StreamParams p;
        p.in = static_cast<const uint4*>(a.in);
        p.out = static_cast<uint4*>(a.out);
        p.n_vec = numel / V;
        p.vec_per_row = p.slots_per_row = 0;
        if (padded) {
            p.vec_per_row = (uint32_t)(a.K / V);
            p.slots_per_row = (uint32_t)(round_up(a.K, a.B) / V);
            p.n_vec = a.rows * (int64_t)p.slots_per_row;
        }
        p.lanes_per_block = quant ? a.B / V : 1;
        p.m = a.m; p.eps = a.eps; p.kdrop = sparse ? a.M - a.N : 0;
        p.seed = a.seed; p.offset = a.offset; p.ctr_base = a.index_base / 4;
        return stoc ? launch_stream_o<DT, true>(p, a.order, a.M, a.tie, st) : launch_stream_o<DT, false>(p, a.order, a.M, a.tie, st);
    }
    if (sparse && a.tie == BFP_TIE_TORCH_CPU && !(a.M == 4 && a.N == 2))
        return set_error(BFP_E_UNSUPPORTED, "BFP_TIE_TORCH_CPU is implemented for N:M = 2:4 only");
    GenericParams g;
    g.in = a.in; g.out = a.out; g.rows = a.rows; g.K = a.K; g.B = quant ? a.B : 1; g.m = a.m; g.eps = a.eps;
    g.N = a.N; g.M = sparse ? a.M : 1; g.tie = a.tie; g.seed = a.seed; g.offset = a.offset; g.index_base = a.index_base;
    return stoc ? launch_generic_o<DT, true>(g, a.order, st) : launch_generic_o<DT, false>(g, a.order, st);
}

int quantize_device(const QuantArgs& a, cudaStream_t st) {
    if (a.in_dtype != BFP_DT_F32) ensure_exp_tables(st);
    switch (a.in_dtype) {
    case BFP_DT_F32: return quantize_dt<BFP_DT_F32>(a, st);
    case BFP_DT_F16: return quantize_dt<BFP_DT_F16>(a, st);
    case BFP_DT_BF16: return quantize_dt<BFP_DT_BF16>(a, st);
    }
    return set_error(BFP_E_ARG, "bad dtype");
}

int debug_exp_table(int dtype, uint16_t out[256]) {
    if (dtype != BFP_DT_F16 && dtype != BFP_DT_BF16) return set_error(BFP_E_ARG, "fp16 / bf16 only");
    ensure_exp_tables(nullptr);
    uint16_t all[2][256];
    const cudaError_t e = cudaMemcpyFromSymbol(all, g_exp_step, sizeof(all));
    if (e != cudaSuccess) return set_errorf(BFP_E_CUDA, "cudaMemcpyFromSymbol: %s", cudaGetErrorString(e));
    for (int i = 0; i < 256; ++i) out[i] = all[dtype == BFP_DT_BF16 ? 1 : 0][i];
    return BFP_OK;
}

int block_exponent_device(const void* in, float* e_out, int64_t rows, int64_t K, int dtype, int B, float eps, cudaStream_t st) {
    if (dtype != BFP_DT_F32) ensure_exp_tables(st);
    const int64_t units = rows * ((K + B - 1) / B);
    if (units == 0) return BFP_OK;
    const int grid = (int)std::min<int64_t>((units + 127) / 128, (int64_t)device_info().sm_count * 16);
    switch (dtype) {
    case BFP_DT_F32: block_exponent_kernel<BFP_DT_F32><<<grid, 128, 0, st>>>(in, e_out, rows, K, B, eps); break;
    case BFP_DT_F16: block_exponent_kernel<BFP_DT_F16><<<grid, 128, 0, st>>>(in, e_out, rows, K, B, eps); break;
    case BFP_DT_BF16: block_exponent_kernel<BFP_DT_BF16><<<grid, 128, 0, st>>>(in, e_out, rows, K, B, eps); break;
    default: return set_error(BFP_E_ARG, "bad dtype");
    }
    count_launch();
    return check_launch("block_exponent_kernel");
}

int debug_cpu_tie_lut(uint8_t out[256]) {
    cudaError_t e = cudaMemcpyFromSymbol(out, c_cpu_tie_lut, 256);
    if (e != cudaSuccess) return set_error(BFP_E_CUDA, cudaGetErrorString(e));
    return BFP_OK;
}

}  // namespace bfp
