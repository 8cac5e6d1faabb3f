// bfp_pack.cu -- quantise (+ N:M sparsify) straight into the packed BFP operand format of the tcgen05 GEMM, and back.
//
// Packed format (DESIGN.md "Packed layout"; new -- the reference only ever materialises dequantised floats):
//   mant    int8  [rows, Kp]           Kp = K rounded up to 16; two's-complement mantissa q, |q| <= 2^m - 1, m <= 7;
//                                      columns >= K are zero.
//   scale_t fp32  [nkb_pad, rows_pad]  TRANSPOSED (block-major): scale_t[kb][row] = 2^(e - m), the block's interval
//                                      (bfp_ops.py:38); value = q * scale.  nkb = ceil(K / B); padding entries are 0.
//                                      Block-major so that the GEMM fetches the 128 (256) row scales of a tile for one
//                                      K-block with a single contiguous bulk copy.
//   A block the fast arithmetic cannot represent (Inf/NaN inside, |x| >= 2^126, all-zero fp16 block: the cases where
//   the reference itself yields NaN or leaves the normal range) gets scale = NaN and zero mantissas, so the GEMM
//   output row becomes NaN exactly where the reference's would.
// Contract: unpack(pack(x)) == float_to_bfp_blocked(x) bit for bit, except that -0.0 unpacks as +0.0.
#include <algorithm>

#include "bfp_internal.h"
#include "bfp_stream.cuh"

namespace bfp {

struct PackParams {
    const uint4* in;
    int8_t* mant;
    float* scale_t;
    int64_t n_vec;
    int64_t Kp;             // mant row stride (== K on the stream path)
    int64_t rows_pad;       // leading dimension of scale_t
    uint32_t nkb;           // blocks per row
    int lanes_per_block, lpb_shift;
    int m;
    float eps;
    int kdrop;
    uint64_t seed, offset;
    uint32_t vec_per_row, slots_per_row;   // padded-row mode (bf16 format only): see StreamParams in bfp_quant.cu; 0 = flat
};

__device__ __forceinline__ int64_t pack_real_vec(const PackParams& p, int64_t g) {
    if (p.slots_per_row == 0u) return g < p.n_vec ? g : -1;
    if (g >= p.n_vec) return -1;
    const uint32_t row = (uint32_t)((uint64_t)g / p.slots_per_row), slot = (uint32_t)((uint64_t)g - (uint64_t)row * p.slots_per_row);
    return slot < p.vec_per_row ? (int64_t)row * p.vec_per_row + slot : -1;
}

__device__ __forceinline__ uint32_t pack4_s8(const float* q) {
    // q are integers in [-127, 127] (or -0.0): cvt.rni.s8 via int conversion, then byte pack
    const int a = __float2int_rn(q[0]), b = __float2int_rn(q[1]), c = __float2int_rn(q[2]), d = __float2int_rn(q[3]);
    return (uint32_t)(a & 0xff) | ((uint32_t)(b & 0xff) << 8) | ((uint32_t)(c & 0xff) << 16) | ((uint32_t)(d & 0xff) << 24);
}

// stream path: same lane mapping as quant_stream_kernel (bfp_quant.cu); requires K % B == 0 so rows never matter for
// the mantissas (Kp == K) and a block's (row, kb) follows from its flat index.
// FMT: 0 = int8 mantissas + block-major fp32 scale table; 1 = dequantised bf16 (exact for m <= 8: q has <= 8 significant
// bits and 2^(e-m) only moves the exponent), the operand format of the exact bf16 tensor-core GEMM.
template <int DT, int ORDER, int M, int KD, bool STOC, int FMT, bool PADDED>
__global__ void __launch_bounds__(kStreamThreads) pack_stream_kernel(const PackParams p) {
    using D = DType<DT>;
    constexpr int V = D::kVec;
    constexpr bool kSparseFirst = ORDER == BFP_ORDER_SPARSIFY_QUANT;
    constexpr bool kSparseLast = ORDER == BFP_ORDER_QUANT_SPARSIFY;
    constexpr int kTileVecs = kStreamThreads * kStreamUnroll;
    const int64_t n_tiles = (p.n_vec + kTileVecs - 1) / kTileVecs;

    pdl_launch_dependents();          // see bfp_stream.cuh: launch overlap only, stream order preserved by pdl_wait
    pdl_wait();

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t tile_base = tile * kTileVecs;
        uint4 raw[kStreamUnroll];
        int64_t rv[kStreamUnroll];                                            // real vector index, -1 = padding / out of range
#pragma unroll
        for (int u = 0; u < kStreamUnroll; ++u) {
            const int64_t g = tile_base + (int)threadIdx.x + u * kStreamThreads;
            rv[u] = PADDED ? pack_real_vec(p, g) : (g < p.n_vec ? g : -1);
            raw[u] = rv[u] >= 0 ? ld_stream(p.in + rv[u]) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < kStreamUnroll; ++u) {
            const int64_t vi = rv[u];
            const bool live = vi >= 0;
            float v[V];
            unpack_vec<DT>(raw[u], v);
            uint32_t amax = 0u;
#pragma unroll
            for (int i = 0; i < V; ++i) amax = max(amax, abs_bits(v[i]));     // unmasked max == masked max (N >= 1)
            if (kSparseFirst) mask_vec<M, KD, BFP_TIE_TORCH_CUDA, V>(v, p.kdrop);
#pragma unroll
            for (int off = 1; off < 32; off <<= 1)
                if (off < p.lanes_per_block) amax = max(amax, __shfl_xor_sync(0xffffffffu, amax, off));
            const BlockScale sc = make_scale<DT>(amax, p.m, p.eps);
            float q[V];
            if (sc.fast) {
                float un[STOC ? V : 1];
                if (STOC) {
#pragma unroll
                    for (int k = 0; k < V / 4; ++k) {
                        const uint4 r = philox4x32_10((uint64_t)(vi * (V / 4) + k), p.offset, p.seed);
                        un[4 * k] = u01(r.x); un[4 * k + 1] = u01(r.y); un[4 * k + 2] = u01(r.z); un[4 * k + 3] = u01(r.w);
                    }
                }
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    const float x = v[i] * sc.inv;
                    const float r = STOC ? rintf((un[i] - 0.5f) + x) : rintf(x);
                    q[i] = fminf(fmaxf(r, -sc.vmax), sc.vmax);                 // integer mantissa
                }
            } else {
#pragma unroll
                for (int i = 0; i < V; ++i) q[i] = 0.0f;
            }
            if (kSparseLast) mask_vec<M, KD, BFP_TIE_TORCH_CUDA, V>(q, p.kdrop);   // same order as masking q * delta
            if (FMT == 1) {
                if (live) {
                    const float d = sc.fast ? sc.delta : __int_as_float(0x7fc00000);
                    uint32_t w[V / 2];
#pragma unroll
                    for (int i = 0; i < V / 2; ++i) {
                        __nv_bfloat162 h = __floats2bfloat162_rn(q[2 * i] * d, q[2 * i + 1] * d);
                        w[i] = *reinterpret_cast<uint32_t*>(&h);
                    }
                    uint16_t* dst = reinterpret_cast<uint16_t*>(p.mant) + vi * V;
                    if (V == 4) *reinterpret_cast<uint2*>(dst) = make_uint2(w[0], w[1]);
                    else *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[V / 2 - 2], w[V / 2 - 1]);
                }
            } else if (live) {
                if (V == 4) {
                    *reinterpret_cast<uint32_t*>(p.mant + vi * 4) = pack4_s8(q);
                } else {
                    *reinterpret_cast<uint2*>(p.mant + vi * 8) = make_uint2(pack4_s8(q), pack4_s8(q + 4));
                }
                if ((threadIdx.x & (p.lanes_per_block - 1)) == 0) {            // block leader: one scale per block
                    const uint32_t gb = (uint32_t)(vi >> p.lpb_shift);
                    const uint32_t row = gb / p.nkb, kb = gb - row * p.nkb;
                    p.scale_t[(int64_t)kb * p.rows_pad + row] = sc.fast ? sc.delta : __int_as_float(0x7fc00000);
                }
            }
        }
    }
}

// generic path: one thread per block; any K (zero-padded tail), any B, N:M groups must nest in blocks for q->s.
struct PackGenericParams {
    const void* in;
    int8_t* mant;
    float* scale_t;
    int64_t rows, K, Kp, rows_pad;
    int B, m;
    float eps;
    int N, M;
    uint64_t seed, offset;
};

template <int DT, int ORDER, bool STOC, int FMT>
__global__ void __launch_bounds__(128) pack_generic_kernel(const PackGenericParams p) {
    const int64_t nkb = (p.K + p.B - 1) / p.B;
    for (int64_t unit = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; unit < p.rows * nkb; unit += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = unit / nkb, kb = unit % nkb;
        auto raw = [&](int64_t c) { return ld_pad<DT>(p.in, row, c, p.K); };
        auto src = [&](int64_t c) {
            const float t = raw(c);
            if (ORDER == BFP_ORDER_SPARSIFY_QUANT) return (c < p.K && nm_dropped(raw, c, p.N, p.M, BFP_TIE_TORCH_CUDA)) ? 0.0f : t;
            return t;
        };
        auto uni = [&](int64_t c) {
            const uint64_t flat = (uint64_t)(row * p.K + c);
            const uint4 r = philox4x32_10(flat >> 2, p.offset, p.seed);
            const uint32_t w = (flat & 3) == 0 ? r.x : ((flat & 3) == 1 ? r.y : ((flat & 3) == 2 ? r.z : r.w));
            return u01(w);
        };
        const int64_t c0 = kb * p.B, c1 = min(p.K, c0 + p.B);
        uint32_t amax = 0u;
        for (int64_t c = c0; c < c1; ++c) amax = max(amax, abs_bits(src(c)));
        const BlockScale sc = make_scale<DT>(amax, p.m, p.eps);
        const float dq = sc.fast ? sc.delta : __int_as_float(0x7fc00000);
        if (FMT == 0) p.scale_t[kb * p.rows_pad + row] = dq;
        auto put = [&](int64_t c, float qv) {
            if (FMT == 0) p.mant[row * p.Kp + c] = (int8_t)__float2int_rn(qv);
            else reinterpret_cast<__nv_bfloat16*>(p.mant)[row * p.Kp + c] = __float2bfloat16_rn(qv * dq);
        };
        auto qof = [&](int64_t c) {
            if (!sc.fast || c >= c1) return 0.0f;
            const float x = src(c) * sc.inv;
            const float r = STOC ? rintf((uni(c) - 0.5f) + x) : rintf(x);
            return fminf(fmaxf(r, -sc.vmax), sc.vmax);
        };
        if (ORDER == BFP_ORDER_QUANT_SPARSIFY) {
            float q[kMaxGroup];
            for (int64_t g0 = c0; g0 < c1; g0 += p.M) {          // groups nest in blocks (host checks B % M == 0)
                for (int j = 0; j < p.M; ++j) q[j] = qof(g0 + j);
                auto qsrc = [&](int64_t c) { return q[c - g0]; };
                for (int j = 0; j < p.M && g0 + j < c1; ++j)
                    put(g0 + j, nm_dropped(qsrc, g0 + j, p.N, p.M, BFP_TIE_TORCH_CUDA) ? 0.0f : q[j]);
            }
        } else {
            for (int64_t c = c0; c < c1; ++c) put(c, qof(c));
        }
    }
}

__global__ void __launch_bounds__(256) unpack_kernel(const int8_t* mant, const float* scale_t, float* out, int64_t rows, int64_t K,
                                                     int64_t Kp, int64_t rows_pad, int B) {
    const int64_t n = rows * K;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / K, c = i - row * K;
        out[i] = (float)mant[row * Kp + c] * scale_t[(c / B) * rows_pad + row];
    }
}

// ---------------------------------------------------------------------------------------------------------------
template <int DT, int ORDER, bool STOC, int FMT>
static int launch_pack_stream(const PackParams& p, bool sparse, cudaStream_t st) {
    const int64_t n_tiles = (p.n_vec + kStreamThreads * kStreamUnroll - 1) / (kStreamThreads * kStreamUnroll);
    if (p.slots_per_row != 0) {                           // padded rows (bf16 format only)
        if constexpr (FMT == 1) {
            static const int occ_ps = kernel_occupancy(pack_stream_kernel<DT, ORDER, 4, 2, STOC, FMT, true>, kStreamThreads);
            static const int occ_pd = kernel_occupancy(pack_stream_kernel<DT, BFP_ORDER_QUANT_ONLY, 0, 0, STOC, FMT, true>, kStreamThreads);
            const int gridp = stream_grid(sparse ? occ_ps : occ_pd, n_tiles);
            if (int rc = sparse ? launch_pdl(pack_stream_kernel<DT, ORDER, 4, 2, STOC, FMT, true>, gridp, kStreamThreads, st, p)
                                : launch_pdl(pack_stream_kernel<DT, BFP_ORDER_QUANT_ONLY, 0, 0, STOC, FMT, true>, gridp, kStreamThreads, st, p))
                return rc;
            count_launch();
            return check_launch("pack_stream_kernel (padded rows)");
        } else {
            return set_error(BFP_E_UNSUPPORTED, "internal: padded rows are implemented for the bf16 operand format only");
        }
    }
    static const int occ_s = kernel_occupancy(pack_stream_kernel<DT, ORDER, 4, 2, STOC, FMT, false>, kStreamThreads);
    static const int occ_d = kernel_occupancy(pack_stream_kernel<DT, BFP_ORDER_QUANT_ONLY, 0, 0, STOC, FMT, false>, kStreamThreads);
    const int grid = stream_grid(sparse ? occ_s : occ_d, n_tiles);
    if (int rc = sparse ? launch_pdl(pack_stream_kernel<DT, ORDER, 4, 2, STOC, FMT, false>, grid, kStreamThreads, st, p)
                        : launch_pdl(pack_stream_kernel<DT, BFP_ORDER_QUANT_ONLY, 0, 0, STOC, FMT, false>, grid, kStreamThreads, st, p))
        return rc;
    count_launch();
    return check_launch("pack_stream_kernel");
}

template <int DT, bool STOC, int FMT>
static int pack_dt(const QuantArgs& a, int8_t* mant, float* scale_t, int64_t Kp, int64_t rows_pad, cudaStream_t st) {
    constexpr int V = DType<DT>::kVec;
    const bool sparse = a.order != BFP_ORDER_QUANT_ONLY;
    const int64_t numel = a.rows * a.K, nkb = (a.K + a.B - 1) / a.B;
    if (numel == 0) return BFP_OK;
    bool fast = (a.K % a.B == 0) && (a.B & (a.B - 1)) == 0 && a.B >= V && a.B <= 32 * V && Kp == a.K &&
                reinterpret_cast<uintptr_t>(a.in) % 16 == 0 && reinterpret_cast<uintptr_t>(mant) % 16 == 0 &&
                a.rows * nkb < (int64_t)1 << 32 && !tuning().force_generic;
    if (sparse) fast = fast && a.M == 4 && a.N == 2;
    // padded-row mode for the bf16 operand format: vector-aligned rows whose length is not a multiple of the block
    bool padded = false;
    if (!fast && FMT == 1 && !tuning().force_generic) {
        padded = (a.K % V == 0) && (a.K % a.B != 0) && (a.B & (a.B - 1)) == 0 && a.B >= V && a.B <= 32 * V && Kp == a.K &&
                 reinterpret_cast<uintptr_t>(a.in) % 16 == 0 && reinterpret_cast<uintptr_t>(mant) % 16 == 0 && a.K / V < (int64_t)1 << 31 &&
                 a.rows < (int64_t)1 << 31 && (!sparse || (a.M == 4 && a.N == 2 && a.K % 4 == 0));
    }
    if (fast || padded) {
        PackParams p;
        p.in = static_cast<const uint4*>(a.in); p.mant = mant; p.scale_t = scale_t; p.n_vec = numel / V; p.Kp = Kp;
        p.vec_per_row = p.slots_per_row = 0;
        if (padded) {
            p.vec_per_row = (uint32_t)(a.K / V);
            p.slots_per_row = (uint32_t)(round_up(a.K, a.B) / V);
            p.n_vec = a.rows * (int64_t)p.slots_per_row;
        }
        p.rows_pad = rows_pad; p.nkb = (uint32_t)nkb; p.lanes_per_block = a.B / V; p.lpb_shift = __builtin_ctz(a.B / V);
        p.m = a.m; p.eps = a.eps; p.kdrop = sparse ? a.M - a.N : 0; p.seed = a.seed; p.offset = a.offset;
        switch (a.order) {
        case BFP_ORDER_QUANT_ONLY: return launch_pack_stream<DT, BFP_ORDER_QUANT_ONLY, STOC, FMT>(p, false, st);
        case BFP_ORDER_SPARSIFY_QUANT: return launch_pack_stream<DT, BFP_ORDER_SPARSIFY_QUANT, STOC, FMT>(p, true, st);
        case BFP_ORDER_QUANT_SPARSIFY: return launch_pack_stream<DT, BFP_ORDER_QUANT_SPARSIFY, STOC, FMT>(p, true, st);
        }
        return set_error(BFP_E_ARG, "bad order");
    }
    if (a.order == BFP_ORDER_QUANT_SPARSIFY && a.B % a.M != 0)
        return set_error(BFP_E_UNSUPPORTED, "packed quantise->sparsify needs N:M groups that nest in blocks (B % M == 0)");
    PackGenericParams g;
    g.in = a.in; g.mant = mant; g.scale_t = scale_t; g.rows = a.rows; g.K = a.K; g.Kp = Kp; g.rows_pad = rows_pad;
    g.B = a.B; g.m = a.m; g.eps = a.eps; g.N = a.N; g.M = sparse ? a.M : 1; g.seed = a.seed; g.offset = a.offset;
    const int64_t units = a.rows * nkb;
    const int grid = (int)std::min<int64_t>((units + 127) / 128, (int64_t)device_info().sm_count * 16);
    switch (a.order) {
    case BFP_ORDER_QUANT_ONLY: pack_generic_kernel<DT, BFP_ORDER_QUANT_ONLY, STOC, FMT><<<grid, 128, 0, st>>>(g); break;
    case BFP_ORDER_SPARSIFY_QUANT: pack_generic_kernel<DT, BFP_ORDER_SPARSIFY_QUANT, STOC, FMT><<<grid, 128, 0, st>>>(g); break;
    case BFP_ORDER_QUANT_SPARSIFY: pack_generic_kernel<DT, BFP_ORDER_QUANT_SPARSIFY, STOC, FMT><<<grid, 128, 0, st>>>(g); break;
    default: return set_error(BFP_E_ARG, "bad order");
    }
    count_launch();
    return check_launch("pack_generic_kernel");
}

template <int FMT>
static int pack_fmt(const QuantArgs& a, int8_t* mant, float* scale_t, int64_t Kp, int64_t rows_pad, cudaStream_t st) {
    const bool stoc = a.rounding == BFP_ROUND_STOCHASTIC;
    switch (a.in_dtype) {
    case BFP_DT_F32: return stoc ? pack_dt<BFP_DT_F32, true, FMT>(a, mant, scale_t, Kp, rows_pad, st) : pack_dt<BFP_DT_F32, false, FMT>(a, mant, scale_t, Kp, rows_pad, st);
    case BFP_DT_F16: return stoc ? pack_dt<BFP_DT_F16, true, FMT>(a, mant, scale_t, Kp, rows_pad, st) : pack_dt<BFP_DT_F16, false, FMT>(a, mant, scale_t, Kp, rows_pad, st);
    case BFP_DT_BF16: return stoc ? pack_dt<BFP_DT_BF16, true, FMT>(a, mant, scale_t, Kp, rows_pad, st) : pack_dt<BFP_DT_BF16, false, FMT>(a, mant, scale_t, Kp, rows_pad, st);
    }
    return set_error(BFP_E_ARG, "bad dtype");
}

int pack_device(const QuantArgs& a, int8_t* mant, float* scale_t, int64_t Kp, int64_t rows_pad, cudaStream_t st) {
    if (a.in_dtype != BFP_DT_F32) ensure_exp_tables(st);
    return pack_fmt<0>(a, mant, scale_t, Kp, rows_pad, st);
}
// dequantised bf16 [rows, Kp] (Kp = K rounded up to 8 elements)
int pack_bf16_device(const QuantArgs& a, void* out_bf16, int64_t Kp, cudaStream_t st) {
    if (a.in_dtype != BFP_DT_F32) ensure_exp_tables(st);
    return pack_fmt<1>(a, static_cast<int8_t*>(out_bf16), nullptr, Kp, 0, st);
}

int unpack_device(const int8_t* mant, const float* scale_t, float* out, int64_t rows, int64_t K, int64_t Kp, int64_t rows_pad, int B,
                  cudaStream_t st) {
    if (rows * K == 0) return BFP_OK;
    const int grid = (int)std::min<int64_t>((rows * K + 255) / 256, (int64_t)device_info().sm_count * 16);
    unpack_kernel<<<grid, 256, 0, st>>>(mant, scale_t, out, rows, K, Kp, rows_pad, B);
    count_launch();
    return check_launch("unpack_kernel");
}

// ---------------------------------------------------------------------------------------------------------------
// 16-bit transpose with zero padding: in [R, C] (row stride ld_in) -> out [C, ld_out], out[c][r] = in[r][c], columns R .. ld_out-1
// zero.  The backward contractions of the BFP linear run over T (wgrad) and N (dgrad): their operands are the transposes
// of the packed bf16 tensors (bfp_ops._BFPLinearTC.backward).  64 x 64 tiles through shared memory, 128-byte rows both ways.
// ---------------------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) transpose16_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, int64_t R, int64_t C,
                                                          int64_t ld_in, int64_t ld_out) {
    __shared__ uint16_t tile[64][66];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t r0 = (int64_t)blockIdx.y * 64, c0 = (int64_t)blockIdx.x * 64;
    const bool vec_in = (ld_in % 2 == 0) && (reinterpret_cast<uintptr_t>(in) % 4 == 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = ty + 8 * i;
        const int64_t gr = r0 + r, gc = c0 + 2 * tx;
        uint32_t v = 0;
        if (gr < R) {
            if (vec_in && gc + 1 < C) v = *reinterpret_cast<const uint32_t*>(in + gr * ld_in + gc);
            else {
                if (gc < C) v = in[gr * ld_in + gc];
                if (gc + 1 < C) v |= (uint32_t)in[gr * ld_in + gc + 1] << 16;
            }
        }
        tile[r][2 * tx] = (uint16_t)(v & 0xffffu);
        tile[r][2 * tx + 1] = (uint16_t)(v >> 16);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = ty + 8 * i;
        const int64_t gc = c0 + c, gr = r0 + 2 * tx;                  // out row gc, out columns gr, gr + 1 (ld_out is even)
        if (gc < C && gr < ld_out) {
            const uint32_t v = (uint32_t)tile[2 * tx][c] | ((uint32_t)tile[2 * tx + 1][c] << 16);   // rows >= R were loaded as zero
            *reinterpret_cast<uint32_t*>(out + gc * ld_out + gr) = v;
        }
    }
}
}  // namespace

int transpose16_device(const void* in, void* out, int64_t R, int64_t C, int64_t ld_in, int64_t ld_out, cudaStream_t s) {
    if (R == 0 || C == 0) return BFP_OK;
    if (ld_out % 2 || ld_out < R || ld_in < C) return set_error(BFP_E_ARG, "transpose: ld_out must be even and >= R, ld_in >= C");
    if (reinterpret_cast<uintptr_t>(out) % 4) return set_error(BFP_E_ALIGN, "transpose: out must be 4-byte aligned");
    const int64_t gx = (C + 63) / 64, gy = (ld_out + 63) / 64;
    if (gy > 65535) return set_error(BFP_E_ARG, "transpose: too many rows");
    transpose16_kernel<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, s>>>(static_cast<const uint16_t*>(in), static_cast<uint16_t*>(out), R, C, ld_in, ld_out);
    count_launch();
    return check_launch("transpose16_kernel");
}

}  // namespace bfp
