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
        // butterfly over the lanes that share this block; every lane of the warp executes every shuffle
#pragma unroll
        for (int off = 1; off < 32; off <<= 1)
            if (off < p.lanes_per_block) amax = max(amax, __shfl_xor_sync(0xffffffffu, amax, off));
        const BlockScale sc = make_scale<DT>(amax, p.m, p.eps);
        float un[STOC ? V : 1];
        if (STOC) {
#pragma unroll
            for (int q = 0; q < V / 4; ++q) {
                const uint4 r = philox4x32_10((uint64_t)(p.ctr_base + vec_index * (V / 4) + q), p.offset, p.seed);
                un[4 * q] = u01(r.x); un[4 * q + 1] = u01(r.y); un[4 * q + 2] = u01(r.z); un[4 * q + 3] = u01(r.w);
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

// ORDER: BFP_ORDER_*.  M / KD / TIE: see mask_vec.  STOC: stochastic rounding (fp32 output).
template <int DT, int ORDER, int M, int KD, int TIE, bool STOC, bool PADDED>
__global__ void __launch_bounds__(kStreamThreads) quant_stream_kernel(const StreamParams p) {
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
            return u01(w);
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
