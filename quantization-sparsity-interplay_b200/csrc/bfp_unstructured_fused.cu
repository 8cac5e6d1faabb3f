// bfp_unstructured_fused.cu -- global magnitude pruning fused with the BFP quantiser: two reads and one write of the tensor.
//
// Replaces float_to_bfp_blocked (bfp_ops.py:124-149) with sparsity_mode == 'unstructured' -- _unstructured_sparsity (:61-71) and
// _no_sparsity_float_to_bfp (:46-59) in either order -- and _unstructured_sparsity alone.  The k = int(numel * frac) entries
// torch.topk(|t|, k, largest=False) returns on torch-CUDA are dropped: every key strictly below the k-th smallest key tau, and of
// the keys equal to tau the first (k - #below) in index order.  key = bit pattern of |value| as fp32 (monotone for fp16 / bf16
// values too; all NaNs share one key above +inf, as in torch's radix select).  With first == 'q' the keys are those of the
// QUANTISED values, recomputed on the fly in every pass.
//
// The multi-pass radix select of bfp_unstructured.cu reads the tensor four times (plus a separate quantiser pass).  Here:
//   1. sample_kernel   one cluster of 8 CTAs.  A stratified sample of ~64 K elements goes into a histogram of the 16 leading key
//                      bits spread over the cluster's shared memory; the bins holding the sample quantiles k/n -+ 5.5 sigma
//                      BRACKET tau (about two per cent of the mass).
//   2. pass_a_kernel   first full read; CTA b owns the contiguous RANGE b of the tensor.  Counts the keys below the bracket, and
//                      looks closer at the keys inside it:
//                        window mode (bracket <= 2048 bins): per-bin counts, kept per range for the keys whose 15 trailing bits
//                          are zero ("round" keys: 0.0, every BFP value with mant_bits <= 8, bf16 data -- where the massive ties
//                          are); the other in-bracket keys are appended to a candidate list as (key, range);
//                        list mode (a wider bracket, e.g. one that straddles the zeros of a ReLU output and its smallest positive
//                          values): every in-bracket key is listed, except one "hot" round key (the zeros) counted per range.
//                      The in-bracket keys of a tile go through per-lane queues, so the per-element code is branch-free.
//   3. refine_kernel   (cooperative launch) window mode: the candidates of the bin that holds the k-th key resolve its 15 trailing
//                      bits; list mode: a three-digit radix select over the list.  Either way -> tau, how many of the keys equal
//                      to tau go (`need`) and how many there are, and -- when only some go -- how many of them each range holds
//                      (from the per-range counts for a round tau, from the list otherwise).  If the bracket missed (probability
//                      ~1e-7) or the list overflowed (massive ties on a non-round value) the same grid runs the radix select over
//                      the whole tensor, with grid-wide barriers -- no host round trip.
//   4. apply_kernel    second read + the write; CTA b owns range b again.  Mask (key < tau; key == tau according to the range:
//                      all of them, none, or -- in the ONE range where the cut falls -- by a running tie rank), BFP quantisation
//                      before or after it, 128-bit stores.  No communication between CTAs.
// = 2 reads + 1 write (12 B / element fp32) for 8 B / element algorithmic.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "bfp_internal.h"
#include "bfp_stream.cuh"

namespace bfp {

namespace {
namespace cg = cooperative_groups;

constexpr int kT = 256;                    // threads per CTA of the full-tensor kernels
constexpr int kWarps = kT / 32;
constexpr int kWin = 2048;                 // window bins (16 leading key bits per bin)
constexpr int kLowBins = 32768;            // the 15 trailing key bits
constexpr int kSampleThreads = 1024;
constexpr int kSampleCtas = 8;              // the sample kernel is one cluster of eight CTAs
constexpr int kSampleVecs = 16384;         // 128-bit vectors in the sample (64 K fp32 / 128 K half elements)
constexpr int kWarpBuf = 256;              // listed keys a warp stages in shared memory between flushes
constexpr int kQ = 16;                     // elements per thread per tile = slots of a lane's in-bracket queue
constexpr int kFbBins = 2048;              // radix digits of the list / whole-tensor select: 11 + 10 + 10 bits
constexpr int kMaxRanges = 2048;
constexpr uint32_t kNoKey = 0xffffffffu;

enum { MODE_S_ONLY = 0, MODE_SQ = 1, MODE_QS = 2 };
enum { PATH_TENSOR = 0, PATH_WINDOW = 1, PATH_LIST = 2 };       // FusedState::path
enum { TIES_ALL = 0, TIES_ROW = 1, TIES_COUNTED = 2 };           // FusedState::tie_src

struct FusedState {
    uint32_t lo_bin, span;                 // bracket = bins [lo_bin, lo_bin + span] of the 16 leading key bits
    uint32_t hot_key;                      // a round key holding > 1/64 of the sample: counted in registers, never queued (kNoKey: none)
    uint32_t path;                         // PATH_*; pass A / refine downgrade it to PATH_TENSOR when the two-read path does not hold
    unsigned int done_a, done_r;
    uint32_t flag_r, pad0;                 // refine: set by the CTA that finished the selection
    unsigned long long below;              // keys below the bracket
    unsigned long long cand_count;         // listed keys (may exceed the capacity: then path = PATH_TENSOR)
    unsigned long long hot_total;          // list mode: keys equal to hot_key
    uint32_t bin, pad1;                    // window mode: the bin holding the k-th smallest key ...
    unsigned long long need_bin, cnt_bin;  // ... its rank inside the bin (1-based), keys in the bin
    unsigned long long impure_in_bin;      // listed keys found in the bin
    uint32_t tau;
    uint32_t tie_src, tie_row;             // where apply_kernel finds the per-range tie counts (TIES_*)
    uint32_t pad2;
    unsigned long long need, ties_total;   // of the ties_total keys equal to tau the first `need` (index order) are dropped
    uint32_t prefix_value, prefix_mask;    // radix select state
    unsigned long long fb_need;
    long long clk[8];                      // sample kernel phase timestamps (BFP_UNSTRUCTURED_TIMING)
    alignas(16) uint32_t win_hist[kWin];   // (the histograms are read with 128-bit loads)
    uint32_t fb_hist[3][kFbBins];
    uint32_t range_count[kMaxRanges];
    uint32_t low_hist[kLowBins];
};
static_assert(offsetof(FusedState, win_hist) % 16 == 0 && offsetof(FusedState, fb_hist) % 16 == 0 && offsetof(FusedState, low_hist) % 16 == 0 &&
              offsetof(FusedState, range_count) % 16 == 0, "histograms are read with 128-bit loads");
constexpr size_t kStateHeaderWords = offsetof(FusedState, fb_hist) / 4;    // zeroed by the sample kernel (the rest by pass A)

struct UParams {
    const uint4* in;
    uint4* out;
    int64_t n_vec, n_tiles;
    unsigned long long k, n;
    int lanes_per_block, m;
    float eps;
    uint64_t seed, offset;
    FusedState* st;
    uint32_t* cta_hist;                    // [row][range]: window mode row d = round keys of bin lo_bin + d; list mode row 0 = hot key
    uint2* cand;                           // (key, range)
    unsigned long long cand_cap;
    int n_ranges, tiles_per_range;
    int force_fallback;
};

__device__ __forceinline__ unsigned long long ld_cg64(const unsigned long long* p) { return __ldcg(p); }
__device__ __forceinline__ uint32_t ld_cg32(const uint32_t* p) { return __ldcg(p); }

// BFP quantisation of one 128-bit vector in place (the arithmetic and Philox counters of process_vec in bfp_quant.cu).  Every lane
// of the warp must call it: the block maximum is a butterfly over the lanes that share the block.
template <int DT, bool STOC>
__device__ __forceinline__ void quantize_vec(float* v, const UParams& p, int64_t vec_index) {
    using D = DType<DT>;
    constexpr int V = D::kVec;
    uint32_t amax = 0u;
#pragma unroll
    for (int i = 0; i < V; ++i) amax = max(amax, abs_bits(v[i]));
#pragma unroll
    for (int off = 1; off < 32; off <<= 1)
        if (off < p.lanes_per_block) amax = max(amax, __shfl_xor_sync(0xffffffffu, amax, off));
    const BlockScale sc = make_scale<DT>(amax, p.m, p.eps);
    float un[STOC ? V : 1];
    if (STOC) {
#pragma unroll
        for (int q = 0; q < V / 4; ++q) {
            const uint4 r = philox4x32_10((uint64_t)(vec_index * (V / 4) + q), p.offset, p.seed);
            un[4 * q] = u01_centered(r.x); un[4 * q + 1] = u01_centered(r.y); un[4 * q + 2] = u01_centered(r.z); un[4 * q + 3] = u01_centered(r.w);
        }
    }
    if (sc.fast) {
#pragma unroll
        for (int i = 0; i < V; ++i) v[i] = quant_elt_fast<STOC>(v[i], sc, STOC ? un[i] : 0.0f);
    } else {
#pragma unroll
        for (int i = 0; i < V; ++i) v[i] = quant_elt_slow<DT, STOC>(v[i], sc.delta, sc.vmax, STOC ? un[i] : 0.0f);
    }
}

// exclusive prefix of `mine` over the CTA (T threads, T / 32 <= 32 warps) and the CTA total; `scratch` holds 32 words
template <int T>
__device__ __forceinline__ unsigned long long block_excl_scan(unsigned long long mine, unsigned long long* scratch, unsigned long long* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long incl = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned long long o = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += o;
    }
    __syncthreads();                        // scratch may still be read from an earlier call
    if (lane == 31) scratch[warp] = incl;
    __syncthreads();
    unsigned long long before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < T / 32; ++w) { const unsigned long long c = scratch[w]; before += w < warp ? c : 0ull; all += c; }
    *total = all;
    return before + incl - mine;
}

template <int T>
__device__ __forceinline__ unsigned long long block_sum(unsigned long long mine, unsigned long long* scratch) {
    unsigned long long total;
    block_excl_scan<T>(mine, scratch, &total);
    return total;
}

// The bin holding the need-th smallest key (1-based) of a 32-bit histogram in global memory, by the whole CTA.  T threads,
// T * PER bins: chunk c = bins [c PER, (c + 1) PER).  Chunk sums are formed with coalesced loads, a block scan over the T chunk
// sums finds the chunk, one warp walks its PER bins: nothing is chained.  `extra_cnt` is added to bin `extra_bin`.
// found = 0 when need is outside [1, total].
struct SelectResult { uint32_t bin; int found; unsigned long long before, count; };
template <int T, int PER>
__device__ __forceinline__ SelectResult block_select(const uint32_t* hist, int nbins, unsigned long long need, uint32_t extra_bin, unsigned long long extra_cnt,
                                                      unsigned long long* scratch, SelectResult* s_res) {
    static_assert(PER == 8 || PER == 128, "chunk = 8 bins (one thread) or 128 bins (one 128-bit load per lane)");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ unsigned long long s_chunk[T];
    unsigned long long c8[PER == 8 ? 8 : 1];
    // four consecutive bins starting at b (a multiple of 4), as 64-bit counts with the extra key mixed in
    auto bins4 = [&](int b, unsigned long long* c) {
        const uint4 w = b < nbins ? __ldcg(reinterpret_cast<const uint4*>(hist + b)) : make_uint4(0u, 0u, 0u, 0u);
        c[0] = w.x; c[1] = w.y; c[2] = w.z; c[3] = w.w;
        if ((extra_bin & ~3u) == (uint32_t)b) {
#pragma unroll
            for (int j = 0; j < 4; ++j) c[j] += (extra_bin & 3u) == (uint32_t)j ? extra_cnt : 0ull;
        }
    };
    if (PER == 8) {
        unsigned long long mine = 0, c[4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            bins4((int)threadIdx.x * 8 + 4 * h, c);
#pragma unroll
            for (int j = 0; j < 4; ++j) { c8[PER == 8 ? 4 * h + j : 0] = c[j]; mine += c[j]; }
        }
        s_chunk[threadIdx.x] = mine;
    } else {
        // warp w sums chunks [32 w, 32 w + 32): one 128-bit load per lane and chunk, eight chunks in flight
        for (int c0 = 0; c0 < 32; c0 += 8) {
            unsigned long long part[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                unsigned long long c[4];
                bins4((warp * 32 + c0 + q) * 128 + lane * 4, c);
                part[q] = c[0] + c[1] + c[2] + c[3];
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) part[q] += __shfl_xor_sync(0xffffffffu, part[q], off);
                if (lane == 0) s_chunk[warp * 32 + c0 + q] = part[q];
            }
        }
    }
    if (threadIdx.x == 0) { s_res->found = 0; s_res->bin = 0; s_res->before = 0; s_res->count = 0; }
    __syncthreads();
    const unsigned long long mine = s_chunk[threadIdx.x];
    unsigned long long total;
    unsigned long long before = block_excl_scan<T>(mine, scratch, &total);
    const bool owner = need >= 1 && need <= total && before < need && need <= before + mine;      // exactly one thread (chunk)
    if (PER == 8) {
        if (owner) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const unsigned long long c = c8[PER == 8 ? j : 0];
                if (before < need && need <= before + c) { s_res->bin = threadIdx.x * 8 + j; s_res->before = before; s_res->count = c; s_res->found = 1; }
                before += c;
            }
        }
        __syncthreads();
    } else {
        __shared__ int s_chunk_sel;
        __shared__ unsigned long long s_before;
        if (threadIdx.x == 0) s_chunk_sel = -1;
        __syncthreads();
        if (owner) { s_chunk_sel = (int)threadIdx.x; s_before = before; }
        __syncthreads();
        if (warp == 0 && s_chunk_sel >= 0) {
            unsigned long long c[4];
            bins4(s_chunk_sel * 128 + lane * 4, c);
            const unsigned long long sum4 = c[0] + c[1] + c[2] + c[3];
            unsigned long long incl = sum4;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned long long o = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= off) incl += o;
            }
            unsigned long long bef = s_before + incl - sum4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (bef < need && need <= bef + c[j]) { s_res->bin = (uint32_t)(s_chunk_sel * 128 + lane * 4 + j); s_res->before = bef; s_res->count = c[j]; s_res->found = 1; }
                bef += c[j];
            }
        }
        __syncthreads();
    }
    return *s_res;
}

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// ---------------------------------------------------------------------------------------------------------------
// 1. sample: one cluster of eight CTAs.  Every CTA histograms its share of the sample into a private 65536-bin table of 16-bit
//    counters in its own shared memory; CTA r then owns bins [8192 r, 8192 r + 8192) and sums the eight tables' slices through
//    distributed shared memory (remote atomics onto the one CTA that owns the populated exponents would serialise).
// ---------------------------------------------------------------------------------------------------------------
template <int DT, int MODE, bool STOC>
__global__ void __cluster_dims__(kSampleCtas, 1, 1) __launch_bounds__(kSampleThreads) sample_kernel(const UParams p) {
    pdl_launch_dependents();
    pdl_wait();
    using D = DType<DT>;
    constexpr int V = D::kVec;
    constexpr int kCoarse = 1024, kFinePer = 65536 / kCoarse;      // coarse bin = 64 fine bins; thread t of a CTA owns coarse bin t
    static_assert(kCoarse == kSampleThreads, "one coarse bin per thread");
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned int rank = cluster.block_rank();
    extern __shared__ unsigned int s_h[];                    // 65536 16-bit counters, two per word: this CTA's share of the sample
    __shared__ unsigned int s_coarse[kCoarse];               // this CTA's coarse histogram (read by CTA 0 through distributed shared memory)
    __shared__ unsigned long long s_scr[32];
    __shared__ unsigned int s_fine[2][kFinePer];             // CTA 0: merged fine counts of the two coarse bins that hold the rank bracket
    __shared__ uint32_t s_cb[2];                             // CTA 0: those coarse bins ...
    __shared__ unsigned long long s_cb_before[2];            // ... and the sample keys before them
    __shared__ unsigned long long s_best;
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < 32768; i += kSampleThreads) s_h[i] = 0u;
    if (rank == 0) {
        uint32_t* w = reinterpret_cast<uint32_t*>(p.st);
        for (int i = tid; i < (int)kStateHeaderWords; i += kSampleThreads) w[i] = 0u;
    }
    if (tid == 0) { s_cb[0] = 0u; s_cb[1] = kCoarse - 1; s_cb_before[0] = 0ull; s_cb_before[1] = 0ull; s_best = 0ull; }
    __syncthreads();
    long long clk[8];
    auto stamp = [&](int i) { if (rank == 0 && tid == 0) clk[i] = clock64(); };
    stamp(0);

    // stratified sample: S_u units (a unit = one vector, or one BFP block when the keys are those of quantised values), unit j
    // taken at a hashed position inside the j-th of S_u equal strata
    const uint32_t L = MODE == MODE_QS ? (uint32_t)p.lanes_per_block : 1u, l_shift = 31 - __clz(L);     // L is a power of two
    const uint32_t n_units = (uint32_t)(p.n_vec >> l_shift);       // n < 2^32 elements
    const uint32_t S_u = min(n_units, (uint32_t)kSampleVecs >> l_shift);
    const uint32_t stride = n_units / S_u;                   // >= 1
    const int S_v = (int)(S_u << l_shift);
    for (int j0 = (int)(rank * kSampleThreads) + (tid & ~31); j0 < S_v; j0 += kSampleCtas * kSampleThreads) {
        const int j = j0 + lane;
        const bool active = j < S_v;
        int64_t vec = 0;
        if (active) {
            const uint32_t unit = (uint32_t)j >> l_shift;
            const uint32_t u = unit * stride + hash32(unit) % stride;
            vec = ((int64_t)u << l_shift) + ((uint32_t)j & (L - 1u));
        }
        float v[V];
        const uint4 raw = active ? ld_stream(p.in + vec) : make_uint4(0u, 0u, 0u, 0u);
        unpack_vec<DT>(raw, v);
        if (MODE == MODE_QS) quantize_vec<DT, STOC>(v, p, vec);
#pragma unroll
        for (int e = 0; e < V; ++e) {
            // one shared-memory atomic per distinct bin of the warp (massive ties -- zeros, quantised values -- would otherwise serialise)
            const uint32_t bin = active ? (topk_key(v[e]) >> 15) : 0xffffffffu;
            const uint32_t same = __match_any_sync(0xffffffffu, bin);
            if (active && lane == __ffs(same) - 1) atomicAdd(&s_h[bin >> 1], (unsigned int)__popc(same) << (16 * (bin & 1u)));
        }
    }
    __syncthreads();
    stamp(1);
    {   // this CTA's coarse histogram: thread t sums its 64 fine counters (32 words, rotated start: bank = lane)
        unsigned int c = 0;
        for (int i = 0; i < 32; ++i) { const unsigned int w = s_h[tid * 32 + ((i + lane) & 31)]; c += (w & 0xffffu) + (w >> 16); }
        s_coarse[tid] = c;
    }
    cluster.sync();
    stamp(2);
    if (rank == 0) {
        // merged coarse histogram (one remote word per CTA and thread), its scan, and the coarse bins of the two bracket ranks
        unsigned long long mine = 0;
#pragma unroll
        for (int r = 0; r < kSampleCtas; ++r) mine += cluster.map_shared_rank(s_coarse, r)[tid];
        unsigned long long total;
        const unsigned long long before = block_excl_scan<kSampleThreads>(mine, s_scr, &total);
        const unsigned long long S = (unsigned long long)S_v * V;     // == total
        long long r_lo, r_hi;
        if ((unsigned long long)S_v == (unsigned long long)p.n_vec) {
            r_lo = r_hi = (long long)p.k - 1;                // the sample is the tensor
        } else {
            const float q = (float)p.k / (float)p.n;
            const float mean = q * (float)S;
            const float sd = sqrtf((float)S * q * (1.0f - q) * (MODE == MODE_QS ? 4.0f : 1.0f));   // design effect: a block shares its scale
            r_lo = (long long)floorf(mean - 5.5f * sd) - 2;
            r_hi = (long long)ceilf(mean + 5.5f * sd) + 2;
        }
        const bool lo_open = r_lo < 0, hi_open = r_hi >= (long long)total;     // the bracket reaches the bottom / the top of the key range
        for (int which = 0; which < 2; ++which) {
            const long long r = which ? r_hi : r_lo;
            if (r >= 0 && (unsigned long long)r < total && before <= (unsigned long long)r && (unsigned long long)r < before + mine) {
                s_cb[which] = (uint32_t)tid; s_cb_before[which] = before;
            }
        }
        __syncthreads();
        // merged fine counts of those two coarse bins
        if (tid < 2 * kFinePer) {
            const int which = tid / kFinePer, f = tid % kFinePer;
            const uint32_t b = s_cb[which] * kFinePer + f;
            unsigned int c = 0;
#pragma unroll
            for (int r = 0; r < kSampleCtas; ++r) { const unsigned int w = cluster.map_shared_rank(s_h, r)[b >> 1]; c += (b & 1u) ? (w >> 16) : (w & 0xffffu); }
            s_fine[which][f] = c;
        }
        __syncthreads();
        stamp(3);
        // hot bin: a fine bin holding more than 1/64 of the sample can only live in a coarse bin that does; those (at most 64) are merged too
        const bool heavy = mine * 64ull > S;
        __shared__ int s_nheavy, s_heavy_list[64];
        if (tid == 0) s_nheavy = 0;
        __syncthreads();
        if (heavy) { const int i = atomicAdd(&s_nheavy, 1); if (i < 64) s_heavy_list[i] = tid; }
        __syncthreads();
        const int nheavy = min(s_nheavy, 64);
        for (int i0 = 0; i0 < nheavy; i0 += kSampleThreads / kFinePer) {       // sixteen heavy coarse bins per round, 64 threads each
            const int i = i0 + tid / kFinePer;
            if (i < nheavy) {
                const uint32_t b = (uint32_t)s_heavy_list[i] * kFinePer + (uint32_t)(tid % kFinePer);
                unsigned int c = 0;
#pragma unroll
                for (int r = 0; r < kSampleCtas; ++r) { const unsigned int w = cluster.map_shared_rank(s_h, r)[b >> 1]; c += (b & 1u) ? (w >> 16) : (w & 0xffffu); }
                if ((unsigned long long)c * 64ull > S) atomicMax(&s_best, ((unsigned long long)c << 32) | b);
            }
        }
        __syncthreads();
        stamp(4);
        if (tid == 0) {
            // the fine bins of the two ranks
            uint32_t lo = 0u, hi = 65535u;
            if (!lo_open) {
                unsigned long long acc = s_cb_before[0];
                for (int f = 0; f < kFinePer; ++f) { if ((unsigned long long)r_lo < acc + s_fine[0][f]) { lo = s_cb[0] * kFinePer + f; break; } acc += s_fine[0][f]; }
            }
            if (!hi_open) {
                unsigned long long acc = s_cb_before[1];
                for (int f = 0; f < kFinePer; ++f) { if ((unsigned long long)r_hi < acc + s_fine[1][f]) { hi = s_cb[1] * kFinePer + f; break; } acc += s_fine[1][f]; }
            }
            hi = max(hi, lo);
            const uint32_t span = hi - lo;
            // sample mass of the bracket: an upper estimate from the rank bracket itself (ranks r_lo .. r_hi plus the two edge bins)
            const unsigned long long mass = (unsigned long long)((hi_open ? (long long)total - 1 : r_hi) - (lo_open ? 0 : r_lo) + 1)
                                            + s_fine[0][lo % kFinePer] + s_fine[1][hi % kFinePer];
            FusedState* st = p.st;
            st->lo_bin = lo; st->span = span;
            const unsigned long long bst = s_best;
            const uint32_t hot_bin = (uint32_t)(bst & 0xffffffffu);
            const bool hot = (bst >> 32) != 0ull && hot_bin >= lo && hot_bin <= hi;
            st->hot_key = hot ? (hot_bin << 15) : kNoKey;
            uint32_t path = span < (uint32_t)kWin ? PATH_WINDOW : PATH_LIST;
            // list mode lists everything in the bracket but the hot key: only worth it (and only fits) when that is a small share
            if (path == PATH_LIST && (mass > (hot ? (bst >> 32) : 0ull) ? mass - (hot ? (bst >> 32) : 0ull) : 0ull) * 12ull > S) path = PATH_TENSOR;
            if (p.force_fallback) path = PATH_TENSOR;
            st->path = path;
            stamp(5);
            for (int i = 0; i < 6; ++i) st->clk[i] = clk[i];
        }
    }
    cluster.sync();                                          // CTA 0 reads the other CTAs' shared memory until here
}

// ---------------------------------------------------------------------------------------------------------------
// 2. pass A
// ---------------------------------------------------------------------------------------------------------------
template <int DT> struct TileCfg { static constexpr int kU = kQ / DType<DT>::kVec; };   // 16 elements per thread per tile

template <int DT, int MODE, bool STOC>
__global__ void __launch_bounds__(kT) pass_a_kernel(const UParams p) {
    pdl_launch_dependents();
    pdl_wait();
    using D = DType<DT>;
    constexpr int V = D::kVec;
    constexpr int U = TileCfg<DT>::kU;
    constexpr int kTileVecs = kT * U;
    __shared__ unsigned int s_pure[kWin], s_imp[kWin];
    __shared__ uint32_t s_q[kQ][kT];
    __shared__ uint32_t s_buf[kWarps][kWarpBuf];
    __shared__ unsigned long long s_scr[32];
    __shared__ SelectResult s_res;
    __shared__ int s_flag;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    FusedState* st = p.st;
    for (int i = tid; i < kWin; i += kT) { s_pure[i] = 0u; s_imp[i] = 0u; }
    {   // state that is first used by the refine kernel
        uint32_t* w = reinterpret_cast<uint32_t*>(st) + kStateHeaderWords;
        const int words = (int)(sizeof(FusedState) / 4 - kStateHeaderWords);
        for (int i = blockIdx.x * kT + tid; i < words; i += gridDim.x * kT) w[i] = 0u;
    }
    __shared__ uint4 s_head;
    if (tid == 0) s_head = __ldcg(reinterpret_cast<const uint4*>(st));     // lo_bin, span, hot_key, path: one read per CTA (other CTAs may downgrade the path meanwhile)
    __syncthreads();
    const int path = (int)s_head.w;
    if (path == PATH_TENSOR) return;
    const bool list = path == PATH_LIST;
    const uint32_t lo_bin = s_head.x, span = s_head.y, hot_key = s_head.z;
    const uint32_t lo_key = lo_bin << 15;
    const uint32_t hi_key = (lo_bin + span + 1u) << 15;                  // exclusive; 2^31 when the bracket reaches the top
    uint32_t below = 0u, hotc = 0u;
    uint32_t used = 0u;                                      // keys staged in this warp's buffer (uniform over the warp)
    bool dead = false;                                       // the list overflowed: stop listing (path is already PATH_TENSOR)
    auto flush = [&]() {
        if (used == 0u) return;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&st->cand_count, (unsigned long long)used);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base + used > p.cand_cap) {
            if (lane == 0) st->path = PATH_TENSOR;
            dead = true;
        } else {
            for (uint32_t i = lane; i < used; i += 32) p.cand[base + i] = make_uint2(s_buf[warp][i], blockIdx.x);
        }
        used = 0u;
        __syncwarp();
    };
    const int64_t t0 = (int64_t)blockIdx.x * p.tiles_per_range, t1 = min(p.n_tiles, t0 + p.tiles_per_range);
    auto range_loop = [&](auto has_hot_c) {
    constexpr bool has_hot = decltype(has_hot_c)::value;
    // the next tile's vectors are requested before this tile's arithmetic starts (one tile ahead, in registers)
    uint4 nxt[U];
    auto fetch = [&](int64_t t) {
        const int64_t base = t * kTileVecs;
        const int r = (int)min((int64_t)kTileVecs, p.n_vec - base);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int li = tid + u * kT;
            nxt[u] = (li < r) ? ld_stream(p.in + base + li) : make_uint4(0u, 0u, 0u, 0u);
        }
    };
    if (t0 < t1) fetch(t0);
    for (int64_t tile = t0; tile < t1; ++tile) {
        const int64_t tile_base = tile * kTileVecs;
        const int rem = (int)min((int64_t)kTileVecs, p.n_vec - tile_base);
        uint4 raw[U];
#pragma unroll
        for (int u = 0; u < U; ++u) raw[u] = nxt[u];
        if (tile + 1 < t1) fetch(tile + 1);
        // phase 1, branch-free: every key goes to this lane's column of s_q; three sign-bit accumulators (one funnel shift per
        // element each) record key < lo_key, key < hi_key and key == hot_key.  Keys are < 2^31, so a - b is negative iff a < b.
        uint32_t acc_lo = 0u, acc_hi = 0u, acc_hot = 0u;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int li = tid + u * kT;
            float v[V];
            unpack_vec<DT>(raw[u], v);
            if (MODE == MODE_QS) quantize_vec<DT, STOC>(v, p, tile_base + li);
#pragma unroll
            for (int e = 0; e < V; ++e) {
                const uint32_t kk = topk_key(v[e]);
                s_q[u * V + e][tid] = kk;
                acc_lo = __funnelshift_l(kk - lo_key, acc_lo, 1);
                acc_hi = __funnelshift_l(kk - hi_key, acc_hi, 1);
                if (has_hot) acc_hot = __funnelshift_l((kk ^ hot_key) - 1u, acc_hot, 1);     // x - 1 is negative iff x == 0 (x < 2^31)
            }
        }
        // bit (15 - j) of the accumulators belongs to element j; padding slots of the tensor's last tile are masked out
        uint32_t real_mask = 0xffffu;
        if (rem < kTileVecs) {
            real_mask = 0u;
#pragma unroll
            for (int u = 0; u < U; ++u) real_mask |= (tid + u * kT < rem) ? (((1u << V) - 1u) << (kQ - V - u * V)) : 0u;
        }
        acc_lo &= real_mask; acc_hot &= real_mask;
        below += __popc(acc_lo);
        hotc += __popc(acc_hot);
        uint32_t inmask = acc_hi & ~acc_lo & ~acc_hot & real_mask;
        // phase 2: the in-bracket keys (a few per cent of the elements), lanes in step
        const uint32_t nmax = __reduce_max_sync(0xffffffffu, (uint32_t)__popc(inmask));
        for (uint32_t it = 0; it < nmax; ++it) {
            const bool act = inmask != 0u;
            const int pos = act ? 31 - __clz(inmask) : 0;            // highest set bit = earliest element
            if (act) inmask ^= 1u << pos;
            const uint32_t kk = act ? s_q[kQ - 1 - pos][tid] : 0u;
            bool listed = act;
            if (!list) {
                const uint32_t d = (kk >> 15) - lo_bin;
                const bool pure = (kk & 0x7fffu) == 0u;
                if (act) atomicAdd(pure ? &s_pure[d] : &s_imp[d], 1u);
                listed = act && !pure;
            }
            const uint32_t mm = __ballot_sync(0xffffffffu, listed);
            if (mm && !dead) {
                const uint32_t cnt = __popc(mm);
                if (used + cnt > (uint32_t)kWarpBuf) flush();
                if (!dead) {
                    if (listed) s_buf[warp][used + __popc(mm & ((1u << lane) - 1u))] = kk;
                    used += cnt;
                    __syncwarp();
                }
            }
        }
    }
    };
    if (hot_key != kNoKey) range_loop(std::true_type{}); else range_loop(std::false_type{});
    if (!dead) flush();
    const unsigned long long bh = block_sum<kT>((unsigned long long)below | ((unsigned long long)hotc << 32), s_scr);   // a range holds < 2^32 elements
    const unsigned long long b_sum = bh & 0xffffffffull, h_sum = bh >> 32;
    if (tid == 0) {
        if (b_sum) atomicAdd(&st->below, b_sum);
        if (list) {
            p.cta_hist[blockIdx.x] = (uint32_t)h_sum;
            if (h_sum) atomicAdd(&st->hot_total, h_sum);
        } else if (h_sum && hot_key != kNoKey) {
            s_pure[(hot_key >> 15) - lo_bin] += (uint32_t)h_sum;
        }
    }
    __syncthreads();
    if (!list) {
        for (uint32_t d = tid; d <= span; d += kT) {
            const uint32_t pc = s_pure[d], tot = pc + s_imp[d];
            p.cta_hist[(size_t)d * p.n_ranges + blockIdx.x] = pc;
            if (tot) atomicAdd(&st->win_hist[d], tot);
        }
        // the last CTA to arrive picks the bin
        __threadfence();
        __syncthreads();
        if (tid == 0) s_flag = atomicAdd(&st->done_a, 1u) == gridDim.x - 1;
        __syncthreads();
        if (!s_flag) return;
        __threadfence();
        const unsigned long long bel = ld_cg64(&st->below);
        const unsigned long long need_w = p.k > bel ? p.k - bel : 0ull;
        const SelectResult r = block_select<kT, kWin / kT>(st->win_hist, kWin, need_w, kNoKey, 0ull, s_scr, &s_res);
        if (tid == 0) {
            if (!r.found) st->path = PATH_TENSOR;
            else { st->bin = lo_bin + r.bin; st->need_bin = need_w - r.before; st->cnt_bin = r.count; }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// 3. refine (cooperative launch)
// ---------------------------------------------------------------------------------------------------------------
// three-digit radix select (11 + 10 + 10 bits) with grid-wide barriers.  count_pass(pv, pm, shift, dm) adds this CTA's keys that
// match the prefix to the shared histogram s_fb; one extra key value with a known multiplicity can be mixed in.
template <class CountPass>
__device__ __forceinline__ bool grid_radix_select(cg::grid_group& grid, FusedState* st, unsigned long long k_init, uint32_t extra_key, unsigned long long extra_cnt,
                                                  unsigned int* s_fb, unsigned long long* s_scr, SelectResult* s_res, CountPass&& count_pass) {
    const int tid = threadIdx.x;
    bool ok = true;
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
        const int sh = pass == 0 ? 20 : (pass == 1 ? 10 : 0), nb = pass == 0 ? 11 : 10;
        for (int i = tid; i < kFbBins; i += kT) s_fb[i] = 0u;
        __syncthreads();
        const uint32_t pv = ld_cg32(&st->prefix_value), pm = ld_cg32(&st->prefix_mask), dm = (1u << nb) - 1u;
        count_pass(pv, pm, sh, dm);
        __syncthreads();
        for (int i = tid; i < kFbBins; i += kT)
            if (s_fb[i]) atomicAdd(&st->fb_hist[pass][i], s_fb[i]);
        __threadfence();
        grid.sync();
        if (blockIdx.x == 0) {
            const unsigned long long need = pass == 0 ? k_init : ld_cg64(&st->fb_need);
            const bool extra_in = extra_cnt != 0ull && (extra_key & pm) == pv;
            const SelectResult r = block_select<kT, kFbBins / kT>(st->fb_hist[pass], 1 << nb, need, extra_in ? ((extra_key >> sh) & dm) : kNoKey,
                                                                   extra_in ? extra_cnt : 0ull, s_scr, s_res);
            if (tid == 0) {
                if (!r.found) {
                    st->fb_need = 0ull;                       // need outside the keys counted: the select fails
                } else {
                    st->prefix_value = pv | (r.bin << sh);
                    st->prefix_mask = pm | (dm << sh);
                    st->fb_need = need - r.before;
                    if (pass == 2) { st->tau = pv | (r.bin << sh); st->need = need - r.before; st->ties_total = r.count; }
                }
            }
            __threadfence();
        }
        grid.sync();
        if (ld_cg64(&st->fb_need) == 0ull) ok = false;        // uniform over the grid
        if (!ok) break;
    }
    return ok;
}

template <int DT, int MODE, bool STOC>
__global__ void __launch_bounds__(kT) refine_kernel(const UParams p) {
    using D = DType<DT>;
    constexpr int V = D::kVec;
    constexpr int U = TileCfg<DT>::kU;
    constexpr int kTileVecs = kT * U;
    __shared__ unsigned int s_fb[kFbBins];
    __shared__ unsigned long long s_scr[32];
    __shared__ SelectResult s_res;
    pdl_launch_dependents();
    pdl_wait();
    cg::grid_group grid = cg::this_grid();
    const int tid = threadIdx.x;
    FusedState* st = p.st;
    uint32_t path = ld_cg32(&st->path);                      // stable: pass A has completed
    const unsigned long long n_c = min(ld_cg64(&st->cand_count), p.cand_cap);
    const unsigned long long gtid = (unsigned long long)blockIdx.x * kT + tid, gthreads = (unsigned long long)gridDim.x * kT;

    // every listed (key, range), four entries (two 128-bit loads) per thread and iteration
    auto for_each_listed = [&](auto&& fn) {
        const unsigned long long n4 = n_c / 4;
        const uint4* c4 = reinterpret_cast<const uint4*>(p.cand);
#pragma unroll 2
        for (unsigned long long i = gtid; i < n4; i += gthreads) {
            const uint4 a = __ldcg(c4 + 2 * i), b = __ldcg(c4 + 2 * i + 1);
            fn(a.x, a.y); fn(a.z, a.w); fn(b.x, b.y); fn(b.z, b.w);
        }
        for (unsigned long long i = n4 * 4 + gtid; i < n_c; i += gthreads) { const uint2 c = __ldcg(p.cand + i); fn(c.x, c.y); }
    };
    // the listed keys equal to tau, counted per range
    auto count_listed_ties = [&](uint32_t tau) {
        for_each_listed([&](uint32_t kk, uint32_t range) { if (kk == tau) atomicAdd(&st->range_count[range], 1u); });
    };

    if (path == PATH_WINDOW) {
        const uint32_t bin = ld_cg32(&st->bin);
        unsigned long long impure = 0;
        for_each_listed([&](uint32_t kk, uint32_t) { if ((kk >> 15) == bin) { atomicAdd(&st->low_hist[kk & 0x7fffu], 1u); ++impure; } });
        const unsigned long long i_sum = block_sum<kT>(impure, s_scr);
        if (tid == 0 && i_sum) atomicAdd(&st->impure_in_bin, i_sum);
        // the last CTA to arrive selects; the others wait for its flag (every CTA of a cooperative launch is resident, so waiting is
        // safe) -- one grid-wide barrier less than two grid.sync()
        __shared__ int s_last;
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = atomicAdd(&st->done_r, 1u) == gridDim.x - 1;
        __syncthreads();
        if (s_last) {
            __threadfence();
            const unsigned long long cnt = ld_cg64(&st->cnt_bin), imp = ld_cg64(&st->impure_in_bin), need_bin = ld_cg64(&st->need_bin);
            const unsigned long long pure = cnt > imp ? cnt - imp : 0ull;        // keys of the bin with zero trailing bits
            const SelectResult r = block_select<kT, kLowBins / kT>(st->low_hist, kLowBins, need_bin, 0u, pure, s_scr, &s_res);
            if (tid == 0) {
                // r.found is guaranteed: need_bin <= cnt_bin = pure + impure
                const unsigned long long need = need_bin - r.before;
                st->tau = (bin << 15) | r.bin;
                st->need = need;
                st->ties_total = r.count;
                st->tie_src = need == r.count ? TIES_ALL : (r.bin == 0u ? TIES_ROW : TIES_COUNTED);
                st->tie_row = bin - ld_cg32(&st->lo_bin);
                __threadfence();
                *reinterpret_cast<volatile uint32_t*>(&st->flag_r) = 1u;
            }
            __syncthreads();
        } else {
            if (tid == 0) { while (*reinterpret_cast<volatile uint32_t*>(&st->flag_r) == 0u) { } __threadfence(); }
            __syncthreads();
        }
        if (ld_cg32(&st->tie_src) == TIES_COUNTED) count_listed_ties(ld_cg32(&st->tau));
        return;
    }

    if (path == PATH_LIST) {
        const unsigned long long bel = ld_cg64(&st->below);
        const uint32_t hot_key = ld_cg32(&st->hot_key);
        const unsigned long long hot_total = ld_cg64(&st->hot_total);
        const bool ok = grid_radix_select(grid, st, p.k > bel ? p.k - bel : 0ull, hot_key, hot_total, s_fb, s_scr, &s_res,
                                          [&](uint32_t pv, uint32_t pm, int sh, uint32_t dm) {
                                              for_each_listed([&](uint32_t kk, uint32_t) { if ((kk & pm) == pv) atomicAdd(&s_fb[(kk >> sh) & dm], 1u); });
                                          });
        if (ok) {
            const uint32_t tau = ld_cg32(&st->tau);
            const bool some = ld_cg64(&st->need) < ld_cg64(&st->ties_total);
            if (blockIdx.x == 0 && tid == 0) { st->tie_src = !some ? TIES_ALL : (tau == hot_key ? TIES_ROW : TIES_COUNTED); st->tie_row = 0u; }
            if (some && tau != hot_key) count_listed_ties(tau);
            return;
        }
        // the bracket missed: start over on the whole tensor
        if (blockIdx.x == 0) {
            for (int i = tid; i < 3 * kFbBins; i += kT) (&st->fb_hist[0][0])[i] = 0u;
            if (tid == 0) { st->prefix_value = 0u; st->prefix_mask = 0u; st->fb_need = 0ull; st->path = PATH_TENSOR; }
            __threadfence();
        }
        grid.sync();
        path = PATH_TENSOR;
    }

    // ---- whole-tensor select: every CTA of the grid takes part ----
    const int64_t n_tiles = p.n_tiles;
    auto for_each_key = [&](int64_t tile, auto&& fn) {
        const int64_t tile_base = tile * kTileVecs;
        const int rem = (int)min((int64_t)kTileVecs, p.n_vec - tile_base);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int li = tid + u * kT;
            const uint4 raw = (li < rem) ? ld_stream(p.in + tile_base + li) : make_uint4(0u, 0u, 0u, 0u);
            float v[V];
            unpack_vec<DT>(raw, v);
            if (MODE == MODE_QS) quantize_vec<DT, STOC>(v, p, tile_base + li);
            if (li < rem) {
#pragma unroll
                for (int e = 0; e < V; ++e) fn(topk_key(v[e]));
            }
        }
    };
    grid_radix_select(grid, st, p.k, kNoKey, 0ull, s_fb, s_scr, &s_res, [&](uint32_t pv, uint32_t pm, int sh, uint32_t dm) {
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
            for_each_key(tile, [&](uint32_t kk) { if ((kk & pm) == pv) atomicAdd(&s_fb[(kk >> sh) & dm], 1u); });
    });
    const uint32_t tau = ld_cg32(&st->tau);
    const bool some = ld_cg64(&st->need) < ld_cg64(&st->ties_total);
    if (blockIdx.x == 0 && tid == 0) { st->tie_src = some ? TIES_COUNTED : TIES_ALL; st->tie_row = 0u; }
    if (some) {
        // one more read: the keys equal to tau per range
        for (int r = blockIdx.x; r < p.n_ranges; r += gridDim.x) {
            unsigned long long c = 0;
            const int64_t t0 = (int64_t)r * p.tiles_per_range, t1 = min(n_tiles, t0 + p.tiles_per_range);
            for (int64_t tile = t0; tile < t1; ++tile) for_each_key(tile, [&](uint32_t kk) { c += kk == tau ? 1u : 0u; });
            const unsigned long long tot = block_sum<kT>(c, s_scr);
            if (tid == 0) st->range_count[r] = (uint32_t)tot;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// 4. apply
// ---------------------------------------------------------------------------------------------------------------
template <int DT, int MODE, bool STOC>
__global__ void __launch_bounds__(kT, 3) apply_kernel(const UParams p) {
    pdl_launch_dependents();
    pdl_wait();
    using D = DType<DT>;
    constexpr int V = D::kVec;
    constexpr int U = TileCfg<DT>::kU;
    constexpr int kTileVecs = kT * U;
    constexpr int kOutVecs = (STOC && V == 8 && MODE != MODE_S_ONLY) ? 2 : 1;
    __shared__ unsigned int s_w[2][kWarps];
    __shared__ unsigned long long s_scr[32];
    __shared__ unsigned char s_keep[kMaxRanges];              // 1: the keys equal to tau of this range stay; 0: they go
    __shared__ int s_boundary;                                // the range where the cut falls (-1: none)
    __shared__ unsigned long long s_running;                  // keys equal to tau before that range
    __shared__ uint4 s_head;
    FusedState* st = p.st;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        s_head = make_uint4(ld_cg32(&st->tau), ld_cg32(&st->tie_src), ld_cg32(&st->tie_row), 0u);
        s_boundary = -1; s_running = 0ull;
    }
    __syncthreads();
    const uint32_t tau = s_head.x, tie_src = s_head.y;
    const unsigned long long need = ld_cg64(&st->need);

    // what happens to the keys equal to tau, range by range: all go (ranges before the cut), none go (after it), or -- in the one
    // range where the cut falls -- the first (need - running) in index order
    if (tie_src != TIES_ALL) {
        const uint32_t* cnt = tie_src == TIES_ROW ? p.cta_hist + (size_t)s_head.z * p.n_ranges : st->range_count;
        unsigned long long carry = 0;
        for (int c0 = 0; c0 < p.n_ranges; c0 += kT) {
            const int r = c0 + tid;
            const unsigned long long here = r < p.n_ranges ? ld_cg32(cnt + r) : 0u;
            unsigned long long total;
            const unsigned long long before = carry + block_excl_scan<kT>(here, s_scr, &total);
            if (r < p.n_ranges) {
                s_keep[r] = before >= need ? 1 : 0;
                if (before < need && need < before + here) { s_boundary = r; s_running = before; }
            }
            carry += total;
        }
        __syncthreads();
    }
    const int boundary = s_boundary;
    const bool all_go = tie_src == TIES_ALL;

    auto store_vec = [&](int64_t vec, const float* v) {
        if (kOutVecs == 1) {
            st_stream(p.out + vec, (STOC && MODE != MODE_S_ONLY) ? pack_vec<BFP_DT_F32>(v) : pack_vec<DT>(v));
        } else {
            st_stream(p.out + vec * 2, pack_vec<BFP_DT_F32>(v));
            st_stream(p.out + vec * 2 + 1, pack_vec<BFP_DT_F32>(v + (V == 8 ? 4 : 0)));
        }
    };

    // the last CTA walks the boundary range in order; the others (all of them when there is none) stream the remaining tiles
    const int n_workers = max(1, (int)gridDim.x - (boundary >= 0 ? 1 : 0));
    if (boundary >= 0 && blockIdx.x == gridDim.x - 1) {
        unsigned long long running = s_running;
        const int64_t t0 = (int64_t)boundary * p.tiles_per_range, t1 = min(p.n_tiles, t0 + p.tiles_per_range);
        uint4 nraw[U];                                        // the next tile's vectors, requested one tile ahead
        auto bfetch = [&](int64_t t) {
            const int64_t base = t * kTileVecs;
            const int r = (int)min((int64_t)kTileVecs, p.n_vec - base);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int li = tid + u * kT;
                nraw[u] = (li < r) ? ld_stream(p.in + base + li) : make_uint4(0u, 0u, 0u, 0u);
            }
        };
        if (t0 < t1) bfetch(t0);
        for (int64_t tile = t0; tile < t1; ++tile) {
            const int64_t tile_base = tile * kTileVecs;
            const int rem = (int)min((int64_t)kTileVecs, p.n_vec - tile_base);
            float v[U][V];
            uint32_t tm[U];                                   // bit e: element e of vector u equals tau
#pragma unroll
            for (int u = 0; u < U; ++u) unpack_vec<DT>(nraw[u], v[u]);
            if (tile + 1 < t1) bfetch(tile + 1);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int li = tid + u * kT;
                if (MODE == MODE_QS) quantize_vec<DT, STOC>(v[u], p, tile_base + li);
                tm[u] = 0u;
#pragma unroll
                for (int e = 0; e < V; ++e) tm[u] |= (li < rem && topk_key(v[u][e]) == tau) ? (1u << e) : 0u;
            }
            // packed scan of the per-vector tie counts (field totals <= 256 * 8 < 65536); vector order inside the tile is u * kT + tid
            uint32_t c01 = __popc(tm[0]) | (__popc(tm[1]) << 16), c23 = U == 4 ? (__popc(tm[U - 2]) | (__popc(tm[U - 1]) << 16)) : 0u;
            uint32_t i01 = c01, i23 = c23;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t o01 = __shfl_up_sync(0xffffffffu, i01, off), o23 = __shfl_up_sync(0xffffffffu, i23, off);
                if (lane >= off) { i01 += o01; i23 += o23; }
            }
            __syncthreads();                                  // s_w of the previous tile has been read
            if (lane == 31) { s_w[0][warp] = i01; s_w[1][warp] = i23; }
            __syncthreads();
            uint32_t wb01 = 0u, wb23 = 0u, t01 = 0u, t23 = 0u;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) {
                const uint32_t a = s_w[0][w], bb = s_w[1][w];
                wb01 += w < warp ? a : 0u; wb23 += w < warp ? bb : 0u; t01 += a; t23 += bb;
            }
            const uint32_t tot[4] = {t01 & 0xffffu, t01 >> 16, t23 & 0xffffu, t23 >> 16};
            const uint32_t ex01 = wb01 + i01 - c01, ex23 = wb23 + i23 - c23;      // exclusive prefixes within each field
            const uint32_t exu[4] = {ex01 & 0xffffu, ex01 >> 16, ex23 & 0xffffu, ex23 >> 16};
            uint32_t before_u = 0u;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int li = tid + u * kT;
                unsigned long long rank = running + before_u + exu[u];
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    const uint32_t kk = topk_key(v[u][e]);
                    bool drop = kk < tau;
                    if (tm[u] & (1u << e)) { drop = rank < need; ++rank; }
                    v[u][e] = drop ? 0.0f : v[u][e];
                }
                before_u += tot[u];
                if (MODE == MODE_SQ) quantize_vec<DT, STOC>(v[u], p, tile_base + li);
                if (li < rem) store_vec(tile_base + li, v[u]);
            }
            running += before_u;
        }
        if (gridDim.x > 1) return;
    }
    if ((int)blockIdx.x >= n_workers) return;

    // streaming tiles: the next tile's vectors are requested before this tile's arithmetic starts
    uint4 nxt[U];
    auto fetch = [&](int64_t t) {
        const int64_t base = t * kTileVecs;
        const int r = (int)min((int64_t)kTileVecs, p.n_vec - base);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int li = tid + u * kT;
            nxt[u] = (li < r) ? ld_stream(p.in + base + li) : make_uint4(0u, 0u, 0u, 0u);
        }
    };
    if ((int64_t)blockIdx.x < p.n_tiles) fetch(blockIdx.x);
    for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += n_workers) {
        const int64_t tile_base = tile * kTileVecs;
        const int rem = (int)min((int64_t)kTileVecs, p.n_vec - tile_base);
        uint4 raw[U];
#pragma unroll
        for (int u = 0; u < U; ++u) raw[u] = nxt[u];
        if (tile + n_workers < p.n_tiles) fetch(tile + n_workers);
        const int range = (int)(tile / p.tiles_per_range);
        if (range == boundary) continue;
        const uint32_t lim = tau + ((all_go || !s_keep[range]) ? 1u : 0u);        // drop key < lim (keys <= 0x7f800001: no overflow)
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int li = tid + u * kT;
            float v[V];
            unpack_vec<DT>(raw[u], v);
            if (MODE == MODE_QS) quantize_vec<DT, STOC>(v, p, tile_base + li);
#pragma unroll
            for (int e = 0; e < V; ++e) v[e] = topk_key(v[e]) < lim ? 0.0f : v[e];
            if (MODE == MODE_SQ) quantize_vec<DT, STOC>(v, p, tile_base + li);
            if (li < rem) store_vec(tile_base + li, v);
        }
    }
}

template <class Kernel>
static cudaError_t launch_one(Kernel kernel, int grid, int threads, size_t smem, cudaStream_t s, bool pdl, bool cooperative, const UParams& p) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)threads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[2];
    int na = 0;
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[na].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0; ++na;
    if (cooperative) { attr[na].id = cudaLaunchAttributeCooperative; attr[na].val.cooperative = 1; ++na; }
    cfg.attrs = attr; cfg.numAttrs = na;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, p);
    if (e != cudaSuccess && cooperative && pdl) {
        // a driver that refuses a cooperative launch with programmatic serialization: plain stream order (the kernel's
        // griddepcontrol.wait is then a no-op)
        cudaGetLastError();
        attr[0].val.programmaticStreamSerializationAllowed = 0;
        e = cudaLaunchKernelEx(&cfg, kernel, p);
    }
    return e;
}

struct WsLayout { size_t state, rows, cand, total; unsigned long long cap; };
static WsLayout ws_layout(int64_t n) {
    WsLayout w;
    w.state = (sizeof(FusedState) + 255) / 256 * 256;
    w.rows = (size_t)kWin * kMaxRanges * 4;                       // per-range counts of every window bin (only the used part is touched)
    w.cap = (unsigned long long)std::max<int64_t>(65536, n / 8);
    w.cand = (size_t)(w.cap * 8 + 255) / 256 * 256;
    w.total = w.state + w.rows + w.cand;
    return w;
}

template <int DT, int MODE, bool STOC>
int run_fused(const UnstructuredArgs& a, cudaStream_t s) {
    constexpr int V = DType<DT>::kVec;
    constexpr int U = TileCfg<DT>::kU;
    const int64_t n = a.n;
    const WsLayout w = ws_layout(n);
    const int sms = device_info().sm_count;
    UParams p = {};
    p.in = static_cast<const uint4*>(a.in);
    p.out = static_cast<uint4*>(a.out);
    p.n_vec = n / V;
    p.n_tiles = (p.n_vec + kT * U - 1) / (kT * U);
    p.k = a.k; p.n = (unsigned long long)n;
    p.lanes_per_block = MODE == MODE_S_ONLY ? 1 : a.B / V;
    p.m = a.m; p.eps = a.eps; p.seed = a.seed; p.offset = a.offset;
    char* base = static_cast<char*>(a.workspace);
    p.st = reinterpret_cast<FusedState*>(base);
    p.cta_hist = reinterpret_cast<uint32_t*>(base + w.state);
    p.cand = reinterpret_cast<uint2*>(base + w.state + w.rows);
    p.cand_cap = w.cap;
    p.force_fallback = tuning().unstructured_force_fallback;
    const bool pdl = tuning().pdl != 0;

    auto k_sample = sample_kernel<DT, MODE, STOC>;
    auto k_a = pass_a_kernel<DT, MODE, STOC>;
    auto k_r = refine_kernel<DT, MODE, STOC>;
    auto k_ap = apply_kernel<DT, MODE, STOC>;
    // per device, once per instantiation: the sample kernel's shared-memory opt-in and the resident CTAs per SM of the refine grid
    struct PerDevice { bool init = false; int occ_r = 0, occ_a = 0, occ_ap = 0; };
    static PerDevice per_device[64];
    PerDevice& pd = per_device[std::max(0, std::min(63, device_info().device))];
    if (!pd.init) {
        cudaFuncSetAttribute(k_sample, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
        pd.occ_r = kernel_occupancy(k_r, kT);
        pd.occ_a = kernel_occupancy(k_a, kT);
        pd.occ_ap = kernel_occupancy(k_ap, kT);
        pd.init = true;
    }
    // contiguous ranges of whole tiles: one CTA of pass A each, all resident at once (a single wave)
    const int64_t want = std::min<int64_t>(kMaxRanges, (int64_t)sms * pd.occ_a);
    p.tiles_per_range = (int)std::max<int64_t>(1, (p.n_tiles + want - 1) / want);
    p.n_ranges = (int)((p.n_tiles + p.tiles_per_range - 1) / p.tiles_per_range);
    static const bool dbg = getenv("BFP_UNSTRUCTURED_TIMING") != nullptr;      // per-phase device times on stderr (tools only)
    cudaEvent_t ev[5]; int nev = 0;
    auto mark = [&] { if (dbg) { cudaEventCreate(&ev[nev]); cudaEventRecord(ev[nev], s); ++nev; } };
    mark();
    cudaError_t e = launch_one(k_sample, kSampleCtas, kSampleThreads, 131072, s, pdl, false, p);
    if (e != cudaSuccess) return set_errorf(BFP_E_CUDA, "unstructured sample kernel: %s", cudaGetErrorString(e));
    count_launch();
    mark();
    e = launch_one(k_a, p.n_ranges, kT, 0, s, pdl, false, p);
    if (e != cudaSuccess) return set_errorf(BFP_E_CUDA, "unstructured pass A: %s", cudaGetErrorString(e));
    count_launch();
    mark();
    const int grid_r = (int)std::max<int64_t>(1, std::min<int64_t>(p.n_tiles, (int64_t)sms * std::min(pd.occ_r, 2)));
    e = launch_one(k_r, grid_r, kT, 0, s, pdl, true, p);
    if (e != cudaSuccess) return set_errorf(BFP_E_CUDA, "unstructured refine kernel (cooperative launch): %s", cudaGetErrorString(e));
    count_launch();
    mark();
    const int grid_ap = (int)std::max<int64_t>(1, std::min<int64_t>(p.n_tiles, (int64_t)sms * pd.occ_ap));
    e = launch_one(k_ap, grid_ap, kT, 0, s, pdl, false, p);
    if (e != cudaSuccess) return set_errorf(BFP_E_CUDA, "unstructured apply kernel: %s", cudaGetErrorString(e));
    count_launch();
    mark();
    if (dbg) {
        cudaStreamSynchronize(s);
        static const char* names[4] = {"sample", "pass_a", "refine", "apply"};
        for (int i = 0; i + 1 < nev; ++i) { float ms = 0; cudaEventElapsedTime(&ms, ev[i], ev[i + 1]); fprintf(stderr, "%s %.1f us  ", names[i], ms * 1e3f); }
        FusedState h;
        cudaMemcpy(&h, p.st, offsetof(FusedState, win_hist), cudaMemcpyDeviceToHost);
        fprintf(stderr, "| sample clk: sampled %lld coarse+sync %lld merge %lld hot %lld rest %lld ", h.clk[1] - h.clk[0], h.clk[2] - h.clk[1], h.clk[3] - h.clk[2],
                h.clk[4] - h.clk[3], h.clk[5] - h.clk[4]);
        fprintf(stderr, "| path %u bracket [%u, +%u] hot %08x below %llu listed %llu bin %u tau %08x need %llu ties %llu tie_src %u ranges %d x %d tiles, refine grid %d\n",
                h.path, h.lo_bin, h.span, h.hot_key, h.below, h.cand_count, h.bin, h.tau, h.need, h.ties_total, h.tie_src, p.n_ranges, p.tiles_per_range, grid_r);
        for (int i = 0; i < nev; ++i) cudaEventDestroy(ev[i]);
    }
    return check_launch("fused unstructured sparsity kernels");
}

template <int DT>
int dispatch_mode(const UnstructuredArgs& a, cudaStream_t s) {
    const bool stoc = a.rounding == BFP_ROUND_STOCHASTIC;
    switch (a.order) {
    case BFP_ORDER_SPARSIFY_ONLY: return run_fused<DT, MODE_S_ONLY, false>(a, s);
    case BFP_ORDER_SPARSIFY_QUANT: return stoc ? run_fused<DT, MODE_SQ, true>(a, s) : run_fused<DT, MODE_SQ, false>(a, s);
    case BFP_ORDER_QUANT_SPARSIFY: return stoc ? run_fused<DT, MODE_QS, true>(a, s) : run_fused<DT, MODE_QS, false>(a, s);
    }
    return set_error(BFP_E_ARG, "order must be SPARSIFY_ONLY, SPARSIFY_QUANT or QUANT_SPARSIFY");
}
}  // namespace

size_t unstructured_fused_workspace_bytes(int64_t n, int /*dtype*/) { return ws_layout(std::max<int64_t>(n, 0)).total; }

// which calls the two-read pipeline takes; everything else composes bfp_unstructured_sparsify and bfp_quantize
bool unstructured_fused_supported(const UnstructuredArgs& a) {
    const int V = a.dtype == BFP_DT_F32 ? 4 : 8;
    if (a.n <= 0 || a.n % V || a.n >= (int64_t(1) << 32)) return false;      // 32-bit counters per bin
    if (reinterpret_cast<uintptr_t>(a.in) % 16 || reinterpret_cast<uintptr_t>(a.out) % 16 || reinterpret_cast<uintptr_t>(a.workspace) % 16) return false;
    if (a.order == BFP_ORDER_SPARSIFY_ONLY) return true;
    const int B = a.B;
    if (B < V || (B & (B - 1)) || B / V > 32) return false;          // a block = 2^j adjacent lanes of a warp
    if (a.K <= 0 || a.K % B) return false;
    return true;
}

int unstructured_fused_device(const UnstructuredArgs& a, cudaStream_t s) {
    if (a.k == 0 || a.k >= (unsigned long long)a.n) return set_error(BFP_E_ARG, "k must be in (0, numel)");
    if (a.dtype != BFP_DT_F32 && a.order != BFP_ORDER_SPARSIFY_ONLY) ensure_exp_tables(s);    // this translation unit's copy of the half-precision exponent table
    switch (a.dtype) {
    case BFP_DT_F32: return dispatch_mode<BFP_DT_F32>(a, s);
    case BFP_DT_F16: return dispatch_mode<BFP_DT_F16>(a, s);
    case BFP_DT_BF16: return dispatch_mode<BFP_DT_BF16>(a, s);
    }
    return set_error(BFP_E_ARG, "bad dtype");
}

}  // namespace bfp
