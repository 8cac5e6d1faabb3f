// bfp_unstructured_fused.cu -- global magnitude pruning fused with the BFP quantiser: two reads and one write of the tensor.
//
// Replaces float_to_bfp_blocked (bfp_ops.py:124-149) with sparsity_mode == 'unstructured' -- _unstructured_sparsity (:61-71) and
// _no_sparsity_float_to_bfp (:46-59) in either order -- and _unstructured_sparsity alone.  The k = int(numel * frac) entries
// torch.topk(|t|, k, largest=False) returns on torch-CUDA are dropped: every key strictly below the k-th smallest key tau, and of
// the keys equal to tau the first (k - #below) in index order.  key = bit pattern of |value| as fp32 (monotone for fp16 / bf16
// values too; NaN largest).  With first == 'q' the keys are those of the QUANTISED values, recomputed on the fly in every pass.
//
// The multi-pass radix select of bfp_unstructured.cu reads the tensor four times (plus a separate quantiser pass).  Here:
//   1. sample_kernel   one CTA.  A stratified sample of ~16 K elements goes into a shared-memory histogram of the 16 leading key
//                      bits; the bins holding the sample quantiles k/n -+ 5.5 sigma bracket tau: a WINDOW of at most 2048 bins
//                      (a few per cent of the tensor's mass).  Also zeroes the workspace.
//   2. pass_a_kernel   first full read.  Counts the keys below the window exactly, histograms the keys inside it at 16-bit
//                      granularity, and appends every in-window key whose 15 trailing bits are not all zero to a candidate list
//                      (a few per cent of n).  Keys with zero trailing bits -- 0.0, every BFP value with mant_bits <= 8, bf16
//                      data: where the massive ties are -- are fully described by their bin and never listed.  The last CTA to
//                      finish picks the bin that holds the k-th key.
//   3. refine_kernel   the candidates of that bin resolve the 15 trailing bits -> tau, how many keys equal to tau go (`need`) and
//                      how many there are (`ties_total`).  A cooperative launch: if the bracket missed (probability ~1e-7), the
//                      window was too wide or the candidate list overflowed (massive ties on a non-round value), the same grid
//                      runs a three-digit radix select over the whole tensor with grid-wide barriers instead -- no host round trip.
//   4. apply_kernel    second read + the write: mask (key < tau, or key == tau and tie rank < need), BFP quantisation before or
//                      after it, 128-bit stores.  When only some of the ties go, tiles are handed out in index order and a chained
//                      scan with decoupled look-back gives every tile the number of ties before it.
// = 2 reads + 1 write (12 B / element fp32) for 8 B / element algorithmic.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstddef>
#include <cstdio>
#include <cstdlib>

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
constexpr int kSampleVecs = 4096;          // 128-bit vectors in the sample (16 K fp32 / 32 K half elements)
constexpr int kWarpBuf = 512;              // candidate keys a warp stages in shared memory between flushes
constexpr int kFbBins = 2048;              // fallback radix digits: 11 + 10 + 10 bits
constexpr uint32_t kNoBin = 0xffffffffu;

enum { MODE_S_ONLY = 0, MODE_SQ = 1, MODE_QS = 2 };

struct FusedState {
    uint32_t lo_bin, span;                 // window = bins [lo_bin, lo_bin + span]
    uint32_t hot_bin;                      // a window bin holding > 1/64 of the sample (counted in registers), or kNoBin
    uint32_t valid;                        // 1 while the two-read path holds; 0 -> refine_kernel selects over the whole tensor
    unsigned int done_a, done_r, tile_counter, pad0;
    unsigned long long below;              // keys in bins below the window
    unsigned long long cand_count;         // candidate keys appended (may exceed the capacity: then valid = 0)
    uint32_t bin, pad1;                    // the bin holding the k-th smallest key
    unsigned long long need_bin, cnt_bin;  // rank of that key inside the bin (1-based), keys in the bin
    unsigned long long impure_in_bin;      // candidates found in the bin
    uint32_t tau, pad2;
    unsigned long long need, ties_total;   // of the ties_total keys equal to tau the first `need` (index order) are dropped
    uint32_t prefix_value, prefix_mask;    // fallback radix select
    unsigned long long fb_need;
    unsigned long long win_hist[kWin];
    unsigned long long low_hist[kLowBins];
    unsigned long long fb_hist[3][kFbBins];
};

struct UParams {
    const uint4* in;
    uint4* out;
    int64_t n_vec, n_tiles;
    unsigned long long k, n;
    int lanes_per_block, m;
    float eps;
    uint64_t seed, offset;
    FusedState* st;
    unsigned long long* tile_state;
    uint32_t* cand;
    unsigned long long cand_cap;
    int force_fallback;
};

__device__ __forceinline__ unsigned long long ld_cg64(const unsigned long long* p) { return __ldcg(p); }
__device__ __forceinline__ uint32_t ld_cg32(const uint32_t* p) { return __ldcg(p); }
__device__ __forceinline__ unsigned long long ld_volatile64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

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

// The bin holding the need-th smallest key (1-based) of a histogram in global memory, by the whole CTA (thread t owns PER
// consecutive bins).  `extra0` is added to bin 0.  Returns (through smem result) bin, keys before it, keys in it; found = 0 when
// need is outside [1, total].
struct SelectResult { uint32_t bin; int found; unsigned long long before, count; };
// T threads, T * PER bins: chunk c = bins [c PER, (c + 1) PER).  Chunk sums are formed by warps with coalesced loads, a block scan
// over the T chunk sums finds the chunk, one warp walks its PER bins.  Every load is coalesced and nothing is chained.
template <int T, int PER>
__device__ __forceinline__ SelectResult block_select(const unsigned long long* hist, int nbins, unsigned long long need, unsigned long long extra0,
                                                      unsigned long long* scratch, SelectResult* s_res) {
    static_assert(PER == 8 || PER % 32 == 0, "chunk = 8 bins (one thread) or a multiple of 32 (one warp pass)");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ unsigned long long s_chunk[T];
    auto bin_at = [&](int b) { return (b < nbins ? ld_cg64(hist + b) : 0ull) + (b == 0 ? extra0 : 0ull); };
    unsigned long long c8[PER == 8 ? 8 : 1];
    if (PER == 8) {
        unsigned long long mine = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) { c8[PER == 8 ? j : 0] = bin_at((int)threadIdx.x * 8 + j); mine += c8[PER == 8 ? j : 0]; }
        s_chunk[threadIdx.x] = mine;
    } else {
        // warp w sums chunks [32 w, 32 w + 32): lane l reads bins l, l + 32, ... of the chunk
        for (int c0 = 0; c0 < 32; c0 += 4) {
            unsigned long long part[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int base = (warp * 32 + c0 + q) * PER;
#pragma unroll
                for (int j = 0; j < PER / 32; ++j) part[q] += bin_at(base + j * 32 + lane);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
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
            unsigned long long run = s_before;
            const int base = s_chunk_sel * PER;
            bool done = false;
            for (int j = 0; j < PER / 32 && !done; ++j) {
                const unsigned long long c = bin_at(base + j * 32 + lane);
                unsigned long long incl = c;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const unsigned long long o = __shfl_up_sync(0xffffffffu, incl, off);
                    if (lane >= off) incl += o;
                }
                const unsigned long long bef = run + incl - c;
                const bool hit = bef < need && need <= bef + c;
                if (hit) { s_res->bin = (uint32_t)(base + j * 32 + lane); s_res->before = bef; s_res->count = c; s_res->found = 1; }
                done = __any_sync(0xffffffffu, hit);
                run += __shfl_sync(0xffffffffu, incl, 31);
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
// 1. sample
// ---------------------------------------------------------------------------------------------------------------
template <int DT, int MODE, bool STOC>
__global__ void __launch_bounds__(kSampleThreads) sample_kernel(const UParams p) {
    pdl_launch_dependents();
    pdl_wait();
    using D = DType<DT>;
    constexpr int V = D::kVec;
    extern __shared__ unsigned int s_h[];                    // 65536 16-bit counters, two per word
    __shared__ unsigned long long s_scr[32];
    __shared__ uint32_t s_lo, s_hi;
    __shared__ unsigned int s_best;
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < 32768; i += kSampleThreads) s_h[i] = 0u;
    {
        uint32_t* w = reinterpret_cast<uint32_t*>(p.st);
        for (int i = tid; i < (int)(sizeof(FusedState) / 4); i += kSampleThreads) w[i] = 0u;
    }
    if (tid == 0) { s_lo = 0u; s_hi = 65535u; s_best = 0u; }
    __syncthreads();

    // stratified sample: S_u units (a unit = one vector, or one BFP block when the keys are those of quantised values), unit j
    // taken at a hashed position inside the j-th of S_u equal strata
    const int L = MODE == MODE_QS ? p.lanes_per_block : 1;
    const int64_t n_units = p.n_vec / L;
    const int64_t S_u = n_units < (int64_t)(kSampleVecs / L) ? n_units : (int64_t)(kSampleVecs / L);
    const int64_t stride = n_units / S_u;                    // >= 1
    const int S_v = (int)(S_u * L);
    for (int j0 = tid & ~31; j0 < S_v; j0 += kSampleThreads) {
        const int j = j0 + lane;
        const bool active = j < S_v;
        int64_t vec = 0;
        if (active) {
            const int64_t unit = j / L;
            const int64_t u = unit * stride + (int64_t)(hash32((uint32_t)unit) % (uint32_t)(stride < 0x7fffffff ? stride : 0x7fffffff));
            vec = u * L + (j % L);
        }
        float v[V];
        const uint4 raw = active ? ld_stream(p.in + vec) : make_uint4(0u, 0u, 0u, 0u);
        unpack_vec<DT>(raw, v);
        if (MODE == MODE_QS) quantize_vec<DT, STOC>(v, p, vec);
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const uint32_t bin = topk_key(v[e]) >> 15;
            // up to four rounds of leader aggregation (massive ties: zeros, quantised values), then plain atomics
            bool todo = active;
#pragma unroll 1
            for (int r = 0; r < 4; ++r) {
                const uint32_t act = __ballot_sync(0xffffffffu, todo);
                if (act == 0u) break;
                const int leader = __ffs(act) - 1;
                const uint32_t lb = __shfl_sync(0xffffffffu, bin, leader);
                const bool same = todo && bin == lb;
                const uint32_t mm = __ballot_sync(0xffffffffu, same);
                if (lane == leader) atomicAdd(&s_h[lb >> 1], (unsigned int)__popc(mm) << (16 * (lb & 1u)));
                todo = todo && !same;
            }
            if (todo) atomicAdd(&s_h[bin >> 1], 1u << (16 * (bin & 1u)));
        }
    }
    __syncthreads();

    // rank bracket of the k-th smallest key inside the sorted sample
    const unsigned long long S = (unsigned long long)S_v * V;
    long long r_lo, r_hi;
    if ((unsigned long long)S_v == (unsigned long long)p.n_vec) {
        r_lo = r_hi = (long long)p.k - 1;                    // the sample is the tensor
    } else {
        const double q = (double)p.k / (double)p.n;
        const double mean = q * (double)S;
        const double sd = sqrt((double)S * q * (1.0 - q) * (MODE == MODE_QS ? 4.0 : 1.0));   // design effect: a block shares its scale
        r_lo = (long long)floor(mean - 5.5 * sd) - 1;
        r_hi = (long long)ceil(mean + 5.5 * sd) + 1;
    }
    // thread t owns bins [64 t, 64 t + 64)
    unsigned long long mine = 0;
    for (int i = 0; i < 32; ++i) { const unsigned int w = s_h[tid * 32 + ((i + lane) & 31)]; mine += (w & 0xffffu) + (w >> 16); }   // rotated: bank = lane
    unsigned long long total;
    const unsigned long long before = block_excl_scan<kSampleThreads>(mine, s_scr, &total);
    for (int which = 0; which < 2; ++which) {
        const long long r = which ? r_hi : r_lo;
        if (r >= 0 && (unsigned long long)r < total && before <= (unsigned long long)r && (unsigned long long)r < before + mine) {
            unsigned long long acc = before;
            for (int i = 0; i < 64; ++i) {
                const unsigned int w = s_h[tid * 32 + (i >> 1)];
                const unsigned int c = (i & 1) ? (w >> 16) : (w & 0xffffu);
                if ((unsigned long long)r < acc + c) { if (which) s_hi = (uint32_t)(tid * 64 + i); else s_lo = (uint32_t)(tid * 64 + i); break; }
                acc += c;
            }
        }
    }
    __syncthreads();
    const uint32_t lo = s_lo, hi = max(s_hi, s_lo), span = hi - lo;
    // hottest window bin of the sample
    for (uint32_t d = tid; d <= span && d < (uint32_t)kWin; d += kSampleThreads) {
        const uint32_t b = lo + d;
        const unsigned int w = s_h[b >> 1];
        const unsigned int c = (b & 1u) ? (w >> 16) : (w & 0xffffu);
        if (c) atomicMax(&s_best, (c << 16) | d);           // d < 2048 < 65536; c <= 32768 fits in 16 bits
    }
    __syncthreads();
    if (tid == 0) {
        FusedState* st = p.st;
        st->lo_bin = lo; st->span = span;
        const unsigned int best = s_best;
        st->hot_bin = ((unsigned long long)(best >> 16) * 64ull > S) ? lo + (best & 0xffffu) : kNoBin;
        st->valid = (span < (uint32_t)kWin && !p.force_fallback) ? 1u : 0u;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// 2. pass A
// ---------------------------------------------------------------------------------------------------------------
template <int DT> struct TileCfg { static constexpr int kU = DType<DT>::kVec == 4 ? 4 : 2; };   // 16 elements per thread per tile

template <int DT, int MODE, bool STOC>
__global__ void __launch_bounds__(kT) pass_a_kernel(const UParams p) {
    pdl_launch_dependents();
    pdl_wait();
    using D = DType<DT>;
    constexpr int V = D::kVec;
    constexpr int U = TileCfg<DT>::kU;
    constexpr int kTileVecs = kT * U;
    __shared__ unsigned int s_hist[kWin];
    __shared__ uint32_t s_buf[kWarps][kWarpBuf];
    __shared__ unsigned long long s_scr[32];
    __shared__ SelectResult s_res;
    __shared__ int s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    FusedState* st = p.st;
    for (int i = tid; i < kWin; i += kT) s_hist[i] = 0u;
    // look-back states of apply_kernel's tiles
    for (int64_t i = (int64_t)blockIdx.x * kT + tid; i < p.n_tiles; i += (int64_t)gridDim.x * kT) p.tile_state[i] = 0ull;
    if (tid == 0) s_last = (int)ld_cg32(&st->valid);       // one read per CTA: other CTAs may clear the flag while this one runs
    __syncthreads();
    if (s_last == 0) return;
    __syncthreads();
    const uint32_t lo = ld_cg32(&st->lo_bin), span = ld_cg32(&st->span), hot = ld_cg32(&st->hot_bin);
    uint32_t below = 0u, hotc = 0u;
    uint32_t used = 0u;                                      // keys staged in this warp's buffer (uniform over the warp)
    bool dead = false;                                       // the candidate list overflowed: stop listing (valid is already 0)
    auto flush = [&]() {
        if (used == 0u) return;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&st->cand_count, (unsigned long long)used);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base + used > p.cand_cap) {
            if (lane == 0) st->valid = 0u;
            dead = true;
        } else {
            for (uint32_t i = lane; i < used; i += 32) p.cand[base + i] = s_buf[warp][i];
        }
        used = 0u;
        __syncwarp();
    };
    const int64_t n_tiles = (p.n_vec + kTileVecs - 1) / kTileVecs;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t tile_base = tile * kTileVecs;
        const int rem = (int)min((int64_t)kTileVecs, p.n_vec - tile_base);
        uint4 raw[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int li = tid + u * kT;
            raw[u] = (li < rem) ? ld_stream(p.in + tile_base + li) : make_uint4(0u, 0u, 0u, 0u);
        }
        uint32_t key[U * V];
        uint32_t cm = 0u;                                    // bit j: element j is a candidate
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int li = tid + u * kT;
            float v[V];
            unpack_vec<DT>(raw[u], v);
            if (MODE == MODE_QS) quantize_vec<DT, STOC>(v, p, tile_base + li);
            const bool real = li < rem;
#pragma unroll
            for (int e = 0; e < V; ++e) {
                const uint32_t kk = topk_key(v[e]);
                key[u * V + e] = kk;
                const uint32_t bin = kk >> 15, d = bin - lo;
                const bool inw = real && d <= span;
                below += (real && bin < lo) ? 1u : 0u;
                if (inw) {
                    if (bin == hot) ++hotc; else atomicAdd(&s_hist[d], 1u);
                    if (kk & 0x7fffu) cm |= 1u << (u * V + e);
                }
            }
        }
        const uint32_t ncand = __popc(cm);
        if (__any_sync(0xffffffffu, ncand != 0u) && !dead) {
            uint32_t incl = ncand;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= off) incl += o;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            if (used + total > (uint32_t)kWarpBuf) flush();
            if (total > (uint32_t)kWarpBuf) {                // more than half of a tile is in-window and not round: give up listing
                if (lane == 0) st->valid = 0u;
                dead = true;
            }
            if (!dead) {
                uint32_t pos = used + incl - ncand;
#pragma unroll
                for (int j = 0; j < U * V; ++j)
                    if (cm & (1u << j)) s_buf[warp][pos++] = key[j];
                used += total;
                __syncwarp();
            }
        }
    }
    if (!dead) flush();
    __syncthreads();
    for (int i = tid; i < kWin; i += kT)
        if (s_hist[i]) atomicAdd(&st->win_hist[i], (unsigned long long)s_hist[i]);
    const unsigned long long b_sum = block_sum<kT>(below, s_scr);
    const unsigned long long h_sum = block_sum<kT>(hotc, s_scr);
    if (tid == 0) {
        if (b_sum) atomicAdd(&st->below, b_sum);
        if (h_sum) atomicAdd(&st->win_hist[hot - lo], h_sum);
    }
    // the last CTA to arrive picks the bin
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(&st->done_a, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const unsigned long long bel = ld_cg64(&st->below);
    const unsigned long long need_w = p.k > bel ? p.k - bel : 0ull;
    const SelectResult r = block_select<kT, kWin / kT>(st->win_hist, kWin, need_w, 0ull, s_scr, &s_res);
    if (tid == 0) {
        if (!r.found) st->valid = 0u;
        else { st->bin = lo + r.bin; st->need_bin = need_w - r.before; st->cnt_bin = r.count; }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// 3. refine (cooperative launch): candidates of the selected bin -> tau; or, when the two-read path does not hold, a radix
//    select over the whole tensor with grid-wide barriers
// ---------------------------------------------------------------------------------------------------------------
template <int DT, int MODE, bool STOC>
__global__ void __launch_bounds__(kT) refine_kernel(const UParams p) {
    using D = DType<DT>;
    constexpr int V = D::kVec;
    constexpr int U = TileCfg<DT>::kU;
    constexpr int kTileVecs = kT * U;
    __shared__ unsigned int s_fb[kFbBins];
    __shared__ unsigned long long s_scr[32];
    __shared__ SelectResult s_res;
    __shared__ int s_last;
    const int tid = threadIdx.x;
    FusedState* st = p.st;
    if (tid == 0) s_last = (int)ld_cg32(&st->valid);       // stable here: pass A has completed
    __syncthreads();
    const bool two_read_path = s_last != 0;
    __syncthreads();
    if (two_read_path) {
        const unsigned long long n_c = min(ld_cg64(&st->cand_count), p.cand_cap);
        const uint32_t bin = ld_cg32(&st->bin);
        unsigned long long impure = 0;
        for (unsigned long long i = (unsigned long long)blockIdx.x * kT + tid; i < n_c; i += (unsigned long long)gridDim.x * kT) {
            const uint32_t kk = __ldcg(p.cand + i);
            if ((kk >> 15) == bin) { atomicAdd(&st->low_hist[kk & 0x7fffu], 1ull); ++impure; }
        }
        const unsigned long long i_sum = block_sum<kT>(impure, s_scr);
        if (tid == 0 && i_sum) atomicAdd(&st->impure_in_bin, i_sum);
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = atomicAdd(&st->done_r, 1u) == gridDim.x - 1;
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        const unsigned long long cnt = ld_cg64(&st->cnt_bin), imp = ld_cg64(&st->impure_in_bin), need_bin = ld_cg64(&st->need_bin);
        const unsigned long long pure = cnt > imp ? cnt - imp : 0ull;        // keys of the bin with zero trailing bits
        const SelectResult r = block_select<kT, kLowBins / kT>(st->low_hist, kLowBins, need_bin, pure, s_scr, &s_res);
        if (tid == 0) {
            // r.found is guaranteed: need_bin <= cnt_bin = pure + impure
            st->tau = (bin << 15) | r.bin;
            st->need = need_bin - r.before;
            st->ties_total = r.count;
        }
        return;
    }
    // ---- fallback: 11 + 10 + 10-bit radix select over the whole tensor, every CTA of the grid takes part ----
    cg::grid_group grid = cg::this_grid();
    const int lane = tid & 31;
    (void)lane;
    const int shifts[3] = {20, 10, 0}, bits[3] = {11, 10, 10};
    const int64_t n_tiles = (p.n_vec + kTileVecs - 1) / kTileVecs;
    for (int pass = 0; pass < 3; ++pass) {
        for (int i = tid; i < kFbBins; i += kT) s_fb[i] = 0u;
        __syncthreads();
        const uint32_t pv = ld_cg32(&st->prefix_value), pm = ld_cg32(&st->prefix_mask), dm = (1u << bits[pass]) - 1u;
        const int sh = shifts[pass];
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
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
                    for (int e = 0; e < V; ++e) {
                        const uint32_t kk = topk_key(v[e]);
                        if ((kk & pm) == pv) atomicAdd(&s_fb[(kk >> sh) & dm], 1u);
                    }
                }
            }
        }
        __syncthreads();
        for (int i = tid; i < kFbBins; i += kT)
            if (s_fb[i]) atomicAdd(&st->fb_hist[pass][i], (unsigned long long)s_fb[i]);
        __threadfence();
        grid.sync();
        if (blockIdx.x == 0) {
            const unsigned long long need = pass == 0 ? p.k : ld_cg64(&st->fb_need);
            const SelectResult r = block_select<kT, kFbBins / kT>(st->fb_hist[pass], 1 << bits[pass], need, 0ull, s_scr, &s_res);
            if (tid == 0) {
                st->prefix_value = pv | (r.bin << sh);
                st->prefix_mask = pm | (dm << sh);
                st->fb_need = need - r.before;
                if (pass == 2) { st->tau = pv | (r.bin << sh); st->need = need - r.before; st->ties_total = r.count; }
            }
            __threadfence();
        }
        grid.sync();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// 4. apply
// ---------------------------------------------------------------------------------------------------------------
template <int DT, int MODE, bool STOC>
__global__ void __launch_bounds__(kT) apply_kernel(const UParams p) {
    pdl_launch_dependents();
    pdl_wait();
    using D = DType<DT>;
    constexpr int V = D::kVec;
    constexpr int U = TileCfg<DT>::kU;                        // 16 elements per thread per tile
    constexpr int kTileVecs = kT * U;
    constexpr int kOutVecs = (STOC && V == 8 && MODE != MODE_S_ONLY) ? 2 : 1;
    __shared__ unsigned int s_w[2][kWarps];
    __shared__ unsigned long long s_excl;
    __shared__ int64_t s_tile;
    FusedState* st = p.st;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tau = ld_cg32(&st->tau);
    const unsigned long long need = ld_cg64(&st->need), ties_total = ld_cg64(&st->ties_total);
    const bool ranked = need < ties_total;                    // only some of the keys equal to tau go: index order decides
    const uint32_t lim = tau + (ranked ? 0u : 1u);            // drop key < lim (keys <= 0x7fffffff: no overflow)
    const int64_t n_tiles = p.n_tiles;

    auto store_vec = [&](int64_t vec, const float* v) {
        if (kOutVecs == 1) {
            st_stream(p.out + vec, (STOC && MODE != MODE_S_ONLY) ? pack_vec<BFP_DT_F32>(v) : pack_vec<DT>(v));
        } else {
            st_stream(p.out + vec * 2, pack_vec<BFP_DT_F32>(v));
            st_stream(p.out + vec * 2 + 1, pack_vec<BFP_DT_F32>(v + (V == 8 ? 4 : 0)));
        }
    };

    if (!ranked) {
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t tile_base = tile * kTileVecs;
            const int rem = (int)min((int64_t)kTileVecs, p.n_vec - tile_base);
            uint4 raw[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int li = tid + u * kT;
                raw[u] = (li < rem) ? ld_stream(p.in + tile_base + li) : make_uint4(0u, 0u, 0u, 0u);
            }
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
        return;
    }

    // ranked ties: tiles in index order, chained scan of the per-tile tie counts with decoupled look-back
    while (true) {
        if (tid == 0) s_tile = (int64_t)atomicAdd(&st->tile_counter, 1u);
        __syncthreads();
        const int64_t tile = s_tile;
        if (tile >= n_tiles) break;
        const int64_t tile_base = tile * kTileVecs;
        const int rem = (int)min((int64_t)kTileVecs, p.n_vec - tile_base);
        float v[U][V];
        uint32_t tm[U];                                       // bit e: element e of vector u equals tau
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int li = tid + u * kT;
            const uint4 raw = (li < rem) ? ld_stream(p.in + tile_base + li) : make_uint4(0u, 0u, 0u, 0u);
            unpack_vec<DT>(raw, v[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int li = tid + u * kT;
            if (MODE == MODE_QS) quantize_vec<DT, STOC>(v[u], p, tile_base + li);
            tm[u] = 0u;
#pragma unroll
            for (int e = 0; e < V; ++e) tm[u] |= (li < rem && topk_key(v[u][e]) == tau) ? (1u << e) : 0u;
        }
        // packed scan of the four per-vector tie counts (field totals <= 256 * 8 < 65536); vector order inside the tile is u * kT + tid
        uint32_t c01 = __popc(tm[0]) | (__popc(tm[1]) << 16), c23 = U == 4 ? (__popc(tm[U - 2]) | (__popc(tm[U - 1]) << 16)) : 0u;
        uint32_t i01 = c01, i23 = c23;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t o01 = __shfl_up_sync(0xffffffffu, i01, off), o23 = __shfl_up_sync(0xffffffffu, i23, off);
            if (lane >= off) { i01 += o01; i23 += o23; }
        }
        if (lane == 31) { s_w[0][warp] = i01; s_w[1][warp] = i23; }
        __syncthreads();
        uint32_t wb01 = 0u, wb23 = 0u, t01 = 0u, t23 = 0u;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const uint32_t a = s_w[0][w], b = s_w[1][w];
            wb01 += w < warp ? a : 0u; wb23 += w < warp ? b : 0u; t01 += a; t23 += b;
        }
        const uint32_t tot[4] = {t01 & 0xffffu, t01 >> 16, t23 & 0xffffu, t23 >> 16};
        const uint32_t tile_total = tot[0] + tot[1] + tot[2] + tot[3];
        const uint32_t ex01 = wb01 + i01 - c01, ex23 = wb23 + i23 - c23;      // exclusive prefixes within each field
        const uint32_t exu[4] = {ex01 & 0xffffu, ex01 >> 16, ex23 & 0xffffu, ex23 >> 16};
        // look-back (warp 0)
        if (warp == 0) {
            unsigned long long excl = 0;
            if (tile == 0) {
                if (lane == 0) st_volatile64(p.tile_state + tile, (2ull << 62) | (unsigned long long)tile_total);
            } else {
                if (lane == 0) st_volatile64(p.tile_state + tile, (1ull << 62) | (unsigned long long)tile_total);
                int64_t j = tile - 1;
                while (true) {
                    const int64_t idx = j - lane;
                    const unsigned long long s = idx >= 0 ? ld_volatile64(p.tile_state + idx) : (2ull << 62);
                    const uint32_t status = (uint32_t)(s >> 62);
                    const uint32_t ready = __ballot_sync(0xffffffffu, status != 0u), pref = __ballot_sync(0xffffffffu, status == 2u);
                    const unsigned long long val = s & ((1ull << 62) - 1ull);
                    if (pref) {
                        const int first = __ffs(pref) - 1;
                        const uint32_t upto = first == 31 ? 0xffffffffu : ((2u << first) - 1u);
                        if ((ready & upto) == upto) {
                            unsigned long long part = lane <= first ? val : 0ull;
#pragma unroll
                            for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
                            excl += part;
                            break;
                        }
                    } else if (ready == 0xffffffffu) {
                        unsigned long long part = val;
#pragma unroll
                        for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
                        excl += part;
                        j -= 32;
                    }
                }
                if (lane == 0) st_volatile64(p.tile_state + tile, (2ull << 62) | (excl + (unsigned long long)tile_total));
            }
            if (lane == 0) s_excl = excl;
        }
        __syncthreads();
        const unsigned long long tile_excl = s_excl;
        uint32_t before_u = 0u;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int li = tid + u * kT;
            unsigned long long rank = tile_excl + before_u + exu[u];
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
    }
}

template <class Kernel>
static cudaError_t launch_one(Kernel kernel, int grid, int threads, size_t smem, cudaStream_t s, bool pdl, bool cooperative, const UParams& p) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)threads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    if (cooperative) { attr[0].id = cudaLaunchAttributeCooperative; attr[0].val.cooperative = 1; }
    else { attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0; }
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, p);
}

struct WsLayout { size_t state, tiles, cand, total; unsigned long long cap; };
static WsLayout ws_layout(int64_t n, int dtype) {
    const int V = dtype == BFP_DT_F32 ? 4 : 8;
    const int64_t n_vec = (n + V - 1) / V;
    WsLayout w;
    w.state = (sizeof(FusedState) + 255) / 256 * 256;
    w.tiles = (size_t)((n_vec + 511) / 512 * 8 + 255) / 256 * 256;
    w.cap = (unsigned long long)std::max<int64_t>(65536, n / 8);
    w.cand = (size_t)(w.cap * 4 + 255) / 256 * 256;
    w.total = w.state + w.tiles + w.cand;
    return w;
}

template <int DT, int MODE, bool STOC>
int run_fused(const UnstructuredArgs& a, cudaStream_t s) {
    constexpr int V = DType<DT>::kVec;
    const int64_t n = a.n;
    const WsLayout w = ws_layout(n, DT);
    UParams p = {};
    p.in = static_cast<const uint4*>(a.in);
    p.out = static_cast<uint4*>(a.out);
    p.n_vec = n / V;
    p.n_tiles = (p.n_vec + kT * TileCfg<DT>::kU - 1) / (kT * TileCfg<DT>::kU);      // apply_kernel's tiles
    p.k = a.k; p.n = (unsigned long long)n;
    p.lanes_per_block = MODE == MODE_S_ONLY ? 1 : a.B / V;
    p.m = a.m; p.eps = a.eps; p.seed = a.seed; p.offset = a.offset;
    char* base = static_cast<char*>(a.workspace);
    p.st = reinterpret_cast<FusedState*>(base);
    p.tile_state = reinterpret_cast<unsigned long long*>(base + w.state);
    p.cand = reinterpret_cast<uint32_t*>(base + w.state + w.tiles);
    p.cand_cap = w.cap;
    p.force_fallback = tuning().unstructured_force_fallback;
    const int sms = device_info().sm_count;
    const bool pdl = tuning().pdl != 0;

    auto k_sample = sample_kernel<DT, MODE, STOC>;
    auto k_a = pass_a_kernel<DT, MODE, STOC>;
    auto k_r = refine_kernel<DT, MODE, STOC>;
    auto k_ap = apply_kernel<DT, MODE, STOC>;
    static const bool dbg = getenv("BFP_UNSTRUCTURED_TIMING") != nullptr;      // per-phase device times on stderr (tools only)
    cudaEvent_t ev[5]; int nev = 0;
    auto mark = [&] { if (dbg) { cudaEventCreate(&ev[nev]); cudaEventRecord(ev[nev], s); ++nev; } };
    mark();
    // per device, once per instantiation: the sample kernel's shared-memory opt-in and the resident CTAs per SM of the three grids
    struct PerDevice { bool init = false; int occ_a = 0, occ_r = 0, occ_ap = 0; };
    static PerDevice per_device[64];
    PerDevice& pd = per_device[std::max(0, std::min(63, device_info().device))];
    if (!pd.init) {
        cudaFuncSetAttribute(k_sample, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
        pd.occ_a = kernel_occupancy(k_a, kT); pd.occ_r = kernel_occupancy(k_r, kT); pd.occ_ap = kernel_occupancy(k_ap, kT);
        pd.init = true;
    }
    cudaError_t e = launch_one(k_sample, 1, kSampleThreads, 131072, s, pdl, false, p);
    if (e != cudaSuccess) return set_errorf(BFP_E_CUDA, "unstructured sample kernel: %s", cudaGetErrorString(e));
    count_launch();
    mark();
    constexpr int UA = TileCfg<DT>::kU;
    const int64_t tiles_a = (p.n_vec + kT * UA - 1) / (kT * UA);
    const int grid_a = (int)std::max<int64_t>(1, std::min<int64_t>(tiles_a, (int64_t)sms * pd.occ_a));
    e = launch_one(k_a, grid_a, kT, 0, s, pdl, false, p);
    if (e != cudaSuccess) return set_errorf(BFP_E_CUDA, "unstructured pass A: %s", cudaGetErrorString(e));
    count_launch();
    mark();
    const int grid_r = (int)std::max<int64_t>(1, std::min<int64_t>(tiles_a, (int64_t)sms * std::min(pd.occ_r, 4)));
    e = launch_one(k_r, grid_r, kT, 0, s, false, true, p);
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
        fprintf(stderr, "| valid %u window [%u, +%u] hot %d below %llu cand %llu bin %u need %llu ties %llu grids %d/%d/%d\n", h.valid, h.lo_bin, h.span, (int)h.hot_bin,
                h.below, h.cand_count, h.bin, h.need, h.ties_total, grid_a, grid_r, grid_ap);
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

size_t unstructured_fused_workspace_bytes(int64_t n, int dtype) { return ws_layout(std::max<int64_t>(n, 0), dtype).total; }

// which calls the two-read pipeline takes; everything else composes bfp_unstructured_sparsify and bfp_quantize
bool unstructured_fused_supported(const UnstructuredArgs& a) {
    const int V = a.dtype == BFP_DT_F32 ? 4 : 8;
    if (a.n <= 0 || a.n % V) return false;
    if (reinterpret_cast<uintptr_t>(a.in) % 16 || reinterpret_cast<uintptr_t>(a.out) % 16 || reinterpret_cast<uintptr_t>(a.workspace) % 16) return false;
    if (a.order == BFP_ORDER_SPARSIFY_ONLY) return true;
    const int B = a.B;
    if (B < V || (B & (B - 1)) || B / V > 32) return false;          // a block = 2^j adjacent lanes of a warp
    if (a.K <= 0 || a.K % B) return false;
    return true;
}

int unstructured_fused_device(const UnstructuredArgs& a, cudaStream_t s) {
    if (a.k == 0 || a.k >= (unsigned long long)a.n) return set_error(BFP_E_ARG, "k must be in (0, numel)");
    switch (a.dtype) {
    case BFP_DT_F32: return dispatch_mode<BFP_DT_F32>(a, s);
    case BFP_DT_F16: return dispatch_mode<BFP_DT_F16>(a, s);
    case BFP_DT_BF16: return dispatch_mode<BFP_DT_BF16>(a, s);
    }
    return set_error(BFP_E_ARG, "bad dtype");
}

}  // namespace bfp
