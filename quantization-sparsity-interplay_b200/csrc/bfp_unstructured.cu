// bfp_unstructured.cu -- global magnitude pruning.
//
// Replaces _unstructured_sparsity (bfp_ops.py:61-71): view the tensor as one row, zero the k = int(numel * frac)
// entries torch.topk(|t|, k, largest=False) returns.  torch-CUDA semantics: everything strictly below the k-th smallest
// magnitude tau is dropped, and of the entries equal to tau the first (k - #below) in index order; NaN is largest.
//
// HBM-bound multi-pass radix select on the 31-bit key |x| (bit pattern of the fp32 value; monotone for fp16/bf16 too):
//   3 x histogram pass (8 + 12 + 11 bits; 128-bit loads; pass 0 on lane-private histograms, later passes warp-aggregated; parallel-scan select,
//                       prefix carried in device memory: no host sync)
//   1 x apply pass      (128-bit loads / stores; drop key <= tau when every key equal to tau goes -- always true for a
//                       unique tau -- else a tie-count pass ranks the ties in index order first)
// = 4 reads + 1 write (20 B/element fp32) against 8 B/element algorithmic; the reference spends 20.8 ms on a 4096x4096
// tensor on the same GPU (profiles/r01_probe_ref_gpu.log).
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "bfp_internal.h"
#include "bfp_stream.cuh"

namespace bfp {

namespace {
constexpr int kBins = 4096;                // largest digit: 12 bits
constexpr int kThreadsU = 256;
constexpr int kMaxCtasU = 1024;

struct SelectState {          // lives in the caller's workspace
    uint32_t prefix_value;    // bits of tau found so far
    uint32_t prefix_mask;     // which bits are fixed
    unsigned long long need;  // how many more (smallest) keys to take inside the current prefix bucket
    unsigned long long ties_total;   // after the last pass: how many keys equal tau (need <= ties_total)
    unsigned long long hist[3][kBins];
    unsigned long long tie_count[kMaxCtasU];
};

template <int DT>
__device__ __forceinline__ uint32_t key_at(const void* in, int64_t i) { return topk_key(DType<DT>::load(in, i)); }

// keys of one 128-bit vector (4 fp32 / 8 half values)
template <int DT>
__device__ __forceinline__ void keys_of(const uint4& raw, uint32_t* k) {
    float v[DType<DT>::kVec];
    unpack_vec<DT>(raw, v);
#pragma unroll
    for (int i = 0; i < DType<DT>::kVec; ++i) k[i] = topk_key(v[i]);
}

template <int DT>
__global__ void __launch_bounds__(kThreadsU) hist_scalar_kernel(const void* in, int64_t n, SelectState* st, int pass, int shift, int bits) {
    pdl_launch_dependents();   // launch overlap only: pdl_wait() orders this kernel after everything earlier on the stream
    pdl_wait();
    __shared__ unsigned int sh[kBins];
    for (int i = threadIdx.x; i < kBins; i += kThreadsU) sh[i] = 0;
    __syncthreads();
    const uint32_t pv = st->prefix_value, pm = st->prefix_mask, dm = (1u << bits) - 1u;
    for (int64_t i = (int64_t)blockIdx.x * kThreadsU + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreadsU) {
        const uint32_t k = key_at<DT>(in, i);
        if ((k & pm) == pv) atomicAdd(&sh[(k >> shift) & dm], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kBins; i += kThreadsU)
        if (sh[i]) atomicAdd(&st->hist[pass][i], (unsigned long long)sh[i]);
}

template <int DT>
__global__ void __launch_bounds__(kThreadsU) hist_kernel(const void* in, int64_t n, SelectState* st, int pass, int shift, int bits) {
    pdl_launch_dependents();   // launch overlap only: pdl_wait() orders this kernel after everything earlier on the stream
    pdl_wait();
    constexpr int V = DType<DT>::kVec;
    __shared__ unsigned int sh[kBins];
    for (int i = threadIdx.x; i < kBins; i += kThreadsU) sh[i] = 0;
    __syncthreads();
    const uint32_t pv = st->prefix_value, pm = st->prefix_mask, dm = (1u << bits) - 1u;
    const int64_t n_vec = n / V;
    const uint4* src = static_cast<const uint4*>(in);
    // whole warps iterate together (the aggregation uses warp-wide votes): the loop bound is rounded up per warp
    const int64_t stride = (int64_t)gridDim.x * kThreadsU;
    for (int64_t base = (int64_t)blockIdx.x * kThreadsU + (threadIdx.x & ~31); base < n_vec; base += stride) {
        const int64_t i = base + (threadIdx.x & 31);
        const bool in_range = i < n_vec;
        uint32_t k[V];
        if (in_range) keys_of<DT>(ld_stream(src + i), k);
        if (in_range) {
#pragma unroll
            for (int j = 0; j < V; ++j)
                if ((k[j] & pm) == pv) atomicAdd(&sh[(k[j] >> shift) & dm], 1u);   // mantissa digits: spread out, few conflicts
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < 32) {               // ragged tail (n not a multiple of the vector width)
        const int64_t i = n_vec * V + threadIdx.x;
        if (i < n) {
            const uint32_t k = key_at<DT>(in, i);
            if ((k & pm) == pv) atomicAdd(&sh[(k >> shift) & dm], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kBins; i += kThreadsU)
        if (sh[i]) atomicAdd(&st->hist[pass][i], (unsigned long long)sh[i]);
}

// First pass (digit = the 8 exponent bits, every element counts): the magnitudes of a weight tensor sit in a handful of
// exponent bins, so a plain shared histogram serialises on them (and match.any aggregation is slower still: 247 us for 45 M
// elements).  Here every LANE owns a private copy of the 256 bins (address = bin * 32 + lane: bank = lane), so the 32
// increments of a warp instruction never conflict; the copies are summed at the end.
template <int DT>
__global__ void __launch_bounds__(kThreadsU) hist0_kernel(const void* in, int64_t n, SelectState* st) {
    pdl_launch_dependents();   // launch overlap only: pdl_wait() orders this kernel after everything earlier on the stream
    pdl_wait();
    constexpr int V = DType<DT>::kVec;
    __shared__ unsigned int sh[256 * 32];
    for (int i = threadIdx.x; i < 256 * 32; i += kThreadsU) sh[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t n_vec = n / V;
    const uint4* src = static_cast<const uint4*>(in);
    for (int64_t i = (int64_t)blockIdx.x * kThreadsU + threadIdx.x; i < n_vec; i += (int64_t)gridDim.x * kThreadsU) {
        uint32_t k[V];
        keys_of<DT>(ld_stream(src + i), k);
#pragma unroll
        for (int j = 0; j < V; ++j) atomicAdd(&sh[(k[j] >> 23) * 32 + lane], 1u);
    }
    if (blockIdx.x == 0) {
        for (int64_t i = n_vec * V + threadIdx.x; i < n; i += kThreadsU) atomicAdd(&sh[(key_at<DT>(in, i) >> 23) * 32 + lane], 1u);
    }
    __syncthreads();
    {   // thread t sums bin t over the 32 lane copies (rotated start: conflict-free)
        unsigned int c = 0;
#pragma unroll 8
        for (int l = 0; l < 32; ++l) c += sh[threadIdx.x * 32 + ((l + lane) & 31)];
        if (c) atomicAdd(&st->hist[0][threadIdx.x], (unsigned long long)c);
    }
}

// one block of 1024 threads, four bins each: inclusive scan of the histogram, then the (unique) bin whose cumulative count
// first reaches `need` fixes its digit into the prefix.  After the last pass it also records how many keys equal tau.
__global__ void __launch_bounds__(1024) select_kernel(SelectState* st, int pass, int shift, int bits, unsigned long long k_init, int last) {
    pdl_launch_dependents();   // launch overlap only: pdl_wait() orders this kernel after everything earlier on the stream
    pdl_wait();
    __shared__ unsigned long long warp_sum[32];
    const int nb = 1 << bits, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const unsigned long long need = pass == 0 ? k_init : st->need;
    unsigned long long c[4], incl = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) { c[j] = 4 * t + j < nb ? st->hist[pass][4 * t + j] : 0ull; incl += c[j]; }
    const unsigned long long mine = incl;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned long long o = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += o;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = warp_sum[lane];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned long long o = __shfl_up_sync(0xffffffffu, w, off);
            if (lane >= off) w += o;
        }
        warp_sum[lane] = w;
    }
    __syncthreads();
    incl += warp ? warp_sum[warp - 1] : 0ull;                 // keys in bins [0, 4t+3]
    // bin d is selected iff before(d) < need <= before(d) + count(d): exactly one bin of one thread
    unsigned long long before = incl - mine;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (4 * t + j < nb && before < need && need <= before + c[j]) {
            st->prefix_value |= (uint32_t)(4 * t + j) << shift;
            st->prefix_mask |= ((1u << bits) - 1u) << shift;
            st->need = need - before;                         // >= 1: rank of tau inside its bucket
            if (last) st->ties_total = c[j];                  // keys == tau
        }
        before += c[j];
    }
}

// contiguous range of CTA b: [b * per, min(n, (b + 1) * per)), per a multiple of the tile so ranges align with tiles.
// Only needed when SOME but not all of the keys equal to tau are dropped (index order then decides); otherwise a no-op.
template <int DT>
__global__ void __launch_bounds__(kThreadsU) tie_count_kernel(const void* in, int64_t n, int64_t per, SelectState* st) {
    pdl_launch_dependents();   // launch overlap only: pdl_wait() orders this kernel after everything earlier on the stream
    pdl_wait();
    if (st->need == st->ties_total) return;
    using D = DType<DT>;
    constexpr int V = D::kVec;
    const uint32_t tau = st->prefix_value;
    const int64_t lo = (int64_t)blockIdx.x * per, hi = min(n, lo + per);
    unsigned int c = 0;
    int64_t i0 = lo;
    if (reinterpret_cast<uintptr_t>(in) % 16 == 0) {          // lo is a multiple of the vector width
        const int64_t nv = (hi - lo) / V;
        const uint4* src = reinterpret_cast<const uint4*>(static_cast<const char*>(in) + lo * sizeof(typename D::T));
        for (int64_t j = threadIdx.x; j < nv; j += kThreadsU) {
            uint32_t k[V];
            keys_of<DT>(ld_stream(src + j), k);
#pragma unroll
            for (int e = 0; e < V; ++e) c += k[e] == tau;
        }
        i0 = lo + nv * V;
    }
    for (int64_t i = i0 + threadIdx.x; i < hi; i += kThreadsU) c += key_at<DT>(in, i) == tau;
    __shared__ unsigned int sh;
    if (threadIdx.x == 0) sh = 0;
    __syncthreads();
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&sh, c);
    __syncthreads();
    if (threadIdx.x == 0) st->tie_count[blockIdx.x] = sh;
}

template <int DT>
__global__ void __launch_bounds__(kThreadsU) apply_kernel(const void* in, void* out, int64_t n, int64_t per, const SelectState* st) {
    pdl_launch_dependents();   // launch overlap only: pdl_wait() orders this kernel after everything earlier on the stream
    pdl_wait();
    using D = DType<DT>;
    constexpr int V = D::kVec;
    const uint32_t tau = st->prefix_value;
    const unsigned long long need = st->need;
    const int64_t lo = (int64_t)blockIdx.x * per, hi = min(n, lo + per);
    const bool all_ties_go = need == st->ties_total;
    if (all_ties_go || st->tie_count[blockIdx.x] == 0) {
        // a pure threshold, 128-bit loads / stores: either every key equal to tau is dropped (always the case when tau is
        // unique) or this CTA's range holds no key equal to tau at all (only the few ranges that do rank them in index order)
        const uint32_t lim = tau + (all_ties_go ? 1u : 0u);   // drop key < lim (keys are <= 0x7fffffff: no overflow)
        const bool vec_ok = (reinterpret_cast<uintptr_t>(in) % 16 == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0) && (lo % V == 0);
        int64_t i0 = lo;
        if (vec_ok) {
            const int64_t nv = (hi - lo) / V;
            const uint4* src = reinterpret_cast<const uint4*>(static_cast<const char*>(in) + lo * sizeof(typename D::T));
            uint4* dst = reinterpret_cast<uint4*>(static_cast<char*>(out) + lo * sizeof(typename D::T));
            for (int64_t j = threadIdx.x; j < nv; j += kThreadsU) {
                float v[V];
                unpack_vec<DT>(ld_stream(src + j), v);
#pragma unroll
                for (int e = 0; e < V; ++e) v[e] = topk_key(v[e]) < lim ? 0.0f : v[e];
                st_stream(dst + j, pack_vec<DT>(v));
            }
            i0 = lo + nv * V;
        }
        for (int64_t i = i0 + threadIdx.x; i < hi; i += kThreadsU) {
            const float v = D::load(in, i);
            D::store(out, i, topk_key(v) < lim ? 0.0f : v);
        }
        return;
    }
    // This CTA's range holds keys equal to tau and only the first `need` of them (in index order, over the whole tensor) go.
    // Tiles of 256 x V elements, thread t owning V consecutive ones; a tile without ties costs one barrier, a tile with ties a
    // block scan of the per-thread tie counts.
    __shared__ unsigned long long s_before;
    __shared__ unsigned int s_warp[kThreadsU / 32];
    if (threadIdx.x == 0) {
        unsigned long long b = 0;
        for (int i = 0; i < (int)blockIdx.x; ++i) b += st->tie_count[i];
        s_before = b;
    }
    __syncthreads();
    unsigned long long before = s_before;                   // ties == tau in all earlier indices
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(in) % 16 == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0);
    for (int64_t base = lo; base < hi; base += (int64_t)kThreadsU * V) {
        const int64_t i0 = base + (int64_t)threadIdx.x * V;
        float v[V];
        const bool full = vec_ok && i0 + V <= hi;
        if (full) {
            unpack_vec<DT>(ld_stream(reinterpret_cast<const uint4*>(static_cast<const char*>(in) + i0 * sizeof(typename D::T))), v);
        } else {
#pragma unroll
            for (int e = 0; e < V; ++e) v[e] = i0 + e < hi ? D::load(in, i0 + e) : 0.0f;
        }
        unsigned int ties = 0;                               // bit e: element e equals tau
#pragma unroll
        for (int e = 0; e < V; ++e) ties |= (i0 + e < hi && topk_key(v[e]) == tau) ? (1u << e) : 0u;
        if (__syncthreads_or(ties != 0u)) {
            // exclusive prefix of the tie counts over the threads of the tile (thread order = index order)
            const unsigned int cnt = __popc(ties);
            unsigned int incl = cnt;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned int o = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= off) incl += o;
            }
            if (lane == 31) s_warp[warp] = incl;
            __syncthreads();
            unsigned int wbefore = 0, total = 0;
            for (int w = 0; w < kThreadsU / 32; ++w) { const unsigned int c = s_warp[w]; wbefore += w < warp ? c : 0; total += c; }
            unsigned long long rank = before + wbefore + (incl - cnt);
#pragma unroll
            for (int e = 0; e < V; ++e) {
                if (ties & (1u << e)) { if (rank < need) v[e] = 0.0f; ++rank; }
            }
            before += total;
            __syncthreads();                                  // s_warp is reused by the next tile with ties
        }
#pragma unroll
        for (int e = 0; e < V; ++e) v[e] = (topk_key(v[e]) < tau) ? 0.0f : v[e];
        if (full) {
            st_stream(reinterpret_cast<uint4*>(static_cast<char*>(out) + i0 * sizeof(typename D::T)), pack_vec<DT>(v));
        } else {
#pragma unroll
            for (int e = 0; e < V; ++e)
                if (i0 + e < hi) D::store(out, i0 + e, v[e]);
        }
    }
}

// every kernel of the pipeline is launched with programmatic stream serialization: its launch latency hides behind the
// predecessor, and griddepcontrol.wait at its top keeps the data dependence.  Measured (tools/ab_unstructured_pdl.py): 97 -> 85 us
// at 16 M elements, neutral at 45 M, 10 % SLOWER at 180 M (early-resident dependents take SM slots from the tail of a long
// kernel), so only tensors up to 32 M elements use it.
template <class... KArgs, class... Args>
static cudaError_t launch_u(bool overlap, void (*kernel)(KArgs...), int grid, int threads, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)threads); cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = overlap ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <int DT>
int run(const void* in, void* out, int64_t n, unsigned long long k, SelectState* st, cudaStream_t s) {
    const int sms = device_info().sm_count;
    cudaError_t e = cudaMemsetAsync(st, 0, sizeof(SelectState), s);
    if (e != cudaSuccess) return set_errorf(BFP_E_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    const bool aligned = reinterpret_cast<uintptr_t>(in) % 16 == 0;
    const int64_t n_hist = aligned ? n : 0;                  // unaligned input: everything goes through the scalar tail path below
    const int64_t vecs = n_hist / DType<DT>::kVec;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((vecs + kThreadsU - 1) / kThreadsU, (int64_t)sms * 8));
    const int shifts[3] = {23, 11, 0}, bits[3] = {8, 12, 11};
    const bool pdl = tuning().pdl && n <= (int64_t(32) << 20);
    static const bool dbg = getenv("BFP_UNSTRUCTURED_TIMING") != nullptr;      // per-phase device times on stderr (tools only)
    cudaEvent_t ev[9]; int nev = 0;
    auto mark = [&] { if (dbg) { cudaEventCreate(&ev[nev]); cudaEventRecord(ev[nev], s); ++nev; } };
    mark();
    for (int p = 0; p < 3; ++p) {
        if (aligned && p == 0) launch_u(pdl, hist0_kernel<DT>, grid, kThreadsU, s, in, n, st);
        else if (aligned) launch_u(pdl, hist_kernel<DT>, grid, kThreadsU, s, in, n, st, p, shifts[p], bits[p]);
        else launch_u(pdl, hist_scalar_kernel<DT>, (int)std::min<int64_t>((n + kThreadsU - 1) / kThreadsU, (int64_t)sms * 8), kThreadsU, s, in, n, st, p, shifts[p], bits[p]);
        count_launch();
        mark();
        launch_u(pdl, select_kernel, 1, 1024, s, st, p, shifts[p], bits[p], k, (int)(p == 2));
        count_launch();
        mark();
    }
    const int ctas = (int)std::min<int64_t>(kMaxCtasU, std::max<int64_t>(1, std::min<int64_t>((n + kThreadsU - 1) / kThreadsU, (int64_t)sms * 4)));
    int64_t per = (n + ctas - 1) / ctas;
    per = (per + kThreadsU * 8 - 1) / (kThreadsU * 8) * (kThreadsU * 8);      // multiple of the tile and of the vector width
    const int ctas_used = (int)((n + per - 1) / per);
    launch_u(pdl, tie_count_kernel<DT>, ctas_used, kThreadsU, s, in, n, per, st);
    count_launch();
    mark();
    launch_u(pdl, apply_kernel<DT>, ctas_used, kThreadsU, s, in, out, n, per, (const SelectState*)st);
    count_launch();
    mark();
    if (dbg) {
        cudaStreamSynchronize(s);
        static const char* names[8] = {"hist0", "select0", "hist1", "select1", "hist2", "select2", "tie_count", "apply"};
        for (int i = 0; i + 1 < nev; ++i) { float ms = 0; cudaEventElapsedTime(&ms, ev[i], ev[i + 1]); fprintf(stderr, "%s %.1f us  ", names[i], ms * 1e3f); }
        fprintf(stderr, "\n");
        for (int i = 0; i < nev; ++i) cudaEventDestroy(ev[i]);
    }
    return check_launch("unstructured sparsity kernels");
}
}  // namespace

size_t unstructured_workspace_bytes() { return sizeof(SelectState); }

int unstructured_device(const void* in, void* out, int64_t n, int dtype, unsigned long long k, void* workspace, cudaStream_t s) {
    if (n == 0) return BFP_OK;
    SelectState* st = static_cast<SelectState*>(workspace);
    if (k == 0 || k >= (unsigned long long)n) {      // nothing / everything dropped: no selection needed
        const size_t bytes = (size_t)n * dtype_size(dtype);
        cudaError_t e = k == 0 ? (in == out ? cudaSuccess : cudaMemcpyAsync(out, in, bytes, cudaMemcpyDeviceToDevice, s))
                               : cudaMemsetAsync(out, 0, bytes, s);
        if (e != cudaSuccess) return set_errorf(BFP_E_CUDA, "%s", cudaGetErrorString(e));
        return BFP_OK;
    }
    switch (dtype) {
    case BFP_DT_F32: return run<BFP_DT_F32>(in, out, n, k, st, s);
    case BFP_DT_F16: return run<BFP_DT_F16>(in, out, n, k, st, s);
    case BFP_DT_BF16: return run<BFP_DT_BF16>(in, out, n, k, st, s);
    }
    return set_error(BFP_E_ARG, "bad dtype");
}

}  // namespace bfp
