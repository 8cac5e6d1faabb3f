// bfp_unstructured.cu -- global magnitude pruning.
//
// Replaces _unstructured_sparsity (bfp_ops.py:61-71): view the tensor as one row, zero the k = int(numel * frac)
// entries torch.topk(|t|, k, largest=False) returns.  torch-CUDA semantics: everything strictly below the k-th smallest
// magnitude tau is dropped, and of the entries equal to tau the first (k - #below) in index order; NaN is largest.
//
// HBM-bound multi-pass radix select on the 31-bit key |x| (bit pattern of the fp32 value; monotone for fp16/bf16 too):
//   3 x histogram pass (11 + 10 + 10 bits, shared-memory histograms, prefix carried in device memory: no host sync)
//   1 x tie-count pass  (per-CTA contiguous range: how many keys == tau)
//   1 x apply pass      (drop key < tau, and key == tau while the running tie rank < need)
// = 5 reads + 1 write (24 B/element fp32) against 8 B/element algorithmic; the reference spends 20.8 ms on a 4096x4096
// tensor on the same GPU (profiles/r01_probe_ref_gpu.log).
#include <algorithm>

#include "bfp_internal.h"
#include "bfp_stream.cuh"

namespace bfp {

namespace {
constexpr int kBins = 2048;
constexpr int kThreadsU = 256;
constexpr int kMaxCtasU = 1024;

struct SelectState {          // lives in the caller's workspace
    uint32_t prefix_value;    // bits of tau found so far
    uint32_t prefix_mask;     // which bits are fixed
    unsigned long long need;  // how many more (smallest) keys to take inside the current prefix bucket
    unsigned long long hist[3][kBins];
    unsigned long long tie_count[kMaxCtasU];
};

template <int DT>
__device__ __forceinline__ uint32_t key_at(const void* in, int64_t i) { return abs_bits(DType<DT>::load(in, i)); }

template <int DT>
__global__ void __launch_bounds__(kThreadsU) hist_kernel(const void* in, int64_t n, SelectState* st, int pass, int shift, int bits) {
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

// one block: find the digit whose bucket contains the need-th smallest key, fix it into the prefix
__global__ void __launch_bounds__(1024) select_kernel(SelectState* st, int pass, int shift, int bits, unsigned long long k_init) {
    __shared__ unsigned long long cum[kBins];
    const int nb = 1 << bits;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) cum[i] = st->hist[pass][i];
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long need = pass == 0 ? k_init : st->need, below = 0;
        int d = 0;
        for (; d < nb - 1; ++d) {
            if (below + cum[d] >= need) break;
            below += cum[d];
        }
        st->prefix_value |= (uint32_t)d << shift;
        st->prefix_mask |= ((1u << bits) - 1u) << shift;
        st->need = need - below;            // >= 1: rank of tau inside its bucket
    }
}

// contiguous range of CTA b: [b * per, min(n, (b + 1) * per)), per a multiple of the tile so ranges align with tiles
template <int DT>
__global__ void __launch_bounds__(kThreadsU) tie_count_kernel(const void* in, int64_t n, int64_t per, SelectState* st) {
    const uint32_t tau = st->prefix_value;
    const int64_t lo = (int64_t)blockIdx.x * per, hi = min(n, lo + per);
    unsigned int c = 0;
    for (int64_t i = lo + threadIdx.x; i < hi; i += kThreadsU) c += key_at<DT>(in, i) == tau;
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
    using D = DType<DT>;
    const uint32_t tau = st->prefix_value;
    const unsigned long long need = st->need;
    __shared__ unsigned long long s_before;
    __shared__ unsigned int s_warp[kThreadsU / 32];
    if (threadIdx.x == 0) {
        unsigned long long b = 0;
        for (int i = 0; i < (int)blockIdx.x; ++i) b += st->tie_count[i];
        s_before = b;
    }
    __syncthreads();
    unsigned long long before = s_before;                   // ties == tau in all earlier indices
    const bool range_has_ties = st->tie_count[blockIdx.x] != 0;
    const int64_t lo = (int64_t)blockIdx.x * per, hi = min(n, lo + per);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = lo; base < hi; base += kThreadsU) {
        const int64_t i = base + threadIdx.x;
        const float v = i < hi ? D::load(in, i) : 0.0f;
        const uint32_t k = abs_bits(v);
        bool drop = i < hi && k < tau;
        if (range_has_ties) {                               // block-uniform branch: index-ordered rank among the ties
            const bool tie = i < hi && k == tau;
            const unsigned int bal = __ballot_sync(0xffffffffu, tie);
            if (lane == 0) s_warp[warp] = __popc(bal);
            __syncthreads();
            unsigned int wbefore = 0, total = 0;
            for (int w = 0; w < kThreadsU / 32; ++w) { const unsigned int c = s_warp[w]; wbefore += w < warp ? c : 0; total += c; }
            const unsigned long long rank = before + wbefore + __popc(bal & ((1u << lane) - 1u));
            if (tie && rank < need) drop = true;
            before += total;
            __syncthreads();
        }
        if (i < hi) D::store(out, i, drop ? 0.0f : v);
    }
}

template <int DT>
int run(const void* in, void* out, int64_t n, unsigned long long k, SelectState* st, cudaStream_t s) {
    const int sms = device_info().sm_count;
    cudaError_t e = cudaMemsetAsync(st, 0, sizeof(SelectState), s);
    if (e != cudaSuccess) return set_errorf(BFP_E_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    const int grid = (int)std::min<int64_t>((n + kThreadsU - 1) / kThreadsU, (int64_t)sms * 8);
    const int shifts[3] = {20, 10, 0}, bits[3] = {11, 10, 10};
    for (int p = 0; p < 3; ++p) {
        hist_kernel<DT><<<grid, kThreadsU, 0, s>>>(in, n, st, p, shifts[p], bits[p]);
        count_launch();
        select_kernel<<<1, 1024, 0, s>>>(st, p, shifts[p], bits[p], k);
        count_launch();
    }
    const int ctas = (int)std::min<int64_t>(kMaxCtasU, std::max<int64_t>(1, std::min<int64_t>((n + kThreadsU - 1) / kThreadsU, (int64_t)sms * 4)));
    int64_t per = (n + ctas - 1) / ctas;
    per = (per + kThreadsU - 1) / kThreadsU * kThreadsU;
    const int ctas_used = (int)((n + per - 1) / per);
    tie_count_kernel<DT><<<ctas_used, kThreadsU, 0, s>>>(in, n, per, st);
    count_launch();
    apply_kernel<DT><<<ctas_used, kThreadsU, 0, s>>>(in, out, n, per, st);
    count_launch();
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
