// bfp_host.cu -- the operator through HOST buffers: row-chunked, four in-order streams, so chunk i+1's H2D copy,
// chunk i's kernel and chunk i-1's D2H copy overlap (both PCIe directions busy).  Chunk sizes are TAPERED: they grow
// geometrically from host_chunk_min_bytes to host_chunk_bytes and shrink again towards the end, because the pipeline's
// fill (first H2D with nothing to overlap) and drain (last D2H) cost one chunk each -- small end chunks keep those
// short, large middle chunks keep the per-copy launch overhead negligible.  This is the end-to-end path bench.py
// times as `e2e`.
#include <algorithm>
#include <mutex>

#include "bfp_internal.h"

namespace bfp {

namespace {
constexpr int kSlots = 4;
struct Staging {
    int device = -1;
    size_t in_cap = 0, out_cap = 0;
    void* d_in[kSlots] = {};
    void* d_out[kSlots] = {};
    cudaStream_t stream[kSlots] = {};
};
Staging g_st;
std::mutex g_mu;

int cuda_fail(cudaError_t e, const char* what) { return set_errorf(BFP_E_CUDA, "%s: %s", what, cudaGetErrorString(e)); }

void release_locked() {
    for (int i = 0; i < kSlots; ++i) {
        if (g_st.d_in[i]) cudaFree(g_st.d_in[i]);
        if (g_st.d_out[i]) cudaFree(g_st.d_out[i]);
        if (g_st.stream[i]) cudaStreamDestroy(g_st.stream[i]);
        g_st.d_in[i] = g_st.d_out[i] = nullptr;
        g_st.stream[i] = nullptr;
    }
    g_st.in_cap = g_st.out_cap = 0;
    g_st.device = -1;
}

int ensure_locked(size_t in_bytes, size_t out_bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    if (g_st.device != dev || g_st.in_cap < in_bytes || g_st.out_cap < out_bytes) {
        if (g_st.device >= 0) { cudaSetDevice(g_st.device); release_locked(); cudaSetDevice(dev); }
        for (int i = 0; i < kSlots; ++i) {
            if ((e = cudaMalloc(&g_st.d_in[i], in_bytes)) != cudaSuccess) { release_locked(); return cuda_fail(e, "cudaMalloc(staging in)"); }
            if ((e = cudaMalloc(&g_st.d_out[i], out_bytes)) != cudaSuccess) { release_locked(); return cuda_fail(e, "cudaMalloc(staging out)"); }
            if ((e = cudaStreamCreateWithFlags(&g_st.stream[i], cudaStreamNonBlocking)) != cudaSuccess) { release_locked(); return cuda_fail(e, "cudaStreamCreate"); }
        }
        g_st.in_cap = in_bytes; g_st.out_cap = out_bytes; g_st.device = dev;
    }
    return BFP_OK;
}
}  // namespace

int host_staging_release() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_st.device >= 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaSetDevice(g_st.device);
        release_locked();
        cudaSetDevice(dev);
    }
    return BFP_OK;
}

int quantize_host(const QuantArgs& a) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (a.rows == 0 || a.K == 0) return BFP_OK;
    const size_t in_es = dtype_size(a.in_dtype), out_es = dtype_size(a.out_dtype);
    // rows per chunk: multiples of 8 rows so every chunk starts on a Philox-counter and 128-bit boundary
    const int64_t row_bytes = std::max<int64_t>(1, (int64_t)(a.K * in_es));
    auto rows_for = [&](int64_t bytes) { return std::max<int64_t>(8, (std::max<int64_t>(1, bytes / row_bytes) / 8) * 8); };
    const int64_t max_rows = std::min(rows_for(tuning().host_chunk_bytes), ((a.rows + 7) / 8) * 8);
    const int64_t min_rows = std::min(rows_for(tuning().host_chunk_min_bytes), max_rows);
    if (int rc = ensure_locked((size_t)max_rows * a.K * in_es, (size_t)max_rows * a.K * out_es)) return rc;
    cudaError_t e;
    int slot = 0;
    int64_t grow = min_rows;
    for (int64_t r0 = 0; r0 < a.rows; slot = (slot + 1) % kSlots) {
        const int64_t left = a.rows - r0;
        // taper: geometric growth at the start, half of what is left towards the end
        int64_t nr = std::min(grow, std::max(min_rows, ((left / 2 + 7) / 8) * 8));
        nr = std::min(std::min(nr, max_rows), left);
        if (left - nr < min_rows / 2) nr = std::min(left, max_rows);       // do not leave a sliver
        grow = std::min(max_rows, grow * 2);
        cudaStream_t st = g_st.stream[slot];
        const char* hin = static_cast<const char*>(a.in) + (size_t)r0 * a.K * in_es;
        char* hout = static_cast<char*>(a.out) + (size_t)r0 * a.K * out_es;
        if ((e = cudaMemcpyAsync(g_st.d_in[slot], hin, (size_t)nr * a.K * in_es, cudaMemcpyHostToDevice, st)) != cudaSuccess)
            return cuda_fail(e, "H2D copy");
        QuantArgs c = a;
        c.in = g_st.d_in[slot]; c.out = g_st.d_out[slot]; c.rows = nr; c.index_base = a.index_base + r0 * a.K;
        if (int rc = quantize_device(c, st)) return rc;
        if ((e = cudaMemcpyAsync(hout, g_st.d_out[slot], (size_t)nr * a.K * out_es, cudaMemcpyDeviceToHost, st)) != cudaSuccess)
            return cuda_fail(e, "D2H copy");
        r0 += nr;
    }
    for (int i = 0; i < kSlots; ++i)
        if ((e = cudaStreamSynchronize(g_st.stream[i])) != cudaSuccess) return cuda_fail(e, "stream sync");
    return BFP_OK;
}

}  // namespace bfp
