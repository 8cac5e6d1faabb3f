// bfp_int.cu -- the SparseGPT-style per-channel symmetric INT-k fake quantiser.
//
// Replaces int_ops.Quantizer.configure/find_params/quantize as reached from _quantize with
// sparsity_num_format == 'int' (bfp_ops.py:111-120; int_ops.py:6-8, 18-31, 33-120 with perchannel=True, sym=True):
//   channel of a WEIGHT  [C, ...]        : dim 0 (int_ops.py:40-43)
//   channel of an ACTIVATION [.., C]     : last dim for 2-D / 3-D, dim 1 for 4-D (int_ops.py:44-50)
//   xmin = min(min_c, 0), xmax = max(max_c, 0); xmax = max(|xmin|, xmax); xmin = -xmax where xmin < 0;
//   all-zero channel -> (-1, +1); scale = (xmax - xmin) / maxq; zero = (maxq + 1) / 2; maxq = 2^bits - 1
//   y = scale * (clamp(round(x / scale) + zero, 0, maxq) - zero)
// Every step is fp32 for every input dtype (the fp32 torch.zeros in find_params promotes half inputs, and the output
// of the reference is fp32), so half inputs are converted exactly and the arithmetic below is literal.
// The tensor is viewed as [A, C, inner]: element i has channel (i / inner) % C.
#include <algorithm>

#include "bfp_internal.h"
#include "bfp_stream.cuh"

namespace bfp {
namespace {

__device__ __forceinline__ int enc(float f) { const int b = __float_as_int(f); return b ^ ((b >> 31) & 0x7fffffff); }   // monotone
__device__ __forceinline__ float dec(int e) { return __int_as_float(e ^ ((e >> 31) & 0x7fffffff)); }

struct IntWs { int* mn; int* mx; float* scale; };

__global__ void int_init_kernel(int* mn, int* mx, int64_t C) {
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < C; c += (int64_t)gridDim.x * blockDim.x) {
        mn[c] = enc(__int_as_float(0x7f800000));
        mx[c] = enc(__int_as_float(0xff800000));
    }
}

// contiguous runs (inner >= 32): one warp per run (a, c)
template <int DT>
__global__ void __launch_bounds__(256) int_minmax_runs_kernel(const void* in, int64_t runs, int64_t C, int64_t inner, int* mn, int* mx) {
    const int lane = threadIdx.x & 31;
    for (int64_t run = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; run < runs; run += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const int64_t base = run * inner;
        float lo = __int_as_float(0x7f800000), hi = __int_as_float(0xff800000);
        for (int64_t j = lane; j < inner; j += 32) { const float v = DType<DT>::load(in, base + j); lo = fminf(lo, v); hi = fmaxf(hi, v); }
#pragma unroll
        for (int off = 16; off; off >>= 1) { lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, off)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, off)); }
        if (lane == 0) { const int64_t c = run % C; atomicMin(&mn[c], enc(lo)); atomicMax(&mx[c], enc(hi)); }
    }
}

// inner == 1: thread per column, a chunk of rows per CTA row (coalesced across the warp)
template <int DT>
__global__ void __launch_bounds__(256) int_minmax_cols_kernel(const void* in, int64_t A, int64_t C, int64_t rows_per_cta, int* mn, int* mx) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const int64_t a0 = (int64_t)blockIdx.y * rows_per_cta, a1 = min(A, a0 + rows_per_cta);
    float lo = __int_as_float(0x7f800000), hi = __int_as_float(0xff800000);
    for (int64_t a = a0; a < a1; ++a) { const float v = DType<DT>::load(in, a * C + c); lo = fminf(lo, v); hi = fmaxf(hi, v); }
    if (a0 < a1) { atomicMin(&mn[c], enc(lo)); atomicMax(&mx[c], enc(hi)); }
}

// anything else (1 < inner < 32): per-element atomics; rare shapes only
template <int DT>
__global__ void __launch_bounds__(256) int_minmax_generic_kernel(const void* in, int64_t n, int64_t C, int64_t inner, int* mn, int* mx) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = DType<DT>::load(in, i);
        const int64_t c = (i / inner) % C;
        atomicMin(&mn[c], enc(v)); atomicMax(&mx[c], enc(v));
    }
}

__global__ void int_scale_kernel(const int* mn, const int* mx, float* scale, int64_t C, float maxq) {
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < C; c += (int64_t)gridDim.x * blockDim.x) {
        float xmin = fminf(dec(mn[c]), 0.0f), xmax = fmaxf(dec(mx[c]), 0.0f);      // int_ops.py:55-56
        xmax = fmaxf(fabsf(xmin), xmax);                                           // :59
        if (xmin < 0.0f) xmin = -xmax;                                             // :60-62
        if (xmin == 0.0f && xmax == 0.0f) { xmin = -1.0f; xmax = 1.0f; }           // :63-65
        scale[c] = __fdiv_rn(xmax - xmin, maxq);                                   // :67
    }
}

template <int DT>
__global__ void __launch_bounds__(256) int_apply_kernel(const void* in, float* out, int64_t n, int64_t C, int64_t inner, const float* scale,
                                                        float maxq, float zero) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float x = DType<DT>::load(in, i);
        const float s = scale[(i / inner) % C];
        const float q = fminf(fmaxf(rintf(__fdiv_rn(x, s)) + zero, 0.0f), maxq);   // int_ops.py:7
        out[i] = s * (q - zero);                                                   // :8
    }
}

// Activations [A, C] (inner == 1) for the tensor-core contraction: the same per-column quantiser, but the fp32 result
// x = s_c (q - zero) is written as three bf16 planes with x == hi + mid + lo (8 + 8 + 8 significant bits; the last plane
// absorbs any remainder to within 2^-25 |x|).  bf16 GEMMs against the integer weight grid then reproduce the fp32 product
// sum.  Layout (one buffer of 3 A C elements), chosen for the K-chunked contraction of include/bfp_b200.h:
//   hi   column segments of width kseg, segment-major: segment s is a contiguous [A, w_s] matrix at offset s * A * kseg
//   mid|lo one [A, 2C] matrix at offset A * C (mid in columns 0 .. C-1, lo in C .. 2C-1)
// 8 columns per thread (kseg % 8 == 0).
template <int DT>
__global__ void __launch_bounds__(256) int_apply_split_kernel(const void* in, __nv_bfloat16* out, int64_t A, int64_t C, int64_t kseg,
                                                              const float* scale, float maxq, float zero) {
    const int64_t cv = C / 8, nv = A * cv;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / cv, c0 = (i - row * cv) * 8;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = DType<DT>::load(in, row * C + c0 + e);
        const float4 s0 = *reinterpret_cast<const float4*>(scale + c0), s1 = *reinterpret_cast<const float4*>(scale + c0 + 4);
        const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        __align__(16) __nv_bfloat16 p[3][8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float q = fminf(fmaxf(rintf(__fdiv_rn(v[e], sc[e])) + zero, 0.0f), maxq);   // int_ops.py:7
            const float x = sc[e] * (q - zero);                                               // :8
            p[0][e] = __float2bfloat16_rn(x);
            const float r1 = x - __bfloat162float(p[0][e]);
            p[1][e] = __float2bfloat16_rn(r1);
            p[2][e] = __float2bfloat16_rn(r1 - __bfloat162float(p[1][e]));
        }
        const int64_t seg = c0 / kseg, ws = min(kseg, C - seg * kseg);
        *reinterpret_cast<uint4*>(out + seg * A * kseg + row * ws + (c0 - seg * kseg)) = *reinterpret_cast<const uint4*>(p[0]);
        __nv_bfloat16* ml = out + A * C + row * 2 * C + c0;
        *reinterpret_cast<uint4*>(ml) = *reinterpret_cast<const uint4*>(p[1]);
        *reinterpret_cast<uint4*>(ml + C) = *reinterpret_cast<const uint4*>(p[2]);
    }
}

// Weights (A == 1: one channel per row of K = inner elements): ONE pass.  A CTA keeps its row in registers (NV 128-bit
// vectors per thread), reduces min / max across the block, derives the scale and applies it to the registers: 8 B/element of
// traffic (4 in + 4 out for fp32) instead of the three passes (12 B/element + atomics) of the general path.
//
// ORDER adds the N:4 magnitude mask of bfp_ops.py:73-91 to the same pass (groups of 4 never straddle a 128-bit vector):
// 1 = mask, then quantise the masked row (first == 's': the min / max see the zeros); 2 = quantise, then mask the quantised
// values (their ties decide).  torch-CUDA tie rule, like the stand-alone N:M kernel.
template <int V>
__device__ __forceinline__ void mask_vec4(float* v, int kdrop) {
#pragma unroll
    for (int g = 0; g < V / 4; ++g) {
        if (kdrop == 2) nm_mask4<2>(v + 4 * g);
        else if (kdrop == 1) nm_mask4<1>(v + 4 * g);
        else nm_mask4<3>(v + 4 * g);
    }
}

template <int DT, int NV, int ORDER>
__global__ void __launch_bounds__(256) int_rows_fused_kernel(const void* in, float* out, int64_t rows, int64_t K, float maxq, float zero, int kdrop) {
    constexpr int V = DType<DT>::kVec;
    __shared__ float s_lo[8], s_hi[8];
    const int64_t nv = K / V;                               // <= 256 * NV, checked by the host
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
        const uint4* src = static_cast<const uint4*>(in) + row * nv;
        uint4 raw[NV];
        float lo = __int_as_float(0x7f800000), hi = __int_as_float(0xff800000);
#pragma unroll
        for (int u = 0; u < NV; ++u) {
            const int64_t j = threadIdx.x + u * 256;
            if (j < nv) raw[u] = ld_stream(src + j);
        }
#pragma unroll
        for (int u = 0; u < NV; ++u) {
            const int64_t j = threadIdx.x + u * 256;
            if (j < nv) {
                float v[V];
                unpack_vec<DT>(raw[u], v);
                if (ORDER == 1) { mask_vec4<V>(v, kdrop); raw[u] = pack_vec<DT>(v); }     // kept values are unchanged, dropped become +0
#pragma unroll
                for (int e = 0; e < V; ++e) { lo = fminf(lo, v[e]); hi = fmaxf(hi, v[e]); }
            }
        }
#pragma unroll
        for (int off = 16; off; off >>= 1) { lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, off)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, off)); }
        if (lane == 0) { s_lo[warp] = lo; s_hi[warp] = hi; }
        __syncthreads();
#pragma unroll
        for (int w = 0; w < 8; ++w) { lo = fminf(lo, s_lo[w]); hi = fmaxf(hi, s_hi[w]); }
        float xmin = fminf(lo, 0.0f), xmax = fmaxf(hi, 0.0f);                          // int_ops.py:55-56
        xmax = fmaxf(fabsf(xmin), xmax);                                               // :59
        if (xmin < 0.0f) xmin = -xmax;                                                 // :60-62
        if (xmin == 0.0f && xmax == 0.0f) { xmin = -1.0f; xmax = 1.0f; }               // :63-65
        const float sc = __fdiv_rn(xmax - xmin, maxq);                                 // :67
        float4* dst = reinterpret_cast<float4*>(out + row * K);
#pragma unroll
        for (int u = 0; u < NV; ++u) {
            const int64_t j = threadIdx.x + u * 256;
            if (j < nv) {
                float v[V];
                unpack_vec<DT>(raw[u], v);
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    const float q = fminf(fmaxf(rintf(__fdiv_rn(v[e], sc)) + zero, 0.0f), maxq);   // int_ops.py:7
                    v[e] = sc * (q - zero);                                                        // :8
                }
                if (ORDER == 2) mask_vec4<V>(v, kdrop);
#pragma unroll
                for (int f4 = 0; f4 < V / 4; ++f4)
                    st_stream(reinterpret_cast<uint4*>(dst + j * (V / 4) + f4), pack_vec<BFP_DT_F32>(v + 4 * f4));
            }
        }
        __syncthreads();                                    // s_lo / s_hi are rewritten by the next row
    }
}

template <int DT>
bool rows_fusable(const void* in, const float* out, int64_t inner) {
    constexpr int V = DType<DT>::kVec;
    return inner % V == 0 && inner / V <= 256 * 16 && reinterpret_cast<uintptr_t>(in) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0;
}

template <int DT, int ORDER>
int run_rows(const void* in, float* out, int64_t C, int64_t inner, int bits, int kdrop, cudaStream_t s) {
    const float maxq = (float)((1ll << bits) - 1), zero = (float)(((1ll << bits)) / 2.0);
    const int64_t nv = inner / DType<DT>::kVec;
    const int grid = (int)std::min<int64_t>(C, (int64_t)device_info().sm_count * 4);
    if (nv <= 256 * 4) int_rows_fused_kernel<DT, 4, ORDER><<<grid, 256, 0, s>>>(in, out, C, inner, maxq, zero, kdrop);
    else if (nv <= 256 * 8) int_rows_fused_kernel<DT, 8, ORDER><<<grid, 256, 0, s>>>(in, out, C, inner, maxq, zero, kdrop);
    else int_rows_fused_kernel<DT, 16, ORDER><<<grid, 256, 0, s>>>(in, out, C, inner, maxq, zero, kdrop);
    count_launch();
    return check_launch("int_rows_fused_kernel");
}

template <int DT>
int run_nm(const void* in, float* out, int64_t C, int64_t K, int bits, int N, int order, cudaStream_t s) {
    if (!rows_fusable<DT>(in, out, K))
        return set_error(BFP_E_UNSUPPORTED, "fused INT + N:4 needs 16-byte aligned buffers, K a multiple of the vector width and K <= 4096 vectors");
    return order == BFP_ORDER_SPARSIFY_QUANT ? run_rows<DT, 1>(in, out, C, K, bits, 4 - N, s) : run_rows<DT, 2>(in, out, C, K, bits, 4 - N, s);
}

template <int DT>
int run(const void* in, float* out, int64_t A, int64_t C, int64_t inner, int bits, IntWs ws, cudaStream_t s, __nv_bfloat16* split_out = nullptr,
        int64_t kseg = 0) {
    const int64_t n = A * C * inner;
    const int sms = device_info().sm_count;
    const float maxq = (float)((1ll << bits) - 1), zero = (float)(((1ll << bits)) / 2.0);
    if (!split_out && A == 1 && rows_fusable<DT>(in, out, inner)) return run_rows<DT, 0>(in, out, C, inner, bits, 0, s);
    int_init_kernel<<<(int)std::min<int64_t>((C + 255) / 256, 1024), 256, 0, s>>>(ws.mn, ws.mx, C);
    count_launch();
    if (inner >= 32) {
        const int64_t runs = A * C;
        const int grid = (int)std::min<int64_t>((runs * 32 + 255) / 256, (int64_t)sms * 16);
        int_minmax_runs_kernel<DT><<<grid, 256, 0, s>>>(in, runs, C, inner, ws.mn, ws.mx);
    } else if (inner == 1) {
        const int64_t col_ctas = (C + 255) / 256;
        int64_t row_ctas = std::max<int64_t>(1, std::min<int64_t>((int64_t)sms * 8 / col_ctas, (A + 63) / 64));
        row_ctas = std::min<int64_t>(row_ctas, 65535);
        const int64_t rows_per_cta = (A + row_ctas - 1) / row_ctas;
        dim3 grid((unsigned)col_ctas, (unsigned)((A + rows_per_cta - 1) / rows_per_cta));
        int_minmax_cols_kernel<DT><<<grid, 256, 0, s>>>(in, A, C, rows_per_cta, ws.mn, ws.mx);
    } else {
        int_minmax_generic_kernel<DT><<<(int)std::min<int64_t>((n + 255) / 256, (int64_t)sms * 16), 256, 0, s>>>(in, n, C, inner, ws.mn, ws.mx);
    }
    count_launch();
    int_scale_kernel<<<(int)std::min<int64_t>((C + 255) / 256, 1024), 256, 0, s>>>(ws.mn, ws.mx, ws.scale, C, maxq);
    count_launch();
    if (split_out) int_apply_split_kernel<DT><<<(int)std::min<int64_t>((n / 8 + 255) / 256, (int64_t)sms * 16), 256, 0, s>>>(in, split_out, A, C, kseg, ws.scale, maxq, zero);
    else int_apply_kernel<DT><<<(int)std::min<int64_t>((n + 255) / 256, (int64_t)sms * 16), 256, 0, s>>>(in, out, n, C, inner, ws.scale, maxq, zero);
    count_launch();
    return check_launch("int quantiser kernels");
}
}  // namespace

size_t int_workspace_bytes(int64_t C) { return (size_t)C * 12 + 64; }

int int_quantize_device(const void* in, float* out, int64_t A, int64_t C, int64_t inner, int dtype, int bits, void* workspace, cudaStream_t s) {
    if (A * C * inner == 0) return BFP_OK;
    IntWs ws;
    ws.mn = static_cast<int*>(workspace); ws.mx = ws.mn + C; ws.scale = reinterpret_cast<float*>(ws.mx + C);
    switch (dtype) {
    case BFP_DT_F32: return run<BFP_DT_F32>(in, out, A, C, inner, bits, ws, s);
    case BFP_DT_F16: return run<BFP_DT_F16>(in, out, A, C, inner, bits, ws, s);
    case BFP_DT_BF16: return run<BFP_DT_BF16>(in, out, A, C, inner, bits, ws, s);
    }
    return set_error(BFP_E_ARG, "bad dtype");
}

int int_quantize_split3_device(const void* in, void* out_bf16, int64_t A, int64_t C, int64_t kseg, int dtype, int bits, void* workspace, cudaStream_t s) {
    if (A * C == 0) return BFP_OK;
    IntWs ws;
    ws.mn = static_cast<int*>(workspace); ws.mx = ws.mn + C; ws.scale = reinterpret_cast<float*>(ws.mx + C);
    __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out_bf16);
    switch (dtype) {
    case BFP_DT_F32: return run<BFP_DT_F32>(in, nullptr, A, C, 1, bits, ws, s, o, kseg);
    case BFP_DT_F16: return run<BFP_DT_F16>(in, nullptr, A, C, 1, bits, ws, s, o, kseg);
    case BFP_DT_BF16: return run<BFP_DT_BF16>(in, nullptr, A, C, 1, bits, ws, s, o, kseg);
    }
    return set_error(BFP_E_ARG, "bad dtype");
}

int int_quantize_nm_device(const void* in, float* out, int64_t C, int64_t K, int dtype, int bits, int N, int order, cudaStream_t s) {
    if (C * K == 0) return BFP_OK;
    switch (dtype) {
    case BFP_DT_F32: return run_nm<BFP_DT_F32>(in, out, C, K, bits, N, order, s);
    case BFP_DT_F16: return run_nm<BFP_DT_F16>(in, out, C, K, bits, N, order, s);
    case BFP_DT_BF16: return run_nm<BFP_DT_BF16>(in, out, C, K, bits, N, order, s);
    }
    return set_error(BFP_E_ARG, "bad dtype");
}

}  // namespace bfp
