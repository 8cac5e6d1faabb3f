// bfp_ocp_mx.cu -- the MX (OCP Microscaling) formats of the reference's mx_layers.py on sm_100a: one fused pass over HBM for
//     quantize_elemwise_op (bfloat rounding)  ->  quantize_mx_op (shared power-of-two scale per block + narrow elements)
// with three outputs: the fake-quantised tensor (what the microxcaling emulation returns), the exact-bf16 operand of the tcgen05
// bf16 GEMMs, or E4M3 bytes + UE8M0 scale atoms -- the operand form of tcgen05.mma.kind::mxf8f6f4.block_scale (bfp_gemm_mx.cu), for
// which an MX block IS the hardware's block: the tensor core applies the shared scales itself.
//
// Reference call sites: /root/reference/src/transformers/bfp/mx_layers.py:23-109 (MXLinear / MXConv2d / MXMatmul over mx.Linear,
// mx.Conv2d, mx.matmul), element-format parameters formats.py:86-123, defaults specs.py:30-66.  The arithmetic lives in
// microsoft/microxcaling (un-vendored, un-pinned: parity unpinned); it is restated here operation by operation in fp32 as that
// library evaluates it on an fp32 tensor (oracle/mx_oracle.py has the same restatement in numpy, tests compare bit for bit):
//     bfloatX:  sign(A) floor(|A| 2^(bits-2-pe) + 0.5) 2^(pe-bits+2),  pe = max(floor(log2 |A|), -126), overflow -> Inf
//     shared:   se = clamp(floor(log2 max|A|) - emax_elem, -127, 127)          (all-zero block: log2 of 2^-126)
//     element:  a = A / 2^se;  pe = max(floor(log2 |a|), min_exp);  r = sign(a) floor(|a| 2^(bits-2-pe) + 0.5) 2^(pe-bits+2);
//               clamp to +-max_norm;  result r 2^se
// 'nearest' in that library is round-half-AWAY-from-zero (floor(|x| + 0.5)), not half-to-even; sign(+-0) = 0, so zeros come out
// +0.0 while negative values that round to zero come out -0.0.  fp16 / bf16 tensors are computed in fp32 and rounded once to
// the dtype (the library's op-by-op half-precision rounding is not reproduced).
#include <algorithm>

#include "bfp_stream.cuh"
#include "bfp_internal.h"

namespace bfp {

namespace ocp {

struct Format { int ebits, mbits, emax; float max_norm; bool e4m3_ok; };
// formats.py:86-123; e4m3_ok: every element value is exactly an E4M3 number (the block-scaled tensor-core operand type)
static bool format_of(int id, Format* f) {
    switch (id) {
        case BFP_MX_INT8: *f = {0, 8, 0, 1.984375f, false}; return true;
        case BFP_MX_INT4: *f = {0, 4, 0, 1.75f, true}; return true;
        case BFP_MX_INT2: *f = {0, 2, 0, 1.0f, true}; return true;
        case BFP_MX_FP8_E5M2: *f = {5, 4, 15, 57344.0f, false}; return true;
        case BFP_MX_FP8_E4M3: *f = {4, 5, 8, 448.0f, true}; return true;
        case BFP_MX_FP6_E3M2: *f = {3, 4, 4, 28.0f, true}; return true;
        case BFP_MX_FP6_E2M3: *f = {2, 5, 2, 7.5f, true}; return true;
        case BFP_MX_FP4_E2M1: *f = {2, 3, 2, 6.0f, true}; return true;
        default: return false;
    }
}

struct Params {
    const void* in;
    void* out;              // fake-quant / bf16 operand: [rows, K] (or [rows, ld_out]); packed: vals uint8 [rows, Kp]
    uint8_t* sf;            // packed: UE8M0 scale atoms of 128-row tiles (bfp_gemm_mx.cu)
    int64_t rows, K, ld_out;
    int64_t n_vec;          // stream kernel: 128-bit input vectors
    int64_t n_row_tiles;    // packed: ceil(rows / tile_rows)
    int tile_rows, atoms;   // packed: rows per scale tile (128 for the activation operand, the GEMM's N tile for the weight), atoms per tile
    int block, lanes_per_block;
    int ebits, mbits, emax, min_exp;
    float max_norm;
    int scale_emax;         // 2^(scale_bits - 1) - 1
    int bfloat;             // 0 / 32: no bfloat rounding, else 10 .. 31
    int flush;              // mx_flush_fp32_subnorms
};

// floor(log2f(s)) for finite s > 0 as torch evaluates it: the exponent field, except within 128 ulp below a power of two, where the
// fp32 logarithm may round up to the next integer (the same libdevice log2f torch-CUDA calls).  Out of line: ~1e-5 of the blocks.
static __device__ __noinline__ int floor_log2_near_pow2(float s) { return (int)floorf(log2f(s)); }

struct Shared { int se; bool nan, zero; };
// shared exponent of a block from the bit pattern of max |A| (mx/mx_ops.py _shared_exponents + the offset / clamp of _quantize_mx)
__device__ __forceinline__ Shared shared_exponent(uint32_t amax_bits, const Params& p) {
    Shared s;
    s.nan = amax_bits >= 0x7f800000u;                                  // Inf or NaN anywhere in the block: 2^se is NaN, the block is NaN
    int k;
    if (amax_bits == 0u) k = -126;                                    // log2(0 + FP32_MIN_NORMAL)
    else if (amax_bits < 0x00800000u) k = -127 - (__clz(amax_bits) - 9);      // subnormal maximum: true floor(log2), always clamped below
    else {
        k = (int)(amax_bits >> 23) - 127;
        if ((amax_bits & 0x7fffffu) >= 0x7fff80u && !s.nan) k = floor_log2_near_pow2(__uint_as_float(amax_bits));
    }
    s.zero = p.flush && k <= -127;                                    // A * (shared_exp > -127)
    int se = k - p.emax;
    if (se > p.scale_emax) s.nan = true;
    s.se = max(se, -p.scale_emax);
    return s;
}
// 2^e for -127 <= e <= 127 (2^-127 is the fp32 denormal 0x00400000)
__device__ __forceinline__ float pow2i(int e) {
    e = min(max(e, -149), 127);                                        // (only NaN-marked blocks ever ask for more)
    return e >= -126 ? __uint_as_float((uint32_t)(e + 127) << 23) : __uint_as_float(0x00400000u >> (-127 - e));
}

// one element: a = A / 2^se already formed.  Returns the quantised element (before the multiplication by 2^se).
__device__ __forceinline__ float quantize_element(float a, const Params& p) {
    const uint32_t ab = __float_as_uint(a) & 0x7fffffffu;
    float y;
    if (p.ebits > 0) {
        const int ea = (int)(ab >> 23) - 127;                          // floor(log2 |a|); zero / subnormal fall below min_exp anyway
        const int pe = max(ea, p.min_exp);
        const float up = pow2i(p.mbits - 2 - pe), down = pow2i(pe - p.mbits + 2);
        y = floorf(__fadd_rn(__uint_as_float(ab) * up, 0.5f)) * down;
    } else {
        const float up = pow2i(p.mbits - 2), down = pow2i(2 - p.mbits);
        y = floorf(__fadd_rn(__uint_as_float(ab) * up, 0.5f)) * down;
    }
    y = fminf(y, p.max_norm);
    if (ab == 0u) return 0.0f;                                         // sign(+-0) = 0
    return (__float_as_uint(a) & 0x80000000u) ? -y : y;
}

// E4M3 byte of a value that is exactly representable in E4M3 (|v| <= 448, multiples of 2^-9 below 2^-6)
__device__ __forceinline__ uint32_t e4m3_byte(float v) {
    const uint32_t b = __float_as_uint(v), a = b & 0x7fffffffu, s = (b >> 24) & 0x80u;
    if (a == 0u) return s;
    const int e = (int)(a >> 23) - 127;
    if (e >= -6) return s | ((uint32_t)(e + 7) << 3) | ((a >> 20) & 7u);
    return s | (uint32_t)(__uint_as_float(a) * 512.0f);                // subnormal: m 2^-9
}

enum { kOutFake = 0, kOutBf16 = 1, kOutPacked = 2 };

// Stream kernel: K % block == 0, block a power-of-two multiple of the 128-bit vector, so the tensor is a flat sequence of vectors and
// a block is 2^j adjacent lanes (block max by butterfly).  Persistent grid, kStreamUnroll independent 128-bit loads per thread.
//   OUT = kOutFake: out has the input dtype;  kOutBf16: bf16 [rows, ld_out];  kOutPacked (block 32 / 64 / 128, K % 128 == 0, a warp
//   covers whole 128-element slabs of one row): E4M3 bytes + one UE8M0 byte per (row, 32 elements) in scale atoms.
template <int DT, int OUT>
__global__ void __launch_bounds__(kStreamThreads) mx_stream_kernel(const Params p) {
    using D = DType<DT>;
    constexpr int V = D::kVec;
    constexpr int kTileVecs = kStreamThreads * kStreamUnroll;
    const int64_t n_tiles = (p.n_vec + kTileVecs - 1) / kTileVecs;
    const uint4* in = static_cast<const uint4*>(p.in);
    const int lane = threadIdx.x & 31;
    const int64_t vec_per_row = p.K / V;
    pdl_launch_dependents();
    pdl_wait();
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t base = tile * kTileVecs;
        uint4 raw[kStreamUnroll];
#pragma unroll
        for (int u = 0; u < kStreamUnroll; ++u) {
            const int64_t vi = base + (int)threadIdx.x + u * kStreamThreads;
            raw[u] = vi < p.n_vec ? ld_stream(in + vi) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < kStreamUnroll; ++u) {
            const int64_t vi = base + (int)threadIdx.x + u * kStreamThreads;
            float v[V];
            unpack_vec<DT>(raw[u], v);
            uint32_t amax = 0u;
#pragma unroll
            for (int i = 0; i < V; ++i) { v[i] = round_bfloat(v[i], p.bfloat); amax = max(amax, __float_as_uint(v[i]) & 0x7fffffffu); }
#pragma unroll
            for (int off = 1; off < 32; off <<= 1)
                if (off < p.lanes_per_block) amax = max(amax, __shfl_xor_sync(0xffffffffu, amax, off));
            const Shared sh = shared_exponent(amax, p);
            const float scale = pow2i(sh.se), inv = pow2i(-sh.se);
            float q[V];                                                // elements before the multiplication by 2^se
#pragma unroll
            for (int i = 0; i < V; ++i) {
                const float a = sh.zero ? 0.0f : v[i] * inv;           // = v / 2^se: the same correctly rounded quotient
                q[i] = quantize_element(a, p);
            }
            if (OUT == kOutPacked) {
                uint32_t bytes[V / 4];
#pragma unroll
                for (int w = 0; w < V / 4; ++w) {
                    uint32_t o = 0u;
#pragma unroll
                    for (int b = 0; b < 4; ++b) o |= (sh.nan ? 0u : e4m3_byte(q[4 * w + b])) << (8 * b);
                    bytes[w] = o;
                }
                const uint32_t sbyte = sh.nan ? 0xffu : (uint32_t)(sh.se + 127);
                const bool live = vi < p.n_vec;
                const int64_t vi_c = live ? vi : 0;
                const int64_t row = vi_c / vec_per_row, col = (vi_c - row * vec_per_row) * V;
                if (live) {
                    uint8_t* dst = static_cast<uint8_t*>(p.out) + row * p.ld_out + col;
                    if (V == 4) *reinterpret_cast<uint32_t*>(dst) = bytes[0];
                    else *reinterpret_cast<uint2*>(dst) = make_uint2(bytes[0], bytes[V == 8 ? 1 : 0]);
                }
                // scale bytes: a warp covers 32 V consecutive elements of one row (K % (32 V) == 0) = 32 V / 128 slabs of four 32-groups;
                // the first lane of each slab gathers its four bytes and writes one word of the row's atom
                constexpr int kLanesPerSlab = 128 / V, kLanesPerGroup = 32 / V;
                uint32_t word = 0u;
#pragma unroll
                for (int g = 0; g < 4; ++g) word |= __shfl_sync(0xffffffffu, sbyte, (lane / kLanesPerSlab) * kLanesPerSlab + g * kLanesPerGroup) << (8 * g);
                if (live && (lane % kLanesPerSlab) == 0) {
                    const int64_t slab = col >> 7, rt = row / p.tile_rows;
                    const int rr = (int)(row - rt * p.tile_rows), r = rr & 127;
                    *reinterpret_cast<uint32_t*>(p.sf + (((slab * p.n_row_tiles + rt) * p.atoms + (rr >> 7)) * 512) + 16 * (r & 31) + 4 * (r >> 5)) = word;
                }
            } else {
#pragma unroll
                for (int i = 0; i < V; ++i) q[i] = sh.nan ? __uint_as_float(0x7fc00000u) : q[i] * scale;
                if (vi < p.n_vec) {
                    if (OUT == kOutFake) {
                        st_stream(static_cast<uint4*>(p.out) + vi, pack_vec<DT>(q));
                    } else {
                        const int64_t row = vi / vec_per_row, col = (vi - row * vec_per_row) * V;
                        __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(p.out) + row * p.ld_out + col;
                        if (V == 8) {
                            st_stream(reinterpret_cast<uint4*>(dst), pack_vec<BFP_DT_BF16>(q));
                        } else {
                            __nv_bfloat162 lo = __floats2bfloat162_rn(q[0], q[1]), hi = __floats2bfloat162_rn(q[2], q[V == 4 ? 3 : 0]);
                            *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
                        }
                    }
                }
            }
        }
    }
}

// Generic kernel: any K / block (ragged last block = the zero padding of _reshape_to_blocks, a whole row as one block when block = 0
// is passed as K).  One thread per block, two sweeps over its elements.  Correctness path (ViT patch embedding: 3 channels per block).
template <int DT, int OUT>
__global__ void __launch_bounds__(256) mx_generic_kernel(const Params p) {
    using D = DType<DT>;
    const int64_t blocks_per_row = (p.K + p.block - 1) / p.block;
    const int64_t n_blocks = p.rows * blocks_per_row;
    for (int64_t bi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; bi < n_blocks; bi += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = bi / blocks_per_row, k0 = (bi - row * blocks_per_row) * p.block, k1 = min(p.K, k0 + p.block);
        uint32_t amax = 0u;
        for (int64_t k = k0; k < k1; ++k) amax = max(amax, __float_as_uint(round_bfloat(D::load(p.in, row * p.K + k), p.bfloat)) & 0x7fffffffu);
        const Shared sh = shared_exponent(amax, p);
        const float scale = pow2i(sh.se), inv = pow2i(-sh.se);
        for (int64_t k = k0; k < k1; ++k) {
            const float v = round_bfloat(D::load(p.in, row * p.K + k), p.bfloat);
            const float q = quantize_element(sh.zero ? 0.0f : v * inv, p);
            const float r = sh.nan ? __uint_as_float(0x7fc00000u) : q * scale;
            if (OUT == kOutFake) D::store(p.out, row * p.K + k, r);
            else static_cast<__nv_bfloat16*>(p.out)[row * p.ld_out + k] = __float2bfloat16_rn(r);
        }
    }
}

// quantize_elemwise_op for the bfloat formats, optionally fused with the bias step of mx/linear.py: out = rb(rb(x) + rb(bias[col]))
template <int DT>
__global__ void __launch_bounds__(256) bfloat_round_kernel(const void* in, void* out, const float* bias, int64_t n, int64_t ncols, int bfloat) {
    using D = DType<DT>;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float x = round_bfloat(D::load(in, i), bfloat);
        if (bias) x = round_bfloat(__fadd_rn(x, round_bfloat(bias[i % ncols], bfloat)), bfloat);
        D::store(out, i, x);
    }
}

}  // namespace ocp

static int fill_params(ocp::Params& p, int elem_format, int scale_bits, int bfloat, int flush) {
    ocp::Format f;
    if (!ocp::format_of(elem_format, &f)) return set_error(BFP_E_ARG, "unknown MX element format (BFP_MX_*)");
    if (scale_bits < 2 || scale_bits > 8) return set_error(BFP_E_UNSUPPORTED, "MX scale_bits must be in [2, 8] (8 = E8M0)");
    if (bfloat != 0 && (bfloat < 10 || bfloat > 32)) return set_error(BFP_E_ARG, "bfloat must be 0 or in [10, 32]");
    p.ebits = f.ebits; p.mbits = f.mbits; p.emax = f.emax; p.max_norm = f.max_norm;
    p.min_exp = f.ebits > 0 ? 2 - (1 << (f.ebits - 1)) : 0;
    p.scale_emax = (1 << (scale_bits - 1)) - 1;
    p.bfloat = bfloat; p.flush = flush ? 1 : 0;
    return BFP_OK;
}

template <int OUT>
static int launch_mx_quant(ocp::Params& p, int dtype, bool stream_ok, cudaStream_t st) {
    using namespace ocp;
    if (stream_ok) {
        const int V = dtype == BFP_DT_F32 ? 4 : 8;
        p.n_vec = p.rows * p.K / V;
        p.lanes_per_block = p.block / V;
        const int64_t n_tiles = (p.n_vec + kStreamThreads * kStreamUnroll - 1) / (kStreamThreads * kStreamUnroll);
        int rc;
        if (dtype == BFP_DT_F32) rc = launch_pdl(mx_stream_kernel<BFP_DT_F32, OUT>, stream_grid(kernel_occupancy(mx_stream_kernel<BFP_DT_F32, OUT>, kStreamThreads), n_tiles), kStreamThreads, st, p);
        else if (dtype == BFP_DT_F16) rc = launch_pdl(mx_stream_kernel<BFP_DT_F16, OUT>, stream_grid(kernel_occupancy(mx_stream_kernel<BFP_DT_F16, OUT>, kStreamThreads), n_tiles), kStreamThreads, st, p);
        else rc = launch_pdl(mx_stream_kernel<BFP_DT_BF16, OUT>, stream_grid(kernel_occupancy(mx_stream_kernel<BFP_DT_BF16, OUT>, kStreamThreads), n_tiles), kStreamThreads, st, p);
        if (rc) return rc;
        count_launch();
        return check_launch("mx_stream_kernel");
    }
    if constexpr (OUT == kOutPacked) {
        return set_error(BFP_E_UNSUPPORTED, "MX block-scaled pack: block_size 32 / 64 / 128, K a multiple of 128 (fp32) or 256 (half), 16-byte aligned buffers");
    } else {
        const int64_t n_blocks = p.rows * ((p.K + p.block - 1) / p.block);
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n_blocks + 255) / 256, (int64_t)device_info().sm_count * 16));
        if (dtype == BFP_DT_F32) mx_generic_kernel<BFP_DT_F32, OUT><<<grid, 256, 0, st>>>(p);
        else if (dtype == BFP_DT_F16) mx_generic_kernel<BFP_DT_F16, OUT><<<grid, 256, 0, st>>>(p);
        else mx_generic_kernel<BFP_DT_BF16, OUT><<<grid, 256, 0, st>>>(p);
        count_launch();
        return check_launch("mx_generic_kernel");
    }
}

static bool pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

int ocp_mx_quantize_device(const void* in, void* out, int64_t rows, int64_t K, int dtype, int out_kind, int64_t ld_out, int block_size, int elem_format,
                           int scale_bits, int bfloat, int flush, cudaStream_t st) {
    if (rows == 0 || K == 0) return BFP_OK;
    ocp::Params p = {};
    if (int rc = fill_params(p, elem_format, scale_bits, bfloat, flush)) return rc;
    p.in = in; p.out = out; p.rows = rows; p.K = K;
    p.block = block_size > 0 ? block_size : (int)std::min<int64_t>(K, INT32_MAX);       // block_size 0: the whole axis shares one exponent
    if (block_size == 0 && K > INT32_MAX) return set_error(BFP_E_UNSUPPORTED, "row too long for a single MX block");
    const int V = dtype == BFP_DT_F32 ? 4 : 8;
    const bool aligned = reinterpret_cast<uintptr_t>(in) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0;
    const bool stream_ok = !tuning().force_generic && aligned && pow2(p.block) && p.block >= V && p.block <= 32 * V && K % p.block == 0;
    if (out_kind == 0) { p.ld_out = K; return launch_mx_quant<ocp::kOutFake>(p, dtype, stream_ok, st); }
    p.ld_out = ld_out;
    if (ld_out < K || ld_out % 8) return set_error(BFP_E_ARG, "bf16 operand: ld_out >= K and a multiple of 8");
    return launch_mx_quant<ocp::kOutBf16>(p, dtype, stream_ok && (ld_out * 2) % 16 == 0, st);
}

int ocp_mx_pack_device(const void* in, uint8_t* vals, uint8_t* sf, int64_t rows, int64_t K, int dtype, int tile_rows, int block_size, int elem_format,
                       int scale_bits, int bfloat, int flush, cudaStream_t st) {
    if (rows == 0 || K == 0) return BFP_OK;
    ocp::Params p = {};
    if (int rc = fill_params(p, elem_format, scale_bits, bfloat, flush)) return rc;
    ocp::Format f;
    ocp::format_of(elem_format, &f);
    if (!f.e4m3_ok) return set_error(BFP_E_UNSUPPORTED, "this MX element format is not a subset of E4M3 (int8, fp8_e5m2): use the exact-bf16 operand");
    if (scale_bits != 8) return set_error(BFP_E_UNSUPPORTED, "the hardware scale is E8M0 (scale_bits 8)");
    const int V = dtype == BFP_DT_F32 ? 4 : 8;
    p.in = in; p.out = vals; p.sf = sf; p.rows = rows; p.K = K; p.ld_out = K; p.block = block_size;
    if (tile_rows < 1) return set_error(BFP_E_ARG, "tile_rows");
    p.tile_rows = tile_rows; p.atoms = (tile_rows + 127) / 128;
    p.n_row_tiles = (rows + tile_rows - 1) / tile_rows;
    const bool ok = (block_size == 32 || block_size == 64 || block_size == 128) && K % (32 * V) == 0 && reinterpret_cast<uintptr_t>(in) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(vals) % 16 == 0 && reinterpret_cast<uintptr_t>(sf) % 16 == 0;
    return launch_mx_quant<ocp::kOutPacked>(p, dtype, ok, st);
}

int bfloat_round_device(const void* in, void* out, const float* bias, int64_t n, int64_t ncols, int dtype, int bfloat, cudaStream_t st) {
    if (n == 0) return BFP_OK;
    if (bfloat != 0 && (bfloat < 10 || bfloat > 32)) return set_error(BFP_E_ARG, "bfloat must be 0 or in [10, 32]");
    if (bias && ncols <= 0) return set_error(BFP_E_ARG, "bias needs the row length");
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)device_info().sm_count * 16));
    using namespace ocp;
    if (dtype == BFP_DT_F32) bfloat_round_kernel<BFP_DT_F32><<<grid, 256, 0, st>>>(in, out, bias, n, ncols, bfloat);
    else if (dtype == BFP_DT_F16) bfloat_round_kernel<BFP_DT_F16><<<grid, 256, 0, st>>>(in, out, bias, n, ncols, bfloat);
    else bfloat_round_kernel<BFP_DT_BF16><<<grid, 256, 0, st>>>(in, out, bias, n, ncols, bfloat);
    count_launch();
    return check_launch("bfloat_round_kernel");
}

}  // namespace bfp
