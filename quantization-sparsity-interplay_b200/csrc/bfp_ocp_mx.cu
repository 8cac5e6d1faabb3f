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

// Per-launch constants of the fast path (functions of the format only).
struct Fast {
    uint32_t bf_half, bf_mask;      // bfloat rounding on the bit pattern of |x|: (a + half) & mask  (identity: 0, ~0)
    int pe_min_b;                   // min_exp + 127
    int up_c;                       // 254 + mbits - 2: 2^(mbits-2-pe) has the bits (up_c - (pe + 127)) << 23
    float up_int, down_int;         // integer formats: 2^(mbits-2), 2^(2-mbits)
    float max_norm;
    int emax, scale_emax;
    int e4m3_sub;                   // packed: the format has values below 2^-6 (fp8_e4m3 only): E4M3 subnormal bytes
    FastDiv vec_per_row, tile_rows;
    uint32_t slab_stride;           // packed: bytes between the atoms of consecutive K slabs = n_row_tiles * atoms * 512
};

// One vector of a block whose shared exponent needs the literal evaluation (zero / subnormal / Inf / NaN maximum, a maximum within
// 128 ulp below a power of two, a scale at the end of the E8M0 range): rare, out of line, everything by value (no local memory on
// the caller's fast path).  lo / hi = the bfloat-rounded values; returns the quantised elements (sign applied), 2^se and its byte.
struct GeneralOut { uint4 lo, hi; float scale; uint32_t sbyte; };
template <int V>
__device__ __noinline__ GeneralOut mx_vec_general(uint4 lo, uint4 hi, uint32_t amax, const Params p) {
    const Shared sh = shared_exponent(amax, p);
    const float inv = pow2i(-sh.se);
    uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
    for (int i = 0; i < V; ++i)
        w[i] = sh.nan ? 0x7fc00000u : __float_as_uint(quantize_element(sh.zero ? 0.0f : __uint_as_float(w[i]) * inv, p));
    GeneralOut o;
    o.lo = make_uint4(w[0], w[1], w[2], w[3]); o.hi = make_uint4(w[4], w[5], w[6], w[7]);
    o.scale = sh.nan ? __uint_as_float(0x7fc00000u) : pow2i(sh.se);
    o.sbyte = sh.nan ? 0xffu : (uint32_t)(sh.se + 127);
    return o;
}

// Stream kernel: K % block == 0, block a power-of-two multiple of the 128-bit vector, so the tensor is a flat sequence of vectors and
// a block is 2^j adjacent lanes (block max by butterfly).  Persistent grid, kStreamUnroll independent 128-bit loads per thread.
//   OUT = kOutFake: out has the input dtype;  kOutBf16: bf16 [rows, K];  kOutPacked (block 32 / 64 / 128, K % 128 == 0, a warp
//   covers whole 128-element slabs of one row): E4M3 bytes + one UE8M0 byte per (row, 32 elements) in scale atoms.
//   FLOATFMT: the element format has exponent bits (private exponent per element) / is a fixed-point grid.
// Fast path per block (all but ~1e-5 of them): finite, normal maximum not within 128 ulp below a power of two, scale strictly inside
// the E8M0 range -- then se comes from the exponent field and every step is the literal fp32 operation of the emulation with the
// powers of two built from exponent bits (a multiplication by 2^-e is the same correctly rounded quotient as the division by 2^e).
template <int DT, int OUT, bool FLOATFMT>
__global__ void __launch_bounds__(kStreamThreads) mx_stream_kernel(const Params p, const Fast f) {
    using D = DType<DT>;
    constexpr int V = D::kVec;
    constexpr int kTileVecs = kStreamThreads * kStreamUnroll;
    const int64_t n_tiles = (p.n_vec + kTileVecs - 1) / kTileVecs;
    const uint4* in = static_cast<const uint4*>(p.in);
    const int lane = threadIdx.x & 31;
    pdl_launch_dependents();
    pdl_wait();
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t base = tile * kTileVecs;
        const int rem = (int)min((int64_t)kTileVecs, p.n_vec - base);
        uint4 raw[kStreamUnroll];
#pragma unroll
        for (int u = 0; u < kStreamUnroll; ++u) {
            const int li = (int)threadIdx.x + u * kStreamThreads;
            raw[u] = li < rem ? ld_stream(in + base + li) : make_uint4(0u, 0u, 0u, 0u);
        }
        uint32_t sf_off = 0u;
        if (OUT == kOutPacked) {
            // byte offset of (row, first slab, group 0) of the 32 V elements this warp covers in step (lane % kStreamUnroll): all lanes of
            // a warp share it, so lane u works it out for step u and the steps fetch it with one shuffle instead of every lane
            // repeating the address arithmetic for every vector
            static_assert(kStreamUnroll <= 32, "one lane per unroll step");
            const uint32_t vw = (uint32_t)base + (threadIdx.x & ~31u) + (uint32_t)(lane % kStreamUnroll) * kStreamThreads;
            const uint32_t row = fastdiv(vw, f.vec_per_row);
            const uint32_t slab = ((vw - row * (uint32_t)(p.K / V)) * V) >> 7;
            const uint32_t rt = fastdiv(row, f.tile_rows), rr = row - rt * (uint32_t)p.tile_rows, ra = rr & 127u;
            sf_off = slab * f.slab_stride + (rt * (uint32_t)p.atoms + (rr >> 7)) * 512u + 16u * (ra & 31u) + 4u * (ra >> 5);
        }
#pragma unroll
        for (int u = 0; u < kStreamUnroll; ++u) {
            const int li = (int)threadIdx.x + u * kStreamThreads;
            float v[V];
            unpack_vec<DT>(raw[u], v);
            uint32_t r[V], sgn[V];
            uint32_t amax = 0u;
#pragma unroll
            for (int i = 0; i < V; ++i) {
                const uint32_t b = __float_as_uint(v[i]), a = b & 0x7fffffffu;
                const uint32_t rr = (a + f.bf_half) & f.bf_mask;                   // bfloat rounding, half away from zero
                r[i] = a >= 0x7f800000u ? a : rr;                                  // Inf / NaN pass through
                sgn[i] = r[i] ? (b & 0x80000000u) : 0u;                            // sign(+-0) = 0: zeros come out +0.0
                amax = max(amax, r[i]);
            }
#pragma unroll
            for (int off = 1; off < 32; off <<= 1)
                if (off < p.lanes_per_block) amax = max(amax, __shfl_xor_sync(0xffffffffu, amax, off));
            const int se = (int)(amax >> 23) - 127 - f.emax;
            const bool fast = amax >= 0x00800000u && amax < 0x7f800000u && (amax & 0x7fffffu) < 0x7fff80u && se > -f.scale_emax && se < f.scale_emax;
            float q[V], scale;
            uint32_t sbyte;
            if (fast) {
                const float inv = __uint_as_float((uint32_t)(127 - se) << 23);
                scale = __uint_as_float((uint32_t)(127 + se) << 23);
                sbyte = (uint32_t)(se + 127);
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    const float ap = __uint_as_float(r[i]) * inv;                  // |x| / 2^se
                    float y;
                    if (FLOATFMT) {
                        const int pe_b = max((int)(__float_as_uint(ap) >> 23), f.pe_min_b);
                        const uint32_t up = (uint32_t)(f.up_c - pe_b) << 23;
                        y = floorf(__fadd_rn(ap * __uint_as_float(up), 0.5f)) * __uint_as_float(0x7f000000u - up);
                    } else {
                        y = floorf(__fadd_rn(ap * f.up_int, 0.5f)) * f.down_int;
                    }
                    q[i] = fminf(y, f.max_norm);
                }
            } else {
                uint32_t vb[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) vb[i] = i < V ? (r[i < V ? i : 0] | (__float_as_uint(v[i < V ? i : 0]) & 0x80000000u)) : 0u;
                const GeneralOut g = mx_vec_general<V>(make_uint4(vb[0], vb[1], vb[2], vb[3]), make_uint4(vb[4], vb[5], vb[6], vb[7]), amax, p);
                const uint32_t gw[8] = {g.lo.x, g.lo.y, g.lo.z, g.lo.w, g.hi.x, g.hi.y, g.hi.z, g.hi.w};
                scale = g.scale; sbyte = g.sbyte;
#pragma unroll
                for (int i = 0; i < V; ++i) { sgn[i] = gw[i] & 0x80000000u; q[i] = __uint_as_float(gw[i] & 0x7fffffffu); }     // quantize_element already applied sign(.)
            }
            if (OUT == kOutPacked) {
                uint32_t bytes[V / 4];
#pragma unroll
                for (int w = 0; w < V / 4; ++w) {
                    uint32_t o = 0u;
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const float y = q[4 * w + b];
                        uint32_t nb = max(__float_as_uint(y) >> 20, 960u) - 960u;                 // normal E4M3: (e + 7) << 3 | m3; 0 for y = 0
                        if (f.e4m3_sub) {
                            const uint32_t sub = (__float_as_uint(y + 0.015625f) >> 20) - 968u;   // y < 2^-6: m = y / 2^-9 rides in the mantissa of y + 2^-6
                            nb = y < 0.015625f ? sub : nb;
                        }
                        o |= (nb | (sgn[4 * w + b] >> 24)) << (8 * b);                           // nb <= 0x7e (448)
                    }
                    bytes[w] = o;
                }
                const bool live = li < rem;
                const uint32_t vi = (uint32_t)(base + li);
                if (live) {
                    uint8_t* dst = static_cast<uint8_t*>(p.out) + (int64_t)vi * V;
                    if (V == 4) *reinterpret_cast<uint32_t*>(dst) = bytes[0];
                    else *reinterpret_cast<uint2*>(dst) = make_uint2(bytes[0], bytes[V == 8 ? 1 : 0]);
                }
                // scale bytes: the first lane of every 32-element group stores the group's byte at the warp's atom address (computed
                // once per tile by lane u for step u, see sf_off) + its slab (16-bit inputs: a warp spans two) + its group
                constexpr int kLanesPerGroup = 32 / V, kLanesPerSlab = 128 / V;
                const uint32_t off = __shfl_sync(0xffffffffu, sf_off, u);
                if (live && (lane % kLanesPerGroup) == 0)
                    p.sf[off + (uint32_t)(lane / kLanesPerSlab) * f.slab_stride + (uint32_t)((lane % kLanesPerSlab) / kLanesPerGroup)] = (uint8_t)sbyte;
            } else {
#pragma unroll
                for (int i = 0; i < V; ++i) q[i] = __uint_as_float(__float_as_uint(q[i] * scale) | sgn[i]);
                if (li < rem) {
                    if (OUT == kOutFake) {
                        st_stream(static_cast<uint4*>(p.out) + base + li, pack_vec<DT>(q));
                    } else {
                        __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(p.out) + (base + li) * V;
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

template <int DT, int OUT>
static int launch_stream(const ocp::Params& p, const ocp::Fast& f, int64_t n_tiles, cudaStream_t st) {
    using namespace ocp;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(kStreamThreads); cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = tuning().pdl ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e;
    if (p.ebits > 0) {
        cfg.gridDim = dim3((unsigned)stream_grid(kernel_occupancy(mx_stream_kernel<DT, OUT, true>, kStreamThreads), n_tiles));
        e = cudaLaunchKernelEx(&cfg, mx_stream_kernel<DT, OUT, true>, p, f);
    } else {
        cfg.gridDim = dim3((unsigned)stream_grid(kernel_occupancy(mx_stream_kernel<DT, OUT, false>, kStreamThreads), n_tiles));
        e = cudaLaunchKernelEx(&cfg, mx_stream_kernel<DT, OUT, false>, p, f);
    }
    if (e != cudaSuccess) return set_errorf(BFP_E_CUDA, "cudaLaunchKernelEx(mx_stream_kernel): %s", cudaGetErrorString(e));
    return BFP_OK;
}

template <int OUT>
static int launch_mx_quant(ocp::Params& p, int dtype, bool stream_ok, cudaStream_t st) {
    using namespace ocp;
    const int V = dtype == BFP_DT_F32 ? 4 : 8;
    if (stream_ok && p.rows * p.K / V < (int64_t)UINT32_MAX) {
        p.n_vec = p.rows * p.K / V;
        p.lanes_per_block = p.block / V;
        Fast f = {};
        const int drop = 32 - p.bfloat;
        f.bf_half = (p.bfloat > 0 && p.bfloat < 32) ? (1u << (drop - 1)) : 0u;
        f.bf_mask = (p.bfloat > 0 && p.bfloat < 32) ? ~((1u << drop) - 1u) : 0xffffffffu;
        f.pe_min_b = p.min_exp + 127;
        f.up_c = 254 + p.mbits - 2;
        f.up_int = (float)(1 << (p.mbits - 2)); f.down_int = 1.0f / (float)(1 << (p.mbits - 2));
        f.max_norm = p.max_norm; f.emax = p.emax; f.scale_emax = p.scale_emax;
        f.e4m3_sub = (p.ebits == 4 && p.mbits == 5) ? 1 : 0;
        f.vec_per_row = make_fastdiv((uint32_t)(p.K / V));
        if (OUT == kOutPacked) {
            f.tile_rows = make_fastdiv((uint32_t)p.tile_rows);
            const int64_t stride = p.n_row_tiles * p.atoms * 512;
            if (stride * (p.K / 128) >= (int64_t)UINT32_MAX) return set_error(BFP_E_UNSUPPORTED, "MX block-scaled pack: scale array beyond 4 GB");
            f.slab_stride = (uint32_t)stride;
        }
        const int64_t n_tiles = (p.n_vec + kStreamThreads * kStreamUnroll - 1) / (kStreamThreads * kStreamUnroll);
        int rc;
        if (dtype == BFP_DT_F32) rc = launch_stream<BFP_DT_F32, OUT>(p, f, n_tiles, st);
        else if (dtype == BFP_DT_F16) rc = launch_stream<BFP_DT_F16, OUT>(p, f, n_tiles, st);
        else rc = launch_stream<BFP_DT_BF16, OUT>(p, f, n_tiles, st);
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
    return launch_mx_quant<ocp::kOutBf16>(p, dtype, stream_ok && ld_out == K, st);
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
