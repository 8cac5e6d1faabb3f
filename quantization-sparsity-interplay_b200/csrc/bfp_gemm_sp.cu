// bfp_gemm_sp.cu -- the BFP linear with 2:4-sparsified weights on the structured-sparse tensor-core path
// (tcgen05.mma.sp.kind::f16, sm_100a).
//
// The reference prunes the weight 2:4 along K (bfp_ops.py:73-91) and then multiplies the *dense* zero-filled tensor
// (bfp_ops.py:187-190).  Here the pruned weight is stored compressed -- the two kept values of every group of four plus
// a 4-bit index nibble -- and the tensor core skips the zeros: one tcgen05.mma.sp consumes 32 logical k in the time a
// dense kind::f16 MMA consumes 16.  Operands are the exact-bf16 BFP form of bfp_gemm.cu (q * 2^(e-m), exact for
// mant_bits <= 8), so the products are exact and the only rounding is the fp32 accumulation.
//
// The sparse operand of tcgen05.mma.sp is always "A" (the 128 TMEM lanes), so the roles are swapped relative to the
// dense kernel: A := W tile [128 out-features x K/2 compressed], B := X tile [256 tokens x K]; the accumulator holds
// D[n, t] = y[t, n]^T and the epilogue writes it back transposed (lane = n, so a warp stores 32 consecutive n: 128 bytes).
//
// Metadata (E): per 128 rows x 128 logical k one 2 KB "atom" in exactly the byte order tcgen05.cp.128x128b moves into
// four TMEM columns (one column per MMA).  Within an atom, for row m = m0 + 8*m1 + 16*m2 and k = k0 + 16*k1 + 32*k2:
//     lane = m0 + 8*k1 + 16*m2,  column = k2,  bit = 16*m1 + k0      (byte 16*lane + 4*k2 + 2*m1 + k0/8)
// i.e. one 16-bit word per (row, 16 logical k) holding four nibbles idx0 | idx1 << 2 (positions of the two kept values).
// (Layout restated from the public CUTLASS headers: Sm1xxGemmSparseConfig::TensorEAtom_MMA_F16 / UMMA::tmem_e_frg.)
//
// Kernel: persistent, one CTA per SM, 384 threads.  warp 0 = TMA producer (W slab 128 x 64 B SWIZZLE_64B, X slab
// 256 x 128 B SWIZZLE_128B per 64 logical k; the E atom every second slab), warp 1 = MMA issuer (tcgen05.cp of E, two
// MMAs per slab), warp 2 = TMEM allocator, warps 4-11 = epilogue.  TMEM: 256 accumulator columns + a 4-atom ring of E
// columns; the single accumulator is drained into registers in one burst and handed back before the stores, so the
// next tile's MMAs wait only for the tcgen05.ld burst, not for the global stores.
#include <cuda.h>

#include <algorithm>

#include "bfp_internal.h"
#include "bfp_tc.cuh"

namespace bfp {
namespace gemm_sp {

using namespace gemm;

constexpr int BW = 128;                      // out-features (rows of W) per tile = TMEM lanes
constexpr int BT = 256;                      // tokens per tile = accumulator columns
constexpr int KS = 64;                       // logical k per smem slab
constexpr int kStages = 5;
constexpr int kSmemW = BW * (KS / 2) * 2;    // 8 KB: 32 kept bf16 per row
constexpr int kSmemX = BT * KS * 2;          // 32 KB
constexpr int kSmemE = 2048;                 // one E atom (128 logical k), filled on even slabs
constexpr int kStageBytes = kSmemW + kSmemX + kSmemE;          // 43008 = 42 * 1024
constexpr int kSmemTotal = kStages * kStageBytes + 1024 + 1024;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 128 + kEpiWarps * 32;
constexpr int kTmemCols = 512;
constexpr int kTmemE = 256;                  // first E column
constexpr int kERing = 4;                    // E atoms resident in TMEM (see the reuse argument in the MMA warp)
// D = F32, A = B = BF16, K-major, N = BT, M = BW, sparse flag (bit 2), sparsity selector (bits 0-1) = 0
constexpr uint32_t kIdescSp = (1u << 2) | (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BT >> 3) << 17) | ((uint32_t)(BW >> 4) << 24);

struct Params {
    const uint8_t* meta;        // [tiles_w][e_atoms][2048]
    const float* bias;          // [N] or nullptr
    float* out;                 // [T][N]
    int T, N;
    int num_k_slabs;            // ceil(K / 64)
    int e_atoms;                // ceil(K / 128)
    int tiles_w, tiles_t;
};

struct Barriers {
    uint64_t full[kStages];
    uint64_t empty[kStages];
    uint64_t tmem_full;
    uint64_t tmem_empty;
    uint32_t tmem_base;
};

__device__ __forceinline__ void mma_sp_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t tmem_e, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.sp.cta_group::1.kind::f16 [%0], %1, %2, [%3], %4, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(tmem_e), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_cp_128x128b(uint32_t tmem_dst, uint64_t smem_desc) {
    asm volatile("tcgen05.cp.cta_group::1.128x128b [%0], %1;" ::"r"(tmem_dst), "l"(smem_desc) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
bfp_gemm_bf16_sp_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    Barriers* bars = reinterpret_cast<Barriers*>(smem + kStages * kStageBytes);
    auto stage_x = [&](int s) { return smem + s * kStageBytes; };
    auto stage_w = [&](int s) { return smem + s * kStageBytes + kSmemX; };
    auto stage_e = [&](int s) { return smem + s * kStageBytes + kSmemX + kSmemW; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = p.tiles_w * p.tiles_t;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
        mbar_init(&bars->tmem_full, 1);
        mbar_init(&bars->tmem_empty, kEpiWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        // ===================================== TMA producer =====================================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int tw = tile % p.tiles_w, tt = tile / p.tiles_w;      // consecutive CTAs share the (larger) X tile
                const uint8_t* meta_row = p.meta + (size_t)tw * p.e_atoms * kSmemE;
                for (int ks = 0; ks < p.num_k_slabs; ++ks) {
                    mbar_wait(&bars->empty[stage], phase ^ 1);
                    const bool with_e = (ks & 1) == 0;
                    mbar_expect_tx(&bars->full[stage], kSmemW + kSmemX + (with_e ? kSmemE : 0));
                    tma_load_2d(stage_x(stage), &map_x, &bars->full[stage], ks * KS, tt * BT);
                    tma_load_2d(stage_w(stage), &map_w, &bars->full[stage], ks * (KS / 2), tw * BW);
                    if (with_e) bulk_load(stage_e(stage), meta_row + (size_t)(ks >> 1) * kSmemE, kSmemE, &bars->full[stage]);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer =======================================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            uint32_t tile_phase = 0;
            uint32_t eslot = kERing - 1;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&bars->tmem_empty, tile_phase ^ 1);              // epilogue has drained the accumulator
                tc_fence_after();
                for (int ks = 0; ks < p.num_k_slabs; ++ks) {
                    mbar_wait(&bars->full[stage], phase);
                    tc_fence_after();
                    if ((ks & 1) == 0) {
                        // E ring reuse: having seen full[] for this slab means the producer saw empty[] of the slab
                        // kStages uses earlier, i.e. every MMA up to 5 slabs back has completed; the atom this slot held
                        // was last read 7-8 slabs back (kERing atoms x 2 slabs), so the copy cannot overtake a reader.
                        eslot = (eslot + 1) & (kERing - 1);
                        tmem_cp_128x128b(tmem_base + kTmemE + eslot * 4, make_smem_desc_k(smem_u32(stage_e(stage)), 0, 128, 128));
                    }
                    const uint64_t dw = make_smem_desc_k(smem_u32(stage_w(stage)), 4, 512, 16);    // SWIZZLE_64B: 8 rows x 64 B
                    const uint64_t dx = make_smem_desc(smem_u32(stage_x(stage)));                  // SWIZZLE_128B
                    const uint32_t ecol = tmem_base + kTmemE + eslot * 4 + (uint32_t)(ks & 1) * 2;
#pragma unroll
                    for (int i = 0; i < 2; ++i)      // 32 logical k per MMA: 16 kept bf16 = 32 B of W (+2), 64 B of X (+4)
                        // the metadata address names an even column; the sparsity selector (idesc bits 0-1) picks the odd one
                        mma_sp_bf16(tmem_base, dw + (uint64_t)(i * 2), dx + (uint64_t)(i * 4), ecol, kIdescSp | (uint32_t)i, (ks | i) != 0);
                    tc_commit(&bars->empty[stage]);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                tc_commit(&bars->tmem_full);
                tile_phase ^= 1;
            }
        }
    } else if (warp >= 4) {
        // ===================================== epilogue =========================================
        const int ew = warp - 4, q = warp & 3, half = ew >> 2;
        const int n_in_tile = q * 32 + lane;
        uint32_t tile_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int tw = tile % p.tiles_w, tt = tile / p.tiles_w;
            mbar_wait(&bars->tmem_full, tile_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * (BT / 2));
            uint32_t r[BT / 2];
#pragma unroll
            for (int c = 0; c < BT / 32; ++c) tmem_ld16(taddr + c * 16, r + c * 16);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->tmem_empty);                  // accumulator is free again
            tile_phase ^= 1;

            const int n = tw * BW + n_in_tile;
            const int t0 = tt * BT + half * (BT / 2);
            if (n < p.N) {
                const float bv = p.bias ? p.bias[n] : 0.0f;
                float* dst = p.out + (int64_t)t0 * p.N + n;
                const int t_left = p.T - t0;
#pragma unroll
                for (int j = 0; j < BT / 2; ++j)
                    if (j < t_left) dst[(int64_t)j * p.N] = __uint_as_float(r[j]) + bv;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}

// ---- 2:4 compressor ------------------------------------------------------------------------------------------------
// One thread per (row, 16 logical k): reads 16 bf16 (two 128-bit loads), writes the 8 kept values (one 128-bit store)
// and the 16-bit metadata word.  Groups with fewer than two non-zeros are padded with a zero position (indices stay
// increasing); a group with more than two non-zeros is not 2:4 -- counted in *violations, first two kept.
__global__ void __launch_bounds__(256)
compress_2to4_bf16_kernel(const uint16_t* __restrict__ w, int64_t ld_w, uint16_t* __restrict__ comp, int64_t ld_c,
                          uint8_t* __restrict__ meta, int rows, int K, int e_atoms, unsigned int* __restrict__ violations) {
    const int halves = e_atoms * 8;                        // 16-k units per row (K padded to 128)
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)rows * halves;
    if (idx >= total) return;
    const int row = (int)(idx / halves), h = (int)(idx % halves);
    const int k_base = h * 16;
    uint16_t v[16];
    if (k_base + 16 <= K) {
        const uint4* src = reinterpret_cast<const uint4*>(w + (int64_t)row * ld_w + k_base);
        const uint4 a = src[0], b = src[1];
        const uint32_t u[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) { v[2 * i] = (uint16_t)(u[i] & 0xffffu); v[2 * i + 1] = (uint16_t)(u[i] >> 16); }
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = (k_base + i < K) ? w[(int64_t)row * ld_w + k_base + i] : (uint16_t)0;
    }
    uint16_t kept[8];
    uint32_t word = 0, bad = 0;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        int i0 = -1, i1 = -1, cnt = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if ((v[4 * g + j] & 0x7fffu) != 0) {           // -0.0 counts as zero
                if (cnt == 0) i0 = j; else if (cnt == 1) i1 = j;
                ++cnt;
            }
        }
        if (cnt > 2) bad = 1;
        if (cnt == 0) { i0 = 0; i1 = 1; }
        else if (cnt == 1) { if (i0 == 3) { i1 = 3; i0 = 0; } else i1 = 3; }
        kept[2 * g] = v[4 * g + i0];
        kept[2 * g + 1] = v[4 * g + i1];
        word |= (uint32_t)(i0 | (i1 << 2)) << (4 * g);
    }
    if (bad) atomicAdd(violations, 1u);
    uint4 o;
    o.x = kept[0] | ((uint32_t)kept[1] << 16); o.y = kept[2] | ((uint32_t)kept[3] << 16);
    o.z = kept[4] | ((uint32_t)kept[5] << 16); o.w = kept[6] | ((uint32_t)kept[7] << 16);
    *reinterpret_cast<uint4*>(comp + (int64_t)row * ld_c + h * 8) = o;
    // metadata position (see the file header)
    const int m = row & 127, tile_w = row >> 7;
    const int m0 = m & 7, m1 = (m >> 3) & 1, m2 = m >> 4;
    const int atom = h >> 3, k1 = h & 1, k2 = (h >> 1) & 3;
    const int lane = m0 + 8 * k1 + 16 * m2;
    uint8_t* dst = meta + ((size_t)tile_w * e_atoms + atom) * 2048 + lane * 16 + k2 * 4 + m1 * 2;
    *reinterpret_cast<uint16_t*>(dst) = (uint16_t)word;
}

}  // namespace gemm_sp

int sp_layout(int64_t rows, int64_t K, int64_t* Kc, int64_t* meta_bytes) {
    if (rows < 0 || K < 0) return set_error(BFP_E_ARG, "negative dimension");
    const int64_t atoms = round_up(K, 128) / 128;
    if (Kc) *Kc = atoms * 64;
    if (meta_bytes) *meta_bytes = round_up(rows, 128) / 128 * atoms * 2048;
    return BFP_OK;
}

int compress_2to4_bf16_device(const void* w_bf16, int64_t rows, int64_t K, int64_t ld_w, void* comp, void* meta, unsigned int* violations,
                              cudaStream_t st) {
    using namespace gemm_sp;
    if (rows == 0 || K == 0) return BFP_OK;
    if (K % 8 != 0 || ld_w % 8 != 0 || ld_w < K) return set_error(BFP_E_ARG, "bf16 operand K and row stride must be multiples of 8");
    if (rows > INT32_MAX || K > INT32_MAX) return set_error(BFP_E_ARG, "dimension too large");
    if (reinterpret_cast<uintptr_t>(w_bf16) % 16 || reinterpret_cast<uintptr_t>(comp) % 16 || reinterpret_cast<uintptr_t>(meta) % 16)
        return set_error(BFP_E_ALIGN, "operands must be 16-byte aligned");
    int64_t Kc, mb;
    sp_layout(rows, K, &Kc, &mb);
    // padding rows of the last 128-row tile and every slot the kernel does not touch: valid "keep 0,1" nibbles
    cudaError_t e = cudaMemsetAsync(meta, 0x44, (size_t)mb, st);
    if (e != cudaSuccess) return set_errorf(BFP_E_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    const int e_atoms = (int)(Kc / 64);
    const int64_t total = rows * e_atoms * 8;
    compress_2to4_bf16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        static_cast<const uint16_t*>(w_bf16), ld_w, static_cast<uint16_t*>(comp), Kc, static_cast<uint8_t*>(meta), (int)rows, (int)K, e_atoms,
        violations);
    count_launch();
    return check_launch("compress_2to4_bf16_kernel");
}

int gemm_bf16_sp_device(const void* x_bf16, const void* w_comp, const void* w_meta, const float* bias, float* out, int64_t T, int64_t N,
                        int64_t Kp, cudaStream_t st) {
    using namespace gemm_sp;
    if (T == 0 || N == 0) return BFP_OK;
    if (Kp % 8 != 0 || Kp <= 0) return set_error(BFP_E_ARG, "bf16 operand K must be a positive multiple of 8");
    if (T > INT32_MAX || N > INT32_MAX || Kp > INT32_MAX) return set_error(BFP_E_ARG, "dimension too large");
    if (reinterpret_cast<uintptr_t>(x_bf16) % 16 || reinterpret_cast<uintptr_t>(w_comp) % 16 || reinterpret_cast<uintptr_t>(w_meta) % 16)
        return set_error(BFP_E_ALIGN, "operands must be 16-byte aligned");
    int64_t Kc, mb;
    sp_layout(N, Kp, &Kc, &mb);
    Params p;
    p.meta = static_cast<const uint8_t*>(w_meta); p.bias = bias; p.out = out; p.T = (int)T; p.N = (int)N;
    p.num_k_slabs = (int)((Kp + KS - 1) / KS);
    p.e_atoms = (int)(Kc / 64);
    p.tiles_w = (int)((N + BW - 1) / BW);
    p.tiles_t = (int)((T + BT - 1) / BT);
    CUtensorMap map_w, map_x;
    if (int rc = make_map_bf16(&map_w, w_comp, N, Kc, Kc * 2, KS / 2, BW, CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
    if (int rc = make_map_bf16(&map_x, x_bf16, T, Kp, Kp * 2, KS, BT, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    const cudaError_t e = cudaFuncSetAttribute(bfp_gemm_bf16_sp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal);
    if (e != cudaSuccess) return set_errorf(BFP_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    const int grid = std::min(p.tiles_w * p.tiles_t, std::max(1, device_info().sm_count));
    bfp_gemm_bf16_sp_kernel<<<grid, kThreads, kSmemTotal, st>>>(map_w, map_x, p);
    count_launch();
    return check_launch("bfp_gemm_bf16_sp_kernel");
}

}  // namespace bfp
