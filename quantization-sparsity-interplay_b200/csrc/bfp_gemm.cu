// bfp_gemm.cu -- the BFP linear contraction on 5th-gen tensor cores (tcgen05, sm_100a).
//
// Replaces the reference's dequantise-then-fp32-GEMM (bfp_ops.py:187-190: F.linear on fake-quantised tensors) by an
// exact restatement on PACKED operands:
//     y[t,n] = bias[n] + sum_kb  sa[kb][t] * sb[kb][n] * ( sum_{k in kb} qa[t,k] * qb[n,k] )
// The inner sums are exact integers (tcgen05.mma.kind::i8, int32 accumulators in TMEM), the scales are exact powers of
// two; only the outer fp32 accumulation over K/B blocks rounds.
//
// Kernel shape (one persistent CTA per SM, 384 threads, warp-specialised):
//   warp 0      TMA producer: per 128-byte K slab, one 2-D tensor copy each of the A tile [128 x 128 B] and the B tile
//               [256 x 128 B] (SWIZZLE_128B) plus 1-D bulk copies of the slab's block scales, all landing on one mbarrier.
//   warp 1      MMA issuer (one elected lane): 128x256x32 tcgen05.mma.kind::i8 per instruction; every BFP block starts
//               a fresh accumulator (enable-input-d = 0) in one of TWO 256-column TMEM buffers and is committed to an
//               mbarrier, so block kb+1 multiplies while block kb is rescaled.
//   warp 2      TMEM allocator (512 columns).
//   warps 4-11  epilogue: tcgen05.ld the int32 tile (warp w reads TMEM lanes 32*(w%4).., 128 columns per warp), convert,
//               FMA with sa[t]*sb[n] into 128 fp32 registers per thread; after the last block add bias and store.
// Barriers: full/empty per smem stage (empty = MMA commit + 8 epilogue warps, because the scales live in the stage),
//           tmem_full/tmem_empty per accumulator buffer.
#include <cuda.h>
#include <cstdio>

#include <algorithm>
#include <mutex>

#include "bfp_internal.h"
#include "bfp_tc.cuh"

namespace bfp {

namespace gemm {

constexpr int BM = 128, BN = 256, BKB = 128;           // tile: rows of A, rows of B, K bytes per stage (int8: 128 k)
constexpr int UMMA_K = 32;                             // k per tcgen05.mma.kind::i8
constexpr int kStages = 4;
constexpr int kMaxBlocksPerStage = 4;                  // B >= 32
constexpr int kEpiWarps = 8;                           // bf16 kernel: the epilogue runs once per tile
constexpr int kThreads = 128 + kEpiWarps * 32;         // 384
constexpr int kEpiWarpsI8 = 8;                         // int8 kernel: 8 or 16 epilogue warps (16 measured ~6 % slower: the
constexpr int kThreadsI8 = 128 + kEpiWarpsI8 * 32;     // rescale is pipe-throughput-bound, not latency-bound)
constexpr int kColsI8 = BN / (kEpiWarpsI8 / 4);        // accumulator columns per epilogue thread: 128 (or 64)
// Registers: the CTA's launch allocation is the pool setmaxnreg can redistribute (an .inc beyond it blocks forever):
//   384 threads x 168 = 64512 >= 128 x 80 + 256 x 208 = 63488;   640 threads x 96 = 61440 >= 128 x 40 + 512 x 104 = 58368.
constexpr int kRegDecI8 = kEpiWarpsI8 == 8 ? 80 : 40, kRegIncI8 = kEpiWarpsI8 == 8 ? 208 : 104;
constexpr int kTmemCols = 512;

constexpr int kSmemA = BM * BKB;                       // 16 KB
constexpr int kSmemB = BN * BKB;                       // 32 KB
constexpr int kSmemScaleA = kMaxBlocksPerStage * BM * 4;   // 2 KB
constexpr int kSmemScaleB = kMaxBlocksPerStage * BN * 4;   // 4 KB
constexpr int kStageBytes = kSmemA + kSmemB + kSmemScaleA + kSmemScaleB;   // 55296 (multiple of 1024)
constexpr int kSmemBarriers = 1024;
constexpr int kSmemTotal = kStages * kStageBytes + kSmemBarriers + 1024;   // + alignment slack

struct Params {
    const float* scale_a;       // [nkb_pad][lda_s]
    const float* scale_b;       // [nkb_pad][ldb_s]
    const float* bias;          // [N] or nullptr
    float* out;                 // [T][N]
    int64_t lda_s, ldb_s;
    int T, N, K;                // K = padded mant row length (multiple of 16)
    int blocks_per_stage;       // 128 / B
    int mmas_per_block;         // B / 32
    int num_k_stages;           // ceil(K / 128)
    int tiles_m, tiles_n;
};

// instruction descriptor (InstrDescriptor): D = S32, A = B = signed int8, both K-major, N >> 3, M >> 4
constexpr uint32_t kIdescI8 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

struct Barriers {
    uint64_t full[kStages];
    uint64_t empty[kStages];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(kThreadsI8, 1)
bfp_gemm_i8_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    Barriers* bars = reinterpret_cast<Barriers*>(smem + kStages * kStageBytes);
    auto stage_a = [&](int s) { return smem + s * kStageBytes; };
    auto stage_b = [&](int s) { return smem + s * kStageBytes + kSmemA; };
    auto stage_sa = [&](int s) { return reinterpret_cast<float*>(smem + s * kStageBytes + kSmemA + kSmemB); };
    auto stage_sb = [&](int s) { return reinterpret_cast<float*>(smem + s * kStageBytes + kSmemA + kSmemB + kSmemScaleA); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = p.tiles_m * p.tiles_n;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1 + kEpiWarpsI8); }
        for (int b = 0; b < 2; ++b) { mbar_init(&bars->tmem_full[b], 1); mbar_init(&bars->tmem_empty[b], kEpiWarpsI8); }
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

    const uint32_t scale_bytes_a = (uint32_t)p.blocks_per_stage * BM * 4, scale_bytes_b = (uint32_t)p.blocks_per_stage * BN * 4;
    const uint32_t stage_tx = kSmemA + kSmemB + scale_bytes_a + scale_bytes_b;

    if (warp == 0) {
        // ===================================== TMA producer =====================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegDecI8));
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int tm = tile % p.tiles_m, tn = tile / p.tiles_m;       // consecutive CTAs share the B tile
                for (int ks = 0; ks < p.num_k_stages; ++ks) {
                    mbar_wait(&bars->empty[stage], phase ^ 1);
                    mbar_expect_tx(&bars->full[stage], stage_tx);
                    tma_load_2d(stage_a(stage), &map_a, &bars->full[stage], ks * BKB, tm * BM);
                    tma_load_2d(stage_b(stage), &map_b, &bars->full[stage], ks * BKB, tn * BN);
                    for (int b = 0; b < p.blocks_per_stage; ++b) {
                        const int64_t kb = (int64_t)ks * p.blocks_per_stage + b;
                        bulk_load(stage_sa(stage) + b * BM, p.scale_a + kb * p.lda_s + (int64_t)tm * BM, BM * 4, &bars->full[stage]);
                        bulk_load(stage_sb(stage) + b * BN, p.scale_b + kb * p.ldb_s + (int64_t)tn * BN, BN * 4, &bars->full[stage]);
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer =======================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegDecI8));
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            int buf = 0; uint32_t buf_phases = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                for (int ks = 0; ks < p.num_k_stages; ++ks) {
                    mbar_wait(&bars->full[stage], phase);
                    tc_fence_after();
                    const uint64_t da = make_smem_desc(smem_u32(stage_a(stage)));
                    const uint64_t db = make_smem_desc(smem_u32(stage_b(stage)));
                    int mma = 0;
                    for (int b = 0; b < p.blocks_per_stage; ++b) {
                        mbar_wait(&bars->tmem_empty[buf], ((buf_phases >> buf) & 1u) ^ 1u);   // epilogue drained this accumulator
                        tc_fence_after();
                        const uint32_t d = tmem_base + (uint32_t)buf * BN;
                        for (int i = 0; i < p.mmas_per_block; ++i, ++mma) {
                            // advance both descriptors by 32 bytes of K inside the 128-byte swizzle atom (+2 in 16-B units)
                            mma_i8(d, da + (uint64_t)(mma * (UMMA_K >> 4)), db + (uint64_t)(mma * (UMMA_K >> 4)), kIdescI8, i > 0);
                        }
                        tc_commit(&bars->tmem_full[buf]);                          // accumulator of this BFP block is complete
                        buf_phases ^= 1u << buf;
                        buf ^= 1;
                    }
                    tc_commit(&bars->empty[stage]);                                // smem slab consumed by the tensor core
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegDecI8));        // warps 2, 3: same warpgroup as the producer / issuer
    } else {
        // ===================================== epilogue =========================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegIncI8));
        const int ew = warp - 4;                    // 0..kEpiWarpsI8-1
        const int q = warp & 3;                     // TMEM lane quarter this warp may access
        const int cg = ew >> 2;                     // which kColsI8 accumulator columns
        const int row_in_tile = q * 32 + lane;
        int stage = 0; uint32_t phase = 0;
        int buf = 0; uint32_t buf_phases = 0;       // bit b = parity of accumulator buffer b
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int tm = tile % p.tiles_m, tn = tile / p.tiles_m;
            uint64_t acc2[kColsI8 / 2];             // fp32 accumulators as f32x2 pairs
#pragma unroll
            for (int i = 0; i < kColsI8 / 2; ++i) acc2[i] = 0ull;
            for (int ks = 0; ks < p.num_k_stages; ++ks) {
                mbar_wait(&bars->full[stage], phase);                              // scales of this slab have landed
                for (int b = 0; b < p.blocks_per_stage; ++b) {
                    const float sa = lds32(smem_u32(stage_sa(stage) + b * BM + row_in_tile));
                    const uint64_t sa2 = pack2(sa, sa);
                    const uint32_t sb_addr = smem_u32(stage_sb(stage) + b * BN + cg * kColsI8);
                    mbar_wait(&bars->tmem_full[buf], (buf_phases >> buf) & 1u);
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + cg * kColsI8);
                    // software pipeline: the load of chunk c+1 is in flight while chunk c is rescaled
                    uint32_t r0[16], r1[16];
                    tmem_ld16(taddr, r0);
#pragma unroll
                    for (int c = 0; c < kColsI8 / 16; c += 2) {
                        tmem_ld_wait();
                        tmem_ld16(taddr + (c + 1) * 16, r1);
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 sv = lds128(sb_addr + (c * 4 + (j >> 2)) * 16);
                            const uint64_t w01 = mul2(sa2, pack2(sv.x, sv.y)), w23 = mul2(sa2, pack2(sv.z, sv.w));
                            rescale_pair_xu(acc2[c * 8 + (j >> 1)], r0[j], r0[j + 1], w01);
                            rescale_pair_xu(acc2[c * 8 + (j >> 1) + 1], r0[j + 2], r0[j + 3], w23);
                        }
                        tmem_ld_wait();
                        if (c + 2 < kColsI8 / 16) tmem_ld16(taddr + (c + 2) * 16, r0);
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 sv = lds128(sb_addr + ((c + 1) * 4 + (j >> 2)) * 16);
                            const uint64_t w01 = mul2(sa2, pack2(sv.x, sv.y)), w23 = mul2(sa2, pack2(sv.z, sv.w));
                            rescale_pair_xu(acc2[(c + 1) * 8 + (j >> 1)], r1[j], r1[j + 1], w01);
                            rescale_pair_xu(acc2[(c + 1) * 8 + (j >> 1) + 1], r1[j + 2], r1[j + 3], w23);
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars->tmem_empty[buf]);
                    buf_phases ^= 1u << buf;
                    buf ^= 1;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->empty[stage]);                   // done with this slab's scales
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            float acc[kColsI8];
#pragma unroll
            for (int i = 0; i < kColsI8 / 2; ++i) { const float2 v = unpack2(acc2[i]); acc[2 * i] = v.x; acc[2 * i + 1] = v.y; }
            // bias + store: thread owns row t, kColsI8 consecutive columns
            const int t = tm * BM + row_in_tile;
            const int n0 = tn * BN + cg * kColsI8;
            if (t < p.T) {
                float* dst = p.out + (int64_t)t * p.N + n0;
                const bool vec_ok = (p.N % 4 == 0) && (n0 + kColsI8 <= p.N);
                if (vec_ok) {
#pragma unroll
                    for (int j = 0; j < kColsI8; j += 4) {
                        float4 o = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
                        if (p.bias) {
                            const float4 bv = *reinterpret_cast<const float4*>(p.bias + n0 + j);
                            o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
                        }
                        *reinterpret_cast<float4*>(dst + j) = o;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < kColsI8; ++j)
                        if (n0 + j < p.N) dst[j] = acc[j] + (p.bias ? p.bias[n0 + j] : 0.0f);
                }
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

// ================================================================================================================
// Exact bf16 path.  For mant_bits <= 8 the dequantised value q * 2^(e-m) is exactly representable in bf16 (q has at most
// 8 significant bits, the power of two only moves the exponent), so products are exact in the tensor core and the only
// rounding is the fp32 accumulation -- the same contract as the int8 path, but with NO per-block rescale: the scales
// ride in the operands' exponents.  MMA-bound for every block size (including B = 16, which kind::i8 cannot express).
// Same tile / smem-byte geometry as the int8 kernel: stage = A[128 x 128 B] + B[256 x 128 B], 64 bf16 of K per stage,
// four 128x256x16 tcgen05.mma.kind::f16 per stage; accumulators ping-pong between two TMEM buffers ACROSS TILES so the
// epilogue of tile i overlaps the main loop of tile i+1.
// ================================================================================================================
// Two tile widths: 128x256 (4 stages of 48 KB, the default) and 128x128 (6 stages of 32 KB; for N <= 128 or forced with
// bfp_set_option("gemm_bf16_tile_n", 128)).
// CG = 2: a cluster pair shares a 256 x 256 tile through tcgen05.mma.cta_group::2 (each CTA stages 128 rows of X and 128
// rows of W per slab: 32 KB instead of 48 KB for the same MMA time, so the operand stream per SM drops by a third).
template <int TBN, int CG> struct Bf16Cfg {
    static constexpr int kRowsB = TBN / CG;                 // W rows staged by one CTA
    static constexpr int kStageBytes = kSmemA + kRowsB * BKB;
    static constexpr int kStages = CG == 2 ? 6 : (TBN == 256 ? 4 : 6);
    static constexpr int kSmemStaging = 8 * 4096;                            // per epilogue warp one [32 t][32 n] fp32 tile (SWIZZLE_128B)
    static constexpr int kSmemTotal = kStages * kStageBytes + kSmemStaging + kSmemBarriers + 1024;
    // D = F32 (1 << 4), A = B = BF16 (1 << 7, 1 << 10), K-major, N >> 3, M >> 4 (M = 128 per CTA)
    static constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TBN >> 3) << 17) | ((uint32_t)((BM * CG) >> 4) << 24);
    static constexpr int kColsPerThread = TBN / 2;          // 8 epilogue warps: 4 lane quarters x 2 column halves
};

struct ParamsBf16 {
    const float* bias;
    float* out;
    int out_dtype;              // BFP_DT_F32, or F16 / BF16: accumulator (+ bias) rounded once in the epilogue
    int out_tma;                // 1 = epilogue writes through smem + TMA stores (needs 16-byte aligned rows)
    int accumulate;             // 1 = out += product (fp32, TMA path only): TMA reduce-add instead of a store
    int batch;                  // > 1: `batch` independent products; operands are stacked along rows ([batch * T, K], [batch * N, K]),
                                // T / N / tiles_* are per entry, the output map is 3-D (rows clipped per entry); TMA path only
    int T, N;
    int num_k_stages;           // ceil(K / 64)
    int tiles_m, tiles_n;       // tiles_m counts CG * 128 rows
};

template <int CG>
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    if constexpr (CG == 1)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

struct BarriersBf16 {
    uint64_t full[6];
    uint64_t empty[6];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
};

template <int TBN, int CG>
__global__ void __launch_bounds__(kThreads, 1)
bfp_gemm_bf16_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                     const __grid_constant__ CUtensorMap map_out, const ParamsBf16 p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    using C = Bf16Cfg<TBN, CG>;
    constexpr int kStagesBf16 = C::kStages, kStageBytesBf16 = C::kStageBytes, kCols = C::kColsPerThread;
    uint8_t* staging = smem + kStagesBf16 * kStageBytesBf16;                  // 1024-aligned (stage sizes are multiples of 1024)
    BarriersBf16* bars = reinterpret_cast<BarriersBf16*>(staging + C::kSmemStaging);
    auto stage_a = [&](int s) { return smem + s * kStageBytesBf16; };
    auto stage_b = [&](int s) { return smem + s * kStageBytesBf16 + kSmemA; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = CG == 1 ? 0u : cluster_ctarank();
    const int unit = (int)blockIdx.x / CG, num_units = (int)gridDim.x / CG;
    const int tiles_per = p.tiles_m * p.tiles_n, num_tiles = tiles_per * p.batch;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStagesBf16; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&bars->tmem_full[b], 1); mbar_init(&bars->tmem_empty[b], kEpiWarps * CG); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        if constexpr (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(kTmemCols));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(kTmemCols));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        }
    }
    tc_fence_before();
    if constexpr (CG == 1) __syncthreads(); else cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            const uint64_t pol_keep = l2_policy_evict_last();                 // operands are re-read by other tiles; the output is not
            for (int tile = unit; tile < num_tiles; tile += num_units) {
                const int bt = tile / tiles_per, rem = tile - bt * tiles_per;
                const int tm = rem % p.tiles_m, tn = rem / p.tiles_m;
                // batched: a tile that runs past its entry's rows reads the next entry's (or zero fill); the output map clips them
                const int a_row = bt * p.T + (tm * CG + (int)rank) * BM;      // this CTA's 128 X rows
                const int b_row = bt * p.N + tn * TBN + (int)rank * C::kRowsB;   // this CTA's share of the W tile
                for (int ks = 0; ks < p.num_k_stages; ++ks) {
                    mbar_wait(&bars->empty[stage], phase ^ 1);
                    if (rank == 0) mbar_expect_tx(&bars->full[stage], (uint32_t)CG * kStageBytesBf16);   // both CTAs' bytes land on the leader
                    const uint32_t bar = CG == 1 ? smem_u32(&bars->full[stage]) : mapa_u32(smem_u32(&bars->full[stage]), 0);
                    tma_load_2d_to_hint<CG>(stage_a(stage), &map_a, bar, ks * 64, a_row, pol_keep);
                    tma_load_2d_to_hint<CG>(stage_b(stage), &map_b, bar, ks * 64, b_row, pol_keep);
                    if (++stage == kStagesBf16) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            int stage = 0; uint32_t phase = 0;
            int buf = 0; uint32_t buf_phase[2] = {0, 0};
            for (int tile = unit; tile < num_tiles; tile += num_units) {
                mbar_wait(&bars->tmem_empty[buf], buf_phase[buf] ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)buf * TBN;
                for (int ks = 0; ks < p.num_k_stages; ++ks) {
                    mbar_wait(&bars->full[stage], phase);
                    tc_fence_after();
                    const uint64_t da = make_smem_desc(smem_u32(stage_a(stage)));
                    const uint64_t db = make_smem_desc(smem_u32(stage_b(stage)));
#pragma unroll
                    for (int i = 0; i < 4; ++i)         // 16 bf16 = 32 bytes of K per MMA: +2 in 16-byte units
                        mma_bf16<CG>(d, da + (uint64_t)(i * 2), db + (uint64_t)(i * 2), C::kIdesc, (ks | i) != 0);
                    commit<CG>(&bars->empty[stage]);
                    if (++stage == kStagesBf16) { stage = 0; phase ^= 1; }
                }
                commit<CG>(&bars->tmem_full[buf]);
                buf_phase[buf] ^= 1;
                buf ^= 1;
            }
        }
    } else if (warp >= 4) {
        const int ew = warp - 4, q = warp & 3, half = ew >> 2;
        const int row_in_tile = q * 32 + lane;
        int buf = 0; uint32_t buf_phase[2] = {0, 0};
        for (int tile = unit; tile < num_tiles; tile += num_units) {
            const int bt = tile / tiles_per, rem = tile - bt * tiles_per;
            const int tm = rem % p.tiles_m, tn = rem / p.tiles_m;
            mbar_wait(&bars->tmem_full[buf], buf_phase[buf]);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * TBN + half * kCols);
            const int t = (tm * CG + (int)rank) * BM + row_in_tile;
            const int n0 = tn * TBN + half * kCols;
            float* dst = p.out + (int64_t)t * p.N + n0;
            const bool vec_ok = (p.N % 4 == 0) && (n0 + kCols <= p.N);
            if (p.out_tma) {
                // 32 columns at a time: the warp's [32 t][32 n] block goes to a swizzled smem tile (16-byte chunk c of row t at
                // chunk c ^ (t & 7): conflict-free 128-bit shared stores) and leaves through one TMA store of full 128-byte rows;
                // rows / columns beyond T / N are clipped by the copy engine.
                const uint64_t pol = l2_policy_evict_first();
                uint8_t* sbuf = staging + ew * 4096;
                if (p.out_dtype != BFP_DT_F32) {
                // half outputs: 64 columns per round = one 128-byte row of 2-byte values per token, same swizzled tile
#pragma unroll
                for (int c = 0; c < kCols / 64; ++c) {
                    uint32_t r[64];
#pragma unroll
                    for (int i = 0; i < 4; ++i) tmem_ld16(taddr + c * 64 + i * 16, r + i * 16);
                    tmem_ld_wait();
                    if (lane == 0) tma_store_wait_read<0>();
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j) {                  // chunk j = columns 8j .. 8j+7
                        uint32_t w[4];
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            const int nb = n0 + c * 64 + 8 * j + 2 * h;
                            float o0 = __uint_as_float(r[8 * j + 2 * h]), o1 = __uint_as_float(r[8 * j + 2 * h + 1]);
                            if (p.bias) { if (nb < p.N) o0 += p.bias[nb]; if (nb + 1 < p.N) o1 += p.bias[nb + 1]; }
                            if (p.out_dtype == BFP_DT_F16) { const __half2 hh = __floats2half2_rn(o0, o1); w[h] = *reinterpret_cast<const uint32_t*>(&hh); }
                            else { const __nv_bfloat162 hh = __floats2bfloat162_rn(o0, o1); w[h] = *reinterpret_cast<const uint32_t*>(&hh); }
                        }
                        *reinterpret_cast<uint4*>(sbuf + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        if (p.batch > 1) tma_store_3d_hint(&map_out, sbuf, n0 + c * 64, (tm * CG + (int)rank) * BM + q * 32, bt, pol);
                        else tma_store_2d_hint(&map_out, sbuf, n0 + c * 64, (tm * CG + (int)rank) * BM + q * 32, pol);
                        tma_store_commit();
                    }
                }
                } else {
#pragma unroll
                for (int c = 0; c < kCols / 32; ++c) {
                    uint32_t r[32];
                    tmem_ld16(taddr + c * 32, r);
                    tmem_ld16(taddr + c * 32 + 16, r + 16);
                    tmem_ld_wait();
                    if (lane == 0) tma_store_wait_read<0>();                  // single staging tile: the previous store must have read it

                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 o = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
                        if (p.bias) {
                            const int nb = n0 + c * 32 + 4 * j;
                            if (nb + 3 < p.N) {
                                const float4 bv = *reinterpret_cast<const float4*>(p.bias + nb);
                                o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
                            } else {
                                if (nb < p.N) o.x += p.bias[nb];
                                if (nb + 1 < p.N) o.y += p.bias[nb + 1];
                                if (nb + 2 < p.N) o.z += p.bias[nb + 2];
                            }
                        }
                        *reinterpret_cast<float4*>(sbuf + lane * 128 + ((j ^ (lane & 7)) << 4)) = o;
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        if (p.batch > 1) tma_store_3d_hint(&map_out, sbuf, n0 + c * 32, (tm * CG + (int)rank) * BM + q * 32, bt, pol);
                        else tma_store_or_add_2d(&map_out, sbuf, n0 + c * 32, (tm * CG + (int)rank) * BM + q * 32, pol, p.accumulate);
                        tma_store_commit();
                    }
                }
                }
            } else {
#pragma unroll
            for (int c = 0; c < kCols / 16; ++c) {
                uint32_t r[16];
                tmem_ld16(taddr + c * 16, r);
                tmem_ld_wait();
                if (t < p.T && p.out_dtype != BFP_DT_F32) {
                    uint16_t* hd = reinterpret_cast<uint16_t*>(p.out) + (int64_t)t * p.N + n0 + c * 16;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int nn = n0 + c * 16 + j;
                        if (nn < p.N) {
                            const float o = __uint_as_float(r[j]) + (p.bias ? p.bias[nn] : 0.0f);
                            hd[j] = p.out_dtype == BFP_DT_F16 ? __half_as_ushort(__float2half_rn(o)) : __bfloat16_as_ushort(__float2bfloat16_rn(o));
                        }
                    }
                } else if (t < p.T) {
                    if (vec_ok) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            float4 o = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                            if (p.bias) {
                                const float4 bv = *reinterpret_cast<const float4*>(p.bias + n0 + c * 16 + j);
                                o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
                            }
                            *reinterpret_cast<float4*>(dst + c * 16 + j) = o;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (n0 + c * 16 + j < p.N) dst[c * 16 + j] = __uint_as_float(r[j]) + (p.bias ? p.bias[n0 + c * 16 + j] : 0.0f);
                    }
                }
            }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (CG == 1) mbar_arrive(&bars->tmem_empty[buf]);
                else mbar_arrive_cluster(mapa_u32(smem_u32(&bars->tmem_empty[buf]), 0));
            }
            buf_phase[buf] ^= 1;
            buf ^= 1;
        }
        if (p.out_tma && lane == 0) tma_store_wait_all<0>();                 // the staging tiles must outlive the last store
    }

    tc_fence_before();
    if constexpr (CG == 1) __syncthreads(); else cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        if constexpr (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}

}  // namespace gemm

int gemm_i8_device(const int8_t* a_mant, const float* a_scale_t, int64_t lda_s, const int8_t* b_mant, const float* b_scale_t,
                   int64_t ldb_s, const float* bias, float* out, int64_t T, int64_t N, int64_t Kp, int block_size, cudaStream_t st) {
    using namespace gemm;
    if (T == 0 || N == 0) return BFP_OK;
    if (block_size < 32 || block_size > 128 || (block_size & (block_size - 1)))
        return set_error(BFP_E_UNSUPPORTED, "bfp_gemm_i8 supports block_size 32, 64, 128 (one MMA is 32 deep)");
    if (Kp % 16 != 0 || Kp <= 0) return set_error(BFP_E_ARG, "packed K must be a positive multiple of 16");
    if (T > INT32_MAX || N > INT32_MAX || Kp > INT32_MAX) return set_error(BFP_E_ARG, "dimension too large");
    if (reinterpret_cast<uintptr_t>(a_mant) % 16 || reinterpret_cast<uintptr_t>(b_mant) % 16 ||
        reinterpret_cast<uintptr_t>(a_scale_t) % 16 || reinterpret_cast<uintptr_t>(b_scale_t) % 16 || lda_s % 4 || ldb_s % 4)
        return set_error(BFP_E_ALIGN, "packed operands must be 16-byte aligned");
    Params p;
    p.scale_a = a_scale_t; p.scale_b = b_scale_t; p.bias = bias; p.out = out; p.lda_s = lda_s; p.ldb_s = ldb_s;
    p.T = (int)T; p.N = (int)N; p.K = (int)Kp;
    p.blocks_per_stage = BKB / block_size; p.mmas_per_block = block_size / UMMA_K;
    p.num_k_stages = (int)((Kp + BKB - 1) / BKB);
    p.tiles_m = (int)((T + BM - 1) / BM); p.tiles_n = (int)((N + BN - 1) / BN);
    if ((int64_t)p.tiles_m * BM > lda_s || (int64_t)p.tiles_n * BN > ldb_s)
        return set_error(BFP_E_ARG, "scale arrays must be padded to the tile size (rows_pad multiple of 256)");
    CUtensorMap map_a, map_b;
    if (int rc = make_map(&map_a, a_mant, T, Kp, BM)) return rc;
    if (int rc = make_map(&map_b, b_mant, N, Kp, BN)) return rc;
    const cudaError_t attr_err = cudaFuncSetAttribute(bfp_gemm_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal);
    if (attr_err != cudaSuccess) return set_errorf(BFP_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
    const int grid = std::min(p.tiles_m * p.tiles_n, device_info().sm_count);
    bfp_gemm_i8_kernel<<<grid, kThreadsI8, kSmemTotal, st>>>(map_a, map_b, p);
    count_launch();
    return check_launch("bfp_gemm_i8_kernel");
}

template <int TBN, int CG>
static int launch_bf16(const CUtensorMap& map_a, const CUtensorMap& map_b, const CUtensorMap& map_out, const gemm::ParamsBf16& p, int units,
                       cudaStream_t st) {
    using namespace gemm;
    using C = Bf16Cfg<TBN, CG>;
    cudaError_t e = cudaFuncSetAttribute(bfp_gemm_bf16_kernel<TBN, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemTotal);
    if (e != cudaSuccess) return set_errorf(BFP_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(units * CG)); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = C::kSmemTotal; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, bfp_gemm_bf16_kernel<TBN, CG>, map_a, map_b, map_out, p);
    if (e != cudaSuccess) return set_errorf(BFP_E_CUDA, "cudaLaunchKernelEx(bfp_gemm_bf16_kernel): %s", cudaGetErrorString(e));
    return BFP_OK;
}

int gemm_bf16_device(const void* a_bf16, const void* b_bf16, const float* bias, float* out, int64_t T, int64_t N, int64_t Kp,
                     cudaStream_t st) {
    return gemm_bf16_ex_device(a_bf16, b_bf16, bias, out, BFP_DT_F32, T, N, Kp, st);
}

int gemm_bf16_ex_device(const void* a_bf16, const void* b_bf16, const float* bias, void* out_v, int out_dtype, int64_t T, int64_t N, int64_t Kp,
                        cudaStream_t st, int accumulate, int64_t batch) {
    using namespace gemm;
    if (T == 0 || N == 0) return BFP_OK;
    if (out_dtype != BFP_DT_F32 && out_dtype != BFP_DT_F16 && out_dtype != BFP_DT_BF16) return set_error(BFP_E_ARG, "bad output dtype");
    float* out = static_cast<float*>(out_v);
    const int out_es = out_dtype == BFP_DT_F32 ? 4 : 2;
    if (Kp % 8 != 0 || Kp <= 0) return set_error(BFP_E_ARG, "bf16 operand K must be a positive multiple of 8");
    if (T > INT32_MAX || N > INT32_MAX || Kp > INT32_MAX) return set_error(BFP_E_ARG, "dimension too large");
    if (reinterpret_cast<uintptr_t>(a_bf16) % 16 || reinterpret_cast<uintptr_t>(b_bf16) % 16)
        return set_error(BFP_E_ALIGN, "bf16 operands must be 16-byte aligned");
    ParamsBf16 p;
    p.bias = bias; p.out = out; p.out_dtype = out_dtype; p.T = (int)T; p.N = (int)N; p.accumulate = accumulate; p.batch = (int)batch;
    if (batch < 1 || batch * T > INT32_MAX || batch * N > INT32_MAX) return set_error(BFP_E_ARG, "bad batch count");
    p.num_k_stages = (int)((Kp + 63) / 64);
    // tile: CTA pairs on 256x256 (cta_group::2) unless the problem is a single 128-row or 128-column strip; the knobs
    // gemm_bf16_cta_group (1 / 2) and gemm_bf16_tile_n (128 / 256, single-CTA mode only) force a variant.
    // (Measured on B200, profiles/r01_gemm_bench_v3_tiles.log: the 128x128 tile is operand-feed-bound at ~1.05 PFLOP/s
    // and loses to 128x256 (1.24-1.45) even where its wave efficiency is better.)
    const int sms = std::max(2, device_info().sm_count);
    int cg = (T > 128 && N > 128) ? 2 : 1;
    if (tuning().gemm_bf16_cta_group == 1 || tuning().gemm_bf16_cta_group == 2) cg = tuning().gemm_bf16_cta_group;
    int tbn = N <= 128 ? 128 : 256;
    if (tuning().gemm_bf16_tile_n == 128 || tuning().gemm_bf16_tile_n == 256) { tbn = tuning().gemm_bf16_tile_n; if (tbn == 128) cg = 1; }
    if (cg == 2) tbn = 256;
    p.tiles_m = (int)((T + BM * cg - 1) / (BM * cg));
    p.tiles_n = (int)((N + tbn - 1) / tbn);
    CUtensorMap map_a, map_b;
    if (int rc = make_map(&map_a, a_bf16, batch * T, Kp * 2, BM, true)) return rc;          // batched operands are stacked along rows
    if (int rc = make_map(&map_b, b_bf16, batch * N, Kp * 2, tbn / cg, true)) return rc;
    CUtensorMap map_out = map_a;
    p.out_tma = ((N * out_es) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0 && (tuning().gemm_out_tma || batch > 1)) ? 1 : 0;
    if (batch > 1) {
        if (!p.out_tma || bias || accumulate)
            return set_error(BFP_E_UNSUPPORTED, "batched GEMM: 16-byte aligned output rows (N * element size % 16 == 0), no bias, no accumulation");
        if (int rc = make_map_out3(&map_out, out, out_dtype, batch, T, N, out_es == 4 ? 32 : 64, 32, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    } else if (p.out_tma) {
        if (int rc = make_map_out(&map_out, out, out_dtype, T, N, N * out_es, out_es == 4 ? 32 : 64, 32, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    }
    if (accumulate && (!p.out_tma || out_dtype != BFP_DT_F32 || bias))
        return set_error(BFP_E_UNSUPPORTED, "accumulating GEMM: fp32 output with 16-byte aligned rows (N % 4 == 0), no bias");
    const int units = (int)std::min<int64_t>((int64_t)p.tiles_m * p.tiles_n * batch, sms / cg);
    int rc;
    if (cg == 2) rc = launch_bf16<256, 2>(map_a, map_b, map_out, p, units, st);
    else if (tbn == 256) rc = launch_bf16<256, 1>(map_a, map_b, map_out, p, units, st);
    else rc = launch_bf16<128, 1>(map_a, map_b, map_out, p, units, st);
    if (rc) return rc;
    count_launch();
    return check_launch("bfp_gemm_bf16_kernel");
}

}  // namespace bfp
