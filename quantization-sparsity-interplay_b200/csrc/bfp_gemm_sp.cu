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
// Kernel: persistent, 384 threads per CTA.  warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warps
// 4-11 = epilogue.  One smem slab = 128 logical k = one E atom = four MMAs: W 128 rows x 128 B (64 kept bf16), X two
// 64-k swizzle atoms of XR rows x 128 B, E 2 KB as 16 TMA rows of 128 B -- every TMA row is a full 128 bytes because the
// copy engine is request-rate-bound (measured ~2 clk per row: a 64-k slab with 64-byte W rows ran at 260 clk per MMA,
// tools/exp_sp_mma_rate.py).  The E atom goes to TMEM by tcgen05.cp.128x128b into an 8-atom ring of columns (odd columns
// are addressed through the sparsity selector); the single 256-column accumulator is drained into registers in one
// tcgen05.ld burst and handed back before the global stores, so the next tile's MMAs wait only for the burst.
// CG = 2 (default): a cluster pair issues cta_group::2 MMAs on a 256 x 256 tile, each CTA staging its own 128 W rows and
// half of X; both CTAs' TMA bytes complete on the leader's mbarrier, tcgen05.commit multicasts the releases.
#include <cuda.h>

#include <algorithm>

#include "bfp_internal.h"
#include "bfp_tc.cuh"

namespace bfp {
namespace gemm_sp {

using namespace gemm;

// CG = 1: one CTA per tile (128 out-features x 256 tokens).  CG = 2: a cluster pair shares a 256 x 256 tile through
// tcgen05.mma.cta_group::2 -- each CTA stages its own 128 W rows and only HALF of the X tile, so the shared-memory feed per
// MMA drops from 20 KB to 12 KB per SM (the 1-CTA kernel is feed-bound: 160 B/clk wanted, 128 B/clk available).
constexpr int kSmemStaging = 8 * 2 * 1024;      // epilogue: per warp two [8 tokens][32 out-features] fp32 tiles for the TMA stores
template <int CG> struct Cfg {
    static constexpr int BW = 128;                       // out-features (rows of W) per CTA = TMEM lanes
    static constexpr int BT = 256;                       // tokens per tile = accumulator columns
    static constexpr int XR = BT / CG;                   // X rows staged by one CTA
    static constexpr int KS = 128;                       // logical k per smem slab = one E atom = four MMAs
    static constexpr int kSmemW = BW * (KS / 2) * 2;     // 16 KB: 64 kept bf16 = one 128-byte swizzle row per W row
    static constexpr int kSmemXAtom = XR * 128;          // one 64-k swizzle atom of X: 32 KB / 16 KB
    static constexpr int kSmemX = 2 * kSmemXAtom;        // two atoms per slab
    static constexpr int kSmemE = 2048;                  // one E atom
    static constexpr int kStageBytes = kSmemW + kSmemX + kSmemE;      // 83968 / 51200, multiples of 1024
    static constexpr int kStages = CG == 1 ? 2 : 4;
    static constexpr int kSmemTotal = kStages * kStageBytes + kSmemStaging + 1024 + 1024;
    // D = F32, A = B = BF16, K-major, N = BT, M = BW * CG, sparse flag (bit 2); bits 0-1 = sparsity selector
    static constexpr uint32_t kIdesc = (1u << 2) | (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BT >> 3) << 17) | ((uint32_t)((BW * CG) >> 4) << 24);
};
constexpr int kMaxStages = 4;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 128 + kEpiWarps * 32;
constexpr int kTmemCols = 512;
constexpr int kTmemE = 256;                  // first E column
constexpr int kERing = 8;                    // E atoms resident in TMEM: kERing > kMaxStages (reuse argument in the MMA warp)

struct Params {
    const float* bias;          // [N] or nullptr
    float* out;                 // [T][N]
    int T, N;
    int num_k_slabs;            // ceil(K / 128)
    int e_atoms;                // ceil(K / 128)
    int tiles_w, tiles_t;       // tiles_w counts CG * 128 rows
    int64_t ld_out;             // row stride of `out` in floats (plain-store path)
    int out_dtype;              // BFP_DT_F32, or BFP_DT_F16 / BFP_DT_BF16: the fp32 accumulator (+ bias) is rounded to it once
    int out_tma;                // 1 = the epilogue stages the tile in smem and writes it with TMA stores (needs N % 4 == 0)
    int accumulate;             // 1 = out += product (fp32, TMA path, one destination): TMA reduce-add instead of a store
    int debug;                  // timing experiments only (wrong results): bit 0 = every tile loads X tile 0, bit 1 = W tile 0
};

// Output destinations of the epilogue's TMA stores.  n = 1: the caller's [T, N] tensor.  n = G > 1: the fused all-gather of
// the column-parallel linear -- destination g is this rank's column slice inside rank g's full [T, N_total] output (own
// HBM or a peer's over NVLink, symmetric memory), so the tile is stored G times straight from the staging smem tile and no
// separate collective (nor the transposing copy after it) is needed.
constexpr int kMaxDests = 8;
struct OutMaps {
    CUtensorMap m[kMaxDests];
    int n;
};

struct Barriers {
    uint64_t full[kMaxStages];
    uint64_t empty[kMaxStages];
    uint64_t tmem_full;
    uint64_t tmem_empty;
    uint32_t tmem_base;
};

template <int CG>
__device__ __forceinline__ void mma_sp_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t tmem_e, uint32_t idesc,
                                            uint32_t accumulate) {
    if constexpr (CG == 1)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %5, 0;\n\t"
            "tcgen05.mma.sp.cta_group::1.kind::f16 [%0], %1, %2, [%3], %4, p;\n\t}"
            ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(tmem_e), "r"(idesc), "r"(accumulate) : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %5, 0;\n\t"
            "tcgen05.mma.sp.cta_group::2.kind::f16 [%0], %1, %2, [%3], %4, p;\n\t}"
            ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(tmem_e), "r"(idesc), "r"(accumulate) : "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_cp_128x128b(uint32_t tmem_dst, uint64_t smem_desc) {
    if constexpr (CG == 1) asm volatile("tcgen05.cp.cta_group::1.128x128b [%0], %1;" ::"r"(tmem_dst), "l"(smem_desc) : "memory");
    else asm volatile("tcgen05.cp.cta_group::2.128x128b [%0], %1;" ::"r"(tmem_dst), "l"(smem_desc) : "memory");
}
// Epilogue store of NC accumulator columns (tokens t0 .. t0+NC-1) of this warp's 32 out-features n0 .. n0+31.
// TMA path: the warp transposes 8 tokens at a time through a private double-buffered smem tile [8 t][32 n] and one lane
// issues a TMA store per tile (8 full 128-byte rows; rows / columns beyond T / N are clipped by the copy engine).  Plain
// path (N % 4 != 0): one 128-byte warp store per token.  Measured on B200: the st.global epilogue cost ~28 % of the whole
// kernel at K = 4096 (it slows the operand stream while it runs, tools/exp_sp_tile_overhead.py).
template <int NC, bool HALF>
__device__ __forceinline__ void store_columns(const Params& p, const OutMaps& outs, float* stg, const uint32_t* r, float bv,
                                              int n0, int lane, int t0) {
    if (p.debug & 8) return;
    if (p.out_tma && HALF) {
        // Half-precision output: 32 out-features are only 64 bytes per token, so the two warps of adjacent lane quarters
        // (same tokens, out-features n0 .. n0+63 together) share one [8 t][64 n] tile of full 128-byte rows -- the unit both
        // the L2 and, for the fused all-gather, NVLink move efficiently -- synchronised by a 64-thread named barrier.
        const uint64_t pol = l2_policy_evict_first();
        const int ew = ((int)threadIdx.x >> 5) - 4, odd = ew & 1, bar_id = 1 + (ew >> 1);
        uint16_t* pstg = reinterpret_cast<uint16_t*>(stg - odd * 512);            // the even warp's 2 KB: two 1 KB tiles
        const bool leader = !odd && lane == 0;
#pragma unroll
        for (int rd = 0; rd < NC / 8; ++rd) {
            uint16_t* hb = pstg + (rd & 1) * 512;
            // round 0 starts on tile 0 again whatever the previous call ended on (NC / 8 may be odd): wait for every earlier read
            if (leader) { if (rd == 0) tma_store_wait_read<0>(); else tma_store_wait_read<1>(); }
            asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");           // tile free (the leader's stores of two rounds ago have read it)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float o = __uint_as_float(r[rd * 8 + j]) + bv;
                hb[j * 64 + odd * 32 + lane] = p.out_dtype == BFP_DT_F16 ? __half_as_ushort(__float2half_rn(o)) : __bfloat16_as_ushort(__float2bfloat16_rn(o));
            }
            fence_proxy_async_smem();
            asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");           // both warps have written
            if (leader) {
                tma_store_2d_hint(&outs.m[0], hb, n0, t0 + rd * 8, pol);          // leader's n0 is the pair's first out-feature
                if (outs.n > 1) {
                    for (int g = 1; g < outs.n; ++g) tma_store_2d_hint(&outs.m[g], hb, n0, t0 + rd * 8, pol);
                }
                tma_store_commit();
            }
        }
    } else if (p.out_tma) {
        const uint64_t pol = l2_policy_evict_first();
#pragma unroll
        for (int rd = 0; rd < NC / 8; ++rd) {
            float* buf = stg + (rd & 1) * 256;
            // the stores that last read this buffer (two rounds ago; one bulk group per round) are done with it; round 0 starts
            // on tile 0 again whatever the previous call ended on (NC / 8 may be odd), so it waits for every earlier read
            if (lane == 0) { if (rd == 0) tma_store_wait_read<0>(); else tma_store_wait_read<1>(); }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j) buf[j * 32 + lane] = __uint_as_float(r[rd * 8 + j]) + bv;
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                tma_store_or_add_2d(&outs.m[0], buf, n0, t0 + rd * 8, pol, p.accumulate);
                if (outs.n > 1) {                                 // fused all-gather: the same tile to every peer's buffer
                    for (int g = 1; g < outs.n; ++g) tma_store_2d_hint(&outs.m[g], buf, n0, t0 + rd * 8, pol);
                }
                tma_store_commit();
            }
        }
    } else {
        const int n = n0 + lane;
        if (n < p.N) {
            const int t_left = p.T - t0;
            if (!HALF) {
                float* dst = p.out + (int64_t)t0 * p.ld_out + n;
#pragma unroll
                for (int j = 0; j < NC; ++j)
                    if (j < t_left) dst[(int64_t)j * p.ld_out] = __uint_as_float(r[j]) + bv;
            } else {
                uint16_t* dst = reinterpret_cast<uint16_t*>(p.out) + (int64_t)t0 * p.ld_out + n;
#pragma unroll
                for (int j = 0; j < NC; ++j) {
                    const float o = __uint_as_float(r[j]) + bv;
                    if (j < t_left) dst[(int64_t)j * p.ld_out] = p.out_dtype == BFP_DT_F16 ? __half_as_ushort(__float2half_rn(o)) : __bfloat16_as_ushort(__float2bfloat16_rn(o));
                }
            }
        }
    }
}

template <int CG, bool HALF>
__global__ void __launch_bounds__(kThreads, 1)
bfp_gemm_bf16_sp_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x,
                        const __grid_constant__ CUtensorMap map_e, const __grid_constant__ OutMaps outs, const Params p) {
    using C = Cfg<CG>;
    constexpr int kStages = C::kStages, kStageBytes = C::kStageBytes, BW = C::BW, BT = C::BT, KS = C::KS;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* staging = reinterpret_cast<float*>(smem + kStages * kStageBytes);
    Barriers* bars = reinterpret_cast<Barriers*>(smem + kStages * kStageBytes + kSmemStaging);
    auto stage_x = [&](int s) { return smem + s * kStageBytes; };
    auto stage_w = [&](int s) { return smem + s * kStageBytes + C::kSmemX; };
    auto stage_e = [&](int s) { return smem + s * kStageBytes + C::kSmemX + C::kSmemW; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = CG == 1 ? 0u : cluster_ctarank();          // position in the CTA pair; 0 issues the MMAs
    const int unit = (int)blockIdx.x / CG, num_units = (int)gridDim.x / CG;   // a unit = one CTA (CG 1) or one pair (CG 2)
    const int num_tiles = p.tiles_w * p.tiles_t;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
        mbar_init(&bars->tmem_full, 1);
        mbar_init(&bars->tmem_empty, kEpiWarps * CG);                 // the leader collects both CTAs' epilogue warps
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
    if constexpr (CG == 1) __syncthreads(); else cluster_sync_all();  // the peer's barriers must exist before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        // ===================================== TMA producer (every CTA) =========================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            int uses = 0;
            const uint64_t pol_keep = l2_policy_evict_last();
            for (int tile = unit; tile < num_tiles; tile += num_units) {
                const int tw = tile % p.tiles_w, tt = tile / p.tiles_w;      // consecutive units share the X tile
                const int w_row = (tw * CG + (int)rank) * BW;                // this CTA's 128 W rows
                const int x_row = tt * BT + (int)rank * C::XR;               // this CTA's share of the X tile
                int e_row = (tw * CG + (int)rank) * p.e_atoms * 16;          // E atoms as 16 rows of 128 B (few, wide TMA rows)
                int w_row_ld = w_row, x_row_ld = x_row;
                if (p.debug & 1) x_row_ld = (int)rank * C::XR;
                if (p.debug & 2) { w_row_ld = (int)rank * BW; e_row = (int)rank * p.e_atoms * 16; }
                for (int ks = 0; ks < p.num_k_slabs; ++ks) {
                    if ((p.debug & 4) && uses++ >= kStages) continue;       // experiment: MMAs re-read the first slabs, no loads
                    mbar_wait(&bars->empty[stage], phase ^ 1);
                    // completion bytes of BOTH CTAs land on the leader's barrier; only the leader posts the expectation
                    if (rank == 0) mbar_expect_tx(&bars->full[stage], (uint32_t)CG * kStageBytes);
                    const uint32_t bar = CG == 1 ? smem_u32(&bars->full[stage]) : mapa_u32(smem_u32(&bars->full[stage]), 0);
                    // every TMA row is a full 128 bytes: the copy engine is request-rate-bound (~2 clk per row), so 64-byte
                    // rows (a 64-k W slab) would halve its throughput
                    tma_load_2d_to_hint<CG>(stage_x(stage), &map_x, bar, ks * KS, x_row_ld, pol_keep);
                    tma_load_2d_to_hint<CG>(stage_x(stage) + C::kSmemXAtom, &map_x, bar, ks * KS + 64, x_row_ld, pol_keep);
                    tma_load_2d_to_hint<CG>(stage_w(stage), &map_w, bar, ks * (KS / 2), w_row_ld, pol_keep);
                    tma_load_2d_to_hint<CG>(stage_e(stage), &map_e, bar, 0, e_row + ks * 16, pol_keep);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer (leader CTA only) =====================
        if (lane == 0 && rank == 0) {
            int stage = 0; uint32_t phase = 0;
            uint32_t tile_phase = 0;
            uint32_t eslot = kERing - 1;
            int uses = 0;
            for (int tile = unit; tile < num_tiles; tile += num_units) {
                mbar_wait(&bars->tmem_empty, tile_phase ^ 1);              // every epilogue warp has drained the accumulator
                tc_fence_after();
                for (int ks = 0; ks < p.num_k_slabs; ++ks) {
                    if (!((p.debug & 4) && uses++ >= kStages)) mbar_wait(&bars->full[stage], phase);
                    tc_fence_after();
                    // E ring reuse: having seen full[] for this slab means the producers saw empty[] of the slab kStages
                    // uses earlier, i.e. every MMA at least kStages slabs back has completed; the atom this slot held was
                    // last read kERing (> kStages) slabs back, so the copy cannot overtake a reader.
                    eslot = (eslot + 1) & (kERing - 1);
                    tmem_cp_128x128b<CG>(tmem_base + kTmemE + eslot * 4, make_smem_desc_k(smem_u32(stage_e(stage)), 0, 128, 128));
                    const uint64_t dw = make_smem_desc(smem_u32(stage_w(stage)));                        // SWIZZLE_128B
                    const uint64_t dx0 = make_smem_desc(smem_u32(stage_x(stage)));
                    const uint64_t dx1 = make_smem_desc(smem_u32(stage_x(stage) + C::kSmemXAtom));
                    const uint32_t ecol = tmem_base + kTmemE + eslot * 4;
#pragma unroll
                    for (int i = 0; i < 4; ++i)      // 32 logical k per MMA: 16 kept bf16 = 32 B of W (+2), 64 B of X (+4, two per atom)
                        // the metadata address names an even column; the sparsity selector (idesc bits 0-1) picks the odd one
                        mma_sp_bf16<CG>(tmem_base, dw + (uint64_t)(i * 2), (i < 2 ? dx0 : dx1) + (uint64_t)((i & 1) * 4), ecol + (uint32_t)(i & 2),
                                        C::kIdesc | (uint32_t)(i & 1), (ks | i) != 0);
                    commit<CG>(&bars->empty[stage]);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                commit<CG>(&bars->tmem_full);
                tile_phase ^= 1;
            }
        }
    } else if (warp >= 4) {
        // ===================================== epilogue (every CTA) =============================
        const int ew = warp - 4, q = warp & 3, half = ew >> 2;
        const int n_in_tile = q * 32 + lane;
        const uint32_t leader_tmem_empty = CG == 1 ? 0u : mapa_u32(smem_u32(&bars->tmem_empty), 0);
        uint32_t tile_phase = 0;
        for (int tile = unit; tile < num_tiles; tile += num_units) {
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
            if (lane == 0) {                                                // accumulator is free again
                if constexpr (CG == 1) mbar_arrive(&bars->tmem_empty); else mbar_arrive_cluster(leader_tmem_empty);
            }
            tile_phase ^= 1;

            const int n0 = (tw * CG + (int)rank) * BW + q * 32;
            const float bv = (p.bias && n0 + lane < p.N) ? p.bias[n0 + lane] : 0.0f;
            store_columns<BT / 2, HALF>(p, outs, staging + ew * 512, r, bv, n0, lane, tt * BT + half * (BT / 2));
        }
        if (p.out_tma && lane == 0) tma_store_wait_all<0>();                 // smem (and the stores) must outlive the CTA's exit
    }

    tc_fence_before();
    if constexpr (CG == 1) __syncthreads(); else cluster_sync_all();       // the peer may still be signalling / reading this CTA
    if (warp == 2) {
        tc_fence_after();
        if constexpr (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}

// ====================================================================================================================
// Wide tile: 256 out-features x 480 tokens per CTA pair, two 240-column accumulators fed by the SAME W / E slab.
// The 256-wide kernel above is bound by the SM's ingest port (50 KB of operands per 128-k slab per SM at 64 B/clk);
// sharing each W slab and E atom between two X halves brings that to 78 KB per 480 tokens = 41.6 KB per 256 (-17 %), and
// 480-token tiles also quantise better on 74 CTA pairs for the LLaMA shapes (4096^3: 1.95 waves instead of 3.46).
//   smem: W/E ring of 3 slots (128 k: 16 KB + 2 KB) + X ring of 5 stages (64 k: 2 chunks x 120 rows x 128 B = 30 KB).
//   TMEM: acc0 = columns [0,240), acc1 = [240,480), E ring = [480,512).
//   MMA:  tcgen05.mma.sp.cta_group::2 with N = 240; per CTA the B operand is 120 rows (chunk c = this CTA's half of
//         accumulator c's tokens: rows t0 + 240 c + 120 rank + [0,120)).
//   Epilogue: warps 4-7 drain acc0, warps 8-11 acc1, 120 columns at a time (the 480 x 128 fp32 tile is the size of the
//         whole register file); the accumulators are handed back after the second tcgen05.ld burst.
// ====================================================================================================================
namespace wide {
constexpr int BW = 128, NT = 240, BT = 2 * NT;           // W rows per CTA, tokens per accumulator, tokens per tile
constexpr int XC = NT / 2;                               // X rows per chunk per CTA (120)
constexpr int kWeSlots = 3, kXStages = 5;
constexpr int kSmemW = BW * 128, kSmemE = 2048, kWeBytes = kSmemW + kSmemE;          // 18 KB
constexpr int kSmemXChunk = XC * 128, kXBytes = 2 * kSmemXChunk;                     // 15 KB, 30 KB
constexpr int kSmemTotal = kWeSlots * kWeBytes + kXStages * kXBytes + kSmemStaging + 1024 + 1024;
constexpr uint32_t kIdesc = (1u << 2) | (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)((BW * 2) >> 4) << 24);
constexpr int kTmemEW = 480;
struct Barriers {
    uint64_t we_full[kWeSlots], we_empty[kWeSlots];
    uint64_t x_full[kXStages], x_empty[kXStages];
    uint64_t tmem_full, tmem_empty;
    uint32_t tmem_base;
};
}  // namespace wide

__device__ __forceinline__ void tmem_ld8(uint32_t addr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(addr));
}

template <bool HALF>
__global__ void __launch_bounds__(kThreads, 1)
bfp_gemm_bf16_sp_wide_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x,
                             const __grid_constant__ CUtensorMap map_e, const __grid_constant__ OutMaps outs, const Params p) {
    using namespace wide;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_x = smem + kWeSlots * kWeBytes;                       // 55296: multiple of 1024
    float* staging = reinterpret_cast<float*>(smem_x + kXStages * kXBytes);
    wide::Barriers* bars = reinterpret_cast<wide::Barriers*>(smem_x + kXStages * kXBytes + kSmemStaging);
    auto slot_w = [&](int s) { return smem + s * kWeBytes; };
    auto slot_e = [&](int s) { return smem + s * kWeBytes + kSmemW; };
    auto stage_x = [&](int s, int c) { return smem_x + s * kXBytes + c * kSmemXChunk; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int unit = (int)blockIdx.x / 2, num_units = (int)gridDim.x / 2;
    const int num_tiles = p.tiles_w * p.tiles_t;
    const int nk128 = p.num_k_slabs;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kWeSlots; ++s) { mbar_init(&bars->we_full[s], 1); mbar_init(&bars->we_empty[s], 1); }
        for (int s = 0; s < kXStages; ++s) { mbar_init(&bars->x_full[s], 1); mbar_init(&bars->x_empty[s], 1); }
        mbar_init(&bars->tmem_full, 1);
        mbar_init(&bars->tmem_empty, kEpiWarps * 2);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        // ===================================== TMA producer (both CTAs) =========================
        if (lane == 0) {
            int ws = 0, xs = 0; uint32_t wphase = 0, xphase = 0;
            const uint64_t pol_keep = l2_policy_evict_last();
            for (int tile = unit; tile < num_tiles; tile += num_units) {
                const int tw = tile % p.tiles_w, tt = tile / p.tiles_w;
                const int w_row = (tw * 2 + (int)rank) * BW;
                const int e_row = (tw * 2 + (int)rank) * p.e_atoms * 16;
                const int x_row0 = tt * BT + (int)rank * XC;                 // chunk 0 (accumulator 0); chunk 1 is NT rows further
                for (int k = 0; k < nk128; ++k) {
                    mbar_wait(&bars->we_empty[ws], wphase ^ 1);
                    if (rank == 0) mbar_expect_tx(&bars->we_full[ws], 2u * kWeBytes);
                    const uint32_t wbar = mapa_u32(smem_u32(&bars->we_full[ws]), 0);
                    tma_load_2d_to_hint<2>(slot_w(ws), &map_w, wbar, k * 64, w_row, pol_keep);
                    tma_load_2d_to_hint<2>(slot_e(ws), &map_e, wbar, 0, e_row + k * 16, pol_keep);
                    if (++ws == kWeSlots) { ws = 0; wphase ^= 1; }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        mbar_wait(&bars->x_empty[xs], xphase ^ 1);
                        if (rank == 0) mbar_expect_tx(&bars->x_full[xs], 2u * kXBytes);
                        const uint32_t xbar = mapa_u32(smem_u32(&bars->x_full[xs]), 0);
                        tma_load_2d_to_hint<2>(stage_x(xs, 0), &map_x, xbar, k * 128 + h * 64, x_row0, pol_keep);
                        tma_load_2d_to_hint<2>(stage_x(xs, 1), &map_x, xbar, k * 128 + h * 64, x_row0 + NT, pol_keep);
                        if (++xs == kXStages) { xs = 0; xphase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer (leader CTA) ==========================
        if (lane == 0 && rank == 0) {
            int ws = 0, xs = 0; uint32_t wphase = 0, xphase = 0, tile_phase = 0, eslot = kERing - 1;
            for (int tile = unit; tile < num_tiles; tile += num_units) {
                mbar_wait(&bars->tmem_empty, tile_phase ^ 1);
                tc_fence_after();
                for (int k = 0; k < nk128; ++k) {
                    mbar_wait(&bars->we_full[ws], wphase);
                    tc_fence_after();
                    // E ring: the slot's previous atom was read kERing (8) slabs ago; we_full of this slab implies the
                    // producers saw we_empty of the slab kWeSlots (3) back, i.e. every MMA at least 3 slabs old is done.
                    eslot = (eslot + 1) & (kERing - 1);
                    const uint32_t ecol = tmem_base + kTmemEW + eslot * 4;
                    tmem_cp_128x128b<2>(ecol, make_smem_desc_k(smem_u32(slot_e(ws)), 0, 128, 128));
                    const uint64_t dw = make_smem_desc(smem_u32(slot_w(ws)));
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        mbar_wait(&bars->x_full[xs], xphase);
                        tc_fence_after();
                        const uint64_t dx0 = make_smem_desc(smem_u32(stage_x(xs, 0)));
                        const uint64_t dx1 = make_smem_desc(smem_u32(stage_x(xs, 1)));
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const int j = h * 2 + i;                         // k32 step inside the 128-k slab
                            const uint32_t acc = (uint32_t)((k | j) != 0);
                            mma_sp_bf16<2>(tmem_base, dw + (uint64_t)(j * 2), dx0 + (uint64_t)(i * 4), ecol + (uint32_t)(j & 2), kIdesc | (uint32_t)(j & 1), acc);
                            mma_sp_bf16<2>(tmem_base + NT, dw + (uint64_t)(j * 2), dx1 + (uint64_t)(i * 4), ecol + (uint32_t)(j & 2), kIdesc | (uint32_t)(j & 1), acc);
                        }
                        commit<2>(&bars->x_empty[xs]);
                        if (++xs == kXStages) { xs = 0; xphase ^= 1; }
                    }
                    commit<2>(&bars->we_empty[ws]);
                    if (++ws == kWeSlots) { ws = 0; wphase ^= 1; }
                }
                commit<2>(&bars->tmem_full);
                tile_phase ^= 1;
            }
        }
    } else if (warp >= 4) {
        // ===================================== epilogue (both CTAs) =============================
        const int q = warp & 3, a = (warp - 4) >> 2;                         // TMEM lane quarter, accumulator
        const int n_in_tile = q * 32 + lane;
        const uint32_t leader_tmem_empty = mapa_u32(smem_u32(&bars->tmem_empty), 0);
        uint32_t tile_phase = 0;
        for (int tile = unit; tile < num_tiles; tile += num_units) {
            const int tw = tile % p.tiles_w, tt = tile / p.tiles_w;
            const int n = (tw * 2 + (int)rank) * BW + n_in_tile;
            const float bv = (p.bias && n < p.N) ? p.bias[n] : 0.0f;
            mbar_wait(&bars->tmem_full, tile_phase);
            tc_fence_after();
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * NT + pass * (NT / 2));
                uint32_t r[NT / 2];                                          // 120 columns
#pragma unroll
                for (int c = 0; c < 7; ++c) tmem_ld16(taddr + c * 16, r + c * 16);
                tmem_ld8(taddr + 112, r + 112);
                tmem_ld_wait();
                if (pass == 1) {                                             // everything this warp owns has left TMEM
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(leader_tmem_empty);
                }
                store_columns<NT / 2, HALF>(p, outs, staging + (warp - 4) * 512, r, bv, (tw * 2 + (int)rank) * BW + q * 32, lane,
                                      tt * BT + a * NT + pass * (NT / 2));
            }
            tile_phase ^= 1;
        }
        if (p.out_tma && lane == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}

// ====================================================================================================================
// Ping-pong tile: 256 out-features x 240 tokens per CTA pair, TWO 240-column accumulators used alternately by consecutive
// tiles, so the epilogue of tile i (tcgen05.ld, bias, transposed TMA stores) runs entirely under the main loop of tile i+1 and
// the MMA issuer never waits for a drain.  The single-accumulator 256-token kernel pays ~3 us per tile for that hand-over
// (tools/exp_sp_tile_overhead.py), which is 15-20 % of a K = 4096 tile.  Same slab structure as the 256-token kernel with
// 120 X rows per CTA: W 16 KB + X 2 x 15 KB + E 2 KB = 48 KB per 128 k, four stages.
//   TMEM: acc0 = columns [0,240), acc1 = [240,480), E ring = [480,512).
// ====================================================================================================================
namespace pp {
constexpr int BW = 128, NT = 240, XC = NT / 2;
constexpr int kStages = 4;
constexpr int kSmemW = BW * 128, kSmemE = 2048, kSmemXAtom = XC * 128;              // 16 KB, 2 KB, 15 KB
constexpr int kStageBytes = kSmemW + 2 * kSmemXAtom + kSmemE;                       // 49152
constexpr int kSmemTotal = kStages * kStageBytes + kSmemStaging + 1024 + 1024;
constexpr uint32_t kIdesc = (1u << 2) | (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)((BW * 2) >> 4) << 24);
constexpr int kTmemEP = 480;
struct Barriers {
    uint64_t full[kStages], empty[kStages];
    uint64_t tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
};
}  // namespace pp

template <bool HALF>
__global__ void __launch_bounds__(kThreads, 1)
bfp_gemm_bf16_sp_pp_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x,
                           const __grid_constant__ CUtensorMap map_e, const __grid_constant__ OutMaps outs, const Params p) {
    using namespace pp;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* staging = reinterpret_cast<float*>(smem + kStages * kStageBytes);
    pp::Barriers* bars = reinterpret_cast<pp::Barriers*>(smem + kStages * kStageBytes + kSmemStaging);
    auto stage_w = [&](int s) { return smem + s * kStageBytes; };
    auto stage_x = [&](int s, int a) { return smem + s * kStageBytes + kSmemW + a * kSmemXAtom; };
    auto stage_e = [&](int s) { return smem + s * kStageBytes + kSmemW + 2 * kSmemXAtom; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int unit = (int)blockIdx.x / 2, num_units = (int)gridDim.x / 2;
    const int num_tiles = p.tiles_w * p.tiles_t;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&bars->tmem_full[b], 1); mbar_init(&bars->tmem_empty[b], kEpiWarps * 2); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            const uint64_t pol_keep = l2_policy_evict_last();
            for (int tile = unit; tile < num_tiles; tile += num_units) {
                const int tw = tile % p.tiles_w, tt = tile / p.tiles_w;
                const int w_row = (tw * 2 + (int)rank) * BW, x_row = tt * NT + (int)rank * XC;
                const int e_row = (tw * 2 + (int)rank) * p.e_atoms * 16;
                for (int ks = 0; ks < p.num_k_slabs; ++ks) {
                    mbar_wait(&bars->empty[stage], phase ^ 1);
                    if (rank == 0) mbar_expect_tx(&bars->full[stage], 2u * kStageBytes);
                    const uint32_t bar = mapa_u32(smem_u32(&bars->full[stage]), 0);
                    tma_load_2d_to_hint<2>(stage_x(stage, 0), &map_x, bar, ks * 128, x_row, pol_keep);
                    tma_load_2d_to_hint<2>(stage_x(stage, 1), &map_x, bar, ks * 128 + 64, x_row, pol_keep);
                    tma_load_2d_to_hint<2>(stage_w(stage), &map_w, bar, ks * 64, w_row, pol_keep);
                    tma_load_2d_to_hint<2>(stage_e(stage), &map_e, bar, 0, e_row + ks * 16, pol_keep);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            int stage = 0; uint32_t phase = 0, eslot = kERing - 1;
            int buf = 0; uint32_t buf_phase[2] = {0, 0};
            for (int tile = unit; tile < num_tiles; tile += num_units) {
                mbar_wait(&bars->tmem_empty[buf], buf_phase[buf] ^ 1);      // drained two tiles ago: normally already free
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)buf * NT;
                for (int ks = 0; ks < p.num_k_slabs; ++ks) {
                    mbar_wait(&bars->full[stage], phase);
                    tc_fence_after();
                    eslot = (eslot + 1) & (kERing - 1);               // reuse distance kERing (8) slabs > kStages (4): see the 256-token kernel
                    const uint32_t ecol = tmem_base + kTmemEP + eslot * 4;
                    tmem_cp_128x128b<2>(ecol, make_smem_desc_k(smem_u32(stage_e(stage)), 0, 128, 128));
                    const uint64_t dw = make_smem_desc(smem_u32(stage_w(stage)));
                    const uint64_t dx0 = make_smem_desc(smem_u32(stage_x(stage, 0))), dx1 = make_smem_desc(smem_u32(stage_x(stage, 1)));
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        mma_sp_bf16<2>(d, dw + (uint64_t)(i * 2), (i < 2 ? dx0 : dx1) + (uint64_t)((i & 1) * 4), ecol + (uint32_t)(i & 2),
                                       kIdesc | (uint32_t)(i & 1), (ks | i) != 0);
                    commit<2>(&bars->empty[stage]);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                commit<2>(&bars->tmem_full[buf]);
                buf_phase[buf] ^= 1;
                buf ^= 1;
            }
        }
    } else if (warp >= 4) {
        const int ew = warp - 4, q = warp & 3, half = ew >> 2;
        int buf = 0; uint32_t buf_phase[2] = {0, 0};
        for (int tile = unit; tile < num_tiles; tile += num_units) {
            const int tw = tile % p.tiles_w, tt = tile / p.tiles_w;
            mbar_wait(&bars->tmem_full[buf], buf_phase[buf]);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * NT + half * (NT / 2));
            uint32_t r[NT / 2];                                              // 120 columns
#pragma unroll
            for (int c = 0; c < 7; ++c) tmem_ld16(taddr + c * 16, r + c * 16);
            tmem_ld8(taddr + 112, r + 112);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&bars->tmem_empty[buf]), 0));
            buf_phase[buf] ^= 1;
            buf ^= 1;
            const int n0 = (tw * 2 + (int)rank) * BW + q * 32;
            const float bv = (p.bias && n0 + lane < p.N) ? p.bias[n0 + lane] : 0.0f;
            store_columns<NT / 2, HALF>(p, outs, staging + ew * 512, r, bv, n0, lane, tt * NT + half * (NT / 2));
        }
        if (p.out_tma && lane == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
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
    void* outs[1] = {out};
    return gemm_bf16_sp_multi_device(x_bf16, w_comp, w_meta, bias, outs, 1, BFP_DT_F32, N, T, N, Kp, st);
}

int gemm_bf16_sp_multi_device(const void* x_bf16, const void* w_comp, const void* w_meta, const float* bias, void* const* out_ptrs, int n_out,
                              int out_dtype, int64_t ld_out, int64_t T, int64_t N, int64_t Kp, cudaStream_t st, int accumulate) {
    using namespace gemm_sp;
    if (T == 0 || N == 0) return BFP_OK;
    if (n_out < 1 || n_out > kMaxDests) return set_error(BFP_E_ARG, "1 to 8 output destinations");
    if (ld_out < N) return set_error(BFP_E_ARG, "output row stride smaller than N");
    if (out_dtype != BFP_DT_F32 && out_dtype != BFP_DT_F16 && out_dtype != BFP_DT_BF16) return set_error(BFP_E_ARG, "bad output dtype");
    float* out = static_cast<float*>(out_ptrs[0]);
    const int out_es = out_dtype == BFP_DT_F32 ? 4 : 2;
    if (Kp % 8 != 0 || Kp <= 0) return set_error(BFP_E_ARG, "bf16 operand K must be a positive multiple of 8");
    if (T > INT32_MAX || N > INT32_MAX || Kp > INT32_MAX) return set_error(BFP_E_ARG, "dimension too large");
    if (reinterpret_cast<uintptr_t>(x_bf16) % 16 || reinterpret_cast<uintptr_t>(w_comp) % 16 || reinterpret_cast<uintptr_t>(w_meta) % 16)
        return set_error(BFP_E_ALIGN, "operands must be 16-byte aligned");
    int64_t Kc, mb;
    sp_layout(N, Kp, &Kc, &mb);
    const int sms = std::max(2, device_info().sm_count);
    // CTA pairs (cta_group::2) unless the problem has a single 128-row W tile or the knob forces one CTA per tile
    int cg = (N > 128) ? 2 : 1;
    if (tuning().gemm_sp_cta_group == 1 || tuning().gemm_sp_cta_group == 2) cg = tuning().gemm_sp_cta_group;
    Params p;
    p.bias = bias; p.out = out; p.T = (int)T; p.N = (int)N; p.ld_out = ld_out; p.out_dtype = out_dtype; p.debug = tuning().gemm_sp_debug; p.accumulate = accumulate;
    p.num_k_slabs = (int)((Kp + 127) / 128);
    p.e_atoms = (int)(Kc / 64);
    p.tiles_w = (int)((N + 128 * cg - 1) / (128 * cg));
    p.tiles_t = (int)((T + 255) / 256);
    // Three pair tiles, chosen from measurements over the nine LLaMA shapes (profiles/r01_gemm_sp_bench.log):
    //   480 tokens, two accumulators sharing each W slab (78 KB instead of 2 x 50 KB of operands per SM per 480 tokens): a wave
    //       costs 1.5-1.95x a 256-token wave plus a larger hand-over, so it is taken for K >= 6144 when it needs fewer than
    //       1/1.7 of the waves;
    //   240 tokens, two accumulators ping-ponged between consecutive tiles (no hand-over stall, fitted cost 1.03 S per wave
    //       against S + 2.9 for the single-accumulator tile, S = 128-k slabs);
    //   256 tokens, one accumulator, otherwise.
    bool wide_tile = false, pp_tile = false;
    if (cg == 2) {
        const int64_t pairs = sms / 2, S = p.num_k_slabs;
        const int64_t tiles256 = (int64_t)p.tiles_w * p.tiles_t, tiles480 = (int64_t)p.tiles_w * ((T + wide::BT - 1) / wide::BT);
        const int64_t tiles240 = (int64_t)p.tiles_w * ((T + pp::NT - 1) / pp::NT);
        const int64_t waves256 = (tiles256 + pairs - 1) / pairs, waves480 = (tiles480 + pairs - 1) / pairs, waves240 = (tiles240 + pairs - 1) / pairs;
        wide_tile = waves480 * 17 < waves256 * 10 && S >= 48;
        pp_tile = !wide_tile && waves240 * 1026 * S < waves256 * (1000 * S + 2900);
        if (tuning().gemm_sp_tile == 256) { wide_tile = false; pp_tile = false; }
        if (tuning().gemm_sp_tile == 480) { wide_tile = true; pp_tile = false; }
        if (tuning().gemm_sp_tile == 240) { wide_tile = false; pp_tile = true; }
        if (p.debug & 7) { wide_tile = false; pp_tile = false; }
    }
    CUtensorMap map_w, map_x, map_e;
    if (int rc = make_map_bf16(&map_w, w_comp, N, Kc, Kc * 2, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    if (int rc = make_map_bf16(&map_x, x_bf16, T, Kp, Kp * 2, 64, (wide_tile || pp_tile) ? wide::XC : 256 / cg, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    if (int rc = make_map_bytes(&map_e, w_meta, mb / 128, 128, 16)) return rc;
    OutMaps map_out;
    map_out.n = n_out;
    for (int g = 0; g < kMaxDests; ++g) map_out.m[g] = map_e;
    bool aligned = (ld_out * out_es) % 16 == 0;
    for (int g = 0; g < n_out; ++g) aligned = aligned && out_ptrs[g] && reinterpret_cast<uintptr_t>(out_ptrs[g]) % 16 == 0;
    p.out_tma = (aligned && (tuning().gemm_out_tma || n_out > 1)) ? 1 : 0;
    if (n_out > 1 && !p.out_tma) return set_error(BFP_E_ALIGN, "multi-destination output needs 16-byte aligned slices and a row stride that is a multiple of 16 bytes");
    if (accumulate && (!p.out_tma || out_dtype != BFP_DT_F32 || bias || n_out != 1))
        return set_error(BFP_E_UNSUPPORTED, "accumulating GEMM: one fp32 output with 16-byte aligned rows (N % 4 == 0), no bias");
    // each map covers exactly the [T, N] slice (row stride ld_out), so the copy engine clips at the slice's edge
    for (int g = 0; g < (p.out_tma ? n_out : 0); ++g)
        if (int rc = make_map_out(&map_out.m[g], out_ptrs[g], out_dtype, T, N, ld_out * out_es, out_es == 4 ? 32 : 64, 8)) return rc;
    if (wide_tile) p.tiles_t = (int)((T + wide::BT - 1) / wide::BT);
    if (pp_tile) p.tiles_t = (int)((T + pp::NT - 1) / pp::NT);
    const int units = std::min(p.tiles_w * p.tiles_t, sms / cg);
    const bool half = out_dtype != BFP_DT_F32;
    const void* kern;
    int smem_bytes;
    if (cg == 1) { kern = half ? (const void*)bfp_gemm_bf16_sp_kernel<1, true> : (const void*)bfp_gemm_bf16_sp_kernel<1, false>; smem_bytes = Cfg<1>::kSmemTotal; }
    else if (wide_tile) { kern = half ? (const void*)bfp_gemm_bf16_sp_wide_kernel<true> : (const void*)bfp_gemm_bf16_sp_wide_kernel<false>; smem_bytes = wide::kSmemTotal; }
    else if (pp_tile) { kern = half ? (const void*)bfp_gemm_bf16_sp_pp_kernel<true> : (const void*)bfp_gemm_bf16_sp_pp_kernel<false>; smem_bytes = pp::kSmemTotal; }
    else { kern = half ? (const void*)bfp_gemm_bf16_sp_kernel<2, true> : (const void*)bfp_gemm_bf16_sp_kernel<2, false>; smem_bytes = Cfg<2>::kSmemTotal; }
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return set_errorf(BFP_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)units * cg); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cg; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        void* args[5] = {(void*)&map_w, (void*)&map_x, (void*)&map_e, (void*)&map_out, (void*)&p};
        e = cudaLaunchKernelExC(&cfg, kern, args);
        if (e != cudaSuccess) return set_errorf(BFP_E_CUDA, "cudaLaunchKernelExC(bfp_gemm_bf16_sp): %s", cudaGetErrorString(e));
    }
    count_launch();
    return check_launch("bfp_gemm_bf16_sp_kernel");
}

}  // namespace bfp
