// bfp_gemm_mx.cu -- the BFP linear for narrow mantissas (HBFP4 / HBFP5: mant_bits <= 4) on the block-scaled FP8-class tensor-core
// path: tcgen05.mma.kind::mxf8f6f4.block_scale (sm_100a), twice the rate of kind::f16.
//
// Replaces the reference's dequantise-then-fp32-GEMM (bfp_ops.py:187-190).  A BFP value is q * 2^(e-m) with |q| <= 2^m - 1.  For
// m <= 4 the integer q is exactly representable in E4M3 (integers up to 16), and the block scale 2^(e-m) is exactly a UE8M0 byte
// (the biased exponent).  The hardware multiplies every 32-element group of K by the product of its two scales before the fp32
// accumulation in TMEM, so for block sizes that are multiples of 32 the instruction computes exactly
//     y[t,n] = sum_k (qa[t,k] 2^pa[t,k/B]) (qb[n,k] 2^pb[n,k/B])
// with exact products and fp32 accumulation -- the same function as the exact-bf16 kind (bfp_gemm.cu), with half the operand bytes
// and no CUDA-core rescale (which is what limits the int8 kind to 13 % of peak).
//
// Operand form ("mx", produced by bfp_mx_from_packed from the int8-mantissa pack of bfp_quantize_pack):
//   vals  uint8 [rows, Kp]  E4M3 byte of q, Kp = K rounded up to 128 (zero filled)
//   sf    uint8 [Kp / 128][row_tiles][atoms][512]  UE8M0 scale of (row, 32-group of K): one 512-byte ATOM per 128 rows and 128 k in
//         the byte order tcgen05.cp.32x128b.warpx4 moves into four TMEM columns: byte 16 (r % 32) + 4 (r / 32) + g for row r of
//         the atom and 32-group g of the slab (layout restated from the public CUTLASS headers: Sm1xxBlockScaledBasicChunk /
//         UMMA::tmem_sf_frg).  Rows are grouped in tiles of `tile_rows` (128 for the A operand, the N tile for B), each tile
//         padded to whole atoms, so a tile's scales for one K slab are one contiguous bulk copy.
//
// Kernel: the warp-specialised persistent structure of bfp_gemm_bf16_kernel (TMA producer warp, one MMA-issuing thread, TMEM
// allocator warp, 8 epilogue warps, smem ring of 128-byte K slabs = 4 MMAs of K = 32).  Per slab the MMA thread first moves the
// slab's scale atoms smem -> TMEM (tcgen05.cp, in order with the MMAs on the tensor pipe), then issues four MMAs whose scale
// factor ids select the 32-group.  TMEM: accumulator(s) + 4 columns of A scales + 4 columns per 128 B rows.
#include <cuda.h>
#include <cstdio>

#include <algorithm>

#include "bfp_internal.h"
#include "bfp_stream.cuh"
#include "bfp_tc.cuh"

namespace bfp {

namespace gemm_mx {
using namespace gemm;

constexpr int BM = 128, BKB = 128;                     // rows of A per CTA, K bytes (= elements) per slab
constexpr int kEpiWarps = 8;
constexpr int kThreads = 128 + kEpiWarps * 32;         // 384
constexpr int kTmemCols = 512;
constexpr int kSmemBarriers = 1024;

template <int TBN, int CG> struct Cfg {
    static constexpr int kRowsB = TBN / CG;                                  // B rows staged by one CTA
    static constexpr int kAtomsB = (TBN + 127) / 128;                        // scale atoms of the whole B tile (every CTA of a pair needs all)
    static constexpr int kSfBytes = 512 + kAtomsB * 512;
    static constexpr int kStageBytes = BM * BKB + kRowsB * BKB + 2048;       // operands + scale atoms (padded to keep 1024-byte alignment)
    static constexpr int kStages = (232448 - 8 * 2048 - 2048) / kStageBytes > 8 ? 8 : (232448 - 8 * 2048 - 2048) / kStageBytes;   // 227 KB of shared memory per CTA
    static constexpr int kBufs = 2 * TBN + 4 + 4 * kAtomsB <= kTmemCols ? 2 : 1;   // accumulator buffers that fit beside the scale columns
    static constexpr int kSfCol = kBufs * TBN;                               // first scale column
    static constexpr int kSmemStaging = 8 * 2048;                             // per epilogue warp one [32 rows][16 cols] fp32 tile
    static constexpr int kSmemTotal = kStages * kStageBytes + kSmemStaging + kSmemBarriers + 1024;
    // block-scaled instruction descriptor (cute InstrDescriptorBlockScaled): a/b format E4M3 (0), K-major, N >> 3 at bit 17,
    // scale format UE8M0 (bit 23), M >> 4 at bit 24; b_sf_id at bits 4-5 and a_sf_id at bits 29-30 are added per MMA
    static constexpr uint32_t kIdesc = ((uint32_t)(TBN >> 3) << 17) | (1u << 23) | ((uint32_t)((BM * CG) >> 4) << 24);
    static_assert(kStageBytes % 1024 == 0, "stages must keep the 1024-byte alignment of SWIZZLE_128B tiles");
    static_assert(kSfBytes <= 2048, "scale atoms must fit their slot");
};

struct Params {
    const float* bias;
    float* out;
    int out_dtype, out_tma;
    int T, N;
    int num_k_stages;           // Kp / 128
    int tiles_m, tiles_n;       // tiles_m counts CG * 128 rows
    int sf_a_tiles;             // row tiles (of 128) in the A scale array
    int b_folded;               // 1 = the B operand carries its block exponents in its E4M3 values and has ONE scale per row (atoms of one slab)
    int out_bfloat;             // MX linear (mx/linear.py): 0 = plain; else out = rb(acc), with a bias rb(rb(acc) + rb(bias)), rb = bfloatX half-away rounding
    int debug;                  // timing experiments (wrong results): bit 1 = no operand loads, bit 2 = no scale copies, bit 3 = no scale loads, bit 4 = no B tile loads, bit 5 = every slab loads k = 0
};

template <int CG>
__device__ __forceinline__ void mma_mx(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t tmem_sfa, uint32_t tmem_sfb,
                                       uint32_t accumulate) {
    if constexpr (CG == 1)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %6, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%4], [%5], p;\n\t}"
            ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(tmem_sfa), "r"(tmem_sfb), "r"(accumulate) : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %6, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%4], [%5], p;\n\t}"
            ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(tmem_sfa), "r"(tmem_sfb), "r"(accumulate) : "memory");
}
// 32 rows x 16 bytes of shared memory -> four TMEM columns of lanes 0-31, replicated to the other three lane quarters
template <int CG>
__device__ __forceinline__ void tmem_cp_sf(uint32_t tmem_dst, uint64_t smem_desc) {
    if constexpr (CG == 1) asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(tmem_dst), "l"(smem_desc) : "memory");
    else asm volatile("tcgen05.cp.cta_group::2.32x128b.warpx4 [%0], %1;" ::"r"(tmem_dst), "l"(smem_desc) : "memory");
}

struct Barriers {
    uint64_t full[8];
    uint64_t empty[8];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
};

template <int TBN, int CG>
__global__ void __launch_bounds__(kThreads, 1)
bfp_gemm_mx_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_sfa,
                   const __grid_constant__ CUtensorMap map_sfb, const __grid_constant__ CUtensorMap map_out, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    using C = Cfg<TBN, CG>;
    constexpr int kStages = C::kStages, kStageBytes = C::kStageBytes, kBufs = C::kBufs;
    uint8_t* staging = smem + kStages * kStageBytes;
    Barriers* bars = reinterpret_cast<Barriers*>(staging + C::kSmemStaging);
    auto stage_a = [&](int s) { return smem + s * kStageBytes; };
    auto stage_b = [&](int s) { return smem + s * kStageBytes + BM * BKB; };
    auto stage_sf = [&](int s) { return smem + s * kStageBytes + BM * BKB + C::kRowsB * BKB; };      // A atom, then the B atoms

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = CG == 1 ? 0u : cluster_ctarank();
    const int unit = (int)blockIdx.x / CG, num_units = (int)gridDim.x / CG;
    const int num_tiles = p.tiles_m * p.tiles_n;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
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
            const uint64_t pol_keep = l2_policy_evict_last();
            for (int tile = unit; tile < num_tiles; tile += num_units) {
                const int tm = tile % p.tiles_m, tn = tile / p.tiles_m;
                const int a_tile = tm * CG + (int)rank;                       // this CTA's 128 A rows
                const int b_row = tn * TBN + (int)rank * C::kRowsB;           // this CTA's share of the B tile
                for (int ks = 0; ks < p.num_k_stages && !(p.debug & 2); ++ks) {
                    mbar_wait(&bars->empty[stage], phase ^ 1);
                    // the bytes of both CTAs of a pair land on the leader's barrier (its MMA thread issues for both); the scale atoms
                    // travel as 2-D tensor copies of 256-byte rows so that they can use the same cta_group::2 completion path.
                    // With a folded B operand (one scale per row, p.b_folded) the B atoms are loaded with the tile's first slab only.
                    const bool load_sfa = !(p.debug & 8);
                    const bool load_sfb = (!p.b_folded || ks == 0) && !(p.debug & 8);
                    const bool load_b = !(p.debug & 16);                                   // timing experiments: no B tile / always the first slab
                    const int kc = (p.debug & 32) ? 0 : ks * BKB;
                    if (rank == 0) mbar_expect_tx(&bars->full[stage], (uint32_t)CG * (uint32_t)(BM * BKB + (load_b ? C::kRowsB * BKB : 0) + (load_sfa ? 512 : 0) + (load_sfb ? C::kAtomsB * 512 : 0)));
                    const uint32_t bar = CG == 1 ? smem_u32(&bars->full[stage]) : mapa_u32(smem_u32(&bars->full[stage]), 0);
                    tma_load_2d_to_hint<CG>(stage_a(stage), &map_a, bar, kc, a_tile * BM, pol_keep);
                    if (load_b) tma_load_2d_to_hint<CG>(stage_b(stage), &map_b, bar, kc, b_row, pol_keep);
                    if (load_sfa) tma_load_2d_to<CG>(stage_sf(stage), &map_sfa, bar, 0, (ks * p.sf_a_tiles + a_tile) * 2);
                    if (load_sfb) tma_load_2d_to<CG>(stage_sf(stage) + 512, &map_sfb, bar, 0, ((p.b_folded ? 0 : ks * p.tiles_n) + tn) * (C::kAtomsB * 2));
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            int stage = 0; uint32_t phase = 0;
            int buf = 0; uint32_t buf_phase[2] = {0, 0};
            const uint32_t sf_col = tmem_base + (uint32_t)C::kSfCol;
            for (int tile = unit; tile < num_tiles; tile += num_units) {
                mbar_wait(&bars->tmem_empty[buf], buf_phase[buf] ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)buf * TBN;
                for (int ks = 0; ks < p.num_k_stages; ++ks) {
                    if (!(p.debug & 2)) mbar_wait(&bars->full[stage], phase);
                    tc_fence_after();
                    // scale atoms of this slab: A -> columns sf_col .. +3, B atom j -> sf_col + 4 + 4 j .. (the tensor pipe executes the
                    // copies and the MMAs in issue order).  A copy costs ~50 clk of the pipe (measured), which is why a folded B operand
                    // -- scales constant along K, copied once per tile -- matters: one copy per slab instead of three.
                    const uint32_t sfs = smem_u32(stage_sf(stage));
                    if (!(p.debug & 4)) {
                        tmem_cp_sf<CG>(sf_col, make_smem_desc_k(sfs, 0, 128, 16));             // no swizzle: 8-row groups 128 bytes apart
                        if (!p.b_folded || ks == 0) {
#pragma unroll
                            for (int j = 0; j < C::kAtomsB; ++j) tmem_cp_sf<CG>(sf_col + 4 + 4 * j, make_smem_desc_k(sfs + 512 + 512 * j, 0, 128, 16));
                        }
                    }
                    const uint64_t da = make_smem_desc(smem_u32(stage_a(stage)));
                    const uint64_t db = make_smem_desc(smem_u32(stage_b(stage)));
#pragma unroll
                    for (int i = 0; i < 4; ++i) {        // 32 elements = 32 bytes of K per MMA: +2 in 16-byte units; scale id = 32-group of the slab
                        const uint32_t idesc = C::kIdesc | ((uint32_t)i << 4) | ((uint32_t)i << 29);
                        mma_mx<CG>(d, da + (uint64_t)(i * 2), db + (uint64_t)(i * 2), idesc, sf_col, sf_col + 4, (ks | i) != 0);
                    }
                    if (!(p.debug & 2)) commit<CG>(&bars->empty[stage]);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                commit<CG>(&bars->tmem_full[buf]);
                buf_phase[buf] ^= 1;
                if (kBufs == 2) buf ^= 1;
            }
        }
    } else if (warp >= 4) {
        // 8 epilogue warps: 4 lane quarters x 2 column halves.  Half 0 takes columns [0, kHalf0), half 1 the rest; a half is drained
        // into registers in one burst, the accumulator handed back, then stored 16 columns at a time (swizzled smem tile + TMA store).
        const int ew = warp - 4, q = warp & 3, half = ew >> 2;
        constexpr int kHalf0 = TBN >= 240 ? 128 : TBN / 2;
        constexpr int kMaxCols = kHalf0 > TBN - kHalf0 ? kHalf0 : TBN - kHalf0;
        const int col0 = half ? kHalf0 : 0, width = half ? TBN - kHalf0 : kHalf0;
        int buf = 0; uint32_t buf_phase[2] = {0, 0};
        const uint64_t pol = l2_policy_evict_first();
        uint8_t* sbuf = staging + ew * 2048;
        for (int tile = unit; tile < num_tiles; tile += num_units) {
            const int tm = tile % p.tiles_m, tn = tile / p.tiles_m;
            mbar_wait(&bars->tmem_full[buf], buf_phase[buf]);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * TBN + col0);
            const int row0 = (tm * CG + (int)rank) * BM + q * 32;
            const int n0 = tn * TBN + col0;
            uint32_t r[kMaxCols];
#pragma unroll
            for (int i = 0; i < kMaxCols / 16; ++i)
                if (i * 16 < width) tmem_ld16(taddr + i * 16, r + i * 16);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (CG == 1) mbar_arrive(&bars->tmem_empty[buf]);
                else mbar_arrive_cluster(mapa_u32(smem_u32(&bars->tmem_empty[buf]), 0));
            }
            // 16 columns at a time: the warp's [32 rows][16 cols] block goes to a 2 KB smem tile (64-byte rows) and leaves through one
            // TMA store; rows / columns beyond T / N are clipped by the copy engine.  (A small staging tile buys a sixth operand stage.)
#pragma unroll
            for (int c = 0; c < kMaxCols / 16; ++c) {
                if (c * 16 < width) {
                    if (lane == 0) tma_store_wait_read<0>();
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float4 o = make_float4(__uint_as_float(r[c * 16 + 4 * j]), __uint_as_float(r[c * 16 + 4 * j + 1]), __uint_as_float(r[c * 16 + 4 * j + 2]),
                                               __uint_as_float(r[c * 16 + 4 * j + 3]));
                        if (p.out_bfloat) { o.x = round_bfloat(o.x, p.out_bfloat); o.y = round_bfloat(o.y, p.out_bfloat); o.z = round_bfloat(o.z, p.out_bfloat); o.w = round_bfloat(o.w, p.out_bfloat); }
                        if (p.bias) {
                            const int nb = n0 + c * 16 + 4 * j;
                            const int rbb = p.out_bfloat;        // round_bfloat(x, 0) is the identity
                            if (nb < p.N) o.x = round_bfloat(o.x + round_bfloat(p.bias[nb], rbb), rbb);
                            if (nb + 1 < p.N) o.y = round_bfloat(o.y + round_bfloat(p.bias[nb + 1], rbb), rbb);
                            if (nb + 2 < p.N) o.z = round_bfloat(o.z + round_bfloat(p.bias[nb + 2], rbb), rbb);
                            if (nb + 3 < p.N) o.w = round_bfloat(o.w + round_bfloat(p.bias[nb + 3], rbb), rbb);
                        }
                        *reinterpret_cast<float4*>(sbuf + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = o;      // SWIZZLE_64B: chunk ^ (row / 2 % 4)
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) { tma_store_2d_hint(&map_out, sbuf, n0 + c * 16, row0, pol); tma_store_commit(); }
                }
            }
            buf_phase[buf] ^= 1;
            if (kBufs == 2) buf ^= 1;
        }
        if (lane == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    if constexpr (CG == 1) __syncthreads(); else cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        if constexpr (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}

// ---- operand conversion: int8 mantissas + fp32 block-major scales (bfp_quantize_pack) -> E4M3 bytes + UE8M0 scale atoms -----------
// E4M3 byte of a * 2^u for an integer 1 <= a <= 16 whose result is a NORMAL E4M3 number (caller checks -6 <= floor(log2 a) + u <= 8)
__device__ __forceinline__ uint32_t e4m3_of(uint32_t a, int u) {
    const int lg = 31 - __clz(a);
    return ((uint32_t)(lg + u + 7) << 3) | (((a << 3) >> lg) & 7u);
}

// FOLDED form (weights): the block exponents ride in the E4M3 values, relative to one reference exponent per row, so the row has a
// single hardware scale and the GEMM copies the B scales once per tile instead of once per K slab.  With ref = pmax - 4 (pmax = the
// largest block exponent of the row among blocks that hold a non-zero mantissa) a mantissa |q| <= 15 of a block with exponent p becomes
// q 2^(p - ref), exactly representable as a normal E4M3 number while -6 <= floor(log2 |q|) + p - ref, i.e. for block exponents up to
// ten octaves below the row's largest.  Rows that need more (or hold NaN-marked blocks) are counted in *violations: use the general form.
__global__ void __launch_bounds__(256) mx_row_ref_kernel(const int8_t* __restrict__ mant, const float* __restrict__ scale_t, int64_t ld_s, int64_t rows,
                                                         int64_t K, int64_t Kp_in, int B, int* __restrict__ ref, unsigned int* __restrict__ violations) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int64_t nblk = (K + B - 1) / B;
    int pmax = -1000;
    bool bad = false;
    for (int64_t b = lane; b < nblk; b += 32) {
        const uint32_t sb = __float_as_uint(scale_t[b * ld_s + row]);
        bool nz = false;
        const int64_t k0 = b * B, k1 = min(K, k0 + B);
        for (int64_t k = k0; k < k1; ++k) nz |= mant[row * Kp_in + k] != 0;
        if (nz) {
            if ((sb & 0x7fffffffu) >= 0x7f800000u || (sb & 0x7fffffu)) bad = true;      // NaN-marked / not a power of two
            pmax = max(pmax, (int)((sb >> 23) & 0xffu) - 127);
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) pmax = max(pmax, __shfl_xor_sync(0xffffffffu, pmax, off));
    int r = pmax == -1000 ? 0 : pmax - 4;                   // all-zero row: any scale
    if (r + 127 < 1 || r + 127 > 254) bad = true;
    // every non-zero mantissa must stay a normal E4M3 number: floor(log2 |q|) + p - r >= -6
    for (int64_t b = lane; b < nblk && !bad; b += 32) {
        const int p = (int)((__float_as_uint(scale_t[b * ld_s + row]) >> 23) & 0xffu) - 127;
        const int64_t k0 = b * B, k1 = min(K, k0 + B);
        for (int64_t k = k0; k < k1; ++k) {
            const int q = mant[row * Kp_in + k];
            const uint32_t a = (uint32_t)(q < 0 ? -q : q);
            if (a && (a > 16u || (31 - __clz(a)) + p - r < -6)) bad = true;
        }
    }
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) { ref[row] = r; if (bad) atomicAdd(violations, 1u); }
}

// one thread per 16 output bytes; ref == nullptr: the general form (plain integer mantissas), else the folded form
__global__ void __launch_bounds__(256) mx_vals_kernel(const int8_t* __restrict__ mant, const float* __restrict__ scale_t, int64_t ld_s, int B,
                                                      const int* __restrict__ ref, uint8_t* __restrict__ vals, int64_t rows, int64_t Kp_in, int64_t Kp_out,
                                                      unsigned int* __restrict__ violations) {
    const int64_t chunks_per_row = Kp_out / 16;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * chunks_per_row) return;
    const int64_t row = i / chunks_per_row, c = i - row * chunks_per_row;
    uint4 w = make_uint4(0u, 0u, 0u, 0u);
    if (c * 16 < Kp_in) w = *reinterpret_cast<const uint4*>(mant + row * Kp_in + c * 16);      // Kp_in is a multiple of 16
    uint32_t in[4] = {w.x, w.y, w.z, w.w}, out[4];
    const int r = ref ? ref[row] : 0;
    bool bad = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t o = 0u;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int q = (int)(int8_t)((in[j] >> (8 * b)) & 0xffu);
            const uint32_t a = (uint32_t)(q < 0 ? -q : q);
            uint32_t e = 0u;
            if (a) {
                int u = 0;
                if (ref) u = (int)((__float_as_uint(scale_t[((c * 16 + j * 4 + b) / B) * ld_s + row]) >> 23) & 0xffu) - 127 - r;
                const int lg = 31 - __clz(a);
                if (a > 16u || lg + u < -6 || lg + u > 8) { bad = true; u = 0; }
                e = e4m3_of(a, u);
            }
            o |= (e | (q < 0 ? 0x80u : 0u)) << (8 * b);
        }
        out[j] = o;
    }
    if (bad && !ref) atomicAdd(violations, 1u);               // (the folded form's violations are counted per row by mx_row_ref_kernel)
    *reinterpret_cast<uint4*>(vals + row * Kp_out + c * 16) = make_uint4(out[0], out[1], out[2], out[3]);
}

// scale_t [nkb][ld_s] fp32 (2^p per BFP block, NaN = unrepresentable) -> atoms; one thread per (slab, tile, atom, row-in-atom).
// ref != nullptr: folded form, n_slabs == 1 and every 32-group carries the row's reference exponent.
__global__ void __launch_bounds__(256) mx_sf_kernel(const float* __restrict__ scale_t, int64_t ld_s, const int* __restrict__ ref, uint8_t* __restrict__ sf,
                                                    int64_t rows, int64_t K, int B, int tile_rows, int atoms, int64_t n_tiles, int64_t n_slabs) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = n_slabs * n_tiles * atoms * 128;
    if (i >= total) return;
    const int r = (int)(i & 127);
    int64_t rest = i >> 7;
    const int atom = (int)(rest % atoms); rest /= atoms;
    const int64_t tile = rest % n_tiles, slab = rest / n_tiles;
    const int in_tile = atom * 128 + r;
    const int64_t row = tile * tile_rows + in_tile;
    uint32_t word = 0u;
    if (in_tile < tile_rows && row < rows) {
        if (ref) {
            word = (uint32_t)(ref[row] + 127) * 0x01010101u;
        } else {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int64_t k = slab * 128 + g * 32;
                uint32_t byte = 0u;
                if (k < K) byte = (__float_as_uint(scale_t[(k / B) * ld_s + row]) >> 23) & 0xffu;      // 2^p -> p + 127; NaN -> 0xff (NaN scale)
                word |= byte << (8 * g);
            }
        }
    }
    uint8_t* atom_base = sf + (((slab * n_tiles + tile) * atoms + atom) * 512);
    *reinterpret_cast<uint32_t*>(atom_base + 16 * (r & 31) + 4 * (r >> 5)) = word;
}

// ---- fused activation pack: float_to_bfp_blocked (quantise only, nearest rounding) straight into the general mx form -----------------
// One pass: lane l of a warp owns one 128-bit vector (4 fp32 / 8 half values); a BFP block is 2^j adjacent lanes (block max by
// butterfly, the quantiser arithmetic of bfp_common.cuh); a warp covers 128 (fp32) or 256 (half) consecutive k of one row, i.e. one
// or two K slabs, and writes their E4M3 bytes (4 / 8 per lane, coalesced) and the row's four scale bytes per slab (one 32-bit store).
// Blocks the packed form cannot represent (slow-path scales: NaN / Inf / denormal range) get the NaN scale 0xff like bfp_quantize_pack.
template <int DT>
__global__ void __launch_bounds__(256) mx_pack_kernel(const uint4* __restrict__ in, uint8_t* __restrict__ vals, uint8_t* __restrict__ sf, int64_t rows, int64_t K,
                                                      int lanes_per_block, int m, float eps, int64_t n_tiles, FastDiv wpr_div) {
    using D = DType<DT>;
    constexpr int V = D::kVec;
    constexpr int kSlabsPerWarp = 32 * V / 128;             // 1 (fp32) or 2 (half)
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int64_t warps_per_row = K / (32 * V);
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t wi = gw; wi < rows * warps_per_row; wi += n_warps) {
        const int64_t row = (int64_t)fastdiv((uint32_t)wi, wpr_div), wk = wi - row * warps_per_row;       // wi < 2^32 (checked by the host)
        const int64_t vec = row * (K / V) + wk * 32 + lane;
        float v[V];
        unpack_vec<DT>(ld_stream(in + vec), v);
        uint32_t amax = 0u;
#pragma unroll
        for (int i = 0; i < V; ++i) amax = max(amax, abs_bits(v[i]));
#pragma unroll
        for (int off = 1; off < 32; off <<= 1)
            if (off < lanes_per_block) amax = max(amax, __shfl_xor_sync(0xffffffffu, amax, off));
        const BlockScale sc = make_scale<DT>(amax, m, eps);
        uint32_t bytes[V / 4];
        uint32_t sbyte = 0xffu;                               // NaN scale unless the block is on the exact path
        if (sc.fast) {
            sbyte = (uint32_t)(sc.p + 127);
#pragma unroll
            for (int w = 0; w < V / 4; ++w) {
                uint32_t o = 0u;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const float t = v[4 * w + b];
                    const float qf = fminf(fmaxf(rintf(t * sc.inv), -sc.vmax), sc.vmax);     // the integer mantissa (bfp_ops.py:40-44 on the grid)
                    // E4M3 byte of an integer |q| <= 15 held as a float: (e + 7) << 3 | m3 = the float's exponent and top mantissa bits
                    // rebased (127 - 7) << 3 = 960; zero (either sign) -> 0
                    const uint32_t qb = __float_as_uint(qf), ab = qb & 0x7fffffffu;
                    const uint32_t nb = max(ab >> 20, 960u) - 960u;
                    o |= (nb | (ab ? (qb >> 24) & 0x80u : 0u)) << (8 * b);
                }
                bytes[w] = o;
            }
        } else {
#pragma unroll
            for (int w = 0; w < V / 4; ++w) bytes[w] = 0u;
        }
        uint8_t* dst = vals + row * K + (wk * 32 + lane) * V;
        if (V == 4) *reinterpret_cast<uint32_t*>(dst) = bytes[0];
        else *reinterpret_cast<uint2*>(dst) = make_uint2(bytes[0], bytes[V == 8 ? 1 : 0]);
        // scale bytes: 32-group g of slab s lives in lanes [(s * 128 + g * 32) / V, ...): gather the four bytes of a slab into lane 0 / 16
#pragma unroll
        for (int sl = 0; sl < kSlabsPerWarp; ++sl) {
            uint32_t word = 0u;
#pragma unroll
            for (int g = 0; g < 4; ++g) word |= __shfl_sync(0xffffffffu, sbyte, (sl * 128 + g * 32) / V) << (8 * g);
            if (lane == 0) {
                const int64_t slab = wk * kSlabsPerWarp + sl, tile = row >> 7;
                const int r = (int)(row & 127);
                *reinterpret_cast<uint32_t*>(sf + ((slab * n_tiles + tile) * 512) + 16 * (r & 31) + 4 * (r >> 5)) = word;
            }
        }
    }
}

}  // namespace gemm_mx

int mx_layout(int64_t rows, int64_t K, int tile_rows, int fold, int64_t* Kp, int64_t* sf_bytes) {
    if (rows < 0 || K < 0 || tile_rows < 1) return set_error(BFP_E_ARG, "bad argument");
    const int64_t kp = round_up(K, 128), atoms = (tile_rows + 127) / 128, tiles = (rows + tile_rows - 1) / tile_rows;
    if (Kp) *Kp = kp;
    if (sf_bytes) *sf_bytes = (fold ? 1 : kp / 128) * tiles * atoms * 512;
    return BFP_OK;
}

int mx_from_packed_device(const int8_t* mant, const float* scale_t, int64_t ld_s, int64_t rows, int64_t K, int block_size, int tile_rows, int fold,
                          uint8_t* vals, uint8_t* sf, int* row_ref, unsigned int* violations, cudaStream_t st) {
    using namespace gemm_mx;
    if (rows == 0 || K == 0) return BFP_OK;
    if (!fold && (block_size % 32 || block_size <= 0))
        return set_error(BFP_E_UNSUPPORTED, "block-scaled operands need a block_size that is a multiple of 32 (one hardware scale per 32 elements); "
                                            "only the folded (weight) form takes other block sizes");
    if (fold && !row_ref) return set_error(BFP_E_ARG, "the folded form needs the row_ref scratch (rows int32)");
    const int64_t kp_in = packed_kp(K), kp_out = round_up(K, 128);
    if (fold) {
        mx_row_ref_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(mant, scale_t, ld_s, rows, K, kp_in, block_size, row_ref, violations);
        count_launch();
    }
    const int64_t chunks = rows * (kp_out / 16);
    mx_vals_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, st>>>(mant, scale_t, ld_s, block_size, fold ? row_ref : nullptr, vals, rows, kp_in, kp_out, violations);
    count_launch();
    const int atoms = (tile_rows + 127) / 128;
    const int64_t tiles = (rows + tile_rows - 1) / tile_rows, slabs = fold ? 1 : kp_out / 128;
    const int64_t total = slabs * tiles * atoms * 128;
    mx_sf_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(scale_t, ld_s, fold ? row_ref : nullptr, sf, rows, K, block_size, tile_rows, atoms, tiles, slabs);
    count_launch();
    return check_launch("mx operand conversion");
}

// float_to_bfp_blocked (quantise only) -> general mx form in one pass.  Needs nearest rounding, mant_bits in [1, 4], a power-of-two
// block_size in {32, 64, 128}, K a multiple of 128 (fp32) / 256 (half) and 16-byte aligned buffers: BFP_E_UNSUPPORTED otherwise (pack
// with bfp_quantize_pack and convert with bfp_mx_from_packed).  Rows beyond `rows` inside the last 128-row tile keep whatever the sf
// buffer held: the caller zero-fills sf once when rows % 128 != 0.
int mx_pack_device(const void* in, int dtype, int64_t rows, int64_t K, int block_size, int mant_bits, float eps, uint8_t* vals, uint8_t* sf, cudaStream_t st) {
    using namespace gemm_mx;
    if (rows == 0 || K == 0) return BFP_OK;
    const int V = dtype == BFP_DT_F32 ? 4 : 8;
    if (mant_bits < 1 || mant_bits > 4) return set_error(BFP_E_UNSUPPORTED, "the block-scaled form holds mant_bits in [1, 4]");
    if ((block_size != 32 && block_size != 64 && block_size != 128) || K % (32 * V) || reinterpret_cast<uintptr_t>(in) % 16 ||
        reinterpret_cast<uintptr_t>(vals) % 16 || reinterpret_cast<uintptr_t>(sf) % 16)
        return set_error(BFP_E_UNSUPPORTED, "fused mx pack: block_size 32 / 64 / 128, K a multiple of 128 (fp32) or 256 (half), 16-byte aligned buffers");
    if (dtype != BFP_DT_F32) ensure_exp_tables(st);          // this translation unit's copy of the half-precision exponent table
    const int64_t n_tiles = (rows + 127) / 128;
    const int64_t warps = rows * (K / (32 * V));
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((warps + 7) / 8, (int64_t)device_info().sm_count * 8));
    const uint4* src = static_cast<const uint4*>(in);
    const int lpb = block_size / V;
    int rc;
    struct P { const uint4* in; uint8_t* vals; uint8_t* sf; int64_t rows, K; int lpb, m; float eps; int64_t n_tiles; };
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(256); cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[0].val.programmaticStreamSerializationAllowed = tuning().pdl ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e;
    if (warps >= (int64_t)UINT32_MAX) return set_error(BFP_E_UNSUPPORTED, "fused mx pack: tensor too large (use bfp_quantize_pack + bfp_mx_from_packed)");
    const FastDiv wpr = make_fastdiv((uint32_t)(K / (32 * V)));
    if (dtype == BFP_DT_F32) e = cudaLaunchKernelEx(&cfg, mx_pack_kernel<BFP_DT_F32>, src, vals, sf, rows, K, lpb, mant_bits, eps, n_tiles, wpr);
    else if (dtype == BFP_DT_F16) e = cudaLaunchKernelEx(&cfg, mx_pack_kernel<BFP_DT_F16>, src, vals, sf, rows, K, lpb, mant_bits, eps, n_tiles, wpr);
    else e = cudaLaunchKernelEx(&cfg, mx_pack_kernel<BFP_DT_BF16>, src, vals, sf, rows, K, lpb, mant_bits, eps, n_tiles, wpr);
    (void)rc;
    if (e != cudaSuccess) return set_errorf(BFP_E_CUDA, "cudaLaunchKernelEx(mx_pack_kernel): %s", cudaGetErrorString(e));
    count_launch();
    return check_launch("mx_pack_kernel");
}

template <int TBN, int CG>
static int launch_mx(const CUtensorMap& map_a, const CUtensorMap& map_b, const CUtensorMap& map_sfa, const CUtensorMap& map_sfb, const CUtensorMap& map_out,
                     const gemm_mx::Params& p, int units, cudaStream_t st) {
    using namespace gemm_mx;
    using C = Cfg<TBN, CG>;
    cudaError_t e = cudaFuncSetAttribute(bfp_gemm_mx_kernel<TBN, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemTotal);
    if (e != cudaSuccess) return set_errorf(BFP_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(units * CG)); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = C::kSmemTotal; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, bfp_gemm_mx_kernel<TBN, CG>, map_a, map_b, map_sfa, map_sfb, map_out, p);
    if (e != cudaSuccess) return set_errorf(BFP_E_CUDA, "cudaLaunchKernelEx(bfp_gemm_mx_kernel): %s", cudaGetErrorString(e));
    return BFP_OK;
}

int gemm_mx_device(const uint8_t* a_vals, const uint8_t* a_sf, const uint8_t* b_vals, const uint8_t* b_sf, int b_tile_rows, int b_folded, const float* bias,
                   float* out, int64_t T, int64_t N, int64_t Kp, cudaStream_t st, int out_bfloat) {
    using namespace gemm_mx;
    if (T == 0 || N == 0) return BFP_OK;
    if (Kp % 128 != 0 || Kp <= 0) return set_error(BFP_E_ARG, "mx operand K must be a positive multiple of 128");
    if (T > INT32_MAX || N > INT32_MAX || Kp > INT32_MAX) return set_error(BFP_E_ARG, "dimension too large");
    if (reinterpret_cast<uintptr_t>(a_vals) % 16 || reinterpret_cast<uintptr_t>(b_vals) % 16 || reinterpret_cast<uintptr_t>(a_sf) % 16 ||
        reinterpret_cast<uintptr_t>(b_sf) % 16 || reinterpret_cast<uintptr_t>(out) % 16 || (N * 4) % 16)
        return set_error(BFP_E_ALIGN, "mx operands and the output must be 16-byte aligned (N a multiple of 4)");
    if (b_tile_rows != 128 && b_tile_rows != 240 && b_tile_rows != 256) return set_error(BFP_E_UNSUPPORTED, "B scale atoms must be tiled by 128, 240 or 256 rows");
    Params p;
    p.bias = bias; p.out = out; p.out_dtype = BFP_DT_F32; p.out_tma = 1;
    p.T = (int)T; p.N = (int)N; p.num_k_stages = (int)(Kp / 128);
    p.b_folded = b_folded ? 1 : 0;
    p.out_bfloat = out_bfloat;
    p.debug = tuning().gemm_mx_variant & 62;
    // CTA pairs (cta_group::2: each CTA stages its 128 A rows and HALF of the B tile) for the 240- and 256-wide tiles whenever there is
    // more than one 128-row strip; bfp_set_option("gemm_mx_variant", 1) forces single CTAs on the 256-wide tile
    const int tbn = b_tile_rows;
    const int cg = (tbn == 240 || (tbn == 256 && T > 128 && !(tuning().gemm_mx_variant & 1))) ? 2 : 1;
    p.tiles_m = (int)((T + BM * cg - 1) / (BM * cg));
    p.tiles_n = (int)((N + tbn - 1) / tbn);
    p.sf_a_tiles = (int)((T + 127) / 128);
    const int atoms_b = (tbn + 127) / 128;
    const int64_t slabs = Kp / 128;
    CUtensorMap map_a, map_b, map_sfa, map_sfb, map_out;
    if (int rc = make_map(&map_a, a_vals, T, Kp, BM)) return rc;
    if (int rc = make_map(&map_b, b_vals, N, Kp, tbn / cg)) return rc;
    if (int rc = make_map_bytes(&map_sfa, a_sf, slabs * p.sf_a_tiles * 2, 256, 2)) return rc;
    if (int rc = make_map_bytes(&map_sfb, b_sf, (b_folded ? 1 : slabs) * p.tiles_n * atoms_b * 2, 256, atoms_b * 2)) return rc;
    if (int rc = make_map_out(&map_out, out, BFP_DT_F32, T, N, N * 4, 16, 32, CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
    const int sms = std::max(2, device_info().sm_count);
    const int units = (int)std::min<int64_t>((int64_t)p.tiles_m * p.tiles_n, sms / cg);
    int rc;
    if (tbn == 240) rc = launch_mx<240, 2>(map_a, map_b, map_sfa, map_sfb, map_out, p, units, st);
    else if (cg == 2) rc = launch_mx<256, 2>(map_a, map_b, map_sfa, map_sfb, map_out, p, units, st);
    else if (tbn == 256) rc = launch_mx<256, 1>(map_a, map_b, map_sfa, map_sfb, map_out, p, units, st);
    else rc = launch_mx<128, 1>(map_a, map_b, map_sfa, map_sfb, map_out, p, units, st);
    if (rc) return rc;
    count_launch();
    return check_launch("bfp_gemm_mx_kernel");
}

}  // namespace bfp
