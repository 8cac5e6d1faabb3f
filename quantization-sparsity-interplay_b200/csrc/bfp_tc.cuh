// bfp_tc.cuh -- tcgen05 / TMA / mbarrier PTX wrappers and tensor-map helpers shared by the tensor-core kernels
// (bfp_gemm.cu: dense int8 + bf16 kinds, bfp_gemm_sp.cu: 2:4 structured-sparse bf16 kind).  sm_100a only.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "bfp_internal.h"

namespace bfp {
namespace gemm {

// ---- PTX wrappers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded wait: a protocol bug must surface as a trap within ~2 s, not as a hung GPU box
__device__ __forceinline__ uint64_t global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint64_t t0 = 0;
    for (uint32_t spin = 0;; ++spin) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity), "r"(20000u) : "memory");      // suspend-time hint (ns)
        if (done) return;
        if ((spin & 1023u) == 1023u) {
            const uint64_t now = global_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 2000000000ull) __trap();
        }
    }
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 r;\n\telect.sync r|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(addr));
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
// packed fp32x2 arithmetic (sm_100): one issue slot for two lanes of the epilogue's rescale
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 unpack2(uint64_t v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ void fma2_acc(uint64_t& acc, uint64_t a, uint64_t b) {      // acc += a * b, in place (no register moves)
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// int32 -> fp32 for |v| < 2^22 without the (slow) conversion pipe: bit pattern of (v + 1.5 * 2^23) as a float is
// 0x4B400000 + v, so one integer add per value and one packed fp32 subtract per pair; exact.
__device__ __forceinline__ uint64_t cvt2_s32(uint32_t a, uint32_t b) {
    uint64_t m;
    asm("mov.b64 %0, {%1, %2};" : "=l"(m) : "r"(a + 0x4B400000u), "r"(b + 0x4B400000u));
    return add2(m, 0xCB400000CB400000ull);              // (-12582912.0f, -12582912.0f)
}
// acc (f32x2) += float(a, b) * w for two int32 accumulator values (I2FP: 64 lanes/clk/SM, tools/microbench/pipe_rates.cu)
__device__ __forceinline__ void rescale_pair_xu(uint64_t& acc, uint32_t a, uint32_t b, uint64_t w) {
    asm("{\n\t.reg .f32 fa, fb;\n\t.reg .b64 f;\n\t"
        "cvt.rn.f32.s32 fa, %1;\n\tcvt.rn.f32.s32 fb, %2;\n\t"
        "mov.b64 f, {fa, fb};\n\t"
        "fma.rn.f32x2 %0, f, %3, %0;\n\t}"
        : "+l"(acc) : "r"(a), "r"(b), "l"(w));
}
// the same with the conversion on the integer + FMA pipes: for |v| < 2^22 the bits of float(v + 1.5 * 2^23) are
// 0x4B400000 + v, so float(v) = as_float(v + 0x4B400000) - 12582912.0f exactly (|block sums| <= 128 * 127^2 < 2^22).
// Measured slower than I2FP here (the integer adds land on the same half-rate pipe); kept for reference.
__device__ __forceinline__ void rescale_pair_magic(uint64_t& acc, uint32_t a, uint32_t b, uint64_t w) {
    asm("{\n\t.reg .b64 m, f;\n\t"
        "mov.b64 m, {%1, %2};\n\t"
        "add.rn.f32x2 f, m, %4;\n\t"
        "fma.rn.f32x2 %0, f, %3, %0;\n\t}"
        : "+l"(acc) : "r"(a + 0x4B400000u), "r"(b + 0x4B400000u), "l"(w), "l"(0xCB400000CB400000ull));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
// start address >> 4 | LBO (ignored for swizzled K-major) = 1 | SBO = 1024 B (8 rows x 128 B) | version 1 | layout 2
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// ---- TMA stores (epilogues): smem tile -> global through the copy engine, full 128-byte rows, no LSU store instructions --
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
// L2 eviction policies: the output tile is written once and never re-read by the kernel (evict_first), the operands are
// re-read by every tile of the other dimension (evict_last) -- without this the output stream pushes the operands out of L2.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_store_2d_hint(const CUtensorMap* map, const void* smem_src, int c0, int c1, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
                 ::"l"(map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "l"(policy) : "memory");
}
// out += tile: the copy engine adds the fp32 smem tile into global memory (element-wise RN add at the L2); used by the
// K-chunked contraction (bfp_gemm_bf16_acc / _sp_acc), where every launch after the first accumulates into the output
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_or_add_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1, uint64_t policy, int accumulate) {
    if (accumulate) tma_reduce_add_2d(map, smem_src, c0, c1);
    else tma_store_2d_hint(map, smem_src, c0, c1, policy);
}
// batched outputs: the third coordinate selects the batch entry, rows beyond that entry's extent are clipped
__device__ __forceinline__ void tma_store_3d_hint(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;"
                 ::"l"(map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "l"(policy) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- CTA-pair (cta_group::2) helpers: CG = 1 degenerates to the single-CTA forms -----------------------------------
// tcgen05.commit: CG = 2 arrives on the barrier at the same offset in both CTAs of the pair
template <int CG>
__device__ __forceinline__ void commit(uint64_t* bar) {
    if constexpr (CG == 1) tc_commit(bar);
    else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                      ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// TMA tile load whose completion bytes land on `bar_addr` (a shared::cluster address; for CG = 2 the pair leader's barrier)
template <int CG>
__device__ __forceinline__ void tma_load_2d_to_hint(void* dst, const CUtensorMap* map, uint32_t bar_addr, int c0, int c1, uint64_t policy) {
    if constexpr (CG == 1)
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
                     ::"r"(smem_u32(dst)), "l"(map), "r"(bar_addr), "r"(c0), "r"(c1), "l"(policy) : "memory");
    else
        asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
                     ::"r"(smem_u32(dst)), "l"(map), "r"(bar_addr), "r"(c0), "r"(c1), "l"(policy) : "memory");
}
template <int CG>
__device__ __forceinline__ void tma_load_2d_to(void* dst, const CUtensorMap* map, uint32_t bar_addr, int c0, int c1) {
    if constexpr (CG == 1)
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(dst)), "l"(map), "r"(bar_addr), "r"(c0), "r"(c1) : "memory");
    else
        asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(dst)), "l"(map), "r"(bar_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {      // shared::cta address -> shared::cluster address in CTA `rank`
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Generic K-major shared-memory matrix descriptor: layout_type 0 = no swizzle (8-row x 16-byte core matrices), 2 = SWIZZLE_128B,
// 4 = SWIZZLE_64B, 6 = SWIZZLE_32B; sbo = bytes between consecutive 8-row groups; lbo = bytes between core matrices along K
// (ignored by the swizzled K-major modes).
__device__ __forceinline__ uint64_t make_smem_desc_k(uint32_t smem_addr, uint32_t layout_type, uint32_t sbo, uint32_t lbo) {
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) |
           (1ull << 46) | ((uint64_t)layout_type << 61);
}

// ---- host side -------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

inline int make_map(CUtensorMap* map, const void* ptr, int64_t rows, int64_t kbytes, int box_rows, bool bf16 = false) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return set_error(BFP_E_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)(bf16 ? kbytes / 2 : kbytes), (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)kbytes};
    cuuint32_t box[2] = {(cuuint32_t)(bf16 ? 64 : 128), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_errorf(BFP_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return BFP_OK;
}

// 2-D bf16 tensor map with an explicit box and swizzle (make_map above is the 128-byte-box special case).
inline int make_map_bf16(CUtensorMap* map, const void* ptr, int64_t rows, int64_t k_elems, int64_t row_stride_bytes, int box_k,
                         int box_rows, CUtensorMapSwizzle swz) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return set_error(BFP_E_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)k_elems, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)row_stride_bytes};
    cuuint32_t box[2] = {(cuuint32_t)box_k, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_errorf(BFP_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return BFP_OK;
}

// 2-D fp32 tensor map without swizzle (output tiles written by TMA stores): `rows` rows of `cols` floats, row stride in bytes.
inline int make_map_f32(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int64_t row_stride_bytes, int box_cols, int box_rows,
                        CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_NONE) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return set_error(BFP_E_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)row_stride_bytes};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swz, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_errorf(BFP_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return BFP_OK;
}

// Output tile map for any of the three output dtypes (no swizzle).
inline int make_map_out(CUtensorMap* map, const void* ptr, int dtype, int64_t rows, int64_t cols, int64_t row_stride_bytes, int box_cols, int box_rows,
                        CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_NONE) {
    if (dtype == BFP_DT_F32) return make_map_f32(map, ptr, rows, cols, row_stride_bytes, box_cols, box_rows, swz);
    EncodeTiledFn fn = encode_fn();
    if (!fn) return set_error(BFP_E_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)row_stride_bytes};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, dtype == BFP_DT_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_errorf(BFP_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return BFP_OK;
}

// 3-D output map of a batched GEMM: `batch` matrices of [rows, cols], contiguous; box = one [box_rows, box_cols] tile of one matrix
inline int make_map_out3(CUtensorMap* map, const void* ptr, int dtype, int64_t batch, int64_t rows, int64_t cols, int box_cols, int box_rows,
                         CUtensorMapSwizzle swz) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return set_error(BFP_E_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    const int es = dtype == BFP_DT_F32 ? 4 : 2;
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)(cols * es), (cuuint64_t)(rows * cols * es)};
    cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapDataType dt = dtype == BFP_DT_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : (dtype == BFP_DT_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
    CUresult r = fn(map, dt, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_errorf(BFP_E_CUDA, "cuTensorMapEncodeTiled (3-D) failed (%d)", (int)r);
    return BFP_OK;
}

// 2-D byte tensor map without swizzle: `rows` rows of `row_bytes` (>= 16) contiguous bytes, box = box_rows whole rows.
inline int make_map_bytes(CUtensorMap* map, const void* ptr, int64_t rows, int row_bytes, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return set_error(BFP_E_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)row_bytes, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
    cuuint32_t box[2] = {(cuuint32_t)row_bytes, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_errorf(BFP_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return BFP_OK;
}

}  // namespace gemm
}  // namespace bfp
