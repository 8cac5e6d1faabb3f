/*
 * bfp_b200.h -- C ABI of libbfp_b200.so: the B200 (sm_100a) implementation of the
 * block-floating-point quantise + N:M sparsify + BFP linear hot path of
 * parsa-epfl/quantization-sparsity-interplay (src/transformers/bfp/bfp_ops.py).
 *
 * This is the drop-in boundary.  The reference has no FFI of its own: its operator
 * API for this path *is* the Python module transformers.bfp.bfp_ops, so every entry
 * point below names the reference function (file:line under
 * /root/reference/src/transformers/bfp/) whose work it replaces; the Python mirror
 * in quantization-sparsity-interplay_b200/bfp_ops.py binds them with ctypes
 * (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - plain pointers and sizes only; all device pointers are caller-owned CUDA
 *     device memory on the current device; `stream` is a cudaStream_t passed as void*.
 *   - no allocation and no synchronisation inside the *_device entry points: work is
 *     enqueued on `stream` and the call returns.  The *_host entry points take HOST
 *     buffers, stage through internal pinned/device buffers and return when the
 *     result is in the host output buffer.
 *   - every call returns 0 on success, else a BFP_E_* code; bfp_last_error() gives
 *     the message for the calling thread.  There is no CPU fallback: without a CUDA
 *     device every compute call fails with BFP_E_CUDA.
 *   - tensors are viewed as [rows, K] row-major contiguous: rows = product of the
 *     leading dims, K = last dim.  Blocks (block_size) and N:M groups run along K and
 *     never straddle rows; a ragged tail is treated as zero-padded
 *     (bfp_ops.py:50-53, :79-82) and the padding is never written.
 */
#ifndef BFP_B200_H_
#define BFP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BFP_B200_VERSION 100 /* 0.1.0 */

/* element types of in/out tensors */
enum { BFP_DT_F32 = 0, BFP_DT_F16 = 1, BFP_DT_BF16 = 2 };
/* bfp_ops.py:16-18 rounding_modes: 'determ' / 'stoc' */
enum { BFP_ROUND_NEAREST = 0, BFP_ROUND_STOCHASTIC = 1 };
/* bfp_ops.py:141-149: first=='s' -> sparsify then quantise; anything else -> quantise then sparsify.
 * QUANT_ONLY: sparsity flag off for this identifier; SPARSIFY_ONLY: sparsity_num_format=='fp32'. */
enum { BFP_ORDER_QUANT_ONLY = 0, BFP_ORDER_SPARSIFY_QUANT = 1, BFP_ORDER_QUANT_SPARSIFY = 2, BFP_ORDER_SPARSIFY_ONLY = 3 };
/* which torch.topk backend's tie-breaking the N:M mask reproduces (bfp_ops.py:84).
 * TORCH_CUDA: the k smallest by (|v|, index) -- what the reference does with device='cuda' (measured on B200).
 * TORCH_CPU : ATen's std::nth_element order; supported for N:M = 2:4 only. */
enum { BFP_TIE_TORCH_CUDA = 0, BFP_TIE_TORCH_CPU = 1 };

enum {
    BFP_OK = 0,
    BFP_E_ARG = 1,          /* invalid argument (the reference would assert / raise) */
    BFP_E_UNSUPPORTED = 2,  /* valid in the reference but not implemented here */
    BFP_E_CUDA = 3,         /* CUDA runtime error (including: no device) */
    BFP_E_ALIGN = 4         /* pointer not aligned to the element size */
};

int bfp_version(void);
const char* bfp_last_error(void);

/* Number of kernels this library has launched since load (all threads); bench.py's gpu_launches. */
uint64_t bfp_launch_count(void);

/* Runtime knobs (also read once from the environment: BFP_STREAM_CTAS_PER_SM, BFP_FORCE_GENERIC, BFP_HOST_CHUNK_MB):
 *   "stream_ctas_per_sm"  resident CTAs per SM the streaming quantiser sizes its grid for (default 8)
 *   "force_generic"       1 = route every call through the generic (ragged-shape) kernel; for tests
 *   "host_chunk_bytes"    largest pipelined chunk of bfp_quantize_host, input bytes (default 16 MiB)
 *   "host_chunk_min_bytes" smallest chunk of its tapered schedule: chunks double from here at the start and halve towards
 *                         the end, so pipeline fill and drain cost one small chunk each (default 1 MiB)
 *   "quant_tma"           1 = TMA-staged variant of the streaming quantiser (cp.async.bulk -> smem ring); slower than the
 *                         default direct 128-bit loads on B200, kept for comparison
 *   "gemm_out_tma"        1 (default) = GEMM epilogues write through smem + TMA stores; 0 = plain st.global
 *   "gemm_sp_tile"        0 = sparse GEMM picks its 256- or 480-token pair tile by cost; 256 / 480 forces one
 *   "gemm_bf16_cta_group" 0 = dense bf16 GEMM uses CTA pairs when T > 128 and N > 128; 1 / 2 forces the mode
 *   "pdl"                 1 (default) = the streaming kernels are launched with programmatic stream serialization
 *   "gemm_sp_cta_group"   0 = bfp_gemm_bf16_sp uses CTA pairs (cta_group::2) when N > 128; 1 / 2 forces the mode
 *   "gemm_bf16_tile_n"    0 = bfp_gemm_bf16 uses its 128x256 tile (128x128 when N <= 128); 128 / 256 forces one
 *   "unstructured_force_fallback" 1 = bfp_unstructured_quantize takes its whole-tensor radix select (the path for a missed
 *                         bracket or massive ties on a non-round value); for tests */
int bfp_set_option(const char* name, int64_t value);

/* sm count, compute capability, L2 bytes of the current device. */
int bfp_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* l2_bytes);

/*
 * float_to_bfp_blocked (bfp_ops.py:124-149) for sparsity_num_format in {'bfp','fp32'} with
 * sparsity_mode 'structured' -- ONE fused kernel: N:M magnitude mask (bfp_ops.py:73-91), shared block exponent
 * ceil(log2(max|t|+eps)) (:29-33), mantissa rounding (:20-27) and clamp (:35-44), in the order `order`.
 *
 *   in_dtype   BFP_DT_*; out_dtype must equal in_dtype for BFP_ROUND_NEAREST and BFP_DT_F32 for
 *              BFP_ROUND_STOCHASTIC (the reference's type promotion, bfp_ops.py:22-23).
 *   block_size > 0, mant_bits in [0, 23] (number of magnitude bits m: HBFP8 -> 7).
 *   rounding   BFP_ROUND_*; stochastic uniforms come from Philox4x32-10 keyed by `seed`, counter =
 *              (flat element index / 4, offset): u = (word >> 8) * 2^-24.
 *   N, M       keep N of every M along K (ignored for BFP_ORDER_QUANT_ONLY); 0 < N <= M <= 64.
 *   out may alias in only when out_dtype == in_dtype and order == BFP_ORDER_QUANT_ONLY.
 */
int bfp_quantize(const void* in, void* out, int64_t rows, int64_t K, int in_dtype, int out_dtype, int block_size,
                 int mant_bits, float eps, int rounding, uint64_t seed, uint64_t offset, int N, int M, int order,
                 int tie_rule, void* stream);

/* _structured_N_M_sparsity (bfp_ops.py:73-91) alone: = bfp_quantize(..., BFP_ORDER_SPARSIFY_ONLY). */
int bfp_nm_sparsify(const void* in, void* out, int64_t rows, int64_t K, int dtype, int N, int M, int tie_rule,
                    void* stream);

/* _unstructured_sparsity (bfp_ops.py:61-71): global magnitude pruning of the whole tensor viewed as one row.  Zeroes the
 * k entries torch.topk(|t|, k, largest=False) returns on torch-CUDA: every entry strictly below the k-th smallest
 * magnitude, and of the entries equal to it the first ones in index order; NaN counts as largest.
 * k = int(numel * sparsity_frac) is computed by the caller (bfp_ops.py:66).  `workspace` is caller-owned device scratch of
 * bfp_unstructured_workspace_bytes() bytes (no allocation inside).  out may alias in. */
size_t bfp_unstructured_workspace_bytes(void);
int bfp_unstructured_sparsify(const void* in, void* out, int64_t numel, int dtype, uint64_t k, void* workspace,
                              void* stream);

/* float_to_bfp_blocked (bfp_ops.py:124-149) with sparsity_mode == 'unstructured' and sparsity_num_format in {'bfp','fp32'}:
 * global magnitude pruning (_unstructured_sparsity, :61-71) and the BFP quantiser (_no_sparsity_float_to_bfp, :46-59) in the
 * order `order` -- BFP_ORDER_SPARSIFY_QUANT (first == 's'), BFP_ORDER_QUANT_SPARSIFY (the k smallest magnitudes of the QUANTISED
 * tensor go) or BFP_ORDER_SPARSIFY_ONLY -- in two reads and one write of the tensor: a sampled bracket of the k-th magnitude, one
 * counting pass, one masking + quantising pass (csrc/bfp_unstructured_fused.cu).  Same results as composing
 * bfp_unstructured_sparsify and bfp_quantize (same Philox counters for stochastic rounding).  k in (0, numel).
 * Needs 16-byte aligned buffers, numel a multiple of 4 (fp32) / 8 (half) and, when quantising, K a multiple of a power-of-two
 * block_size of 4..128 (fp32) / 8..256 (half) elements: BFP_E_UNSUPPORTED otherwise (compose the two calls).  out must not alias
 * in.  `workspace`: bfp_unstructured_quantize_workspace_bytes(numel, in_dtype) bytes of caller-owned device scratch
 * (numel bytes + 17 MB), 16-byte aligned; numel < 2^32; no allocation and no host synchronisation inside. */
size_t bfp_unstructured_quantize_workspace_bytes(int64_t numel, int dtype);
int bfp_unstructured_quantize(const void* in, void* out, int64_t rows, int64_t K, int in_dtype, int out_dtype, uint64_t k,
                              int order, int block_size, int mant_bits, float eps, int rounding, uint64_t seed, uint64_t offset,
                              void* workspace, size_t workspace_bytes, void* stream);

/* The 'int' number format (_quantize with sparsity_num_format == 'int', bfp_ops.py:111-120): SparseGPT's per-channel
 * symmetric min/max INT-`bits` fake quantiser (int_ops.py Quantizer.configure/find_params/quantize, perchannel, sym).
 * The tensor is viewed as [A, C, inner]; element i belongs to channel (i / inner) % C:
 *   weight [C, ...]               : A = 1, inner = numel / C           (int_ops.py:40-43)
 *   activation [.., C] (2-D, 3-D) : A = numel / C, inner = 1           (int_ops.py:47-50)
 *   activation [N, C, H, W]       : A = N, inner = H * W               (int_ops.py:44-46)
 * Output is fp32 for every input dtype (the reference promotes).  workspace: bfp_int_workspace_bytes(C) bytes of
 * caller-owned device scratch. */
size_t bfp_int_workspace_bytes(int64_t C);
int bfp_int_quantize(const void* in, float* out, int64_t A, int64_t C, int64_t inner, int dtype, int bits,
                     void* workspace, void* stream);

/* Activations [A, C] (channel = last dim, int_ops.py:47-50) quantised like bfp_int_quantize and written as three bf16
 * planes whose sum is the fp32 fake-quantised value: the operands of bf16 tensor-core GEMMs against the weight's integer
 * grid (BFPLinear with sparsity_num_format == 'int').  out_bf16 holds 3 A C elements:
 *   hi     column segments of width kseg, segment-major: segment s is a contiguous [A, min(kseg, C - s kseg)] matrix at
 *          element offset s * A * kseg;
 *   mid|lo one [A, 2C] matrix at element offset A * C.
 * The contraction is chunked accordingly (mid|lo first, then one bfp_gemm_bf16_acc per hi segment).  C and kseg must be
 * multiples of 8; workspace as for bfp_int_quantize, 16-byte aligned. */
int bfp_int_quantize_split3(const void* in, void* out_bf16, int64_t A, int64_t C, int64_t kseg, int dtype, int bits,
                            void* workspace, void* stream);

/* 2-D weights [C, K] with N:4 structured sparsity and the 'int' format in ONE pass (float_to_bfp_blocked with
 * sparsity_num_format == 'int', sparsity_mode == 'structured', bfp_ops.py:143-149): order BFP_ORDER_SPARSIFY_QUANT =
 * _quantize(_sparsify(t)), BFP_ORDER_QUANT_SPARSIFY = _sparsify(_quantize(t)).  torch-CUDA tie rule.  Needs M == 4,
 * 0 < N < 4, 16-byte aligned buffers, K a multiple of 4 (fp32) / 8 (half) and K <= 16384 (fp32) / 32768 (half);
 * BFP_E_UNSUPPORTED otherwise (compose bfp_nm_sparsify and bfp_int_quantize). */
int bfp_int_quantize_nm(const void* in, float* out, int64_t C, int64_t K, int dtype, int bits, int N, int M, int order,
                        void* stream);

/* get_exponent (bfp_ops.py:29-33): exp_out[rows, ceil(K/block_size)] fp32, in the arithmetic of `dtype`. */
int bfp_block_exponent(const void* in, float* exp_out, int64_t rows, int64_t K, int dtype, int block_size, float eps,
                       void* stream);

/*
 * The same operator through HOST buffers (what a caller holding CPU tensors sees): the tensor is split into row
 * chunks; chunk i+1's host->device copy, chunk i's kernel and chunk i-1's device->host copy overlap on three
 * streams.  host_in / host_out should be page-locked for full PCIe speed (pageable memory works, slower).
 * Returns after host_out is complete.
 */
int bfp_quantize_host(const void* host_in, void* host_out, int64_t rows, int64_t K, int in_dtype, int out_dtype,
                      int block_size, int mant_bits, float eps, int rounding, uint64_t seed, uint64_t offset, int N,
                      int M, int order, int tie_rule);

/* Frees the staging buffers bfp_quantize_host keeps between calls. */
int bfp_host_staging_release(void);

/* ------------------------------------------------------------------------------------------------------------------
 * Packed BFP operands and the tensor-core BFP linear (replaces the F.linear on dequantised tensors of
 * bfp_ops.py:187-190 / BFPLinear.forward :278-287).
 *
 * Packed layout of a [rows, K] tensor quantised with (block_size B, mant_bits m <= 7):
 *   mant    int8 [rows, Kp]            Kp = K rounded up to 16; integer mantissas q, |q| <= 2^m - 1; columns >= K zero
 *   scale_t fp32 [nkb_pad, rows_pad]   block-major: scale_t[kb][row] = 2^(e - m) (the block's interval); value = q*scale
 *                                      nkb_pad = ceil(Kp / 128) * (128 / B); rows_pad = rows rounded up to 256;
 *                                      padding entries must be 0.  NaN marks a block the packed form cannot represent
 *                                      (the reference yields NaN or leaves the normal range there).
 * bfp_packed_layout fills the three sizes.  mant / scale_t must be 16-byte aligned; when Kp != K or padding exists the
 * caller zero-fills the buffers before packing.
 */
int bfp_packed_layout(int64_t rows, int64_t K, int block_size, int64_t* Kp, int64_t* rows_pad, int64_t* nkb_pad);

/* float_to_bfp_blocked (bfp_ops.py:124-149) straight into the packed form: same arguments as bfp_quantize (orders
 * QUANT_ONLY / SPARSIFY_QUANT / QUANT_SPARSIFY; N:M ties follow BFP_TIE_TORCH_CUDA).
 * Contract: bfp_unpack(bfp_quantize_pack(x)) == bfp_quantize(x) bit for bit, except that -0.0 unpacks as +0.0. */
int bfp_quantize_pack(const void* in, int8_t* mant, float* scale_t, int64_t rows, int64_t K, int in_dtype,
                      int block_size, int mant_bits, float eps, int rounding, uint64_t seed, uint64_t offset, int N,
                      int M, int order, void* stream);

/* packed -> fp32 [rows, K] */
int bfp_unpack(const int8_t* mant, const float* scale_t, float* out, int64_t rows, int64_t K, int block_size,
               void* stream);

/* out[T, N] (fp32) = A_packed[T, K] . B_packed[N, K]^T + bias[N]:
 *   out[t,n] = bias[n] + sum_kb a_scale_t[kb][t] * b_scale_t[kb][n] * sum_{k in kb} a_mant[t,k] * b_mant[n,k]
 * tcgen05.mma.kind::i8 with int32 TMEM accumulators per BFP block, fp32 rescale + accumulation in registers.
 * Both operands must be packed with the same block_size (32, 64 or 128) and the layout above; bias may be NULL. */
int bfp_gemm_i8(const int8_t* a_mant, const float* a_scale_t, const int8_t* b_mant, const float* b_scale_t,
                const float* bias, float* out, int64_t T, int64_t N, int64_t K, int block_size, void* stream);

/* Exact bf16 variant of the BFP linear.  For mant_bits <= 8 the dequantised value q * 2^(e-m) is exactly representable in
 * bf16, so the operands carry their block scales in their exponents and the contraction needs no per-block rescale:
 *   bfp_quantize_pack_bf16: float_to_bfp_blocked straight to bf16 [rows, Kp], Kp = K rounded up to 8 (caller zero-fills
 *                           the buffer when Kp != K); unrepresentable blocks become NaN.  Any block_size (also 16).
 *   bfp_gemm_bf16:          out[T,N] (fp32) = A[T,K] . B[N,K]^T + bias, tcgen05.mma.kind::f16 with fp32 TMEM accumulators.
 * Same result contract as bfp_gemm_i8 (exact products, fp32 accumulation order differs). */
int bfp_quantize_pack_bf16(const void* in, void* out_bf16, int64_t rows, int64_t K, int in_dtype, int block_size,
                           int mant_bits, float eps, int rounding, uint64_t seed, uint64_t offset, int N, int M,
                           int order, void* stream);
int bfp_gemm_bf16(const void* a_bf16, const void* b_bf16, const float* bias, float* out, int64_t T, int64_t N, int64_t K,
                  void* stream);
/* ... with the output dtype chosen: BFP_DT_F32, or BFP_DT_F16 / BFP_DT_BF16 (accumulator + bias rounded once in the epilogue). */
int bfp_gemm_bf16_ex(const void* a_bf16, const void* b_bf16, const float* bias, void* out, int out_dtype, int64_t T, int64_t N,
                     int64_t K, void* stream);
/* Transposed copy of a 16-bit matrix with zero padding: out[c][r] = in[r][c] for r < rows, c < cols; out columns
 * rows .. ld_out-1 are zero.  in has row stride ld_in (>= cols), out [cols, ld_out] with ld_out even and >= rows.  The
 * backward contractions of the BFP linear (bfp_ops.py:168-185: dgrad over N, wgrad over T) take their operands from the
 * packed bf16 tensors through this. */
int bfp_transpose_pad_16(const void* in, void* out, int64_t rows, int64_t cols, int64_t ld_in, int64_t ld_out, void* stream);

/* `batch` independent products out[b] = A[b] . B[b]^T in ONE launch (F_matmul_bfp on [..., M, K] x [..., K, N] operands,
 * bfp_ops.py:240-245: GPT-2 style attention matmuls, modeling_gpt2.py:205-207): A [batch, T, K], B [batch, N, K] bf16
 * contiguous (K a multiple of 8), out [batch, T, N] of out_dtype.  The operands are read as stacked rows through 2-D tensor
 * maps and the output is written through a 3-D map that clips each entry's rows, so T and N need not be tile multiples.
 * Needs N * sizeof(out element) % 16 == 0 and a 16-byte aligned `out`; BFP_E_UNSUPPORTED otherwise. */
int bfp_gemm_bf16_batched(const void* a_bf16, const void* b_bf16, void* out, int out_dtype, int64_t batch, int64_t T, int64_t N,
                          int64_t K, void* stream);

/* out[T,N] (fp32) += A . B^T: the K-chunked contraction.  The tensor cores truncate (round toward zero) each time a group
 * of products joins the fp32 accumulator, so a long contraction whose partial sums are not exactly representable drifts by
 * about 2^-25 per MMA step (measured: 1.7e-5 relative at K = 12288).  BFP operands normally sum exactly; operands that do
 * not (the 'int' format's three-plane activations, bfp_int_quantize_split3) are contracted in chunks of K: the first chunk
 * through bfp_gemm_bf16 / bfp_gemm_bf16_sp, the rest through these, whose epilogue adds the tile into `out` with TMA
 * reduce-add (one correctly rounded fp32 add per chunk).  Needs N % 4 == 0 and a 16-byte aligned `out`. */
int bfp_gemm_bf16_acc(const void* a_bf16, const void* b_bf16, float* out, int64_t T, int64_t N, int64_t K, void* stream);
int bfp_gemm_bf16_sp_acc(const void* x_bf16, const void* w_comp, const void* w_meta, float* out, int64_t T, int64_t N, int64_t K,
                         void* stream);

/* Block-scaled FP8-class variant of the BFP linear for narrow mantissas (HBFP4 / HBFP5: mant_bits <= 4, block_size a multiple
 * of 32): tcgen05.mma.kind::mxf8f6f4.block_scale, twice the tensor-core rate of the bf16 kind.  The integer mantissa q (|q| <= 16)
 * is exact in E4M3 and the block scale 2^(e-m) is exactly a UE8M0 byte, which the hardware applies per 32 elements of K before
 * the fp32 accumulation: same result contract as bfp_gemm_bf16 (exact products, fp32 accumulation order differs).  Replaces the
 * F.linear on fake-quantised tensors of bfp_ops.py:187-190.
 *   bfp_mx_layout:      Kp = K rounded up to 128; sf_bytes of the scale array for rows grouped in tiles of tile_rows (128 for
 *                       the activation operand; the GEMM's N tile -- 128, 240 or 256 -- for the weight).
 *   bfp_mx_from_packed: int8 mantissas + fp32 block-major scales (bfp_quantize_pack output, same rows / K / block_size) ->
 *                       vals uint8 [rows, Kp] (E4M3 bytes, every byte written) + sf (one 512-byte atom per 128 rows x 128 k in the
 *                       order tcgen05.cp.32x128b.warpx4 moves it to TMEM, csrc/bfp_gemm_mx.cu).
 *                       fold = 0: the general form -- integer mantissas, one scale per 32 elements (block_size % 32 == 0).
 *                       fold = 1: the weight form -- the block exponents are folded into the E4M3 values relative to one reference
 *                       exponent per row (row_ref: rows int32 of scratch), so the row has a single scale, the GEMM copies the B
 *                       scales once per tile instead of once per K slab (a copy costs ~50 clk of the tensor pipe), and any
 *                       block_size works.  Exact while every block exponent of a row is within ten octaves of its largest.
 *                       *violations (device uint32, caller-zeroed) counts what the form cannot hold (mantissas beyond +-16; for
 *                       fold = 1 rows whose exponent spread is too wide or that hold NaN-marked blocks): then use the other form
 *                       or bfp_gemm_bf16.
 *   bfp_gemm_mx:        out[T,N] (fp32) = A . B^T + bias.  A in the general form (tile_rows 128), B in either (b_folded).  K is
 *                       the logical K of both operands; N % 4 == 0; 16-byte aligned buffers. */
int bfp_mx_layout(int64_t rows, int64_t K, int tile_rows, int fold, int64_t* Kp, int64_t* sf_bytes);
int bfp_mx_from_packed(const int8_t* mant, const float* scale_t, int64_t rows, int64_t K, int block_size, int tile_rows, int fold,
                       void* vals, void* sf, int32_t* row_ref, uint32_t* violations, void* stream);
/* float_to_bfp_blocked (quantise only, nearest rounding; bfp_ops.py:46-59) of an activation tensor straight into the general mx form in
 * one pass (vals [rows, K] + scale atoms of 128-row tiles).  Needs mant_bits in [1, 4], block_size 32 / 64 / 128, K a multiple of 128
 * (fp32) / 256 (half), 16-byte aligned buffers: BFP_E_UNSUPPORTED otherwise (bfp_quantize_pack + bfp_mx_from_packed).  The caller
 * zero-fills sf once when rows % 128 != 0 (the rows of the last tile that do not exist are never written). */
int bfp_quantize_pack_mx(const void* in, void* vals, void* sf, int64_t rows, int64_t K, int in_dtype, int block_size, int mant_bits,
                         float eps, void* stream);
int bfp_gemm_mx(const void* a_vals, const void* a_sf, const void* b_vals, const void* b_sf, int b_tile_rows, int b_folded,
                const float* bias, float* out, int64_t T, int64_t N, int64_t K, void* stream);

/* ---- MX (OCP Microscaling) formats: the arithmetic behind the reference's mx_layers.py -----------------------------------------
 * /root/reference/src/transformers/bfp/mx_layers.py:23-109 wraps mx.Linear / mx.Conv2d / mx.matmul of microsoft/microxcaling (not
 * vendored, not pinned: parity unpinned -- restated from the library's published emulation and the OCP MX v1.0 specification,
 * oracle/mx_oracle.py).  Element formats: formats.py:24-33 (the enum values below are the reference's ElemFormat values),
 * parameters formats.py:86-123.  Every rounding is the library's 'nearest' = half away from zero.
 *   bfp_bfloat_round:     quantize_elemwise_op for bfloatX (mx/elemwise_ops.py _quantize_bfloat): out = rb(in), or with a bias row
 *                         (fp32 [ncols], n % ncols == 0) the bias step of mx/linear.py: out = rb(rb(in) + rb(bias[col])).
 *   bfp_ocp_mx_quantize:  rb (bfloat, 0 = none) then quantize_mx_op along the last dim: zero-padded blocks of block_size (0 = the whole
 *                         row), shared scale 2^se with se = clamp(floor(log2 max|x|) - emax_elem, +-(2^(scale_bits-1) - 1)), elements
 *                         rounded to elem_format with saturation.  out_kind 0: fake-quantised tensor in the input dtype [rows, K];
 *                         out_kind 1: exact bf16 operand [rows, ld_out] for bfp_gemm_bf16 / bfp_compress_2to4_bf16 (every MX value
 *                         has <= 8 significant bits; columns K .. ld_out-1 are the caller's zero fill).
 *   bfp_ocp_mx_pack:      the same quantiser straight into the operand form of tcgen05.mma.kind::mxf8f6f4.block_scale (bfp_gemm_mx):
 *                         vals = E4M3 byte of every element, sf = the blocks' UE8M0 scale bytes in atoms of tile_rows-row tiles
 *                         (bfp_mx_layout, fold = 0).  Formats whose values are E4M3 numbers (fp8_e4m3, fp6_*, fp4_e2m1, int4, int2),
 *                         block_size 32 / 64 / 128, K % 128 == 0 (fp32) / 256 (half), scale_bits 8; else BFP_E_UNSUPPORTED (use
 *                         out_kind 1).  A NaN / Inf block gets the NaN scale 0xff, as the emulation makes the whole block NaN.
 *   bfp_gemm_mx_round:    bfp_gemm_mx whose epilogue applies the output steps of mx/linear.py: out = rb(acc), then with a bias
 *                         out = rb(out + rb(bias)) (bfloat 0 = plain bfp_gemm_mx). */
enum { BFP_MX_INT8 = 1, BFP_MX_INT4 = 2, BFP_MX_INT2 = 3, BFP_MX_FP8_E5M2 = 4, BFP_MX_FP8_E4M3 = 5, BFP_MX_FP6_E3M2 = 6,
       BFP_MX_FP6_E2M3 = 7, BFP_MX_FP4_E2M1 = 8 };
int bfp_bfloat_round(const void* in, void* out, const float* bias, int64_t n, int64_t ncols, int dtype, int bfloat, void* stream);
int bfp_ocp_mx_quantize(const void* in, void* out, int64_t rows, int64_t K, int in_dtype, int out_kind, int64_t ld_out, int block_size,
                        int elem_format, int scale_bits, int bfloat, int flush_fp32_subnorms, void* stream);
int bfp_ocp_mx_pack(const void* in, void* vals, void* sf, int64_t rows, int64_t K, int in_dtype, int tile_rows, int block_size,
                    int elem_format, int scale_bits, int bfloat, int flush_fp32_subnorms, void* stream);
int bfp_gemm_mx_round(const void* a_vals, const void* a_sf, const void* b_vals, const void* b_sf, int b_tile_rows, int b_folded,
                      const float* bias, float* out, int64_t T, int64_t N, int64_t K, int bfloat, void* stream);

/* 2:4 structured-sparse variant of the BFP linear for weights pruned by _structured_N_M_sparsity with N=2, M=4
 * (bfp_ops.py:73-91; the reference then multiplies the zero-filled dense tensor, bfp_ops.py:187-190).  The pruned
 * exact-bf16 weight is stored compressed and the tensor core skips the zeros (tcgen05.mma.sp.kind::f16: 32 logical k per
 * MMA instead of 16):
 *   bfp_sp_layout:          sizes of the compressed form of a [rows, K] operand: Kc = kept bf16 per row (K padded to 128,
 *                           halved), meta_bytes = ceil(rows/128) * ceil(K/128) * 2048.
 *   bfp_compress_2to4_bf16: w_bf16 [rows, Kp] (bfp_quantize_pack_bf16 output, Kp = K rounded up to 8) -> w_comp bf16
 *                           [rows, Kc] + w_meta (4-bit index nibbles in the order tcgen05.cp moves them to TMEM, see
 *                           csrc/bfp_gemm_sp.cu).  *violations (device uint32, caller-zeroed) counts groups of four with
 *                           more than two non-zeros, i.e. input that is not 2:4; the result is then not equivalent.
 *   bfp_gemm_bf16_sp:       out[T,N] (fp32) = x[T,K] . W[N,K]^T + bias from the compressed W.  Same result contract as
 *                           bfp_gemm_bf16 on the uncompressed operand (exact products, fp32 accumulation order differs). */
int bfp_sp_layout(int64_t rows, int64_t K, int64_t* Kc, int64_t* meta_bytes);
int bfp_compress_2to4_bf16(const void* w_bf16, int64_t rows, int64_t K, void* w_comp, void* w_meta, uint32_t* violations,
                           void* stream);
int bfp_gemm_bf16_sp(const void* x_bf16, const void* w_comp, const void* w_meta, const float* bias, float* out, int64_t T,
                     int64_t N, int64_t K, void* stream);

/* Column-parallel BFP linear with the all-gather fused into the GEMM epilogue (SURVEY.md section 8 e: the 65B-class linear
 * split by output rows over G GPUs).  Same contraction as bfp_gemm_bf16_sp on this rank's weight shard W[N, K]; the
 * result tile is written n_out times: out_slices[g] points at this rank's column slice inside GPU g's full
 * [T, ld_out] fp32 output (its own memory or a peer's mapped over NVLink, e.g. torch symmetric memory), row stride ld_out
 * elements of out_dtype (BFP_DT_F32 / F16 / BF16).  Every slice pointer and the row stride in bytes must be multiples of 16.  The caller orders the kernel against the
 * peers with a barrier before (buffers free) and after (stores visible) -- no NCCL call, no transposing copy. */
int bfp_gemm_bf16_sp_gather(const void* x_bf16, const void* w_comp, const void* w_meta, const float* bias,
                            void* const* out_slices, int n_out, int out_dtype, int64_t ld_out, int64_t T, int64_t N,
                            int64_t K, void* stream);

/* bfp_gemm_bf16_sp with an explicit output dtype and row stride: out_dtype BFP_DT_F32, or BFP_DT_F16 / BFP_DT_BF16 for half
 * precision modules -- the fp32 accumulator (+ bias) is rounded to the dtype once in the epilogue, which is what the library
 * HGEMM the reference calls on fake-quantised fp16 tensors does, and halves the output traffic.  ld_out in elements. */
int bfp_gemm_bf16_sp_ex(const void* x_bf16, const void* w_comp, const void* w_meta, const float* bias, void* out,
                        int out_dtype, int64_t ld_out, int64_t T, int64_t N, int64_t K, void* stream);

/* The 256-entry table behind BFP_TIE_TORCH_CPU for 2:4 (index = c0 + 4*c1 + 16*c2 + 64*c3 with
 * c_i = #{j : |v_j| < |v_i|}; value = 4-bit drop mask, 0xff = unreachable).  Exposed for the tests. */
int bfp_debug_cpu_tie_lut(uint8_t out[256]);

/* The table behind the half-precision block exponent (csrc/bfp_common.cuh): for s = 2^k (1 + f 2^-mb) in fp16 (mb = 10) or
 * bf16 (mb = 7), ceil(log2(s) rounded to the dtype) = k + (f + 1 >= out[k + 128]); 0 = not tabulated (the kernels then evaluate
 * the formula).  Exposed for the tests, which compare it with torch's own log2 on every (k, f). */
int bfp_debug_exp_table(int dtype, uint16_t out[256]);

#ifdef __cplusplus
}
#endif
#endif /* BFP_B200_H_ */
