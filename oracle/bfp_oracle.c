/*
 * oracle/bfp_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the block-floating-point + N:M hot path of the
 * reference, src/transformers/bfp/bfp_ops.py.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product path (quantization-sparsity-interplay_b200/) never does.
 *
 * Parity status: PINNED against the reference itself.  The reference holds no
 * golden vectors for this path, so tests/golden/make_golden.py
 * imports the reference's bfp_ops.py in the build container, records its
 * outputs as fixtures, and tests/test_oracle.py checks this file against them
 * (and live against the reference when /root/reference is present).
 *
 * Each function cites the reference lines it follows (paths relative to
 * /root/reference/src/transformers/bfp/).
 *
 * Arithmetic model: the reference runs torch elementwise ops in the tensor's
 * dtype.  For fp16/bf16 tensors torch computes each op in fp32 and rounds the
 * result to the tensor dtype (round-to-nearest-even), so every intermediate
 * below is passed through rnd(dtype, .).  dtype codes: 0 = fp32, 1 = fp16,
 * 2 = bf16.  Half/bf16 tensors travel as uint16_t bit patterns.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

enum { DT_F32 = 0, DT_F16 = 1, DT_BF16 = 2 };

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* ---- software fp16 / bf16 <-> fp32 (round-to-nearest-even) ------------- */
static inline float bf16_to_f32(uint16_t h) { return u2f((uint32_t)h << 16); }
static inline uint16_t f32_to_bf16(float f) {
    uint32_t u = f2u(f);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x0040u); /* quiet NaN */
    uint32_t lsb = (u >> 16) & 1u;
    u += 0x7fffu + lsb;
    return (uint16_t)(u >> 16);
}
static inline float f16_to_f32(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1fu, man = h & 0x3ffu;
    if (exp == 0) {
        if (man == 0) return u2f(sign);
        float v = (float)man * 5.9604644775390625e-08f; /* man * 2^-24 */
        return u2f(f2u(v) | sign);
    }
    if (exp == 31) return u2f(sign | 0x7f800000u | (man << 13));
    return u2f(sign | ((exp + 112u) << 23) | (man << 13));
}
static inline uint16_t f32_to_f16(float f) {
    uint32_t u = f2u(f), sign = (u >> 16) & 0x8000u, a = u & 0x7fffffffu;
    if (a > 0x7f800000u) return (uint16_t)(sign | 0x7e00u);          /* NaN */
    if (a >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);         /* >= 65520 -> inf */
    if (a < 0x33000001u) return (uint16_t)sign;                      /* <= 2^-25 -> 0 (tie to even) */
    if (a < 0x38800000u) {                                           /* subnormal result */
        float v = u2f(a) * 16777216.0f;                              /* / 2^-24, exact */
        float r = rintf(v);                                          /* RNE under default mode */
        return (uint16_t)(sign | (uint32_t)r);
    }
    uint32_t lsb = (a >> 13) & 1u;
    a += 0xfffu + lsb;
    return (uint16_t)(sign | ((a - 0x38000000u) >> 13));
}
static inline float rnd(int dt, float x) {
    if (dt == DT_F16) return f16_to_f32(f32_to_f16(x));
    if (dt == DT_BF16) return bf16_to_f32(f32_to_bf16(x));
    return x;
}
static inline float load_elt(const void *p, int dt, int64_t i) {
    if (dt == DT_F32) return ((const float *)p)[i];
    if (dt == DT_F16) return f16_to_f32(((const uint16_t *)p)[i]);
    return bf16_to_f32(((const uint16_t *)p)[i]);
}
static inline void store_elt(void *p, int dt, int64_t i, float v) {
    if (dt == DT_F32) ((float *)p)[i] = v;
    else if (dt == DT_F16) ((uint16_t *)p)[i] = f32_to_f16(v);
    else ((uint16_t *)p)[i] = f32_to_bf16(v);
}

static inline uint32_t order_key(float absv);

/* torch.max / torch.min (binary, elementwise): NaN in either operand propagates. */
static inline float t_max(float a, float b) { if (a != a) return a; if (b != b) return b; return a > b ? a : b; }
static inline float t_min(float a, float b) { if (a != a) return a; if (b != b) return b; return a < b ? a : b; }

/* ---- Philox4x32-10 (Salmon et al., SC'11): the counter-based generator the
 * CUDA kernel uses for stochastic rounding.  Restated here so the test can
 * feed the oracle the exact uniforms the kernel drew.
 * counter = (elt_index/4 lo, elt_index/4 hi, offset lo, offset hi), key = seed. */
static inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
/* 32-bit word of the stream for flat element index i: word (i & 3) of the block with counter (i >> 2, offset), key = seed */
uint32_t oracle_philox_word(uint64_t seed, uint64_t offset, uint64_t i) {
    uint32_t c[4] = { (uint32_t)(i >> 2), (uint32_t)((i >> 2) >> 32), (uint32_t)offset, (uint32_t)(offset >> 32) };
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    return c[i & 3];
}
/* uniform in [0,1): the word truncated (round toward zero) to 24 significant bits, times 2^-32 -- what the kernel's
 * cvt.rz.f32.u32 followed by an exact power-of-two scale gives.  Words >= 2^31 yield (w >> 8) * 2^-24. */
float oracle_philox_uniform(uint64_t seed, uint64_t offset, uint64_t i) {
    uint32_t w = oracle_philox_word(seed, offset, i);
    if (w) {
        int drop = 32 - __builtin_clz(w) - 24;
        if (drop > 0) w &= ~((1u << drop) - 1u);
    }
    return (float)w * 2.3283064365386963e-10f;            /* exact: <= 24 significant bits, power-of-two scale */
}
void oracle_philox_fill(float *u, int64_t n, uint64_t seed, uint64_t offset) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) u[i] = oracle_philox_uniform(seed, offset, (uint64_t)i);
}

/* ---- BFP quantiser ------------------------------------------------------
 * bfp_ops.py:29-33  get_exponent:   ceil(log2(max|t| + eps)) per block
 * bfp_ops.py:35-44  _convert_blocked_float_to_bfp
 * bfp_ops.py:20-27  round_tensor ('determ': round-half-even; 'stoc': rint(u-0.5+x), u fp32)
 * bfp_ops.py:46-59  _no_sparsity_float_to_bfp: right-pad last dim with zeros to a
 *                   multiple of block_size, blocks never straddle rows, narrow back.
 *
 * in/out: [rows, K] contiguous.  rounding: 0 = 'determ', 1 = 'stoc'.
 * For 'stoc' the caller supplies the uniforms rand_u[rows*K] (fp32, [0,1));
 * the output dtype is then fp32 for every input dtype (torch type promotion:
 * the fp32 random tensor is the in-place destination, bfp_ops.py:22-23,42-44).
 * out_dt is therefore in_dt for determ and DT_F32 for stoc.
 */
static void quant_block(const float *t, float *y, int n, int B, int dt, int m, float eps, int stoc, const float *u) {
    /* n = real elements in this block (<= B); the rest are zero padding. */
    float a = 0.0f;
    int nan_seen = 0;
    for (int i = 0; i < n; ++i) {
        float v = fabsf(t[i]);
        if (v != v) nan_seen = 1;
        if (v > a) a = v;
    }
    if (nan_seen) a = NAN;                                   /* torch.max propagates NaN */
    (void)B;
    float s = rnd(dt, a + eps);                              /* :33  max_v + epsilon   */
    float e = ceilf(rnd(dt, log2f(s)));                      /* :33  .log2().ceil()    */
    float p = rnd(dt, e - (float)m);                         /* :38  exp - mant_bits   */
    float delta = rnd(dt, powf(2.0f, p));                    /* :38  interval          */
    float vmax = rnd(dt, rnd(dt, powf(2.0f, e)) - delta);    /* :39  max_v             */
    for (int i = 0; i < n; ++i) {
        float x = rnd(dt, t[i] / delta);                     /* :40  t / interval      */
        float r, yy;
        if (stoc) {
            float smp = u[i] - 0.5f;                         /* :22  rand - 0.5 (fp32) */
            r = rintf(smp + x);                              /* :23  add_(t).round()   */
            yy = r * delta;                                  /* :42  fp32 *= interval  */
        } else {
            r = rintf(x);                                    /* :25  t.round()         */
            yy = rnd(dt, r * delta);                         /* :42                    */
        }
        y[i] = t_min(t_max(yy, -vmax), vmax);                /* :44                    */
    }
}

int oracle_bfp_quantize(const void *in, void *out, int64_t rows, int64_t K, int dt, int B, int m, float eps,
                        int stoc, const float *rand_u) {
    if (B <= 0 || rows < 0 || K < 0) return 1;
    const int64_t nblk = (K + B - 1) / B;
    const int out_dt = stoc ? DT_F32 : dt;
#pragma omp parallel
    {
        float *tb = (float *)malloc(sizeof(float) * (size_t)B * 2);
        float *yb = tb + B;
#pragma omp for schedule(static)
        for (int64_t blk = 0; blk < rows * nblk; ++blk) {
            int64_t r = blk / nblk, kb = blk % nblk;
            int64_t base = r * K + kb * B;
            int n = (int)((kb * B + B <= K) ? B : (K - kb * B));
            for (int i = 0; i < n; ++i) tb[i] = load_elt(in, dt, base + i);
            quant_block(tb, yb, n, B, dt, m, eps, stoc, stoc ? rand_u + base : NULL);
            for (int i = 0; i < n; ++i) store_elt(out, out_dt, base + i, yb[i]);
        }
        free(tb);
    }
    return 0;
}

/* per-block exponent only (bfp_ops.py:29-33), as floats in the tensor dtype. */
int oracle_bfp_exponent(const void *in, float *exp_out, int64_t rows, int64_t K, int dt, int B, float eps) {
    if (B <= 0) return 1;
    const int64_t nblk = (K + B - 1) / B;
#pragma omp parallel for schedule(static)
    for (int64_t blk = 0; blk < rows * nblk; ++blk) {
        int64_t r = blk / nblk, kb = blk % nblk, base = r * K + kb * B;
        int n = (int)((kb * B + B <= K) ? B : (K - kb * B));
        float a = 0.0f; int nan_seen = 0;
        for (int i = 0; i < n; ++i) { float v = fabsf(load_elt(in, dt, base + i)); if (v != v) nan_seen = 1; if (v > a) a = v; }
        if (nan_seen) a = NAN;
        exp_out[blk] = ceilf(rnd(dt, log2f(rnd(dt, a + eps))));
    }
    return 0;
}

/* ---- N:M structured sparsity -------------------------------------------
 * bfp_ops.py:73-91 _structured_N_M_sparsity: right-pad the last dim with zeros
 * to a multiple of M, view(-1, M), topk(|t|, k=M-N, largest=False) picks the
 * dropped positions, where(mask==0, 0, t): dropped -> +0.0, kept bits intact.
 *
 * torch.topk tie-breaking is implementation-defined; tie_rule selects which
 * torch backend is restated:
 *   0 = TORCH_CUDA: radix-select gather = the k smallest by (|v|, index)
 *       (measured on B200, profiles/r01_probe_ref_gpu.json: equals the
 *       lowest-index rule for (M,k) in {(4,1),(4,2),(4,3),(8,4),(8,6),(2,1),(16,8)}).
 *   1 = TORCH_CPU: ATen TopKImpl.h topk_impl_loop: std::nth_element over
 *       (value,index) pairs with a value-only comparator, first k slots
 *       (implemented in nm_cpu_rule.cpp with the same libstdc++ algorithm).
 * NaN sorts as largest in both.
 */
extern void oracle_topk_smallest_cpu_rule(const float *absv, int M, int k, int *idx_out); /* nm_cpu_rule.cpp */

static inline uint32_t order_key(float absv) {     /* monotone key, NaN largest */
    uint32_t u = f2u(absv) & 0x7fffffffu;
    return u;                                       /* |v| bits: inf=0x7f800000 < NaN patterns */
}

int oracle_nm_sparsify(const void *in, void *out, int64_t rows, int64_t K, int dt, int N, int M, int tie_rule) {
    if (!(N > 0 && M > 0 && N <= M)) return 1;      /* bfp_ops.py:74 assert */
    if (M > 4096) return 2;
    const int64_t ngrp = (K + M - 1) / M;
    const int k = M - N;
#pragma omp parallel
    {
        float *v = (float *)malloc(sizeof(float) * (size_t)M * 2);
        float *av = v + M;
        int *idx = (int *)malloc(sizeof(int) * (size_t)M);
        unsigned char *drop = (unsigned char *)malloc((size_t)M);
#pragma omp for schedule(static)
        for (int64_t g = 0; g < rows * ngrp; ++g) {
            int64_t r = g / ngrp, gi = g % ngrp, base = r * K + gi * M;
            int n = (int)((gi * M + M <= K) ? M : (K - gi * M));
            for (int i = 0; i < M; ++i) { v[i] = (i < n) ? load_elt(in, dt, base + i) : 0.0f; av[i] = fabsf(v[i]); }
            memset(drop, 0, (size_t)M);
            if (k > 0) {
                if (tie_rule == 1) {
                    oracle_topk_smallest_cpu_rule(av, M, k, idx);
                    for (int j = 0; j < k; ++j) drop[idx[j]] = 1;
                } else {
                    for (int i = 0; i < M; ++i) {
                        int rank = 0; uint32_t ki = order_key(av[i]);
                        for (int j = 0; j < M; ++j) {
                            uint32_t kj = order_key(av[j]);
                            rank += (kj < ki) || (kj == ki && j < i);
                        }
                        drop[i] = rank < k;
                    }
                }
            }
            for (int i = 0; i < n; ++i) store_elt(out, dt, base + i, drop[i] ? 0.0f : v[i]);
        }
        free(v); free(idx); free(drop);
    }
    return 0;
}

/* ---- unstructured (global) magnitude sparsity ------------------------------------
 * bfp_ops.py:61-71 _unstructured_sparsity: view(1, -1), topk(|t|, k = int(numel*frac), largest=False), dropped -> +0.0.
 * torch-CUDA tie order (the only one restated: torch-CPU's nth_element order on a 16 M-element slice is not a contract
 * anyone can rely on): the k smallest by (|v|, index).  k is passed in (the caller evaluates int(numel*frac) in double
 * precision exactly as Python does).
 */
typedef struct { uint32_t key; int64_t idx; } kv_t;
static int kv_cmp(const void *a, const void *b) {
    const kv_t *x = (const kv_t *)a, *y = (const kv_t *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx);
}
int oracle_unstructured_sparsify(const void *in, void *out, int64_t n, int dt, int64_t k) {
    if (k < 0) return 1;
    if (k > n) k = n;
    kv_t *kv = (kv_t *)malloc(sizeof(kv_t) * (size_t)(n > 0 ? n : 1));
    /* torch's CUDA radix select maps every NaN to one all-ones key (TopKTypeConfig::convert): NaNs tie, index order decides */
    for (int64_t i = 0; i < n; ++i) { uint32_t key = order_key(fabsf(load_elt(in, dt, i))); kv[i].key = key > 0x7f800000u ? 0x7f800001u : key; kv[i].idx = i; }
    qsort(kv, (size_t)n, sizeof(kv_t), kv_cmp);
    for (int64_t i = 0; i < n; ++i) store_elt(out, dt, i, load_elt(in, dt, i));
    for (int64_t j = 0; j < k; ++j) store_elt(out, dt, kv[j].idx, 0.0f);
    free(kv);
    return 0;
}

/* ---- 'int' number format ------------------------------------------------------------
 * bfp_ops.py:111-120 -> int_ops.py Quantizer.configure(bits) :18-31 (maxq = 2^bits - 1, perchannel, sym),
 * find_params(x, weight) :33-120, quantize :6-8, 111-114.  View [A, C, inner], channel of element i = (i/inner) % C.
 * All arithmetic fp32 (find_params' fp32 zeros promote half inputs); output fp32.
 */
int oracle_int_quantize(const void *in, float *out, int64_t A, int64_t C, int64_t inner, int dt, int bits) {
    const int64_t n = A * C * inner;
    float *mn = (float *)malloc(sizeof(float) * (size_t)(C > 0 ? C : 1) * 3), *mx = mn + C, *sc = mx + C;
    for (int64_t c = 0; c < C; ++c) { mn[c] = INFINITY; mx[c] = -INFINITY; }
    for (int64_t i = 0; i < n; ++i) {
        float v = load_elt(in, dt, i); int64_t c = (i / inner) % C;
        if (v < mn[c]) mn[c] = v;
        if (v > mx[c]) mx[c] = v;
    }
    const float maxq = (float)((1ll << bits) - 1), zero = (float)((double)(1ll << bits) / 2.0);   /* :23, :69 */
    for (int64_t c = 0; c < C; ++c) {
        float xmin = mn[c] < 0.0f ? mn[c] : 0.0f, xmax = mx[c] > 0.0f ? mx[c] : 0.0f;             /* :55-56 */
        if (fabsf(xmin) > xmax) xmax = fabsf(xmin);                                                /* :59 */
        if (xmin < 0.0f) xmin = -xmax;                                                             /* :60-62 */
        if (xmin == 0.0f && xmax == 0.0f) { xmin = -1.0f; xmax = 1.0f; }                           /* :63-65 */
        sc[c] = (xmax - xmin) / maxq;                                                              /* :67 */
    }
    for (int64_t i = 0; i < n; ++i) {
        float x = load_elt(in, dt, i), s = sc[(i / inner) % C];
        float q = rintf(x / s) + zero;                                                             /* :7 */
        q = q < 0.0f ? 0.0f : (q > maxq ? maxq : q);
        out[i] = s * (q - zero);                                                                   /* :8 */
    }
    free(mn);
    return 0;
}

/* ---- float_to_bfp_blocked for the 'bfp' + structured case ---------------
 * bfp_ops.py:124-149: first == 's' -> Q(S(t)), anything else -> S(Q(t)).
 * order: 0 = quantise only, 1 = sparsify then quantise, 2 = quantise then sparsify,
 * 3 = sparsify only (sparsity_num_format == 'fp32').  tmp: scratch of rows*K
 * elements of the widest intermediate dtype (4 bytes each is always enough).
 */
int oracle_float_to_bfp_blocked(const void *in, void *out, void *tmp, int64_t rows, int64_t K, int dt, int B, int m,
                                float eps, int stoc, const float *rand_u, int N, int M, int order, int tie_rule) {
    const int q_out_dt = stoc ? DT_F32 : dt;
    switch (order) {
    case 0: return oracle_bfp_quantize(in, out, rows, K, dt, B, m, eps, stoc, rand_u);
    case 3: return oracle_nm_sparsify(in, out, rows, K, dt, N, M, tie_rule);
    case 1: {
        int rc = oracle_nm_sparsify(in, tmp, rows, K, dt, N, M, tie_rule);
        if (rc) return rc;
        return oracle_bfp_quantize(tmp, out, rows, K, dt, B, m, eps, stoc, rand_u);
    }
    case 2: {
        int rc = oracle_bfp_quantize(in, tmp, rows, K, dt, B, m, eps, stoc, rand_u);
        if (rc) return rc;
        return oracle_nm_sparsify(tmp, out, rows, K, q_out_dt, N, M, tie_rule);
    }
    default: return 3;
    }
}

/* ---- packed form ---------------------------------------------------------
 * Not a reference function: the reference only ever returns dequantised floats.
 * The packed layout is the product's GEMM operand format; its
 * contract is unpack(pack(x)) == float_to_bfp_blocked(x) up to the sign of zero.
 * This restates the contract from the fake-quant output y and exponent e:
 *   q = y / 2^(e-m)  (exact integer, |q| <= 2^m - 1),  exp byte = e.
 */
int oracle_bfp_pack_from_fakequant(const float *y, const float *e_blk, int8_t *mant, int8_t *exp_out, int64_t rows,
                                   int64_t K, int B, int m) {
    const int64_t nblk = (K + B - 1) / B;
    if (m > 7) return 1;
#pragma omp parallel for schedule(static)
    for (int64_t blk = 0; blk < rows * nblk; ++blk) {
        int64_t r = blk / nblk, kb = blk % nblk, base = r * K + kb * B;
        int n = (int)((kb * B + B <= K) ? B : (K - kb * B));
        float e = e_blk[blk];
        exp_out[blk] = (int8_t)e;
        for (int i = 0; i < n; ++i) mant[base + i] = (int8_t)ldexpf(y[base + i], (int)(m - e));
    }
    return 0;
}

/* ---- BFP linear ----------------------------------------------------------
 * bfp_ops.py:187-190 / 278-287: F.linear on the dequantised operands, bias not
 * quantised.  The reference's fp32 GEMM summation order is the library's; the
 * oracle accumulates in double and rounds once, so it is the exact value the
 * stated tolerance (relative error <= 1e-5) is measured against.
 * xq [T,K], wq [Nout,K] fp32, bias [Nout] or NULL, out [T,Nout] fp32.
 */
int oracle_linear_f64acc(const float *xq, const float *wq, const float *bias, float *out, int64_t T, int64_t Nout,
                         int64_t K) {
#pragma omp parallel for schedule(static) collapse(2)
    for (int64_t t = 0; t < T; ++t)
        for (int64_t n = 0; n < Nout; ++n) {
            double acc = 0.0;
            const float *xr = xq + t * K, *wr = wq + n * K;
            for (int64_t k = 0; k < K; ++k) acc += (double)xr[k] * (double)wr[k];
            if (bias) acc += (double)bias[n];
            out[t * Nout + n] = (float)acc;
        }
    return 0;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int oracle_version(void) { return 1; }
