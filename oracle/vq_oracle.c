/*
 * vq_oracle.c -- CPU restatement of the reference vector quantiser.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity checker for the CUDA path.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it; the product
 * (vq_vae_gan_diffusion_b200/) never does.
 *
 * It restates /root/reference/network/vqvae/submodule/codebook.py::CodeBook.forward
 * (codebook.py:47-111) and the autograd backward PyTorch derives from it, in plain C:
 *
 *   codebook.py:62-66   NCHW -> (N, D) rows, N = B*H*W in (b, h, w) order
 *   codebook.py:70-79   d[n,k] = fl( fl(|z_n|^2 + |e_k|^2) - fl(2 * fl(z_n . e_k)) )     (fp32)
 *   codebook.py:82      idx[n] = first index of the row minimum (torch.argmin; a NaN distance counts as the minimum)
 *   codebook.py:85      e = E[idx]
 *   codebook.py:96-103  loss = mean((e - z)^2 + beta * mean((e - z)^2))
 *   codebook.py:106     z_q = fl(z + fl(e - z))            (straight-through value)
 *   codebook.py:109     returned as NHWC memory viewed NCHW -> here written as (N, D) rows
 *   backward            grad_z = g_out + g_loss * 2 (z - e) / (N D)
 *                       grad_E[idx[n]] += g_loss * beta * 2 (e - z) / (N D)
 *   histogram           bincount(idx, minlength=K)   (not in the reference; defined on its indices)
 *
 * Parity pinning: the reference ships no tests and no golden vectors.
 * This restatement is pinned against outputs of the reference itself, generated in the build
 * container by tests/golden/make_golden.py (imports /root/reference) and committed under
 * tests/golden/.  See tests/test_oracle_golden.py.
 *
 * Accumulation order.  The reference's reductions run inside ATen/BLAS whose summation order
 * is implementation defined (MKL sgemm on CPU, cuBLAS sgemm on GPU).  The oracle fixes ONE
 * order -- the "canonical order" -- and the CUDA path reproduces exactly this order in its
 * exact-fp32 re-rank, so CUDA-vs-oracle indices are bit-exact by construction, while
 * oracle-vs-reference differences are confined to rows that are exact fp32 ties or lie within
 * the rounding band of the reference formula (classified in tests/parity.py):
 *
 *   dot(x, y) over D terms: four partial sums, partial j accumulates the terms d == j (mod 4)
 *   in ascending d with one fused multiply-add each; result = (p0 + p1) + (p2 + p3).
 *   |x|^2 = dot(x, x).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define VQO_API __attribute__((visibility("default")))

/* canonical-order dot product; x has element stride sx, y has element stride sy */
static inline float vqo_dot(const float* x, int64_t sx, const float* y, int64_t sy, int D) {
    float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
    int d = 0;
    for (; d + 3 < D; d += 4) {
        p0 = fmaf(x[(d + 0) * sx], y[(d + 0) * sy], p0);
        p1 = fmaf(x[(d + 1) * sx], y[(d + 1) * sy], p1);
        p2 = fmaf(x[(d + 2) * sx], y[(d + 2) * sy], p2);
        p3 = fmaf(x[(d + 3) * sx], y[(d + 3) * sy], p3);
    }
    if (d < D) { p0 = fmaf(x[d * sx], y[d * sy], p0); d++; }
    if (d < D) { p1 = fmaf(x[d * sx], y[d * sy], p1); d++; }
    if (d < D) { p2 = fmaf(x[d * sx], y[d * sy], p2); d++; }
    return (p0 + p1) + (p2 + p3);
}

/* fp32 distance of the reference formula, codebook.py:70-79 */
static inline float vqo_dist(float z2, float e2, float dot) {
    volatile float s = z2 + e2;     /* fl(|z|^2 + |e|^2) */
    volatile float t = 2.0f * dot;  /* fl(2 * dot) (exact) */
    return s - t;
}

VQO_API int vq_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* |row|^2 of a (R, D) row-major matrix in canonical order */
VQO_API void vq_oracle_row_norms(const float* X, int64_t R, int D, float* out) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < R; r++) out[r] = vqo_dot(X + r * D, 1, X + r * D, 1, D);
}

/* torch.argmin bookkeeping for one more distance (codebook.py:82): NaN counts as smaller than every number, the FIRST
 * minimal element wins; n_best counts the codes attaining the minimum (NaN == NaN for this purpose) */
static inline void vqo_argmin_step(int64_t k, float dist, float* best, int64_t* best_k, int* n_best) {
    if (k == 0) { *best = dist; *best_k = k; *n_best = 1; }
    else if (isnan(*best)) { if (isnan(dist)) (*n_best)++; }
    else if (isnan(dist) || dist < *best) { *best = dist; *best_k = k; *n_best = 1; }
    else if (dist == *best) (*n_best)++;
}

/* scalar argmin stage: the definition */
static void vqo_argmin_scalar(const float* z_nchw, int64_t N, int64_t HW, int D, const float* E, const float* e2, int K,
                              int64_t* idx, float* dist_min, uint64_t* tie_rows) {
    uint64_t ties = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : ties)
    for (int64_t n = 0; n < N; n++) {
        const int64_t b = n / HW, hw = n % HW;
        const float* zr = z_nchw + b * (int64_t)D * HW + hw; /* element stride HW */
        const float z2 = vqo_dot(zr, HW, zr, HW, D);
        float best = INFINITY;
        int64_t best_k = 0;
        int n_best = 0;
        for (int k = 0; k < K; k++) {
            const float dot = vqo_dot(zr, HW, E + (int64_t)k * D, 1, D);
            vqo_argmin_step(k, vqo_dist(z2, e2[k], dot), &best, &best_k, &n_best);
        }
        idx[n] = best_k;
        if (dist_min) dist_min[n] = best;
        if (n_best > 1) ties++;
    }
    *tie_rows = ties;
}

/*
 * Vectorised argmin stage (AVX2 + FMA, 16 codes per step): the SAME arithmetic per (row, code) pair as the scalar stage
 * -- every SIMD lane runs vqo_dot's four fma chains over d == j (mod 4) in ascending d, then (p0 + p1) + (p2 + p3), then
 * vqo_dist -- with the codes of a block spread over the lanes, so the results are bit-identical (tests/test_oracle_golden.py
 * checks that on every case).  It exists so that the GPU tests can put ALL rows of the benchmarked configs
 * (262 144 latents x 16 384 codes) through the oracle in seconds.  D must be a multiple of 4 (else the caller uses the
 * scalar stage); the K % 16 remainder codes go through the scalar code.
 */
#if defined(__x86_64__)
#include <immintrin.h>
__attribute__((target("avx2,fma")))
static void vqo_argmin_avx2(const float* z_nchw, int64_t N, int64_t HW, int D, const float* E, const float* e2, int K,
                            int64_t* idx, float* dist_min, uint64_t* tie_rows) {
    const int kb_n = K / 16;                                  /* full blocks of 16 codes */
    /* Et[kb][d][16]: block-transposed codebook, so that one row value meets 16 codes per fma pair */
    float* Et = (float*)aligned_alloc(64, sizeof(float) * (size_t)(kb_n > 0 ? kb_n : 1) * (size_t)D * 16);
    if (!Et) { vqo_argmin_scalar(z_nchw, N, HW, D, E, e2, K, idx, dist_min, tie_rows); return; }
#pragma omp parallel for schedule(static)
    for (int kb = 0; kb < kb_n; kb++)
        for (int d = 0; d < D; d++)
            for (int l = 0; l < 16; l++) Et[((size_t)kb * D + d) * 16 + l] = E[(size_t)(kb * 16 + l) * D + d];
    uint64_t ties = 0;
#pragma omp parallel reduction(+ : ties)
    {
        float* zrow = (float*)aligned_alloc(64, sizeof(float) * (size_t)((D + 15) / 16 * 16));
#pragma omp for schedule(dynamic, 16)
        for (int64_t n = 0; n < N; n++) {
            const int64_t b = n / HW, hw = n % HW;
            const float* zr = z_nchw + b * (int64_t)D * HW + hw;
            for (int d = 0; d < D; d++) zrow[d] = zr[(int64_t)d * HW];
            const float z2 = vqo_dot(zrow, 1, zrow, 1, D);
            const __m256 z2v = _mm256_set1_ps(z2), two = _mm256_set1_ps(2.0f);
            float best = INFINITY;
            int64_t best_k = 0;
            int n_best = 0;
            for (int kb = 0; kb < kb_n; kb++) {
                const float* et = Et + (size_t)kb * D * 16;
                __m256 a0 = _mm256_setzero_ps(), a1 = a0, a2 = a0, a3 = a0, b0 = a0, b1 = a0, b2 = a0, b3 = a0;
                for (int d = 0; d < D; d += 4) {
                    const __m256 x0 = _mm256_set1_ps(zrow[d]), x1 = _mm256_set1_ps(zrow[d + 1]);
                    const __m256 x2 = _mm256_set1_ps(zrow[d + 2]), x3 = _mm256_set1_ps(zrow[d + 3]);
                    a0 = _mm256_fmadd_ps(x0, _mm256_load_ps(et + (d + 0) * 16), a0);
                    b0 = _mm256_fmadd_ps(x0, _mm256_load_ps(et + (d + 0) * 16 + 8), b0);
                    a1 = _mm256_fmadd_ps(x1, _mm256_load_ps(et + (d + 1) * 16), a1);
                    b1 = _mm256_fmadd_ps(x1, _mm256_load_ps(et + (d + 1) * 16 + 8), b1);
                    a2 = _mm256_fmadd_ps(x2, _mm256_load_ps(et + (d + 2) * 16), a2);
                    b2 = _mm256_fmadd_ps(x2, _mm256_load_ps(et + (d + 2) * 16 + 8), b2);
                    a3 = _mm256_fmadd_ps(x3, _mm256_load_ps(et + (d + 3) * 16), a3);
                    b3 = _mm256_fmadd_ps(x3, _mm256_load_ps(et + (d + 3) * 16 + 8), b3);
                }
                /* (p0 + p1) + (p2 + p3);  fl(fl(z2 + e2) - fl(2 dot)) */
                const __m256 dota = _mm256_add_ps(_mm256_add_ps(a0, a1), _mm256_add_ps(a2, a3));
                const __m256 dotb = _mm256_add_ps(_mm256_add_ps(b0, b1), _mm256_add_ps(b2, b3));
                float dist[16] __attribute__((aligned(32)));
                _mm256_store_ps(dist, _mm256_sub_ps(_mm256_add_ps(z2v, _mm256_loadu_ps(e2 + kb * 16)), _mm256_mul_ps(two, dota)));
                _mm256_store_ps(dist + 8, _mm256_sub_ps(_mm256_add_ps(z2v, _mm256_loadu_ps(e2 + kb * 16 + 8)), _mm256_mul_ps(two, dotb)));
                for (int l = 0; l < 16; l++) vqo_argmin_step((int64_t)kb * 16 + l, dist[l], &best, &best_k, &n_best);
            }
            for (int k = kb_n * 16; k < K; k++) {
                const float dot = vqo_dot(zrow, 1, E + (int64_t)k * D, 1, D);
                vqo_argmin_step(k, vqo_dist(z2, e2[k], dot), &best, &best_k, &n_best);
            }
            idx[n] = best_k;
            if (dist_min) dist_min[n] = best;
            if (n_best > 1) ties++;
        }
        free(zrow);
    }
    free(Et);
    *tie_rows = ties;
}
static int vqo_have_avx2(void) { return __builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma"); }
#else
static int vqo_have_avx2(void) { return 0; }
#endif

/* 1 when vq_oracle_forward(..., fast = 1) really runs the vectorised stage on this host */
VQO_API int vq_oracle_has_fast_path(void) { return vqo_have_avx2(); }

/*
 * Full forward.  z_nchw is contiguous (B, D, HW).  Any output pointer may be NULL.
 *   zq_nhwc  (N, D)   straight-through value fl(z + fl(e - z))
 *   idx      (N)      int64 argmin, first minimum
 *   loss     (1)      fp32
 *   hist     (K)      int64 bincount(idx)
 *   dist_min (N)      fp32 minimal distance (diagnostic)
 *   tie_rows (1)      rows whose minimal fp32 distance is attained by >= 2 codes
 *   fast              0: scalar argmin stage (the definition); 1: the vectorised stage when the host has AVX2 + FMA
 */
VQO_API int vq_oracle_forward(const float* z_nchw, int64_t B, int64_t HW, int D,
                              const float* E, int K, float beta,
                              float* zq_nhwc, int64_t* idx, float* loss, int64_t* hist,
                              float* dist_min, uint64_t* tie_rows, int fast) {
    if (B < 0 || HW < 0 || D <= 0 || K <= 0) return -1;
    const int64_t N = B * HW;
    float* e2 = (float*)malloc(sizeof(float) * (size_t)K);
    int64_t* idx_local = idx ? idx : (int64_t*)malloc(sizeof(int64_t) * (size_t)(N > 0 ? N : 1));
    if (!e2 || !idx_local) return -2;
    vq_oracle_row_norms(E, K, D, e2);

    uint64_t ties = 0;
#if defined(__x86_64__)
    if (fast && D % 4 == 0 && vqo_have_avx2()) vqo_argmin_avx2(z_nchw, N, HW, D, E, e2, K, idx_local, dist_min, &ties);
    else
#endif
        vqo_argmin_scalar(z_nchw, N, HW, D, E, e2, K, idx_local, dist_min, &ties);

    double sq_sum = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : sq_sum)
    for (int64_t n = 0; n < N; n++) {
        const int64_t b = n / HW, hw = n % HW;
        const float* zr = z_nchw + b * (int64_t)D * HW + hw;
        const float* e = E + idx_local[n] * D;
        for (int d = 0; d < D; d++) {
            const float zv = zr[(int64_t)d * HW];
            volatile float diff = e[d] - zv;               /* fl(e - z) */
            if (zq_nhwc) zq_nhwc[n * D + d] = zv + diff;   /* fl(z + fl(e - z)), codebook.py:106 */
            volatile float sq = diff * diff;
            sq_sum += (double)sq;
        }
    }
    if (tie_rows) *tie_rows = ties;
    if (loss) {
        /* mean(a + beta*mean(b)) with a == b elementwise, codebook.py:96-103 */
        const double m = (N > 0) ? sq_sum / ((double)N * (double)D) : NAN;
        *loss = (float)(m + (double)beta * m);
    }
    if (hist) {
        memset(hist, 0, sizeof(int64_t) * (size_t)K);
        for (int64_t n = 0; n < N; n++) hist[idx_local[n]]++;
    }
    free(e2);
    if (!idx) free(idx_local);
    return 0;
}

/*
 * Nearest code of row-major vectors under the broadcast-difference recipe of
 * /root/reference/network/continous_vq_diffusion/v_vq_diffusion.py:114-123:
 *     distances = torch.sum((x.unsqueeze(2) - codebook.unsqueeze(0).unsqueeze(0)) ** 2, dim=-1);  distances.argmin(-1)
 * i.e. d[n,k] = sum_d fl(fl(x_nd - e_kd)^2), each term rounded before it is added (no fused multiply-add), summed here
 * in the canonical order (four partial sums over d == j (mod 4), ascending, (p0 + p1) + (p2 + p3)); first minimum,
 * NaN counts as the minimum (torch.argmin).
 *   x_rows (N, D) row-major, E (K, D); idx (N) int64; dist_min (N) and tie_rows optional.
 */
static inline float vqo_diffsq(const float* x, const float* y, int D) {
    float p[4] = {0.f, 0.f, 0.f, 0.f};
    for (int d = 0; d < D; d++) {
        volatile float diff = x[d] - y[d];
        volatile float sq = diff * diff;
        volatile float acc = p[d & 3] + sq;
        p[d & 3] = acc;
    }
    volatile float a = p[0] + p[1], b = p[2] + p[3];
    return a + b;
}

VQO_API int vq_oracle_nearest_diffsq(const float* x_rows, int64_t N, int D, const float* E, int K,
                                     int64_t* idx, float* dist_min, uint64_t* tie_rows) {
    if (N < 0 || D <= 0 || K <= 0 || !idx) return -1;
    uint64_t ties = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : ties)
    for (int64_t n = 0; n < N; n++) {
        const float* x = x_rows + n * (int64_t)D;
        float best = INFINITY;
        int64_t best_k = 0;
        int n_best = 0;
        for (int k = 0; k < K; k++) {
            const float dist = vqo_diffsq(x, E + (int64_t)k * D, D);
            if (k == 0) { best = dist; best_k = k; n_best = 1; }
            else if (isnan(best)) { if (isnan(dist)) n_best++; }
            else if (isnan(dist) || dist < best) { best = dist; best_k = k; n_best = 1; }
            else if (dist == best) n_best++;
        }
        idx[n] = best_k;
        if (dist_min) dist_min[n] = best;
        if (n_best > 1) ties++;
    }
    if (tie_rows) *tie_rows = ties;
    return 0;
}

/*
 * Nearest table row under the recipe of VQGaussianDiffusion3DWrapper.gaussian_to_indices
 * (/root/reference/network/vqDiffusion/submodule/diffusion_gaussian3d.py:543-570):
 *     table = F.normalize(table, p=2, dim=-1); x = F.normalize(x, p=2, dim=-1)        :560-563
 *     distances = torch.cdist(x, table); distances.argmin(-1)                          :566-569
 * F.normalize divides by max(|v|_2, 1e-12) (clamp_min keeps a NaN).  torch.cdist (p = 2, more than 25 rows: the matmul
 * path, aten/src/ATen/native/Distance.cpp _euclidean_dist) evaluates the squared distance as ONE matrix product of the
 * augmented vectors [-2 x, |x|^2, 1] . [y, 1, |y|^2], then clamp_min(0).sqrt().  Canonical restatement: the D + 2 terms run
 * through vqo_dot's four fma chains (term d goes to chain d mod 4, ascending), so the two norm terms are the last terms of
 * chains D mod 4 and (D + 1) mod 4; norms are canonical dot products; first minimum of the square roots, NaN counts as
 * the minimum.
 *   x_rows (N, D), table (K, D) raw (normalised here); idx (N); dist_min (N), tie_rows, table_hat (K, D) optional outputs.
 */
static inline float vqo_normalize_denom(float norm2) {
    const float nrm = sqrtf(norm2);
    return (nrm < 1e-12f) ? 1e-12f : nrm;
}

static inline float vqo_cdist(const float* x, float xn, const float* y, float yn, int D) {
    float p[4] = {0.f, 0.f, 0.f, 0.f};
    for (int d = 0; d < D; d++) p[d & 3] = fmaf(-2.0f * x[d], y[d], p[d & 3]);
    p[D & 3] = fmaf(xn, 1.0f, p[D & 3]);
    p[(D + 1) & 3] = fmaf(1.0f, yn, p[(D + 1) & 3]);
    volatile float a = p[0] + p[1], b = p[2] + p[3];
    volatile float d2 = a + b;
    return sqrtf(d2 < 0.0f ? 0.0f : d2);                  /* clamp_min(0).sqrt(); a NaN stays a NaN */
}

VQO_API int vq_oracle_nearest_cdist(const float* x_rows, int64_t N, int D, const float* table, int K,
                                    int64_t* idx, float* dist_min, uint64_t* tie_rows, float* table_hat) {
    if (N < 0 || D <= 0 || K <= 0 || !idx) return -1;
    float* th = table_hat ? table_hat : (float*)malloc(sizeof(float) * (size_t)K * (size_t)D);
    float* tn = (float*)malloc(sizeof(float) * (size_t)K);
    if (!th || !tn) return -2;
#pragma omp parallel for schedule(static)
    for (int k = 0; k < K; k++) {
        const float* t = table + (int64_t)k * D;
        const float den = vqo_normalize_denom(vqo_dot(t, 1, t, 1, D));
        for (int d = 0; d < D; d++) th[(int64_t)k * D + d] = t[d] / den;
        tn[k] = vqo_dot(th + (int64_t)k * D, 1, th + (int64_t)k * D, 1, D);
    }
    uint64_t ties = 0;
#pragma omp parallel reduction(+ : ties)
    {
        float* xh = (float*)malloc(sizeof(float) * (size_t)D);
#pragma omp for schedule(dynamic, 16)
        for (int64_t n = 0; n < N; n++) {
            const float* x = x_rows + n * (int64_t)D;
            const float den = vqo_normalize_denom(vqo_dot(x, 1, x, 1, D));
            for (int d = 0; d < D; d++) xh[d] = x[d] / den;
            const float xn = vqo_dot(xh, 1, xh, 1, D);
            float best = INFINITY;
            int64_t best_k = 0;
            int n_best = 0;
            for (int k = 0; k < K; k++)
                vqo_argmin_step(k, vqo_cdist(xh, xn, th + (int64_t)k * D, tn[k], D), &best, &best_k, &n_best);
            idx[n] = best_k;
            if (dist_min) dist_min[n] = best;
            if (n_best > 1) ties++;
        }
        free(xh);
    }
    if (tie_rows) *tie_rows = ties;
    if (!table_hat) free(th);
    free(tn);
    return 0;
}

/*
 * Distances of selected (row, code) pairs in canonical order -- lets the tests classify a
 * disagreement without recomputing whole rows.
 */
VQO_API void vq_oracle_pair_dist(const float* z_nchw, int64_t B, int64_t HW, int D,
                                 const float* E, const int64_t* rows, const int64_t* codes,
                                 int64_t npairs, float* out) {
    (void)B;
    for (int64_t i = 0; i < npairs; i++) {
        const int64_t n = rows[i], b = n / HW, hw = n % HW;
        const float* zr = z_nchw + b * (int64_t)D * HW + hw;
        const float* e = E + codes[i] * D;
        out[i] = vqo_dist(vqo_dot(zr, HW, zr, HW, D), vqo_dot(e, 1, e, 1, D), vqo_dot(zr, HW, e, 1, D));
    }
}

/*
 * Backward of the reference forward as PyTorch autograd derives it.
 *   gout     upstream gradient on z_q, logical shape (B, D, HW), element strides gs[3] = {b, d, hw};
 *            may be NULL (treated as zero)
 *   g_loss   upstream gradient on the scalar loss
 *   n_global number of latents the loss mean ran over (= N on one device; the global N when the
 *            batch is sharded, so that summed shard gradients equal the single-device gradient)
 *   grad_z   (B, D, HW) contiguous NCHW
 *   grad_E   (K, D), overwritten
 */
VQO_API int vq_oracle_backward(const float* gout, const int64_t* gs, float g_loss,
                               const float* z_nchw, const int64_t* idx, const float* E,
                               int64_t B, int64_t HW, int D, int K, float beta, int64_t n_global,
                               float* grad_z, float* grad_E) {
    const int64_t N = B * HW;
    if (n_global <= 0) n_global = N;
    const float coef = (float)(2.0 * (double)g_loss / ((double)n_global * (double)D));
    const float coef_e = beta * coef;
    if (grad_E) memset(grad_E, 0, sizeof(float) * (size_t)K * (size_t)D);
    if (grad_z) {
#pragma omp parallel for schedule(static)
        for (int64_t n = 0; n < N; n++) {
            const int64_t b = n / HW, hw = n % HW;
            const float* e = E + idx[n] * D;
            for (int d = 0; d < D; d++) {
                const int64_t off = b * (int64_t)D * HW + (int64_t)d * HW + hw;
                const float g = gout ? gout[b * gs[0] + d * gs[1] + hw * gs[2]] : 0.f;
                grad_z[off] = fmaf(coef, z_nchw[off] - e[d], g);
            }
        }
    }
    if (grad_E) {
        /* serial, ascending n: a deterministic summation order for the scatter-add */
        for (int64_t n = 0; n < N; n++) {
            const int64_t b = n / HW, hw = n % HW;
            const float* e = E + idx[n] * D;
            float* ge = grad_E + idx[n] * D;
            for (int d = 0; d < D; d++)
                ge[d] += coef_e * (e[d] - z_nchw[b * (int64_t)D * HW + (int64_t)d * HW + hw]);
        }
    }
    return 0;
}
