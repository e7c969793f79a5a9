/*
 * spa_oracle.c -- TEST INFRASTRUCTURE ONLY.  Not part of the shipped product.
 *
 * A scalar fp64 CPU restatement of the hot path of omkuprin7/ldpc-simulator
 * (python_ldpc_app), used only as the parity checker by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference
 * legs.  Nothing under ldpc-simulator_b200/ may import, link or call it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function
 * here against the .npz files under tests/golden/, which tests/golden/make_golden.py produced
 * by running the UNMODIFIED reference (imported from /root/reference) on
 * seeded inputs.  The reference ships no golden vectors of its own for this
 * path (python_ldpc_app/tests/test_integration.py:13-75,135-166 assert only
 * ranges).
 *
 * Each function cites the reference lines (relative to python_ldpc_app/) whose
 * behaviour it restates.  The reference is Python; this is a fresh C
 * formulation of the same arithmetic, not a translation of its data
 * structures (scipy LIL/CSR objects become flat CSR edge arrays).
 *
 * Build: see oracle/Makefile  (gcc -O2 -pthread -shared -fPIC).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

/* spa_decoder.py:140-146 -- the value substituted for tanh() outside +-17.5,
 * and spa_decoder.py:167 -- the clip applied to the atanh argument.          */
#define SPA_TANH_ARG_LIMIT 17.5
#define SPA_UNIT_CLIP      0.99999999999999878
/* spa_decoder.py:159 -- below this |tanh| the leave-one-out is a real product */
#define SPA_SMALL_TANH     1e-10
/* spa_decoder.py:218 -- magnitude gate of the "normalized LLR" metric         */
#define SPA_NORM_GATE      7.0

/* -------------------------------------------------------------------------- */
/* Column view of a CSR pattern: for every column, its edges (CSR positions) in
 * ascending row order.  This is the order in which scipy's csr_matvec on
 * E_csc.transpose() accumulates the column sum (spa_decoder.py:177-182).     */
static void build_column_view(int m, int n, const int32_t *rp, const int32_t *ci,
                              int32_t *cp, int32_t *cedge)
{
    int64_t nnz = rp[m];
    memset(cp, 0, sizeof(int32_t) * (size_t)(n + 1));
    for (int64_t e = 0; e < nnz; ++e) cp[ci[e] + 1]++;
    for (int j = 0; j < n; ++j) cp[j + 1] += cp[j];
    int32_t *fill = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    memcpy(fill, cp, sizeof(int32_t) * (size_t)n);
    for (int i = 0; i < m; ++i)
        for (int32_t e = rp[i]; e < rp[i + 1]; ++e) cedge[fill[ci[e]]++] = e;
    free(fill);
}

/* spa_decoder.py:133-146 */
static inline double tanh_half_clipped(double msg)
{
    double h = msg / 2.0;
    if (h > SPA_TANH_ARG_LIMIT) return SPA_UNIT_CLIP;
    if (h < -SPA_TANH_ARG_LIMIT) return -SPA_UNIT_CLIP;
    return tanh(h);
}

/* np.clip semantics (NaN passes through), spa_decoder.py:167 */
static inline double clip_unit(double r)
{
    if (r < -SPA_UNIT_CLIP) return -SPA_UNIT_CLIP;
    if (r > SPA_UNIT_CLIP) return SPA_UNIT_CLIP;
    return r;
}

/*
 * One frame of SPA_Decoder.decode (spa_decoder.py:63-280).
 *
 *   rp/ci       CSR pattern of the decoding graph (columns ascending in a row,
 *               which is the COO-of-CSR order of spa_decoder.py:41-61)
 *   cp/cedge    column view from build_column_view
 *   llr         n channel LLRs (p_data_buffer._channel_data, :88)
 *   z_out       n bytes: z = (posterior < 0), i.e. the COMPLEMENT of the
 *               decided bits, exactly what lands in _decoded_data (:233,245)
 *   conv_it     0-based index of the check-node pass after which the syndrome
 *               vanished, -1 if it never did (:65,232)
 *   post_out    n posteriors of the exit iteration, may be NULL
 *   post_trace  [max_iter][n] posteriors of every executed pass, may be NULL
 *   norm_out    the value left in _d_summarize_normalized_llr (:237-239,
 *               249-251) when calc_norm, may be NULL
 *   scratch     2*nnz + 3*n doubles
 *   odd_fix     0 = the reference.  1 = test-only variant matching the product's
 *               non-default LDPC_FLAG_FIX_ODD_SIGN switch (see the check-node pass)
 * returns 1 when the frame converged (Result.OK, :241), 0 otherwise (:253).
 *
 * max_iter must be >= 1: with max_iter <= 0 the reference loops until the
 * syndrome vanishes (:104,244), possibly forever; callers reject that.
 */
static int decode_one(int m, int n, const int32_t *rp, const int32_t *ci,
                      const int32_t *cp, const int32_t *cedge,
                      const double *llr, int max_iter, int calc_norm, int k_norm,
                      uint8_t *z_out, int32_t *conv_it, double *post_out,
                      double *post_trace, double *norm_out, double *scratch, int odd_fix)
{
    const int64_t nnz = rp[m];
    double *M = scratch;            /* variable->check messages, CSR edge order */
    double *E = scratch + nnz;      /* check->variable messages                 */
    double *post = E + nnz;
    double *prior = post + n;       /* "a priori" vector of the metric (:94,273) */
    double *tbuf = prior + n;       /* tanh values of one check row (<= n)      */
    double last_norm = 0.0;

    *conv_it = -1;                                            /* :65 */
    for (int i = 0; i < m; ++i)                               /* :88-90 */
        for (int32_t e = rp[i]; e < rp[i + 1]; ++e) M[e] = llr[ci[e]];
    memset(E, 0, sizeof(double) * (size_t)nnz);
    memcpy(prior, llr, sizeof(double) * (size_t)n);          /* :94 */
    memset(z_out, 0, (size_t)n);                              /* :70 */

    for (int it = 0; it < max_iter; ++it) {
        /* ---- check-node pass, spa_decoder.py:114-168 ---- */
        for (int i = 0; i < m; ++i) {
            const int32_t a = rp[i], b = rp[i + 1];
            const int d = b - a;
            if (d == 0) continue;                             /* :115-122 */
            double total = 1.0;
            for (int q = 0; q < d; ++q) {
                tbuf[q] = tanh_half_clipped(M[a + q]);
                total *= tbuf[q];                             /* :151-152 */
            }
            for (int q = 0; q < d; ++q) {
                double r;
                if (fabs(tbuf[q]) > SPA_SMALL_TANH) {         /* :159-161 */
                    r = total / tbuf[q];
                } else {                                      /* :162-164 */
                    r = 1.0;
                    for (int u = 0; u < d; ++u)
                        if (u != q) r *= tbuf[u];
                }
                E[a + q] = 2.0 * atanh(clip_unit(r));         /* :167-168 */
                /* NOT the reference (odd_fix = 0 everywhere the reference is restated): the product's
                 * non-default fix_odd_check_sign switch negates the extrinsics of odd-degree checks. */
                if (odd_fix && (d & 1)) E[a + q] = -E[a + q];
            }
        }
        /* ---- posterior, hard decision, spa_decoder.py:173-188 ---- */
        for (int j = 0; j < n; ++j) {
            double s = 0.0;
            for (int32_t q = cp[j]; q < cp[j + 1]; ++q) s += E[cedge[q]];
            post[j] = llr[j] + s;
            z_out[j] = (uint8_t)(post[j] < 0.0);
        }
        if (post_trace) memcpy(post_trace + (size_t)it * n, post, sizeof(double) * (size_t)n);
        /* ---- syndrome of the complemented decisions, :191-204 ---- */
        int finished = 1;
        for (int i = 0; i < m; ++i) {
            unsigned par = 0;
            for (int32_t e = rp[i]; e < rp[i + 1]; ++e) par ^= (unsigned)(z_out[ci[e]] ^ 1u);
            if (par & 1u) { finished = 0; break; }
        }
        /* ---- "normalized LLR" metric, :210-228 ---- */
        if (calc_norm) {
            int changes = 0;
            for (int j = 0; j < k_norm; ++j) {
                if (fabs(post[j]) > SPA_NORM_GATE) continue;
                if (prior[j] * post[j] < 0.0) ++changes;
            }
            last_norm = k_norm > 0 ? (double)changes / (double)k_norm : 0.0;
        }
        /* ---- exits: converged first (:231-241), then budget (:244-253) ---- */
        if (finished || it == max_iter - 1) {
            if (finished) *conv_it = it;
            if (post_out) memcpy(post_out, post, sizeof(double) * (size_t)n);
            if (norm_out) *norm_out = calc_norm ? last_norm : 0.0;
            return finished;
        }
        /* ---- variable-node pass, :260-268 ---- */
        for (int i = 0; i < m; ++i)
            for (int32_t e = rp[i]; e < rp[i + 1]; ++e) M[e] = post[ci[e]] - E[e];
        if (calc_norm) memcpy(prior, post, sizeof(double) * (size_t)n);   /* :273-274 */
    }
    return 0; /* unreachable for max_iter >= 1 */
}

/* Scratch doubles decode_one needs. */
static size_t scratch_doubles(int64_t nnz, int n) { return (size_t)(2 * nnz + 3 * (int64_t)n + 8); }

/*
 * Batch driver: frames are independent (the reference decodes one per call,
 * main.py:295-312, and fans frames over processes, main.py:248-256), so they
 * are handed out to POSIX threads through an atomic cursor.
 * llr [F][n] row-major; z [F][n]; conv_it [F]; ok [F]; post [F][n] or NULL;
 * norm [F] or NULL.  nthreads < 1 = all online cores.
 * Returns 0, or -1 on bad arguments / allocation failure.
 */
typedef struct {
    int odd_fix;
    double *trace;          /* [F][max_iter][n] posteriors of every executed pass, or NULL */
    int m, n, max_iter, calc_norm, k_norm;
    const int32_t *rp, *ci, *cp, *cedge;
    int64_t F, nnz;
    const double *llr;
    uint8_t *z, *ok;
    int32_t *conv_it;
    double *post, *norm;
    int64_t cursor;
    int fail;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *job = (batch_job *)arg;
    double *scratch = (double *)malloc(sizeof(double) * scratch_doubles(job->nnz, job->n));
    if (!scratch) { __atomic_store_n(&job->fail, 1, __ATOMIC_RELAXED); return NULL; }
    for (;;) {
        int64_t f = __atomic_fetch_add(&job->cursor, 1, __ATOMIC_RELAXED);
        if (f >= job->F) break;
        int32_t cit;
        double nl = 0.0;
        const int n = job->n;
        int good = decode_one(job->m, n, job->rp, job->ci, job->cp, job->cedge,
                              job->llr + (size_t)f * n, job->max_iter, job->calc_norm, job->k_norm,
                              job->z + (size_t)f * n, &cit,
                              job->post ? job->post + (size_t)f * n : NULL,
                              job->trace ? job->trace + (size_t)f * job->max_iter * n : NULL, &nl, scratch,
                              job->odd_fix);
        job->conv_it[f] = cit;
        job->ok[f] = (uint8_t)good;
        if (job->norm) job->norm[f] = nl;
    }
    free(scratch);
    return NULL;
}

int spa_oracle_decode_batch_ex(int m, int n, const int32_t *rp, const int32_t *ci,
                               int64_t F, const double *llr, int max_iter,
                               int calc_norm, int k_norm,
                               uint8_t *z, int32_t *conv_it, uint8_t *ok,
                               double *post, double *norm, int nthreads,
                               int odd_fix, double *post_trace);

int spa_oracle_decode_batch(int m, int n, const int32_t *rp, const int32_t *ci,
                            int64_t F, const double *llr, int max_iter,
                            int calc_norm, int k_norm,
                            uint8_t *z, int32_t *conv_it, uint8_t *ok,
                            double *post, double *norm, int nthreads)
{
    return spa_oracle_decode_batch_ex(m, n, rp, ci, F, llr, max_iter, calc_norm, k_norm, z, conv_it, ok,
                                      post, norm, nthreads, 0, NULL);
}

/* Same, plus odd_fix (see decode_one; 0 = the reference) and post_trace [F][max_iter][n] (or NULL):
 * the posterior of every executed pass of every frame (rows of passes that were not executed are
 * left untouched). */
int spa_oracle_decode_batch_ex(int m, int n, const int32_t *rp, const int32_t *ci,
                               int64_t F, const double *llr, int max_iter,
                               int calc_norm, int k_norm,
                               uint8_t *z, int32_t *conv_it, uint8_t *ok,
                               double *post, double *norm, int nthreads,
                               int odd_fix, double *post_trace)
{
    if (m < 0 || n <= 0 || F < 0 || max_iter < 1 || !rp || !ci || !llr || !z || !conv_it || !ok)
        return -1;
    const int64_t nnz = rp[m];
    int32_t *cp = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n + 1));
    int32_t *cedge = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
    if (!cp || !cedge) { free(cp); free(cedge); return -1; }
    build_column_view(m, n, rp, ci, cp, cedge);
    if (nthreads < 1) {
        long online = sysconf(_SC_NPROCESSORS_ONLN);
        nthreads = online > 0 ? (int)online : 1;
    }
    if ((int64_t)nthreads > F) nthreads = F > 0 ? (int)F : 1;
    batch_job job = { odd_fix, post_trace, m, n, max_iter, calc_norm, k_norm, rp, ci, cp, cedge, F, nnz, llr,
                      z, ok, conv_it, post, norm, 0, 0 };
    if (nthreads == 1) {
        batch_worker(&job);
    } else {
        pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
        int started = 0;
        if (!tid) job.fail = 1;
        for (int t = 0; tid && t < nthreads; ++t) {
            if (pthread_create(&tid[t], NULL, batch_worker, &job) != 0) break;
            ++started;
        }
        if (tid && started == 0) batch_worker(&job);
        for (int t = 0; t < started; ++t) pthread_join(tid[t], NULL);
        free(tid);
    }
    free(cp);
    free(cedge);
    return job.fail ? -1 : 0;
}

/* Single frame with the posterior of every executed pass (for the KAT checks). */
int spa_oracle_decode_trace(int m, int n, const int32_t *rp, const int32_t *ci,
                            const double *llr, int max_iter, uint8_t *z, int32_t *conv_it,
                            double *post_trace /* [max_iter][n] */)
{
    if (m < 0 || n <= 0 || max_iter < 1) return -1;
    const int64_t nnz = rp[m];
    int32_t *cp = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n + 1));
    int32_t *cedge = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
    double *scratch = (double *)malloc(sizeof(double) * scratch_doubles(nnz, n));
    if (!cp || !cedge || !scratch) { free(cp); free(cedge); free(scratch); return -1; }
    build_column_view(m, n, rp, ci, cp, cedge);
    int good = decode_one(m, n, rp, ci, cp, cedge, llr, max_iter, 0, 0, z, conv_it, NULL,
                          post_trace, NULL, scratch, 0);
    free(cp); free(cedge); free(scratch);
    return good;
}

/* -------------------------------------------------------------------------- */
/* channel.py:102-125 (mode 1): sigma = 1/sqrt(2*speed*10^(snr/10)).          */
double spa_oracle_sigma(double speed, double snr_db)
{
    return 1.0 / sqrt(2.0 * speed * pow(10.0, snr_db * 0.1));
}

/*
 * channel.py:38-81, mode 1, modulation 1 (BPSK) or 2 (amplitude 0.7), with the noise samples given
 * by the caller as unit normals g[]:  symbol = -1 for bit 0, +1 for bit 1
 * (:49); noise = g * sigma^2 when sigma_sq_quirk (the reference passes
 * sigma**2 as the *standard deviation*, :68) else g * sigma; y = symbol +
 * noise (:76); LLR = 2*y/sigma^2 (:80).
 */
void spa_oracle_channel_llr(int64_t count, const uint8_t *bits, const double *g,
                            double sigma, int sigma_sq_quirk, double amp, double *llr)
{
    const double s2 = sigma * sigma;
    const double dev = sigma_sq_quirk ? s2 : sigma;
    for (int64_t q = 0; q < count; ++q) {
        double sym = bits[q] == 0 ? -amp : amp;     /* amp = 1 (modulation 1, :49) or 0.7 (modulation 2, :51) */
        double y = sym + dev * g[q];
        llr[q] = 2.0 * y / s2;
    }
}

/*
 * main.py:314-339 folded over a batch: counters[0]=frames, [1]=failed frames
 * (FER numerator, :319-321), [2]=info-bit errors counted ONLY in failed frames
 * with the decoder output un-complemented (:326-330), [3]=sum of
 * convergence_iteration over converged frames, [4]=number of converged frames
 * (:336-339).  data [F][k] are the transmitted info bits (NULL = all zero).
 */
void spa_oracle_count_errors(int64_t F, int n, int k, const uint8_t *z, const uint8_t *ok,
                             const int32_t *conv_it, const uint8_t *data, uint64_t counters[5])
{
    for (int64_t f = 0; f < F; ++f) {
        counters[0] += 1;
        if (!ok[f]) {
            counters[1] += 1;
            for (int j = 0; j < k; ++j) {
                unsigned est = (unsigned)(z[(size_t)f * n + j] ^ 1u);
                unsigned sent = data ? data[(size_t)f * k + j] : 0u;
                if (est != sent) counters[2] += 1;
            }
        }
        if (conv_it[f] >= 0) {
            counters[3] += (uint64_t)conv_it[f];
            counters[4] += 1;
        }
    }
}

/* -------------------------------------------------------------------------- */
/*
 * encoder_decoder_data.py:13-183 + 269-317: GF(2) Gauss-Jordan of H into
 * [A | I] with the reference's pivoting rule restated on a dense byte matrix:
 * columns are visited left to right; the pivot is the first row at or below
 * the cursor with a 1 in that column (:45-48); that row is swapped up
 * (:56-62); rows below are cleared (:80-118); after the sweep every pivot
 * column is cleared above its pivot (:129-174).  When the rank r is smaller
 * than m only the first r rows are kept (:289-305, a second elimination of an
 * already reduced matrix changes nothing).  The column permutation is
 * "non-pivot columns in ascending order, then the pivot columns in pivot
 * order" (:307-313) and H_std[:, c] = H_reduced[:, perm[c]] (:315).
 *
 *   dense_in  [m][n] bytes (0/1), not modified
 *   dense_out [m][n] bytes, first *rank_out rows valid, columns permuted
 *   perm_out  [n]
 */
int spa_oracle_standard_form(int m, int n, const uint8_t *dense_in, uint8_t *dense_out,
                             int32_t *perm_out, int32_t *rank_out)
{
    uint8_t *w = (uint8_t *)malloc((size_t)m * n);
    int32_t *piv = (int32_t *)malloc(sizeof(int32_t) * (size_t)(m > 0 ? m : 1));
    uint8_t *is_piv = (uint8_t *)calloc((size_t)n, 1);
    if (!w || !piv || !is_piv) { free(w); free(piv); free(is_piv); return -1; }
    memcpy(w, dense_in, (size_t)m * n);
    int cur = 0, npiv = 0;
    for (int c = 0; c < n && npiv < m; ++c) {
        int p = -1;
        for (int r = cur; r < m; ++r) if (w[(size_t)r * n + c]) { p = r; break; }
        if (p < 0) continue;
        if (p > cur)
            for (int x = 0; x < n; ++x) {
                uint8_t t = w[(size_t)p * n + x];
                w[(size_t)p * n + x] = w[(size_t)cur * n + x];
                w[(size_t)cur * n + x] = t;
            }
        for (int r = cur + 1; r < m; ++r)
            if (w[(size_t)r * n + c])
                for (int x = 0; x < n; ++x) w[(size_t)r * n + x] ^= w[(size_t)cur * n + x];
        piv[npiv++] = c;
        ++cur;
    }
    for (int d = 0; d < npiv; ++d) {
        int c = piv[d];
        for (int r = 0; r < d; ++r)
            if (w[(size_t)r * n + c])
                for (int x = 0; x < n; ++x) w[(size_t)r * n + x] ^= w[(size_t)d * n + x];
    }
    for (int d = 0; d < npiv; ++d) is_piv[piv[d]] = 1;
    int q = 0;
    for (int c = 0; c < n; ++c) if (!is_piv[c]) perm_out[q++] = c;
    for (int d = 0; d < npiv; ++d) perm_out[q++] = piv[d];
    memset(dense_out, 0, (size_t)m * n);
    for (int r = 0; r < npiv; ++r)
        for (int c = 0; c < n; ++c) dense_out[(size_t)r * n + c] = w[(size_t)r * n + perm_out[c]];
    *rank_out = npiv;
    free(w); free(piv); free(is_piv);
    return 0;
}

/*
 * encoder_decoder_data.py:319-344 + data_buffer.py:47-82: with H_std = [A | I_r],
 * G = [I_k | A^T] and the codeword of u is G^T u mod 2 = [u | A u].
 * h_std [r][n] dense bytes, u [k] with k = n - r, cw [n].
 */
void spa_oracle_encode(int r, int n, const uint8_t *h_std, const uint8_t *u, uint8_t *cw)
{
    const int k = n - r;
    for (int j = 0; j < k; ++j) cw[j] = u[j] & 1u;
    for (int i = 0; i < r; ++i) {
        unsigned acc = 0;
        for (int j = 0; j < k; ++j) acc ^= (unsigned)(h_std[(size_t)i * n + j] & u[j]);
        cw[k + i] = (uint8_t)(acc & 1u);
    }
}
