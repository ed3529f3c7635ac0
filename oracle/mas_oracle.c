/*
 * oracle/mas_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, scalar loops) of the reference's Monotonic
 * Alignment Search.  It exists to check the CUDA path and to serve as the
 * reported CPU baseline ("port") in bench.py.  Nothing under isp-tts_b200/
 * may import, link or call it: the product path is CUDA-only and fails
 * loudly without its extension.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file
 * bit-for-bit against tests/golden/mas_*.npz, which oracle/gen_golden.py
 * produced by running the reference's own numba `b_mas`
 * (/root/reference/tts/modules/aligner/mas.py:30-35) in the build container.
 *
 * What follows what:
 *   oracle_mas_width1  <- mas_width1, tts/modules/aligner/mas.py:8-26
 *   oracle_b_mas       <- b_mas,      tts/modules/aligner/mas.py:30-35
 *   durations          <- attn_hard.sum(dim=1),
 *                         tts/models/acoustic/modules/alignment.py:275
 *
 * Differences from the reference, on purpose:
 *   - the input is NOT mutated (the reference accumulates Q in place,
 *     mas.py:11-14; its GPU route clones first, alignment.py:321 -- that is
 *     the behaviour we target);
 *   - only the backpointer comparison is kept per cell (1 byte), not an int16
 *     index: prev_ind[i][j] == j - (Q[i-1][j-1] >= Q[i-1][j])  (mas.py:17).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* One utterance.  x: n rows (mel frames) x m cols (text tokens), row stride
 * ldx floats.  out: int16, row stride ldo, only the [n x m] window is
 * written (caller pre-zeroes).  path (optional): token index per frame.
 * scratch: 2*m floats + n*m bytes. */
static void oracle_mas_width1(const float *x, int64_t ldx, int64_t n, int64_t m,
                              int16_t *out, int64_t ldo, int16_t *path,
                              float *q_prev, float *q_cur, uint8_t *diag)
{
    /* mas.py:11  log_p[0, 1:] = -inf ; Q[0,0] = x[0,0] */
    q_prev[0] = x[0];
    for (int64_t j = 1; j < m; ++j) q_prev[j] = -INFINITY;

    for (int64_t i = 1; i < n; ++i) {
        const float *xi = x + i * ldx;
        uint8_t *di = diag + i * m;
        /* mas.py:12  log_p[:, 0] = cumsum(log_p[:, 0]): sequential fp32 adds */
        q_cur[0] = q_prev[0] + xi[0];
        di[0] = 0; /* prev_ind[:, 0] stays 0 (mas.py:16) */
        for (int64_t j = 1; j < m; ++j) {
            float a = q_prev[j - 1], b = q_prev[j];
            /* mas.py:17  ties (and -inf >= -inf) pick the diagonal j-1 */
            di[j] = (uint8_t)(a >= b);
            /* mas.py:14  one fp32 add per cell on top of the max */
            q_cur[j] = xi[j] + (a >= b ? a : b);
        }
        float *t = q_prev; q_prev = q_cur; q_cur = t;
    }

    /* mas.py:20-24 backtrack from the last token on the last frame */
    int64_t j = m - 1;
    for (int64_t i = n - 1; i >= 0; --i) {
        out[i * ldo + j] = 1;
        if (path) path[i] = (int16_t)j;
        if (i > 0) j -= diag[i * m + j];
    }
}

/*
 * b_attn_map: (B, T1, T2) fp32 contiguous; in_lens: text lengths (B);
 * out_lens: mel lengths (B).  attn_out: (B, T1, T2) int16, fully written
 * (zero outside each window).  durations (optional): (B, T2) int64 =
 * attn_out.sum(axis=1).  nthreads <= 0 -> all OpenMP threads (the
 * reference's prange over utterances, mas.py:32).  Returns 0, or -1 on a
 * length outside [1, T].
 */
int oracle_b_mas(const float *b_attn_map, int64_t B, int64_t T1, int64_t T2,
                 const int64_t *in_lens, const int64_t *out_lens,
                 int16_t *attn_out, int64_t *durations, int nthreads)
{
    for (int64_t b = 0; b < B; ++b) {
        if (in_lens[b] < 1 || in_lens[b] > T2 || out_lens[b] < 1 || out_lens[b] > T1)
            return -1;
    }
    memset(attn_out, 0, (size_t)(B * T1 * T2) * sizeof(int16_t));
    if (durations) memset(durations, 0, (size_t)(B * T2) * sizeof(int64_t));
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    int failed = 0;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int64_t b = 0; b < B; ++b) {
        int64_t n = out_lens[b], m = in_lens[b];
        float *q = (float *)malloc((size_t)(2 * m) * sizeof(float));
        uint8_t *diag = (uint8_t *)malloc((size_t)(n * m));
        int16_t *path = (int16_t *)malloc((size_t)n * sizeof(int16_t));
        if (!q || !diag || !path) {
            failed = 1;
        } else {
            oracle_mas_width1(b_attn_map + b * T1 * T2, T2, n, m,
                              attn_out + b * T1 * T2, T2, path, q, q + m, diag);
            if (durations)
                for (int64_t i = 0; i < n; ++i) durations[b * T2 + path[i]] += 1;
        }
        free(q); free(diag); free(path);
    }
    return failed ? -2 : 0;
}

/* Accumulated Q of one utterance (what the reference leaves in its mutated
 * input, mas.py:11-14) -- used by tests to pin the DP arithmetic itself. */
void oracle_mas_accumulate(const float *x, int64_t ldx, int64_t n, int64_t m, float *Q)
{
    Q[0] = x[0];
    for (int64_t j = 1; j < m; ++j) Q[j] = -INFINITY;
    for (int64_t i = 1; i < n; ++i) {
        const float *p = Q + (i - 1) * m;
        float *c = Q + i * m;
        const float *xi = x + i * ldx;
        c[0] = p[0] + xi[0];
        for (int64_t j = 1; j < m; ++j) {
            float a = p[j - 1], b = p[j];
            c[j] = xi[j] + (a >= b ? a : b);
        }
    }
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
