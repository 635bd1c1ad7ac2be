/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported/linked by the product path.
 *
 * CPU restatement of the DTW the reference calls at
 *   /root/reference/whisper_timestamped/transcribe.py:1572-1581
 *       alignment = dtw.dtw(weights, step_pattern=dtw.stepPattern.symmetric1)
 * and of the "jumps" extraction at transcribe.py:1648-1652.
 *
 * The arithmetic itself lives in the third-party package `dtw-python` (unpinned in
 * /root/reference/requirements.txt:2; latest known 1.5.x), which is NOT vendored in
 * /root/reference and not installed in this image.  This file restates its published
 * algorithm (C `computeCM` + Python `backtrack`):
 *
 *   symmetric1 = three single-step patterns, scanned in this order
 *       p1: (i-1, j-1)   p2: (i, j-1)   p3: (i-1, j)      each cost = cm[pred] + 1.0*lm[i,j]
 *   cm[0,0] = lm[0,0]; cells filled column-major; out-of-range predecessors are NaN and
 *   never win; argmin uses strict '<' starting from +inf, so ties keep the EARLIER pattern.
 *   Backtrack from (n-1, m-1) following the stored pattern index until (0,0).
 *
 * PARITY UNPINNED: the reference ships no golden vector for this boundary;
 * the oracle is pinned against a brute-force optimal-path enumeration on small matrices and
 * against hand-computed tie cases (tests/test_oracle_dtw.py), not against dtw-python itself.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* Fill cost matrix cm[n*m] (row-major, n tokens x m frames) and step matrix sm[n*m]
 * from local costs lm[n*m] (float64).  Follows dtw-python computeCM for symmetric1. */
void oracle_dtw_fill(const double *lm, int n, int m, double *cm, int32_t *sm)
{
    const int di[3] = {1, 0, 1};
    const int dj[3] = {1, 1, 0};
    for (long k = 0; k < (long)n * m; ++k) { cm[k] = NAN; sm[k] = 0; }
    cm[0] = lm[0];
    for (int j = 0; j < m; ++j) {
        for (int i = 0; i < n; ++i) {
            if (!isnan(cm[(long)i * m + j])) continue;       /* already initialised (0,0) */
            double clist[3] = {NAN, NAN, NAN};
            for (int p = 0; p < 3; ++p) {
                int ii = i - di[p], jj = j - dj[p];
                if (ii >= 0 && jj >= 0) {
                    clist[p] = cm[(long)ii * m + jj];         /* starting cost  */
                    clist[p] += 1.0 * lm[(long)i * m + j];    /* one step       */
                }
            }
            int best = -1; double bestv = INFINITY;
            for (int p = 0; p < 3; ++p)
                if (clist[p] < bestv) { bestv = clist[p]; best = p; }  /* NaN never < */
            if (best >= 0) { cm[(long)i * m + j] = bestv; sm[(long)i * m + j] = best + 1; }
        }
    }
}

/* Backtrack (dtw-python `_backtrack`): returns path length; index1/index2 ascending. */
int oracle_dtw_backtrack(const int32_t *sm, int n, int m, int32_t *index1, int32_t *index2)
{
    const int di[4] = {0, 1, 0, 1};
    const int dj[4] = {0, 1, 1, 0};
    int cap = n + m;
    int32_t *ri = (int32_t *)malloc(sizeof(int32_t) * cap);
    int32_t *rj = (int32_t *)malloc(sizeof(int32_t) * cap);
    int i = n - 1, j = m - 1, len = 0;
    ri[len] = i; rj[len] = j; ++len;
    while (i > 0 || j > 0) {
        int s = sm[(long)i * m + j];
        if (s < 1 || s > 3) { free(ri); free(rj); return -1; }
        i -= di[s]; j -= dj[s];
        ri[len] = i; rj[len] = j; ++len;
    }
    for (int k = 0; k < len; ++k) { index1[k] = ri[len - 1 - k]; index2[k] = rj[len - 1 - k]; }
    free(ri); free(rj);
    return len;
}

/* jumps per transcribe.py:1648-1652: first frame of every token row on the path, then
 * the last frame index.  jumps has n+1 entries. */
void oracle_jumps(const int32_t *index1, const int32_t *index2, int len, int n, int32_t *jumps)
{
    int t = 0;
    for (int k = 0; k < len; ++k) {
        int jump = (k == 0) ? 1 : (index1[k] - index1[k - 1]);   /* np.diff, padded with 1 */
        if (jump != 0) jumps[t++] = index2[k];                    /* astype(bool)           */
    }
    jumps[t++] = index2[len - 1];
    (void)n;
}

/* Convenience: whole thing on one matrix.  Returns path length (or -1). */
int oracle_dtw(const double *lm, int n, int m, int32_t *index1, int32_t *index2, int32_t *jumps,
               double *distance)
{
    double *cm = (double *)malloc(sizeof(double) * (size_t)n * m);
    int32_t *sm = (int32_t *)malloc(sizeof(int32_t) * (size_t)n * m);
    oracle_dtw_fill(lm, n, m, cm, sm);
    if (distance) *distance = cm[(long)n * m - 1];
    int len = oracle_dtw_backtrack(sm, n, m, index1, index2);
    if (len > 0 && jumps) oracle_jumps(index1, index2, len, n, jumps);
    free(cm); free(sm);
    return len;
}

/* Batched variant used by the CPU baseline leg of bench.py: matrices stored back to back as
 * float32 (the product's cost layout), widened to float64 like transcribe.py:1550. */
int oracle_dtw_batch_f32(const float *cost, const int64_t *off, const int32_t *T, const int32_t *F,
                         int nseg, int32_t *jumps, const int64_t *joff)
{
    for (int s = 0; s < nseg; ++s) {
        int n = T[s], m = F[s];
        double *lm = (double *)malloc(sizeof(double) * (size_t)n * m);
        for (long k = 0; k < (long)n * m; ++k) lm[k] = (double)cost[off[s] + k];
        int32_t *i1 = (int32_t *)malloc(sizeof(int32_t) * (n + m));
        int32_t *i2 = (int32_t *)malloc(sizeof(int32_t) * (n + m));
        int len = oracle_dtw(lm, n, m, i1, i2, jumps + joff[s], NULL);
        free(lm); free(i1); free(i2);
        if (len < 0) return -1;
    }
    return 0;
}
