/* CPU ORACLE -- test infrastructure, not product code.
 *
 * Rectangular linear sum assignment: a C restatement of the algorithm that
 * scipy.optimize.linear_sum_assignment documents and implements (D. F. Crouse, "On
 * implementing 2D rectangular assignment algorithms", IEEE TAES 52(4), 2016: shortest
 * augmenting paths with dual variables).  The reference calls it at
 * lib/modeling/matcher.py:93 (one 10 x n_f problem per frame) and :158 (one Q x n_v
 * problem per video).  scipy is a third-party dependency of the reference
 * (requirements.txt:3, unpinned; the build container has scipy 1.18.1) whose source is not
 * under /root/reference, so the algorithm is restated from the paper and scipy's documented
 * behaviour, and pinned against scipy 1.18.1 outputs in tests/golden/lsap_*.npz
 * (tie-breaking included: constant matrices give the identity, rows come back ascending,
 * tall matrices are solved on the transpose).
 *
 * Build: make -C oracle    ->  oracle/_build/liblsap_oracle.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LSAP_OK 0
#define LSAP_INFEASIBLE -1
#define LSAP_INVALID -2

/* One shortest-augmenting-path search starting from free row `start`.  Returns the sink
 * column (or -1 if infeasible) and the path length through *min_val. */
static int64_t augment(int64_t nc, const double *cost, const double *u, const double *v,
                       int64_t *path, const int64_t *row4col, double *spc, int64_t start,
                       uint8_t *SR, uint8_t *SC, int64_t *remaining, double *min_val)
{
    double best = 0.0;
    int64_t n_rem = nc, i = start, sink = -1;
    /* candidate columns are kept in reverse order so that a constant matrix yields the
       identity assignment */
    for (int64_t t = 0; t < nc; ++t) remaining[t] = nc - t - 1;
    /* SR is cleared by the caller (it is indexed by row, not by column) */
    for (int64_t j = 0; j < nc; ++j) { SC[j] = 0; spc[j] = INFINITY; }

    while (sink == -1) {
        int64_t pick = -1;
        double lowest = INFINITY;
        SR[i] = 1;
        for (int64_t t = 0; t < n_rem; ++t) {
            int64_t j = remaining[t];
            double r = best + cost[i * nc + j] - u[i] - v[j];
            if (r < spc[j]) { path[j] = i; spc[j] = r; }
            /* among equal minima prefer a column that is still unassigned */
            if (spc[j] < lowest || (spc[j] == lowest && row4col[j] == -1)) {
                lowest = spc[j];
                pick = t;
            }
        }
        best = lowest;
        if (best == INFINITY) return -1;
        int64_t j = remaining[pick];
        if (row4col[j] == -1) sink = j; else i = row4col[j];
        SC[j] = 1;
        remaining[pick] = remaining[--n_rem];
    }
    *min_val = best;
    return sink;
}

/* cost: nr x nc row-major doubles.  a, b: outputs of length min(nr, nc). */
int lsap_solve(int64_t nr, int64_t nc, const double *cost_in, int64_t *a, int64_t *b)
{
    if (nr == 0 || nc == 0) return LSAP_OK;
    int transpose = nc < nr;
    double *tmp = NULL;
    const double *cost = cost_in;
    if (transpose) {
        tmp = (double *)malloc(sizeof(double) * (size_t)(nr * nc));
        for (int64_t i = 0; i < nr; ++i)
            for (int64_t j = 0; j < nc; ++j) tmp[j * nr + i] = cost_in[i * nc + j];
        int64_t t = nr; nr = nc; nc = t;
        cost = tmp;
    }
    for (int64_t k = 0; k < nr * nc; ++k)
        if (cost[k] != cost[k] || cost[k] == -INFINITY) { free(tmp); return LSAP_INVALID; }

    double *u = (double *)calloc((size_t)nr, sizeof(double));
    double *v = (double *)calloc((size_t)nc, sizeof(double));
    double *spc = (double *)malloc(sizeof(double) * (size_t)nc);
    int64_t *path = (int64_t *)malloc(sizeof(int64_t) * (size_t)nc);
    int64_t *col4row = (int64_t *)malloc(sizeof(int64_t) * (size_t)nr);
    int64_t *row4col = (int64_t *)malloc(sizeof(int64_t) * (size_t)nc);
    int64_t *remaining = (int64_t *)malloc(sizeof(int64_t) * (size_t)nc);
    uint8_t *SR = (uint8_t *)malloc((size_t)nr);
    uint8_t *SC = (uint8_t *)malloc((size_t)nc);
    for (int64_t j = 0; j < nc; ++j) { path[j] = -1; row4col[j] = -1; }
    for (int64_t i = 0; i < nr; ++i) col4row[i] = -1;

    int rc = LSAP_OK;
    for (int64_t cur = 0; cur < nr; ++cur) {
        double mv = 0.0;
        memset(SR, 0, (size_t)nr);
        int64_t sink = augment(nc, cost, u, v, path, row4col, spc, cur, SR, SC, remaining, &mv);
        if (sink < 0) { rc = LSAP_INFEASIBLE; break; }
        /* dual update */
        u[cur] += mv;
        for (int64_t i = 0; i < nr; ++i)
            if (SR[i] && i != cur) u[i] += mv - spc[col4row[i]];
        for (int64_t j = 0; j < nc; ++j)
            if (SC[j]) v[j] -= mv - spc[j];
        /* flip the path */
        int64_t j = sink;
        for (;;) {
            int64_t i = path[j];
            row4col[j] = i;
            int64_t prev = col4row[i];
            col4row[i] = j;
            j = prev;
            if (i == cur) break;
        }
    }
    if (rc == LSAP_OK) {
        if (transpose) {
            /* rows of the original matrix are the columns here: emit them ascending */
            int64_t k = 0;
            for (int64_t j = 0; j < nc; ++j)
                if (row4col[j] != -1) { a[k] = j; b[k] = row4col[j]; ++k; }
        } else {
            for (int64_t i = 0; i < nr; ++i) { a[i] = i; b[i] = col4row[i]; }
        }
    }
    free(tmp); free(u); free(v); free(spc); free(path); free(col4row); free(row4col);
    free(remaining); free(SR); free(SC);
    return rc;
}

/* Batched entry: `n` problems packed back to back (row-major), shapes in nr[]/nc[]. */
int lsap_solve_batch(int64_t n, const int64_t *nr, const int64_t *nc, const double *cost,
                     int64_t *a, int64_t *b)
{
    int64_t coff = 0, ooff = 0;
    for (int64_t p = 0; p < n; ++p) {
        int rc = lsap_solve(nr[p], nc[p], cost + coff, a + ooff, b + ooff);
        if (rc != LSAP_OK) return rc;
        coff += nr[p] * nc[p];
        ooff += nr[p] < nc[p] ? nr[p] : nc[p];
    }
    return LSAP_OK;
}
