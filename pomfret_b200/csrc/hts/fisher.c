/* hts-shim Fisher exact test for a 2x2 table (reference call site
 * blockjoin.c:3926; only `two < 0.001` is consumed).
 * Hypergeometric point probabilities are evaluated directly through lgamma;
 * the two-sided p-value sums every table whose probability does not exceed
 * the observed one (relative slack 1e-8, the convention used by klib/htslib
 * and by scipy.stats.fisher_exact). */
#include <math.h>
#include "htslib/kfunc.h"

static double log_choose(int n, int k) {
    if (k < 0 || k > n) return -INFINITY;
    return lgamma((double)n + 1.0) - lgamma((double)k + 1.0) - lgamma((double)(n - k) + 1.0);
}

/* P(X = x) for X ~ Hypergeometric(population n, successes row1, draws col1) */
static double hyper_pmf(int x, int row1, int col1, int n) {
    return exp(log_choose(row1, x) + log_choose(n - row1, col1 - x) - log_choose(n, col1));
}

double kt_fisher_exact(int n11, int n12, int n21, int n22, double *_left, double *_right, double *two) {
    const int row1 = n11 + n12, col1 = n11 + n21, n = n11 + n12 + n21 + n22;
    int hi = col1 < row1 ? col1 : row1;
    int lo = row1 + col1 - n;
    if (lo < 0) lo = 0;
    *two = *_left = *_right = 1.0;
    if (lo == hi) return 1.0;
    const double q = hyper_pmf(n11, row1, col1, n);
    double left = 0.0, right = 0.0, twosided = 0.0;
    for (int x = lo; x <= hi; x++) {
        double p = hyper_pmf(x, row1, col1, n);
        if (x <= n11) left += p;
        if (x >= n11) right += p;
        if (p < 1.00000001 * q) twosided += p;
    }
    if (twosided > 1.0) twosided = 1.0;
    if (left > 1.0) left = 1.0;
    if (right > 1.0) right = 1.0;
    *_left = left;
    *_right = right;
    *two = twosided;
    return q;
}
