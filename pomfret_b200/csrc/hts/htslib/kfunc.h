/* hts-shim: Fisher's exact test (reference call site blockjoin.c:3926). */
#ifndef POMFRET_HTS_SHIM_KFUNC_H
#define POMFRET_HTS_SHIM_KFUNC_H
#ifdef __cplusplus
extern "C" {
#endif
/* 2x2 table  n11 n12 / n21 n22.  Returns the table probability; writes the
 * left-tail, right-tail and two-sided p-values. */
double kt_fisher_exact(int n11, int n12, int n21, int n22, double *_left, double *_right, double *two);
#ifdef __cplusplus
}
#endif
#endif
