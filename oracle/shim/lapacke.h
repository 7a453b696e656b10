/* oracle/shim/lapacke.h -- TEST INFRASTRUCTURE ONLY.
 * The reference calls Fortran-ABI dgetrf_/dgetri_ (conservationRoutines.cpp:215-216 etc.) to
 * invert a 5x5 (or 2x2 / 1x1) matrix once at start-up.  lapack_shim.cpp provides them with
 * partially pivoted LU (same algorithm class as LAPACK's). */
#ifndef LP_ORACLE_SHIM_LAPACKE_H
#define LP_ORACLE_SHIM_LAPACKE_H
#ifdef __cplusplus
extern "C" {
#endif
void dgetrf_(int *m, int *n, double *a, int *lda, int *ipiv, int *info);
void dgetri_(int *n, double *a, int *lda, int *ipiv, double *work, int *lwork, int *info);
#ifdef __cplusplus
}
#endif
#endif
