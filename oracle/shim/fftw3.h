/* oracle/shim/fftw3.h -- TEST INFRASTRUCTURE ONLY.
 * Minimal FFTW3 stand-in for the reference build: the only plans the reference makes are two
 * in-place 3-D complex DFTs on the global `temp` (LP_ompi.cpp:327-328) executed at
 * collisionRoutines_1.cpp:304,346,382.  Semantics reproduced: unnormalised
 *   Y[k] = sum_j X[j] exp(sign * 2*pi*i * j*k / n)   per dimension, row-major [n0][n1][n2].
 * Implemented as three passes of a dense n-point DFT (exact to round-off for any n; n <= 32 here,
 * so cost is negligible beside ComputeQ).  Implementation lives in fftw_shim.cpp. */
#ifndef LP_ORACLE_SHIM_FFTW3_H
#define LP_ORACLE_SHIM_FFTW3_H
#include <stddef.h>
typedef double fftw_complex[2];
struct lp_shim_plan;
typedef struct lp_shim_plan *fftw_plan;
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#ifdef __cplusplus
extern "C" {
#endif
void *fftw_malloc(size_t n);
void fftw_free(void *p);
int fftw_init_threads(void);
void fftw_plan_with_nthreads(int n);
fftw_plan fftw_plan_dft_3d(int n0, int n1, int n2, fftw_complex *in, fftw_complex *out, int sign, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);
#ifdef __cplusplus
}
#endif
#endif
