/* oracle/shim/lapack_shim.cpp -- TEST INFRASTRUCTURE ONLY.  See lapacke.h in this directory.
 * Column-major, 1-based pivots like LAPACK; only tiny n is ever used. */
#include "lapacke.h"
#include <cmath>
#include <vector>
extern "C" void dgetrf_(int *m, int *n, double *a, int *lda, int *ipiv, int *info)
{
  const int M = *m, Nn = *n, L = *lda;
  *info = 0;
  const int K = M < Nn ? M : Nn;
  for (int j = 0; j < K; j++) {
    int p = j; double best = fabs(a[j + j * L]);
    for (int i = j + 1; i < M; i++) if (fabs(a[i + j * L]) > best) { best = fabs(a[i + j * L]); p = i; }
    ipiv[j] = p + 1;
    if (best == 0.) { if (*info == 0) *info = j + 1; continue; }
    if (p != j) for (int c = 0; c < Nn; c++) { double t = a[j + c * L]; a[j + c * L] = a[p + c * L]; a[p + c * L] = t; }
    for (int i = j + 1; i < M; i++) a[i + j * L] /= a[j + j * L];
    for (int c = j + 1; c < Nn; c++) for (int i = j + 1; i < M; i++) a[i + c * L] -= a[i + j * L] * a[j + c * L];
  }
}
extern "C" void dgetri_(int *n, double *a, int *lda, int *ipiv, double *, int *, int *info)
{
  const int Nn = *n, L = *lda;
  *info = 0;
  std::vector<double> inv(Nn * Nn, 0.), col(Nn);
  /* Solve A X = I column by column using P A = L U. */
  for (int c = 0; c < Nn; c++) {
    for (int i = 0; i < Nn; i++) col[i] = (i == c) ? 1. : 0.;
    for (int i = 0; i < Nn; i++) { int p = ipiv[i] - 1; if (p != i) { double t = col[i]; col[i] = col[p]; col[p] = t; } }
    for (int i = 0; i < Nn; i++) for (int k = 0; k < i; k++) col[i] -= a[i + k * L] * col[k];
    for (int i = Nn - 1; i >= 0; i--) { for (int k = i + 1; k < Nn; k++) col[i] -= a[i + k * L] * col[k]; col[i] /= a[i + i * L]; }
    for (int i = 0; i < Nn; i++) inv[i + c * Nn] = col[i];
  }
  for (int c = 0; c < Nn; c++) for (int i = 0; i < Nn; i++) a[i + c * L] = inv[i + c * Nn];
}
