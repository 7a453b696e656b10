/* oracle/shim/fftw_shim.cpp -- TEST INFRASTRUCTURE ONLY.  See fftw3.h in this directory. */
#include "fftw3.h"
#include <cmath>
#include <cstdlib>
#include <vector>
struct lp_shim_plan {
  int n[3];
  fftw_complex *buf;
  int sign;
  std::vector<double> tw[3]; /* tw[d][2*(j*k mod n)] = cos, sin of sign*2*pi*(jk mod n)/n */
};
extern "C" {
void *fftw_malloc(size_t n) { return malloc(n); }
void fftw_free(void *p) { free(p); }
int fftw_init_threads(void) { return 1; }
void fftw_plan_with_nthreads(int) {}
fftw_plan fftw_plan_dft_3d(int n0, int n1, int n2, fftw_complex *in, fftw_complex *out, int sign, unsigned)
{
  if (in != out) return NULL; /* the reference only plans in place */
  lp_shim_plan *p = new lp_shim_plan;
  p->n[0] = n0; p->n[1] = n1; p->n[2] = n2; p->buf = in; p->sign = sign;
  for (int d = 0; d < 3; d++) {
    int n = p->n[d];
    p->tw[d].resize(2 * n);
    for (int r = 0; r < n; r++) {
      double ang = sign * 2.0 * M_PI * (double)r / (double)n;
      p->tw[d][2 * r] = cos(ang);
      p->tw[d][2 * r + 1] = sin(ang);
    }
  }
  return p;
}
static void dft_axis(fftw_complex *x, int n, long stride, long nlines, const long *starts, const std::vector<double> &tw)
{
  #pragma omp parallel
  {
    std::vector<double> tmp(2 * n);
    #pragma omp for
    for (long L = 0; L < nlines; L++) {
      fftw_complex *base = x + starts[L];
      for (int k = 0; k < n; k++) {
        double sr = 0., si = 0.;
        for (int j = 0; j < n; j++) {
          int r = (int)(((long)j * k) % n);
          double c = tw[2 * r], s = tw[2 * r + 1];
          double xr = base[j * stride][0], xi = base[j * stride][1];
          sr += xr * c - xi * s;
          si += xr * s + xi * c;
        }
        tmp[2 * k] = sr; tmp[2 * k + 1] = si;
      }
      for (int k = 0; k < n; k++) { base[k * stride][0] = tmp[2 * k]; base[k * stride][1] = tmp[2 * k + 1]; }
    }
  }
}
void fftw_execute(const fftw_plan p)
{
  const int n0 = p->n[0], n1 = p->n[1], n2 = p->n[2];
  std::vector<long> starts;
  /* axis 2 (contiguous) */
  starts.clear();
  for (long a = 0; a < n0; a++) for (long b = 0; b < n1; b++) starts.push_back((a * n1 + b) * n2);
  dft_axis(p->buf, n2, 1, (long)starts.size(), starts.data(), p->tw[2]);
  /* axis 1 */
  starts.clear();
  for (long a = 0; a < n0; a++) for (long c = 0; c < n2; c++) starts.push_back(a * n1 * n2 + c);
  dft_axis(p->buf, n1, n2, (long)starts.size(), starts.data(), p->tw[1]);
  /* axis 0 */
  starts.clear();
  for (long b = 0; b < n1; b++) for (long c = 0; c < n2; c++) starts.push_back(b * n2 + c);
  dft_axis(p->buf, n0, (long)n1 * n2, (long)starts.size(), starts.data(), p->tw[0]);
}
void fftw_destroy_plan(fftw_plan p) { delete p; }
}
