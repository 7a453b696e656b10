// Host check of the two-thread third of scripts/proto/half_third.cu: both halves are run side by side (the shuffle becomes
// an assignment) and compared with the definition X[3q + r] = sum_{n < 32} x[n] exp(-2 pi i n (3q + r) / 48).
//   g++ -O2 -std=c++17 -I../../landau-poisson-solver_b200/csrc half_third_check.cpp -o /tmp/half_third_check && /tmp/half_third_check
#include "fc3.cuh"
#include <cmath>
#include <cstdio>
#include <cstdlib>
using namespace fc3;

struct Half { double2 y1[2][4]; };
static void first_stage(const double2 *x, int r, int h, Half &H)
{
  constexpr int L = 16, M = 48;
  for (int c = 0; c < 2; c++) {
    const int n2 = 2 * h + c;
    double2 v[4];
    for (int n1 = 0; n1 < 4; n1++) {
      const int l = 4 * n1 + n2;
      const double2 a0 = x[l], a1 = x[l + L];
      double2 b;
      if (r == 0) b = cadd(a0, a1);
      else {
        const double sg = r == 1 ? LP_SQ3H : -LP_SQ3H;
        b = make_double2(a0.x - 0.5 * a1.x + sg * a1.y, a0.y - 0.5 * a1.y - sg * a1.x);
        const int t = (r * l) % M;
        const double cs = Tw<M>::c(t), sn = -Tw<M>::s(t);
        b = make_double2(b.x * cs - b.y * sn, b.x * sn + b.y * cs);
      }
      v[n1] = b;
    }
    dft4<-1>(v[0], v[1], v[2], v[3]);
    for (int k1 = 0; k1 < 4; k1++) {
      const int t = (3 * n2 * k1) % M;
      const double cs = Tw<M>::c(t), sn = -Tw<M>::s(t);
      H.y1[c][k1] = make_double2(v[k1].x * cs - v[k1].y * sn, v[k1].x * sn + v[k1].y * cs);
    }
  }
}
int main()
{
  double worst = 0.;
  for (int r = 0; r < 3; r++) {
    double2 x[32];
    for (int n = 0; n < 32; n++) x[n] = make_double2(std::sin(0.37 * n + r) + 0.1 * n, std::cos(1.3 * n - r));
    Half H[2];
    first_stage(x, r, 0, H[0]); first_stage(x, r, 1, H[1]);
    double2 X[16];
    for (int h = 0; h < 2; h++)
      for (int kk = 0; kk < 2; kk++) {
        const int k1 = 2 * h + kk;
        double2 mine[4] = {H[0].y1[0][k1], H[0].y1[1][k1], H[1].y1[0][k1], H[1].y1[1][k1]};   // n2 = 0..3 (two of them "shuffled in")
        dft4<-1>(mine[0], mine[1], mine[2], mine[3]);
        for (int k2 = 0; k2 < 4; k2++) X[k1 + 4 * k2] = mine[k2];
      }
    for (int q = 0; q < 16; q++) {
      double re = 0., im = 0.;
      for (int n = 0; n < 32; n++) {
        const double a = -2. * M_PI * n * (3 * q + r) / 48.;
        re += x[n].x * std::cos(a) - x[n].y * std::sin(a);
        im += x[n].x * std::sin(a) + x[n].y * std::cos(a);
      }
      worst = std::fmax(worst, std::fmax(std::fabs(re - X[q].x), std::fabs(im - X[q].y)));
    }
  }
  std::printf("max |error| of the two-thread third against the definition: %.3e\n", worst);
  return worst < 1e-12 ? 0 : 1;
}
