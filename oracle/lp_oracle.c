/* oracle/lp_oracle.c -- TEST INFRASTRUCTURE ONLY (see lp_oracle.h).
 *
 * CPU restatement of the reference's collision + advection hot path.  Every routine names the
 * reference lines it follows (paths relative to /root/reference/source).  It is written for
 * checking, not speed: loops keep the reference's summation order where that order is
 * observable (ComputeQ's omega loop, the conservation dot products, the IntModes projection in
 * "direct" mode); the default projection uses the separable 1-D tables that the literal
 * IntModes factorises into, which tests/test_oracle.py verifies against the literal form.
 *
 * The product never links or loads this file.
 */
#include "lp_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

struct lpo_ctx {
  int Nx, Nv, N, homogeneous, gamma, direct_intmodes;
  int fandl;                      /* FullandLinear: electron-ion term Q(f, ions) next to Q(f,f) */
  /* reference test 1: Doping = True (non-uniform background, Dirichlet walls), LinearLandau = True (Q(f,M)),
   * MassConsOnly = True */
  int doping, a_i, b_i, linear, mass_only;
  double NL, NH, eps, T_L, T_R, CCt_mass;
  double *dirL, *dirR;            /* DirichletBC coefficients at the left / right wall: 6 per velocity cell */
  double *mhat;                   /* DFTMaxwell: ncell * N^3 complex */
  double CCt_lin[4];              /* (C C^T)^-1 of the mass and energy rows (conservationRoutines.cpp:222-238) */
  int size_v, size_ft, ncell;
  double Lv, Lx, nu, dt, dv, dx, scalev, scaleL, scale3;
  double L_eta, h_v, h_eta;
  double *v, *eta, *wt;           /* N each */
  double *Sh;                     /* 7 * N^3: A, S11, S22, S33, S12, S13, S23 at omega */
  double *C5;                     /* 5 * N^3: C1_5[0], C2[1], C2[2], C2[3], C1_5[4] */
  double CCt[25];                 /* holds (C C^T)^-1 like the reference after dgetri */
  double *T1, *M1, *S1;           /* N*Nv complex each: 1-D IntModes factors [k*Nv + j] */
  int *node_cell;                 /* N: spectral node -> DG cell index per dimension */
  double *node_xi;                /* N: (v[l] - Gridv(j))/dv */
};

/* ------------------------------------------------------------------------------------------ */
/* grids: advection_1.cpp:15-21 */
static double gridv(const lpo_ctx *c, double m) { return -c->Lv + (m + 0.5) * c->dv; }
static double gridx(const lpo_ctx *c, double m) { return (m + 0.5) * c->dx; }

/* ------------------------------------------------------------------------------------------ */
/* Landau (gamma = -3) kernel symbols: collisionRoutines_1.cpp:18-36 */
static double s1hat(double R, double k1, double k2, double k3)
{
  if (k1 == 0. && k2 == 0. && k3 == 0.) return sqrt(1. / (2 * M_PI)) * R * R;
  double r2 = k1 * k1 + k2 * k2 + k3 * k3;
  return sqrt(2.0 / M_PI) * (1 - cos(R * sqrt(r2))) / r2;
}
static double s233hat(double R, double k1, double k2, double k3)
{
  double r = sqrt(k1 * k1 + k2 * k2 + k3 * k3);
  if (r == 0.) return sqrt(1. / (2. * M_PI)) * R * R / 3.;
  double Rr = R * r;
  return sqrt(2. / M_PI) * ((k1 * k1 + k2 * k2) * (Rr - sin(Rr)) / Rr
                            - k3 * k3 * (Rr + Rr * cos(Rr) - 2. * sin(Rr)) / Rr) / pow(r, 4.);
}
static double s213hat(double R, double k1, double k2, double k3)
{
  double r = sqrt(k1 * k1 + k2 * k2 + k3 * k3);
  if (k1 == 0. || k3 == 0.) return 0.;
  double Rr = R * r;
  return -sqrt(2 / M_PI) * k1 * k3 * (2. * Rr + Rr * cos(Rr) - 3. * sin(Rr)) / (R * pow(r, 5.));
}
/* Maxwell molecules (gamma = 0): collisionRoutines_1.cpp:38-66 */
static double s1hat_mm(double R, double k1, double k2, double k3)
{
  double r = sqrt(k1 * k1 + k2 * k2 + k3 * k3), Rr = R * r, cR = cos(Rr), sR = sin(Rr);
  if (r == 0.) return 2 * sqrt(1. / (2. * M_PI)) * pow(R, 5.) / 5.;
  return sqrt(2. / M_PI) * (-Rr * Rr * Rr * cR + 3 * Rr * Rr * sR + 6 * Rr * cR - 6 * sR) / pow(r, 5.);
}
static double s233hat_mm(double R, double k1, double k2, double k3)
{
  double r = sqrt(k1 * k1 + k2 * k2 + k3 * k3), Rr = R * r, cR = cos(Rr), sR = sin(Rr);
  if (r == 0.) return 2 * sqrt(1. / (2. * M_PI)) * pow(R, 5.) / 15.;
  return sqrt(2. / M_PI) * ((k1 * k1 + k2 * k2) * (-Rr * Rr * sR - 3 * Rr * cR + 3 * sR)
                            + k3 * k3 * (-Rr * Rr * Rr * cR + 5 * Rr * Rr * sR + 12 * Rr * cR - 12 * sR)) / pow(r, 7.);
}
static double s213hat_mm(double R, double k1, double k2, double k3)
{
  double r = sqrt(k1 * k1 + k2 * k2 + k3 * k3), Rr = R * r, cR = cos(Rr), sR = sin(Rr);
  if (k1 == 0. || k3 == 0.) return 0.;
  return sqrt(2. / M_PI) * k1 * k3 * (-Rr * Rr * Rr * cR + 6 * Rr * Rr * sR + 15 * Rr * cR - 15 * sR) / (pow(r, 7.));
}
/* hard spheres (gamma = 1): collisionRoutines_1.cpp:68-96 */
static double s1hat_hs(double R, double k1, double k2, double k3)
{
  double r = sqrt(k1 * k1 + k2 * k2 + k3 * k3), Rr = R * r, cR = cos(Rr), sR = sin(Rr);
  if (r == 0.) return sqrt(1. / (2. * M_PI)) * pow(R, 6.) / 3.;
  return sqrt(2. / M_PI) * (4. * (Rr * Rr - 6.) * Rr * sR - (Rr * Rr * (Rr * Rr - 12.) + 24.) * cR + 24.) / pow(r, 6.);
}
static double s233hat_hs(double R, double k1, double k2, double k3)
{
  double r = sqrt(k1 * k1 + k2 * k2 + k3 * k3), Rr = R * r, cR = cos(Rr), sR = sin(Rr);
  if (r == 0.) return sqrt(1. / (2. * M_PI)) * pow(R, 6.) / 9.;
  return sqrt(2. / M_PI) * ((k1 * k1 + k2 * k2) * ((8. - Rr * Rr) * Rr * sR + 4. * (2. - Rr * Rr) * cR - 8.)
                            + k3 * k3 * ((Rr * Rr * (20. - Rr * Rr) - 40.) * cR + (6. * Rr * Rr - 40.) * Rr * sR + 40.)) / pow(r, 8.);
}
static double s213hat_hs(double R, double k1, double k2, double k3)
{
  double r = sqrt(k1 * k1 + k2 * k2 + k3 * k3), Rr = R * r, cR = cos(Rr), sR = sin(Rr);
  if (k1 == 0. || k3 == 0.) return 0.;
  return sqrt(2. / M_PI) * k1 * k3 * ((Rr * Rr * (24. - Rr * Rr) - 48.) * cR + (7. * Rr * Rr - 48.) * Rr * sR + 48.) / (pow(r, 8.));
}
/* the symmetric 3x3 symbol matrix at omega: collisionRoutines_1.cpp:106-136 */
typedef double (*sym_fn)(double, double, double, double);
static void shat_matrix_g(int gamma, double R, double k1, double k2, double k3, double S[3][3])
{
  sym_fn f1 = gamma == 0 ? s1hat_mm : gamma == 1 ? s1hat_hs : s1hat;
  sym_fn f233 = gamma == 0 ? s233hat_mm : gamma == 1 ? s233hat_hs : s233hat;
  sym_fn f213 = gamma == 0 ? s213hat_mm : gamma == 1 ? s213hat_hs : s213hat;
  double s1 = f1(R, k1, k2, k3);
  S[0][0] = s1 - f233(R, k2, k3, k1);
  S[1][1] = s1 - f233(R, k1, k3, k2);
  S[2][2] = s1 - f233(R, k1, k2, k3);
  S[0][1] = -f213(R, k1, k3, k2);
  S[0][2] = -f213(R, k1, k2, k3);
  S[1][2] = -f213(R, k2, k1, k3);
  S[1][0] = S[0][1]; S[2][0] = S[0][2]; S[2][1] = S[1][2];
}
/* gHat3: collisionRoutines_1.cpp:98-161; gamma = -3 (:137-146), gamma = 0, 1 (:148-157) */
static double ghat3_from(int gamma, const double S[3][3], double A, int r_is_zero, const double z[3], const double k[3])
{
  double res = 0.;
  if (gamma != -3) {
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) res += S[i][j] * (2. * k[j] - z[j]) * z[i];
    return res;
  }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) res += S[i][j] * (z[i] - k[i]) * (z[j] - k[j]);
  return r_is_zero ? -res : A - res;
}
double lpo_gHat3(const lpo_ctx *c, double z1, double z2, double z3, double k1, double k2, double k3)
{
  double S[3][3], z[3] = {z1, z2, z3}, k[3] = {k1, k2, k3};
  double R = c->Lv, r = sqrt(k1 * k1 + k2 * k2 + k3 * k3);
  shat_matrix_g(c->gamma, R, k1, k2, k3, S);
  double A = (r == 0.) ? 0. : sqrt(8. / M_PI) * (R * r - sin(R * r)) / (R * r);
  return ghat3_from(c->gamma, S, A, r == 0., z, k);
}
static double weight_at(const lpo_ctx *c, int i, int j, int k, int l, int m, int n)
{
  const int N3 = c->size_ft, w = n + c->N * (m + c->N * l);
  double S[3][3];
  S[0][0] = c->Sh[1 * N3 + w]; S[1][1] = c->Sh[2 * N3 + w]; S[2][2] = c->Sh[3 * N3 + w];
  S[0][1] = S[1][0] = c->Sh[4 * N3 + w];
  S[0][2] = S[2][0] = c->Sh[5 * N3 + w];
  S[1][2] = S[2][1] = c->Sh[6 * N3 + w];
  double z[3] = {c->eta[i], c->eta[j], c->eta[k]}, kk[3] = {c->eta[l], c->eta[m], c->eta[n]};
  int rz = (kk[0] == 0. && kk[1] == 0. && kk[2] == 0.);
  return ghat3_from(c->gamma, S, c->Sh[w], rz, z, kk);
}
/* row xi of generate_conv_weights' table: collisionRoutines_1.cpp:220-237 */
void lpo_weight_row(const lpo_ctx *c, int xi, double *row)
{
  const int N = c->N;
  int k = xi % N, j = (xi / N) % N, i = xi / (N * N);
  for (int l = 0; l < N; l++)
    for (int m = 0; m < N; m++)
      for (int n = 0; n < N; n++) row[n + N * (m + N * l)] = weight_at(c, i, j, k, l, m, n);
}

/* ------------------------------------------------------------------------------------------ */
/* conservation tables: conservationRoutines.cpp:159-216 */
static double sinc_(double x) { return x == 0.0 ? 1.0 : sin(x) / x; }
static void invert_small(double *a, int n)
{ /* Gauss-Jordan with partial pivoting; stands in for dgetrf_/dgetri_ (:215-216) */
  double inv[25];
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) inv[i * n + j] = (i == j);
  for (int col = 0; col < n; col++) {
    int p = col;
    for (int r = col + 1; r < n; r++) if (fabs(a[r * n + col]) > fabs(a[p * n + col])) p = r;
    if (p != col) for (int j = 0; j < n; j++) {
      double t = a[col * n + j]; a[col * n + j] = a[p * n + j]; a[p * n + j] = t;
      t = inv[col * n + j]; inv[col * n + j] = inv[p * n + j]; inv[p * n + j] = t;
    }
    double d = a[col * n + col];
    for (int j = 0; j < n; j++) { a[col * n + j] /= d; inv[col * n + j] /= d; }
    for (int r = 0; r < n; r++) if (r != col) {
      double f = a[r * n + col];
      for (int j = 0; j < n; j++) { a[r * n + j] -= f * a[col * n + j]; inv[r * n + j] -= f * inv[col * n + j]; }
    }
  }
  memcpy(a, inv, sizeof(double) * n * n);
}
static void build_conservation(lpo_ctx *c)
{
  const int N = c->N, N3 = c->size_ft;
  const double L = c->Lv, *eta = c->eta;
  for (int q = 0; q < N3; q++) {
    int k = q % N, j = (q / N) % N, i = q / (N * N);
    double e[3] = {eta[i], eta[j], eta[k]}, sc[3], a[3];
    for (int d = 0; d < 3; d++) {
      sc[d] = sinc_(L * e[d]);
      a[d] = (e[d] != 0) ? ((e[d] * e[d] * L * L - 2) * sin(e[d] * L) + 2 * e[d] * L * cos(e[d] * L)) / (e[d] * e[d] * e[d] * L)
                         : L * L / 3.;
    }
    c->C5[0 * N3 + q] = sc[0] * sc[1] * sc[2];
    c->C5[4 * N3 + q] = 0.5 * (a[0] * sc[1] * sc[2] + a[1] * sc[0] * sc[2] + a[2] * sc[0] * sc[1]);
    c->C5[1 * N3 + q] = (e[0] != 0) ? -sc[1] * sc[2] * (sinc_(e[0] * L) - cos(e[0] * L)) / e[0] : 0.;
    c->C5[2 * N3 + q] = (e[1] != 0) ? -sc[0] * sc[2] * (sinc_(e[1] * L) - cos(e[1] * L)) / e[1] : 0.;
    c->C5[3 * N3 + q] = (e[2] != 0) ? -sc[0] * sc[1] * (sinc_(e[2] * L) - cos(e[2] * L)) / e[2] : 0.;
  }
  /* CCt[i][j] = sum_k C1_5[i]C1_5[j] + C2[i]C2[j] (:202-210): rows 0,4 are real-part rows, rows
   * 1..3 imaginary-part rows, and the complementary arrays are zero, so cross terms vanish. */
  for (int i = 0; i < 5; i++)
    for (int j = 0; j < 5; j++) {
      int ri = (i == 0 || i == 4), rj = (j == 0 || j == 4);
      double t = 0.;
      if (ri == rj) for (int q = 0; q < N3; q++) t += c->C5[i * N3 + q] * c->C5[j * N3 + q];
      c->CCt[i * 5 + j] = t;
    }
  invert_small(c->CCt, 5);
  /* CCt_linear: rows C1_5[0], C1_5[4] only (conservationRoutines.cpp:222-238) */
  double a = 0., b = 0., d = 0.;
  for (int q = 0; q < N3; q++) { a += c->C5[q] * c->C5[q]; b += c->C5[q] * c->C5[4 * N3 + q]; d += c->C5[4 * N3 + q] * c->C5[4 * N3 + q]; }
  c->CCt_lin[0] = a; c->CCt_lin[1] = b; c->CCt_lin[2] = b; c->CCt_lin[3] = d;
  invert_small(c->CCt_lin, 2);
  /* createCCtAndPivot_OnlyMass (conservationRoutines.cpp:311-349): M = 1, CCt = (sum C1_1^2)^-1 */
  c->CCt_mass = 1. / a;
}
void lpo_get_conservation(const lpo_ctx *c, double *C5, double *CCt25)
{
  memcpy(C5, c->C5, sizeof(double) * 5 * c->size_ft);
  memcpy(CCt25, c->CCt, sizeof(double) * 25);
}

/* ------------------------------------------------------------------------------------------ */
/* 1-D factors of IntModes (collisionRoutines_1.cpp:408-562): for Fourier node k and DG cell j,
 *   T = int_cell e^{i eta v} dv, M = int e^{i eta v}(v-c)/dv dv, S = int e^{i eta v}((v-c)/dv)^2 dv */
static void intmodes_1d(const lpo_ctx *c, int k, int j, double T[2], double Mm[2], double S[2])
{
  const double e = c->eta[k], dv = c->dv;
  const double vc = gridv(c, (double)j), vl = gridv(c, j - 0.5), vr = gridv(c, j + 0.5);
  if (e != 0.) {
    T[0] = (sin(e * vr) - sin(e * vl)) / e;
    T[1] = (cos(e * vl) - cos(e * vr)) / e;
    double a_re = (vr * sin(e * vr) - vl * sin(e * vl)) / e + (cos(e * vr) - cos(e * vl)) / e / e;
    double a_im = (sin(e * vr) - sin(e * vl)) / e / e + (vl * cos(e * vl) - vr * cos(e * vr)) / e;
    Mm[0] = (a_re - vc * T[0]) / dv;
    Mm[1] = (a_im - vc * T[1]) / dv;
    S[0] = ((vr * vr * sin(e * vr) - vl * vl * sin(e * vl) - 2 * a_im) / e - 2 * vc * a_re + vc * vc * T[0]) / dv / dv;
    S[1] = ((vl * vl * cos(e * vl) - vr * vr * cos(e * vr) + 2 * a_re) / e - 2 * vc * a_im + vc * vc * T[1]) / dv / dv;
  } else {
    T[0] = dv; T[1] = 0.; Mm[0] = 0.; Mm[1] = 0.; S[0] = dv / 12.; S[1] = 0.;
  }
}
static void cmul3(const double a[2], const double b[2], const double d[2], double out[2])
{ /* d * (a * b), the association IntModes uses (:443-444) */
  double pr = a[0] * b[0] - a[1] * b[1], pi = a[0] * b[1] + b[0] * a[1];
  out[0] = d[0] * pr - d[1] * pi;
  out[1] = d[0] * pi + d[1] * pr;
}
void lpo_IntModes(const lpo_ctx *c, int k1, int k2, int k3, int j1, int j2, int j3, double *out)
{
  double T1[2], M1[2], S1[2], T2[2], M2[2], S2[2], T3[2], M3[2], S3[2], t[2], u[2], w[2];
  intmodes_1d(c, k1, j1, T1, M1, S1);
  intmodes_1d(c, k2, j2, T2, M2, S2);
  intmodes_1d(c, k3, j3, T3, M3, S3);
  cmul3(T1, T2, T3, out + 0);
  cmul3(M1, T2, T3, out + 2);
  cmul3(T1, M2, T3, out + 4);
  cmul3(T1, T2, M3, out + 6);
  cmul3(S1, T2, T3, t);
  cmul3(T1, S2, T3, u);
  cmul3(T1, T2, S3, w);
  out[8] = t[0] + u[0] + w[0];
  out[9] = t[1] + u[1] + w[1];
}

/* ------------------------------------------------------------------------------------------ */
lpo_ctx *lpo_create(int Nx, int Nv, int N, double Lv, double Lx, double nu, double dt,
                    int homogeneous, int gamma)
{
  if (gamma != -3 && gamma != 0 && gamma != 1) return NULL; /* InputParsing.cpp:202-238 */
  lpo_ctx *c = (lpo_ctx *)calloc(1, sizeof(lpo_ctx));
  c->Nx = Nx; c->Nv = Nv; c->N = N; c->homogeneous = homogeneous; c->gamma = gamma;
  c->Lv = Lv; c->Lx = Lx; c->nu = nu; c->dt = dt;
  /* LP_ompi.cpp:169-184 */
  c->size_v = Nv * Nv * Nv;
  c->size_ft = N * N * N;
  c->ncell = homogeneous ? 1 : Nx;
  c->dv = 2. * Lv / Nv;
  c->dx = Lx / Nx;
  c->scalev = c->dv * c->dv * c->dv;
  c->scaleL = 8 * Lv * Lv * Lv;
  /* LP_ompi.cpp:359-373 */
  c->scale3 = pow(1.0 / sqrt(2.0 * M_PI), 3.0);
  c->L_eta = 0.5 * (double)(N - 1) * M_PI / Lv;
  c->h_v = 2.0 * Lv / (double)(N - 1);
  c->h_eta = 2.0 * c->L_eta / (double)N;
  c->v = (double *)malloc(sizeof(double) * N);
  c->eta = (double *)malloc(sizeof(double) * N);
  c->wt = (double *)malloc(sizeof(double) * N);
  for (int i = 0; i < N; i++) {
    c->eta[i] = -c->L_eta + (double)i * c->h_eta;
    c->v[i] = -Lv + (double)i * c->h_v;
    c->wt[i] = (i == 0 || i == N - 1) ? 0.5 : 1.0; /* SetInit_1.cpp:19-27 */
  }
  const int N3 = c->size_ft;
  c->Sh = (double *)malloc(sizeof(double) * 7 * N3);
  #pragma omp parallel for
  for (int w = 0; w < N3; w++) {
    int n = w % N, m = (w / N) % N, l = w / (N * N);
    double k1 = c->eta[l], k2 = c->eta[m], k3 = c->eta[n], S[3][3];
    double r = sqrt(k1 * k1 + k2 * k2 + k3 * k3);
    shat_matrix_g(gamma, Lv, k1, k2, k3, S);
    c->Sh[w] = (r == 0.) ? 0. : sqrt(8. / M_PI) * (Lv * r - sin(Lv * r)) / (Lv * r);
    c->Sh[1 * N3 + w] = S[0][0]; c->Sh[2 * N3 + w] = S[1][1]; c->Sh[3 * N3 + w] = S[2][2];
    c->Sh[4 * N3 + w] = S[0][1]; c->Sh[5 * N3 + w] = S[0][2]; c->Sh[6 * N3 + w] = S[1][2];
  }
  c->C5 = (double *)malloc(sizeof(double) * 5 * N3);
  build_conservation(c);
  c->T1 = (double *)malloc(sizeof(double) * 2 * N * Nv);
  c->M1 = (double *)malloc(sizeof(double) * 2 * N * Nv);
  c->S1 = (double *)malloc(sizeof(double) * 2 * N * Nv);
  for (int k = 0; k < N; k++)
    for (int j = 0; j < Nv; j++)
      intmodes_1d(c, k, j, c->T1 + 2 * (k * Nv + j), c->M1 + 2 * (k * Nv + j), c->S1 + 2 * (k * Nv + j));
  /* node -> cell map, SetInit_1.cpp:399-408 (identical expression: (int)((l*h_v)/dv), clip) */
  c->node_cell = (int *)malloc(sizeof(int) * N);
  c->node_xi = (double *)malloc(sizeof(double) * N);
  for (int l = 0; l < N; l++) {
    int j = (int)((l * c->h_v) / c->dv);
    if (j == Nv) j = Nv - 1;
    c->node_cell[l] = j;
    c->node_xi[l] = (c->v[l] - gridv(c, (double)j)) / c->dv;
  }
  return c;
}
void lpo_destroy(lpo_ctx *c)
{
  if (!c) return;
  free(c->v); free(c->eta); free(c->wt); free(c->Sh); free(c->C5);
  free(c->T1); free(c->M1); free(c->S1); free(c->node_cell); free(c->node_xi);
  free(c->dirL); free(c->dirR); free(c->mhat);
  free(c);
}
void lpo_set_direct_intmodes(lpo_ctx *c, int on) { c->direct_intmodes = on; }
void lpo_get_grids(const lpo_ctx *c, double *v, double *eta, double *wt)
{
  memcpy(v, c->v, sizeof(double) * c->N);
  memcpy(eta, c->eta, sizeof(double) * c->N);
  memcpy(wt, c->wt, sizeof(double) * c->N);
}
void lpo_set_num_threads(int n)
{
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int lpo_num_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------------------------------ */
/* setInit_spectral: SetInit_1.cpp:396-436 */
void lpo_setInit_spectral(const lpo_ctx *c, const double *U, double *f)
{
  const int N = c->N, Nv = c->Nv;
  for (int cell = 0; cell < c->ncell; cell++)
    for (int l = 0; l < N; l++)
      for (int m = 0; m < N; m++)
        for (int n = 0; n < N; n++) {
          int j1 = c->node_cell[l], j2 = c->node_cell[m], j3 = c->node_cell[n];
          size_t k = (size_t)cell * c->size_v + (j1 * Nv * Nv + j2 * Nv + j3);
          double x1 = c->node_xi[l], x2 = c->node_xi[m], x3 = c->node_xi[n];
          f[(size_t)cell * c->size_ft + l * N * N + m * N + n] =
              U[k * 6 + 0] + U[k * 6 + 2] * x1 + U[k * 6 + 3] * x2 + U[k * 6 + 4] * x3
              + U[k * 6 + 5] * (x1 * x1 + x2 * x2 + x3 * x3);
        }
}

/* ------------------------------------------------------------------------------------------ */
/* unnormalised 3-D DFT, Y[k] = sum_j X[j] exp(sign 2 pi i jk/N) per axis (FFTW semantics of
 * the plans made at LP_ompi.cpp:327-328), as three dense passes */
static void dft3(double *x, int N, int sign)
{
  double *tw = (double *)malloc(sizeof(double) * 2 * N), *tmp = (double *)malloc(sizeof(double) * 2 * N);
  for (int r = 0; r < N; r++) { tw[2 * r] = cos(sign * 2.0 * M_PI * r / N); tw[2 * r + 1] = sin(sign * 2.0 * M_PI * r / N); }
  const long strides[3] = {1, N, (long)N * N};
  for (int ax = 0; ax < 3; ax++) {
    long st = strides[ax], s_a = strides[(ax + 1) % 3], s_b = strides[(ax + 2) % 3];
    for (int a = 0; a < N; a++)
      for (int b = 0; b < N; b++) {
        double *base = x + 2 * (a * s_a + b * s_b);
        for (int k = 0; k < N; k++) {
          double sr = 0., si = 0.;
          for (int j = 0; j < N; j++) {
            int r = (int)(((long)j * k) % N);
            double xr = base[2 * j * st], xi = base[2 * j * st + 1];
            sr += xr * tw[2 * r] - xi * tw[2 * r + 1];
            si += xr * tw[2 * r + 1] + xi * tw[2 * r];
          }
          tmp[2 * k] = sr; tmp[2 * k + 1] = si;
        }
        for (int k = 0; k < N; k++) { base[2 * k * st] = tmp[2 * k]; base[2 * k * st + 1] = tmp[2 * k + 1]; }
      }
  }
  free(tw); free(tmp);
}
/* fft3D: collisionRoutines_1.cpp:285-319 */
void lpo_fft3D(const lpo_ctx *c, const double *in, double *out)
{
  const int N = c->N;
  double *t = (double *)malloc(sizeof(double) * 2 * c->size_ft);
  const double hv3 = c->scale3 * c->h_v * c->h_v * c->h_v;
  for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) for (int k = 0; k < N; k++) {
    int q = k + N * (j + N * i);
    double s = ((double)i + (double)j + (double)k) * c->L_eta * c->h_v, w = hv3 * c->wt[i] * c->wt[j] * c->wt[k];
    t[2 * q] = w * (cos(s) * in[2 * q] - sin(s) * in[2 * q + 1]);
    t[2 * q + 1] = w * (cos(s) * in[2 * q + 1] + sin(s) * in[2 * q]);
  }
  dft3(t, N, -1);
  for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) for (int k = 0; k < N; k++) {
    int q = k + N * (j + N * i);
    double s = c->Lv * (c->eta[i] + c->eta[j] + c->eta[k]);
    out[2 * q] = cos(s) * t[2 * q] - sin(s) * t[2 * q + 1];
    out[2 * q + 1] = cos(s) * t[2 * q + 1] + sin(s) * t[2 * q];
  }
  free(t);
}
/* FS: collisionRoutines_1.cpp:363-398 */
void lpo_FS(const lpo_ctx *c, const double *in, double *out)
{
  const int N = c->N;
  double *t = (double *)malloc(sizeof(double) * 2 * c->size_ft);
  for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) for (int k = 0; k < N; k++) {
    int q = k + N * (j + N * i);
    double s = -(((double)i + (double)j + (double)k) * c->Lv * c->h_eta);
    t[2 * q] = cos(s) * in[2 * q] - sin(s) * in[2 * q + 1];
    t[2 * q + 1] = cos(s) * in[2 * q + 1] + sin(s) * in[2 * q];
  }
  dft3(t, N, +1);
  for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) for (int k = 0; k < N; k++) {
    int q = k + N * (j + N * i);
    double s = -(c->L_eta * (c->v[i] + c->v[j] + c->v[k]));
    out[2 * q] = (cos(s) * t[2 * q] - sin(s) * t[2 * q + 1]) / c->scaleL / c->scale3;
    out[2 * q + 1] = (cos(s) * t[2 * q + 1] + sin(s) * t[2 * q]) / c->scaleL / c->scale3;
  }
  free(t);
}

/* ------------------------------------------------------------------------------------------ */
/* ComputeQ: collisionRoutines_1.cpp:691-774 (same window logic, same omega order) */
static void window(int N, int i, int *s, int *e)
{
  if (i < N / 2) { *s = 0; *e = i + N / 2 + 1; } else { *s = i - N / 2 + 1; *e = N; }
}
void lpo_ComputeQ(const lpo_ctx *c, const double *f, double *qHat)
{
  const int N = c->N, N3 = c->size_ft;
  double *in = (double *)malloc(sizeof(double) * 2 * N3), *fh = (double *)malloc(sizeof(double) * 2 * N3);
  for (int q = 0; q < N3; q++) { in[2 * q] = f[q]; in[2 * q + 1] = 0.; }
  lpo_fft3D(c, in, fh);
  const double pref = c->h_eta * c->h_eta * c->h_eta, *wt = c->wt;
  #pragma omp parallel for collapse(2) schedule(dynamic)
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++)
      for (int k = 0; k < N; k++) {
        int si, ei, sj, ej, sk, ek;
        window(N, i, &si, &ei); window(N, j, &sj, &ej); window(N, k, &sk, &ek);
        double t0 = 0., t1 = 0.;
        for (int l = si; l < ei; l++)
          for (int m = sj; m < ej; m++)
            for (int n = sk; n < ek; n++) {
              int x = i + N / 2 - l, y = j + N / 2 - m, z = k + N / 2 - n;
              int a = n + N * (m + N * l), b = z + N * (y + N * x);
              double W = weight_at(c, i, j, k, l, m, n);
              t0 += pref * wt[l] * wt[m] * wt[n] * W * (fh[2 * a] * fh[2 * b] - fh[2 * a + 1] * fh[2 * b + 1]);
              t1 += pref * wt[l] * wt[m] * wt[n] * W * (fh[2 * a] * fh[2 * b + 1] + fh[2 * a + 1] * fh[2 * b]);
            }
        int q = k + N * (j + N * i);
        qHat[2 * q] = t0; qHat[2 * q + 1] = t1;
      }
  free(in); free(fh);
}

/* ComputeQLinear (collisionRoutines_1.cpp:1185-1269): Q(f, M) -- the first factor of every pair is the stored
 * transform of the initial Maxwellian, the second the transform of f */
void lpo_ComputeQLinear(const lpo_ctx *c, const double *f, const double *mh, double *qHat)
{
  const int N = c->N, N3 = c->size_ft;
  double *in = (double *)malloc(sizeof(double) * 2 * N3), *fh = (double *)malloc(sizeof(double) * 2 * N3);
  for (int q = 0; q < N3; q++) { in[2 * q] = f[q]; in[2 * q + 1] = 0.; }
  lpo_fft3D(c, in, fh);
  const double pref = c->h_eta * c->h_eta * c->h_eta, *wt = c->wt;
  #pragma omp parallel for collapse(2) schedule(dynamic)
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++)
      for (int k = 0; k < N; k++) {
        int si, ei, sj, ej, sk, ek;
        window(N, i, &si, &ei); window(N, j, &sj, &ej); window(N, k, &sk, &ek);
        double t0 = 0., t1 = 0.;
        for (int l = si; l < ei; l++)
          for (int m = sj; m < ej; m++)
            for (int n = sk; n < ek; n++) {
              int x = i + N / 2 - l, y = j + N / 2 - m, z = k + N / 2 - n;
              int a = n + N * (m + N * l), b = z + N * (y + N * x);
              double W = weight_at(c, i, j, k, l, m, n);
              t0 += pref * wt[l] * wt[m] * wt[n] * (W * (mh[2 * a] * fh[2 * b] - mh[2 * a + 1] * fh[2 * b + 1]));
              t1 += pref * wt[l] * wt[m] * wt[n] * (W * (mh[2 * a] * fh[2 * b + 1] + mh[2 * a + 1] * fh[2 * b]));
            }
        int q = k + N * (j + N * i);
        qHat[2 * q] = t0; qHat[2 * q + 1] = t1;
      }
  free(in); free(fh);
}
/* the operator the time loop applies to cell `cell` (LP_ompi.cpp:692-703) */
static void compute_q_cell(const lpo_ctx *c, const double *f, int cell, double *qHat)
{
  if (c->linear) lpo_ComputeQLinear(c, f, c->mhat + (size_t)cell * 2 * c->size_ft, qHat);
  else lpo_ComputeQ(c, f, qHat);
}
/* ComputeDFTofMaxwellian (collisionRoutines_1.cpp:1169-1183): the transform of the state held in U when it is called
 * (the initial condition), per cell.  The reference calls fft3D inside its fill loop; the last call wins. */
void lpo_set_linear_landau(lpo_ctx *c, const double *U)
{
  const int N3 = c->size_ft;
  free(c->mhat); c->mhat = NULL; c->linear = 0;
  if (!U) return;
  double *f = (double *)malloc(sizeof(double) * (size_t)c->ncell * N3), *in = (double *)malloc(sizeof(double) * 2 * N3);
  c->mhat = (double *)malloc(sizeof(double) * (size_t)c->ncell * 2 * N3);
  lpo_setInit_spectral(c, U, f);
  for (int cell = 0; cell < c->ncell; cell++) {
    for (int q = 0; q < N3; q++) { in[2 * q] = f[(size_t)cell * N3 + q]; in[2 * q + 1] = 0.; }
    lpo_fft3D(c, in, c->mhat + (size_t)cell * 2 * N3);
  }
  free(f); free(in);
  c->linear = 1;
}
void lpo_set_mass_cons_only(lpo_ctx *c, int on) { c->mass_only = on; }

/* conserveAllMoments_Normal + solveWithCCt: conservationRoutines.cpp:131-156, 32-58 */
void lpo_conserveMoments(const lpo_ctx *c, double *qHat)
{
  const int N3 = c->size_ft;
  if (c->mass_only) {
    /* conserveMass_Normal: conservationRoutines.cpp:290-309 (C1_1 = C1_5[0]) */
    double t = 0.;
    for (int q = 0; q < N3; q++) t += qHat[2 * q] * c->C5[q];
    const double lam0 = c->CCt_mass * t;
    for (int q = 0; q < N3; q++) qHat[2 * q] -= (c->C5[q] * lam0);
    return;
  }
  double lam[5] = {0, 0, 0, 0, 0}, b[5];
  for (int q = 0; q < N3; q++) {
    lam[0] += qHat[2 * q] * c->C5[0 * N3 + q];
    lam[1] += qHat[2 * q + 1] * c->C5[1 * N3 + q];
    lam[2] += qHat[2 * q + 1] * c->C5[2 * N3 + q];
    lam[3] += qHat[2 * q + 1] * c->C5[3 * N3 + q];
    lam[4] += qHat[2 * q] * c->C5[4 * N3 + q];
  }
  for (int i = 0; i < 5; i++) { b[i] = 0.; for (int j = 0; j < 5; j++) b[i] += c->CCt[j + i * 5] * lam[j]; }
  for (int q = 0; q < N3; q++) {
    qHat[2 * q] -= (c->C5[0 * N3 + q] * b[0] + c->C5[4 * N3 + q] * b[4]);
    qHat[2 * q + 1] -= (c->C5[1 * N3 + q] * b[1] + c->C5[2 * N3 + q] * b[2] + c->C5[3 * N3 + q] * b[3]);
  }
}

/* ------------------------------------------------------------------------------------------ */
/* Fourier -> DG projection of Qc[k] (already = nu*(q0/2 + (q1+q2+q3)/6)): the kt loop of
 * RK4_Inhomo/RK4_Homo, collisionRoutines_1.cpp:946-984 / 1128-1166 */
static void project_direct(const lpo_ctx *c, const double *Qc, double *tp /* 5*Nv^3 */)
{
  const int N = c->N, Nv = c->Nv;
  #pragma omp parallel for schedule(dynamic)
  for (int kt = 0; kt < c->size_v; kt++) {
    int j3 = kt % Nv, j2 = (kt / Nv) % Nv, j1 = kt / (Nv * Nv);
    double t[5] = {0, 0, 0, 0, 0}, IM[10];
    for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) for (int k = 0; k < N; k++) {
      int q = k + N * (j + N * i);
      lpo_IntModes(c, i, j, k, j1, j2, j3, IM);
      for (int l = 0; l < 5; l++) t[l] += IM[2 * l] * Qc[2 * q] - IM[2 * l + 1] * Qc[2 * q + 1];
    }
    for (int l = 0; l < 5; l++) tp[5 * kt + l] = t[l];
  }
}
/* y[a,b,j] = sum_k tab[k,j] * x[a,b,k] along the last axis of an [na][nb][N] complex array */
static void contract_last(const double *tab, int N, int Nv, const double *x, long nab, double *y)
{
  #pragma omp parallel for
  for (long ab = 0; ab < nab; ab++)
    for (int j = 0; j < Nv; j++) {
      double sr = 0., si = 0.;
      for (int k = 0; k < N; k++) {
        double tr = tab[2 * (k * Nv + j)], ti = tab[2 * (k * Nv + j) + 1];
        double xr = x[2 * (ab * N + k)], xi = x[2 * (ab * N + k) + 1];
        sr += tr * xr - ti * xi; si += tr * xi + ti * xr;
      }
      y[2 * (ab * Nv + j)] = sr; y[2 * (ab * Nv + j) + 1] = si;
    }
}
/* contract the middle axis: x[a][k][j3] -> y[a][j2][j3] */
static void contract_mid(const double *tab, int N, int Nv, const double *x, int na, int nlast, double *y, int accumulate)
{
  #pragma omp parallel for collapse(2)
  for (int a = 0; a < na; a++)
    for (int j = 0; j < Nv; j++)
      for (int z = 0; z < nlast; z++) {
        double sr = 0., si = 0.;
        for (int k = 0; k < N; k++) {
          double tr = tab[2 * (k * Nv + j)], ti = tab[2 * (k * Nv + j) + 1];
          const double *xx = x + 2 * (((long)a * N + k) * nlast + z);
          sr += tr * xx[0] - ti * xx[1]; si += tr * xx[1] + ti * xx[0];
        }
        double *yy = y + 2 * (((long)a * Nv + j) * nlast + z);
        if (accumulate) { yy[0] += sr; yy[1] += si; } else { yy[0] = sr; yy[1] = si; }
      }
}
static void project_separable(const lpo_ctx *c, const double *Qc, double *tp)
{
  const int N = c->N, Nv = c->Nv;
  const long nA = (long)N * N * Nv, nB = (long)N * Nv * Nv;
  double *AT = (double *)malloc(sizeof(double) * 2 * nA), *AM = (double *)malloc(sizeof(double) * 2 * nA),
         *AS = (double *)malloc(sizeof(double) * 2 * nA);
  double *BTT = (double *)malloc(sizeof(double) * 2 * nB), *BMT = (double *)malloc(sizeof(double) * 2 * nB),
         *BTM = (double *)malloc(sizeof(double) * 2 * nB), *BS = (double *)malloc(sizeof(double) * 2 * nB);
  contract_last(c->T1, N, Nv, Qc, (long)N * N, AT);
  contract_last(c->M1, N, Nv, Qc, (long)N * N, AM);
  contract_last(c->S1, N, Nv, Qc, (long)N * N, AS);
  contract_mid(c->T1, N, Nv, AT, N, Nv, BTT, 0);
  contract_mid(c->M1, N, Nv, AT, N, Nv, BMT, 0);
  contract_mid(c->T1, N, Nv, AM, N, Nv, BTM, 0);
  contract_mid(c->S1, N, Nv, AT, N, Nv, BS, 0);
  contract_mid(c->T1, N, Nv, AS, N, Nv, BS, 1);
  const int P = Nv * Nv;
  #pragma omp parallel for
  for (int j1 = 0; j1 < Nv; j1++)
    for (int p = 0; p < P; p++) {
      double t[5] = {0, 0, 0, 0, 0};
      for (int k = 0; k < N; k++) {
        const double *T = c->T1 + 2 * (k * Nv + j1), *Mm = c->M1 + 2 * (k * Nv + j1), *S = c->S1 + 2 * (k * Nv + j1);
        const double *btt = BTT + 2 * ((long)k * P + p), *bmt = BMT + 2 * ((long)k * P + p),
                     *btm = BTM + 2 * ((long)k * P + p), *bs = BS + 2 * ((long)k * P + p);
        t[0] += T[0] * btt[0] - T[1] * btt[1];
        t[1] += Mm[0] * btt[0] - Mm[1] * btt[1];
        t[2] += T[0] * bmt[0] - T[1] * bmt[1];
        t[3] += T[0] * btm[0] - T[1] * btm[1];
        t[4] += S[0] * btt[0] - S[1] * btt[1] + T[0] * bs[0] - T[1] * bs[1];
      }
      for (int l = 0; l < 5; l++) tp[5 * (j1 * P + p) + l] = t[l];
    }
  free(AT); free(AM); free(AS); free(BTT); free(BMT); free(BTM); free(BS);
}

/* RK4_Inhomo / RK4_Homo: collisionRoutines_1.cpp:903-985 / 1087-1167.  The third stage has no
 * dt factor in the reference (:940, :1122) -- reproduced. */
void lpo_RK4(const lpo_ctx *c, const double *f, int cell, const double *qHat, const double *U,
             double *dU, double *q123)
{
  const int N3 = c->size_ft;
  const double dt = c->dt, nu = c->nu;
  double *out = (double *)malloc(sizeof(double) * 2 * N3);
  double *Q = (double *)malloc(sizeof(double) * N3), *Q1 = (double *)malloc(sizeof(double) * N3),
         *f1 = (double *)malloc(sizeof(double) * N3);
  double *q1 = (double *)malloc(sizeof(double) * 2 * N3), *q2 = (double *)malloc(sizeof(double) * 2 * N3),
         *q3 = (double *)malloc(sizeof(double) * 2 * N3);
  lpo_FS(c, qHat, out);
  for (int i = 0; i < N3; i++) { Q[i] = out[2 * i]; f1[i] = f[i] + dt * Q[i] * nu; }
  compute_q_cell(c, f1, cell, q1); lpo_conserveMoments(c, q1);   /* RK4Linear (:1272-1350) has the same stage logic */
  lpo_FS(c, q1, out);
  for (int i = 0; i < N3; i++) { Q1[i] = out[2 * i]; f1[i] = f[i] + 0.5 * dt * Q[i] * nu + 0.5 * dt * Q1[i] * nu; }
  compute_q_cell(c, f1, cell, q2); lpo_conserveMoments(c, q2);
  lpo_FS(c, q2, out);
  for (int i = 0; i < N3; i++) { Q1[i] = out[2 * i]; f1[i] = f[i] + 0.5 * Q[i] * nu + 0.5 * Q1[i] * nu; }
  compute_q_cell(c, f1, cell, q3); lpo_conserveMoments(c, q3);
  if (q123) {
    memcpy(q123, q1, sizeof(double) * 2 * N3);
    memcpy(q123 + 2 * N3, q2, sizeof(double) * 2 * N3);
    memcpy(q123 + 4 * N3, q3, sizeof(double) * 2 * N3);
  }
  double *Qc = out; /* nu*(q0/2 + (q1+q2+q3)/6), :957-958 */
  for (int i = 0; i < 2 * N3; i++) Qc[i] = nu * (0.5 * qHat[i] + (q1[i] + q2[i] + q3[i]) / 6.);
  double *tp = (double *)malloc(sizeof(double) * 5 * c->size_v);
  if (c->direct_intmodes) project_direct(c, Qc, tp); else project_separable(c, Qc, tp);
  const double sc = c->scalev, sL = c->scaleL, s3 = c->scale3;
  for (int kt = 0; kt < c->size_v; kt++) {
    size_t kv = (size_t)cell * c->size_v + kt;
    double t0 = U[kv * 6 + 0] + U[kv * 6 + 5] / 4. + dt * tp[5 * kt + 0] / sc / sL / s3;
    double t2 = U[kv * 6 + 2] + dt * tp[5 * kt + 1] * 12. / sc / sL / s3;
    double t3 = U[kv * 6 + 3] + dt * tp[5 * kt + 2] * 12. / sc / sL / s3;
    double t4 = U[kv * 6 + 4] + dt * tp[5 * kt + 3] * 12. / sc / sL / s3;
    double t5 = U[kv * 6 + 0] / 4. + U[kv * 6 + 5] * 19. / 240. + dt * tp[5 * kt + 4] / sc / sL / s3;
    dU[5 * kt + 0] = 19 * t0 / 4. - 15 * t5;
    dU[5 * kt + 4] = 60 * t5 - 15 * t0;
    dU[5 * kt + 1] = t2; dU[5 * kt + 2] = t3; dU[5 * kt + 3] = t4;
  }
  free(out); free(Q); free(Q1); free(f1); free(q1); free(q2); free(q3); free(tp);
}

/* ------------------------------------------------------------------------------------------ */
/* FullandLinear variant (reference test 3): ComputeQ_FandL collisionRoutines_1.cpp:605-689,
 * gHat3_linear :193-218, conserveAllMoments_FandL conservationRoutines.cpp:102-129,
 * RK4_FandL_Inhomo/_Homo collisionRoutines_1.cpp:800-901 / 987-1085. */
void lpo_set_fandl(lpo_ctx *c, int on) { c->fandl = on; }
static double weight_lin_at(const lpo_ctx *c, int i, int j, int k, int l, int m, int n)
{
  const int N3 = c->size_ft, w = n + c->N * (m + c->N * l);
  double S[3][3];
  S[0][0] = c->Sh[1 * N3 + w]; S[1][1] = c->Sh[2 * N3 + w]; S[2][2] = c->Sh[3 * N3 + w];
  S[0][1] = S[1][0] = c->Sh[4 * N3 + w];
  S[0][2] = S[2][0] = c->Sh[5 * N3 + w];
  S[1][2] = S[2][1] = c->Sh[6 * N3 + w];
  double z[3] = {c->eta[i], c->eta[j], c->eta[k]}, kk[3] = {c->eta[l], c->eta[m], c->eta[n]};
  double res = 0.;
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++) res += S[a][b] * z[a] * (z[b] - kk[b]);
  return -res;
}
void lpo_ComputeQ_FandL(const lpo_ctx *c, const double *f, double *qHat, double *qLin)
{
  const int N = c->N, N3 = c->size_ft;
  double *in = (double *)malloc(sizeof(double) * 2 * N3), *fh = (double *)malloc(sizeof(double) * 2 * N3);
  for (int q = 0; q < N3; q++) { in[2 * q] = f[q]; in[2 * q + 1] = 0.; }
  lpo_fft3D(c, in, fh);
  const double pref = c->h_eta * c->h_eta * c->h_eta, *wt = c->wt, s3 = c->scale3;
  #pragma omp parallel for collapse(2) schedule(dynamic)
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++)
      for (int k = 0; k < N; k++) {
        int si, ei, sj, ej, sk, ek;
        window(N, i, &si, &ei); window(N, j, &sj, &ej); window(N, k, &sk, &ek);
        double t0 = 0., t1 = 0., t01 = 0., t11 = 0.;
        for (int l = si; l < ei; l++)
          for (int m = sj; m < ej; m++)
            for (int n = sk; n < ek; n++) {
              int x = i + N / 2 - l, y = j + N / 2 - m, z = k + N / 2 - n;
              int a = n + N * (m + N * l), b = z + N * (y + N * x);
              double W = weight_at(c, i, j, k, l, m, n), W1 = weight_lin_at(c, i, j, k, l, m, n);
              double pw = pref * wt[l] * wt[m] * wt[n];
              t0 += pw * (W * (fh[2 * a] * fh[2 * b] - fh[2 * a + 1] * fh[2 * b + 1]) + s3 * W1 * fh[2 * b]);
              t1 += pw * (W * (fh[2 * a] * fh[2 * b + 1] + fh[2 * a + 1] * fh[2 * b]) + s3 * W1 * fh[2 * b + 1]);
              t01 += pw * s3 * W1 * fh[2 * b];
              t11 += pw * s3 * W1 * fh[2 * b + 1];
            }
        int q = k + N * (j + N * i);
        qHat[2 * q] = t0; qHat[2 * q + 1] = t1;
        qLin[2 * q] = t01; qLin[2 * q + 1] = t11;
      }
  free(in); free(fh);
}
void lpo_conserveMoments_FandL(const lpo_ctx *c, double *qHat, double *qLin)
{
  const int N3 = c->size_ft;
  double tp[2] = {0., 0.}, b[2];
  for (int q = 0; q < N3; q++) { tp[0] += qLin[2 * q] * c->C5[q]; tp[1] += qLin[2 * q] * c->C5[4 * N3 + q]; }
  lpo_conserveMoments(c, qHat);
  for (int i = 0; i < 2; i++) { b[i] = 0.; for (int j = 0; j < 2; j++) b[i] += c->CCt_lin[j + i * 2] * tp[j]; }
  for (int q = 0; q < N3; q++) qLin[2 * q] -= (c->C5[q] * b[0] + c->C5[4 * N3 + q] * b[1]);
}
static void project_update(const lpo_ctx *c, const double *Qc, int cell, const double *U, double *dU)
{
  const double dt = c->dt, sc = c->scalev, sL = c->scaleL, s3 = c->scale3;
  double *tp = (double *)malloc(sizeof(double) * 5 * c->size_v);
  if (c->direct_intmodes) project_direct(c, Qc, tp); else project_separable(c, Qc, tp);
  for (int kt = 0; kt < c->size_v; kt++) {
    size_t kv = (size_t)cell * c->size_v + kt;
    double t0 = U[kv * 6 + 0] + U[kv * 6 + 5] / 4. + dt * tp[5 * kt + 0] / sc / sL / s3;
    double t2 = U[kv * 6 + 2] + dt * tp[5 * kt + 1] * 12. / sc / sL / s3;
    double t3 = U[kv * 6 + 3] + dt * tp[5 * kt + 2] * 12. / sc / sL / s3;
    double t4 = U[kv * 6 + 4] + dt * tp[5 * kt + 3] * 12. / sc / sL / s3;
    double t5 = U[kv * 6 + 0] / 4. + U[kv * 6 + 5] * 19. / 240. + dt * tp[5 * kt + 4] / sc / sL / s3;
    dU[5 * kt + 0] = 19 * t0 / 4. - 15 * t5;
    dU[5 * kt + 4] = 60 * t5 - 15 * t0;
    dU[5 * kt + 1] = t2; dU[5 * kt + 2] = t3; dU[5 * kt + 3] = t4;
  }
  free(tp);
}
/* qHat, qLin: conserved first-stage spectra; qHat is overwritten by qHat + qLin as in the reference */
void lpo_RK4_FandL(const lpo_ctx *c, const double *f, int cell, double *qHat, const double *qLin, const double *U, double *dU)
{
  const int N3 = c->size_ft;
  const double dt = c->dt, nu = c->nu;
  double *out = (double *)malloc(sizeof(double) * 2 * N3);
  double *Q = (double *)malloc(sizeof(double) * N3), *Q1 = (double *)malloc(sizeof(double) * N3), *f1 = (double *)malloc(sizeof(double) * N3);
  double *q[3], *ql = (double *)malloc(sizeof(double) * 2 * N3);
  for (int s = 0; s < 3; s++) q[s] = (double *)malloc(sizeof(double) * 2 * N3);
  for (int i = 0; i < 2 * N3; i++) qHat[i] += qLin[i];
  lpo_FS(c, qHat, out);
  for (int i = 0; i < N3; i++) { Q[i] = out[2 * i]; f1[i] = f[i] + dt * Q[i] * nu; }
  for (int s = 0; s < 3; s++) {
    lpo_ComputeQ_FandL(c, f1, q[s], ql);
    lpo_conserveMoments_FandL(c, q[s], ql);
    for (int i = 0; i < 2 * N3; i++) q[s][i] += ql[i];
    if (s < 2) {
      lpo_FS(c, q[s], out);
      /* both later stage vectors carry dt here (:842, :858), unlike RK4_Inhomo's third stage */
      for (int i = 0; i < N3; i++) { Q1[i] = out[2 * i]; f1[i] = f[i] + 0.5 * dt * Q[i] * nu + 0.5 * dt * Q1[i] * nu; }
    }
  }
  for (int i = 0; i < 2 * N3; i++) out[i] = nu * (0.5 * qHat[i] + (q[0][i] + q[1][i] + q[2][i]) / 6.);
  project_update(c, out, cell, U, dU);
  free(out); free(Q); free(Q1); free(f1); free(ql);
  for (int s = 0; s < 3; s++) free(q[s]);
}

/* collision branch of the time loop incl. the scatter into U: LP_ompi.cpp:669-754 */
void lpo_collide_step(const lpo_ctx *c, double *U)
{
  const int N3 = c->size_ft, sv = c->size_v;
  double *f = (double *)malloc(sizeof(double) * (size_t)c->ncell * N3);
  double *q = (double *)malloc(sizeof(double) * 2 * N3);
  double *dU = (double *)malloc(sizeof(double) * 5 * (size_t)c->ncell * sv);
  lpo_setInit_spectral(c, U, f);
  for (int cell = 0; cell < c->ncell; cell++) {
    if (c->fandl) {
      double *ql = (double *)malloc(sizeof(double) * 2 * N3);
      lpo_ComputeQ_FandL(c, f + (size_t)cell * N3, q, ql);
      lpo_conserveMoments_FandL(c, q, ql);
      lpo_RK4_FandL(c, f + (size_t)cell * N3, cell, q, ql, U, dU + 5 * (size_t)cell * sv);
      free(ql);
      continue;
    }
    compute_q_cell(c, f + (size_t)cell * N3, cell, q);
    lpo_conserveMoments(c, q);
    lpo_RK4(c, f + (size_t)cell * N3, cell, q, U, dU + 5 * (size_t)cell * sv, NULL);
  }
  for (size_t kv = 0; kv < (size_t)c->ncell * sv; kv++) {
    U[kv * 6 + 0] = dU[kv * 5]; U[kv * 6 + 5] = dU[kv * 5 + 4];
    U[kv * 6 + 2] = dU[kv * 5 + 1]; U[kv * 6 + 3] = dU[kv * 5 + 2]; U[kv * 6 + 4] = dU[kv * 5 + 3];
  }
  free(f); free(q); free(dU);
}

/* ------------------------------------------------------------------------------------------ */
/* Field integrals, periodic ("Normal") path.  The reference recomputes nested sums over the
 * whole mesh (FieldCalculations.cpp:223-243 computePhi_x_0_Normal, :60-73 computeC_rho,
 * :356-409 Int_E/E1st/E2nd_Normal); they collapse to per-x-cell sums
 *   m_i = scalev * sum_j (U0 + U5/4),  s_i = scalev * sum_j U1,  P_i = sum_{q<i} m_q. */
typedef struct { double ce, *cp, *iE, *iE1, *iE2, *m, *s; } field_t;
static void field_alloc(field_t *F, int Nx)
{
  F->cp = (double *)malloc(sizeof(double) * 6 * Nx);
  F->iE = F->cp + Nx; F->iE1 = F->cp + 2 * Nx; F->iE2 = F->cp + 3 * Nx; F->m = F->cp + 4 * Nx; F->s = F->cp + 5 * Nx;
}
static void field_compute(const lpo_ctx *c, const double *U, field_t *F)
{
  const int Nx = c->Nx, sv = c->size_v;
  const double dx = c->dx, Lx = c->Lx;
  #pragma omp parallel for
  for (int i = 0; i < Nx; i++) {
    double a = 0., b = 0.;
    for (int j = 0; j < sv; j++) {
      size_t k = (size_t)i * sv + j;
      a += U[k * 6 + 0] + U[k * 6 + 5] / 4.;
      b += U[k * 6 + 1];
    }
    F->m[i] = a * c->scalev; F->s[i] = b * c->scalev;
  }
  double P = 0., acc = 0.;
  for (int i = 0; i < Nx; i++) { F->cp[i] = dx * P; acc += P + 0.5 * F->m[i] - F->s[i] / 12.; P += F->m[i]; }
  if (c->doping) {
    /* computePhi_x_0_Doping, Int_E_Doping, Int_E1st_Doping, Int_E2nd_Doping: FieldCalculations.cpp:427-450, 585-676 */
    const double NL = c->NL, NH = c->NH, eps = c->eps, a_val = (c->a_i + 1) * dx, b_val = (c->b_i + 1) * dx, Phi_Lx = 1;
    const double tmp = acc * dx * dx;
    F->ce = Phi_Lx / Lx + 0.5 * NH * Lx / eps + (NL - NH) * (b_val - a_val) / eps - (0.5 * (NL - NH) * (b_val * b_val - a_val * a_val) + tmp) / (Lx * eps);
    P = 0.;
    for (int i = 0; i < Nx; i++) {
      const double ND = (i <= c->a_i || i > c->b_i) ? NH : NL, xi = gridx(c, (double)i), c2 = F->s[i] * dx / 2.;
      double r = -(P + 0.5 * F->m[i] - F->s[i] / 12.) * dx * dx + ND * xi * dx;
      if (i > c->a_i) r += (NH - NL) * a_val * dx;
      if (i > c->b_i) r += (NL - NH) * b_val * dx;
      F->iE[i] = r / eps - F->ce * dx;
      F->iE1[i] = (ND - F->m[i]) * dx * dx / (12. * eps);
      r = (-F->cp[i] + (F->m[i] * gridx(c, i - 0.5) + 0.25 * c2)) * dx / 12. + (ND - F->m[i]) * dx * xi / 12. - c2 * dx / 80.;
      if (i > c->a_i) r += (NH - NL) * a_val * dx / 12.;
      if (i > c->b_i) r += (NL - NH) * b_val * dx / 12.;
      F->iE2[i] = r / eps - F->ce * dx / 12.;
      P += F->m[i];
    }
    return;
  }
  F->ce = 0.5 * Lx - acc * dx * dx / Lx;
  P = 0.;
  for (int i = 0; i < Nx; i++) {
    double xi = gridx(c, (double)i), c2 = F->s[i] * dx / 2.;
    F->iE[i] = -F->ce * dx - (P + 0.5 * F->m[i] - F->s[i] / 12.) * dx * dx + xi * dx;
    F->iE1[i] = (1 - F->m[i]) * dx * dx / 12.;
    F->iE2[i] = (-F->cp[i] - F->ce + (F->m[i] * gridx(c, i - 0.5) + 0.25 * c2)) * dx / 12.
                + (1 - F->m[i]) * dx * xi / 12. - c2 * dx / 80.;
    P += F->m[i];
  }
}
void lpo_field(const lpo_ctx *c, const double *U, double *out)
{
  field_t F; field_alloc(&F, c->Nx);
  field_compute(c, U, &F);
  out[0] = F.ce;
  memcpy(out + 1, F.cp, sizeof(double) * 4 * c->Nx);
  free(F.cp);
}

/* DG right-hand side for one (x,v) cell: I1 - I2 - I3 + I5 then the mass-matrix inverse,
 * advection_1.cpp:72-103 (I1, I2), :284-321 (I3_Normal), :323-390 (I5), :440-452 (H). */
static void dg_rhs(const lpo_ctx *c, const double *U, const field_t *F, size_t k, double H[6])
{
  const int Nv = c->Nv, sv = c->size_v, Nx = c->Nx;
  const double dv = c->dv, dv2 = dv * dv, dv3 = dv2 * dv;
  const int jm = (int)(k % sv), i = (int)(k / sv);
  const int j1 = jm / (Nv * Nv);
  const double c1 = gridv(c, (double)j1), E = F->iE[i], E1 = F->iE1[i], E2 = F->iE2[i];
  const double *u = U + 6 * k;
  double tp[6] = {0, 0, 0, 0, 0, 0};
  /* I1 */
  tp[1] += dv3 * (c1 * u[0] + dv * u[2] / 12. + u[5] * c1 / 4.);
  /* I2 (Int_fE: FieldCalculations.cpp:126-135) */
  tp[2] -= ((u[0] + u[5] / 4.) * E + u[1] * E1) * c->scalev / dv;
  tp[5] -= u[2] * dv2 * E / 6.;
  /* I3: upwind in x on the sign of the v1 cell index */
  {
    const double *R, *L; double ur, ul;
    if (j1 < Nv / 2) {
      int ir = i + 1;
      if (ir == Nx && c->doping) R = c->dirR + 6 * (size_t)jm;       /* I3_Doping: DirichletBC at the right wall (advection_1.cpp:230-233) */
      else { if (ir == Nx) ir = 0; R = U + 6 * ((size_t)ir * sv + jm); }
      L = u; ur = -R[1]; ul = -L[1];
    } else {
      int il = i - 1;
      if (il == -1 && c->doping) L = c->dirL + 6 * (size_t)jm;       /* DirichletBC at the left wall (:254-257) */
      else { if (il == -1) il = Nx - 1; L = U + 6 * ((size_t)il * sv + jm); }
      R = u; ur = R[1]; ul = L[1];
    }
    tp[0] -= dv3 * ((R[0] + 0.5 * ur - L[0] - 0.5 * ul) * c1 + (R[2] - L[2]) * dv / 12. + (R[5] - L[5]) * c1 / 4.);
    tp[1] -= 0.5 * dv3 * ((R[0] + 0.5 * ur + L[0] + 0.5 * ul) * c1 + (R[2] + L[2]) * dv / 12. + (R[5] + L[5]) * c1 / 4.);
    tp[2] -= dv2 * (((R[0] - L[0]) * dv2 + (ur - ul) * 0.5 * dv2 + (R[2] - L[2]) * dv * c1) / 12. + (R[5] - L[5]) * dv2 * 19. / 720.);
    tp[3] -= (R[3] - L[3]) * c1 * dv3 / 12.;
    tp[4] -= (R[4] - L[4]) * c1 * dv3 / 12.;
    tp[5] -= dv3 * ((R[0] + 0.5 * ur - L[0] - 0.5 * ul) * c1 / 4. + (R[2] - L[2]) * dv * 19. / 720. + (R[5] - L[5]) * c1 * 19. / 240.);
  }
  /* I5: upwind in v1 on the sign of the cell-integrated field; zero flux through |v1| = Lv */
  {
    static const double zero6[6] = {0, 0, 0, 0, 0, 0};
    const double *R, *L; double ur, ul;
    if (E > 0) {
      L = u; ul = -L[2];
      if (j1 + 1 < Nv) { R = u + 6 * (size_t)(Nv * Nv); ur = -R[2]; } else { R = zero6; ur = 0.; }
    } else {
      R = u; ur = R[2];
      if (j1 - 1 > -1) { L = u - 6 * (size_t)(Nv * Nv); ul = L[2]; } else { L = zero6; ul = 0.; }
    }
    double gR = R[0] + 0.5 * ur + R[5] * 5. / 12., gL = L[0] + 0.5 * ul + L[5] * 5. / 12.;
    tp[0] += dv2 * (gR - gL) * E + dv2 * (R[1] - L[1]) * E1;
    tp[1] += dv2 * ((gR - gL) * E1 + (R[1] - L[1]) * E2);
    tp[2] += 0.5 * (dv2 * (gR + gL) * E + dv2 * (R[1] + L[1]) * E1);
    tp[3] += (R[3] - L[3]) * E * dv2 / 12.;
    tp[4] += (R[4] - L[4]) * E * dv2 / 12.;
    tp[5] += dv2 * (((R[0] + 0.5 * ur - L[0] - 0.5 * ul) * 5. / 12. + (R[5] - L[5]) * 133. / 720.) * E + (R[1] - L[1]) * E1 * 5. / 12.);
  }
  const double dxs = c->dx * c->scalev;
  H[0] = (19 * tp[0] / 4. - 15 * tp[5]) / dxs;
  H[5] = (60 * tp[5] - 15 * tp[0]) / dxs;
  for (int l = 1; l < 5; l++) H[l] = tp[l] * 12. / dxs;
}
/* RK3: advection_1.cpp:412-576 */
void lpo_RK3(const lpo_ctx *c, double *U)
{
  const size_t n = (size_t)c->Nx * c->size_v;
  const double dt = c->dt;
  double *U1 = (double *)malloc(sizeof(double) * 6 * n), *U2 = (double *)malloc(sizeof(double) * 6 * n);
  field_t F; field_alloc(&F, c->Nx);
  field_compute(c, U, &F);
  #pragma omp parallel for
  for (size_t k = 0; k < n; k++) { double H[6]; dg_rhs(c, U, &F, k, H); for (int l = 0; l < 6; l++) U1[6 * k + l] = U[6 * k + l] + dt * H[l]; }
  field_compute(c, U1, &F);
  #pragma omp parallel for
  for (size_t k = 0; k < n; k++) { double H[6]; dg_rhs(c, U1, &F, k, H); for (int l = 0; l < 6; l++) U2[6 * k + l] = 0.75 * U[6 * k + l] + 0.25 * U1[6 * k + l] + 0.25 * dt * H[l]; }
  field_compute(c, U2, &F);
  #pragma omp parallel for
  for (size_t k = 0; k < n; k++) { double H[6]; dg_rhs(c, U2, &F, k, H); for (int l = 0; l < 6; l++) U1[6 * k + l] = U[6 * k + l] / 3. + U2[6 * k + l] * 2. / 3. + dt * H[l] * 2. / 3.; }
  memcpy(U, U1, sizeof(double) * 6 * n);
  free(U1); free(U2); free(F.cp);
}

/* ------------------------------------------------------------------------------------------ */
/* initial conditions: SetInit_1.cpp:43-49 (Mw), :35-40 (f_2Gauss), :68-123 (SetInit_LD),
 * :175-258 (SetInit_4H), :261-325 (SetInit_4H_Homo); 5-point Gauss-Legendre, advection_1.cpp:12-13 */
static const double GW[5] = {0.5688888888888889, 0.4786286704993665, 0.4786286704993665, 0.2369268850561891, 0.2369268850561891};
static const double GT[5] = {0., -0.5384693101056831, 0.5384693101056831, -0.9061798459386640, 0.9061798459386640};
static double maxwellian(double v1, double v2, double v3, double T)
{
  double r2 = v1 * v1 + v2 * v2 + v3 * v3;
  return exp(-r2 / (2 * T)) / (2 * M_PI * T * sqrt(2 * T * M_PI));
}
static double two_gauss(double v1, double v2, double v3)
{
  double sig = M_PI / 10;
  return 0.5 * (exp(-((v1 - 2 * sig) * (v1 - 2 * sig) + v2 * v2 + v3 * v3) / (2 * sig * sig))
                + exp(-((v1 + 2 * sig) * (v1 + 2 * sig) + v2 * v2 + v3 * v3) / (2 * sig * sig)))
         / (2 * M_PI * sig * sig * sqrt(2 * M_PI * sig * sig));
}
static double maxwellian_x(double x) { double T = 0.4; return exp(-x * x / (2 * T)) / sqrt(2 * T * M_PI); }
/* cell moments of a velocity profile against the DG test functions: tmp0..tmp4 of the reference */
static void vcell_moments_T(const lpo_ctx *c, int kind, double T, double s1, double s2, double s3, int j1, int j2, int j3, double t[5]);
static void vcell_moments(const lpo_ctx *c, int kind, double s1, double s2, double s3, int j1, int j2, int j3, double t[5])
{
  vcell_moments_T(c, kind, 0.4, s1, s2, s3, j1, j2, j3, t);
}
static void vcell_moments_T(const lpo_ctx *c, int kind, double T, double s1, double s2, double s3, int j1, int j2, int j3, double t[5])
{
  const double dv = c->dv;
  for (int l = 0; l < 5; l++) t[l] = 0.;
  for (int m1 = 0; m1 < 5; m1++) for (int m2 = 0; m2 < 5; m2++) for (int m3 = 0; m3 < 5; m3++) {
    double a = gridv(c, (double)j1) + 0.5 * dv * GT[m1] + s1, b = gridv(c, (double)j2) + 0.5 * dv * GT[m2] + s2,
           d = gridv(c, (double)j3) + 0.5 * dv * GT[m3] + s3;
    double tp = GW[m1] * GW[m2] * GW[m3] * (kind == 1 ? two_gauss(a, b, d) : maxwellian(a, b, d, T));
    t[0] += tp; t[1] += tp * 0.5 * GT[m1]; t[2] += tp * 0.5 * GT[m2]; t[3] += tp * 0.5 * GT[m3];
    t[4] += tp * 0.25 * (GT[m1] * GT[m1] + GT[m2] * GT[m2] + GT[m3] * GT[m3]);
  }
  for (int l = 0; l < 5; l++) t[l] = t[l] * 0.5 * 0.5 * 0.5;
}
void lpo_SetInit_LD(const lpo_ctx *c, double *U, double a, double kw, int twostream)
{
  const int Nv = c->Nv; const double dx = c->dx;
  for (int j1 = 0; j1 < Nv; j1++) for (int j2 = 0; j2 < Nv; j2++) for (int j3 = 0; j3 < Nv; j3++) {
    double t[5]; vcell_moments(c, twostream ? 1 : 0, 0., 0., 0., j1, j2, j3, t);
    for (int i = 0; i < c->Nx; i++) {
      size_t k = (size_t)i * c->size_v + (j1 * Nv * Nv + j2 * Nv + j3);
      double xp = gridx(c, i + 0.5), xm = gridx(c, i - 0.5);
      double xf = dx + (sin(kw * xp) - sin(kw * xm)) * a / kw;
      double tp0 = xf * t[0] / dx, tp5 = xf * t[4] / dx;
      U[k * 6 + 0] = 19 * tp0 / 4. - 15 * tp5;
      U[k * 6 + 5] = 60 * tp5 - 15 * tp0;
      U[k * 6 + 1] = (0.5 * (sin(kw * xp) + sin(kw * xm)) + (cos(kw * xp) - cos(kw * xm)) / (kw * dx)) * (a / kw) * t[0] * 12. / dx;
      U[k * 6 + 2] = xf * t[1] * 12 / dx; U[k * 6 + 3] = xf * t[2] * 12 / dx; U[k * 6 + 4] = xf * t[3] * 12 / dx;
    }
  }
}
/* DG coefficients of ND * Maxwellian(T) on velocity cell (j1,j2,j3): the body shared by SetInit_ND
 * (SetInit_1.cpp:125-173) and DirichletBC (advection_1.cpp:24-69) */
static void nd_maxwellian_cell(const lpo_ctx *c, double ND, double T, int j1, int j2, int j3, double u[6])
{
  double t[5]; vcell_moments_T(c, 0, T, 0., 0., 0., j1, j2, j3, t);
  const double tp0 = ND * t[0], tp5 = ND * t[4];
  u[0] = 19 * tp0 / 4. - 15 * tp5;
  u[5] = 60 * tp5 - 15 * tp0;
  u[1] = 0;
  u[2] = ND * t[1] * 12; u[3] = ND * t[2] * 12; u[4] = ND * t[3] * 12;
}
static double doping_profile(const lpo_ctx *c, int i) { return (i <= c->a_i || i > c->b_i) ? c->NH : c->NL; }   /* FieldCalculations.cpp:413-425 */
/* Doping = True: LP_ompi.cpp:157-166 (a_i, b_i), ReadDopingParameters */
void lpo_set_doping(lpo_ctx *c, double NL, double NH, double eps, double T_L, double T_R)
{
  const int Nv = c->Nv;
  c->doping = 1; c->NL = NL; c->NH = NH; c->eps = eps; c->T_L = T_L; c->T_R = T_R;
  c->a_i = c->Nx / 3 - 1; c->b_i = 2 * c->Nx / 3 - 1;
  free(c->dirL); free(c->dirR);
  c->dirL = (double *)malloc(sizeof(double) * 6 * c->size_v);
  c->dirR = (double *)malloc(sizeof(double) * 6 * c->size_v);
  for (int j1 = 0; j1 < Nv; j1++) for (int j2 = 0; j2 < Nv; j2++) for (int j3 = 0; j3 < Nv; j3++) {
    const size_t j = (size_t)(j1 * Nv * Nv + j2 * Nv + j3);
    nd_maxwellian_cell(c, doping_profile(c, 0), T_L, j1, j2, j3, c->dirL + 6 * j);
    nd_maxwellian_cell(c, doping_profile(c, c->Nx - 1), T_R, j1, j2, j3, c->dirR + 6 * j);
  }
}
void lpo_SetInit_ND(const lpo_ctx *c, double *U)
{
  const int Nv = c->Nv;
  for (int j1 = 0; j1 < Nv; j1++) for (int j2 = 0; j2 < Nv; j2++) for (int j3 = 0; j3 < Nv; j3++)
    for (int i = 0; i < c->Nx; i++) {
      size_t k = (size_t)i * c->size_v + (j1 * Nv * Nv + j2 * Nv + j3);
      nd_maxwellian_cell(c, doping_profile(c, i), c->T_R, j1, j2, j3, U + 6 * k);
    }
}
void lpo_SetInit_4H(const lpo_ctx *c, double *U)
{
  const int Nv = c->Nv; const double C = 1., dx = c->dx;
  for (int p = 0; p < 4; p++) {
    double sv_ = C * pow(-1, p), sx = C * pow(-1, (int)(p / 2));
    for (int j1 = 0; j1 < Nv; j1++) for (int j2 = 0; j2 < Nv; j2++) for (int j3 = 0; j3 < Nv; j3++) {
      double t[5]; vcell_moments(c, 0, sv_, sv_, sv_, j1, j2, j3, t);
      for (int i = 0; i < c->Nx; i++) {
        double x0 = 0., x1 = 0.;
        for (int m = 0; m < 5; m++) { double tpx = GW[m] * maxwellian_x(gridx(c, (double)i) + 0.5 * dx * GT[m] - c->Lx / 2 + sx); x0 += tpx; x1 += tpx * 0.5 * GT[m]; }
        x0 *= 0.5; x1 *= 0.5;
        size_t k = (size_t)i * c->size_v + (j1 * Nv * Nv + j2 * Nv + j3);
        double tp0 = x0 * t[0], tp5 = x0 * t[4];
        double add[6] = {19 * tp0 / 4. - 15 * tp5, x1 * t[0] * 12, x0 * t[1] * 12, x0 * t[2] * 12, x0 * t[3] * 12, 60 * tp5 - 15 * tp0};
        for (int l = 0; l < 6; l++) U[k * 6 + l] = (p == 0 ? 0. : U[k * 6 + l]) + add[l];
      }
    }
  }
  for (size_t q = 0; q < 6 * (size_t)c->Nx * c->size_v; q++) U[q] = U[q] / 4;
}
void lpo_SetInit_4H_Homo(const lpo_ctx *c, double *U)
{
  const int Nv = c->Nv; const double C = 0.02;
  for (int p = 0; p < 4; p++) {
    double s1 = C * pow(-1, (int)(p / 2)), s23 = C * pow(-1, p);
    for (int j1 = 0; j1 < Nv; j1++) for (int j2 = 0; j2 < Nv; j2++) for (int j3 = 0; j3 < Nv; j3++) {
      double t[5]; vcell_moments(c, 0, s1, s23, s23, j1, j2, j3, t);
      size_t k = (size_t)(j1 * Nv * Nv + j2 * Nv + j3);
      double add[6] = {19 * t[0] / 4. - 15 * t[4], 0., t[1] * 12, t[2] * 12, t[3] * 12, 60 * t[4] - 15 * t[0]};
      for (int l = 0; l < 6; l++) if (l != 1) U[k * 6 + l] = (p == 0 ? 0. : U[k * 6 + l]) + add[l];
      if (p == 0) U[k * 6 + 1] = 0.; /* the reference leaves U1 unset (malloc); 0 is what it reads in practice */
    }
  }
  for (size_t q = 0; q < 6 * (size_t)c->size_v; q++) U[q] = U[q] / 4;
}

/* moments: MomentCalculations.cpp:23-131 (mass, momentum, KiE), :201-230 (computeEleE) */
void lpo_moments(const lpo_ctx *c, const double *U, double *out)
{
  const int Nv = c->Nv, sv = c->size_v; const double dv = c->dv, dx = c->dx;
  const size_t n = (size_t)c->ncell * sv;
  double ms = 0., p1 = 0., p2 = 0., p3 = 0., ke = 0.;
  for (size_t k = 0; k < n; k++) {
    int j = (int)(k % sv), j3 = j % Nv, j2 = (j / Nv) % Nv, j1 = j / (Nv * Nv);
    double c1 = gridv(c, (double)j1), c2 = gridv(c, (double)j2), c3 = gridv(c, (double)j3), r2 = c1 * c1 + c2 * c2 + c3 * c3;
    const double *u = U + 6 * k;
    ms += u[0] + u[5] / 4.;
    p1 += c1 * dv * u[0] + u[2] * dv * dv / 12. + u[5] * c1 * dv / 4.;
    p2 += c2 * dv * u[0] + u[3] * dv * dv / 12. + u[5] * c2 * dv / 4.;
    p3 += c3 * dv * u[0] + u[4] * dv * dv / 12. + u[5] * c3 * dv / 4.;
    ke += u[0] * (r2 + dv * dv / 4.) * dv + (c1 * u[2] + c2 * u[3] + c3 * u[4]) * dv * dv / 6. + u[5] * (dv * dv * dv * 19. / 240. + r2 * dv / 4.);
  }
  const double xs = c->homogeneous ? 1. : dx;
  out[0] = ms * xs * c->scalev;
  out[1] = p1 * xs * dv * dv; out[2] = p2 * xs * dv * dv; out[3] = p3 * xs * dv * dv;
  out[4] = 0.5 * ke * xs * dv * dv;
  out[5] = 0.;
  if (!c->homogeneous) {
    field_t F; field_alloc(&F, c->Nx); field_compute(c, U, &F);
    const double Lx = c->Lx, ce = F.ce;
    double t4 = 0., t5 = 0., t6 = 0.;
    for (int i = 0; i < c->Nx; i++) {
      double cc = dx * dx * (0.5 * F.m[i] - F.s[i] / 12.), cp1 = F.cp[i];
      double tp1 = F.m[i] / c->scalev, tp2 = F.s[i] / c->scalev, xi = gridx(c, (double)i), xl = gridx(c, i - 0.5), xr = gridx(c, i + 0.5);
      t4 += dx * cp1 + cc;
      t5 += dx * xi * cp1;
      t5 += c->scalev * (tp1 * ((pow(xr, 3) - pow(xl, 3)) / 3. - xl * xi * dx) - tp2 * dx * dx * xi / 12.);
      tp2 *= dx / 2.;
      t6 += cp1 * cp1 * dx + 2 * cp1 * cc + pow(dv, 6) * (tp1 * tp1 * dx * dx * dx / 3. + tp2 * tp2 * dx / 30. - tp1 * tp2 * dx * dx / 6.);
    }
    out[5] = 0.5 * (ce * ce * Lx + Lx * Lx * Lx / 3. - ce * Lx * Lx + 2 * ce * t4 - 2 * t5 + t6);
    free(F.cp);
  }
}

/* Per-step diagnostics the reference runs on rank 0: computeEntropy_Inhomo/_Homo
 * (EntropyCalculations.cpp:23-77 / 79-122), FindNegVals (NegativityChecks.cpp:24-160, cell averages by
 * the same Gauss rule) and computeKiEratio (MomentCalculations.cpp:133-199).
 * out4 = entropy, sum of KiE terms over cells with f_avg >= 0, over cells with f_avg < 0, number of
 * negative cells. */
void lpo_diagnostics(const lpo_ctx *c, const double *U, double *out4)
{
  const int Nv = c->Nv, sv = c->size_v, nxq = c->homogeneous ? 1 : 5;
  const double dv = c->dv;
  const size_t n = (size_t)c->ncell * sv;
  double ent = 0., kpos = 0., kneg = 0., nneg = 0.;
  for (size_t k = 0; k < n; k++) {
    const double *u = U + 6 * k;
    int j = (int)(k % sv), j3 = j % Nv, j2 = (j / Nv) % Nv, j1 = j / (Nv * Nv);
    double e = 0., avg = 0.;
    for (int a = 0; a < nxq; a++) for (int b = 0; b < 5; b++) for (int cc = 0; cc < 5; cc++) for (int d = 0; d < 5; d++) {
      double xs = c->homogeneous ? 0. : 0.5 * GT[a], x1 = 0.5 * GT[b], x2 = 0.5 * GT[cc], x3 = 0.5 * GT[d];
      double f = u[0] + (c->homogeneous ? 0. : u[1] * xs) + u[2] * x1 + u[3] * x2 + u[4] * x3 + u[5] * (x1 * x1 + x2 * x2 + x3 * x3);
      double w = (c->homogeneous ? 1. : GW[a]) * GW[b] * GW[cc] * GW[d];
      if (f > 0) e += w * f * log(f);
      avg += w * f;
    }
    ent += e;
    double c1 = gridv(c, (double)j1), c2 = gridv(c, (double)j2), c3 = gridv(c, (double)j3), r2 = c1 * c1 + c2 * c2 + c3 * c3;
    double ke = u[0] * (r2 + dv * dv / 4.) * dv + (c1 * u[2] + c2 * u[3] + c3 * u[4]) * dv * dv / 6. + u[5] * (dv * dv * dv * 19. / 240. + r2 * dv / 4.);
    if (avg < 0) { kneg += ke; nneg += 1.; } else kpos += ke;
  }
  out4[0] = ent * 0.5 * dv * 0.5 * dv * 0.5 * dv * (c->homogeneous ? 1. : 0.5 * c->dx);
  out4[1] = kpos; out4[2] = kneg; out4[3] = nneg;
}

/* one pass of the while(t<nT) body without diagnostics: LP_ompi.cpp:662-813 */
void lpo_step(const lpo_ctx *c, double *U)
{
  if (!c->homogeneous) lpo_RK3(c, U);
  if (c->nu > 0.) lpo_collide_step(c, U);
}
