/* oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Function-level access to the UNMODIFIED reference.  oracle/Makefile compiles the reference's
 * own translation units where they lie under /root/reference/source (LP_ompi.cpp with
 * -Dmain=ref_main so its globals are kept but its main() is renamed) against oracle/shim/, and
 * links them with this file into oracle/_ref/libref.so.  This file contains no reference code:
 * it only sets the reference's global variables the way its main() does (LP_ompi.cpp:169-375,
 * 359-373) and forwards extern "C" calls to the reference's functions, so tests can feed the
 * reference arbitrary inputs and read back element-wise results (the golden .dc files pin only
 * 8 digits of 5 scalars).
 */
#include "LP_ompi.h"
#include <cstring>
#include <vector>

static double **g_W = NULL;     /* conv_weights (reference layout: W[xi][omega]) */
static double **g_f = NULL;     /* f[chunk_Nx][N^3] */
static fftw_complex *g_qHat = NULL;
static int g_ready = 0;

extern "C" {

/* Mirrors LP_ompi.cpp:169-184 (sizes), :205-225 (chunking, 1 rank), :227-243 and :245-348
 * (allocation), :359-375 (grids, trapezoid weights, conservation matrices) and :424 (weights).
 * build_weights=0 skips the N^6 table (needed only for ComputeQ/RK4). */
int ref_setup(int Nx_, int Nv_, int N_, double Lv_, double Lx_, double nu_, double dt_,
              int homogeneous, int gamma, int build_weights)
{
  if (g_ready) { /* re-configuration: drop the big table, leak the small arrays (test process only) */
    if (g_W) { for (int i = 0; i < size_ft; i++) free(g_W[i]); free(g_W); g_W = NULL; }
    g_ready = 0;
  }
  myrank_mpi = 0; nprocs_mpi = 1;
  Nx = Nx_; Nv = Nv_; N = N_; Lv = Lv_; Lx = Lx_; nu = nu_; dt = dt_;
  Homogeneous = homogeneous != 0;
  FullandLinear = false; LinearLandau = false; MassConsOnly = false; Doping = false;
  Damping = false; TwoStream = false; FourHump = false; TwoHump = false;
  First = true; Second = false;
  size_v = Nv * Nv * Nv;
  size = Homogeneous ? size_v : Nx * size_v;
  size_ft = N * N * N;
  dv = 2. * Lv / Nv;
  dx = Lx / Nx;
  scalev = dv * dv * dv;
  L_v = Lv; R_v = Lv;
  scaleL = 8 * Lv * Lv * Lv;
  M = 5;
  chunksize_dg = size; chunksize_ft = size_ft;
  if (Homogeneous) { chunk_Nx = 1; nprocs_Nx = 1; } else { chunk_Nx = Nx; nprocs_Nx = 1; }

  U1 = (double *)malloc(size * 6 * sizeof(double));
  Utmp = (double *)malloc(chunksize_dg * 6 * sizeof(double));
  output_buffer_vp = (double *)malloc(chunksize_dg * 6 * sizeof(double));
  cp = (double *)malloc(Nx * sizeof(double));
  intE = (double *)malloc(Nx * sizeof(double));
  intE1 = (double *)malloc(Nx * sizeof(double));
  intE2 = (double *)malloc(Nx * sizeof(double));
  fNegVals = (int *)malloc(size * sizeof(int));
  fAvgVals = (double *)malloc(size * sizeof(double));

  C1_5 = (double **)malloc(M * sizeof(double *));
  C2 = (double **)malloc(M * sizeof(double *));
  for (int i = 0; i < M; i++) {
    C1_5[i] = (double *)malloc(size_ft * sizeof(double));
    C2[i] = (double *)malloc(size_ft * sizeof(double));
  }
  g_f = (double **)malloc(chunk_Nx * sizeof(double *));
  for (int i = 0; i < chunk_Nx; i++) g_f[i] = (double *)malloc(size_ft * sizeof(double));
  Q = (double *)malloc(size_ft * sizeof(double));
  f1 = (double *)malloc(size_ft * sizeof(double));
  Q1 = (double *)malloc(size_ft * sizeof(double));
  Utmp_coll = (double *)malloc((size_t)chunk_Nx * size_v * 5 * sizeof(double));
  temp = (fftw_complex *)fftw_malloc(size_ft * sizeof(fftw_complex));
  g_qHat = (fftw_complex *)fftw_malloc(size_ft * sizeof(fftw_complex));
  fftw_init_threads();
  fftw_plan_with_nthreads(omp_get_max_threads());
  p_forward = fftw_plan_dft_3d(N, N, N, temp, temp, FFTW_FORWARD, FFTW_MEASURE);
  p_backward = fftw_plan_dft_3d(N, N, N, temp, temp, FFTW_BACKWARD, FFTW_MEASURE);
  wtN = (double *)malloc(N * sizeof(double));
  v = (double *)malloc(N * sizeof(double));
  eta = (double *)malloc(N * sizeof(double));
  Q1_fft = (fftw_complex *)fftw_malloc(size_ft * sizeof(fftw_complex));
  Q2_fft = (fftw_complex *)fftw_malloc(size_ft * sizeof(fftw_complex));
  Q3_fft = (fftw_complex *)fftw_malloc(size_ft * sizeof(fftw_complex));
  fftOut = (fftw_complex *)fftw_malloc(size_ft * sizeof(fftw_complex));
  fftIn = (fftw_complex *)fftw_malloc(size_ft * sizeof(fftw_complex));

  scale = 1.0 / sqrt(2.0 * M_PI);
  scale3 = pow(scale, 3.0);
  L_eta = 0.5 * (double)(N - 1) * PI / L_v;
  h_v = 2.0 * L_v / (double)(N - 1);
  h_eta = 2.0 * L_eta / (double)(N);
  for (int i = 0; i < N; i++) {
    eta[i] = -L_eta + (double)i * h_eta;
    v[i] = -L_v + (double)i * h_v;
  }
  trapezoidalRule(N, wtN);
  createCCtAndPivot();
  if (build_weights) {
    g_W = (double **)malloc(size_ft * sizeof(double *));
    for (int i = 0; i < size_ft; i++) g_W[i] = (double *)malloc(size_ft * sizeof(double));
    generate_conv_weights(g_W, gamma);
  }
  g_ready = 1;
  return 0;
}

/* --- accessors for tables the reference built ------------------------------------------- */
void ref_get_grids(double *v_out, double *eta_out, double *wt_out)
{
  memcpy(v_out, v, N * sizeof(double));
  memcpy(eta_out, eta, N * sizeof(double));
  memcpy(wt_out, wtN, N * sizeof(double));
}
void ref_get_conservation(double *C_out /* 5*N^3: C1_5[0], C2[1..3], C1_5[4] */, double *CCt_out /* 25 */)
{
  memcpy(C_out + 0 * size_ft, C1_5[0], size_ft * sizeof(double));
  memcpy(C_out + 1 * size_ft, C2[1], size_ft * sizeof(double));
  memcpy(C_out + 2 * size_ft, C2[2], size_ft * sizeof(double));
  memcpy(C_out + 3 * size_ft, C2[3], size_ft * sizeof(double));
  memcpy(C_out + 4 * size_ft, C1_5[4], size_ft * sizeof(double));
  memcpy(CCt_out, CCt, 25 * sizeof(double));
}
/* one row W[xi][.] of the reference's table */
int ref_get_weight_row(int xi, double *row) { if (!g_W) return 1; memcpy(row, g_W[xi], size_ft * sizeof(double)); return 0; }
double ref_gHat3(double e1, double e2, double e3, double k1, double k2, double k3, int gamma)
{ return gHat3(e1, e2, e3, k1, k2, k3, gamma); }
void ref_IntModes(int k1, int k2, int k3, int j1, int j2, int j3, double *out10) { IntModes(k1, k2, k3, j1, j2, j3, out10); }

/* --- collision path ---------------------------------------------------------------------- */
void ref_fft3D(const double *in /* N^3 x2 */, double *out) { fft3D((fftw_complex *)in, (fftw_complex *)out); }
void ref_FS(const double *in, double *out) { FS((fftw_complex *)in, (fftw_complex *)out); }
int ref_ComputeQ(const double *f, double *qHat) { if (!g_W) return 1; ComputeQ((double *)f, (fftw_complex *)qHat, g_W); return 0; }
void ref_conserveMoments(double *qHat) { conserveMoments((fftw_complex *)qHat); }
void ref_setInit_spectral(const double *U, double *f_out /* chunk_Nx * N^3 */)
{
  setInit_spectral((double *)U, g_f);
  for (int i = 0; i < chunk_Nx; i++) memcpy(f_out + (size_t)i * size_ft, g_f[i], size_ft * sizeof(double));
}
/* The collision branch of the time loop, LP_ompi.cpp:669-754 with one rank: U is updated in place. */
int ref_collide_step(double *U)
{
  if (!g_W) return 1;
  setInit_spectral(U, g_f);
  int ncell = Homogeneous ? 1 : Nx;
  for (int l = 0; l < ncell; l++) {
    ComputeQ(g_f[l % chunk_Nx], g_qHat, g_W);
    conserveMoments(g_qHat);
    RK4(g_f[l % chunk_Nx], l, g_qHat, g_W, U, Utmp_coll);
  }
  for (int l = 0; l < ncell; l++)
    for (int k = 0; k < size_v; k++) {
      size_t kv = (size_t)l * size_v + k;
      U[kv * 6 + 0] = Utmp_coll[kv * 5];
      U[kv * 6 + 5] = Utmp_coll[kv * 5 + 4];
      U[kv * 6 + 2] = Utmp_coll[kv * 5 + 1];
      U[kv * 6 + 3] = Utmp_coll[kv * 5 + 2];
      U[kv * 6 + 4] = Utmp_coll[kv * 5 + 3];
    }
  return 0;
}
/* stage spectra of the last RK4 call (globals Q1_fft..Q3_fft) for element-wise checks */
void ref_get_stage_spectra(double *q1, double *q2, double *q3)
{
  memcpy(q1, Q1_fft, size_ft * sizeof(fftw_complex));
  memcpy(q2, Q2_fft, size_ft * sizeof(fftw_complex));
  memcpy(q3, Q3_fft, size_ft * sizeof(fftw_complex));
}

/* --- advection path ----------------------------------------------------------------------- */
void ref_RK3(double *U) { RK3(U); }
/* field integrals for one stage: out = ce, then cp[Nx], intE[Nx], intE1[Nx], intE2[Nx] (advection_1.cpp:419-429) */
void ref_field(const double *U_in, double *out)
{
  double *U = (double *)U_in;
  ce = computePhi_x_0(U);
  for (int i = 0; i < Nx; i++) { cp[i] = computeC_rho(U, i); intE[i] = Int_E(U, i); intE1[i] = Int_E1st(U, i); }
  for (int i = 0; i < Nx; i++) intE2[i] = Int_E2nd(U, i);
  out[0] = ce;
  memcpy(out + 1, cp, Nx * sizeof(double));
  memcpy(out + 1 + Nx, intE, Nx * sizeof(double));
  memcpy(out + 1 + 2 * Nx, intE1, Nx * sizeof(double));
  memcpy(out + 1 + 3 * Nx, intE2, Nx * sizeof(double));
}

/* --- initial conditions and diagnostics ---------------------------------------------------- */
void ref_SetInit_LD(double *U, double A, double k, int twostream)
{ A_amp = A; k_wave = k; Damping = !twostream; TwoStream = twostream != 0; SetInit_LD(U); Damping = false; TwoStream = false; }
void ref_SetInit_4H(double *U) { SetInit_4H(U); }
void ref_SetInit_4H_Homo(double *U) { SetInit_4H_Homo(U); }
void ref_SetInit_2H(double *U) { SetInit_2H(U); }
/* out[0..5] = mass, P1, P2, P3, KiE, EleE (EleE = 0 when homogeneous) -- LP_ompi.cpp:820-827 */
void ref_moments(const double *U_in, double *out)
{
  double *U = (double *)U_in, a[3];
  out[0] = computeMass(U);
  computeMomentum(U, a);
  out[1] = a[0]; out[2] = a[1]; out[3] = a[2];
  out[4] = computeKiE(U);
  out[5] = Homogeneous ? 0. : computeEleE(U);
}
double ref_entropy(const double *U) { return computeEntropy((double *)U); }
/* out4 = entropy, KiEneg/KiEpos ratio, number of negative cells, 0 -- LP_ompi.cpp:819,829,846 */
void ref_diagnostics(const double *U_in, double *out4)
{
  double *U = (double *)U_in;
  FindNegVals(U, fNegVals, fAvgVals);
  out4[0] = computeEntropy(U);
  out4[1] = computeKiEratio(U, fNegVals);
  double n = 0.;
  for (int k = 0; k < size; k++) n += fNegVals[k];
  out4[2] = n; out4[3] = 0.;
}
int ref_num_threads(void) { return omp_get_max_threads(); }
/* launchers such as torchrun export OMP_NUM_THREADS=1: bench.py sets the thread count explicitly (before ref_setup, which sizes the FFT plans) */
void ref_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }
}
