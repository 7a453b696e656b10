// ComputeQ as seven zero-padded linear convolutions, register-resident transforms (computeq_variant 0/2).
//
//   Wt(xi,omega) = G0(omega) - sum_{p=1..6} G_p(omega) mono_p(E(beta)),   beta = xi + N/2 - omega,
//   mono = {e1^2, e2^2, e3^2, e1 e2, e1 e3, e2 e3},  e_a = E(beta_a) = eta[beta_a] - eta[N/2]
//   => Qhat[xi] = sum_p ( u_p (*) v_p )[xi + N/2],   u_p = G_p fhat,  v_p = -mono_p(E) fhat  (v_0 = fhat)
// (the sum ComputeQ evaluates pair by pair, collisionRoutines_1.cpp:706-773), on cyclic transforms of size
// M = 3N/2 = 3L per dimension: the linear convolution lives on [0, 2N-2], only s = xi + N/2 in [L, 3L) is
// wanted, and with period M the aliases of that window fall outside the support.
//
// One length-M line is three length-L transforms held entirely in registers (no shuffles, no shared-memory
// butterflies):  forward, inputs n < 2L non-zero (zero padding costs nothing):
//     X[3q + r] = FFT_L( y_r )[q],   y_r[l] = (x[l] + w3^r x[l+L]) w_M^(r l),     r = 0,1,2
//   inverse, only outputs n = l + L s, s = 1,2 are formed (truncation costs nothing):
//     x[l + L s] = sum_r conj(w3)^(s r) t_r[l],   t_r[l] = conj(w_M^(r l)) IFFT_L( Z[3q + r] )[l]
// Every transformed axis is stored at "position" r*L + q; products are position-wise, so no permutation
// pass exists anywhere.
//
// Pipeline (B cells per launch); the separable monomials of v_p are applied where the index is at hand:
//   F1  CTA (cell, y):  z-lines of u_0..u_6 and of E(z)^m fhat, m = 0,1,2  ->  Z[cell][a][kz][y][x]   (10 arrays)
//   F2  CTA (cell, kz): for p = 0..6: y- then x-transform of u_p and v_p (E(y)^k on the way in, -E(x)^k on the
//                       way out of the y stage), accumulate uh*vh in registers; inverse x, inverse y
//                       ->  C[cell][kz][xo][yo]
//   F3  CTA (cell, xo): inverse z, scale M^-3  ->  Qhat[cell][xo][yo][zo]
// Only Z (10 N^2 M complex per cell) and C (N^2 M) make a round trip through HBM/L2.
//
// The per-thread code is __host__ __device__ and organised in barrier-separated phases so that the same
// source runs under a CPU thread-loop emulator (tests/emul/fc3_emul.cpp) -- this container has no GPU.
#pragma once
#include "fc_twiddles.h"

#ifdef __CUDACC__
#define LP_HD __host__ __device__ __forceinline__
#else
#define LP_HD inline
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
#endif

namespace fc3 {

#define LP_SQ3H 0.86602540378443864676   // sqrt(3)/2

// ---- twiddles exp(+2 pi i t / M), t < M: __constant__ on the device (folded into c[bank][imm] operands once the
// loops are unrolled), plain arrays on the host
#ifdef __CUDACC__
__device__ __constant__ double d_twc8[8] = LP_TWC_8, d_tws8[8] = LP_TWS_8;
__device__ __constant__ double d_twc16[16] = LP_TWC_16, d_tws16[16] = LP_TWS_16;
__device__ __constant__ double d_twc32[32] = LP_TWC_32, d_tws32[32] = LP_TWS_32;
__device__ __constant__ double d_twc12[12] = LP_TWC_12, d_tws12[12] = LP_TWS_12;
__device__ __constant__ double d_twc24[24] = LP_TWC_24, d_tws24[24] = LP_TWS_24;
__device__ __constant__ double d_twc36[36] = LP_TWC_36, d_tws36[36] = LP_TWS_36;
__device__ __constant__ double d_twc48[48] = LP_TWC_48, d_tws48[48] = LP_TWS_48;
#endif
static const double h_twc8[8] = LP_TWC_8, h_tws8[8] = LP_TWS_8;
static const double h_twc16[16] = LP_TWC_16, h_tws16[16] = LP_TWS_16;
static const double h_twc32[32] = LP_TWC_32, h_tws32[32] = LP_TWS_32;
static const double h_twc12[12] = LP_TWC_12, h_tws12[12] = LP_TWS_12;
static const double h_twc24[24] = LP_TWC_24, h_tws24[24] = LP_TWS_24;
static const double h_twc36[36] = LP_TWC_36, h_tws36[36] = LP_TWS_36;
static const double h_twc48[48] = LP_TWC_48, h_tws48[48] = LP_TWS_48;

template <int M> struct Tw;
#ifdef __CUDA_ARCH__
#define LP_TW_SPEC(M)                                                  \
  template <> struct Tw<M> {                                           \
    static LP_HD double c(int t) { return d_twc##M[t]; }               \
    static LP_HD double s(int t) { return d_tws##M[t]; }               \
  };
#else
#define LP_TW_SPEC(M)                                                  \
  template <> struct Tw<M> {                                           \
    static LP_HD double c(int t) { return h_twc##M[t]; }               \
    static LP_HD double s(int t) { return h_tws##M[t]; }               \
  };
#endif
LP_TW_SPEC(8) LP_TW_SPEC(12) LP_TW_SPEC(16) LP_TW_SPEC(24) LP_TW_SPEC(32) LP_TW_SPEC(36) LP_TW_SPEC(48)

LP_HD double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
LP_HD double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
LP_HD double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// v * exp(SIGN 2 pi i t / M); t is a compile-time constant once the callers' loops are unrolled
template <int M, int SIGN>
LP_HD double2 mul_tw(double2 v, int t)
{
  t %= M;
  if (t == 0) return v;
  if (4 * t == M) return SIGN > 0 ? make_double2(-v.y, v.x) : make_double2(v.y, -v.x);
  if (2 * t == M) return make_double2(-v.x, -v.y);
  if (4 * t == 3 * M) return SIGN > 0 ? make_double2(v.y, -v.x) : make_double2(-v.y, v.x);
  const double c = Tw<M>::c(t), s = SIGN > 0 ? Tw<M>::s(t) : -Tw<M>::s(t);
  return make_double2(v.x * c - v.y * s, v.x * s + v.y * c);
}

// ---- small DFTs, sign of the exponent = SIGN
template <int SIGN>
LP_HD void dft2(double2 &a, double2 &b) { const double2 t = csub(a, b); a = cadd(a, b); b = t; }
template <int SIGN>
LP_HD void dft3(double2 &a0, double2 &a1, double2 &a2)
{
  const double2 s = cadd(a1, a2), d = csub(a1, a2);
  const double2 m = make_double2(a0.x - 0.5 * s.x, a0.y - 0.5 * s.y);
  a0 = cadd(a0, s);
  // SIGN i (sqrt3/2) d = SIGN (-(sqrt3/2) d.y, (sqrt3/2) d.x)
  const double cx = SIGN > 0 ? -LP_SQ3H : LP_SQ3H;
  a1 = make_double2(m.x + cx * d.y, m.y - cx * d.x);
  a2 = make_double2(m.x - cx * d.y, m.y + cx * d.x);
}
template <int SIGN>
LP_HD void dft4(double2 &a0, double2 &a1, double2 &a2, double2 &a3)
{
  const double2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = csub(a1, a3);
  // SIGN i t3
  const double2 it3 = SIGN > 0 ? make_double2(-t3.y, t3.x) : make_double2(t3.y, -t3.x);
  a0 = cadd(t0, t2); a2 = csub(t0, t2);
  a1 = cadd(t1, it3); a3 = csub(t1, it3);
}
template <int R, int SIGN>
LP_HD void dftR(double2 *v)
{
  if constexpr (R == 2) dft2<SIGN>(v[0], v[1]);
  if constexpr (R == 3) dft3<SIGN>(v[0], v[1], v[2]);
  if constexpr (R == 4) dft4<SIGN>(v[0], v[1], v[2], v[3]);
}

template <int L> struct Fac;
template <> struct Fac<4>  { static constexpr int R1 = 4, R2 = 1; };
template <> struct Fac<8>  { static constexpr int R1 = 2, R2 = 4; };
template <> struct Fac<12> { static constexpr int R1 = 3, R2 = 4; };
template <> struct Fac<16> { static constexpr int R1 = 4, R2 = 4; };

// L-point DFT in registers, natural order in and out:  X[k] = sum_n x[n] exp(SIGN 2 pi i n k / L),
// L = R1 R2, n = R2 n1 + n2, k = k1 + R1 k2.
template <int L, int SIGN>
LP_HD void fft_small(double2 (&x)[L])
{
  constexpr int R1 = Fac<L>::R1, R2 = Fac<L>::R2, M = 3 * L;
  double2 y[L];
  #pragma unroll
  for (int n2 = 0; n2 < R2; n2++) {
    double2 v[R1];
    #pragma unroll
    for (int n1 = 0; n1 < R1; n1++) v[n1] = x[R2 * n1 + n2];
    dftR<R1, SIGN>(v);
    #pragma unroll
    for (int k1 = 0; k1 < R1; k1++) y[k1 * R2 + n2] = mul_tw<M, SIGN>(v[k1], 3 * n2 * k1);   // w_L^(n2 k1) = w_M^(3 n2 k1)
  }
  #pragma unroll
  for (int k1 = 0; k1 < R1; k1++) {
    double2 v[R2 > 1 ? R2 : 1];
    #pragma unroll
    for (int n2 = 0; n2 < R2; n2++) v[n2] = y[k1 * R2 + n2];
    if (R2 > 1) dftR<R2, SIGN>(v);
    #pragma unroll
    for (int k2 = 0; k2 < R2; k2++) x[k1 + R1 * k2] = v[k2];
  }
}

// N-point DFT in registers for the shifted transforms fft3D / FS (N = 8, 16, 24, 32; any N = 2^a 3^b works), natural
// order in and out, X[k] = sum_n x[n] exp(SIGN 2 pi i n k / N); twiddles from the size-TB table (N divides TB).
template <int N> struct Rad { static constexpr int R1 = (N % 4 == 0) ? 4 : (N % 3 == 0) ? 3 : 2; };
template <int N, int SIGN, int TB>
LP_HD void fftN(double2 (&x)[N])
{
  if constexpr (N <= 4) {
    dftR<N, SIGN>(x);
  } else {
    constexpr int R1 = Rad<N>::R1, R2 = N / R1;
    double2 y[N];
    #pragma unroll
    for (int n2 = 0; n2 < R2; n2++) {
      double2 v[R1];
      #pragma unroll
      for (int n1 = 0; n1 < R1; n1++) v[n1] = x[R2 * n1 + n2];
      dftR<R1, SIGN>(v);
      #pragma unroll
      for (int k1 = 0; k1 < R1; k1++) y[k1 * R2 + n2] = mul_tw<TB, SIGN>(v[k1], (TB / N) * n2 * k1);
    }
    #pragma unroll
    for (int k1 = 0; k1 < R1; k1++) {
      double2 v[R2];
      #pragma unroll
      for (int n2 = 0; n2 < R2; n2++) v[n2] = y[k1 * R2 + n2];
      fftN<R2, SIGN, TB>(v);
      #pragma unroll
      for (int k2 = 0; k2 < R2; k2++) x[k1 + R1 * k2] = v[k2];
    }
  }
}

// forward pre-stage of sub-transform r: y[l] = (a0[l] + w3^r a1[l]) w_M^(r l), w3 = exp(-2 pi i/3), w_M = exp(-2 pi i/M)
template <int L>
LP_HD void fwd_pre(const double2 (&a0)[L], const double2 (&a1)[L], int r, double2 (&y)[L])
{
  constexpr int M = 3 * L;
  if (r == 0) {
    #pragma unroll
    for (int l = 0; l < L; l++) y[l] = cadd(a0[l], a1[l]);
  } else if (r == 1) {
    #pragma unroll
    for (int l = 0; l < L; l++) {
      const double2 b = make_double2(a0[l].x - 0.5 * a1[l].x + LP_SQ3H * a1[l].y, a0[l].y - 0.5 * a1[l].y - LP_SQ3H * a1[l].x);
      y[l] = mul_tw<M, -1>(b, l);
    }
  } else {
    #pragma unroll
    for (int l = 0; l < L; l++) {
      const double2 b = make_double2(a0[l].x - 0.5 * a1[l].x - LP_SQ3H * a1[l].y, a0[l].y - 0.5 * a1[l].y + LP_SQ3H * a1[l].x);
      y[l] = mul_tw<M, -1>(b, 2 * l);
    }
  }
}
// forward line third: X[3q + r], q < L, from the 2L non-zero inputs (a0 = x[0..L), a1 = x[L..2L))
template <int L>
LP_HD void fwd_third(const double2 (&a0)[L], const double2 (&a1)[L], int r, double2 (&y)[L])
{
  fwd_pre<L>(a0, a1, r, y);
  fft_small<L, -1>(y);
}
// inverse line third, in place: z[q] = Z[3q + r]  ->  t_r[l] = conj(w_M^(r l)) IFFT_L(z)[l]
template <int L>
LP_HD void inv_third(double2 (&z)[L], int r)
{
  constexpr int M = 3 * L;
  fft_small<L, +1>(z);
  if (r == 1) {
    #pragma unroll
    for (int l = 0; l < L; l++) z[l] = mul_tw<M, +1>(z[l], l);
  } else if (r == 2) {
    #pragma unroll
    for (int l = 0; l < L; l++) z[l] = mul_tw<M, +1>(z[l], 2 * l);
  }
}
// x[l + L s] = t0 + w^s t1 + w^(2s) t2, w = exp(+2 pi i/3), s = 1 or 2
LP_HD double2 inv_combine(double2 t0, double2 t1, double2 t2, int s)
{
  const double2 S = cadd(t1, t2), D = csub(t1, t2);
  const double2 m = make_double2(t0.x - 0.5 * S.x, t0.y - 0.5 * S.y);
  const double c = (s == 1) ? LP_SQ3H : -LP_SQ3H;
  return make_double2(m.x - c * D.y, m.y + c * D.x);
}

// ---- line halves: two threads per M-point line, H = M/2 = 3L/2 outputs each (F2Q: the x stage of a three-warp CTA is then
// exactly N x 3 / ... = 2M tasks, and a line's 2L inputs are read twice instead of three times)
// forward: y[q] = X[2q + h], q < H, from the 2L non-zero inputs:  b[n] = (x[n] + (-1)^h x[n + H]) w_M^(h n)  (x[n + H] exists
// for n < L/2), X[2q + h] = sum_{n < H} b[n] exp(-2 pi i n q / H)
template <int L>
LP_HD void fwd_half(const double2 (&a0)[L], const double2 (&a1)[L], int h, double2 (&y)[3 * L / 2])
{
  constexpr int M = 3 * L, H = M / 2, P = L / 2;      // pairs: n < P  (n + H = L + (n + P))
  if (h == 0) {
    #pragma unroll
    for (int n = 0; n < P; n++) y[n] = cadd(a0[n], a1[n + P]);
    #pragma unroll
    for (int n = P; n < L; n++) y[n] = a0[n];
    #pragma unroll
    for (int n = L; n < H; n++) y[n] = a1[n - L];
  } else {
    #pragma unroll
    for (int n = 0; n < P; n++) y[n] = mul_tw<M, -1>(csub(a0[n], a1[n + P]), n);
    #pragma unroll
    for (int n = P; n < L; n++) y[n] = mul_tw<M, -1>(a0[n], n);
    #pragma unroll
    for (int n = L; n < H; n++) y[n] = mul_tw<M, -1>(a1[n - L], n);
  }
  fftN<H, -1, M>(y);
}
// inverse, in place: z[q] = Z[2q + h]  ->  t_h[n] = conj(w_M^(h n)) IFFT_H(z)[n];  x[n] = t_0[n] + t_1[n], x[n + H] = t_0[n] - t_1[n]
template <int L>
LP_HD void inv_half(double2 (&z)[3 * L / 2], int h)
{
  constexpr int M = 3 * L, H = M / 2;
  fftN<H, +1, M>(z);
  if (h == 1) {
    #pragma unroll
    for (int n = 0; n < H; n++) z[n] = mul_tw<M, +1>(z[n], n);
  }
}

// 16-byte asynchronous global -> shared copy (plain copy under the emulator)
LP_HD void cp16(double2 *dst_smem, const double2 *src)
{
#ifdef __CUDA_ARCH__
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src) : "memory");
#else
  *dst_smem = *src;
#endif
}
LP_HD void cp_wait_all()
{
#ifdef __CUDA_ARCH__
  asm volatile("cp.async.wait_all;\n" ::: "memory");
#endif
}

// v_p = sgn * E(x)^xp E(y)^yp E(z)^zp fhat
LP_HD int zpow_of(int p) { return p == 3 ? 2 : (p == 5 || p == 6) ? 1 : 0; }
LP_HD int ypow_of(int p) { return p == 2 ? 2 : (p == 4 || p == 6) ? 1 : 0; }
LP_HD int xpow_of(int p) { return p == 1 ? 2 : (p == 4 || p == 5) ? 1 : 0; }
LP_HD double ipow(double e, int k) { return k == 0 ? 1. : k == 1 ? e : e * e; }

// =========================================================================================================
// F1: CTA = (cell, y), NT = 6N threads; thread = (array slot a' < 2, r, x); five rounds cover the 10 arrays.
//   shared: FS[N][N+1] (fhat slab [x][z]), sE[N]
//   Gt: [7][y][z][x] (x fastest) so that the lanes of a warp read consecutive doubles
template <int L>
struct F1 {
  static constexpr int N = 2 * L, M = 3 * L, P = N + 1, NT = 6 * N;
  static constexpr int SMEM_C2 = N * P;   // double2 count; plus N doubles
  static LP_HD void load(int tid, int cell, int y, const double2 *fhat, const double *E, double2 *FS, double *sE)
  {
    const double2 *src = fhat + (long long)cell * N * N * N;
    for (int idx = tid; idx < N * N; idx += NT) {
      const int x = idx / N, z = idx % N;
      FS[x * P + z] = src[((long long)x * N + y) * N + z];
    }
    for (int i = tid; i < N; i += NT) sE[i] = E[i];
  }
  // the y slab of a stored spectrum (the Maxwellian of the linear operator) into FM[x][z]
  // cell_stride: N^3 for one spectrum per cell, 0 when every cell uses the same one
  static LP_HD void load_slab(int tid, int cell, int y, const double2 *spec, double2 *FM, long long cell_stride = (long long)N * N * N)
  {
    const double2 *src = spec + (long long)cell * cell_stride;
    for (int idx = tid; idx < N * N; idx += NT) {
      const int x = idx / N, z = idx % N;
      FM[x * P + z] = src[((long long)x * N + y) * N + z];
    }
  }
  // stage the seven kernel-symbol slabs of this y into shared memory: Gs[a][z][x]
  static LP_HD void issue_g(int tid, int y, const double *Gt, double *Gs)
  {
    constexpr int H = N * N / 2;   // double2 chunks per slab
    for (int idx = tid; idx < 7 * H; idx += NT) {
      const int a = idx / H, e = idx % H;
      cp16(reinterpret_cast<double2 *>(Gs) + idx, reinterpret_cast<const double2 *>(Gt + ((long long)a * N + y) * N * N) + e);
    }
  }
  // Gs + a*astride + z*N + x: the slab of array a (shared: astride = N*N; straight from Gt + y*N*N: astride = N^3)
  // rounds [round_begin, round_end) of the five (two arrays each): a launch with few cells gives every round its own CTA
  // FU: the slab the seven u arrays are built from -- FS itself for Q(f,f); the stored transform of the Maxwellian for
  // the linear operator Q(f,M) (ComputeQLinear, collisionRoutines_1.cpp:1185-1269: first factor M, second factor f)
  static LP_HD void lines(int tid, int cell, int y, const double *Gs, long long astride, const double2 *FS, const double *sE, double2 *Z,
                          int round_begin = 0, int round_end = 5, const double2 *FU = nullptr)
  {
    if (!FU) FU = FS;
    const int x = tid % N, r = (tid / N) % 3, slot = tid / (3 * N);
    #pragma unroll 1
    for (int round = round_begin; round < round_end; round++) {
      const int a = round * 2 + slot;
      double2 a0[L], a1[L], yv[L];
      if (a < 7) {
        const double *g = Gs + a * astride + x;
        #pragma unroll
        for (int l = 0; l < L; l++) {
          const double g0 = g[l * N], g1 = g[(l + L) * N];
          const double2 f0 = FU[x * P + l], f1 = FU[x * P + l + L];
          a0[l] = make_double2(g0 * f0.x, g0 * f0.y);
          a1[l] = make_double2(g1 * f1.x, g1 * f1.y);
        }
      } else {
        const int k = a - 7;
        #pragma unroll
        for (int l = 0; l < L; l++) {
          const double e0 = ipow(sE[l], k), e1 = ipow(sE[l + L], k);
          const double2 f0 = FS[x * P + l], f1 = FS[x * P + l + L];
          a0[l] = make_double2(e0 * f0.x, e0 * f0.y);
          a1[l] = make_double2(e1 * f1.x, e1 * f1.y);
        }
      }
      fwd_third<L>(a0, a1, r, yv);
      double2 *o = Z + (((long long)cell * 10 + a) * M + r * L) * (N * N) + y * N + x;
      #pragma unroll
      for (int q = 0; q < L; q++) o[(long long)q * (N * N)] = yv[q];
    }
  }
};

// =========================================================================================================
// F2: CTA = (cell, kz position), NT = 6N threads.
//   shared: IN[2][N*N] (u_p plane, v source plane, [y][x]; later T2[3L][N+1]),
//           Y[2][N][M+1] (y-transformed u and v, [x][ky]; later T[3L][M+1]), sE[N]
//   y stage: thread = (array, r, x);   x stage / inverse x: thread = (r, ky) -- 3M of the 6N threads
template <int L>
struct F2 {
  static constexpr int N = 2 * L, M = 3 * L, PY = M + 1, PN = N + 1, NT = 6 * N;
  static constexpr int IN_C2 = 2 * N * N, Y_C2 = 2 * N * PY;
  static_assert(3 * L * PN <= IN_C2, "T2 must fit the input-plane buffer");
  static_assert(3 * L * PY <= Y_C2, "T must fit the Y buffer");
  // x-stage task of a thread: the 3M tasks (r, ky) on consecutive threads.  L = 16: 144 tasks fill four and a half of
  // the six warps (the sixth issues nothing: an idle warp costs no FP64 slot, an idle lane does); the one warp that
  // holds r = 0 and r = 1 runs both pre-stages, the cheap r = 0 one (additions only) at half mask
  static LP_HD bool xtask(int tid, int &r, int &ky)
  {
    r = tid / M; ky = tid % M;
    if (r > 2) { r = 2; return false; }
    return true;
  }
  // product range [p_begin, p_end) of split sp of nsplit (p = 0 and 1 stay together: p = 1 reuses p = 0's y transform)
  static LP_HD void psplit(int sp, int nsplit, int &p_begin, int &p_end)
  {
    if (nsplit <= 1) { p_begin = 0; p_end = 7; return; }
    p_begin = sp == 0 ? 0 : sp == 1 ? 3 : 5;
    p_end = sp == 0 ? 3 : sp == 1 ? 5 : 7;
  }
  static LP_HD const double2 *plane(const double2 *Z, int cell, int a, int kz) { return Z + (((long long)cell * 10 + a) * M + kz) * (N * N); }
  static LP_HD void issue_loads(int tid, int cell, int kz, int p, const double2 *Z, double2 *IN)
  {
    const double2 *su = plane(Z, cell, p, kz), *sv = plane(Z, cell, 7 + zpow_of(p), kz);
    const int lim = (p == 1) ? N * N : 2 * N * N;      // p = 1 reuses the y-transformed v of p = 0
    for (int idx = tid; idx < lim; idx += NT) {
      const int arr = idx / (N * N), e = idx % (N * N);
      cp16(IN + idx, (arr ? sv : su) + e);
    }
  }
  static LP_HD void ystage(int tid, int p, const double2 *IN, const double *sE, double2 *Y)
  {
    const int x = tid % N, r = (tid / N) % 3, arr = tid / (3 * N);
    if (arr == 1 && p == 1) {
      // v_1 = -E(x)^2 fhat has the y transform of v_0 = fhat, which this thread stored at p = 0: rescale in place
      const double sx = -ipow(sE[x], 2);
      double2 *dst = Y + N * PY + x * PY + r * L;
      #pragma unroll
      for (int q = 0; q < L; q++) { const double2 v = dst[q]; dst[q] = make_double2(v.x * sx, v.y * sx); }
      return;
    }
    const double2 *src = IN + arr * N * N + x;
    double2 a0[L], a1[L], yv[L];
    #pragma unroll
    for (int l = 0; l < L; l++) { a0[l] = src[l * N]; a1[l] = src[(l + L) * N]; }
    const int yp = ypow_of(p);
    if (arr == 1 && yp) {
      #pragma unroll
      for (int l = 0; l < L; l++) {
        const double e0 = ipow(sE[l], yp), e1 = ipow(sE[l + L], yp);
        a0[l].x *= e0; a0[l].y *= e0; a1[l].x *= e1; a1[l].y *= e1;
      }
    }
    fwd_third<L>(a0, a1, r, yv);
    if (arr == 1 && p > 0) {
      const double sx = -ipow(sE[x], xpow_of(p));
      #pragma unroll
      for (int q = 0; q < L; q++) { yv[q].x *= sx; yv[q].y *= sx; }
    }
    double2 *dst = Y + arr * N * PY + x * PY + r * L;
    #pragma unroll
    for (int q = 0; q < L; q++) dst[q] = yv[q];
  }
  static LP_HD void xstage(int tid, const double2 *Y, double2 (&acc)[L])
  {
    int r, ky;
    if (!xtask(tid, r, ky)) return;
    double2 a0[L], a1[L], uh[L], vh[L];
    #pragma unroll
    for (int l = 0; l < L; l++) { a0[l] = Y[l * PY + ky]; a1[l] = Y[(l + L) * PY + ky]; }
    fwd_third<L>(a0, a1, r, uh);
    #pragma unroll
    for (int l = 0; l < L; l++) { a0[l] = Y[N * PY + l * PY + ky]; a1[l] = Y[N * PY + (l + L) * PY + ky]; }
    fwd_third<L>(a0, a1, r, vh);
    #pragma unroll
    for (int q = 0; q < L; q++) {
      acc[q].x += uh[q].x * vh[q].x - uh[q].y * vh[q].y;
      acc[q].y += uh[q].x * vh[q].y + uh[q].y * vh[q].x;
    }
  }
  // inverse x of the accumulated products: T[(r L + l)][ky]  (T aliases Y: call after a barrier)
  static LP_HD void xinverse(int tid, double2 (&acc)[L], double2 *T)
  {
    int r, ky;
    if (!xtask(tid, r, ky)) return;
    inv_third<L>(acc, r);
    #pragma unroll
    for (int l = 0; l < L; l++) T[(r * L + l) * PY + ky] = acc[l];
  }
  // inverse y of the N kept x rows: thread = (ry, xo) -> T2[(ry L + l')][xo]
  static LP_HD void yinverse(int tid, const double2 *T, double2 *T2)
  {
    if (tid >= 3 * N) return;
    const int xo = tid % N, ry = tid / N, l = xo % L, s = xo / L + 1;
    double2 c[L];
    #pragma unroll
    for (int q = 0; q < L; q++) {
      const int kp = ry * L + q;
      c[q] = inv_combine(T[l * PY + kp], T[(L + l) * PY + kp], T[(2 * L + l) * PY + kp], s);
    }
    inv_third<L>(c, ry);
    #pragma unroll
    for (int lp = 0; lp < L; lp++) T2[(ry * L + lp) * PN + xo] = c[lp];
  }
  static LP_HD void store(int tid, int cell, int kz, const double2 *T2, double2 *C)
  {
    double2 *o = C + ((long long)cell * M + kz) * (N * N);
    for (int idx = tid; idx < N * N; idx += NT) {
      const int xo = idx / N, yo = idx % N, lp = yo % L, s = yo / L + 1;
      o[idx] = inv_combine(T2[lp * PN + xo], T2[(L + lp) * PN + xo], T2[(2 * L + lp) * PN + xo], s);
    }
  }
};

// =========================================================================================================
// F2Q: the same plane as F2 on a CTA of 3N threads (three warps at L = 16), four CTAs per SM.  One array at a time: y stage
// of u_p (thirds, thread = (r, x)), x stage of u_p (halves, thread = (h, pos): the transform stays in registers), y stage
// of v_p, x stage of v_p, multiply-accumulate.  shared: IN[N][N] (one plane) | Y[N][PY] (one array); T and T2 alias both.
//   Y[x][r L + q] holds y' = 3q + r (as in F2); pos = r L + q is the x stage's line index
template <int L>
struct F2Q {
  static constexpr int N = 2 * L, M = 3 * L, H = M / 2, PY = M + 1, PN = N + 1, NT = 3 * N;
  static constexpr int IN_C2 = N * N, Y_C2 = N * PY;
  static_assert(2 * M == NT, "x-stage tasks (two halves of M lines) = threads");
  static_assert(2 * H * PY <= IN_C2 + Y_C2, "T (inverse-x halves) must fit the two buffers");
  static_assert(3 * L * PN <= IN_C2 + Y_C2, "T2 must fit the two buffers");
  static LP_HD const double2 *plane(const double2 *Z, int cell, int p, int arr, int kz)
  {
    return Z + (((long long)cell * 10 + (arr ? 7 + zpow_of(p) : p)) * M + kz) * (N * N);
  }
  // y stage of one array (arr = 0: u_p, 1: the v source of p, weighted by E(y)^yp before and -E(x)^xp after the transform)
  static LP_HD void ystage1(int tid, int p, int arr, const double2 *IN, const double *sE, double2 *Y)
  {
    const int x = tid % N, r = tid / N;
    const double2 *src = IN + x;
    double2 a0[L], a1[L], yv[L];
    #pragma unroll
    for (int l = 0; l < L; l++) { a0[l] = src[l * N]; a1[l] = src[(l + L) * N]; }
    const int yp = ypow_of(p);
    if (arr == 1 && yp) {
      #pragma unroll
      for (int l = 0; l < L; l++) {
        const double e0 = ipow(sE[l], yp), e1 = ipow(sE[l + L], yp);
        a0[l].x *= e0; a0[l].y *= e0; a1[l].x *= e1; a1[l].y *= e1;
      }
    }
    fwd_third<L>(a0, a1, r, yv);
    if (arr == 1 && p > 0) {
      const double sx = -ipow(sE[x], xpow_of(p));
      #pragma unroll
      for (int q = 0; q < L; q++) { yv[q].x *= sx; yv[q].y *= sx; }
    }
    double2 *dst = Y + x * PY + r * L;
    #pragma unroll
    for (int q = 0; q < L; q++) dst[q] = yv[q];
  }
  // v_1 = -E(x)^2 fhat from the y transform of v_0 = fhat that is in Y: every thread rescales the third it wrote
  static LP_HD void rescale_v1(int tid, const double *sE, double2 *Y)
  {
    const int x = tid % N, r = tid / N;
    const double sx = -ipow(sE[x], 2);
    double2 *dst = Y + x * PY + r * L;
    #pragma unroll
    for (int q = 0; q < L; q++) { const double2 v = dst[q]; dst[q] = make_double2(v.x * sx, v.y * sx); }
  }
  // x stage of the array in Y: out[q] = transform at x' = 2q + h of line pos
  static LP_HD void xhalf(int tid, const double2 *Y, double2 (&out)[H])
  {
    const int h = tid / M, pos = tid % M;
    double2 a0[L], a1[L];
    #pragma unroll
    for (int l = 0; l < L; l++) { a0[l] = Y[l * PY + pos]; a1[l] = Y[(l + L) * PY + pos]; }
    fwd_half<L>(a0, a1, h, out);
  }
  // inverse x of the accumulated products: T[(h H + n)][pos]  (T aliases IN | Y: call after a barrier)
  static LP_HD void xinverse(int tid, double2 (&acc)[H], double2 *T)
  {
    const int h = tid / M, pos = tid % M;
    inv_half<L>(acc, h);
    #pragma unroll
    for (int n = 0; n < H; n++) T[(h * H + n) * PY + pos] = acc[n];
  }
  // inverse y of the N kept x rows (x index xo + L of M): thread = (ry, xo).  The two halves are combined on the way in;
  // the results go to T2, which aliases T: every thread loads, barrier, every thread stores
  static LP_HD void yinverse_load(int tid, const double2 *T, double2 (&c)[L])
  {
    const int xo = tid % N, ry = tid / N, n = xo + L, np = n < H ? n : n - H;
    #pragma unroll
    for (int q = 0; q < L; q++) {
      const int kp = ry * L + q;
      const double2 t0 = T[np * PY + kp], t1 = T[(H + np) * PY + kp];
      c[q] = n < H ? cadd(t0, t1) : csub(t0, t1);
    }
  }
  static LP_HD void yinverse_store(int tid, double2 (&c)[L], double2 *T2)
  {
    const int xo = tid % N, ry = tid / N;
    inv_third<L>(c, ry);
    #pragma unroll
    for (int lp = 0; lp < L; lp++) T2[(ry * L + lp) * PN + xo] = c[lp];
  }
  static LP_HD void store(int tid, int cell, int kz, const double2 *T2, double2 *C)
  {
    double2 *o = C + ((long long)cell * M + kz) * (N * N);
    for (int idx = tid; idx < N * N; idx += NT) {
      const int xo = idx / N, yo = idx % N, lp = yo % L, s = yo / L + 1;
      o[idx] = inv_combine(T2[lp * PN + xo], T2[(L + lp) * PN + xo], T2[(2 * L + lp) * PN + xo], s);
    }
  }
};

// =========================================================================================================
// F3: CTA = (cell, xo), NT = 3N threads; thread = (rz, yo).  shared: T3[3L][N+1]
template <int L>
struct F3 {
  static constexpr int N = 2 * L, M = 3 * L, PN = N + 1, NT = 3 * N;
  static constexpr int SMEM_C2 = 3 * L * PN;
  // nsplit partial C arrays (split_stride apart) are summed on the way in: F2 may split its seven products over
  // nsplit CTAs per (cell, kz) when few cells are in flight (the inverse transforms are linear)
  template <int NSPLIT>
  static LP_HD void zinverse(int tid, int cell, int xo, const double2 *C, double2 *T3, long long split_stride)
  {
    const int yo = tid % N, rz = tid / N;
    double2 c[L];
    #pragma unroll
    for (int q = 0; q < L; q++) {
      const double2 *src = C + (((long long)cell * M + rz * L + q) * N + xo) * N + yo;
      double2 v = src[0];
      #pragma unroll
      for (int sp = 1; sp < NSPLIT; sp++) v = cadd(v, src[sp * split_stride]);
      c[q] = v;
    }
    inv_third<L>(c, rz);
    #pragma unroll
    for (int l = 0; l < L; l++) T3[(rz * L + l) * PN + yo] = c[l];
  }
  // C5 (nullable): the five conservation rows, planar [m][N^3]; s accumulates this thread's share of the five dot
  // products of conserveAllMoments_Normal (conservationRoutines.cpp:137-144) over the values it stores
  static LP_HD void store(int tid, int cell, int xo, const double2 *T3, double2 *q, const double *C5, double (&s)[5])
  {
    const double sc = 1.0 / ((double)M * M * M);
    constexpr int N3 = N * N * N;
    double2 *o = q + (long long)cell * N3 + (long long)xo * N * N;
    // unrolled with the bound as a predicate: the loads of the conservation rows of all the iterations go out together
    // (the runtime-bound loop fetched them one iteration at a time, eleven L2 round trips per CTA)
    constexpr int IT = (N * N + NT - 1) / NT;
    #pragma unroll
    for (int it = 0; it < IT; it++) {
      const int idx = tid + it * NT;
      if (idx < N * N) {
        const int yo = idx / N, zo = idx % N, l = zo % L, sh = zo / L + 1;
        const double2 v = inv_combine(T3[l * PN + yo], T3[(L + l) * PN + yo], T3[(2 * L + l) * PN + yo], sh);
        const double2 w = make_double2(v.x * sc, v.y * sc);
        o[idx] = w;
        if (C5) {
          const int g = xo * N * N + idx;
          s[0] += w.x * C5[g]; s[1] += w.y * C5[N3 + g]; s[2] += w.y * C5[2 * N3 + g];
          s[3] += w.y * C5[3 * N3 + g]; s[4] += w.x * C5[4 * N3 + g];
        }
      }
    }
  }
};

} // namespace fc3
