// Two-thread thirds for the y/x stage kernel of the FFT-convolution ComputeQ (L = 16, N = 32, M = 48).
//
// STATUS: groundwork for the next version of k_fc3_f2_tmem.  The task algebra below is
// checked on the CPU by the thread-loop emulator (tests/emul, tests/test_fc3_emul.py); the experimental kernel
// k_fc3_f2h_tmem (fftconv.cu, LPGPU_F2_HALF=1, off by default) uses it: correct on hardware, not yet faster.
//
// Why: one thread per 16-point third (fc3.cuh) needs 168 registers, so 12 warps fit an SM and the FP64 pipe idles half
// the time.  Here a PAIR of adjacent threads shares a third.  With n = 4 n1 + n2 and k = k1 + 4 k2,
//   forward   X[k1 + 4 k2] = sum_n2 w4^(n2 k2) [ w16^(n2 k1) sum_n1 y[4 n1 + n2] w4^(n1 k1) ]
// thread h of the pair owns the residue classes n2 = 2h, 2h+1: it loads only their 8 of the 16 (a0, a1) input pairs, runs
// the radix-3 pre-stage and the first radix-4 stage (over n1) on them, hands the partner the four values that belong to
// its k1 (one shuffle exchange of 4 complex numbers), and finishes X for k1 = 2h, 2h+1 (radix-4 over n2).  The inverse is
// the mirror image: radix-4 over k2 on the owned k1, exchange, radix-4 over k1 for the owned n2, so the same thread that
// loaded x[4 n1 + n2] on the way in produces it on the way out.  Half the loads, half the registers (compile-only probe:
// 63 / 80 registers at 384 threads per CTA, scripts/proto/half_third.cu), every lane busy in the x stage (288 tasks).
//
// Every function is split at the exchange into an `_a` and a `_b` part so that the emulator (which runs the threads of
// a CTA one after the other, phase by phase) can stand in for the shuffle with an array.
#pragma once
#include "fc3.cuh"

namespace fc3 {

struct Half16 {
  static constexpr int L = 16, M = 48;
  // v * w_M^(SIGN t), t = t0 for the h = 0 thread of a pair and t1 for the h = 1 thread.  Both are compile-time
  // constants once the callers' loops are unrolled, so the factor is a select between two immediates -- the first
  // version looked t up in the constant bank with a per-lane index (two addresses per warp, serialised) and ran at
  // half the speed of the one-thread kernel
  template <int SIGN>
  static LP_HD double2 tw2(double2 v, int t0, int t1, int h)
  {
    t0 %= M; t1 %= M;
    const double c = h ? Tw<M>::c(t1) : Tw<M>::c(t0), sp = h ? Tw<M>::s(t1) : Tw<M>::s(t0), s = SIGN > 0 ? sp : -sp;
    return make_double2(v.x * c - v.y * s, v.x * s + v.y * c);
  }
  // input slot i = 4 c + n1 of thread h  <->  line index l = 4 n1 + 2 h + c   (c = 0, 1; n1 = 0..3)
  static LP_HD int l_of(int h, int i) { return 4 * (i & 3) + 2 * h + (i >> 2); }
  // output slot o = 4 kk + k2 of thread h  <->  spectral index k = 2 h + kk + 4 k2   (kk = 0, 1; k2 = 0..3)
  static LP_HD int k_of(int h, int o) { return 2 * h + (o >> 2) + 4 * (o & 3); }

  // ---- forward third r of a line: a0[i] = x[l_of(h, i)], a1[i] = x[l_of(h, i) + L]  ->  X[3 k + r] at slot o
  // part a: pre-stage, radix-4 over n1, twiddle; keep[2 kk + c] stays, send[2 kk + c] goes to the partner
  static LP_HD void fwd_a(const double2 (&a0)[8], const double2 (&a1)[8], int r, int h, double2 (&keep)[4], double2 (&send)[4])
  {
    #pragma unroll
    for (int c = 0; c < 2; c++) {
      double2 v[4];
      #pragma unroll
      for (int n1 = 0; n1 < 4; n1++) {
        const double2 p = a0[4 * c + n1], q = a1[4 * c + n1];
        const int l0 = 4 * n1 + c, l1 = l0 + 2;               // the line index of this slot for h = 0 / h = 1
        if (r == 0) v[n1] = cadd(p, q);
        else if (r == 1) {                                    // w3^r = -1/2 -/+ i sqrt(3)/2
          const double2 b = make_double2(p.x - 0.5 * q.x + LP_SQ3H * q.y, p.y - 0.5 * q.y - LP_SQ3H * q.x);
          v[n1] = tw2<-1>(b, l0, l1, h);
        } else {
          const double2 b = make_double2(p.x - 0.5 * q.x - LP_SQ3H * q.y, p.y - 0.5 * q.y + LP_SQ3H * q.x);
          v[n1] = tw2<-1>(b, 2 * l0, 2 * l1, h);
        }
      }
      dft4<-1>(v[0], v[1], v[2], v[3]);
      #pragma unroll
      for (int k1 = 0; k1 < 4; k1++) {
        const double2 y = tw2<-1>(v[k1], 3 * c * k1, 3 * (2 + c) * k1, h);   // w16^(n2 k1) = w48^(3 n2 k1), n2 = 2 h + c
        const int kk = k1 & 1;
        if ((k1 >> 1) == h) keep[2 * kk + c] = y; else send[2 * kk + c] = y;
      }
    }
  }
  // part b: got = the partner's send; radix-4 over n2 for k1 = 2 h + kk
  static LP_HD void fwd_b(const double2 (&keep)[4], const double2 (&got)[4], int h, double2 (&out)[8])
  {
    #pragma unroll
    for (int kk = 0; kk < 2; kk++) {
      double2 m0 = h == 0 ? keep[2 * kk] : got[2 * kk], m1 = h == 0 ? keep[2 * kk + 1] : got[2 * kk + 1];      // n2 = 0, 1
      double2 m2 = h == 0 ? got[2 * kk] : keep[2 * kk], m3 = h == 0 ? got[2 * kk + 1] : keep[2 * kk + 1];      // n2 = 2, 3
      dft4<-1>(m0, m1, m2, m3);
      out[4 * kk] = m0; out[4 * kk + 1] = m1; out[4 * kk + 2] = m2; out[4 * kk + 3] = m3;
    }
  }
  // ---- inverse third r: z[o] = Z[3 k_of(h, o) + r]  ->  t_r[l] = conj(w_M^(r l)) IFFT_L(z)[l] at slot i (l = l_of(h, i))
  // part a: radix-4 over k2 on the owned k1, twiddle w16^(+n2 k1); values of the partner's n2 are sent
  static LP_HD void inv_a(const double2 (&z)[8], int h, double2 (&keep)[4], double2 (&send)[4])
  {
    #pragma unroll
    for (int kk = 0; kk < 2; kk++) {
      const int k1 = 2 * h + kk;
      double2 v0 = z[4 * kk], v1 = z[4 * kk + 1], v2 = z[4 * kk + 2], v3 = z[4 * kk + 3];
      dft4<+1>(v0, v1, v2, v3);                               // index n2 = 0..3
      const double2 v[4] = {v0, v1, v2, v3};
      #pragma unroll
      for (int n2 = 0; n2 < 4; n2++) {
        const double2 y = tw2<+1>(v[n2], 3 * n2 * kk, 3 * n2 * (2 + kk), h);    // k1 = 2 h + kk
        const int c = n2 & 1;
        if ((n2 >> 1) == h) keep[2 * c + kk] = y; else send[2 * c + kk] = y;
      }
    }
  }
  // part b: radix-4 over k1 for the owned n2 = 2 h + c, then the twist conj(w_M^(r l))
  static LP_HD void inv_b(const double2 (&keep)[4], const double2 (&got)[4], int r, int h, double2 (&out)[8])
  {
    #pragma unroll
    for (int c = 0; c < 2; c++) {
      double2 m0 = h == 0 ? keep[2 * c] : got[2 * c], m1 = h == 0 ? keep[2 * c + 1] : got[2 * c + 1];          // k1 = 0, 1
      double2 m2 = h == 0 ? got[2 * c] : keep[2 * c], m3 = h == 0 ? got[2 * c + 1] : keep[2 * c + 1];          // k1 = 2, 3
      dft4<+1>(m0, m1, m2, m3);                               // index n1 = 0..3
      const double2 m[4] = {m0, m1, m2, m3};
      #pragma unroll
      for (int n1 = 0; n1 < 4; n1++) {
        const int l0 = 4 * n1 + c, l1 = l0 + 2;
        out[4 * c + n1] = r == 0 ? m[n1] : r == 1 ? tw2<+1>(m[n1], l0, l1, h) : tw2<+1>(m[n1], 2 * l0, 2 * l1, h);
      }
    }
  }
};

// =========================================================================================================
// F2H: the tasks of the y/x stage kernel on half thirds.  CTA = (cell, kz), NT = 384 threads, the shared arrays of
// fc3::F2<16> (IN, Y and their aliases T, T2).
//   y stage          thread = (array, r, x, h)      384 tasks
//   x stage, x inv   thread = (r, ky, h)            288 tasks = nine full warps
//   y inverse        thread = (ry, xo, h)           192 tasks
struct F2H {
  typedef F2<16> K;
  static constexpr int L = 16, N = 32, M = 48, PY = K::PY, PN = K::PN, NT = 384;
  struct Ex { double2 keep[4], send[4]; };        // what a thread holds across the exchange

  static LP_HD void ytask(int tid, int &h, int &x, int &r, int &arr) { h = tid & 1; x = (tid >> 1) % N; r = (tid / (2 * N)) % 3; arr = tid / (6 * N); }
  static LP_HD bool xtask(int tid, int &h, int &r, int &ky) { h = tid & 1; const int t = tid >> 1; r = t / M; ky = t % M; if (r > 2) { r = 2; return false; } return true; }
  static LP_HD bool yitask(int tid, int &h, int &ry, int &xo) { h = tid & 1; const int t = tid >> 1; xo = t % N; ry = t / N; if (ry > 2) { ry = 2; return false; } return true; }

  // y stage, part a (same role as F2::ystage up to the exchange); p = 1 reuses the y transform of p = 0 for v
  static LP_HD void ystage_a(int tid, int p, const double2 *IN, const double *sE, Ex &e)
  {
    int h, x, r, arr; ytask(tid, h, x, r, arr);
    if (arr == 1 && p == 1) return;
    const double2 *src = IN + arr * N * N + x;
    double2 a0[8], a1[8];
    const int yp = ypow_of(p);
    #pragma unroll
    for (int i = 0; i < 8; i++) {
      const int l = Half16::l_of(h, i);
      a0[i] = src[l * N]; a1[i] = src[(l + L) * N];
      if (arr == 1 && yp) {
        const double e0 = ipow(sE[l], yp), e1 = ipow(sE[l + L], yp);
        a0[i].x *= e0; a0[i].y *= e0; a1[i].x *= e1; a1[i].y *= e1;
      }
    }
    Half16::fwd_a(a0, a1, r, h, e.keep, e.send);
  }
  static LP_HD void ystage_b(int tid, int p, const double *sE, const Ex &e, const double2 (&got)[4], double2 *Y)
  {
    int h, x, r, arr; ytask(tid, h, x, r, arr);
    double2 *dst = Y + arr * N * PY + x * PY + r * L;
    if (arr == 1 && p == 1) {
      const double sx = -ipow(sE[x], 2);            // v_1 = -E(x)^2 fhat: rescale the transform stored at p = 0
      #pragma unroll
      for (int o = 0; o < 8; o++) { const int k = Half16::k_of(h, o); const double2 v = dst[k]; dst[k] = make_double2(v.x * sx, v.y * sx); }
      return;
    }
    double2 out[8];
    Half16::fwd_b(e.keep, got, h, out);
    const double sx = (arr == 1 && p > 0) ? -ipow(sE[x], xpow_of(p)) : 1.;
    #pragma unroll
    for (int o = 0; o < 8; o++) dst[Half16::k_of(h, o)] = make_double2(out[o].x * sx, out[o].y * sx);
  }
  // x stage: transform of array `arr` (0: u, 1: v) of this thread's line
  static LP_HD void xfwd_a(int tid, int arr, const double2 *Y, Ex &e)
  {
    int h, r, ky; xtask(tid, h, r, ky);
    const double2 *src = Y + arr * N * PY + ky;
    double2 a0[8], a1[8];
    #pragma unroll
    for (int i = 0; i < 8; i++) { const int l = Half16::l_of(h, i); a0[i] = src[l * PY]; a1[i] = src[(l + L) * PY]; }
    Half16::fwd_a(a0, a1, r, h, e.keep, e.send);
  }
  static LP_HD void xfwd_b(int tid, const Ex &e, const double2 (&got)[4], double2 (&out)[8])
  {
    int h, r, ky; xtask(tid, h, r, ky);
    Half16::fwd_b(e.keep, got, h, out);
  }
  static LP_HD void product(const double2 (&uh)[8], const double2 (&vh)[8], double2 (&acc)[8])
  {
    #pragma unroll
    for (int o = 0; o < 8; o++) { acc[o].x += uh[o].x * vh[o].x - uh[o].y * vh[o].y; acc[o].y += uh[o].x * vh[o].y + uh[o].y * vh[o].x; }
  }
  // inverse x of the accumulated products -> T[(r L + l)][ky]   (T aliases Y)
  static LP_HD void xinv_a(int tid, const double2 (&acc)[8], Ex &e) { int h, r, ky; xtask(tid, h, r, ky); Half16::inv_a(acc, h, e.keep, e.send); }
  static LP_HD void xinv_b(int tid, const Ex &e, const double2 (&got)[4], double2 *T)
  {
    int h, r, ky;
    if (!xtask(tid, h, r, ky)) return;
    double2 out[8];
    Half16::inv_b(e.keep, got, r, h, out);
    #pragma unroll
    for (int i = 0; i < 8; i++) T[(r * L + Half16::l_of(h, i)) * PY + ky] = out[i];
  }
  // inverse y of the N kept x rows -> T2[(ry L + l')][xo]   (T2 aliases IN)
  static LP_HD void yinv_a(int tid, const double2 *T, Ex &e)
  {
    int h, ry, xo; yitask(tid, h, ry, xo);
    const int l = xo % L, s = xo / L + 1;
    double2 z[8];
    #pragma unroll
    for (int o = 0; o < 8; o++) {
      const int kp = ry * L + Half16::k_of(h, o);
      z[o] = inv_combine(T[l * PY + kp], T[(L + l) * PY + kp], T[(2 * L + l) * PY + kp], s);
    }
    Half16::inv_a(z, h, e.keep, e.send);
  }
  static LP_HD void yinv_b(int tid, const Ex &e, const double2 (&got)[4], double2 *T2)
  {
    int h, ry, xo;
    if (!yitask(tid, h, ry, xo)) return;
    double2 out[8];
    Half16::inv_b(e.keep, got, ry, h, out);
    #pragma unroll
    for (int i = 0; i < 8; i++) T2[(ry * L + Half16::l_of(h, i)) * PN + xo] = out[i];
  }
};

} // namespace fc3
