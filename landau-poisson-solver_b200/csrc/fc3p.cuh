// Task algebra of the warp-specialised, persistent y/x-stage kernel of the FFT-convolution ComputeQ (k_fc3_f2s,
// fftconv.cu; L = 16, N = 32, M = 48).  Same transforms as fc3::F2 (fc3.cuh), cut into the tasks of the two warp roles:
//
//   y-warps (3): warp j owns sub-transform r = j of every line along y, lane = x.  Per product p it transforms the
//                column x of the u plane, then of the v plane, into the Y buffer of this product (double-buffered),
//                one product ahead of the x-warps; at the end of a plane it runs the inverse y transform and stores C.
//   x-warps (5): 144 tasks (r, ky): transform the two lines along x of Y, multiply, accumulate over p; at the end of a
//                plane inverse x into T.
//
// The functions are __host__ __device__ and free of synchronisation, so the CPU thread-loop emulator (tests/emul) runs
// them in dependency order and checks the index algebra without a GPU; the kernel adds the mbarrier pipeline around them.
#pragma once
#include "fc3.cuh"

namespace fc3 {

struct F2P {
  static constexpr int L = 16, N = 32, M = 48, PY = M + 1, PN = N + 1;
  static constexpr int NYT = 96;            // y-role threads: t = 32 j + x
  static constexpr int NXT = 160;           // x-role threads (five warps); tasks t < 144
  static constexpr int PLANE_C2 = N * N;    // one staged input plane [y][x]
  static constexpr int YARR_C2 = N * PY;    // one y-transformed array [x][ky]
  static constexpr int YBUF_C2 = 2 * YARR_C2;
  static constexpr int T_C2 = M * PY;       // inverse-x output [kx'][ky]; T2 [ky'][xo] (pitch PN) aliases it
  static_assert(M * PN <= T_C2, "T2 must fit in T");

  // ---- y role: one sub-transform of column x of array `arr` (0 = u_p, 1 = v source of p) -> Ybuf[arr][x][r L + q]
  // plane: the staged input plane [y][x]; for arr = 1 the y- and x-monomials of v_p are applied on the way in / out.
  static LP_HD void ythird(int arr, int r, int x, int p, const double2 *plane, const double *sE, double2 *Ybuf)
  {
    const double2 *src = plane + x;
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
    double2 *dst = Ybuf + arr * YARR_C2 + x * PY + r * L;
    #pragma unroll
    for (int q = 0; q < L; q++) dst[q] = yv[q];
  }
  // p = 1: v_1 = -E(x)^2 fhat has the y transform of v_0 = fhat, which this thread wrote into the other Y buffer at p = 0
  static LP_HD void yrescale(int r, int x, const double *sE, const double2 *Yprev, double2 *Ybuf)
  {
    const double sx = -ipow(sE[x], 2);
    const double2 *s = Yprev + YARR_C2 + x * PY + r * L;
    double2 *dst = Ybuf + YARR_C2 + x * PY + r * L;
    #pragma unroll
    for (int q = 0; q < L; q++) { const double2 v = s[q]; dst[q] = make_double2(v.x * sx, v.y * sx); }
  }
  // ---- x role: task t < 144 = (r, ky); the other lanes of the fifth warp repeat the last task and store nothing
  static LP_HD bool xtask(int t, int &r, int &ky)
  {
    const bool ok = t < 3 * M;
    if (!ok) t = 3 * M - 1;
    r = t / M; ky = t % M;
    return ok;
  }
  static LP_HD void xload(const double2 *Yarr, int ky, double2 (&a0)[L], double2 (&a1)[L])
  {
    #pragma unroll
    for (int l = 0; l < L; l++) { a0[l] = Yarr[l * PY + ky]; a1[l] = Yarr[(l + L) * PY + ky]; }
  }
  static LP_HD void xinverse_store(bool ok, int r, int ky, double2 (&acc)[L], double2 *T)
  {
    inv_third<L>(acc, r);
    if (!ok) return;
    #pragma unroll
    for (int l = 0; l < L; l++) T[(r * L + l) * PY + ky] = acc[l];
  }
  // ---- y role at the end of a plane: inverse y of the N kept x rows.  t = 32 ry + xo.
  // part a reads T; (all y-role threads synchronise); part b writes T2 over T; (synchronise); store reads T2.
  static LP_HD void yinv_a(int t, const double2 *T, double2 (&c)[L])
  {
    const int xo = t % N, ry = t / N, l = xo % L, s = xo / L + 1;
    #pragma unroll
    for (int q = 0; q < L; q++) {
      const int kp = ry * L + q;
      c[q] = inv_combine(T[l * PY + kp], T[(L + l) * PY + kp], T[(2 * L + l) * PY + kp], s);
    }
    inv_third<L>(c, ry);
  }
  static LP_HD void yinv_b(int t, const double2 (&c)[L], double2 *T2)
  {
    const int xo = t % N, ry = t / N;
    #pragma unroll
    for (int lp = 0; lp < L; lp++) T2[(ry * L + lp) * PN + xo] = c[lp];
  }
  static LP_HD void store(int t, const double2 *T2, double2 *Cplane)
  {
    for (int idx = t; idx < N * N; idx += NYT) {
      const int xo = idx / N, yo = idx % N, lp = yo % L, s = yo / L + 1;
      Cplane[idx] = inv_combine(T2[lp * PN + xo], T2[(L + lp) * PN + xo], T2[(2 * L + lp) * PN + xo], s);
    }
  }
  // the staged loads of one plane in consumption order: (p, arr) for h = 0 .. 12 (p = 1 has no v plane)
  static constexpr int LOADS_PER_PLANE = 13;
  static LP_HD void load_of(int h, int &p, int &arr)
  {
    // h: 0 u0, 1 v0, 2 u1, 3 u2, 4 v2, 5 u3, 6 v3, ...
    if (h < 2) { p = 0; arr = h; return; }
    if (h == 2) { p = 1; arr = 0; return; }
    p = 2 + (h - 3) / 2; arr = (h - 3) % 2;
  }
  static LP_HD const double2 *load_src(const double2 *Z, int cell, int kz, int p, int arr)
  {
    return F2<L>::plane(Z, cell, arr ? 7 + zpow_of(p) : p, kz);
  }
};

} // namespace fc3
