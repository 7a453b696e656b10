// CPU thread-loop emulator of the FFT-convolution ComputeQ pipeline (landau-poisson-solver_b200/csrc/fc3.cuh).
// Test infrastructure only: the per-thread phase functions of the CUDA kernels are __host__ __device__;
// here every barrier-separated phase is run for all thread ids of a CTA in turn, CTA by CTA, so the index
// arithmetic and the transform algebra are checked in a container without a GPU.  fc3_direct is the plain
// O(N^6) sum the pipeline must reproduce (collisionRoutines_1.cpp:706-773 with the folded weight).
#include <cstring>
#include <vector>
#include "../../landau-poisson-solver_b200/csrc/fc3.cuh"

// mhat (nullable): the linear operator Q(f, M) -- the u arrays come from the stored Maxwellian transform
template <int L>
static void emulate(int B, const double2 *fhat, const double *G7, const double *E, double2 *q, int nsplit, const double2 *mhat = nullptr, bool quarter = false)
{
  using namespace fc3;
  constexpr int N = 2 * L, M = 3 * L;
  std::vector<double> Gt((size_t)7 * N * N * N);
  for (int a = 0; a < 7; a++)
    for (int x = 0; x < N; x++)
      for (int y = 0; y < N; y++)
        for (int z = 0; z < N; z++) Gt[(((size_t)a * N + y) * N + z) * N + x] = G7[(size_t)7 * (z + N * (y + N * x)) + a];
  const long long split_stride = (long long)B * M * N * N;
  std::vector<double2> Z((size_t)B * 10 * M * N * N), C((size_t)nsplit * split_stride);
  {
    typedef F1<L> K;
    std::vector<double2> FS(K::SMEM_C2), FM(K::SMEM_C2);
    std::vector<double> sE(N), Gs((size_t)7 * N * N);
    for (int cell = 0; cell < B; cell++)
      for (int y = 0; y < N; y++) {
        for (int t = 0; t < K::NT; t++) K::issue_g(t, y, Gt.data(), Gs.data());
        for (int t = 0; t < K::NT; t++) K::load(t, cell, y, fhat, E, FS.data(), sE.data());
        if (mhat) for (int t = 0; t < K::NT; t++) K::load_slab(t, cell, y, mhat, FM.data());
        for (int t = 0; t < K::NT; t++) K::lines(t, cell, y, Gs.data(), N * N, FS.data(), sE.data(), Z.data(), 0, 5, mhat ? FM.data() : nullptr);
      }
  }
  if (quarter) {
    // F2Q (k_fc3_f2q): one array at a time on 3N threads; IN | Y are one contiguous region that T and T2 alias
    typedef F2Q<L> K;
    std::vector<double2> buf(K::IN_C2 + K::Y_C2);
    double2 *IN = buf.data(), *Y = IN + K::IN_C2;
    std::vector<double> sE(E, E + N);
    struct Half { double2 a[K::H]; };
    struct Line { double2 a[L]; };
    std::vector<Half> acc(K::NT), uh(K::NT), vh(K::NT);
    std::vector<Line> cl(K::NT);
    for (int cell = 0; cell < B; cell++)
      for (int kz = 0; kz < M; kz++) {
        std::memset(acc.data(), 0, sizeof(Half) * acc.size());
        // fourteen array passes as in the kernel: product 1 takes its v factor from product 0's Y (rescaled in place) and
        // transforms it FIRST; every other product transforms u first
        for (int k = 0; k < 14; k++) {
          const int p = k < 2 ? 0 : k < 4 ? 1 : 2 + (k - 4) / 2;
          const bool rescale = k == 2, first = k == 0 || k == 2 || (k >= 4 && !((k - 4) & 1));
          const int arr = k == 3 ? 0 : k < 2 ? k : (k - 4) & 1;
          if (rescale) {
            for (int t = 0; t < K::NT; t++) K::rescale_v1(t, sE.data(), Y);
          } else {
            std::memcpy(IN, K::plane(Z.data(), cell, p, arr, kz), sizeof(double2) * N * N);     // the bulk copy
            for (int t = 0; t < K::NT; t++) K::ystage1(t, p, arr, IN, sE.data(), Y);
          }
          for (int t = 0; t < K::NT; t++) {
            if (first) { K::xhalf(t, Y, uh[t].a); continue; }
            K::xhalf(t, Y, vh[t].a);
            for (int q = 0; q < K::H; q++) {
              acc[t].a[q].x += uh[t].a[q].x * vh[t].a[q].x - uh[t].a[q].y * vh[t].a[q].y;
              acc[t].a[q].y += uh[t].a[q].x * vh[t].a[q].y + uh[t].a[q].y * vh[t].a[q].x;
            }
          }
        }
        for (int t = 0; t < K::NT; t++) K::xinverse(t, acc[t].a, buf.data());
        for (int t = 0; t < K::NT; t++) K::yinverse_load(t, buf.data(), cl[t].a);
        for (int t = 0; t < K::NT; t++) K::yinverse_store(t, cl[t].a, buf.data());
        for (int t = 0; t < K::NT; t++) K::store(t, cell, kz, buf.data(), C.data());
      }
  } else {
    typedef F2<L> K;
    std::vector<double2> IN(K::IN_C2), Y(K::Y_C2);
    std::vector<double> sE(E, E + N);
    struct Acc { double2 a[L]; };
    std::vector<Acc> acc(K::NT);
    for (int cell = 0; cell < B; cell++)
      for (int kz = 0; kz < M; kz++)
        for (int sp = 0; sp < nsplit; sp++) {
          int p0, p1;
          K::psplit(sp, nsplit, p0, p1);
          std::memset(acc.data(), 0, sizeof(Acc) * acc.size());
          for (int t = 0; t < K::NT; t++) K::issue_loads(t, cell, kz, p0, Z.data(), IN.data());
          for (int p = p0; p < p1; p++) {
            for (int t = 0; t < K::NT; t++) K::ystage(t, p, IN.data(), sE.data(), Y.data());
            if (p + 1 < p1) for (int t = 0; t < K::NT; t++) K::issue_loads(t, cell, kz, p + 1, Z.data(), IN.data());
            for (int t = 0; t < K::NT; t++) K::xstage(t, Y.data(), acc[t].a);
          }
          for (int t = 0; t < K::NT; t++) K::xinverse(t, acc[t].a, Y.data());
          for (int t = 0; t < K::NT; t++) K::yinverse(t, Y.data(), IN.data());
          for (int t = 0; t < K::NT; t++) K::store(t, cell, kz, IN.data(), C.data() + sp * split_stride);
        }
  }
  {
    typedef F3<L> K;
    std::vector<double2> T3(K::SMEM_C2);
    for (int cell = 0; cell < B; cell++)
      for (int xo = 0; xo < N; xo++) {
        for (int t = 0; t < K::NT; t++) {
          if (nsplit == 3) K::template zinverse<3>(t, cell, xo, C.data(), T3.data(), split_stride);
          else K::template zinverse<1>(t, cell, xo, C.data(), T3.data(), split_stride);
        }
        double dummy[5] = {0., 0., 0., 0., 0.};
        for (int t = 0; t < K::NT; t++) K::store(t, cell, xo, T3.data(), q, nullptr, dummy);
      }
  }
}

extern "C" int fc3_emulate_split(int N, int B, const double *fhat, const double *G7, const double *E, double *q, int nsplit)
{
  const double2 *f = reinterpret_cast<const double2 *>(fhat);
  double2 *o = reinterpret_cast<double2 *>(q);
  switch (N) {
    case 8: emulate<4>(B, f, G7, E, o, nsplit); return 0;
    case 16: emulate<8>(B, f, G7, E, o, nsplit); return 0;
    case 24: emulate<12>(B, f, G7, E, o, nsplit); return 0;
    case 32: emulate<16>(B, f, G7, E, o, nsplit); return 0;
  }
  return 1;
}

extern "C" int fc3_emulate_linear(int N, int B, const double *fhat, const double *mhat, const double *G7, const double *E, double *q)
{
  const double2 *f = reinterpret_cast<const double2 *>(fhat), *m = reinterpret_cast<const double2 *>(mhat);
  double2 *o = reinterpret_cast<double2 *>(q);
  switch (N) {
    case 8: emulate<4>(B, f, G7, E, o, 1, m); return 0;
    case 16: emulate<8>(B, f, G7, E, o, 1, m); return 0;
    case 24: emulate<12>(B, f, G7, E, o, 1, m); return 0;
    case 32: emulate<16>(B, f, G7, E, o, 1, m); return 0;
  }
  return 1;
}

extern "C" int fc3_emulate(int N, int B, const double *fhat, const double *G7, const double *E, double *q)
{
  return fc3_emulate_split(N, B, fhat, G7, E, q, 1);
}
// the four-CTAs-per-SM variant of the y/x stage (F2Q): halves along x, one array at a time
extern "C" int fc3_emulate_quarter(int N, int B, const double *fhat, const double *G7, const double *E, double *q)
{
  const double2 *f = reinterpret_cast<const double2 *>(fhat);
  double2 *o = reinterpret_cast<double2 *>(q);
  switch (N) {
    case 8: emulate<4>(B, f, G7, E, o, 1, nullptr, true); return 0;
    case 16: emulate<8>(B, f, G7, E, o, 1, nullptr, true); return 0;
    case 24: emulate<12>(B, f, G7, E, o, 1, nullptr, true); return 0;
    case 32: emulate<16>(B, f, G7, E, o, 1, nullptr, true); return 0;
  }
  return 1;
}

static int direct_sum(int N, int B, const double *fhat, const double *mhat, const double *G7, const double *E, double *q);
extern "C" int fc3_direct(int N, int B, const double *fhat, const double *G7, const double *E, double *q) { return direct_sum(N, B, fhat, fhat, G7, E, q); }
// ComputeQLinear's pair sum (collisionRoutines_1.cpp:1259-1260): first factor from the Maxwellian, second from f
extern "C" int fc3_direct_linear(int N, int B, const double *fhat, const double *mhat, const double *G7, const double *E, double *q) { return direct_sum(N, B, fhat, mhat, G7, E, q); }
static int direct_sum(int N, int B, const double *fhat, const double *mhat, const double *G7, const double *E, double *q)
{
  const int H = N / 2, N3 = N * N * N;
  for (int cell = 0; cell < B; cell++) {
    const double *fh = fhat + (size_t)2 * N3 * cell, *mh = mhat + (size_t)2 * N3 * cell;
    #pragma omp parallel for collapse(2)
    for (int i = 0; i < N; i++)
      for (int j = 0; j < N; j++)
        for (int k = 0; k < N; k++) {
          double t0 = 0., t1 = 0.;
          for (int l = 0; l < N; l++) {
            const int x = i + H - l; if (x < 0 || x >= N) continue;
            for (int m = 0; m < N; m++) {
              const int y = j + H - m; if (y < 0 || y >= N) continue;
              for (int n = 0; n < N; n++) {
                const int z = k + H - n; if (z < 0 || z >= N) continue;
                const size_t w = n + (size_t)N * (m + N * l), b = z + (size_t)N * (y + N * x);
                const double *g = G7 + 7 * w;
                const double e1 = E[x], e2 = E[y], e3 = E[z];
                const double W = g[0] - (g[1] * e1 * e1 + g[2] * e2 * e2 + g[3] * e3 * e3 + g[4] * e1 * e2 + g[5] * e1 * e3 + g[6] * e2 * e3);
                const double ar = mh[2 * w], ai = mh[2 * w + 1], br = fh[2 * b], bi = fh[2 * b + 1];
                t0 += W * (ar * br - ai * bi);
                t1 += W * (ar * bi + ai * br);
              }
            }
          }
          const size_t o = (size_t)2 * (N3 * (size_t)cell + k + N * (j + N * i));
          q[o] = t0; q[o + 1] = t1;
        }
  }
  return 0;
}

// in-register N-point DFT of fc3.cuh against the definition
template <int N, int SIGN>
static void run_fftN(const double *in, double *out)
{
  double2 x[N];
  for (int i = 0; i < N; i++) x[i] = make_double2(in[2 * i], in[2 * i + 1]);
  fc3::fftN<N, SIGN, N>(x);
  for (int i = 0; i < N; i++) { out[2 * i] = x[i].x; out[2 * i + 1] = x[i].y; }
}
extern "C" int fc3_fftN(int N, int sign, const double *in, double *out)
{
  switch (N * sign) {
    case 8: run_fftN<8, 1>(in, out); return 0;
    case -8: run_fftN<8, -1>(in, out); return 0;
    case 16: run_fftN<16, 1>(in, out); return 0;
    case -16: run_fftN<16, -1>(in, out); return 0;
    case 24: run_fftN<24, 1>(in, out); return 0;
    case -24: run_fftN<24, -1>(in, out); return 0;
    case 32: run_fftN<32, 1>(in, out); return 0;
    case -32: run_fftN<32, -1>(in, out); return 0;
  }
  return 1;
}

// the line halves of F2Q at L = 16: in = 32 complex inputs (the non-zero two thirds of a 48-point line);
// fwd: out[h*24 + q] = X[2q + h];  inv: in = 48 complex X (same layout), out[n] = sum_k X[k] exp(+2 pi i n k / 48), n < 48,
// rebuilt from the two inverse halves as the kernel does (t_0 + t_1, t_0 - t_1)
extern "C" int fc3_half_lines(int inverse, const double *in, double *out)
{
  using namespace fc3;
  constexpr int L = 16, H = 24;
  if (!inverse) {
    double2 a0[L], a1[L];
    for (int l = 0; l < L; l++) { a0[l] = make_double2(in[2 * l], in[2 * l + 1]); a1[l] = make_double2(in[2 * (l + L)], in[2 * (l + L) + 1]); }
    for (int h = 0; h < 2; h++) {
      double2 y[H];
      fwd_half<L>(a0, a1, h, y);
      for (int q = 0; q < H; q++) { out[2 * (h * H + q)] = y[q].x; out[2 * (h * H + q) + 1] = y[q].y; }
    }
    return 0;
  }
  double2 t[2][H];
  for (int h = 0; h < 2; h++) {
    for (int q = 0; q < H; q++) t[h][q] = make_double2(in[2 * (h * H + q)], in[2 * (h * H + q) + 1]);
    inv_half<L>(t[h], h);
  }
  for (int n = 0; n < H; n++) {
    const double2 p = cadd(t[0][n], t[1][n]), m = csub(t[0][n], t[1][n]);
    out[2 * n] = p.x; out[2 * n + 1] = p.y; out[2 * (n + H)] = m.x; out[2 * (n + H) + 1] = m.y;
  }
  return 0;
}
