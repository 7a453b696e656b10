// Collision-path kernels (sm_100a): DG -> spectral sampling, shifted 3-D transforms, the
// weighted spectral convolution ComputeQ, the conservation projection, the RK stage updates and
// the Fourier -> DG projection.  Reference lines are cited per kernel; paths are relative to
// /root/reference/source.
#include "lpgpu_internal.h"
#include "fc3.cuh"

#define LP_LAUNCHED(c)                                  \
  do {                                                  \
    (c)->launches++;                                    \
    LP_CUDA(cudaGetLastError());                        \
  } while (0)
#define LP_TRY_RC(expr)                                 \
  do {                                                  \
    int rc_ = (expr);                                   \
    if (rc_ != LPGPU_OK) return rc_;                    \
  } while (0)

// ---------------------------------------------------------------------------------------------
// complex helpers (double2 = (re, im))
__device__ __forceinline__ double2 cmul(double2 a, double2 b)
{ return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ void cfma(double2 &acc, double2 a, double2 b)
{
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}

// ---------------------------------------------------------------------------------------------
// layout changes between the reference's AoS U[6k+l] and the device's plane-major
// U[(p*6 + c)*sv + j] (p = local x cell + 1; planes 0 and ncell+1 are x halos)
__global__ void k_aos_to_planes(const double *__restrict__ aos, double *__restrict__ planes, long long n, int sv)
{
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  long long cell = t / sv; int j = (int)(t % sv);
  const double *s = aos + 6 * t;
  double *d = planes + ((cell + 1) * 6) * (long long)sv + j;
  #pragma unroll
  for (int c = 0; c < 6; c++) d[(long long)c * sv] = s[c];
}
__global__ void k_planes_to_aos(const double *__restrict__ planes, double *__restrict__ aos, long long n, int sv)
{
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  long long cell = t / sv; int j = (int)(t % sv);
  const double *s = planes + ((cell + 1) * 6) * (long long)sv + j;
  double *d = aos + 6 * t;
  #pragma unroll
  for (int c = 0; c < 6; c++) d[c] = s[(long long)c * sv];
}
int lp_launch_aos_to_planes(lpgpu_ctx *c, const double *aos, double *planes)
{
  long long n = (long long)c->ncell * c->sv;
  k_aos_to_planes<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(aos, planes, n, c->sv);
  LP_LAUNCHED(c);
  return LPGPU_OK;
}
int lp_launch_planes_to_aos(lpgpu_ctx *c, const double *planes, double *aos)
{
  long long n = (long long)c->ncell * c->sv;
  k_planes_to_aos<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(planes, aos, n, c->sv);
  LP_LAUNCHED(c);
  return LPGPU_OK;
}

// ---------------------------------------------------------------------------------------------
// setInit_spectral (SetInit_1.cpp:396-436): evaluate the DG polynomial at the N^3 spectral nodes
__global__ void k_sample(const double *__restrict__ planes, double *__restrict__ f, const int *__restrict__ node_cell,
                         const double *__restrict__ node_xi, int N, int Nv, int sv, long long total)
{
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int N3 = N * N * N;
  long long cell = t / N3; int q = (int)(t % N3);
  int n = q % N, m = (q / N) % N, l = q / (N * N);
  int j = (node_cell[l] * Nv + node_cell[m]) * Nv + node_cell[n];
  double x1 = node_xi[l], x2 = node_xi[m], x3 = node_xi[n];
  const double *u = planes + ((cell + 1) * 6) * (long long)sv + j;
  f[t] = u[0] + u[2LL * sv] * x1 + u[3LL * sv] * x2 + u[4LL * sv] * x3 + u[5LL * sv] * (x1 * x1 + x2 * x2 + x3 * x3);
}
int lp_launch_sample(lpgpu_ctx *c, const double *planes, double *f, int ncell)
{
  long long total = (long long)ncell * c->N3;
  k_sample<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(planes, f, c->d_node_cell, c->d_node_xi, c->p.N, c->p.Nv, c->sv, total);
  LP_LAUNCHED(c);
  return LPGPU_OK;
}

// ---------------------------------------------------------------------------------------------
// Shifted transforms: fft3D (collisionRoutines_1.cpp:285-319) and FS (:363-398).  Both are
//   pre-phase (by s = i+j+k)  ->  unnormalised 3-D DFT (sign -1 / +1)  ->  post-phase (by (i,j,k))
// with the phase factors tabulated on the host from the reference's own double expressions
// (tables.cpp) and applied here with un-fused multiplies/subtractions, so the reference's rounding of
// the phases is reproduced; only the DFT itself (dense N-point sums out of shared memory, N <= 32, vs
// FFTW's butterflies) differs, at the 1e-16 level.  k_dft_jk covers axes (1,2) of one x-slab,
// k_dft_i axis 0 of one y-slab.  fft3D = jk (pre) then i (post); FS = i (pre) then jk (post + epilogue).
//
// EPI: 0 complex out; 1..3 = FS real part + RK stage update (RK4_Inhomo/RK4_Homo,
// collisionRoutines_1.cpp:910-941 / 1094-1123); 4 = FS real part stored as (re, 0).
__device__ __forceinline__ int bitrev(int v, int bits) { return (int)(__brev((unsigned)v) >> (32 - bits)); }
// N-point decimation-in-time FFT across N consecutive lanes (N = 2^k <= 32): lane l supplies element bitrev(l) and
// returns output l.  W[t] = W_N^t carries the sign of the transform.  All 32 lanes of the warp must call it.
__device__ __forceinline__ double2 lanes_fft(double2 v, int l, int N, const double2 *W)
{
  for (int h = 1; h < N; h <<= 1) {
    const bool up = (l & h) != 0;
    const double2 m = up ? cmul(v, W[(l & (h - 1)) * (N / (2 * h))]) : v;
    const double2 o = make_double2(__shfl_xor_sync(0xffffffffu, m.x, h), __shfl_xor_sync(0xffffffffu, m.y, h));
    v = up ? make_double2(o.x - m.x, o.y - m.y) : make_double2(o.x + m.x, o.y + m.y);
  }
  return v;
}

template <bool IN_REAL, int EPI, bool PRE, bool POST>
__global__ void __launch_bounds__(1024) k_dft_jk(const double *__restrict__ in, double *__restrict__ out,
                                                 const double2 *__restrict__ Wm, int N, int logn, PhaseTabs ph, FsEpilogue ep)
{
  extern __shared__ double2 sm2[];
  const int P = N + 1;
  double2 *X = sm2, *Y = X + N * P, *F = Y + N * P;
  const long long slab = blockIdx.x;                 // cell*N + i
  const int i = (int)(slab % N);
  const int tid = threadIdx.x, r = tid / N, cc = tid % N;
  const long long g = slab * N * N + tid;
  double2 x = IN_REAL ? make_double2(in[g], 0.) : reinterpret_cast<const double2 *>(in)[g];
  if (PRE) {
    x = phase_mul(ph.pre[i + r + cc], x);
    if (ph.wt) { const double fac = ph.c3 * ph.wt[i] * ph.wt[r] * ph.wt[cc]; x.x = __dmul_rn(fac, x.x); x.y = __dmul_rn(fac, x.y); }
  }
  X[r * P + cc] = x;
  F[r * P + cc] = Wm[tid];
  __syncthreads();
  double2 acc = make_double2(0., 0.);
  if (logn) {
    // power-of-two N: each line is a decimation-in-time FFT across N lanes (one element per lane, partners by
    // __shfl_xor); bit-reversed input index, natural output.  F row 1 holds the twiddles W_N^t.
    acc = lanes_fft(X[r * P + bitrev(cc, logn)], cc, N, F + P);            // axis 2: line = row r, lane = cc
    Y[r * P + cc] = acc;
    __syncthreads();
    acc = lanes_fft(Y[bitrev(cc, logn) * P + r], cc, N, F + P);            // axis 1: line = column r, lane = output row cc
    __syncthreads();
    X[cc * P + r] = acc;
    __syncthreads();
    acc = X[r * P + cc];
  } else {
    for (int j = 0; j < N; j++) cfma(acc, F[cc * P + j], X[r * P + j]);   // axis 2
    Y[r * P + cc] = acc;
    __syncthreads();
    acc = make_double2(0., 0.);
    for (int j = 0; j < N; j++) cfma(acc, F[r * P + j], Y[j * P + cc]);   // axis 1
  }
  if (POST) acc = phase_mul(ph.post[(i * N + r) * N + cc], acc);
  if (EPI == 0) {
    reinterpret_cast<double2 *>(out)[g] = acc;
  } else {
    const double Q = acc.x / ep.scaleL / ep.scale3;
    if (EPI == 4) reinterpret_cast<double2 *>(out)[g] = make_double2(Q, 0.);
    if (EPI == 1) { ep.Qv[g] = Q; ep.f1[g] = ep.f[g] + ep.dt * Q * ep.nu; }
    if (EPI == 2) ep.f1[g] = ep.f[g] + 0.5 * ep.dt * ep.Qv[g] * ep.nu + 0.5 * ep.dt * Q * ep.nu;
    if (EPI == 3) ep.f1[g] = ep.f[g] + 0.5 * ep.Qv[g] * ep.nu + 0.5 * Q * ep.nu;   // no dt: reference quirk (:940, :1122)
  }
}

// axis 0: block = (cell, j); slab X[i][k] = in[cell][i][j][k]
template <bool PRE, bool POST>
__global__ void __launch_bounds__(1024) k_dft_i(const double2 *__restrict__ in, double2 *__restrict__ out,
                                                const double2 *__restrict__ Wm, int N, int logn, PhaseTabs ph)
{
  extern __shared__ double2 sm2[];
  const int P = N + 1;
  double2 *X = sm2, *F = X + N * P;
  const long long cell = blockIdx.x / N; const int j = blockIdx.x % N;
  const int tid = threadIdx.x, i = tid / N, k = tid % N;
  const long long g = ((cell * N + i) * N + j) * N + k;
  double2 x = in[g];
  if (PRE) x = phase_mul(ph.pre[i + j + k], x);
  X[i * P + k] = x;
  F[i * P + k] = Wm[tid];
  __syncthreads();
  double2 acc = make_double2(0., 0.);
  if (logn) {
    // thread (i,k) takes lane k of the line "column i": transforms along the slab's first index
    acc = lanes_fft(X[bitrev(k, logn) * P + i], k, N, F + P);             // output index k of column i
    __syncthreads();
    X[k * P + i] = acc;
    __syncthreads();
    acc = X[i * P + k];
  } else {
    for (int a = 0; a < N; a++) cfma(acc, F[i * P + a], X[a * P + k]);
  }
  if (POST) acc = phase_mul(ph.post[(i * N + j) * N + k], acc);
  out[g] = acc;
}

// ---- register-resident lines (N = 8, 16, 24, 32): one thread owns a whole N-point line (fc3::fftN), a CTA of N
// threads owns one N x N slab.  k_tf_jk: coalesced load (+ pre-phase) -> shared, lines along k, lines along j,
// post-phase / epilogue, coalesced store.  k_tf_i: the line of thread k runs along i at fixed (j, k), so loads and
// stores are coalesced as they are and no shared memory is needed.
template <int EPI>
__device__ __forceinline__ void tf_epilogue(double2 acc, long long g, double *out, const FsEpilogue &ep)
{
  if (EPI == 0) {
    reinterpret_cast<double2 *>(out)[g] = acc;
  } else {
    // f and (in stages 2, 3) Qv are only read here: non-coherent loads, so that the loads of later lines are not held
    // behind the stores of earlier ones (one warp per slab: the memory latency is hidden by loads in flight, not by warps)
    const double Q = acc.x * ep.inv;
    if (EPI == 4) reinterpret_cast<double2 *>(out)[g] = make_double2(Q, 0.);
    if (EPI == 1) { ep.Qv[g] = Q; ep.f1[g] = __ldg(ep.f + g) + ep.dt * Q * ep.nu; }
    if (EPI == 2) ep.f1[g] = __ldg(ep.f + g) + 0.5 * ep.dt * __ldg(ep.Qv + g) * ep.nu + 0.5 * ep.dt * Q * ep.nu;
    if (EPI == 3) ep.f1[g] = __ldg(ep.f + g) + 0.5 * __ldg(ep.Qv + g) * ep.nu + 0.5 * Q * ep.nu;   // no dt: reference quirk (:940, :1122)
  }
}
template <int N, int SIGN, bool IN_REAL, int EPI, bool PRE, bool POST>
__global__ void __launch_bounds__(N) k_tf_jk(const double *__restrict__ in, double *__restrict__ out, PhaseTabs ph, FsEpilogue ep)
{
  constexpr int P = N + 1;
  __shared__ double2 X[N * P];
  const long long slab = blockIdx.x;                 // cell*N + i
  const int i = (int)(slab % N), t = threadIdx.x;
  #pragma unroll 16
  for (int j = 0; j < N; j++) {
    const long long g = slab * N * N + j * N + t;
    double2 x = IN_REAL ? make_double2(__ldg(in + g), 0.) : __ldg(reinterpret_cast<const double2 *>(in) + g);
    if (PRE) {
      x = phase_mul(__ldg(ph.pre + i + j + t), x);
      if (ph.wt) { const double fac = ph.c3 * ph.wt[i] * ph.wt[j] * ph.wt[t]; x.x = __dmul_rn(fac, x.x); x.y = __dmul_rn(fac, x.y); }
    }
    X[j * P + t] = x;
  }
  __syncthreads();
  double2 v[N];
  #pragma unroll
  for (int k = 0; k < N; k++) v[k] = X[t * P + k];          // row j = t
  fc3::fftN<N, SIGN, N>(v);
  #pragma unroll
  for (int k = 0; k < N; k++) X[t * P + k] = v[k];
  __syncthreads();
  #pragma unroll
  for (int j = 0; j < N; j++) v[j] = X[j * P + t];          // column k = t
  fc3::fftN<N, SIGN, N>(v);
  #pragma unroll
  for (int j = 0; j < N; j++) {
    double2 acc = v[j];
    if (POST) acc = phase_mul(__ldg(ph.post + (i * N + j) * N + t), acc);
    tf_epilogue<EPI>(acc, slab * N * N + j * N + t, out, ep);
  }
}
template <int N, int SIGN, bool PRE, bool POST>
__global__ void __launch_bounds__(N) k_tf_i(const double2 *__restrict__ in, double2 *__restrict__ out, PhaseTabs ph)
{
  const long long cell = blockIdx.x / N; const int j = blockIdx.x % N, t = threadIdx.x;
  double2 v[N];
  #pragma unroll
  for (int i = 0; i < N; i++) {
    v[i] = in[((cell * N + i) * N + j) * N + t];
    if (PRE) v[i] = phase_mul(ph.pre[i + j + t], v[i]);
  }
  fc3::fftN<N, SIGN, N>(v);
  #pragma unroll
  for (int i = 0; i < N; i++) {
    double2 acc = v[i];
    if (POST) acc = phase_mul(ph.post[(i * N + j) * N + t], acc);
    out[((cell * N + i) * N + j) * N + t] = acc;
  }
}
// FS's last two passes, the RK stage update and fft3D's first two passes of the NEXT stage in one kernel: both work on the
// slab (cell, i) and thread t owns the same elements (j, t) at the end of the inverse transform and at the start of the
// forward one, so the stage input f1 never goes to memory (it was 8 B written + 8 B read per node and a launch per stage).
// in/out: the FS spectrum after its i pass / the forward spectrum before its i pass; the same buffer, slab by slab.
template <int N, int EPI>
__global__ void __launch_bounds__(N, 12) k_tf_jk_fs_fft(const double2 *__restrict__ in, double2 *__restrict__ out, PhaseTabs inv, PhaseTabs fwd, FsEpilogue ep)
{
  constexpr int P = N + 1;
  __shared__ double2 X[N * P];
  const long long slab = blockIdx.x;                 // cell*N + i
  const int i = (int)(slab % N), t = threadIdx.x;
  // the epilogue's f and Qv rows: start them towards L2 now, they are needed two transform passes later
  for (int l = t; l < N * N / 16; l += N) {
    asm volatile("prefetch.global.L2 [%0];" :: "l"(ep.f + slab * N * N + l * 16));
    if (EPI != 1) asm volatile("prefetch.global.L2 [%0];" :: "l"(ep.Qv + slab * N * N + l * 16));
  }
  #pragma unroll 16
  for (int j = 0; j < N; j++) X[j * P + t] = __ldg(in + slab * N * N + j * N + t);
  __syncthreads();
  double2 v[N];
  #pragma unroll
  for (int k = 0; k < N; k++) v[k] = X[t * P + k];          // row j = t
  fc3::fftN<N, +1, N>(v);
  #pragma unroll
  for (int k = 0; k < N; k++) X[t * P + k] = v[k];
  __syncthreads();
  #pragma unroll
  for (int j = 0; j < N; j++) v[j] = X[j * P + t];          // column k = t
  fc3::fftN<N, +1, N>(v);
  // Column t of X is this thread's alone from the load above to the barrier below.  The column is parked there so that
  // the registers can hold the epilogue's operands (post-phase, f, Qv) for N/2 rows at a time: two memory round trips
  // per slab; fetched row by row next to the 2N live transform registers, every row waited for its own loads.
  #pragma unroll
  for (int j = 0; j < N; j++) X[j * P + t] = v[j];
  constexpr int H = N / 2;
  #pragma unroll
  for (int j0 = 0; j0 < N; j0 += H) {
    double2 pz[H]; double ff[H], qq[H];
    #pragma unroll
    for (int jj = 0; jj < H; jj++) {
      const int j = j0 + jj;
      const long long g = slab * N * N + j * N + t;
      pz[jj] = __ldg(inv.post + (i * N + j) * N + t);
      ff[jj] = __ldg(ep.f + g);
      qq[jj] = EPI == 1 ? 0. : __ldg(ep.Qv + g);
    }
    #pragma unroll
    for (int jj = 0; jj < H; jj++) {
      const int j = j0 + jj;
      const long long g = slab * N * N + j * N + t;
      const double2 acc = phase_mul(pz[jj], X[j * P + t]);
      const double Q = acc.x * ep.inv;
      double f1;
      if (EPI == 1) { ep.Qv[g] = Q; f1 = ff[jj] + ep.dt * Q * ep.nu; }
      else if (EPI == 2) f1 = ff[jj] + 0.5 * ep.dt * qq[jj] * ep.nu + 0.5 * ep.dt * Q * ep.nu;
      else f1 = ff[jj] + 0.5 * qq[jj] * ep.nu + 0.5 * Q * ep.nu;   // no dt: reference quirk (:940, :1122)
      // fft3D's pre-phase and quadrature weights on the new stage input, exactly as k_tf_jk<fwd> applies them
      double2 x = phase_mul(__ldg(fwd.pre + i + j + t), make_double2(f1, 0.));
      const double fac = fwd.c3 * fwd.wt[i] * fwd.wt[j] * fwd.wt[t];
      x.x = __dmul_rn(fac, x.x); x.y = __dmul_rn(fac, x.y);
      X[j * P + t] = x;
    }
  }
  __syncthreads();
  #pragma unroll
  for (int k = 0; k < N; k++) v[k] = X[t * P + k];
  fc3::fftN<N, -1, N>(v);
  #pragma unroll
  for (int k = 0; k < N; k++) X[t * P + k] = v[k];
  __syncthreads();
  #pragma unroll
  for (int j = 0; j < N; j++) v[j] = X[j * P + t];
  fc3::fftN<N, -1, N>(v);
  #pragma unroll
  for (int j = 0; j < N; j++) out[slab * N * N + j * N + t] = v[j];
}
static bool lp_reg_lines(int N)
{
  static const bool dense_only = getenv("LPGPU_DFT_DENSE") != nullptr;   // developer knob: the shared-memory DFT kernels
  return !dense_only && (N == 8 || N == 16 || N == 24 || N == 32);
}

// log2(N) when N is a power of two >= 8 and the block is whole warps (shuffle FFT lines), else 0 (dense N-point sums)
static int lp_log2_pow2(int N)
{
  static const bool dense_only = getenv("LPGPU_DFT_DENSE") != nullptr;   // developer knob
  if (dense_only || (N & (N - 1)) || N < 8 || N > 32) return 0;
  int l = 0;
  while ((1 << l) < N) l++;
  return l;
}
static size_t dft_smem(int N, int arrays) { return (size_t)arrays * N * (N + 1) * sizeof(double2); }

template <bool IN_REAL, int EPI, bool PRE, bool POST>
static int launch_jk(lpgpu_ctx *c, const double *in, double *out, const double *Wm, int B, PhaseTabs ph, FsEpilogue ep)
{
  const int N = c->p.N;
  if (lp_reg_lines(N)) {
    const bool fwd = Wm == c->d_Wfwd;
#define LP_TF_JK(NN)                                                                                                          \
    if (fwd) k_tf_jk<NN, -1, IN_REAL, EPI, PRE, POST><<<B * NN, NN, 0, c->stream>>>(in, out, ph, ep);                         \
    else k_tf_jk<NN, +1, IN_REAL, EPI, PRE, POST><<<B * NN, NN, 0, c->stream>>>(in, out, ph, ep)
    if (N == 32) { LP_TF_JK(32); } else if (N == 24) { LP_TF_JK(24); } else if (N == 16) { LP_TF_JK(16); } else { LP_TF_JK(8); }
#undef LP_TF_JK
    LP_LAUNCHED(c);
    return LPGPU_OK;
  }
  const size_t smem = dft_smem(N, 3);
  auto kern = k_dft_jk<IN_REAL, EPI, PRE, POST>;
  LP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<B * N, N * N, smem, c->stream>>>(in, out, reinterpret_cast<const double2 *>(Wm), N, lp_log2_pow2(N), ph, ep);
  LP_LAUNCHED(c);
  return LPGPU_OK;
}
template <bool PRE, bool POST>
static int launch_i(lpgpu_ctx *c, const double *in, double *out, const double *Wm, int B, PhaseTabs ph)
{
  const int N = c->p.N;
  if (lp_reg_lines(N)) {
    const bool fwd = Wm == c->d_Wfwd;
    const double2 *i2 = reinterpret_cast<const double2 *>(in);
    double2 *o2 = reinterpret_cast<double2 *>(out);
#define LP_TF_I(NN)                                                                                   \
    if (fwd) k_tf_i<NN, -1, PRE, POST><<<B * NN, NN, 0, c->stream>>>(i2, o2, ph);                     \
    else k_tf_i<NN, +1, PRE, POST><<<B * NN, NN, 0, c->stream>>>(i2, o2, ph)
    if (N == 32) { LP_TF_I(32); } else if (N == 24) { LP_TF_I(24); } else if (N == 16) { LP_TF_I(16); } else { LP_TF_I(8); }
#undef LP_TF_I
    LP_LAUNCHED(c);
    return LPGPU_OK;
  }
  const size_t smem = dft_smem(N, 2);
  auto kern = k_dft_i<PRE, POST>;
  LP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<B * N, N * N, smem, c->stream>>>(reinterpret_cast<const double2 *>(in), reinterpret_cast<double2 *>(out),
                                          reinterpret_cast<const double2 *>(Wm), N, lp_log2_pow2(N), ph);
  LP_LAUNCHED(c);
  return LPGPU_OK;
}

// first pass of fft3D only (pre-phase, transforms along k and j) into c->d_tmp; the fused ComputeQ (fftconv.cu, F1)
// runs the pass along i and the post-phase itself
int lp_launch_fft3d_jk(lpgpu_ctx *c, const double *in, bool in_real, int B)
{
  FsEpilogue ep = {};
  PhaseTabs pre = {reinterpret_cast<const double2 *>(c->d_pre_fwd), nullptr, c->d_wt, c->tab.c3_fwd};
  return in_real ? launch_jk<true, 0, true, false>(c, in, c->d_tmp, c->d_Wfwd, B, pre, ep)
                 : launch_jk<false, 0, true, false>(c, in, c->d_tmp, c->d_Wfwd, B, pre, ep);
}
int lp_launch_fft3d(lpgpu_ctx *c, const double *in, bool in_real, double *out, int B)
{
  PhaseTabs post = {nullptr, reinterpret_cast<const double2 *>(c->d_post_fwd), nullptr, 0.};
  int rc = lp_launch_fft3d_jk(c, in, in_real, B);
  if (rc) return rc;
  return launch_i<false, true>(c, c->d_tmp, out, c->d_Wfwd, B, post);
}

int lp_launch_fs(lpgpu_ctx *c, const double *q, int mode, double *out_complex, int B)
{
  FsEpilogue ep;
  ep.scaleL = c->tab.scaleL; ep.scale3 = c->tab.scale3; ep.inv = 1. / c->tab.scaleL / c->tab.scale3;
  ep.dt = c->p.dt; ep.nu = c->p.nu; ep.f = c->d_f; ep.Qv = c->d_Qv; ep.f1 = c->d_f1;
  PhaseTabs pre = {reinterpret_cast<const double2 *>(c->d_pre_inv), nullptr, nullptr, 0.};
  PhaseTabs post = {nullptr, reinterpret_cast<const double2 *>(c->d_post_inv), nullptr, 0.};
  int rc = launch_i<true, false>(c, q, c->d_tmp, c->d_Winv, B, pre);
  if (rc) return rc;
  switch (mode) {
    case 0: return launch_jk<false, 4, false, true>(c, c->d_tmp, out_complex, c->d_Winv, B, post, ep);
    case 1: return launch_jk<false, 1, false, true>(c, c->d_tmp, nullptr, c->d_Winv, B, post, ep);
    case 2: return launch_jk<false, 2, false, true>(c, c->d_tmp, nullptr, c->d_Winv, B, post, ep);
    case 3: return launch_jk<false, 3, false, true>(c, c->d_tmp, nullptr, c->d_Winv, B, post, ep);
  }
  lp_set_error("lp_launch_fs: bad mode");
  return LPGPU_EINVAL;
}

// ---------------------------------------------------------------------------------------------
// ComputeQ (collisionRoutines_1.cpp:691-774):
//   Qhat[xi] = sum_{omega in win(xi)} h_eta^3 wt(omega) gHat3(xi, omega) fhat[omega] fhat[xi + N/2 - omega]
// Variant 1 ("simple"): one thread per xi, weight rebuilt per pair from the 7 folded symbols.
// Kept as the on-device cross-check for the tiled kernel.
__device__ __forceinline__ void lp_window(int N, int i, int &s, int &e)
{
  if (i < N / 2) { s = 0; e = i + N / 2 + 1; } else { s = i - N / 2 + 1; e = N; }
}
__global__ void __launch_bounds__(128) k_computeQ_simple(const double2 *__restrict__ fhat, double2 *__restrict__ q,
                                                         const double *__restrict__ G, const double *__restrict__ eta,
                                                         int N, long long total)
{
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int N3 = N * N * N, H = N / 2;
  const long long cell = t / N3; const int xi = (int)(t % N3);
  const int k = xi % N, j = (xi / N) % N, i = xi / (N * N);
  const double2 *fh = fhat + cell * N3;
  int si, ei, sj, ej, sk, ek;
  lp_window(N, i, si, ei); lp_window(N, j, sj, ej); lp_window(N, k, sk, ek);
  double t0 = 0., t1 = 0.;
  for (int l = si; l < ei; l++) {
    const int x = i + H - l; const double e1 = eta[i] - eta[l];
    for (int m = sj; m < ej; m++) {
      const int y = j + H - m; const double e2 = eta[j] - eta[m];
      for (int n = sk; n < ek; n++) {
        const int z = k + H - n; const double e3 = eta[k] - eta[n];
        const int w = n + N * (m + N * l);
        const double *g = G + 7LL * w;
        const double W = g[0] - (g[1] * e1 * e1 + g[2] * e2 * e2 + g[3] * e3 * e3 + g[4] * e1 * e2 + g[5] * e1 * e3 + g[6] * e2 * e3);
        const double2 a = fh[w], b = fh[z + N * (y + N * x)];
        t0 += W * (a.x * b.x - a.y * b.y);
        t1 += W * (a.x * b.y + a.y * b.x);
      }
    }
  }
  q[t] = make_double2(t0, t1);
}

int lp_launch_computeQ_tiled(lpgpu_ctx *c, const double *fhat, double *q, int B);   // computeq.cu


int lp_launch_computeQ(lpgpu_ctx *c, const double *fhat, double *q, int B)
{
  // 0 = fastest validated path (FFT convolutions when N is a power of two, else the tiled direct sum),
  // 1 = simple direct kernel, 2 = FFT convolutions, 3 = tiled direct sum
  const int variant = c->p.computeq_variant;
  if (c->p.linear_landau && !c->have_mhat) { lp_set_error("LinearLandau: call lpgpu_set_maxwellian first"); return LPGPU_EINVAL; }
  if (variant == 0 || variant == 2) {
    int rc = lp_launch_computeQ_fftconv(c, fhat, q, B, false, nullptr);
    if (rc != -1) return rc;   // -1: N is not a power of two -> direct kernels
  }
  if (c->p.linear_landau) { lp_set_error("LinearLandau needs the FFT-convolution pipeline"); return LPGPU_EINVAL; }
  if (variant != 1) {
    int rc = lp_launch_computeQ_tiled(c, fhat, q, B);
    if (rc != -1) return rc;   // -1: size not covered by the tiled kernel -> simple kernel
  }
  long long total = (long long)B * c->N3;
  k_computeQ_simple<<<(unsigned)((total + 127) / 128), 128, 0, c->stream>>>(
      reinterpret_cast<const double2 *>(fhat), reinterpret_cast<double2 *>(q), c->d_G, c->d_eta, c->p.N, total);
  LP_LAUNCHED(c);
  return LPGPU_OK;
}

// ---------------------------------------------------------------------------------------------
// conserveAllMoments_Normal + solveWithCCt (conservationRoutines.cpp:131-156, 32-58).
// One block per cell: five dot products (fixed-order tree reduction), lambda = CCt^-1-applied,
// rank-5 correction.
__global__ void __launch_bounds__(512) k_conserve(double2 *__restrict__ q, const double *__restrict__ C5,
                                                  const double *__restrict__ CCt, int N3)
{
  __shared__ double red[5][16];
  __shared__ double lam[5];
  double2 *qc = q + (long long)blockIdx.x * N3;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  double s[5] = {0., 0., 0., 0., 0.};
  for (int idx = tid; idx < N3; idx += blockDim.x) {
    const double2 v = qc[idx];
    s[0] += v.x * C5[idx];
    s[1] += v.y * C5[N3 + idx];
    s[2] += v.y * C5[2 * N3 + idx];
    s[3] += v.y * C5[3 * N3 + idx];
    s[4] += v.x * C5[4 * N3 + idx];
  }
  #pragma unroll
  for (int m = 0; m < 5; m++) {
    double v = s[m];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) red[m][wid] = v;
  }
  __syncthreads();
  if (tid == 0) {
    double tot[5];
    for (int m = 0; m < 5; m++) { double v = 0.; for (int w = 0; w < nw; w++) v += red[m][w]; tot[m] = v; }
    for (int a = 0; a < 5; a++) { double v = 0.; for (int b = 0; b < 5; b++) v += CCt[b + a * 5] * tot[b]; lam[a] = v; }
  }
  __syncthreads();
  const double b0 = lam[0], b1 = lam[1], b2 = lam[2], b3 = lam[3], b4 = lam[4];
  for (int idx = tid; idx < N3; idx += blockDim.x) {
    double2 v = qc[idx];
    v.x -= (C5[idx] * b0 + C5[4 * N3 + idx] * b4);
    v.y -= (C5[N3 + idx] * b1 + C5[2 * N3 + idx] * b2 + C5[3 * N3 + idx] * b3);
    qc[idx] = v;
  }
}
// Two-kernel form for large spectra: CH blocks per cell compute partial dot products, then CH blocks per
// cell fold them (fixed order), apply CCt and correct their slice.  One block per cell leaves 116 of 148
// SMs idle when a GPU holds 32 cells.
#define LP_CONS_CH 8
#define LP_APPLY_CH 32   // blocks per cell of the correction pass (independent of the number of partial sums folded)
__global__ void __launch_bounds__(256) k_conserve_dots(const double2 *__restrict__ q, const double *__restrict__ C5,
                                                       double *__restrict__ part, int N3)
{
  __shared__ double red[5][8];
  const long long cell = blockIdx.x; const int ch = blockIdx.y;
  const double2 *qc = q + cell * N3;
  const int per = (N3 + LP_CONS_CH - 1) / LP_CONS_CH, lo = ch * per, hi = min(N3, lo + per);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  double s[5] = {0., 0., 0., 0., 0.};
  for (int idx = lo + tid; idx < hi; idx += blockDim.x) {
    const double2 v = qc[idx];
    s[0] += v.x * C5[idx]; s[1] += v.y * C5[N3 + idx]; s[2] += v.y * C5[2 * N3 + idx];
    s[3] += v.y * C5[3 * N3 + idx]; s[4] += v.x * C5[4 * N3 + idx];
  }
  #pragma unroll
  for (int m = 0; m < 5; m++) {
    double v = s[m];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) red[m][wid] = v;
  }
  __syncthreads();
  if (tid < 5) { double v = 0.; for (int w = 0; w < 8; w++) v += red[tid][w]; part[(cell * LP_CONS_CH + ch) * 5 + tid] = v; }
}
__global__ void __launch_bounds__(256) k_conserve_apply(double2 *__restrict__ q, const double *__restrict__ C5,
                                                        const double *__restrict__ CCt, const double *__restrict__ part, int N3, int nparts)
{
  __shared__ double lam[5];
  const long long cell = blockIdx.x; const int ch = blockIdx.y;
  if (threadIdx.x == 0) {
    double tot[5];
    for (int m = 0; m < 5; m++) { double v = 0.; for (int k = 0; k < nparts; k++) v += part[(cell * nparts + k) * 5 + m]; tot[m] = v; }
    for (int a = 0; a < 5; a++) { double v = 0.; for (int b = 0; b < 5; b++) v += CCt[b + a * 5] * tot[b]; lam[a] = v; }
  }
  __syncthreads();
  const double b0 = lam[0], b1 = lam[1], b2 = lam[2], b3 = lam[3], b4 = lam[4];
  double2 *qc = q + cell * N3;
  const int per = (N3 + gridDim.y - 1) / gridDim.y, lo = ch * per, hi = min(N3, lo + per);
  #pragma unroll 4
  for (int idx = lo + threadIdx.x; idx < hi; idx += blockDim.x) {
    double2 v = qc[idx];
    v.x -= (C5[idx] * b0 + C5[4 * N3 + idx] * b4);
    v.y -= (C5[N3 + idx] * b1 + C5[2 * N3 + idx] * b2 + C5[3 * N3 + idx] * b3);
    qc[idx] = v;
  }
}
int lp_launch_conserve(lpgpu_ctx *c, double *q, int B)
{
  if (c->N3 < 4096) {
    k_conserve<<<B, 512, 0, c->stream>>>(reinterpret_cast<double2 *>(q), c->d_C5, c->d_CCt, c->N3);
    LP_LAUNCHED(c);
    return LPGPU_OK;
  }
  k_conserve_dots<<<dim3(B, LP_CONS_CH), 256, 0, c->stream>>>(reinterpret_cast<const double2 *>(q), c->d_C5, c->d_lam, c->N3);
  LP_LAUNCHED(c);
  k_conserve_apply<<<dim3(B, LP_APPLY_CH), 256, 0, c->stream>>>(reinterpret_cast<double2 *>(q), c->d_C5, c->d_CCt, c->d_lam, c->N3, LP_CONS_CH);
  LP_LAUNCHED(c);
  return LPGPU_OK;
}

// ---------------------------------------------------------------------------------------------
// FullandLinear variant (reference test 3).  ComputeQ_FandL (collisionRoutines_1.cpp:605-689): next to the
// quadratic sum, qHat_linear[xi] = sum_w h_eta^3 wt(w) scale3 gHat3_linear(xi, w) fhat[xi + N/2 - w]; qHat receives
// both.  One thread per xi: the direct form, kept for sizes without an FFT-convolution pipeline and as the cross-check.
__global__ void __launch_bounds__(128) k_computeQ_fandl(const double2 *__restrict__ fhat, double2 *__restrict__ q, double2 *__restrict__ ql,
                                                        const double *__restrict__ G, const double *__restrict__ Gl,
                                                        const double *__restrict__ eta, int N, double scale3, long long total)
{
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int N3 = N * N * N, H = N / 2;
  const long long cell = t / N3; const int xi = (int)(t % N3);
  const int k = xi % N, j = (xi / N) % N, i = xi / (N * N);
  const double2 *fh = fhat + cell * N3;
  int si, ei, sj, ej, sk, ek;
  lp_window(N, i, si, ei); lp_window(N, j, sj, ej); lp_window(N, k, sk, ek);
  double t0 = 0., t1 = 0., t01 = 0., t11 = 0.;
  for (int l = si; l < ei; l++) {
    const int x = i + H - l; const double e1 = eta[i] - eta[l];
    for (int m = sj; m < ej; m++) {
      const int y = j + H - m; const double e2 = eta[j] - eta[m];
      for (int n = sk; n < ek; n++) {
        const int z = k + H - n; const double e3 = eta[k] - eta[n];
        const int w = n + N * (m + N * l);
        const double *g = G + 7LL * w, *gl = Gl + 3LL * w;
        const double quad = g[1] * e1 * e1 + g[2] * e2 * e2 + g[3] * e3 * e3 + g[4] * e1 * e2 + g[5] * e1 * e3 + g[6] * e2 * e3;
        const double W = g[0] - quad, W1 = -(scale3 * quad + gl[0] * e1 + gl[1] * e2 + gl[2] * e3);
        const double2 a = fh[w], b = fh[z + N * (y + N * x)];
        const double l0 = W1 * b.x, l1 = W1 * b.y;
        t0 += W * (a.x * b.x - a.y * b.y) + l0;
        t1 += W * (a.x * b.y + a.y * b.x) + l1;
        t01 += l0; t11 += l1;
      }
    }
  }
  q[t] = make_double2(t0, t1);
  ql[t] = make_double2(t01, t11);
}
// ComputeQ_FandL through the FFT-convolution pipeline (N = 8, 16, 24, 32).  With e = xi - omega,
//   gHat3_linear(xi, omega) = - sum_ij S_ij(omega) e_i e_j - sum_j (sum_i S_ij(omega) omega_i) e_j      (collisionRoutines_1.cpp:193-218)
// and the sum ComputeQ_FandL takes over omega has NO fhat(omega) factor (:662), so qHat_linear is a sum of convolutions of
// FIXED symbols with monomials of e times fhat -- the pipeline with a first factor of ones instead of fhat:
//   pass A   symbols {0, scale3 G_1 .. scale3 G_6}, second factor fhat:         - scale3 sum_p G_p mono_p(e)
//   pass B_j symbols {-Gl_j, 0 .. 0}, second factor E_j fhat (j = 1, 2, 3):     - Gl_j e_j
// qHat = ComputeQ(f) + qHat_linear.  Five pipeline passes instead of one O(N^6) sum per evaluation.
__global__ void k_mul_E(const double2 *__restrict__ fhat, double2 *__restrict__ out, const double *__restrict__ E, int N, int axis, long long total)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int w = (int)(t % ((long long)N * N * N));
  const int idx = axis == 0 ? w / (N * N) : axis == 1 ? (w / N) % N : w % N;
  const double e = E[idx];
  const double2 v = fhat[t];
  out[t] = make_double2(e * v.x, e * v.y);
}
__global__ void k_add_to(double2 *__restrict__ acc, const double2 *__restrict__ x, long long total)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < total) { const double2 a = acc[t], b = x[t]; acc[t] = make_double2(a.x + b.x, a.y + b.y); }
}
int lp_launch_computeQ_fandl(lpgpu_ctx *c, const double *fhat, double *q, double *ql, int B)
{
  const long long total = (long long)B * c->N3;
  const bool direct_only = c->p.computeq_variant == 1 || c->p.computeq_variant == 3;   // the O(N^6) kernel on request (cross-check)
  if (!direct_only && c->d_GtLin && lp_fc3_available(c)) {
    const unsigned grid = (unsigned)((total + 255) / 256);
    const size_t tab = (size_t)7 * c->N3;
    const double *E = c->d_Etab + LP_ETAB_PAD;
    LP_TRY_RC(lp_launch_computeQ_fftconv(c, fhat, q, B, false, nullptr));
    LP_TRY_RC(lp_launch_fftconv_with(c, fhat, ql, B, c->d_GtLin, c->d_ones, 0));
    for (int j = 0; j < 3; j++) {
      k_mul_E<<<grid, 256, 0, c->stream>>>(reinterpret_cast<const double2 *>(fhat), reinterpret_cast<double2 *>(c->d_fl_g), E, c->p.N, j, total);
      LP_LAUNCHED(c);
      LP_TRY_RC(lp_launch_fftconv_with(c, c->d_fl_g, c->d_fl_tmp, B, c->d_GtLin + (1 + j) * tab, c->d_ones, 0));
      k_add_to<<<grid, 256, 0, c->stream>>>(reinterpret_cast<double2 *>(ql), reinterpret_cast<const double2 *>(c->d_fl_tmp), total);
      LP_LAUNCHED(c);
    }
    k_add_to<<<grid, 256, 0, c->stream>>>(reinterpret_cast<double2 *>(q), reinterpret_cast<const double2 *>(ql), total);
    LP_LAUNCHED(c);
    return LPGPU_OK;
  }
  k_computeQ_fandl<<<(unsigned)((total + 127) / 128), 128, 0, c->stream>>>(
      reinterpret_cast<const double2 *>(fhat), reinterpret_cast<double2 *>(q), reinterpret_cast<double2 *>(ql), c->d_G, c->d_Gl, c->d_eta,
      c->p.N, c->tab.scale3, total);
  LP_LAUNCHED(c);
  return LPGPU_OK;
}
// conserveAllMoments_FandL (conservationRoutines.cpp:102-129): five rows on qHat, mass and energy rows on qHat_linear,
// then qHat += qHat_linear (first loop of RK4_FandL_*, collisionRoutines_1.cpp:806-810).  One block per cell.
__global__ void __launch_bounds__(512) k_conserve_fandl(double2 *__restrict__ q, double2 *__restrict__ ql, const double *__restrict__ C5,
                                                        const double *__restrict__ CCt, const double *__restrict__ CCt_lin, int N3)
{
  __shared__ double red[7][16];
  __shared__ double lam[7];
  double2 *qc = q + (long long)blockIdx.x * N3, *qlc = ql + (long long)blockIdx.x * N3;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  double s[7] = {0., 0., 0., 0., 0., 0., 0.};
  for (int idx = tid; idx < N3; idx += blockDim.x) {
    const double2 v = qc[idx], u = qlc[idx];
    s[0] += v.x * C5[idx]; s[1] += v.y * C5[N3 + idx]; s[2] += v.y * C5[2 * N3 + idx];
    s[3] += v.y * C5[3 * N3 + idx]; s[4] += v.x * C5[4 * N3 + idx];
    s[5] += u.x * C5[idx]; s[6] += u.x * C5[4 * N3 + idx];
  }
  #pragma unroll
  for (int m = 0; m < 7; m++) {
    double v = s[m];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) red[m][wid] = v;
  }
  __syncthreads();
  if (tid == 0) {
    double tot[7];
    for (int m = 0; m < 7; m++) { double v = 0.; for (int w = 0; w < nw; w++) v += red[m][w]; tot[m] = v; }
    for (int a = 0; a < 5; a++) { double v = 0.; for (int b = 0; b < 5; b++) v += CCt[b + a * 5] * tot[b]; lam[a] = v; }
    for (int a = 0; a < 2; a++) lam[5 + a] = CCt_lin[0 + a * 2] * tot[5] + CCt_lin[1 + a * 2] * tot[6];
  }
  __syncthreads();
  for (int idx = tid; idx < N3; idx += blockDim.x) {
    double2 v = qc[idx], u = qlc[idx];
    v.x -= (C5[idx] * lam[0] + C5[4 * N3 + idx] * lam[4]);
    v.y -= (C5[N3 + idx] * lam[1] + C5[2 * N3 + idx] * lam[2] + C5[3 * N3 + idx] * lam[3]);
    u.x -= (C5[idx] * lam[5] + C5[4 * N3 + idx] * lam[6]);
    qlc[idx] = u;
    qc[idx] = make_double2(v.x + u.x, v.y + u.y);
  }
}
int lp_launch_conserve_fandl(lpgpu_ctx *c, double *q, double *ql, int B)
{
  k_conserve_fandl<<<B, 512, 0, c->stream>>>(reinterpret_cast<double2 *>(q), reinterpret_cast<double2 *>(ql), c->d_C5, c->d_CCt, c->d_CCt_lin, c->N3);
  LP_LAUNCHED(c);
  return LPGPU_OK;
}

int lp_launch_conserve_from_parts(lpgpu_ctx *c, double *q, const double *part, int B)
{
  k_conserve_apply<<<dim3(B, LP_APPLY_CH), 256, 0, c->stream>>>(reinterpret_cast<double2 *>(q), c->d_C5, c->d_CCt, part, c->N3, c->p.N);
  LP_LAUNCHED(c);
  return LPGPU_OK;
}

// conserveAllMoments_Normal folded into the first pass of FS: fold the partial dot products (fixed order), lambda =
// CCt-applied, correct this thread's line of q in place (the projection reads the conserved spectra later), then
// pre-phase and transform along i as k_tf_i does.
template <int N>
__global__ void __launch_bounds__(N) k_tf_i_conserve(double2 *__restrict__ q, double2 *__restrict__ out, PhaseTabs ph,
                                                     const double *__restrict__ C5, const double *__restrict__ CCt, const double *__restrict__ part)
{
  __shared__ double tot[5];
  const long long cell = blockIdx.x / N; const int j = blockIdx.x % N, t = threadIdx.x;
  constexpr int N3 = N * N * N;
  if (t < 5) { double v = 0.; for (int k = 0; k < N; k++) v += part[(cell * N + k) * 5 + t]; tot[t] = v; }
  __syncthreads();
  double lam[5];
  #pragma unroll
  for (int a = 0; a < 5; a++) { double v = 0.; for (int b = 0; b < 5; b++) v += CCt[b + a * 5] * tot[b]; lam[a] = v; }
  double2 v[N];
  #pragma unroll
  for (int i = 0; i < N; i++) {
    const int idx = (i * N + j) * N + t;
    double2 x = q[cell * N3 + idx];
    x.x -= (C5[idx] * lam[0] + C5[4 * N3 + idx] * lam[4]);
    x.y -= (C5[N3 + idx] * lam[1] + C5[2 * N3 + idx] * lam[2] + C5[3 * N3 + idx] * lam[3]);
    q[cell * N3 + idx] = x;
    v[i] = phase_mul(ph.pre[i + j + t], x);
  }
  fc3::fftN<N, +1, N>(v);
  #pragma unroll
  for (int i = 0; i < N; i++) out[cell * N3 + (i * N + j) * N + t] = v[i];
}
static int launch_fs_second(lpgpu_ctx *c, int mode, double *out_complex, int B, FsEpilogue ep)
{
  PhaseTabs post = {nullptr, reinterpret_cast<const double2 *>(c->d_post_inv), nullptr, 0.};
  switch (mode) {
    case 0: return launch_jk<false, 4, false, true>(c, c->d_tmp, out_complex, c->d_Winv, B, post, ep);
    case 1: return launch_jk<false, 1, false, true>(c, c->d_tmp, nullptr, c->d_Winv, B, post, ep);
    case 2: return launch_jk<false, 2, false, true>(c, c->d_tmp, nullptr, c->d_Winv, B, post, ep);
    case 3: return launch_jk<false, 3, false, true>(c, c->d_tmp, nullptr, c->d_Winv, B, post, ep);
  }
  lp_set_error("lp_launch_fs: bad mode");
  return LPGPU_EINVAL;
}
// next_fwd: the caller's next kernel is the fused ComputeQ on c->d_tmp (the next RK stage): run FS's last passes, the stage
// update and fft3D's first passes as one kernel and leave the forward spectrum in c->d_tmp (f1 is not stored)
int lp_launch_fs_conserving(lpgpu_ctx *c, double *q, const double *part, int mode, int B, bool next_fwd)
{
  const int N = c->p.N;
  if (!lp_reg_lines(N)) {          // dense transforms: correction and FS as separate launches
    int rc = lp_launch_conserve_from_parts(c, q, part, B);
    if (!rc) rc = lp_launch_fs(c, q, mode, nullptr, B);
    if (rc != LPGPU_OK || !next_fwd) return rc;
    return lp_launch_fft3d_jk(c, c->d_f1, true, B);
  }
  FsEpilogue ep;
  ep.scaleL = c->tab.scaleL; ep.scale3 = c->tab.scale3; ep.inv = 1. / c->tab.scaleL / c->tab.scale3;
  ep.dt = c->p.dt; ep.nu = c->p.nu; ep.f = c->d_f; ep.Qv = c->d_Qv; ep.f1 = c->d_f1;
  PhaseTabs pre = {reinterpret_cast<const double2 *>(c->d_pre_inv), nullptr, nullptr, 0.};
  double2 *q2 = reinterpret_cast<double2 *>(q), *o2 = reinterpret_cast<double2 *>(c->d_tmp);
  if (N == 32) k_tf_i_conserve<32><<<B * 32, 32, 0, c->stream>>>(q2, o2, pre, c->d_C5, c->d_CCt, part);
  else if (N == 24) k_tf_i_conserve<24><<<B * 24, 24, 0, c->stream>>>(q2, o2, pre, c->d_C5, c->d_CCt, part);
  else if (N == 16) k_tf_i_conserve<16><<<B * 16, 16, 0, c->stream>>>(q2, o2, pre, c->d_C5, c->d_CCt, part);
  else k_tf_i_conserve<8><<<B * 8, 8, 0, c->stream>>>(q2, o2, pre, c->d_C5, c->d_CCt, part);
  LP_LAUNCHED(c);
  static const bool unfused = getenv("LPGPU_FS_UNFUSED") != nullptr;   // developer knob
  if (next_fwd && !unfused && mode >= 1 && mode <= 3) {
    PhaseTabs inv = {nullptr, reinterpret_cast<const double2 *>(c->d_post_inv), nullptr, 0.};
    PhaseTabs fwd = {reinterpret_cast<const double2 *>(c->d_pre_fwd), nullptr, c->d_wt, c->tab.c3_fwd};
    // (twelve one-warp CTAs per SM need the whole shared-memory carve-out)
#define LP_FSFFT(NN, EE) do { cudaFuncSetAttribute(k_tf_jk_fs_fft<NN, EE>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); \
                              k_tf_jk_fs_fft<NN, EE><<<B * NN, NN, 0, c->stream>>>(o2, o2, inv, fwd, ep); } while (0)
#define LP_FSFFT_N(NN) do { if (mode == 1) LP_FSFFT(NN, 1); else if (mode == 2) LP_FSFFT(NN, 2); else LP_FSFFT(NN, 3); } while (0)
    if (N == 32) LP_FSFFT_N(32); else if (N == 24) LP_FSFFT_N(24); else if (N == 16) LP_FSFFT_N(16); else LP_FSFFT_N(8);
#undef LP_FSFFT_N
#undef LP_FSFFT
    LP_LAUNCHED(c);
    return LPGPU_OK;
  }
  int rc = launch_fs_second(c, mode, nullptr, B, ep);
  if (rc != LPGPU_OK || !next_fwd) return rc;
  return lp_launch_fft3d_jk(c, c->d_f1, true, B);
}

// ---------------------------------------------------------------------------------------------
// Fourier -> DG projection: the kt loop of RK4_Inhomo / RK4_Homo (collisionRoutines_1.cpp:946-984,
// 1128-1166).  IntModes (:408-562) factorises per dimension into T, M, S (tables.cpp), so
//   tp_l[j1,j2,j3] = Re sum_{k1,k2,k3} IntM_l(k,j) Qc[k],   Qc = nu (q0/2 + (q1+q2+q3)/6)
// is three chained 1-D contractions.  k_project_slab contracts k3 then k2 for one k1 slab;
// k_project_final contracts k1 and applies the DG update in place on U (LP_ompi.cpp:727-753 scatter folded).
__global__ void __launch_bounds__(1024) k_project_slab(const double2 *__restrict__ q0, const double2 *__restrict__ q1,
                                                       const double2 *__restrict__ q2, const double2 *__restrict__ q3,
                                                       double2 *__restrict__ Bbuf, const double2 *__restrict__ tT,
                                                       const double2 *__restrict__ tM, const double2 *__restrict__ tS,
                                                       int N, int Nv, double nu)
{
  extern __shared__ double2 sm2[];
  const int P = N + 1, PV = Nv + 1;
  double2 *Xs = sm2;                 // [N][P]
  double2 *As = Xs + N * P;          // [3][N][PV]
  const long long slab = blockIdx.x; // cell*N + k1
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int t = tid; t < N * N; t += nt) {
    const long long g = slab * N * N + t;
    const double2 a = q0[g], b = q1[g], c = q2[g], d = q3[g];
    Xs[(t / N) * P + (t % N)] = make_double2(nu * (0.5 * a.x + (b.x + c.x + d.x) / 6.), nu * (0.5 * a.y + (b.y + c.y + d.y) / 6.));
  }
  __syncthreads();
  for (int t = tid; t < N * Nv; t += nt) {
    const int k2 = t / Nv, j3 = t % Nv;
    double2 aT = make_double2(0., 0.), aM = aT, aS = aT;
    for (int k3 = 0; k3 < N; k3++) {
      const double2 x = Xs[k2 * P + k3];
      cfma(aT, tT[k3 * Nv + j3], x); cfma(aM, tM[k3 * Nv + j3], x); cfma(aS, tS[k3 * Nv + j3], x);
    }
    As[(0 * N + k2) * PV + j3] = aT; As[(1 * N + k2) * PV + j3] = aM; As[(2 * N + k2) * PV + j3] = aS;
  }
  __syncthreads();
  const int Pq = Nv * Nv;
  for (int t = tid; t < Pq; t += nt) {
    const int j2 = t / Nv, j3 = t % Nv;
    double2 bTT = make_double2(0., 0.), bMT = bTT, bTM = bTT, bS = bTT;
    for (int k2 = 0; k2 < N; k2++) {
      const double2 T2 = tT[k2 * Nv + j2], M2 = tM[k2 * Nv + j2], S2 = tS[k2 * Nv + j2];
      const double2 at = As[(0 * N + k2) * PV + j3], am = As[(1 * N + k2) * PV + j3], as = As[(2 * N + k2) * PV + j3];
      cfma(bTT, T2, at); cfma(bMT, M2, at); cfma(bTM, T2, am); cfma(bS, S2, at); cfma(bS, T2, as);
    }
    double2 *o = Bbuf + slab * 4 * Pq + t;
    o[0] = bTT; o[Pq] = bMT; o[2 * Pq] = bTM; o[3 * Pq] = bS;
  }
}

__global__ void __launch_bounds__(256) k_project_final(const double2 *__restrict__ Bbuf, double *__restrict__ planes,
                                                       const double2 *__restrict__ tT, const double2 *__restrict__ tM,
                                                       const double2 *__restrict__ tS, int N, int Nv, int sv, double dt,
                                                       double scalev, double scaleL, double scale3)
{
  const int Pq = Nv * Nv;
  const long long cell = blockIdx.y; const int j1 = blockIdx.z;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Pq) return;
  double tp0 = 0., tp2 = 0., tp3 = 0., tp4 = 0., tp5 = 0.;
  for (int k1 = 0; k1 < N; k1++) {
    const double2 T1 = tT[k1 * Nv + j1], M1 = tM[k1 * Nv + j1], S1 = tS[k1 * Nv + j1];
    const double2 *b = Bbuf + (cell * N + k1) * 4 * Pq + p;
    const double2 btt = b[0], bmt = b[Pq], btm = b[2 * Pq], bs = b[3 * Pq];
    tp0 += T1.x * btt.x - T1.y * btt.y;
    tp2 += M1.x * btt.x - M1.y * btt.y;
    tp3 += T1.x * bmt.x - T1.y * bmt.y;
    tp4 += T1.x * btm.x - T1.y * btm.y;
    tp5 += S1.x * btt.x - S1.y * btt.y + T1.x * bs.x - T1.y * bs.y;
  }
  double *u = planes + ((cell + 1) * 6) * (long long)sv + (long long)j1 * Pq + p;
  const double U0 = u[0], U2 = u[2LL * sv], U3 = u[3LL * sv], U4 = u[4LL * sv], U5 = u[5LL * sv];
  const double t0 = U0 + U5 / 4. + dt * tp0 / scalev / scaleL / scale3;
  const double t2 = U2 + dt * tp2 * 12. / scalev / scaleL / scale3;
  const double t3 = U3 + dt * tp3 * 12. / scalev / scaleL / scale3;
  const double t4 = U4 + dt * tp4 * 12. / scalev / scaleL / scale3;
  const double t5 = U0 / 4. + U5 * 19. / 240. + dt * tp5 / scalev / scaleL / scale3;
  u[0] = 19 * t0 / 4. - 15 * t5;
  u[5LL * sv] = 60 * t5 - 15 * t0;
  u[2LL * sv] = t2; u[3LL * sv] = t3; u[4LL * sv] = t4;      // U[6k+1] is not touched by collisions
}

// ---- register-tiled projection (N % 4 == 0, Nv % 8 == 0).  The simple kernels above issue one or more loads per
// complex multiply-add and are bound by the L1/shared-memory pipe; here every loaded table entry / intermediate
// feeds four (slab) or eight (final) outputs held in registers, so the FP64 pipe is the limit.
__global__ void __launch_bounds__(256, 2) k_project_slab_t(const double2 *__restrict__ q0, const double2 *__restrict__ q1,
                                                        const double2 *__restrict__ q2, const double2 *__restrict__ q3,
                                                        double2 *__restrict__ Bbuf, const double2 *__restrict__ tT,
                                                        const double2 *__restrict__ tM, const double2 *__restrict__ tS,
                                                        int N, int Nv, double nu)
{
  extern __shared__ double2 sm2[];
  const int P = N + 1, PV = Nv + 1;
  double2 *Xs = sm2;                 // [N][P]
  double2 *As = Xs + N * P;          // [3][N][PV]
  const long long slab = blockIdx.x; // cell*N + k1
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int t = tid; t < N * N; t += nt) {
    const long long g = slab * N * N + t;
    const double2 a = q0[g], b = q1[g], c = q2[g], d = q3[g];
    Xs[(t / N) * P + (t % N)] = make_double2(nu * (0.5 * a.x + (b.x + c.x + d.x) * (1. / 6.)), nu * (0.5 * a.y + (b.y + c.y + d.y) * (1. / 6.)));
  }
  __syncthreads();
  // contract k3: As[tab][k2][j3], thread = (j3, four consecutive k2)
  for (int it = tid; it < (N / 4) * Nv; it += nt) {
    const int j3 = it % Nv, k20 = 4 * (it / Nv);
    double2 aT[4], aM[4], aS[4];
    #pragma unroll
    for (int kk = 0; kk < 4; kk++) aT[kk] = aM[kk] = aS[kk] = make_double2(0., 0.);
    #pragma unroll 2
    for (int k3 = 0; k3 < N; k3++) {
      const double2 T3 = tT[k3 * Nv + j3], M3 = tM[k3 * Nv + j3], S3 = tS[k3 * Nv + j3];
      #pragma unroll
      for (int kk = 0; kk < 4; kk++) {
        const double2 x = Xs[(k20 + kk) * P + k3];
        cfma(aT[kk], T3, x); cfma(aM[kk], M3, x); cfma(aS[kk], S3, x);
      }
    }
    #pragma unroll
    for (int kk = 0; kk < 4; kk++) {
      As[(0 * N + k20 + kk) * PV + j3] = aT[kk]; As[(1 * N + k20 + kk) * PV + j3] = aM[kk]; As[(2 * N + k20 + kk) * PV + j3] = aS[kk];
    }
  }
  __syncthreads();
  // contract k2: thread = (j3, four consecutive j2)
  const int Pq = Nv * Nv;
  for (int it = tid; it < (Nv / 4) * Nv; it += nt) {
    const int j3 = it % Nv, j20 = 4 * (it / Nv);
    double2 bTT[4], bMT[4], bTM[4], bS[4];
    #pragma unroll
    for (int jj = 0; jj < 4; jj++) bTT[jj] = bMT[jj] = bTM[jj] = bS[jj] = make_double2(0., 0.);
    #pragma unroll 2
    for (int k2 = 0; k2 < N; k2++) {
      const double2 at = As[(0 * N + k2) * PV + j3], am = As[(1 * N + k2) * PV + j3], as = As[(2 * N + k2) * PV + j3];
      #pragma unroll
      for (int jj = 0; jj < 4; jj++) {
        const double2 T2 = tT[k2 * Nv + j20 + jj], M2 = tM[k2 * Nv + j20 + jj], S2 = tS[k2 * Nv + j20 + jj];
        cfma(bTT[jj], T2, at); cfma(bMT[jj], M2, at); cfma(bTM[jj], T2, am); cfma(bS[jj], S2, at); cfma(bS[jj], T2, as);
      }
    }
    #pragma unroll
    for (int jj = 0; jj < 4; jj++) {
      double2 *o = Bbuf + slab * 4 * Pq + (j20 + jj) * Nv + j3;
      o[0] = bTT[jj]; o[Pq] = bMT[jj]; o[2 * Pq] = bTM[jj]; o[3 * Pq] = bS[jj];
    }
  }
}
// contract k1 and apply the DG update (divisions by scalev, scaleL, scale3, 4, 240 of collisionRoutines_1.cpp:973-983
// folded into correctly rounded factors).  CTA = (cell, 32 consecutive p = (j2, j3)); the whole [k1][4][32] tile of
// the intermediate is fetched with one burst of cp.async (64 KB in flight, read exactly once), then warp w contracts
// k1 for its eight j1 values: lanes read the tile from shared memory, the table entries are warp-uniform loads.
__global__ void __launch_bounds__(128) k_project_final_t(const double2 *__restrict__ Bbuf, double *__restrict__ planes,
                                                         const double2 *__restrict__ tT, const double2 *__restrict__ tM,
                                                         const double2 *__restrict__ tS, int N, int Nv, int sv, double fac)
{
  extern __shared__ double2 Bs[];          // [k1][4][32] | T, M, S [k1][Nv]
  const int Pq = Nv * Nv;
  const long long cell = blockIdx.y; const int p0 = 32 * blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  double2 *sT = Bs + N * 128, *sM = sT + N * Nv, *sS = sM + N * Nv;
  for (int idx = tid; idx < N * 4 * 32; idx += blockDim.x) {
    const int k1 = idx >> 7, w = (idx >> 5) & 3, pp = idx & 31;
    fc3::cp16(Bs + idx, Bbuf + ((cell * N + k1) * 4 + w) * Pq + p0 + pp);
  }
  for (int idx = tid; idx < N * Nv; idx += blockDim.x) { fc3::cp16(sT + idx, tT + idx); fc3::cp16(sM + idx, tM + idx); fc3::cp16(sS + idx, tS + idx); }
  fc3::cp_wait_all();
  __syncthreads();
  const int p = p0 + lane;
  for (int j10 = 8 * warp; j10 < Nv; j10 += 8 * nw) {
    double tp0[8], tp2[8], tp3[8], tp4[8], tp5[8];
    #pragma unroll
    for (int a = 0; a < 8; a++) tp0[a] = tp2[a] = tp3[a] = tp4[a] = tp5[a] = 0.;
    #pragma unroll 2
    for (int k1 = 0; k1 < N; k1++) {
      const double2 *b = Bs + k1 * 128 + lane;
      const double2 btt = b[0], bmt = b[32], btm = b[64], bs = b[96];
      #pragma unroll
      for (int a = 0; a < 8; a++) {
        const double2 T1 = sT[k1 * Nv + j10 + a], M1 = sM[k1 * Nv + j10 + a], S1 = sS[k1 * Nv + j10 + a];
        tp0[a] = fma(T1.x, btt.x, tp0[a]); tp0[a] = fma(-T1.y, btt.y, tp0[a]);
        tp2[a] = fma(M1.x, btt.x, tp2[a]); tp2[a] = fma(-M1.y, btt.y, tp2[a]);
        tp3[a] = fma(T1.x, bmt.x, tp3[a]); tp3[a] = fma(-T1.y, bmt.y, tp3[a]);
        tp4[a] = fma(T1.x, btm.x, tp4[a]); tp4[a] = fma(-T1.y, btm.y, tp4[a]);
        tp5[a] = fma(S1.x, btt.x, tp5[a]); tp5[a] = fma(-S1.y, btt.y, tp5[a]);
        tp5[a] = fma(T1.x, bs.x, tp5[a]); tp5[a] = fma(-T1.y, bs.y, tp5[a]);
      }
    }
    #pragma unroll
    for (int a = 0; a < 8; a++) {
      double *u = planes + ((cell + 1) * 6) * (long long)sv + (long long)(j10 + a) * Pq + p;
      const double U0 = u[0], U2 = u[2LL * sv], U3 = u[3LL * sv], U4 = u[4LL * sv], U5 = u[5LL * sv];
      const double t0 = U0 + U5 * 0.25 + tp0[a] * fac;
      const double t2 = U2 + tp2[a] * (12. * fac);
      const double t3 = U3 + tp3[a] * (12. * fac);
      const double t4 = U4 + tp4[a] * (12. * fac);
      const double t5 = U0 * 0.25 + U5 * (19. / 240.) + tp5[a] * fac;
      u[0] = 19 * t0 * 0.25 - 15 * t5;
      u[5LL * sv] = 60 * t5 - 15 * t0;
      u[2LL * sv] = t2; u[3LL * sv] = t3; u[4LL * sv] = t4;      // U[6k+1] is not touched by collisions
    }
  }
}

// ---- Hermitian-folded projection.  The projection keeps only the real part, and the 1-D factors satisfy
// T(-k) = conj T(k) (same for M, S), so  Re IntM(-k) X[-k] = Re IntM(k) conj(X[-k]):  the k1 < 0 half of the spectrum is
// folded onto the k1 > 0 half *before* the contractions,
//   Y[k1][k2][k3] = X[k1][k2][k3] + conj X[-k1][-k2][-k3]     (k1 > 0),
// and N/2 + 1 slabs are contracted instead of N (k1 = -N/2 and k1 = 0 have no partner slab and stay as they are).  Exact
// for any X -- nothing is assumed about X being Hermitian.  The wave numbers run over [-N/2, N/2 - 1], so the partner of
// an entry with k2 = -N/2 or k3 = -N/2 has the wave number +N/2 that the arrays do not hold: k2 and k3 are therefore
// contracted over N + 1 entries, entry N standing for +N/2 with the table row conj(row 0) (tTx, tMx, tSx: [N+1][Nv]).
// slab s of a cell: s = 0 -> index 0 (k1 = -N/2), s = 1 -> index N/2 (k1 = 0), s >= 2 -> index N/2 + s - 1 folded with N - index.
__device__ __forceinline__ int fold_index(int s, int N) { return s == 0 ? 0 : N / 2 + s - 1; }

__global__ void __launch_bounds__(256, 2) k_project_slab_h(const double2 *__restrict__ q0, const double2 *__restrict__ q1,
                                                           const double2 *__restrict__ q2, const double2 *__restrict__ q3,
                                                           double2 *__restrict__ Bbuf, const double2 *__restrict__ tTx,
                                                           const double2 *__restrict__ tMx, const double2 *__restrict__ tSx,
                                                           int N, int Nv, double nu)
{
  extern __shared__ double2 sm2[];
  const int K = N + 1, P = K | 1, PV = Nv + 1, NH = N / 2 + 1;
  double2 *Xs = sm2;                 // [K][P]
  double2 *As = Xs + K * P;          // [3][K][PV]
  double2 *Ps = As + 3 * K * PV;     // [N/4][3][Nv]: partial sums of row N of As
  const long long cell = blockIdx.x / NH;
  const int s = blockIdx.x % NH, i1 = fold_index(s, N);
  const bool paired = s >= 2;
  const int tid = threadIdx.x, nt = blockDim.x;
  const long long base = (cell * N + i1) * N * N, pbase = (cell * N + (N - i1)) * N * N;
  for (int t = tid; t < K * K; t += nt) {
    const int i2 = t / K, i3 = t % K;
    double2 y = make_double2(0., 0.);
    if (i2 < N && i3 < N) {
      const long long g = base + i2 * N + i3;
      const double2 a = q0[g], b = q1[g], c = q2[g], d = q3[g];
      y = make_double2(nu * (0.5 * a.x + (b.x + c.x + d.x) * (1. / 6.)), nu * (0.5 * a.y + (b.y + c.y + d.y) * (1. / 6.)));
    }
    if (paired && i2 > 0 && i3 > 0) {
      const long long g = pbase + (N - i2) * N + (N - i3);
      const double2 a = q0[g], b = q1[g], c = q2[g], d = q3[g];
      y.x += nu * (0.5 * a.x + (b.x + c.x + d.x) * (1. / 6.));
      y.y -= nu * (0.5 * a.y + (b.y + c.y + d.y) * (1. / 6.));
    }
    Xs[i2 * P + i3] = y;
  }
  __syncthreads();
  // contract k3: As[tab][k2][j3], thread = (j3, four consecutive k2 < N); row k2 = N is shared out: the thread of group g
  // adds up k3 = 4g .. 4g+3 (group 0 also k3 = N) and the partial sums are folded after the barrier
  for (int it = tid; it < (N / 4) * Nv; it += nt) {
    const int j3 = it % Nv, g = it / Nv, k20 = 4 * g;
    double2 aT[4], aM[4], aS[4];
    #pragma unroll
    for (int kk = 0; kk < 4; kk++) aT[kk] = aM[kk] = aS[kk] = make_double2(0., 0.);
    #pragma unroll 2
    for (int k3 = 0; k3 < K; k3++) {
      const double2 T3 = tTx[k3 * Nv + j3], M3 = tMx[k3 * Nv + j3], S3 = tSx[k3 * Nv + j3];
      #pragma unroll
      for (int kk = 0; kk < 4; kk++) {
        const double2 x = Xs[(k20 + kk) * P + k3];
        cfma(aT[kk], T3, x); cfma(aM[kk], M3, x); cfma(aS[kk], S3, x);
      }
    }
    #pragma unroll
    for (int kk = 0; kk < 4; kk++) {
      As[(0 * K + k20 + kk) * PV + j3] = aT[kk]; As[(1 * K + k20 + kk) * PV + j3] = aM[kk]; As[(2 * K + k20 + kk) * PV + j3] = aS[kk];
    }
    double2 pT = make_double2(0., 0.), pM = pT, pS = pT;
    for (int kk = 0; kk < (g == 0 ? 5 : 4); kk++) {
      const int k3 = kk < 4 ? k20 + kk : N;
      const double2 x = Xs[N * P + k3];
      cfma(pT, tTx[k3 * Nv + j3], x); cfma(pM, tMx[k3 * Nv + j3], x); cfma(pS, tSx[k3 * Nv + j3], x);
    }
    Ps[(g * 3 + 0) * Nv + j3] = pT; Ps[(g * 3 + 1) * Nv + j3] = pM; Ps[(g * 3 + 2) * Nv + j3] = pS;
  }
  __syncthreads();
  for (int t = tid; t < 3 * Nv; t += nt) {
    const int w = t / Nv, j3 = t % Nv;
    double2 a = make_double2(0., 0.);
    for (int g = 0; g < N / 4; g++) { const double2 x = Ps[(g * 3 + w) * Nv + j3]; a.x += x.x; a.y += x.y; }
    As[(w * K + N) * PV + j3] = a;
  }
  __syncthreads();
  // contract k2: thread = (j3, four consecutive j2)
  const int Pq = Nv * Nv;
  for (int it = tid; it < (Nv / 4) * Nv; it += nt) {
    const int j3 = it % Nv, j20 = 4 * (it / Nv);
    double2 bTT[4], bMT[4], bTM[4], bS[4];
    #pragma unroll
    for (int jj = 0; jj < 4; jj++) bTT[jj] = bMT[jj] = bTM[jj] = bS[jj] = make_double2(0., 0.);
    #pragma unroll 2
    for (int k2 = 0; k2 < K; k2++) {
      const double2 at = As[(0 * K + k2) * PV + j3], am = As[(1 * K + k2) * PV + j3], as = As[(2 * K + k2) * PV + j3];
      #pragma unroll
      for (int jj = 0; jj < 4; jj++) {
        const double2 T2 = tTx[k2 * Nv + j20 + jj], M2 = tMx[k2 * Nv + j20 + jj], S2 = tSx[k2 * Nv + j20 + jj];
        cfma(bTT[jj], T2, at); cfma(bMT[jj], M2, at); cfma(bTM[jj], T2, am); cfma(bS[jj], S2, at); cfma(bS[jj], T2, as);
      }
    }
    #pragma unroll
    for (int jj = 0; jj < 4; jj++) {
      double2 *o = Bbuf + (long long)blockIdx.x * 4 * Pq + (j20 + jj) * Nv + j3;
      o[0] = bTT[jj]; o[Pq] = bMT[jj]; o[2 * Pq] = bTM[jj]; o[3 * Pq] = bS[jj];
    }
  }
}
// k1 contraction over the N/2 + 1 folded slabs + the DG update; otherwise k_project_final_t.
__global__ void __launch_bounds__(128) k_project_final_h(const double2 *__restrict__ Bbuf, double *__restrict__ planes,
                                                         const double2 *__restrict__ tT, const double2 *__restrict__ tM,
                                                         const double2 *__restrict__ tS, int N, int Nv, int sv, double fac)
{
  extern __shared__ double2 Bs[];          // [s][4][32] | T, M, S [s][Nv]
  const int Pq = Nv * Nv, NH = N / 2 + 1;
  const long long cell = blockIdx.y; const int p0 = 32 * blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  double2 *sT = Bs + NH * 128, *sM = sT + NH * Nv, *sS = sM + NH * Nv;
  for (int idx = tid; idx < NH * 4 * 32; idx += blockDim.x) {
    const int s = idx >> 7, w = (idx >> 5) & 3, pp = idx & 31;
    fc3::cp16(Bs + idx, Bbuf + ((cell * NH + s) * 4 + w) * Pq + p0 + pp);
  }
  for (int idx = tid; idx < NH * Nv; idx += blockDim.x) {
    const int g = fold_index(idx / Nv, N) * Nv + idx % Nv;
    fc3::cp16(sT + idx, tT + g); fc3::cp16(sM + idx, tM + g); fc3::cp16(sS + idx, tS + g);
  }
  fc3::cp_wait_all();
  __syncthreads();
  const int p = p0 + lane;
  for (int j10 = 8 * warp; j10 < Nv; j10 += 8 * nw) {
    double tp0[8], tp2[8], tp3[8], tp4[8], tp5[8];
    #pragma unroll
    for (int a = 0; a < 8; a++) tp0[a] = tp2[a] = tp3[a] = tp4[a] = tp5[a] = 0.;
    #pragma unroll 2
    for (int k1 = 0; k1 < NH; k1++) {
      const double2 *b = Bs + k1 * 128 + lane;
      const double2 btt = b[0], bmt = b[32], btm = b[64], bs = b[96];
      #pragma unroll
      for (int a = 0; a < 8; a++) {
        const double2 T1 = sT[k1 * Nv + j10 + a], M1 = sM[k1 * Nv + j10 + a], S1 = sS[k1 * Nv + j10 + a];
        tp0[a] = fma(T1.x, btt.x, tp0[a]); tp0[a] = fma(-T1.y, btt.y, tp0[a]);
        tp2[a] = fma(M1.x, btt.x, tp2[a]); tp2[a] = fma(-M1.y, btt.y, tp2[a]);
        tp3[a] = fma(T1.x, bmt.x, tp3[a]); tp3[a] = fma(-T1.y, bmt.y, tp3[a]);
        tp4[a] = fma(T1.x, btm.x, tp4[a]); tp4[a] = fma(-T1.y, btm.y, tp4[a]);
        tp5[a] = fma(S1.x, btt.x, tp5[a]); tp5[a] = fma(-S1.y, btt.y, tp5[a]);
        tp5[a] = fma(T1.x, bs.x, tp5[a]); tp5[a] = fma(-T1.y, bs.y, tp5[a]);
      }
    }
    #pragma unroll
    for (int a = 0; a < 8; a++) {
      double *u = planes + ((cell + 1) * 6) * (long long)sv + (long long)(j10 + a) * Pq + p;
      const double U0 = u[0], U2 = u[2LL * sv], U3 = u[3LL * sv], U4 = u[4LL * sv], U5 = u[5LL * sv];
      const double t0 = U0 + U5 * 0.25 + tp0[a] * fac;
      const double t2 = U2 + tp2[a] * (12. * fac);
      const double t3 = U3 + tp3[a] * (12. * fac);
      const double t4 = U4 + tp4[a] * (12. * fac);
      const double t5 = U0 * 0.25 + U5 * (19. / 240.) + tp5[a] * fac;
      u[0] = 19 * t0 * 0.25 - 15 * t5;
      u[5LL * sv] = 60 * t5 - 15 * t0;
      u[2LL * sv] = t2; u[3LL * sv] = t3; u[4LL * sv] = t4;      // U[6k+1] is not touched by collisions
    }
  }
}

int lp_launch_project(lpgpu_ctx *c, double *planes, int B)
{
  const int N = c->p.N, Nv = c->p.Nv;
  const size_t smem = ((size_t)N * (N + 1) + (size_t)3 * N * (Nv + 1)) * sizeof(double2);
  int threads = N * Nv > Nv * Nv ? N * Nv : Nv * Nv;
  if (threads > 1024) threads = 1024;
  LP_CUDA(cudaFuncSetAttribute(k_project_slab, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const double2 *T = reinterpret_cast<const double2 *>(c->d_T), *M = reinterpret_cast<const double2 *>(c->d_M),
                *S = reinterpret_cast<const double2 *>(c->d_S);
  static const bool simple_only = getenv("LPGPU_PROJECT_SIMPLE") != nullptr;   // developer knob
  static const bool unfolded = getenv("LPGPU_PROJECT_UNFOLDED") != nullptr;    // developer knob: contract all N slabs
  if (!simple_only && !unfolded && c->project_fold && N % 4 == 0 && Nv % 8 == 0) {
    const int K = N + 1, NH = N / 2 + 1;
    const size_t smemh = ((size_t)K * (K | 1) + (size_t)3 * K * (Nv + 1) + (size_t)(N / 4) * 3 * Nv) * sizeof(double2);
    LP_CUDA(cudaFuncSetAttribute(k_project_slab_h, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemh));
    k_project_slab_h<<<B * NH, 256, smemh, c->stream>>>(
        reinterpret_cast<const double2 *>(c->d_q[0]), reinterpret_cast<const double2 *>(c->d_q[1]),
        reinterpret_cast<const double2 *>(c->d_q[2]), reinterpret_cast<const double2 *>(c->d_q[3]),
        reinterpret_cast<double2 *>(c->d_B), reinterpret_cast<const double2 *>(c->d_Tx), reinterpret_cast<const double2 *>(c->d_Mx),
        reinterpret_cast<const double2 *>(c->d_Sx), N, Nv, c->p.nu);
    LP_LAUNCHED(c);
    const double fac = c->p.dt / c->tab.scalev / c->tab.scaleL / c->tab.scale3;
    const size_t smemf = ((size_t)NH * 4 * 32 + (size_t)3 * NH * Nv) * sizeof(double2);
    LP_CUDA(cudaFuncSetAttribute(k_project_final_h, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemf));
    k_project_final_h<<<dim3(Nv * Nv / 32, B), 128, smemf, c->stream>>>(reinterpret_cast<const double2 *>(c->d_B), planes, T, M, S, N, Nv, c->sv, fac);
    LP_LAUNCHED(c);
    return LPGPU_OK;
  }
  if (!simple_only && N % 4 == 0 && Nv % 8 == 0) {
    LP_CUDA(cudaFuncSetAttribute(k_project_slab_t, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_project_slab_t<<<B * N, 256, smem, c->stream>>>(
        reinterpret_cast<const double2 *>(c->d_q[0]), reinterpret_cast<const double2 *>(c->d_q[1]),
        reinterpret_cast<const double2 *>(c->d_q[2]), reinterpret_cast<const double2 *>(c->d_q[3]),
        reinterpret_cast<double2 *>(c->d_B), T, M, S, N, Nv, c->p.nu);
    LP_LAUNCHED(c);
    const double fac = c->p.dt / c->tab.scalev / c->tab.scaleL / c->tab.scale3;
    const size_t smemf = ((size_t)N * 4 * 32 + (size_t)3 * N * Nv) * sizeof(double2);
    LP_CUDA(cudaFuncSetAttribute(k_project_final_t, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemf));
    k_project_final_t<<<dim3(Nv * Nv / 32, B), 128, smemf, c->stream>>>(reinterpret_cast<const double2 *>(c->d_B), planes, T, M, S, N, Nv, c->sv, fac);
    LP_LAUNCHED(c);
    return LPGPU_OK;
  }
  k_project_slab<<<B * N, threads, smem, c->stream>>>(
      reinterpret_cast<const double2 *>(c->d_q[0]), reinterpret_cast<const double2 *>(c->d_q[1]),
      reinterpret_cast<const double2 *>(c->d_q[2]), reinterpret_cast<const double2 *>(c->d_q[3]),
      reinterpret_cast<double2 *>(c->d_B), T, M, S, N, Nv, c->p.nu);
  LP_LAUNCHED(c);
  dim3 grid((Nv * Nv + 255) / 256, B, Nv);
  k_project_final<<<grid, 256, 0, c->stream>>>(reinterpret_cast<const double2 *>(c->d_B), planes, T, M, S, N, Nv, c->sv,
                                               c->p.dt, c->tab.scalev, c->tab.scaleL, c->tab.scale3);
  LP_LAUNCHED(c);
  return LPGPU_OK;
}
