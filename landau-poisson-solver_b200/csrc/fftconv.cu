// ComputeQ as seven zero-padded linear convolutions (computeq_variant 0/2) -- the O(N^3 log N) form of the
// same sum (identity checked against the reference's table to 1.5e-16); the algebra is in fc3.cuh.
//
// Two implementations live here:
//  * the register-resident pipeline of fc3.cuh (N = 8, 16, 24, 32; the default path): kernels k_fc3_f1 / k_fc3_f2 /
//    k_fc3_f2_tmem / k_fc3_f3, fused with fft3D's last pass, the conservation dot products, and -- for N = 32 -- with
//    the product accumulators and the waiting u transform parked in tensor memory;
//  * a generic shared-memory version for any other even N whose M = 3N/2 is 2^a 3^b (k_fc_fwd_yz / k_fc_x /
//    k_fc_inv_yz): decimation-in-frequency stages on shared-memory lines (natural in, digit-reversed out; the
//    inverse runs the conjugate-transposed stages in reverse order, so products are formed position-wise and no
//    permutation pass exists), 14 arrays of N M^2 through HBM between the y-z and the x kernel.
// Sizes whose M has another prime factor use the tiled direct kernel (computeq.cu).
#include "lpgpu_internal.h"
#include "fc3.cuh"

#define LP_LAUNCHED(c)                                  \
  do {                                                  \
    (c)->launches++;                                    \
    LP_CUDA(cudaGetLastError());                        \
  } while (0)

namespace {

struct FcPlan { int M, nst, radix[8]; };

__device__ __forceinline__ double2 cxmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cxmulc(double2 a, double2 b) { return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }   // a * conj(b)
__device__ __forceinline__ double2 cxadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 cxsub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
#define LP_SQRT3_2 0.86602540378443864676

// One forward (DIF) butterfly of radix r on the block of length L = r*Ls that starts at p: elements p[q*Ls*st];
// outputs are multiplied by w_L^(j k) = tw[j*k*step], tw[t] = exp(-2 pi i t / M), step = M / L.
__device__ __forceinline__ void fwd_bfly(double2 *p, int r, int stride, int j, int step, const double2 *tw)
{
  if (r == 2) {
    const double2 a = p[0], b = p[stride];
    p[0] = cxadd(a, b);
    p[stride] = cxmul(cxsub(a, b), tw[j * step]);
  } else {
    const double2 a0 = p[0], a1 = p[stride], a2 = p[2 * stride];
    const double2 t1 = cxadd(a1, a2), t2 = make_double2(a0.x - 0.5 * t1.x, a0.y - 0.5 * t1.y);
    const double2 d = cxsub(a1, a2), t3 = make_double2(LP_SQRT3_2 * d.x, LP_SQRT3_2 * d.y);
    p[0] = cxadd(a0, t1);
    p[stride] = cxmul(make_double2(t2.x + t3.y, t2.y - t3.x), tw[j * step]);        // t2 - i t3
    p[2 * stride] = cxmul(make_double2(t2.x - t3.y, t2.y + t3.x), tw[2 * j * step]); // t2 + i t3
  }
}
// The conjugate-transposed butterfly (inverse): conj twiddles first, then the inverse r-point DFT.
__device__ __forceinline__ void inv_bfly(double2 *p, int r, int stride, int j, int step, const double2 *tw)
{
  if (r == 2) {
    const double2 y0 = p[0], y1 = cxmulc(p[stride], tw[j * step]);
    p[0] = cxadd(y0, y1);
    p[stride] = cxsub(y0, y1);
  } else {
    const double2 y0 = p[0], y1 = cxmulc(p[stride], tw[j * step]), y2 = cxmulc(p[2 * stride], tw[2 * j * step]);
    const double2 t1 = cxadd(y1, y2), t2 = make_double2(y0.x - 0.5 * t1.x, y0.y - 0.5 * t1.y);
    const double2 d = cxsub(y1, y2), t3 = make_double2(LP_SQRT3_2 * d.x, LP_SQRT3_2 * d.y);
    p[0] = cxadd(y0, t1);
    p[stride] = make_double2(t2.x - t3.y, t2.y + t3.x);        // t2 + i t3
    p[2 * stride] = make_double2(t2.x + t3.y, t2.y - t3.x);    // t2 - i t3
  }
}

// A whole line (elements x[i*st]) by one warp: lanes share the M/r butterflies of each stage.
template <bool FWD>
__device__ __forceinline__ void line_fft_warp(double2 *x, int st, const FcPlan &pl, const double2 *tw, int lane)
{
  const int M = pl.M;
  if (FWD) {
    int L = M;
    for (int s = 0; s < pl.nst; s++) {
      const int r = pl.radix[s], Ls = L / r, step = M / L;
      for (int it = lane; it < M / r; it += 32) {
        const int blk = it / Ls, j = it - blk * Ls;
        fwd_bfly(x + (blk * L + j) * st, r, Ls * st, j, step, tw);
      }
      __syncwarp();
      L = Ls;
    }
  } else {
    int Ls = 1;
    for (int s = pl.nst - 1; s >= 0; s--) {
      const int r = pl.radix[s], L = Ls * r, step = M / L;
      for (int it = lane; it < M / r; it += 32) {
        const int blk = it / Ls, j = it - blk * Ls;
        inv_bfly(x + (blk * L + j) * st, r, Ls * st, j, step, tw);
      }
      __syncwarp();
      Ls = L;
    }
  }
}
// All M columns of a [M][P] plane by the whole block: lanes = columns, warps share the butterflies of a stage.
template <bool FWD>
__device__ __forceinline__ void columns_fft_block(double2 *plane, int P, const FcPlan &pl, const double2 *tw)
{
  const int M = pl.M, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (FWD) {
    int L = M;
    for (int s = 0; s < pl.nst; s++) {
      const int r = pl.radix[s], Ls = L / r, step = M / L;
      for (int it = warp; it < M / r; it += nw) {
        const int blk = it / Ls, j = it - blk * Ls;
        for (int col = lane; col < M; col += 32) fwd_bfly(plane + (blk * L + j) * P + col, r, Ls * P, j, step, tw);
      }
      __syncthreads();
      L = Ls;
    }
  } else {
    int Ls = 1;
    for (int s = pl.nst - 1; s >= 0; s--) {
      const int r = pl.radix[s], L = Ls * r, step = M / L;
      for (int it = warp; it < M / r; it += nw) {
        const int blk = it / Ls, j = it - blk * Ls;
        for (int col = lane; col < M; col += 32) inv_bfly(plane + (blk * L + j) * P + col, r, Ls * P, j, step, tw);
      }
      __syncthreads();
      Ls = L;
    }
  }
}

// F1: padded y-z plane of u_p (p < 7) or v_{p-7}, transformed along z and y
__global__ void __launch_bounds__(256) k_fc_fwd_yz(const double2 *__restrict__ fhat, double2 *__restrict__ Fxy, const double *__restrict__ G,
                                                   const double *__restrict__ Etab, const double2 *__restrict__ twg, int N, FcPlan pl)
{
  extern __shared__ double2 smf[];
  const int M = pl.M, P = M + 1;
  double2 *plane = smf, *tw = plane + M * P;
  const int x = blockIdx.x, p = blockIdx.y; const long long cell = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  for (int t = tid; t < M; t += blockDim.x) tw[t] = twg[t];
  for (int t = tid; t < M * P; t += blockDim.x) plane[t] = make_double2(0., 0.);
  __syncthreads();
  const double *E = Etab + LP_ETAB_PAD;
  const double ex = E[x];
  for (int t = tid; t < N * N; t += blockDim.x) {
    const int y = t / N, z = t % N;
    const long long w = ((long long)x * N + y) * N + z;
    const double2 f = fhat[cell * N * N * N + w];
    double m;
    if (p < 7) m = G[7 * w + p];
    else {
      const double ey = E[y], ez = E[z];
      switch (p - 7) {
        case 0: m = 1.; break;
        case 1: m = -ex * ex; break;
        case 2: m = -ey * ey; break;
        case 3: m = -ez * ez; break;
        case 4: m = -ex * ey; break;
        case 5: m = -ex * ez; break;
        default: m = -ey * ez; break;
      }
    }
    plane[y * P + z] = make_double2(m * f.x, m * f.y);
  }
  __syncthreads();
  for (int y = warp; y < N; y += nw) line_fft_warp<true>(plane + y * P, 1, pl, tw, lane);
  __syncthreads();
  columns_fft_block<true>(plane, P, pl, tw);
  double2 *o = Fxy + ((cell * 14 + p) * N + x) * (long long)(M * M);
  for (int t = tid; t < M * M; t += blockDim.x) o[t] = plane[(t / M) * P + (t % M)];
}

// F2: x-lines of the 14 arrays at 8 consecutive kz: transform, multiply-accumulate over p, inverse transform
__global__ void __launch_bounds__(256) k_fc_x(const double2 *__restrict__ Fxy, double2 *__restrict__ Cx, const double2 *__restrict__ twg, int N, FcPlan pl)
{
  extern __shared__ double2 smf[];
  const int M = pl.M, P = M + 1, H = N / 2;
  double2 *U = smf, *V = U + 8 * P, *tw = V + 8 * P;
  const int kz0 = blockIdx.x * 8, ky = blockIdx.y; const long long cell = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;   // 8 warps: warp = line
  const int nkz = (M - kz0 < 8) ? M - kz0 : 8;                       // M need not be a multiple of 8
  for (int t = tid; t < M; t += blockDim.x) tw[t] = twg[t];
  double2 acc[2] = {make_double2(0., 0.), make_double2(0., 0.)};
  const long long MM = (long long)M * M;
  for (int p = 0; p < 7; p++) {
    __syncthreads();
    const double2 *su = Fxy + ((cell * 14 + p) * N) * MM + (long long)ky * M + kz0;
    const double2 *sv = Fxy + ((cell * 14 + p + 7) * N) * MM + (long long)ky * M + kz0;
    for (int t = tid; t < 8 * M; t += blockDim.x) {
      const int x = t >> 3, kzi = t & 7;
      const bool live = x < N && kzi < nkz;
      U[kzi * P + x] = live ? su[x * MM + kzi] : make_double2(0., 0.);
      V[kzi * P + x] = live ? sv[x * MM + kzi] : make_double2(0., 0.);
    }
    __syncthreads();
    line_fft_warp<true>(U + warp * P, 1, pl, tw, lane);
    line_fft_warp<true>(V + warp * P, 1, pl, tw, lane);
    #pragma unroll
    for (int q = 0; q < 2; q++) {
      const int idx = lane + 32 * q;
      if (idx < M) { const double2 pr = cxmul(U[warp * P + idx], V[warp * P + idx]); acc[q].x += pr.x; acc[q].y += pr.y; }
    }
  }
  __syncwarp();
  #pragma unroll
  for (int q = 0; q < 2; q++) { const int idx = lane + 32 * q; if (idx < M) U[warp * P + idx] = acc[q]; }
  __syncwarp();
  line_fft_warp<false>(U + warp * P, 1, pl, tw, lane);
  __syncthreads();
  double2 *o = Cx + (cell * N) * MM + (long long)ky * M + kz0;
  for (int t = tid; t < 8 * N; t += blockDim.x) {
    const int xo = t >> 3, kzi = t & 7;
    if (kzi < nkz) o[xo * MM + kzi] = U[kzi * P + xo + H];
  }
}

// F3: inverse along y and z of one x' plane, scale, extract the [N/2, N/2+N) window
__global__ void __launch_bounds__(256) k_fc_inv_yz(const double2 *__restrict__ Cx, double2 *__restrict__ q, const double2 *__restrict__ twg, int N, FcPlan pl)
{
  extern __shared__ double2 smf[];
  const int M = pl.M, P = M + 1, H = N / 2;
  double2 *plane = smf, *tw = plane + M * P;
  const int xo = blockIdx.x; const long long cell = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  for (int t = tid; t < M; t += blockDim.x) tw[t] = twg[t];
  const double2 *s = Cx + (cell * N + xo) * (long long)(M * M);
  for (int t = tid; t < M * M; t += blockDim.x) plane[(t / M) * P + (t % M)] = s[t];
  __syncthreads();
  columns_fft_block<false>(plane, P, pl, tw);
  for (int y = H + warp; y < H + N; y += nw) line_fft_warp<false>(plane + y * P, 1, pl, tw, lane);
  __syncthreads();
  const double sc = 1.0 / ((double)M * M * M);
  double2 *o = q + (cell * N + xo) * (long long)(N * N);
  for (int t = tid; t < N * N; t += blockDim.x) {
    const double2 v = plane[(t / N + H) * P + (t % N) + H];
    o[t] = make_double2(v.x * sc, v.y * sc);
  }
}

// ---- mbarrier / bulk-copy (TMA) helpers.  The precomputed kernel symbols of F1 (seven N x N slabs per CTA) and the
// planes of k_fc3_f2_tmem (two 16 KB planes per product) arrive as cp.async.bulk copies issued by ONE thread and completed
// on an mbarrier in bytes, instead of thousands of 16-byte cp.async spread over the CTA's threads.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *b, unsigned count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(b)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *b, unsigned bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long *b, unsigned parity)
{
  asm volatile("{\n"
               ".reg .pred P1;\n"
               "LAB_WAIT:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
               "@P1 bra DONE;\n"
               "bra LAB_WAIT;\n"
               "DONE:\n"
               "}\n" :: "r"(smem_u32(b)), "r"(parity) : "memory");
}
// one contiguous block global -> shared through the TMA unit; completion is counted in bytes on `bar`
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src, unsigned bytes, unsigned long long *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               :: "r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// Register-resident pipeline (fc3.cuh): F1 z-lines -> F2 (y, x, product, inverse x, inverse y) -> F3 inverse z.
// FUSED: `in` is the output of the first fft3D pass (lp_launch_fft3d_jk); the N lines along i of this y slab are
// transformed here (one thread per line, as k_tf_i) and post-phased straight into the shared fhat slab -- fhat
// itself never goes to memory.  The kernel symbols of this y are staged with cp.async while that happens.
// mhat (nullable): LinearLandau -- the seven u arrays are built from the stored transform of the Maxwellian instead of
// fhat (its y slab sits in a second shared array behind the symbols; the launcher sizes the dynamic shared memory)
template <int L, bool FUSED>
__global__ void __launch_bounds__(fc3::F1<L>::NT, 2) k_fc3_f1(const double2 *__restrict__ in, const double *__restrict__ Gt,
                                                           const double *__restrict__ E, double2 *__restrict__ Z, const double2 *__restrict__ post,
                                                           const double2 *__restrict__ mhat, long long mstride)
{
  typedef fc3::F1<L> K;
  constexpr int N = K::N;
  extern __shared__ __align__(128) double2 smf[];
  double2 *FS = smf;
  double *Gs = reinterpret_cast<double *>(FS + K::SMEM_C2), *sE = Gs + 7 * N * N;
  double2 *FM = mhat ? reinterpret_cast<double2 *>(sE + N) : nullptr;
  const int y = blockIdx.x, cell = blockIdx.y, tid = threadIdx.x;
  __shared__ __align__(8) unsigned long long s_bar;
  if (tid == 0) {                            // the seven symbol slabs Gt[a][y][.][.] of this y: one bulk copy (TMA) each
    constexpr unsigned SLAB_BYTES = N * N * sizeof(double);
    mbar_init(&s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    mbar_expect_tx(&s_bar, 7 * SLAB_BYTES);
    for (int a = 0; a < 7; a++) bulk_load(Gs + a * N * N, Gt + ((long long)a * N + y) * N * N, SLAB_BYTES, &s_bar);
  }
  if (mhat) K::load_slab(tid, cell, y, mhat, FM, mstride);
  if (FUSED) {
    // the N lines along i of this slab, four threads per line: thread (j, k) forms the decimated sequence
    // y_j[n] = (sum_s x[n + s N/4] (-i)^(j s)) w_N^(j n) and transforms it (N/4 points): X[4q + j]
    if (tid < 4 * N) {
      constexpr int Q = N / 4;
      const int j = tid / N, k = tid % N;
      const double2 *src = in + (((long long)cell * N) * N + y) * N + k;
      double2 v[Q], ph[Q];
      #pragma unroll
      for (int q = 0; q < Q; q++) ph[q] = __ldg(post + ((4 * q + j) * N + y) * N + k);     // issued with the data loads, not after the transform
      // every load of the line in flight before the first use: ONE memory round trip per CTA instead of Q dependent ones
      // (ncu r02m: the adds below held 25 % of the kernel's warp samples on the long scoreboard, eight loads at a time)
      double2 xin[4][Q];
      #pragma unroll
      for (int n = 0; n < Q; n++) {
        #pragma unroll
        for (int s4 = 0; s4 < 4; s4++) xin[s4][n] = __ldg(src + (long long)(n + s4 * Q) * N * N);
      }
      #pragma unroll
      for (int n = 0; n < Q; n++) {
        const double2 x0 = xin[0][n], x1 = xin[1][n], x2 = xin[2][n], x3 = xin[3][n];
        double2 t;
        if (j == 0) t = fc3::cadd(fc3::cadd(x0, x2), fc3::cadd(x1, x3));
        else if (j == 2) t = fc3::csub(fc3::cadd(x0, x2), fc3::cadd(x1, x3));
        else {
          const double2 a = fc3::csub(x0, x2), b = fc3::csub(x1, x3);       // j = 1: a - i b,  j = 3: a + i b
          t = (j == 1) ? make_double2(a.x + b.y, a.y - b.x) : make_double2(a.x - b.y, a.y + b.x);
        }
        v[n] = t;
      }
      if (j == 1) {
        #pragma unroll
        for (int n = 0; n < Q; n++) v[n] = fc3::mul_tw<N, -1>(v[n], n);
      } else if (j == 2) {
        #pragma unroll
        for (int n = 0; n < Q; n++) v[n] = fc3::mul_tw<N, -1>(v[n], 2 * n);
      } else if (j == 3) {
        #pragma unroll
        for (int n = 0; n < Q; n++) v[n] = fc3::mul_tw<N, -1>(v[n], 3 * n);
      }
      fc3::fftN<Q, -1, N>(v);
      #pragma unroll
      for (int q = 0; q < Q; q++) {
        const int i = 4 * q + j;
        FS[i * K::P + k] = phase_mul(ph[q], v[q]);
      }
    }
    for (int t = tid; t < N; t += K::NT) sE[t] = E[t];
  } else {
    K::load(tid, cell, y, in, E, FS, sE);
  }
  __syncthreads();                           // the fhat slab is complete (and s_bar initialised)
  mbar_wait(&s_bar, 0u);                     // the symbols have landed
  if (gridDim.z == 1) K::lines(tid, cell, y, Gs, N * N, FS, sE, Z, 0, 5, FM);
  else K::lines(tid, cell, y, Gs, N * N, FS, sE, Z, blockIdx.z, blockIdx.z + 1, FM);     // few cells: one round per CTA
}
template <int L>
__global__ void __launch_bounds__(fc3::F2<L>::NT, 1) k_fc3_f2(const double2 *__restrict__ Z, const double *__restrict__ E, double2 *__restrict__ C,
                                                              long long split_stride)
{
  typedef fc3::F2<L> K;
  extern __shared__ double2 smf[];
  double2 *IN = smf, *Y = IN + K::IN_C2;
  double *sE = reinterpret_cast<double *>(Y + K::Y_C2);
  const int kz = blockIdx.x, cell = blockIdx.y, tid = threadIdx.x;
  int p_begin, p_end;
  K::psplit(blockIdx.z, gridDim.z, p_begin, p_end);       // gridDim.z = 3: the seven products split over three CTAs
  C += blockIdx.z * split_stride;
  K::issue_loads(tid, cell, kz, p_begin, Z, IN);
  for (int i = tid; i < K::N; i += K::NT) sE[i] = E[i];
  double2 acc[L];
  #pragma unroll
  for (int q = 0; q < L; q++) acc[q] = make_double2(0., 0.);
  #pragma unroll 1
  for (int p = p_begin; p < p_end; p++) {
    fc3::cp_wait_all();
    __syncthreads();                       // planes of p landed; every x-stage read of Y from p-1 is done
    K::ystage(tid, p, IN, sE, Y);
    __syncthreads();                       // Y complete; IN consumed
    if (p + 1 < p_end) K::issue_loads(tid, cell, kz, p + 1, Z, IN);
    K::xstage(tid, Y, acc);
  }
  __syncthreads();
  K::xinverse(tid, acc, Y);                // T aliases Y
  __syncthreads();
  K::yinverse(tid, Y, IN);                 // T2 aliases IN
  __syncthreads();
  K::store(tid, cell, kz, IN, C);
}
// ---- F2 with the product accumulators parked in tensor memory (L = 16) ------------------------------------
// The x-stage thread needs acc (16 complex) across the seven p iterations, next to uh, vh and the butterflies of
// the transform in flight: > 200 live registers, one CTA per SM.  TMEM (256 KB per SM, idle in this kernel) takes
// acc instead: every thread owns 64 32-bit columns of its own lane (tcgen05.ld/st .32x32b), read-modify-written
// four complex values at a time when the products are formed.  Registers drop to <= 168 -> two CTAs per SM.
// With PARK the u transform waits in TMEM too while the v transform runs (another 64 columns per thread, the two
// CTAs of an SM then use all 512 columns): no spills at 168 registers and 6 % faster.
#define LP_TM_R16(r) "{%" #r "0, %" #r "1, %" #r "2, %" #r "3, %" #r "4, %" #r "5, %" #r "6, %" #r "7, %" #r "8, %" #r "9, %" #r "10, %" #r "11, %" #r "12, %" #r "13, %" #r "14, %" #r "15}"
__device__ __forceinline__ void tmem_ld4c(unsigned taddr, double2 (&a)[4])
{
  unsigned r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
               "tcgen05.wait::ld.sync.aligned;\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
  #pragma unroll
  for (int i = 0; i < 4; i++) a[i] = make_double2(__hiloint2double((int)r[4 * i + 1], (int)r[4 * i]), __hiloint2double((int)r[4 * i + 3], (int)r[4 * i + 2]));
}
__device__ __forceinline__ void tmem_st4c(unsigned taddr, const double2 (&a)[4])
{
  unsigned r[16];
  #pragma unroll
  for (int i = 0; i < 4; i++) {
    r[4 * i] = (unsigned)__double2loint(a[i].x); r[4 * i + 1] = (unsigned)__double2hiint(a[i].x);
    r[4 * i + 2] = (unsigned)__double2loint(a[i].y); r[4 * i + 3] = (unsigned)__double2hiint(a[i].y);
  }
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
               :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                  "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

// PARK: also park the u transform of the x stage in TMEM while the v transform runs (128 instead of 64 columns per thread)
#define LP_F2TM_PER (PARK ? 128 : 64)
#define LP_F2TM_COLS (2 * LP_F2TM_PER)   // 6 warps: lane quarters 0..3 twice -> two column groups
template <bool PARK, int NSPLIT>
__global__ void __launch_bounds__(192, 2) k_fc3_f2_tmem(const double2 *__restrict__ Z, const double *__restrict__ E, double2 *__restrict__ C,
                                                        long long split_stride)
{
  constexpr int L = 16;
  typedef fc3::F2<L> K;
  extern __shared__ double2 smf[];
  __shared__ unsigned s_tmem;
  __shared__ __align__(8) unsigned long long s_bar;        // completion of the bulk loads of one product's planes
  double2 *IN = smf, *Y = IN + K::IN_C2;
  double *sE = reinterpret_cast<double *>(Y + K::Y_C2);
  const int kz = blockIdx.x, cell = blockIdx.y, tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(&s_tmem);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(dst), "r"((unsigned)LP_F2TM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  int p_begin, p_end;
  K::psplit(blockIdx.z, NSPLIT, p_begin, p_end);       // gridDim.z = 3: the seven products split over three CTAs
  C += blockIdx.z * split_stride;
  // the planes u_p and (p != 1) the v source of p: one 16 KB bulk copy (TMA) each into IN, completed on s_bar in bytes
  constexpr unsigned PLANE_BYTES = K::N * K::N * sizeof(double2);
  auto issue_planes = [&](int p) {
    mbar_expect_tx(&s_bar, p == 1 ? PLANE_BYTES : 2 * PLANE_BYTES);
    bulk_load(IN, K::plane(Z, cell, p, kz), PLANE_BYTES, &s_bar);
    if (p != 1) bulk_load(IN + K::N * K::N, K::plane(Z, cell, 7 + fc3::zpow_of(p), kz), PLANE_BYTES, &s_bar);   // p = 1 reuses the y-transformed v of p = 0
  };
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    issue_planes(p_begin);
  }
  for (int i = tid; i < K::N; i += K::NT) sE[i] = E[i];
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const unsigned tbase = s_tmem;
  const unsigned tacc = tbase + ((unsigned)((warp & 3) * 32) << 16) + (unsigned)((warp >> 2) * LP_F2TM_PER);
  // x-stage task: 3M = 144 of them on the first XW = 5 warps; every lane of those warps runs (the tcgen05 instructions
  // are warp-wide), the idle half of the fifth on a duplicate line; the sixth warp skips the stage
  constexpr int XW = (3 * K::M + 31) / 32;
  int r, ky;
  const bool xok = K::xtask(tid, r, ky);
  #pragma unroll 1
  for (int p = p_begin; p < p_end; p++) {
    mbar_wait(&s_bar, (unsigned)(p - p_begin) & 1u);   // planes of p landed
    __syncthreads();                       // every x-stage read of Y from p-1 is done
    K::ystage(tid, p, IN, sE, Y);
    __syncthreads();                       // Y complete; IN consumed
    if (p + 1 < p_end && tid == 0) issue_planes(p + 1);
    if (warp < XW) {
      double2 a0[L], a1[L], uh[L], vh[L];
      #pragma unroll
      for (int l = 0; l < L; l++) { a0[l] = Y[l * K::PY + ky]; a1[l] = Y[(l + L) * K::PY + ky]; }
      fc3::fwd_third<L>(a0, a1, r, uh);
      __syncwarp();                         // the warp holding two r values diverged in the pre-stage; tcgen05.* is warp-wide
      if constexpr (PARK) {
        #pragma unroll
        for (int c4 = 0; c4 < L / 4; c4++) { double2 t4[4] = {uh[4 * c4], uh[4 * c4 + 1], uh[4 * c4 + 2], uh[4 * c4 + 3]}; tmem_st4c(tacc + 64 + 16 * c4, t4); }
        tmem_wait_st();
      }
      #pragma unroll
      for (int l = 0; l < L; l++) { a0[l] = Y[K::N * K::PY + l * K::PY + ky]; a1[l] = Y[K::N * K::PY + (l + L) * K::PY + ky]; }
      fc3::fwd_third<L>(a0, a1, r, vh);
      __syncwarp();
      #pragma unroll
      for (int c4 = 0; c4 < L / 4; c4++) {
        double2 a[4];
        if (p > p_begin) tmem_ld4c(tacc + 16 * c4, a);
        else { a[0] = a[1] = a[2] = a[3] = make_double2(0., 0.); }
        double2 u4[4];
        if constexpr (PARK) tmem_ld4c(tacc + 64 + 16 * c4, u4);
        #pragma unroll
        for (int i = 0; i < 4; i++) {
          const int q = 4 * c4 + i;
          const double2 uu = PARK ? u4[i] : uh[q];
          a[i].x += uu.x * vh[q].x - uu.y * vh[q].y;
          a[i].y += uu.x * vh[q].y + uu.y * vh[q].x;
        }
        tmem_st4c(tacc + 16 * c4, a);
      }
      tmem_wait_st();
    }
  }
  __syncthreads();
  if (warp < XW) {
    double2 acc[L];
    #pragma unroll
    for (int c4 = 0; c4 < L / 4; c4++) {
      double2 a[4];
      tmem_ld4c(tacc + 16 * c4, a);
      #pragma unroll
      for (int i = 0; i < 4; i++) acc[4 * c4 + i] = a[i];
    }
    if (xok) {
      fc3::inv_third<L>(acc, r);
      #pragma unroll
      for (int l = 0; l < L; l++) Y[(r * L + l) * K::PY + ky] = acc[l];      // T aliases Y
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tbase), "r"((unsigned)LP_F2TM_COLS) : "memory");
  K::yinverse(tid, Y, IN);                 // T2 aliases IN
  __syncthreads();
  K::store(tid, cell, kz, IN, C);
}
// ---------------------------------------------------------------------------------------------------------
// k_fc3_f2q: the same plane on THREE warps, four CTAs per SM (fc3::F2Q).  Inside one CTA the warps load together and
// compute together, so a CTA keeps the FP64 pipe busy 36 % and the shared-memory pipe 38 % of its time and two co-resident
// CTAs overlap those only as two independent streams do (profiles/r02_f2_kernels.md); four streams overlap more.  What
// makes four fit: one array at a time (41 KB of shared memory: one input plane + one Y array; the next plane is fetched
// into the input buffer while the x stage runs out of Y), the x stage as two halves per line (96 tasks = 96 threads, a
// line's inputs read twice instead of three times), the first transform of a product kept across the second array's y
// stage in registers, 32 spare TMEM columns and a 12 KB strip of shared memory (see below), and the 48 x 48 accumulators in
// 96 of the CTA's 128 TMEM columns (4 x 128 = all 512).
__global__ void __launch_bounds__(96, 4) k_fc3_f2q(const double2 *__restrict__ Z, const double *__restrict__ E, double2 *__restrict__ C)
{
  constexpr int L = 16;
  typedef fc3::F2Q<L> K;
  extern __shared__ double2 smf[];
  __shared__ unsigned s_tmem;
  __shared__ __align__(8) unsigned long long s_bar;
  double2 *IN = smf, *Y = IN + K::IN_C2;
  double *sE = reinterpret_cast<double *>(Y + K::Y_C2);
  const int kz = blockIdx.x, cell = blockIdx.y, tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(&s_tmem);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(dst), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  constexpr unsigned PLANE_BYTES = K::N * K::N * sizeof(double2);
  // The thirteen planes a CTA reads, in order: u_0, v-source of 0, u_1, then (u_p, v-source of p) for p = 2..6.  Product 1 has
  // no plane of its own for v: v_1 = -E(x)^2 fhat has the y transform of v_0 = fhat, which is still in Y after product 0 --
  // it is rescaled in place and transformed along x FIRST (so v_1's transform is the parked factor of product 1).
  auto plane_of = [&](int i) {
    const int p = i < 2 ? 0 : i == 2 ? 1 : 2 + (i - 3) / 2, arr = i < 2 ? i : i == 2 ? 0 : (i - 3) & 1;
    return K::plane(Z, cell, p, arr, kz);
  };
  auto issue_plane = [&](int i) {      // one thread: plane i into IN through the TMA unit, plane i + 1 on its way to L2
    mbar_expect_tx(&s_bar, PLANE_BYTES);
    bulk_load(IN, plane_of(i), PLANE_BYTES, &s_bar);
    if (i + 1 < 13) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" :: "l"(plane_of(i + 1)), "r"(PLANE_BYTES) : "memory");
  };
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    issue_plane(0);
  }
  for (int i = tid; i < K::N; i += K::NT) sE[i] = E[i];
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const unsigned tbase = s_tmem;
  const unsigned tacc = tbase + ((unsigned)(warp * 32) << 16);
  // fourteen array passes with ONE copy of the y- and x-stage code (two copies were 4096 instructions and lost 10 % of the
  // issue slots to instruction fetch with four CTAs in different phases).  The first transform of a product waits for the
  // second in registers, except its last sixteen values: eight go to the 32 TMEM columns the accumulators leave free, eight
  // to a strip of shared memory of their own (12 KB: four CTAs still fit an SM) -- with all 24 in registers the next y
  // stage spilled most of them to local memory, and the four CTAs then ran no faster than two of the six-warp kernel.
  constexpr int UR = K::H - 16;
  double2 uh[UR];
  double2 *upark = reinterpret_cast<double2 *>(sE + K::N) + tid;       // [8][NT]
  int issued = 1, waited = 0;
  #pragma unroll 1
  for (int k = 0; k < 14; k++) {
    const int p = k < 2 ? 0 : k < 4 ? 1 : 2 + (k - 4) / 2;
    const bool rescale = k == 2, first = k == 0 || k == 2 || (k >= 4 && !((k - 4) & 1));
    const int arr = k == 3 ? 0 : k < 2 ? k : (k - 4) & 1;
    if (rescale) {
      K::rescale_v1(tid, sE, Y);
    } else {
      mbar_wait(&s_bar, (unsigned)(waited++) & 1u);    // the plane landed
      K::ystage1(tid, p, arr, IN, sE, Y);
    }
    __syncthreads();                        // Y complete, IN consumed
    if (!rescale && issued < 13) {
      if (tid == 0) { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); issue_plane(issued); }   // arrives while the x stage runs
      issued++;
    }
    double2 t[K::H];
    K::xhalf(tid, Y, t);
    __syncwarp();                           // the warp holding both h diverged in the pre-stage; tcgen05.* is warp-wide
    if (first) {
      #pragma unroll
      for (int q = 0; q < UR; q++) uh[q] = t[q];
      #pragma unroll
      for (int q = 0; q < 8; q++) upark[q * K::NT] = t[UR + q];
      { double2 t4[4] = {t[UR + 8], t[UR + 9], t[UR + 10], t[UR + 11]}; tmem_st4c(tacc + 96, t4); }
      { double2 t4[4] = {t[UR + 12], t[UR + 13], t[UR + 14], t[UR + 15]}; tmem_st4c(tacc + 112, t4); }
      tmem_wait_st();
    } else {
      #pragma unroll
      for (int c4 = 0; c4 < K::H / 4; c4++) {
        double2 a[4];
        if (p > 0) tmem_ld4c(tacc + 16 * c4, a);
        else { a[0] = a[1] = a[2] = a[3] = make_double2(0., 0.); }
        double2 u4[4];
        if (4 * c4 >= UR + 8) tmem_ld4c(tacc + 96 + 4 * (4 * c4 - UR - 8), u4);
        else if (4 * c4 >= UR) {
          #pragma unroll
          for (int i = 0; i < 4; i++) u4[i] = upark[(4 * c4 - UR + i) * K::NT];
        }
        #pragma unroll
        for (int i = 0; i < 4; i++) {
          const int q = 4 * c4 + i;
          const double2 uu = 4 * c4 >= UR ? u4[i] : uh[q < UR ? q : 0];
          a[i].x += uu.x * t[q].x - uu.y * t[q].y;
          a[i].y += uu.x * t[q].y + uu.y * t[q].x;
        }
        tmem_st4c(tacc + 16 * c4, a);
      }
      tmem_wait_st();
    }
    __syncthreads();                        // every read of Y is done before the next pass writes it
  }
  {
    double2 acc[K::H];
    #pragma unroll
    for (int c4 = 0; c4 < K::H / 4; c4++) {
      double2 a[4];
      tmem_ld4c(tacc + 16 * c4, a);
      #pragma unroll
      for (int i = 0; i < 4; i++) acc[4 * c4 + i] = a[i];
    }
    K::xinverse(tid, acc, smf);             // T aliases IN | Y: no copy is in flight, every reader passed the barrier above
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tbase), "r"(128u) : "memory");
  {
    double2 c[L];
    K::yinverse_load(tid, smf, c);
    __syncthreads();                        // T2 aliases T
    K::yinverse_store(tid, c, smf);
  }
  __syncthreads();
  K::store(tid, cell, kz, smf, C);
}
// part (nullable): per (cell, xo) partial dot products of the conservation rows with the stored spectrum,
// [cell][xo][5]; folded in a fixed order by the kernels that apply the correction (collision.cu)
template <int L, int NSPLIT>
__global__ void __launch_bounds__(fc3::F3<L>::NT) k_fc3_f3(const double2 *__restrict__ C, double2 *__restrict__ q,
                                                           const double *__restrict__ C5, double *__restrict__ part, long long split_stride)
{
  typedef fc3::F3<L> K;
  __shared__ double2 T3[K::SMEM_C2];
  __shared__ double red[5][4];
  const int xo = blockIdx.x, cell = blockIdx.y, tid = threadIdx.x;
  K::template zinverse<NSPLIT>(tid, cell, xo, C, T3, split_stride);
  __syncthreads();
  double s[5] = {0., 0., 0., 0., 0.};
  K::store(tid, cell, xo, T3, q, part ? C5 : nullptr, s);
  if (part) {
    const int lane = tid & 31, wid = tid >> 5, nw = (K::NT + 31) / 32;
    #pragma unroll
    for (int m = 0; m < 5; m++) {
      double v = s[m];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if (lane == 0) red[m][wid] = v;
    }
    __syncthreads();
    if (tid < 5) { double v = 0.; for (int w = 0; w < nw; w++) v += red[tid][w]; part[((long long)cell * K::N + xo) * 5 + tid] = v; }
  }
}
template <int L>
int launch_fc3(lpgpu_ctx *c, const double2 *fh, double2 *Z, double2 *C, double2 *qo, int nb, bool fused_i, double *part, const double2 *mhat,
               const double *Gt, long long mstride)
{
  typedef fc3::F2<L> K2;
  constexpr int N = 2 * L, M = 3 * L;
  const size_t smem2 = (size_t)(K2::IN_C2 + K2::Y_C2) * sizeof(double2) + N * sizeof(double);
  const size_t smem1_max = (size_t)2 * fc3::F1<L>::SMEM_C2 * sizeof(double2) + (size_t)(7 * N * N + N) * sizeof(double);
  const size_t smem1 = smem1_max - (mhat ? 0 : (size_t)fc3::F1<L>::SMEM_C2 * sizeof(double2));
  if (!c->fc3_attr) {
    LP_CUDA(cudaFuncSetAttribute(k_fc3_f1<L, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1_max));
    LP_CUDA(cudaFuncSetAttribute(k_fc3_f1<L, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1_max));
    LP_CUDA(cudaFuncSetAttribute(k_fc3_f2<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    if (L == 16) {
      LP_CUDA(cudaFuncSetAttribute(k_fc3_f2_tmem<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      LP_CUDA(cudaFuncSetAttribute(k_fc3_f2_tmem<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      LP_CUDA(cudaFuncSetAttribute(k_fc3_f2_tmem<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    }
    c->fc3_attr = true;
  }
  const double *E = c->d_Etab + LP_ETAB_PAD;
  const double2 *post = reinterpret_cast<const double2 *>(c->d_post_fwd);
  const dim3 g1(N, nb, nb * N * 2 <= 148 ? 5 : 1);
  if (fused_i) k_fc3_f1<L, true><<<g1, fc3::F1<L>::NT, smem1, c->stream>>>(fh, Gt, E, Z, post, mhat, mstride);
  else k_fc3_f1<L, false><<<g1, fc3::F1<L>::NT, smem1, c->stream>>>(fh, Gt, E, Z, post, mhat, mstride);
  LP_LAUNCHED(c);
  static const bool no_tmem = getenv("LPGPU_FC_NO_TMEM") != nullptr;   // developer knob: accumulators in registers, 1 CTA per SM
  const bool prof2 = c->prof_on == 2 && c->prof_used + 2 <= c->prof_ev.size();
  if (prof2) LP_CUDA(cudaEventRecord(c->prof_ev[c->prof_used], c->stream));
  static const bool park_uh = getenv("LPGPU_F2_NO_PARK_UH") == nullptr;   // developer knob: keep the u transform in registers (6 % slower)
  // few cells in flight (the homogeneous single cell: M CTAs on 148 SMs): split the seven products over three CTAs per
  // (cell, kz); each writes its partial C and F3 sums them (the inverse transforms are linear)
  const int nsplit = (nb * M * 2 <= 148) ? 3 : 1;
  const long long split_stride = (long long)nb * M * N * N;
  const dim3 g2(M, nb, nsplit);
  static const bool quarter = getenv("LPGPU_F2_SIX_WARPS") == nullptr;   // developer knob: LPGPU_F2_SIX_WARPS=1 runs the round-1 kernel (six warps, two CTAs per SM)
  if (L == 16 && !no_tmem && nsplit == 1 && quarter) {
    typedef fc3::F2Q<16> KQ;
    static const int padq = getenv("LPGPU_F2Q_PAD_KB") ? atoi(getenv("LPGPU_F2Q_PAD_KB")) : 0;   // experiment: fewer CTAs per SM
    const size_t smemq = (size_t)(KQ::IN_C2 + KQ::Y_C2 + 8 * KQ::NT) * sizeof(double2) + KQ::N * sizeof(double) + (size_t)padq * 1024;
    LP_CUDA(cudaFuncSetAttribute(k_fc3_f2q, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemq));
    LP_CUDA(cudaFuncSetAttribute(k_fc3_f2q, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    k_fc3_f2q<<<dim3(M, nb), KQ::NT, smemq, c->stream>>>(Z, E, C);
  }
  else if (L == 16 && !no_tmem && nsplit == 3) k_fc3_f2_tmem<true, 3><<<g2, K2::NT, smem2, c->stream>>>(Z, E, C, split_stride);
  else if (L == 16 && !no_tmem && park_uh) k_fc3_f2_tmem<true, 1><<<g2, K2::NT, smem2, c->stream>>>(Z, E, C, split_stride);
  else if (L == 16 && !no_tmem) k_fc3_f2_tmem<false, 1><<<g2, K2::NT, smem2, c->stream>>>(Z, E, C, split_stride);
  else k_fc3_f2<L><<<g2, K2::NT, smem2, c->stream>>>(Z, E, C, split_stride);
  LP_LAUNCHED(c);
  if (prof2) { LP_CUDA(cudaEventRecord(c->prof_ev[c->prof_used + 1], c->stream)); c->prof_used += 2; }
  if (nsplit == 3) k_fc3_f3<L, 3><<<dim3(N, nb), fc3::F3<L>::NT, 0, c->stream>>>(C, qo, c->d_C5, part, split_stride);
  else k_fc3_f3<L, 1><<<dim3(N, nb), fc3::F3<L>::NT, 0, c->stream>>>(C, qo, c->d_C5, part, split_stride);
  LP_LAUNCHED(c);
  return LPGPU_OK;
}

// factor M into radix-3 and radix-2 stages (3s first); false if another prime divides M
bool make_plan(int M, FcPlan &pl)
{
  pl.M = M; pl.nst = 0;
  int m = M;
  while (m % 3 == 0 && pl.nst < 8) { pl.radix[pl.nst++] = 3; m /= 3; }
  while (m % 2 == 0 && pl.nst < 8) { pl.radix[pl.nst++] = 2; m /= 2; }
  return m == 1 && M <= 64 && M >= 2;
}

} // namespace

// returns -1 when M = 3N/2 is not of the form 2^a 3^b (caller uses the tiled direct kernel)
static bool fc3_knobs_off()
{
  static const bool off = getenv("LPGPU_FFT_GENERIC") != nullptr;
  return off;
}
bool lp_fc3_available(const lpgpu_ctx *c)
{
  const int N = c->p.N, v = c->p.computeq_variant;
  return (v == 0 || v == 2) && !fc3_knobs_off() && (N == 32 || N == 24 || N == 16 || N == 8);
}
// fused_i: `fhat` is the output of lp_launch_fft3d_jk (only valid when lp_fc3_available)
// work arrays, twiddles and re-laid symbols of the FFT-convolution pipeline (idempotent; first use allocates)
int lp_fc_prepare(lpgpu_ctx *c)
{
  const int N = c->p.N, M = 3 * N / 2;
  FcPlan pl;
  if ((N & 1) || !make_plan(M, pl)) return -1;
  const size_t plane_bytes = ((size_t)M * (M + 1) + M) * sizeof(double2);
  if (!c->d_fc1) {
    // chunk of cells whose transformed arrays fit a fixed budget: 10 arrays of N^2 M for the register-resident pipeline
    // (7.9 MB per cell at N = 32), 14 of N M^2 for the shared-memory fallback
    const bool fc3_size = (N == 32 || N == 24 || N == 16 || N == 8) && !fc3_knobs_off();
    const size_t per_cell = fc3_size ? (size_t)10 * N * N * M * sizeof(double2) : (size_t)14 * N * M * M * sizeof(double2);
    // Transformed planes of one chunk: up to a third of the free device memory, 64 GB at most (7.9 MB per cell at N = 32:
    // the 512 cells of the largest BASELINE config take 4 GB), so that in practice one launch covers every local cell.
    // Measured in round 1: chunks small enough for F2 to read F1's output out of the 126 MB L2 (96 MB) are 15 % slower
    // than one big launch -- the kernels are not HBM-limited, launch tails are.  LPGPU_FC_CHUNK_MB overrides (tests).
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = (size_t)8 << 30; }
    size_t auto_mb = (free_b / 3) >> 20;
    if (auto_mb > 65536) auto_mb = 65536;
    if (auto_mb < 256) auto_mb = 256;
    const size_t budget_mb = getenv("LPGPU_FC_CHUNK_MB") ? (size_t)atoi(getenv("LPGPU_FC_CHUNK_MB")) : auto_mb;
    size_t chunk = (budget_mb << 20) / per_cell;
    if (chunk < 1) chunk = 1;
    if (chunk > c->cap_cells) chunk = c->cap_cells;
    // equal chunks: a 2-cell tail after a 62-cell chunk would run nearly empty grids
    { const size_t nchunks = (c->cap_cells + chunk - 1) / chunk; chunk = (c->cap_cells + nchunks - 1) / nchunks; }
    c->fc_chunk = (int)chunk;
    LP_CUDA(cudaMalloc((void **)&c->d_fc1, per_cell * chunk));
    LP_CUDA(cudaMalloc((void **)&c->d_fc2, (size_t)N * M * M * sizeof(double2) * 2 * chunk));   // room for the three partial C arrays of a split launch (3 N^2 M = 2 N M^2)
    std::vector<double> tw(2 * M);   // exp(-2 pi i t / M), t < M
    for (int t = 0; t < M; t++) { const long double a = 2.0L * M_PIl * t / M; tw[2 * t] = (double)cosl(a); tw[2 * t + 1] = (double)(-sinl(a)); }
    LP_CUDA(cudaMalloc((void **)&c->d_fctw, 2 * M * sizeof(double)));
    LP_CUDA(cudaMemcpy(c->d_fctw, tw.data(), 2 * M * sizeof(double), cudaMemcpyHostToDevice));
    {
      // kernel symbols re-laid as Gt[a][y][z][x] (x fastest): the lanes of an F1 warp are consecutive x
      std::vector<double> gt((size_t)7 * c->N3);
      for (int a = 0; a < 7; a++)
        for (int x = 0; x < N; x++)
          for (int y = 0; y < N; y++)
            for (int z = 0; z < N; z++) gt[(((size_t)a * N + y) * N + z) * N + x] = c->tab.G[(size_t)7 * (z + N * (y + N * x)) + a];
      LP_CUDA(cudaMalloc((void **)&c->d_Gt, gt.size() * sizeof(double)));
      LP_CUDA(cudaMemcpy(c->d_Gt, gt.data(), gt.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    LP_CUDA(cudaFuncSetAttribute(k_fc_fwd_yz, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plane_bytes));
    LP_CUDA(cudaFuncSetAttribute(k_fc_inv_yz, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plane_bytes));
  }
  return LPGPU_OK;
}
// The pipeline with another set of symbols and another first factor (register-resident pipeline only, unfused):
//   q[xi] = sum_p ( (Gt_p FU) (*) v_p )[xi + N/2],  v_p the seven monomials of E times fhat,  FU = `first` (one spectrum shared by
// all cells when first_stride = 0).  ComputeQ_FandL's linear part is built from such passes (collision.cu).
static int fftconv_run(lpgpu_ctx *c, const double *fhat, double *q, int B, bool fused_i, double *part, const double *Gt, const double *first, long long first_stride);
int lp_launch_fftconv_with(lpgpu_ctx *c, const double *fhat, double *q, int B, const double *Gt, const double *first, long long first_stride)
{
  if (!lp_fc3_available(c) || !Gt || !first) { lp_set_error("lp_launch_fftconv_with: needs the fc3 pipeline, a symbol table and a first factor"); return LPGPU_EINVAL; }
  return fftconv_run(c, fhat, q, B, false, nullptr, Gt, first, first_stride);
}
int lp_launch_computeQ_fftconv(lpgpu_ctx *c, const double *fhat, double *q, int B, bool fused_i, double *part)
{
  return fftconv_run(c, fhat, q, B, fused_i, part, nullptr, nullptr, 0);
}
static int fftconv_run(lpgpu_ctx *c, const double *fhat, double *q, int B, bool fused_i, double *part, const double *Gt_other, const double *first, long long first_stride)
{
  if ((fused_i || part) && !lp_fc3_available(c)) { lp_set_error("fused ComputeQ needs the fc3 pipeline"); return LPGPU_EINVAL; }
  const int N = c->p.N, M = 3 * N / 2;
  FcPlan pl;
  if ((N & 1) || !make_plan(M, pl)) return -1;
  const size_t plane_bytes = ((size_t)M * (M + 1) + M) * sizeof(double2);
  const size_t line_bytes = ((size_t)16 * (M + 1) + M) * sizeof(double2);
  { const int rc = lp_fc_prepare(c); if (rc != LPGPU_OK) return rc; }
  const double2 *tw = reinterpret_cast<const double2 *>(c->d_fctw);
  const bool prof = c->prof_on == 1 && c->prof_used + 2 <= c->prof_ev.size();
  if (prof) LP_CUDA(cudaEventRecord(c->prof_ev[c->prof_used], c->stream));
  for (int b0 = 0; b0 < B; b0 += c->fc_chunk) {
    const int nb = B - b0 < c->fc_chunk ? B - b0 : c->fc_chunk;
    const double2 *fh = reinterpret_cast<const double2 *>(fhat) + (size_t)b0 * c->N3;
    double2 *qo = reinterpret_cast<double2 *>(q) + (size_t)b0 * c->N3;
    double2 *F1 = reinterpret_cast<double2 *>(c->d_fc1), *F2 = reinterpret_cast<double2 *>(c->d_fc2);
    static const bool generic_only = getenv("LPGPU_FFT_GENERIC") != nullptr;   // developer knob: force the shared-memory stages
    if (!generic_only && (N == 32 || N == 24 || N == 16 || N == 8)) {
      double *pp = part ? part + (size_t)b0 * N * 5 : nullptr;
      // LinearLandau: cell b of this call pairs with cell b of the context's stored Maxwellian transforms
      const double2 *mh = (c->p.linear_landau && c->have_mhat) ? reinterpret_cast<const double2 *>(c->d_mhat) + (size_t)b0 * c->N3 : nullptr;
      long long ms = c->N3;
      const double *Gt = c->d_Gt;
      if (first) { mh = reinterpret_cast<const double2 *>(first) + (size_t)b0 * first_stride; ms = first_stride; Gt = Gt_other; }
      int rc = N == 32 ? launch_fc3<16>(c, fh, F1, F2, qo, nb, fused_i, pp, mh, Gt, ms) : N == 24 ? launch_fc3<12>(c, fh, F1, F2, qo, nb, fused_i, pp, mh, Gt, ms)
             : N == 16 ? launch_fc3<8>(c, fh, F1, F2, qo, nb, fused_i, pp, mh, Gt, ms) : launch_fc3<4>(c, fh, F1, F2, qo, nb, fused_i, pp, mh, Gt, ms);
      if (rc != LPGPU_OK) return rc;
      continue;
    }
    k_fc_fwd_yz<<<dim3(N, 14, nb), 256, plane_bytes, c->stream>>>(fh, F1, c->d_G, c->d_Etab, tw, N, pl);
    LP_LAUNCHED(c);
    k_fc_x<<<dim3((M + 7) / 8, M, nb), 256, line_bytes, c->stream>>>(F1, F2, tw, N, pl);
    LP_LAUNCHED(c);
    k_fc_inv_yz<<<dim3(N, nb), 256, plane_bytes, c->stream>>>(F2, qo, tw, N, pl);
    LP_LAUNCHED(c);
  }
  if (prof) { LP_CUDA(cudaEventRecord(c->prof_ev[c->prof_used + 1], c->stream)); c->prof_used += 2; }
  return LPGPU_OK;
}
