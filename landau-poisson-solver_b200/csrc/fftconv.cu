// ComputeQ as seven zero-padded linear convolutions (computeq_variant = 2) -- the O(N^3 log N) form of the
// same sum (identity checked against the reference's table to 1.5e-16):
//
//   Wt(xi,omega) = G0(omega) - sum_{p=1..6} G_p(omega) mono_p(E(beta)),   beta = xi + N/2 - omega,
//   mono = {e1^2, e2^2, e3^2, e1 e2, e1 e3, e2 e3},  e_a = E(beta_a) = eta[beta_a] - eta[N/2]
//   => Qhat[xi] = sum_p ( u_p (*) v_p )[xi + N/2],   u_p = G_p fhat,  v_p = h_p(E) fhat   (linear convolution)
//
// computed with cyclic transforms of size M = 2N per dimension (no aliasing since M >= 2N-1):
//   F1  per (cell, x, p): build the padded y-z plane of u_p / v_p, DIF-FFT along z (N non-zero rows) and y
//   F2  per (cell, ky, 8 kz): DIF-FFT along x of the 14 lines, sum_p u_p v_p, inverse DIT along x, keep N outputs
//   F3  per (cell, x'): inverse DIT along y and z, scale by M^-3, extract the N x N window
// Forward transforms are decimation-in-frequency (natural in, bit-reversed out), inverse transforms are
// decimation-in-time (bit-reversed in, natural out), so no permutation pass exists anywhere: products are
// formed position-wise in bit-reversed order.  Radix-2 butterflies on shared-memory lines, correctly rounded
// twiddles from the host.  Power-of-two N only (N = 8, 16, 32); other sizes use the tiled direct kernel.
#include "lpgpu_internal.h"

#define LP_LAUNCHED(c)                                  \
  do {                                                  \
    (c)->launches++;                                    \
    LP_CUDA(cudaGetLastError());                        \
  } while (0)

namespace {

__device__ __forceinline__ double2 cxmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cxmulc(double2 a, double2 b) { return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }   // a * conj(b)

// one DIF (forward) butterfly: span h, butterfly index b in [0, M/2); element i of the line is x[i*st]
__device__ __forceinline__ void dif_bfly(double2 *x, int st, int M, int h, int b, const double2 *tw)
{
  const int j = b & (h - 1), i0 = ((b - j) << 1) + j, i1 = i0 + h;
  const double2 a = x[i0 * st], c = x[i1 * st];
  x[i0 * st] = make_double2(a.x + c.x, a.y + c.y);
  x[i1 * st] = cxmul(make_double2(a.x - c.x, a.y - c.y), tw[j * (M / (2 * h))]);
}
// one DIT (inverse) butterfly: twiddle conjugated
__device__ __forceinline__ void dit_bfly(double2 *x, int st, int M, int h, int b, const double2 *tw)
{
  const int j = b & (h - 1), i0 = ((b - j) << 1) + j, i1 = i0 + h;
  const double2 a = x[i0 * st], t = cxmulc(x[i1 * st], tw[j * (M / (2 * h))]);
  x[i0 * st] = make_double2(a.x + t.x, a.y + t.y);
  x[i1 * st] = make_double2(a.x - t.x, a.y - t.y);
}
// whole line by one warp: lanes = butterflies (M/2 <= 32)
template <bool FWD>
__device__ __forceinline__ void line_fft_warp(double2 *x, int M, const double2 *tw, int lane)
{
  if (FWD) {
    for (int h = M / 2; h >= 1; h >>= 1) { if (lane < M / 2) dif_bfly(x, 1, M, h, lane, tw); __syncwarp(); }
  } else {
    for (int h = 1; h <= M / 2; h <<= 1) { if (lane < M / 2) dit_bfly(x, 1, M, h, lane, tw); __syncwarp(); }
  }
}
// all M columns of a [M][P] plane by the whole block: lanes = columns, warps share the butterflies of a stage
template <bool FWD>
__device__ __forceinline__ void columns_fft_block(double2 *plane, int M, int P, const double2 *tw)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (FWD) {
    for (int h = M / 2; h >= 1; h >>= 1) {
      for (int b = warp; b < M / 2; b += nw)
        for (int col = lane; col < M; col += 32) dif_bfly(plane + col, P, M, h, b, tw);
      __syncthreads();
    }
  } else {
    for (int h = 1; h <= M / 2; h <<= 1) {
      for (int b = warp; b < M / 2; b += nw)
        for (int col = lane; col < M; col += 32) dit_bfly(plane + col, P, M, h, b, tw);
      __syncthreads();
    }
  }
}

// F1: padded y-z plane of u_p (p < 7) or v_{p-7}, transformed along z and y
__global__ void __launch_bounds__(256) k_fc_fwd_yz(const double2 *__restrict__ fhat, double2 *__restrict__ Fxy, const double *__restrict__ G,
                                                   const double *__restrict__ Etab, const double2 *__restrict__ twg, int N)
{
  extern __shared__ double2 smf[];
  const int M = 2 * N, P = M + 1;
  double2 *plane = smf, *tw = plane + M * P;
  const int x = blockIdx.x, p = blockIdx.y; const long long cell = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  for (int t = tid; t < M / 2; t += blockDim.x) tw[t] = twg[t];
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
  for (int y = warp; y < N; y += nw) line_fft_warp<true>(plane + y * P, M, tw, lane);
  __syncthreads();
  columns_fft_block<true>(plane, M, P, tw);
  double2 *o = Fxy + ((cell * 14 + p) * N + x) * (long long)(M * M);
  for (int t = tid; t < M * M; t += blockDim.x) o[t] = plane[(t / M) * P + (t % M)];
}

// F2: x-lines of the 14 arrays at 8 consecutive kz: transform, multiply-accumulate over p, inverse transform
__global__ void __launch_bounds__(256) k_fc_x(const double2 *__restrict__ Fxy, double2 *__restrict__ Cx, const double2 *__restrict__ twg, int N)
{
  extern __shared__ double2 smf[];
  const int M = 2 * N, P = M + 1, H = N / 2;
  double2 *U = smf, *V = U + 8 * P, *tw = V + 8 * P;
  const int kz0 = blockIdx.x * 8, ky = blockIdx.y; const long long cell = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;   // 8 warps: warp = line
  for (int t = tid; t < M / 2; t += blockDim.x) tw[t] = twg[t];
  for (int t = tid; t < 8 * P; t += blockDim.x) { U[t] = make_double2(0., 0.); V[t] = make_double2(0., 0.); }
  double2 acc[2] = {make_double2(0., 0.), make_double2(0., 0.)};
  const long long MM = (long long)M * M;
  for (int p = 0; p < 7; p++) {
    __syncthreads();
    const double2 *su = Fxy + ((cell * 14 + p) * N) * MM + (long long)ky * M + kz0;
    const double2 *sv = Fxy + ((cell * 14 + p + 7) * N) * MM + (long long)ky * M + kz0;
    for (int t = tid; t < 8 * N; t += blockDim.x) {
      const int x = t >> 3, kzi = t & 7;
      U[kzi * P + x] = su[x * MM + kzi];
      V[kzi * P + x] = sv[x * MM + kzi];
    }
    for (int t = tid; t < 8 * N; t += blockDim.x) {           // re-zero the padded half (previous transform filled it)
      const int x = N + (t >> 3), kzi = t & 7;
      U[kzi * P + x] = make_double2(0., 0.); V[kzi * P + x] = make_double2(0., 0.);
    }
    __syncthreads();
    line_fft_warp<true>(U + warp * P, M, tw, lane);
    line_fft_warp<true>(V + warp * P, M, tw, lane);
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
  line_fft_warp<false>(U + warp * P, M, tw, lane);
  __syncthreads();
  double2 *o = Cx + (cell * N) * MM + (long long)ky * M + kz0;
  for (int t = tid; t < 8 * N; t += blockDim.x) {
    const int xo = t >> 3, kzi = t & 7;
    o[xo * MM + kzi] = U[kzi * P + xo + H];
  }
}

// F3: inverse along y and z of one x' plane, scale, extract the [N/2, N/2+N) window
__global__ void __launch_bounds__(256) k_fc_inv_yz(const double2 *__restrict__ Cx, double2 *__restrict__ q, const double2 *__restrict__ twg, int N)
{
  extern __shared__ double2 smf[];
  const int M = 2 * N, P = M + 1, H = N / 2;
  double2 *plane = smf, *tw = plane + M * P;
  const int xo = blockIdx.x; const long long cell = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  for (int t = tid; t < M / 2; t += blockDim.x) tw[t] = twg[t];
  const double2 *s = Cx + (cell * N + xo) * (long long)(M * M);
  for (int t = tid; t < M * M; t += blockDim.x) plane[(t / M) * P + (t % M)] = s[t];
  __syncthreads();
  columns_fft_block<false>(plane, M, P, tw);
  for (int y = H + warp; y < H + N; y += nw) line_fft_warp<false>(plane + y * P, M, tw, lane);
  __syncthreads();
  const double sc = 1.0 / ((double)M * M * M);
  double2 *o = q + (cell * N + xo) * (long long)(N * N);
  for (int t = tid; t < N * N; t += blockDim.x) {
    const double2 v = plane[(t / N + H) * P + (t % N) + H];
    o[t] = make_double2(v.x * sc, v.y * sc);
  }
}

} // namespace

// returns -1 when N is not a power of two (caller uses the tiled direct kernel)
int lp_launch_computeQ_fftconv(lpgpu_ctx *c, const double *fhat, double *q, int B)
{
  const int N = c->p.N, M = 2 * N;
  if (N & (N - 1)) return -1;
  const size_t plane_bytes = ((size_t)M * (M + 1) + M / 2) * sizeof(double2);
  const size_t line_bytes = ((size_t)16 * (M + 1) + M / 2) * sizeof(double2);
  if (!c->d_fc1) {
    // chunk of cells whose 14 transformed arrays fit a fixed budget (29 MB per cell at N = 32)
    const size_t per_cell = (size_t)14 * N * M * M * sizeof(double2);
    size_t chunk = (size_t)1 << 30;
    chunk /= per_cell;
    if (chunk < 1) chunk = 1;
    if (chunk > c->cap_cells) chunk = c->cap_cells;
    c->fc_chunk = (int)chunk;
    LP_CUDA(cudaMalloc((void **)&c->d_fc1, per_cell * chunk));
    LP_CUDA(cudaMalloc((void **)&c->d_fc2, (size_t)N * M * M * sizeof(double2) * chunk));
    std::vector<double> tw(M);   // exp(-2 pi i k / M), k < M/2
    for (int k = 0; k < M / 2; k++) { const long double a = 2.0L * M_PIl * k / M; tw[2 * k] = (double)cosl(a); tw[2 * k + 1] = (double)(-sinl(a)); }
    LP_CUDA(cudaMalloc((void **)&c->d_fctw, M * sizeof(double)));
    LP_CUDA(cudaMemcpy(c->d_fctw, tw.data(), M * sizeof(double), cudaMemcpyHostToDevice));
    LP_CUDA(cudaFuncSetAttribute(k_fc_fwd_yz, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plane_bytes));
    LP_CUDA(cudaFuncSetAttribute(k_fc_inv_yz, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plane_bytes));
  }
  const double2 *tw = reinterpret_cast<const double2 *>(c->d_fctw);
  const bool prof = c->prof_on && c->prof_used + 2 <= c->prof_ev.size();
  if (prof) LP_CUDA(cudaEventRecord(c->prof_ev[c->prof_used], c->stream));
  for (int b0 = 0; b0 < B; b0 += c->fc_chunk) {
    const int nb = B - b0 < c->fc_chunk ? B - b0 : c->fc_chunk;
    const double2 *fh = reinterpret_cast<const double2 *>(fhat) + (size_t)b0 * c->N3;
    double2 *qo = reinterpret_cast<double2 *>(q) + (size_t)b0 * c->N3;
    double2 *F1 = reinterpret_cast<double2 *>(c->d_fc1), *F2 = reinterpret_cast<double2 *>(c->d_fc2);
    k_fc_fwd_yz<<<dim3(N, 14, nb), 256, plane_bytes, c->stream>>>(fh, F1, c->d_G, c->d_Etab, tw, N);
    LP_LAUNCHED(c);
    k_fc_x<<<dim3(M / 8, M, nb), 256, line_bytes, c->stream>>>(F1, F2, tw, N);
    LP_LAUNCHED(c);
    k_fc_inv_yz<<<dim3(N, nb), 256, plane_bytes, c->stream>>>(F2, qo, tw, N);
    LP_LAUNCHED(c);
  }
  if (prof) { LP_CUDA(cudaEventRecord(c->prof_ev[c->prof_used + 1], c->stream)); c->prof_used += 2; }
  return LPGPU_OK;
}
