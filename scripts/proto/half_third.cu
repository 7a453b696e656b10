// Feasibility probe for DESIGN.md section 8 item 1 (not part of the library): a 16-point third of the 48-point line
// transform shared by TWO threads.  Thread h of a pair owns the residue classes n2 = 2h, 2h+1 of the 4 x 4 split
// (n = 4 n1 + n2): it loads its 8 of the 16 (a0, a1) input pairs, runs the radix-3 pre-stage and the first radix-4 stage
// for them, swaps four complex values with its partner by shuffles and finishes the outputs k = k1 + 4 k2 for
// k1 = 2h, 2h+1.  Compile with -Xptxas -v to read the register count:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Xptxas -v -I../../landau-poisson-solver_b200/csrc -c half_third.cu -o /dev/null
#include "fc3.cuh"
using namespace fc3;

template <int SIGN>
__device__ __forceinline__ double2 shfl_xor_c(double2 v, int m)
{
  return make_double2(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m));
}
// X[k1 + 4 k2], k1 in {2h, 2h+1}, from the 2L = 32 inputs of the line (L = 16, M = 48); partner = lane ^ 1
__device__ __forceinline__ void fwd_half_third(const double2 *__restrict__ src, int stride, int r, int h, double2 (&out)[8])
{
  constexpr int L = 16, M = 48;
  double2 y1[2][4];                                     // [n2 - 2h][k1] after the first stage
  #pragma unroll
  for (int c = 0; c < 2; c++) {
    const int n2 = 2 * h + c;
    double2 v[4];
    #pragma unroll
    for (int n1 = 0; n1 < 4; n1++) {
      const int l = 4 * n1 + n2;
      const double2 a0 = src[l * stride], a1 = src[(l + L) * stride];
      double2 b;
      if (r == 0) b = cadd(a0, a1);
      else {
        const double sg = r == 1 ? LP_SQ3H : -LP_SQ3H;
        b = make_double2(a0.x - 0.5 * a1.x + sg * a1.y, a0.y - 0.5 * a1.y - sg * a1.x);
        // w_M^(r l): l is not a compile-time constant here (n2 depends on h), so the twiddle comes from the table
        const int t = (r * l) % M;
        const double cs = Tw<M>::c(t), sn = -Tw<M>::s(t);
        b = make_double2(b.x * cs - b.y * sn, b.x * sn + b.y * cs);
      }
      v[n1] = b;
    }
    dft4<-1>(v[0], v[1], v[2], v[3]);
    #pragma unroll
    for (int k1 = 0; k1 < 4; k1++) {
      const int t = (3 * n2 * k1) % M;                  // w_16^(n2 k1) = w_48^(3 n2 k1)
      const double cs = Tw<M>::c(t), sn = -Tw<M>::s(t);
      y1[c][k1] = make_double2(v[k1].x * cs - v[k1].y * sn, v[k1].x * sn + v[k1].y * cs);
    }
  }
  // keep k1 = 2h, 2h+1; send the other two k1 of both residue classes to the partner
  #pragma unroll
  for (int kk = 0; kk < 2; kk++) {
    double2 mine[4];                                    // n2 = 0..3 for k1 = 2h + kk
    #pragma unroll
    for (int c = 0; c < 2; c++) {
      const double2 keep = h == 0 ? y1[c][kk] : y1[c][2 + kk];
      const double2 send = h == 0 ? y1[c][2 + kk] : y1[c][kk];
      const double2 got = shfl_xor_c<0>(send, 1);
      mine[c] = h == 0 ? keep : got;                    // n2 = c     (owned by the h = 0 thread)
      mine[2 + c] = h == 0 ? got : keep;                // n2 = 2 + c (owned by the h = 1 thread)
    }
    dft4<-1>(mine[0], mine[1], mine[2], mine[3]);
    #pragma unroll
    for (int k2 = 0; k2 < 4; k2++) out[4 * kk + k2] = mine[k2];       // X[(2h + kk) + 4 k2]
  }
}
// y stage of F2 with half-third tasks: thread = (array, r, x, h); 384 threads per CTA
__global__ void __launch_bounds__(384, 2) k_probe_ystage(const double2 *__restrict__ IN, double2 *__restrict__ Y)
{
  const int tid = threadIdx.x, h = tid & 1, x = (tid >> 1) & 31, r = (tid >> 6) % 3, arr = tid / 192;
  double2 out[8];
  fwd_half_third(IN + ((long long)blockIdx.x * 2 + arr) * 1024 + x, 32, r, h, out);
  double2 *dst = Y + (((long long)blockIdx.x * 2 + arr) * 32 + x) * 49 + r * 16;
  #pragma unroll
  for (int kk = 0; kk < 2; kk++)
    #pragma unroll
    for (int k2 = 0; k2 < 4; k2++) dst[(2 * h + kk) + 4 * k2] = out[4 * kk + k2];
}
// x stage: two half-thirds (u, v) and the product into eight register accumulators over seven products
__global__ void __launch_bounds__(384, 2) k_probe_xstage(const double2 *__restrict__ Y, double2 *__restrict__ C)
{
  const int tid = threadIdx.x, h = tid & 1, task = tid >> 1, r = task / 48 > 2 ? 2 : task / 48, ky = task % 48;
  double2 acc[8];
  #pragma unroll
  for (int q = 0; q < 8; q++) acc[q] = make_double2(0., 0.);
  #pragma unroll 1
  for (int p = 0; p < 7; p++) {
    double2 uh[8], vh[8];
    fwd_half_third(Y + ((long long)(blockIdx.x * 7 + p) * 2) * 32 * 49 + ky, 49, r, h, uh);
    fwd_half_third(Y + ((long long)(blockIdx.x * 7 + p) * 2 + 1) * 32 * 49 + ky, 49, r, h, vh);
    #pragma unroll
    for (int q = 0; q < 8; q++) { acc[q].x += uh[q].x * vh[q].x - uh[q].y * vh[q].y; acc[q].y += uh[q].x * vh[q].y + uh[q].y * vh[q].x; }
  }
  #pragma unroll
  for (int q = 0; q < 8; q++) C[((long long)blockIdx.x * 192 + task) * 16 + 8 * h + q] = acc[q];
}
// the same with the accumulators (and the waiting u transform) outside the register file, as tensor memory holds them in
// k_fc3_f2_tmem: here a read-modify-write of C stands in for the tcgen05.ld/st pair
__global__ void __launch_bounds__(384, 2) k_probe_xstage_parked(const double2 *__restrict__ Y, double2 *__restrict__ C, double2 *__restrict__ P)
{
  const int tid = threadIdx.x, h = tid & 1, task = tid >> 1, r = task / 48 > 2 ? 2 : task / 48, ky = task % 48;
  double2 *acc = C + ((long long)blockIdx.x * 192 + task) * 16 + 8 * h, *park = P + ((long long)blockIdx.x * 192 + task) * 16 + 8 * h;
  #pragma unroll 1
  for (int p = 0; p < 7; p++) {
    double2 w[8];
    fwd_half_third(Y + ((long long)(blockIdx.x * 7 + p) * 2) * 32 * 49 + ky, 49, r, h, w);
    #pragma unroll
    for (int q = 0; q < 8; q++) park[q] = w[q];
    fwd_half_third(Y + ((long long)(blockIdx.x * 7 + p) * 2 + 1) * 32 * 49 + ky, 49, r, h, w);
    #pragma unroll
    for (int q = 0; q < 8; q++) {
      const double2 u = park[q]; double2 a = acc[q];
      a.x += u.x * w[q].x - u.y * w[q].y; a.y += u.x * w[q].y + u.y * w[q].x;
      acc[q] = a;
    }
  }
}
