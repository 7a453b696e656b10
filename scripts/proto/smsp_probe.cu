// Hardware probe for the redesign of the ComputeQ y/x-stage kernel (not part of the library).
//   (1) which SM sub-partition (warp slot % 4) the warps of two co-resident 192-thread CTAs land on;
//   (2) how fast an SM runs the 16-point "third" (fc3::fwd_third<16>) as a function of warps per CTA and CTAs per SM:
//       a) from registers only, b) with the y-stage's shared-memory traffic (32 LDS.128 in, 16 STS.128 out per third),
//       c) with the x stage's (64 LDS.128 in per pair of thirds, products kept in registers).
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../landau-poisson-solver_b200/csrc smsp_probe.cu -o smsp_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "fc3.cuh"
using namespace fc3;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void k_map(int *out, long long spin)
{
  extern __shared__ double2 sm[];
  unsigned smid, warpid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
  const long long t0 = clock64();
  while (clock64() - t0 < spin) { }
  if ((threadIdx.x & 31) == 0) {
    int *o = out + 4 * (blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32);
    o[0] = smid; o[1] = warpid; o[2] = blockIdx.x; o[3] = threadIdx.x / 32;
  }
  if (spin < 0) sm[threadIdx.x] = make_double2(0, 0);
}

// MODE 0: registers only; 1: y-stage-like (smem in, smem out); 2: x-stage-like (two thirds from smem + products)
template <int MODE, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_thirds(double2 *out, int iters)
{
  constexpr int L = 16, N = 32, PY = 49;
  extern __shared__ double2 sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r = warp % 3;
  // every warp owns a private [32][49]-ish region it keeps re-reading/re-writing (no barriers: throughput only)
  double2 *mine = sm + warp * (N * 33);
  for (int i = lane; i < N * 33; i += 32) mine[i] = make_double2(1e-3 * (i % 7), 1e-3 * (i % 5));
  __syncwarp();
  double2 acc[L];
  #pragma unroll
  for (int q = 0; q < L; q++) acc[q] = make_double2(0., 0.);
  double2 a0[L], a1[L], y[L];
  #pragma unroll
  for (int l = 0; l < L; l++) { a0[l] = make_double2(1e-3 * (l + lane), 1e-3); a1[l] = make_double2(1e-3, 1e-3 * (l - lane)); }
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) {
      fwd_third<L>(a0, a1, r, y);
      #pragma unroll
      for (int l = 0; l < L; l++) { a0[l] = make_double2(y[l].x * 0.25, y[l].y * 0.25); a1[l] = make_double2(y[(l + 3) % L].y * 0.25, y[(l + 5) % L].x * 0.25); }
    } else if (MODE == 1) {
      #pragma unroll
      for (int l = 0; l < L; l++) { a0[l] = mine[l * N + lane]; a1[l] = mine[(l + L) * N + lane]; }
      fwd_third<L>(a0, a1, r, y);
      #pragma unroll
      for (int q = 0; q < L; q++) mine[q * N + lane] = make_double2(y[q].x * 0.25, y[q].y * 0.25);
    } else if (MODE == 3 || (MODE == 5 && warp >= 4)) {
      // y-like with every load live: the two halves of the region swap roles every iteration
      const int o = (it & 1) * 16 * N;
      #pragma unroll
      for (int l = 0; l < L; l++) { a0[l] = mine[(o + l * N + lane) % (32 * N)]; a1[l] = mine[(o + (l + L) * N + lane) % (32 * N)]; }
      fwd_third<L>(a0, a1, r, y);
      #pragma unroll
      for (int q = 0; q < L; q++) mine[((o ^ (16 * N)) + q * N + lane) % (32 * N)] = make_double2(y[q].x * 0.25, y[q].y * 0.25);
    } else if (MODE == 4 || MODE == 5) {
      // x-like: two thirds from 64 live loads, products folded into one complex accumulator (no register pressure from acc)
      double2 u[L];
      const int o = (it & 1) * N;
      #pragma unroll
      for (int l = 0; l < L; l++) { a0[l] = mine[o + l * N + lane]; a1[l] = mine[o + (l + L - 1) * N + lane]; }
      fwd_third<L>(a0, a1, r, u);
      #pragma unroll
      for (int l = 0; l < L; l++) { a0[l] = mine[N - o + l * N + lane]; a1[l] = mine[N - o + (l + L - 1) * N + lane]; }
      fwd_third<L>(a0, a1, r, y);
      #pragma unroll
      for (int q = 0; q < L; q++) {
        acc[0].x += u[q].x * y[q].x - u[q].y * y[q].y;
        acc[0].y += u[q].x * y[q].y + u[q].y * y[q].x;
      }
      mine[(it & 15) * N + lane] = acc[0];
    } else {
      double2 u[L];
      #pragma unroll
      for (int l = 0; l < L; l++) { a0[l] = mine[l * N + lane]; a1[l] = mine[(l + L) * N + lane]; }
      fwd_third<L>(a0, a1, r, u);
      #pragma unroll
      for (int l = 0; l < L; l++) { a0[l] = mine[(l + 1) * N + lane]; a1[l] = mine[((l + L + 1) % N) * N + lane]; }
      fwd_third<L>(a0, a1, r, y);
      #pragma unroll
      for (int q = 0; q < L; q++) {
        acc[q].x += u[q].x * y[q].x - u[q].y * y[q].y;
        acc[q].y += u[q].x * y[q].y + u[q].y * y[q].x;
      }
    }
  }
  double2 s = make_double2(0., 0.);
  #pragma unroll
  for (int q = 0; q < L; q++) { s.x += acc[q].x + a0[q].x + y[q].x; s.y += acc[q].y + a1[q].y + y[q].y; }
  if (s.x == 1.2345) out[0] = s;
}

template <int MODE, int NT, int MINB>
void run(const char *name, int iters, double2 *d, int nsm)
{
  const size_t smem_warps = (size_t)(NT / 32) * 32 * 33 * sizeof(double2);
  // force MINB CTAs per SM through the shared-memory footprint too
  size_t smem = smem_warps;
  const size_t cap = (size_t)227 * 1024 / MINB - 1024;
  if (smem > cap) { printf("%-40s skipped (smem)\n", name); return; }
  if (MINB == 1 && smem < 120 * 1024) smem = 120 * 1024;
  if (MINB == 2 && smem < 80 * 1024) smem = 80 * 1024;
  auto kern = k_thirds<MODE, NT, MINB>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaFuncAttributes fa;
  CK(cudaFuncGetAttributes(&fa, kern));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem));
  const int grid = nsm * occ * 4;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  kern<<<grid, NT, smem>>>(d, 8);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaEventRecord(e0));
    kern<<<grid, NT, smem>>>(d, iters);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  const double thirds = (double)grid * NT * iters * (MODE == 2 || MODE == 4 ? 2. : MODE == 5 ? 1.5 : 1.);
  // textbook flop of a third: a 48-point line is 5 M log2 M = 1340 flop -> 447 per third
  printf("%-40s regs %3d occ %d warps/SM %2d : %8.3f ms  %7.2f Gthirds/s  %6.2f TFLOP/s(textbook)\n", name, fa.numRegs, occ, occ * NT / 32, best,
         thirds / best * 1e-6, thirds * 446.8 / best * 1e-9);
}

int main()
{
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int nsm = prop.multiProcessorCount;
  printf("%s, %d SMs\n", prop.name, nsm);
  // ---- (1) warp-slot map
  for (int nt : {192, 256, 384}) {
    const int nb = 2 * nsm, nw = nt / 32;
    int *d; CK(cudaMalloc(&d, sizeof(int) * 4 * nb * nw));
    CK(cudaFuncSetAttribute(k_map, cudaFuncAttributeMaxDynamicSharedMemorySize, 84 * 1024));
    k_map<<<nb, nt, 84 * 1024>>>(d, 2000000);
    CK(cudaDeviceSynchronize());
    std::vector<int> h(4 * nb * nw);
    CK(cudaMemcpy(h.data(), d, h.size() * sizeof(int), cudaMemcpyDeviceToHost));
    for (int sm : {0, 1, 77}) {
      printf("threads/CTA %d, SM %d: (block, warp)->slot%%4 :", nt, sm);
      int per[4] = {0, 0, 0, 0};
      for (int i = 0; i < nb * nw; i++)
        if (h[4 * i] == sm) { printf(" (%d,%d)->%d[%d]", h[4 * i + 2], h[4 * i + 3], h[4 * i + 1] % 4, h[4 * i + 1]); per[h[4 * i + 1] % 4]++; }
      printf("  | per sub-partition %d %d %d %d\n", per[0], per[1], per[2], per[3]);
    }
    CK(cudaFree(d));
  }
  // ---- (2) thirds throughput
  double2 *d; CK(cudaMalloc(&d, 64));
  const int it = 2000;
  run<0, 128, 1>("regs  128x1", it, d, nsm);
  run<0, 256, 1>("regs  256x1", it, d, nsm);
  run<0, 192, 2>("regs  192x2", it, d, nsm);
  run<0, 384, 1>("regs  384x1", it, d, nsm);
  run<0, 256, 2>("regs  256x2 (<=128 regs)", it, d, nsm);
  run<0, 512, 1>("regs  512x1 (<=128 regs)", it, d, nsm);
  run<3, 128, 1>("ylive 128x1", it, d, nsm);
  run<3, 256, 1>("ylive 256x1", it, d, nsm);
  run<3, 384, 1>("ylive 384x1", it, d, nsm);
  run<4, 128, 1>("xlive 128x1", it / 2, d, nsm);
  run<4, 256, 1>("xlive 256x1", it / 2, d, nsm);
  run<4, 384, 1>("xlive 384x1", it / 2, d, nsm);
  run<5, 256, 1>("mixed 256x1 (warps 0-3 x, 4-7 y)", it / 2, d, nsm);
  run<1, 256, 1>("ylike 256x1", it, d, nsm);
  run<1, 192, 2>("ylike 192x2", it, d, nsm);
  run<1, 384, 1>("ylike 384x1", it, d, nsm);
  run<1, 256, 2>("ylike 256x2 (<=128 regs)", it, d, nsm);
  run<2, 256, 1>("xlike 256x1", it / 2, d, nsm);
  run<2, 192, 2>("xlike 192x2", it / 2, d, nsm);
  run<2, 384, 1>("xlike 384x1", it / 2, d, nsm);
  run<2, 256, 2>("xlike 256x2 (<=128 regs)", it / 2, d, nsm);
  return 0;
}
