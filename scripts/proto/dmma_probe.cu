// FP64 tensor-core (mma.sync m8n8k4 f64) throughput on this GPU next to the DFMA rate: is DMMA worth using for the dense
// contractions of the Fourier -> DG projection?  (not part of the library)
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a dmma_probe.cu -o dmma_probe
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
template <int NACC>
__global__ void __launch_bounds__(256) k_dmma(double *out, int iters, double a0, double b0)
{
  double c[NACC][2];
  for (int i = 0; i < NACC; i++) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = threadIdx.x * 2e-3 - i; }
  double a = a0 + threadIdx.x * 1e-9, b = b0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; it++) {
    #pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0.;
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  if (s == 1.2345) out[0] = s;
}
__global__ void __launch_bounds__(256) k_dfma(double *out, int iters, double a, double b)
{
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  if (x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 == 1.2345) out[0] = x0;
}
int main()
{
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  double *d; CK(cudaMalloc(&d, 8));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int blocks = p.multiProcessorCount * 8, iters = 1 << 14;
  for (int rep = 0; rep < 3; rep++) {
    float ms;
    CK(cudaEventRecord(e0)); k_dfma<<<blocks, 256>>>(d, iters, 0.999999, 1e-7); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    const double dfma = 2.0 * 8 * iters * 256.0 * blocks / (ms * 1e-3) / 1e12;
    CK(cudaEventRecord(e0)); k_dmma<8><<<blocks, 256>>>(d, iters, 0.999999, 1e-7); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    const double dmma8 = 512.0 * 8 * iters * 8.0 * blocks / (ms * 1e-3) / 1e12;
    CK(cudaEventRecord(e0)); k_dmma<2><<<blocks, 256>>>(d, iters, 0.999999, 1e-7); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    const double dmma2 = 512.0 * 2 * iters * 8.0 * blocks / (ms * 1e-3) / 1e12;
    printf("%s: DFMA %.2f TFLOP/s | DMMA m8n8k4, 8 independent accumulators per warp %.2f TFLOP/s, 2 accumulators %.2f TFLOP/s\n", p.name, dfma, dmma8, dmma2);
  }
  return 0;
}
