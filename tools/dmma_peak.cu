// FP64 tensor-core (mma.sync m8n8k4 f64) vs FP64 FMA issue rate on this GPU: decides whether the dense cross-moment
// kernel should move to DMMA.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_peak tools/dmma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k_dmma(int iters, double seed, double* sink) {
  double c[8][2];
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = seed * (threadIdx.x + i);
  double a = seed + threadIdx.x * 1e-9, b = seed - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) *sink = s;
}

__global__ void __launch_bounds__(256) k_dfma(int iters, double seed, double* sink) {
  double c[16];
  for (int i = 0; i < 16; ++i) c[i] = seed * (threadIdx.x + i);
  double a = 1.0 + seed * 1e-9, b = seed * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
  for (int i = 0; i < 16; ++i) s += c[i];
  if (s == 12345.678) *sink = s;
}

int main() {
  double* sink;
  cudaMalloc(&sink, 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20000, grid = 148 * 8;
  float ms;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    k_dmma<<<grid, 256>>>(iters, 1.000001, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl = (double)grid * 8 * iters * 8 * 512.0;  // warps * iters * mma per iter * flops per mma
    printf("DMMA m8n8k4: %.2f TFLOP/s (%.2f ms)\n", fl / ms * 1e-9, ms);
    cudaEventRecord(e0);
    k_dfma<<<grid, 256>>>(iters, 1.000001, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl2 = (double)grid * 256 * iters * 16 * 2.0;
    printf("DFMA       : %.2f TFLOP/s (%.2f ms)\n", fl2 / ms * 1e-9, ms);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
