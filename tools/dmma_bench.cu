// Micro-benchmark: fp64 FMA pipe vs fp64 tensor (mma.sync.m8n8k4.f64) throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dfma(double* out, int iters) {
  double a[8], x = threadIdx.x * 1e-3, y = 1.0000001;
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], y, x);
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmma(double* out, int iters) {
  double c[8][2];
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0;
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double* out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(double));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters = 20000; float ms;
  for (int th : {256, 512, 1024}) {
    k_dfma<<<148 * 2, th>>>(out, 1000); cudaDeviceSynchronize();
    cudaEventRecord(e0); k_dfma<<<148 * 2, th>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("DFMA  threads=%4d: %.2f TFLOP/s\n", th, 2.0 * 148 * 2 * th * 8.0 * iters / (ms * 1e-3) / 1e12);
    k_dmma<<<148 * 2, th>>>(out, 1000); cudaDeviceSynchronize();
    cudaEventRecord(e0); k_dmma<<<148 * 2, th>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("DMMA  threads=%4d: %.2f TFLOP/s\n", th, 2.0 * 148 * 2 * (th / 32) * 8.0 * 256 * iters / (ms * 1e-3) / 1e12);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
