// Micro-benchmark: HBM write bandwidth of the generator's store pattern (no compute).
// grid (RCHUNKS, TILES); CTA = 1024 TOAs x 16 rows; thread = 1 TOA x 4 rows per step.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, size_t ld, int ntoa_tile) {
  const int tile = blockIdx.y, r0 = blockIdx.x * 16;
  for (int tt = threadIdx.x; tt < ntoa_tile; tt += 256) {
    double* orow = out + size_t(r0) * ld + size_t(tile) * ntoa_tile + tt;
    for (int rg = 0; rg < 4; ++rg, orow += 4 * ld) {
      const double v = tt * 1e-9 + rg;
      if (MODE == 0) { __stcs(orow, v); __stcs(orow + ld, v + 1); __stcs(orow + 2 * ld, v + 2); __stcs(orow + 3 * ld, v + 3); }
      else { orow[0] = v; orow[ld] = v + 1; orow[2 * ld] = v + 2; orow[3 * ld] = v + 3; }
    }
  }
}
__global__ void fill(double* out, size_t n) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) out[i] = i * 1e-9;
}
int main() {
  const int tiles = 667, ntoa = 960, R = 256;
  const size_t ld = size_t(tiles) * ntoa;
  double* out; cudaMalloc(&out, ld * R * sizeof(double));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); k<0><<<dim3(R / 16, tiles), 256>>>(out, ld, ntoa); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); printf("pattern __stcs : %.3f ms  %.1f GB/s\n", ms, ld * R * 8.0 / ms / 1e6);
    cudaEventRecord(e0); k<1><<<dim3(R / 16, tiles), 256>>>(out, ld, ntoa); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); printf("pattern plain  : %.3f ms  %.1f GB/s\n", ms, ld * R * 8.0 / ms / 1e6);
    cudaEventRecord(e0); fill<<<148 * 8, 1024>>>(out, ld * R); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); printf("linear fill    : %.3f ms  %.1f GB/s\n", ms, ld * R * 8.0 / ms / 1e6);
    cudaEventRecord(e0); cudaMemsetAsync(out, 0, ld * R * 8); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); printf("cudaMemset     : %.3f ms  %.1f GB/s\n", ms, ld * R * 8.0 / ms / 1e6);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
