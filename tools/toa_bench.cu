// Micro-benchmark: the generator's TOA stage in isolation, pieces added one at a time (clk per warp-iteration
// per SM sub-partition, 8 resident warps).  MODE 0: RNG+cvt+white FMA; 1: + Cs loads + Horner; 2: + streaming
// stores; 3: + per-TOA statics from global.
#include <cstdio>
#include <cuda_runtime.h>
#include "../pta_replicator_b200/csrc/ptar_rng.cuh"
using namespace ptar;
constexpr int CSS = 50;
template <int MODE>
__global__ void __launch_bounds__(256, 4) k(double* out, size_t ld, const double* w1g, const double* w2g, const double* dtg,
                                            const unsigned short* elg, int nrchunk, PhiloxKeys K) {
  extern __shared__ double Cs[];
  for (int i = threadIdx.x; i < 64 * CSS; i += 256) Cs[i] = 1e-7 * i;
  __syncthreads();
  const int tile = blockIdx.y, r0 = blockIdx.x * 16;
  double sink = 0.0;
  for (int tt = threadIdx.x; tt < 960; tt += 256) {
    const size_t gi = size_t(tile) * 960 + tt;
    double w1 = 1.1e-6, w2 = 0.7e-6, dt = 0.3;
    int el = (tt >> 5) & 31;
    if (MODE >= 3) { w1 = w1g[gi]; w2 = w2g[gi]; dt = dtg[gi]; el = elg[gi]; }
    double* orow = out + size_t(r0) * ld + gi;
    for (int rg = 0; rg < 4; ++rg, orow += 4 * ld) {
      float n1[4], n2[4];
      normals4(n1, (unsigned)gi, 1, 3, blockIdx.x * 4 + rg, K);
      normals4(n2, (unsigned)gi, 2, 3, blockIdx.x * 4 + rg, K);
      double v[4] = {0, 0, 0, 0};
      if (MODE >= 1) {
        const double2* c2 = reinterpret_cast<const double2*>(Cs + el * CSS + rg * 12);
        const double2 q0 = c2[0], q1 = c2[1], q2 = c2[2], q3 = c2[3], q4 = c2[4], q5 = c2[5];
        v[0] = fma(dt, fma(dt, q1.x, q0.y), q0.x); v[1] = fma(dt, fma(dt, q2.y, q2.x), q1.y);
        v[2] = fma(dt, fma(dt, q4.x, q3.y), q3.x); v[3] = fma(dt, fma(dt, q5.y, q5.x), q4.y);
      }
#pragma unroll
      for (int l = 0; l < 4; ++l) v[l] = fma(w1, (double)n1[l], fma(w2, (double)n2[l], v[l]));
      if (MODE >= 2) { __stcs(orow, v[0]); __stcs(orow + ld, v[1]); __stcs(orow + 2 * ld, v[2]); __stcs(orow + 3 * ld, v[3]); }
      else sink += v[0] + v[1] + v[2] + v[3];
    }
  }
  if (MODE < 2) out[size_t(blockIdx.y) * gridDim.x * 256 + blockIdx.x * 256 + threadIdx.x] = sink;
}
template <int MODE> void run(const char* name, double* out, size_t ld, double* w1, double* w2, double* dt, unsigned short* el) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
  const int tiles = 667, R = 256;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 53776);
  k<MODE><<<dim3(R / 16, tiles), 256, 53776>>>(out, ld, w1, w2, dt, el, R / 16, philox_keys(1)); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<dim3(R / 16, tiles), 256, 53776>>>(out, ld, w1, w2, dt, el, R / 16, philox_keys(1)); cudaEventRecord(e1);
  cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
  const double iters_per_smsp = double(tiles) * 960 * R / 4 / 32 / (148 * 4);
  printf("%-44s %.3f ms -> %.0f clk per warp-iteration per SMSP\n", name, ms, ms * 1e-3 * 1.965e9 / iters_per_smsp);
}
int main() {
  const size_t ld = 667 * 960; const int R = 256;
  double *out, *w1, *w2, *dt; unsigned short* el;
  cudaMalloc(&out, ld * R * 8); cudaMalloc(&w1, ld * 8); cudaMalloc(&w2, ld * 8); cudaMalloc(&dt, ld * 8); cudaMalloc(&el, ld * 2);
  cudaMemset(w1, 0, ld * 8); cudaMemset(w2, 0, ld * 8); cudaMemset(dt, 0, ld * 8); cudaMemset(el, 0, ld * 2);
  run<0>("RNG + cvt + white FMA", out, ld, w1, w2, dt, el);
  run<1>("+ 6 LDS.128 + Horner", out, ld, w1, w2, dt, el);
  run<2>("+ 4 streaming stores", out, ld, w1, w2, dt, el);
  run<3>("+ per-TOA statics (whole TOA stage)", out, ld, w1, w2, dt, el);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
