// Micro-benchmark: cost of the throughput-mode RNG pieces per warp on sm_100a (clk per warp-op per SMSP).
#include <cstdio>
#include <cuda_runtime.h>
#include "../pta_replicator_b200/csrc/ptar_rng.cuh"
using namespace ptar;
template <int MODE>
__global__ void __launch_bounds__(256, 4) k(float* out, int iters, PhiloxKeys K) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  float acc = 0.f; uint32_t accu = 0;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {  // two independent Philox calls
      uint4 a = philox4x32_10(make_uint4(t, 1, it, 0), K), b = philox4x32_10(make_uint4(t, 2, it, 0), K);
      accu ^= a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w;
    } else if (MODE == 1) {  // 4 Box-Muller pairs from cheap words
      uint32_t w = t * 2654435761u + it * 40503u;
      float n0, n1, s = 0.f;
#pragma unroll
      for (int p = 0; p < 4; ++p) { box_muller(w + p * 0x9E3779B9u, (w ^ 0x5bd1e995u) + p * 77u, n0, n1); s += n0 + n1; }
      acc += s;
    } else if (MODE == 2) {  // full: 2 Philox + 4 BM
      float n[4], m[4];
      normals4(n, t, 1, 3, it, K); normals4(m, t, 2, 3, it, K);
      acc += n[0] + n[1] + n[2] + n[3] + m[0] + m[1] + m[2] + m[3];
    } else {  // full + f32->f64 + 8 DFMA (the white-noise part of the TOA stage)
      float n[4], m[4];
      normals4(n, t, 1, 3, it, K); normals4(m, t, 2, 3, it, K);
      double v = 0.0;
#pragma unroll
      for (int l = 0; l < 4; ++l) v = fma(1.1e-6, double(n[l]), fma(0.7e-6, double(m[l]), v));
      acc += float(v);
    }
  }
  out[t] = acc + float(accu);
}
template <int MODE> void run(const char* name, float* out) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
  const int iters = 2000, blocks = 148 * 4;
  k<MODE><<<blocks, 256>>>(out, 10, philox_keys(1)); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(out, iters, philox_keys(1)); cudaEventRecord(e1); cudaEventSynchronize(e1);
  cudaEventElapsedTime(&ms, e0, e1);
  // 8 warps/CTA x 4 CTAs/SM = 32 warps/SM = 8 per SMSP
  const double clk = ms * 1e-3 * 1.965e9;  // assuming 1965 MHz
  printf("%-34s %.3f ms  -> %.1f clk per warp-iteration per SMSP\n", name, ms, clk / (iters * 8.0));
}
int main() {
  float* out; cudaMalloc(&out, 148 * 4 * 256 * sizeof(float));
  run<0>("2 x Philox4x32-10", out); run<1>("4 x Box-Muller pair", out); run<2>("2 Philox + 4 BM (8 normals)", out);
  run<3>("8 normals + 8 cvt + 8 DFMA", out);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
