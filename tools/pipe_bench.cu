// Micro-benchmark: do the pipes the TOA stage leans on overlap?  Measures clk per warp-instruction per SMSP for
// pure streams (IMAD.WIDE, IMAD.HI, IMAD, MUFU, F2F.F64.F32, DFMA, LOP3, FMUL) and for 1:1 interleaved pairs with
// 8 warps per SMSP and 4 independent chains per thread.  sum-like pair times = shared issue/dispatch resource,
// max-like = independent pipes.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_bench pipe_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum Op { WIDE, MHI, MLO, MUFU, CVT, DFMA, LOP, FMUL, NONE, DMMA };

template <int OP>
__device__ __forceinline__ void op(uint32_t& a, uint32_t& b, float& f, double& d) {
  if (OP == WIDE) {
    asm volatile("{.reg .b64 t; mov.b64 t, {%0, %1}; mad.wide.u32 t, %0, 0xD2511F53, t; mov.b64 {%0, %1}, t;}" : "+r"(a), "+r"(b));
  } else if (OP == MHI) {
    asm volatile("mul.hi.u32 %0, %0, 0xD2511F53;" : "+r"(a));
  } else if (OP == MLO) {
    asm volatile("mad.lo.u32 %0, %0, 0xD2511F53, %1;" : "+r"(a) : "r"(b));
  } else if (OP == MUFU) {
    asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f));
  } else if (OP == CVT) {
    // the converted value feeds an integer accumulator and the input changes every time (2 LOP3 ride along)
    asm volatile("{.reg .f64 t; .reg .b32 lo, hi, x; cvt.f64.f32 t, %0; mov.b64 {lo, hi}, t; xor.b32 %1, %1, hi;"
                 " mov.b32 x, %0; xor.b32 x, x, 0x1; mov.b32 %0, x;}" : "+f"(f), "+r"(a));
  } else if (OP == DFMA) {
    asm volatile("fma.rn.f64 %0, %0, %0, %0;" : "+d"(d));
  } else if (OP == LOP) {
    asm volatile("add.u32 %0, %0, %1; xor.b32 %1, %1, %0;" : "+r"(b), "+r"(a));  // IADD3 + LOP3 (both alu pipe)
  } else if (OP == FMUL) {
    asm volatile("mul.ftz.f32 %0, %0, %0;" : "+f"(f));
  } else if (OP == DMMA) {  // D(8x8) = A(8x4) B(4x8) + C on the FP64 tensor path: 256 FMA per warp instruction
    asm volatile("{.reg .f64 c1; mov.f64 c1, %0; mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, c1}, {%1}, {%2}, {%0, c1};}"
                 : "+d"(d) : "d"(1.0e-3), "d"(1.0e-3));
  }
}

template <int A, int B>
__global__ void __launch_bounds__(256, 4) k(float* out, int iters) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t a[4], b[4], a2[4], b2[4];
  float f[4], g[4];
  double d[4], e[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    a[c] = t + c; b[c] = t * 3 + c; a2[c] = t * 5 + c; b2[c] = t * 7 + c; f[c] = 1.5f + t + c; g[c] = 2.5f + t + c; d[c] = 1.0 + 1e-9 * t; e[c] = 1.0 + 2e-9 * t;
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        op<A>(a[c], b[c], f[c], d[c]);
        if (B != NONE) op<B>(a2[c], b2[c], g[c], e[c]);  // second stream on its own registers
      }
    }
  }
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) acc += float(a[c] ^ b[c] ^ a2[c] ^ b2[c]) + f[c] + g[c] + float(d[c]) + float(e[c]);
  out[t] = acc;
}

template <int A, int B>
double run(const char* name, float* out) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  const int iters = 4000, blocks = 148 * 4;
  k<A, B><<<blocks, 256>>>(out, 10);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<A, B><<<blocks, 256>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  cudaEventElapsedTime(&ms, e0, e1);
  const double clk = ms * 1e-3 * 1.965e9;                     // assuming 1965 MHz
  const double per = clk / (double(iters) * 32.0 * 8.0);      // 32 ops of each stream per iteration, 8 warps per SMSP
  printf("%-28s %8.3f ms  %6.2f clk per warp-op(pair) per SMSP\n", name, ms, per);
  return per;
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 4 * 256 * sizeof(float));
  run<WIDE, NONE>("IMAD.WIDE", out);
  run<MHI, NONE>("IMAD.HI", out);
  run<MLO, NONE>("IMAD (lo)", out);
  run<MUFU, NONE>("MUFU.LG2", out);
  run<CVT, NONE>("F2F.F64.F32", out);
  run<DFMA, NONE>("DFMA", out);
  run<LOP, NONE>("IADD3+LOP3", out);
  run<FMUL, NONE>("FMUL", out);
  run<WIDE, MUFU>("IMAD.WIDE + MUFU", out);
  run<WIDE, CVT>("IMAD.WIDE + F2F", out);
  run<WIDE, DFMA>("IMAD.WIDE + DFMA", out);
  run<WIDE, LOP>("IMAD.WIDE + IADD3+LOP3", out);
  run<WIDE, FMUL>("IMAD.WIDE + FMUL", out);
  run<MUFU, CVT>("MUFU + F2F", out);
  run<MUFU, DFMA>("MUFU + DFMA", out);
  run<MUFU, FMUL>("MUFU + FMUL", out);
  run<MUFU, LOP>("MUFU + IADD3+LOP3", out);
  run<CVT, DFMA>("F2F + DFMA", out);
  run<MHI, MLO>("IMAD.HI + IMAD(lo)", out);
  run<MHI, MUFU>("IMAD.HI + MUFU", out);
  run<DFMA, LOP>("DFMA + IADD3+LOP3", out);
  run<DMMA, NONE>("DMMA.884", out);
  run<DMMA, WIDE>("DMMA + IMAD.WIDE", out);
  run<DMMA, MUFU>("DMMA + MUFU", out);
  run<DMMA, LOP>("DMMA + IADD3+LOP3", out);
  run<DMMA, FMUL>("DMMA + FMUL", out);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
