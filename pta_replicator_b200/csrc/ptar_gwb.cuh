// GWB stage kernels (red_noise.py:235-285) and the small dense Cholesky.
//
// The reference colours complex white draws with the Cholesky factor of the ORF, scales by
// sqrt(C(f)), Hermitian-packs, inverse-FFTs (length 2Nf-2 = 2 x prime) and keeps npts samples.
// All of that is ONE real linear map per pulsar row:  grid = A . z  with
//   parity mode      A = T  [npts x 2(Nf-2)]  (pruned inverse DFT x sqrt(C)/dt; z = the
//                    reference's own draws, so results match the reference to fp64 rounding)
//   throughput mode  A = L  [npts x npts] lower triangular with L L^T = T T^T (same Gaussian
//                    law from npts instead of 2(Nf-2) draws per pulsar; 10-20x fewer flops)
// and the ORF mixing M commutes with it, so it is applied to the draws first (ptar_gwb_mix).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ptar.h"
#include "ptar_rng.cuh"

namespace ptar {

// ---------------------------------------------------------------------------------------
// Zm[r][p][j] = sum_{q<=p} M[p][q] z[r][q][j]
// CTA: 32 columns j x 4 realizations (one Philox realization group).  Warp = 4 pulsar rows,
// lanes = columns; the draws of all pulsars for the 32 columns sit in shared memory.
constexpr int MIX_JT = 32;

template <bool INJECT>
__global__ void __launch_bounds__(1024) gwb_mix_kernel(double* __restrict__ Zm, const double* __restrict__ M,
                                                        const double* __restrict__ zin, int P, int J, int64_t nreal,
                                                        uint64_t seed, int64_t real0) {
  extern __shared__ __align__(16) double zs[];  // [P][32][4]
  const int j0 = blockIdx.x * MIX_JT;
  const int64_t rbase = int64_t(blockIdx.y) * 4;
  const int tid = threadIdx.x, nth = blockDim.x;
  for (int idx = tid; idx < P * MIX_JT; idx += nth) {
    const int q = idx / MIX_JT, jj = idx % MIX_JT, j = j0 + jj;
    double z[4] = {0, 0, 0, 0};
    if (j < J) {
      if (INJECT) {
#pragma unroll
        for (int l = 0; l < 4; ++l)
          if (rbase + l < nreal) z[l] = zin[((rbase + l) * P + q) * J + j];
      } else {
        float n[4];
        normals4(n, j, PTAR_K_GWB, q, static_cast<uint64_t>(real0 + rbase) >> 2, philox_keys(seed));
#pragma unroll
        for (int l = 0; l < 4; ++l) z[l] = static_cast<double>(n[l]);
      }
    }
    double2* d = reinterpret_cast<double2*>(zs + size_t(idx) * 4);
    d[0] = make_double2(z[0], z[1]);
    d[1] = make_double2(z[2], z[3]);
  }
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5, nwarps = nth >> 5;
  const int j = j0 + lane;
  for (int pb = warp; pb * 4 < P; pb += nwarps) {
    const int p0 = pb * 4;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int l = 0; l < 4; ++l) acc[a][l] = 0.0;
    const int qmax = min(P, p0 + 4);
    for (int q = 0; q < qmax; ++q) {
      const double2 za = *reinterpret_cast<const double2*>(zs + (size_t(q) * MIX_JT + lane) * 4);
      const double2 zb = *reinterpret_cast<const double2*>(zs + (size_t(q) * MIX_JT + lane) * 4 + 2);
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int p = p0 + a;
        const double m = (p < P && q <= p) ? __ldg(M + size_t(p) * P + q) : 0.0;
        acc[a][0] = fma(m, za.x, acc[a][0]);
        acc[a][1] = fma(m, za.y, acc[a][1]);
        acc[a][2] = fma(m, zb.x, acc[a][2]);
        acc[a][3] = fma(m, zb.y, acc[a][3]);
      }
    }
    if (j < J) {
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int p = p0 + a;
        if (p < P) {
#pragma unroll
          for (int l = 0; l < 4; ++l)
            if (rbase + l < nreal) Zm[((rbase + l) * P + p) * J + j] = acc[a][l];
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// ORF mixing on the fp64 tensor path.  Per CTA: 32 grid columns x 4 realizations = 128 rows
// (row = l*32 + jj), D[row][p] = sum_q z[q][row] * M[p][q]; K = q padded to a multiple of 4,
// N = p padded to a multiple of 8.  Draws z (Philox or injected) are staged in shared memory as
// zs[q][row] (row stride 132 -> the 8x4 A-fragment gather is conflict free), M as ms[p][q] (row
// stride KP+4).  Warp w owns rows 16w..16w+15 (2 m-tiles) and all n-tiles; k-steps beyond an
// n-tile's last pulsar are skipped (M is lower triangular).
constexpr int MX_ROWS = 128, MX_ZS = MX_ROWS + 4;

template <bool INJECT>
__global__ void __launch_bounds__(256) gwb_mix_dmma_kernel(double* __restrict__ Zm, const double* __restrict__ M,
                                                            const double* __restrict__ zin, int P, int J,
                                                            int64_t nreal, const PhiloxKeys K, int64_t real0) {
  extern __shared__ __align__(16) double mx_smem[];
  const int KP = (P + 3) & ~3, NP = (P + 7) & ~7, MS = KP + 4;
  double* zs = mx_smem;                      // [KP][132]
  double* ms = mx_smem + size_t(KP) * MX_ZS;  // [NP][MS]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j0 = blockIdx.x * MIX_JT;
  const int64_t rbase = int64_t(blockIdx.y) * 4;
  for (int idx = tid; idx < NP * MS; idx += 256) {
    const int p = idx / MS, q = idx % MS;
    ms[idx] = (p < P && q <= p) ? __ldg(M + size_t(p) * P + q) : 0.0;
  }
  for (int idx = tid; idx < KP * MIX_JT; idx += 256) {
    const int q = idx / MIX_JT, jj = idx % MIX_JT, j = j0 + jj;
    double z[4] = {0, 0, 0, 0};
    if (q < P && j < J) {
      if (INJECT) {
#pragma unroll
        for (int l = 0; l < 4; ++l)
          if (rbase + l < nreal) z[l] = zin[((rbase + l) * P + q) * J + j];
      } else {
        float n[4];
        normals4(n, j, PTAR_K_GWB, q, static_cast<uint64_t>(real0 + rbase) >> 2, K);
#pragma unroll
        for (int l = 0; l < 4; ++l) z[l] = static_cast<double>(n[l]);
      }
    }
#pragma unroll
    for (int l = 0; l < 4; ++l) zs[size_t(q) * MX_ZS + l * MIX_JT + jj] = z[l];
  }
  __syncthreads();
  const int fr = lane >> 2, fk = lane & 3;
  const int row0 = warp * 16;
  const int n_nt = NP / 8;
  for (int nt = 0; nt < n_nt; ++nt) {
    double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    const int kmax = min(KP, (nt * 8 + 8 + 3) & ~3);  // q <= p < nt*8+8
    for (int k0 = 0; k0 < kmax; k0 += 4) {
      const double b = ms[(nt * 8 + fr) * MS + k0 + fk];
      const double a0 = zs[size_t(k0 + fk) * MX_ZS + row0 + fr];
      const double a1 = zs[size_t(k0 + fk) * MX_ZS + row0 + 8 + fr];
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(acc[0][0]), "+d"(acc[0][1]) : "d"(a0), "d"(b));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(acc[1][0]), "+d"(acc[1][1]) : "d"(a1), "d"(b));
    }
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int row = row0 + t * 8 + fr;
      const int l = row >> 5, jj = row & 31, j = j0 + jj;
      if (j < J && rbase + l < nreal) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int p = nt * 8 + 2 * fk + u;
          if (p < P) Zm[((rbase + l) * P + p) * J + j] = acc[t][u];
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// G[c][n] = sum_j A[n][j] * Z[c][j]   (both operands K-contiguous; fp64 FMA pipe)
// 64x64 tile, BK = 16, 256 threads, 4x4 register micro-tile, register-prefetched next tile.
constexpr int SY_BM = 64, SY_BN = 64, SY_BK = 16, SY_PAD = 2;

__global__ void __launch_bounds__(256) gwb_synth_kernel(double* __restrict__ G, const double* __restrict__ A,
                                                         int64_t lda, const double* __restrict__ Z, int npts, int J,
                                                         int64_t ncols, int lower_tri) {
  __shared__ __align__(16) double As[SY_BK][SY_BM + SY_PAD];
  __shared__ __align__(16) double Bs[SY_BK][SY_BN + SY_PAD];
  const int tid = threadIdx.x;
  const int n0 = blockIdx.x * SY_BM;
  const int64_t c0 = int64_t(blockIdx.y) * SY_BN;
  const int lrow = tid >> 2, lq = (tid & 3) * 4;  // loader: row 0..63, k-quarter 0,4,8,12
  const int ty = tid >> 4, tx = tid & 15;         // compute: rows n0+ty*4.., cols c0+tx*4..
  const int kend = lower_tri ? min(J, n0 + SY_BM) : J;

  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[i][k] = 0.0;

  double ra[4], rb[4];
  auto fetch = [&](int k0) {
    const int n = n0 + lrow;
    const int64_t c = c0 + lrow;
    const int k = k0 + lq;
    if (n < npts && k < J) {
      const double2 u = *reinterpret_cast<const double2*>(A + size_t(n) * lda + k);
      const double2 v = *reinterpret_cast<const double2*>(A + size_t(n) * lda + k + 2);
      ra[0] = u.x; ra[1] = u.y; ra[2] = v.x; ra[3] = v.y;
    } else {
      ra[0] = ra[1] = ra[2] = ra[3] = 0.0;
    }
    if (c < ncols && k < J) {
      const double2 u = *reinterpret_cast<const double2*>(Z + size_t(c) * J + k);
      const double2 v = *reinterpret_cast<const double2*>(Z + size_t(c) * J + k + 2);
      rb[0] = u.x; rb[1] = u.y; rb[2] = v.x; rb[3] = v.y;
    } else {
      rb[0] = rb[1] = rb[2] = rb[3] = 0.0;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < kend; k0 += SY_BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[lq + i][lrow] = ra[i];
      Bs[lq + i][lrow] = rb[i];
    }
    __syncthreads();
    if (k0 + SY_BK < kend) fetch(k0 + SY_BK);
#pragma unroll
    for (int k = 0; k < SY_BK; ++k) {
      const double2 a01 = *reinterpret_cast<const double2*>(&As[k][ty * 4]);
      const double2 a23 = *reinterpret_cast<const double2*>(&As[k][ty * 4 + 2]);
      const double2 b01 = *reinterpret_cast<const double2*>(&Bs[k][tx * 4]);
      const double2 b23 = *reinterpret_cast<const double2*>(&Bs[k][tx * 4 + 2]);
      const double a[4] = {a01.x, a01.y, a23.x, a23.y};
      const double b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int m = 0; m < 4; ++m) acc[i][m] = fma(a[i], b[m], acc[i][m]);
    }
    __syncthreads();
  }
  // G[c][n]: thread owns n = n0+ty*4..+3 (contiguous) for 4 columns
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int64_t c = c0 + tx * 4 + m;
    if (c >= ncols) continue;
    const int n = n0 + ty * 4;
    double* g = G + size_t(c) * npts + n;
    if (n + 3 < npts && ((size_t(c) * npts + n) & 1) == 0) {
      *reinterpret_cast<double2*>(g) = make_double2(acc[0][m], acc[1][m]);
      *reinterpret_cast<double2*>(g + 2) = make_double2(acc[2][m], acc[3][m]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (n + i < npts) g[i] = acc[i][m];
    }
  }
}

// ---------------------------------------------------------------------------------------
// Same GEMM on the fp64 tensor path (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4; tcgen05 has no fp64
// kind).  D[c][n] = sum_j Z[c][j] * A[n][j]:  mma A-operand = Z (row-major, K contiguous),
// B-operand = A^T ("col": K contiguous per n).  CTA tile 128 (c) x 64 (n), BK = 16, 8 warps as
// 4 (c) x 2 (n), each a 32 x 32 warp tile = 4 x 4 DMMA tiles; operands staged with 16-byte
// cp.async through a 3-stage ring; rows padded to 20 doubles so the 8x4 fragment loads are
// conflict free.  The lower-triangular A of the throughput mode makes the n-tiles unequal
// (k runs to n0+64): the heaviest tiles are scheduled first.
constexpr int DM_BC = 128, DM_BN = 64, DM_BK = 16, DM_S = DM_BK + 4, DM_STAGES = 3;
constexpr size_t DM_SMEM = sizeof(double) * DM_STAGES * (DM_BC + DM_BN) * DM_S;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  const int bytes = valid ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(256, 2) gwb_synth_dmma_kernel(double* __restrict__ G, const double* __restrict__ A,
                                                                 int64_t lda, const double* __restrict__ Z, int npts,
                                                                 int J, int64_t ncols, int lower_tri, int c_tiles,
                                                                 int n_tiles) {
  extern __shared__ __align__(16) double dm_smem[];
  double* Zs = dm_smem;                                   // [STAGES][128][20]
  double* As = dm_smem + size_t(DM_STAGES) * DM_BC * DM_S;  // [STAGES][64][20]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nt = n_tiles - 1 - int(blockIdx.x / c_tiles);  // heavy (large n0) tiles first
  const int ct = blockIdx.x % c_tiles;
  const int n0 = nt * DM_BN;
  const int64_t c0 = int64_t(ct) * DM_BC;
  const int kend = lower_tri ? min(J, n0 + DM_BN) : J;
  const int nk = (kend + DM_BK - 1) / DM_BK;
  const int wc = (warp & 3) * 32, wn = (warp >> 2) * 32;
  const int fr = lane >> 2, fk = lane & 3;

  auto load_stage = [&](int kt, int st) {
    const int k0 = kt * DM_BK;
    // Z: 128 rows x 8 chunks ; A: 64 rows x 8 chunks (chunk = 2 doubles)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int ch = tid + i * 256;
      const int row = ch >> 3, kc = (ch & 7) * 2;
      const int64_t c = c0 + row;
      const bool ok = (c < ncols) && (k0 + kc < J);
      cp_async16(Zs + (size_t(st) * DM_BC + row) * DM_S + kc, Z + (ok ? size_t(c) * J + k0 + kc : 0), ok);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int ch = tid + i * 256;
      const int row = ch >> 3, kc = (ch & 7) * 2;
      const int n = n0 + row;
      const bool ok = (n < npts) && (k0 + kc < J);
      cp_async16(As + (size_t(st) * DM_BN + row) * DM_S + kc, A + (ok ? size_t(n) * lda + k0 + kc : 0), ok);
    }
  };

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int m = 0; m < 4; ++m) acc[i][m][0] = acc[i][m][1] = 0.0;

#pragma unroll
  for (int s = 0; s < DM_STAGES - 1; ++s) {
    if (s < nk) load_stage(s, s);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int kt = 0; kt < nk; ++kt) {
    asm volatile("cp.async.wait_group %0;" ::"n"(DM_STAGES - 2) : "memory");
    __syncthreads();
    {  // prefetch tile kt + STAGES - 1 into the slot freed in the previous iteration
      const int nxt = kt + DM_STAGES - 1;
      if (nxt < nk) load_stage(nxt, nxt % DM_STAGES);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const int st = kt % DM_STAGES;
    const double* zs = Zs + size_t(st) * DM_BC * DM_S;
    const double* as = As + size_t(st) * DM_BN * DM_S;
#pragma unroll
    for (int kk = 0; kk < DM_BK; kk += 4) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = zs[(wc + i * 8 + fr) * DM_S + kk + fk];
#pragma unroll
      for (int m = 0; m < 4; ++m) b[m] = as[(wn + m * 8 + fr) * DM_S + kk + fk];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int m = 0; m < 4; ++m)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(acc[i][m][0]), "+d"(acc[i][m][1])
                       : "d"(a[i]), "d"(b[m]));
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  const bool vec_ok = (npts & 1) == 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t c = c0 + wc + i * 8 + fr;
    if (c >= ncols) continue;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int n = n0 + wn + m * 8 + 2 * fk;
      double* g = G + size_t(c) * npts + n;
      if (vec_ok && n + 1 < npts) {
        *reinterpret_cast<double2*>(g) = make_double2(acc[i][m][0], acc[i][m][1]);
      } else {
        if (n < npts) g[0] = acc[i][m][0];
        if (n + 1 < npts) g[1] = acc[i][m][1];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Lower Cholesky, one CTA per matrix, left-looking; every inner product is a warp-shuffle
// reduction (north_star: "warp-shuffle reductions for the small dense Cholesky").
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256) cholesky_kernel(double* __restrict__ Lall, const double* __restrict__ Aall,
                                                        int n, int* __restrict__ info) {
  double* L = Lall + size_t(blockIdx.x) * n * n;
  const double* A = Aall + size_t(blockIdx.x) * n * n;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  __shared__ double s_diag;
  __shared__ int s_bad;
  if (tid == 0) s_bad = 0;
  for (int idx = tid; idx < n * n; idx += blockDim.x) {
    const int i = idx / n, j = idx % n;
    L[idx] = (j <= i) ? A[idx] : 0.0;
  }
  __syncthreads();
  for (int j = 0; j < n; ++j) {
    if (warp == 0) {
      double s = 0.0;
      for (int k = lane; k < j; k += 32) s = fma(L[size_t(j) * n + k], L[size_t(j) * n + k], s);
      s = warp_sum(s);
      if (lane == 0) {
        const double d = L[size_t(j) * n + j] - s;
        if (!(d > 0.0)) {
          s_bad = j + 1;
          s_diag = 1.0;
        } else {
          s_diag = sqrt(d);
        }
      }
    }
    __syncthreads();
    if (s_bad) break;
    const double ljj = s_diag;
    for (int i = j + 1 + warp; i < n; i += nwarps) {
      double s = 0.0;
      for (int k = lane; k < j; k += 32) s = fma(L[size_t(i) * n + k], L[size_t(j) * n + k], s);
      s = warp_sum(s);
      if (lane == 0) L[size_t(i) * n + j] = (L[size_t(i) * n + j] - s) / ljj;
    }
    if (tid == 0) L[size_t(j) * n + j] = ljj;
    __syncthreads();
  }
  if (tid == 0 && info) info[blockIdx.x] = s_bad;
}

}  // namespace ptar
