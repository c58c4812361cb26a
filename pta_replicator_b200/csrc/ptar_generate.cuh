// Fused residual generator for sm_100a.
//
//   out[r][i] = w1_i z1 + w2_i z2                      white_noise.py:105-109
//             + ecorr_e z_{bucket(e)}                  white_noise.py:182
//             + sum_j F(t_i)_j sqrt(prior_j) z_j       red_noise.py:126-128
//             + lerp(G[r][psr], t_i)                   red_noise.py:286-287
//             + det_i                                  deterministic.py:160-165
//
// One CTA = one tile (<=1024 time-sorted TOAs of one pulsar, <=64 kernel-epochs) x RC
// realizations.  Work that is constant inside an epoch is done once per (epoch, realization)
// in an "epoch stage" and kept in shared memory:
//   * the Fourier projection F_e . a_r -- a [64 x J] x [J x 3RC] fp64 GEMM on the FMA pipe
//     whose three right-hand sides are the coefficient vector and its first two time
//     derivatives, so that TOAs inside an epoch (sub-band TOAs <1 s apart) are reached by a
//     2nd-order Taylor step whose remainder is below fp64 rounding of the direct sum (the
//     host picks nd per tile from |omega_max * dt|; single-TOA epochs are exact, nd = 1);
//   * the ECORR draw of the epoch's bucket.
// The basis tile F [J][64] is fetched with one bulk-async (TMA) copy that overlaps the Philox
// generation of the coefficients.  The "TOA stage" then streams the realizations: per TOA two
// Philox/Box-Muller normals, a 3-term Horner step, a 2-point GWB interpolation read through
// L1, and one 32-byte store -- the only HBM traffic that scales with R x N_toa.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ptar.h"
#include "ptar_rng.cuh"

namespace ptar {

constexpr int GEN_THREADS = 256;
constexpr int EP = PTAR_TILE_EPOCHS;  // 64

__host__ __device__ inline size_t gen_smem_bytes(int J, int RC) {
  // mbarrier (16 B) + Fs[J][64] + As[3][J][RC] + Cs[64][3RC+1]
  return 16 + sizeof(double) * (size_t(J) * EP + size_t(3) * J * RC + size_t(EP) * (3 * RC + 1));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

template <int RC, bool INJECT>
__global__ void __launch_bounds__(GEN_THREADS, (RC <= 16 ? 2 : 1)) gen_kernel(const ptar_gen_params P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw);
  double* Fs = reinterpret_cast<double*>(smem_raw + 16);
  const int J = P.J;
  double* As = Fs + size_t(J) * EP;
  double* Cs = As + size_t(3) * J * RC;
  constexpr int CSS = 3 * RC + 1;  // odd stride: epochs land in different banks

  const int tid = threadIdx.x;
  const ptar_tile tile = P.tiles[blockIdx.x];
  const int r0 = blockIdx.y * RC;                 // first realization (local to this call)
  const int nr = min(RC, P.nreal - r0);
  const uint32_t flags = P.flags;
  const bool has_red = (flags & PTAR_F_RED) && J > 0;
  const bool has_ecorr = (flags & PTAR_F_ECORR) != 0;
  const bool has_epoch = has_red || has_ecorr;
  const int nd = has_red ? tile.nd : 1;
  const uint32_t psr = static_cast<uint32_t>(tile.psr);
  const uint64_t rgroup0 = static_cast<uint64_t>(P.real0 + r0) >> 2;  // realization-lane streams

  // ---- epoch stage -------------------------------------------------------------------
  if (has_red) {
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      const uint32_t bytes = static_cast<uint32_t>(sizeof(double) * J * EP);
      const double* src = P.Ftile + size_t(blockIdx.x) * J * EP;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(bytes)
                   : "memory");
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
              smem_u32(Fs)),
          "l"(src), "r"(bytes), "r"(smem_u32(mbar))
          : "memory");
    }
    // coefficients a = sqrt(prior) * z for RC realizations (overlaps the bulk copy)
    const double* scale = P.rn_scale + size_t(psr) * J;
    if (INJECT) {
      for (int idx = tid; idx < J * RC; idx += GEN_THREADS) {
        const int j = idx / RC, r = idx % RC;
        As[idx] = (r < nr) ? scale[j] * P.zrn[(size_t(r0 + r) * P.n_psr + psr) * J + j] : 0.0;
      }
    } else {
      constexpr int RG = RC / 4;
      for (int idx = tid; idx < J * RG; idx += GEN_THREADS) {
        const int j = idx / RG, rg = idx % RG;
        float n[4];
        normals4(n, j, PTAR_K_RED, psr, rgroup0 + rg, P.seed);
        const double s = scale[j];
#pragma unroll
        for (int l = 0; l < 4; ++l) As[j * RC + rg * 4 + l] = s * static_cast<double>(n[l]);
      }
    }
    __syncthreads();
    if (nd > 1) {  // time derivatives of the coefficient vector (folded 1/2 in the 2nd)
      const double* om = P.rn_omega + size_t(psr) * (J / 2);
      const double sgn_even = P.rn_convention ? 1.0 : -1.0;
      for (int idx = tid; idx < J * RC; idx += GEN_THREADS) {
        const int j = idx / RC, r = idx % RC;
        const double w = om[j >> 1];
        const double partner = As[(j ^ 1) * RC + r];
        As[J * RC + idx] = ((j & 1) ? -sgn_even : sgn_even) * w * partner;
        As[2 * J * RC + idx] = -0.5 * w * w * As[idx];
      }
    }
  }
  if (has_epoch) {
    // Cs[e][r][0] starts from the ECORR draw of the epoch's bucket (or 0)
    if (has_ecorr) {
      if (INJECT) {
        for (int idx = tid; idx < EP * RC; idx += GEN_THREADS) {
          const int e = idx / RC, r = idx % RC;
          double v = 0.0;
          if (e < tile.n_ep && r < nr) {
            const int ge = tile.ep_start + e;
            v = P.ep_ecorr[ge] * P.zb[size_t(r0 + r) * P.n_bucket_total + P.psr_bucket_off[psr] + P.ep_bucket[ge]];
          }
          Cs[e * CSS + r * 3] = v;
        }
      } else {
        constexpr int RG = RC / 4;
        for (int idx = tid; idx < EP * RG; idx += GEN_THREADS) {
          const int e = idx / RG, rg = idx % RG;
          float n[4] = {0.f, 0.f, 0.f, 0.f};
          double ec = 0.0;
          if (e < tile.n_ep) {
            const int ge = tile.ep_start + e;
            ec = P.ep_ecorr[ge];
            normals4(n, P.ep_bucket[ge], PTAR_K_ECORR, psr, rgroup0 + rg, P.seed);
          }
#pragma unroll
          for (int l = 0; l < 4; ++l) Cs[e * CSS + (rg * 4 + l) * 3] = ec * static_cast<double>(n[l]);
        }
      }
    } else {
      for (int idx = tid; idx < EP * RC; idx += GEN_THREADS) Cs[(idx / RC) * CSS + (idx % RC) * 3] = 0.0;
    }
    __syncthreads();
  }
  if (has_red) {
    // wait for the basis tile
    {
      uint32_t done = 0;
      while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(mbar))
            : "memory");
      }
    }
    // C[e][r][d] += sum_j F[j][e] * A[d][j][r]; thread = 4 epochs x RPT realizations x nd
    constexpr int RPT = RC / 16;
    const int rl = (tid & 15) * RPT;
    const int e0 = (tid >> 4) * 4;
    double acc[4][RPT][3];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int u = 0; u < RPT; ++u) acc[i][u][0] = acc[i][u][1] = acc[i][u][2] = 0.0;
    if (e0 < tile.n_ep) {
      for (int j = 0; j < J; ++j) {
        const double2 f01 = *reinterpret_cast<const double2*>(Fs + j * EP + e0);
        const double2 f23 = *reinterpret_cast<const double2*>(Fs + j * EP + e0 + 2);
        const double f[4] = {f01.x, f01.y, f23.x, f23.y};
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          if (d < nd) {
            double a[RPT];
#pragma unroll
            for (int u = 0; u < RPT; ++u) a[u] = As[(d * J + j) * RC + rl + u];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int u = 0; u < RPT; ++u) acc[i][u][d] = fma(f[i], a[u], acc[i][u][d]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int u = 0; u < RPT; ++u) {
          double* c = Cs + (e0 + i) * CSS + (rl + u) * 3;
          c[0] += acc[i][u][0];
          c[1] = acc[i][u][1];
          c[2] = acc[i][u][2];
        }
    }
    __syncthreads();
  }

  // ---- TOA stage ---------------------------------------------------------------------
  const int t4 = tid * 4;
  if (t4 >= tile.n_toa) return;
  const size_t gi = size_t(tile.toa_start) + t4;
  double w1[4] = {0, 0, 0, 0}, w2[4] = {0, 0, 0, 0}, dt[4] = {0, 0, 0, 0}, gw[4] = {0, 0, 0, 0}, det[4] = {0, 0, 0, 0};
  int el[4] = {0, 0, 0, 0}, gx[4] = {0, 0, 0, 0};
  if (flags & PTAR_F_WHITE) {
    const double4 a = *reinterpret_cast<const double4*>(P.w1 + gi);
    w1[0] = a.x; w1[1] = a.y; w1[2] = a.z; w1[3] = a.w;
    if (!(flags & PTAR_F_WHITE1)) {
      const double4 b = *reinterpret_cast<const double4*>(P.w2 + gi);
      w2[0] = b.x; w2[1] = b.y; w2[2] = b.z; w2[3] = b.w;
    }
  }
  if (has_epoch) {
    const ushort4 e = *reinterpret_cast<const ushort4*>(P.eloc + gi);
    el[0] = e.x; el[1] = e.y; el[2] = e.z; el[3] = e.w;
    if (nd > 1) {
      const double4 a = *reinterpret_cast<const double4*>(P.dtau + gi);
      dt[0] = a.x; dt[1] = a.y; dt[2] = a.z; dt[3] = a.w;
    }
  }
  const bool has_gwb = (flags & PTAR_F_GWB) && P.npts > 0;
  if (has_gwb) {
    const ushort4 g = *reinterpret_cast<const ushort4*>(P.gidx + gi);
    gx[0] = g.x; gx[1] = g.y; gx[2] = g.z; gx[3] = g.w;
    const double4 a = *reinterpret_cast<const double4*>(P.gw + gi);
    gw[0] = a.x; gw[1] = a.y; gw[2] = a.z; gw[3] = a.w;
  }
  if (flags & PTAR_F_DET) {
    const double4 a = *reinterpret_cast<const double4*>(P.det + gi);
    det[0] = a.x; det[1] = a.y; det[2] = a.z; det[3] = a.w;
  }
  const bool two_draws = (flags & PTAR_F_WHITE) && !(flags & PTAR_F_WHITE1);
  const uint32_t wblock = static_cast<uint32_t>(tile.toa_local0 + t4) >> 2;
  const int nvalid = min(4, tile.n_toa - t4);

  for (int r = 0; r < nr; ++r) {
    double v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = det[k];
    if (has_epoch) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const double* c = Cs + el[k] * CSS + r * 3;
        v[k] += (nd > 1) ? fma(dt[k], fma(dt[k], c[2], c[1]), c[0]) : c[0];
      }
    }
    if (has_gwb) {
      const double* Gr = P.G + (size_t(r0 + r) * P.n_psr + psr) * P.npts;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const double g0 = __ldg(Gr + gx[k]), g1 = __ldg(Gr + gx[k] + 1);
        v[k] += fma(gw[k], g1 - g0, g0);
      }
    }
    if (flags & PTAR_F_WHITE) {
      if (INJECT) {
        const double4 a = *reinterpret_cast<const double4*>(P.z1 + size_t(r0 + r) * P.ld_out + gi);
        v[0] = fma(w1[0], a.x, v[0]); v[1] = fma(w1[1], a.y, v[1]);
        v[2] = fma(w1[2], a.z, v[2]); v[3] = fma(w1[3], a.w, v[3]);
        if (two_draws) {
          const double4 b = *reinterpret_cast<const double4*>(P.z2 + size_t(r0 + r) * P.ld_out + gi);
          v[0] = fma(w2[0], b.x, v[0]); v[1] = fma(w2[1], b.y, v[1]);
          v[2] = fma(w2[2], b.z, v[2]); v[3] = fma(w2[3], b.w, v[3]);
        }
      } else {
        const uint64_t rid = static_cast<uint64_t>(P.real0 + r0 + r);
        float n[4];
        normals4(n, wblock, PTAR_K_WHITE1, psr, rid, P.seed);
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = fma(w1[k], static_cast<double>(n[k]), v[k]);
        if (two_draws) {
          normals4(n, wblock, PTAR_K_WHITE2, psr, rid, P.seed);
#pragma unroll
          for (int k = 0; k < 4; ++k) v[k] = fma(w2[k], static_cast<double>(n[k]), v[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k >= nvalid) v[k] = 0.0;
    double* o = P.out + size_t(r0 + r) * P.ld_out + gi;
    *reinterpret_cast<double4*>(o) = make_double4(v[0], v[1], v[2], v[3]);
  }
}

}  // namespace ptar
