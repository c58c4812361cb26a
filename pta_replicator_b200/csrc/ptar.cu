// libptar_b200.so -- C ABI (include/ptar.h) over the sm_100a kernels.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/ptar.h"
#include "ptar_generate.cuh"
#include "ptar_gwb.cuh"
#include "ptar_gwb_i8.cuh"
#include "ptar_rng.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, const char* detail = "") {
  snprintf(g_err, sizeof(g_err), fmt, detail);
  return code;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of a kernel: remember, per kernel and
// device, that the opt-in has been made (one process may drive several GPUs).
constexpr int kMaxDevices = 64;
struct SmemOptIn {
  bool done[kMaxDevices] = {};
};
template <class Kernel>
int opt_in_smem(Kernel k, SmemOptIn& state, size_t bytes, const char* what) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = -1;
  if (dev >= 0 && state.done[dev]) return 0;
  const cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
    return -100;
  }
  if (dev >= 0) state.done[dev] = true;
  return 0;
}

int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return -100;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------
// red_noise.py:92-101 -- operation order kept: (2*pi * t) * f, then sin / cos.
__global__ void fourier_basis_kernel(double* __restrict__ out, const int64_t* __restrict__ row_off,
                                     int64_t col_stride, const double* __restrict__ tprime,
                                     const int32_t* __restrict__ row_psr, const double* __restrict__ freqs,
                                     const double* __restrict__ phase, int K, int convention, int64_t nrows) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= nrows * K) return;
  const int64_t row = idx / K;
  const int k = static_cast<int>(idx % K);
  const double f = freqs[size_t(row_psr[row]) * K + k];
  double arg = __dmul_rn(__dmul_rn(6.283185307179586, tprime[row]), f);
  if (phase) arg = __dadd_rn(arg, phase[size_t(row_psr[row]) * K + k]);   // pshift (red_noise.py:83-84)
  double s, c;
  sincos(arg, &s, &c);
  double* o = out + row_off[row];
  o[int64_t(2 * k) * col_stride] = convention ? c : s;
  o[int64_t(2 * k + 1) * col_stride] = convention ? s : c;
}

// deterministic.py:98-163.  src = {w0, fac1, fac2, fac3, phase0(orbital), w053, incfac1,
// incfac2, cos2psi, sin2psi}; psr_par[p] = {fplus, fcross, cosMu, pd_seconds}.
__global__ void cgw_kernel(double* __restrict__ out, const double* __restrict__ t, const int32_t* __restrict__ psr_of_toa,
                           const double* __restrict__ psr_par, const double* __restrict__ src, int mode, int psr_term,
                           int accumulate, int64_t n) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double w0 = src[0], fac1 = src[1], fac2 = src[2], fac3 = src[3], phase0 = src[4], w053 = src[5];
  const double inc1 = src[6], inc2 = src[7], c2p = src[8], s2p = src[9];
  const double* pp = psr_par + size_t(psr_of_toa[i]) * 4;
  const double fplus = pp[0], fcross = pp[1], cosMu = pp[2], pd = pp[3];
  const double toa = t[i];
  const double tp = toa - pd * (1 - cosMu);
  double omega, omega_p, phase, phase_p;
  if (mode == 0) {
    omega = w0 * pow(1 - fac1 * toa, -3.0 / 8);
    omega_p = w0 * pow(1 - fac1 * tp, -3.0 / 8);
    phase = phase0 + fac2 * (w053 - pow(omega, -5.0 / 3));
    phase_p = phase0 + fac2 * (w053 - pow(omega_p, -5.0 / 3));
  } else if (mode == 1) {
    omega = w0;
    omega_p = w0 * pow(1 + fac1 * pd * (1 - cosMu), -3.0 / 8);
    phase = phase0 + omega * toa;
    phase_p = phase0 + fac2 * (w053 - pow(omega_p, -5.0 / 3)) + omega_p * toa;
  } else {
    omega = w0;
    omega_p = omega;
    phase = phase0 + omega * toa;
    phase_p = phase0 + omega * tp;
  }
  const double At = sin(2 * phase) * inc1, Bt = cos(2 * phase) * inc2;
  const double Atp = sin(2 * phase_p) * inc1, Btp = cos(2 * phase_p) * inc2;
  const double alpha = fac3 / cbrt(omega), alpha_p = fac3 / cbrt(omega_p);
  const double rplus = alpha * (At * c2p + Bt * s2p), rcross = alpha * (-At * s2p + Bt * c2p);
  const double rplus_p = alpha_p * (Atp * c2p + Btp * s2p), rcross_p = alpha_p * (-Atp * s2p + Btp * c2p);
  const double res = psr_term ? fplus * (rplus_p - rplus) + fcross * (rcross_p - rcross)
                              : -fplus * rplus - fcross * rcross;
  out[i] = accumulate ? out[i] + res : res;
}

// ---------------------------------------------------------------------------------------
// Catalog of continuous-wave sources (deterministic.py:188-561; the reference's numba loops).
// cw_prefactor_kernel: the per-source scalars of :340-376 for one pulsar direction ->
//   pre[s][16] = {w0, fac1, fac2, fac3, phase0/2, w0^-5/3, incfac1, incfac2, cos2psi, sin2psi, fplus, fcross,
//                 cosMu, pd_sec, -, -}
// cw_catalog_kernel: thread = one TOA, CTA = 128 TOAs x one slice of the catalog staged through shared memory;
// NaN contributions (a binary that has already merged, :433-438) are dropped; slices are summed in a fixed order
// by cw_reduce_kernel so the result is deterministic.
__global__ void cw_prefactor_kernel(double* __restrict__ pre, const double* __restrict__ cat, int64_t n_src,
                                    double px, double py, double pz, double pdist_kpc, double pphase, int use_pphase) {
  const int64_t s = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (s >= n_src) return;
  const double gwtheta = cat[s], gwphi = cat[n_src + s];
  const double mc = cat[2 * n_src + s] * 4.925838061995516e-06;     // SOLAR2S (constants.py:6)
  const double dist = cat[3 * n_src + s] * 102927125054338.98;      // MPC2S   (constants.py:8)
  const double fgw = cat[4 * n_src + s], phase0 = cat[5 * n_src + s] / 2, psi = cat[6 * n_src + s], inc = cat[7 * n_src + s];
  const double w0 = 3.141592653589793 * fgw;
  const double cgt = cos(gwtheta), cgp = cos(gwphi), sgt = sin(gwtheta), sgp = sin(gwphi);
  const double mx = sgp, my = -cgp;
  const double nx = -cgt * cgp, ny = -cgt * sgp, nz = sgt;
  const double ox = -sgt * cgp, oy = -sgt * sgp, oz = -cgt;
  const double mc53 = pow(mc, 5.0 / 3);
  const double mdp = mx * px + my * py + 0.0 * pz, ndp = nx * px + ny * py + nz * pz, odp = ox * px + oy * py + oz * pz;
  const double cosMu = -odp;
  double pd = use_pphase ? pphase / (2 * 3.141592653589793 * fgw * (1 - cosMu)) / 102927125054.33899 : pdist_kpc;
  pd *= 102927125054.33899;                                          // KPC2S (constants.py:7)
  double* o = pre + s * 16;
  o[0] = w0;
  o[1] = 256.0 / 5 * mc53 * pow(w0, 8.0 / 3);
  o[2] = 1.0 / 32 / mc53;
  o[3] = mc53 / dist;
  o[4] = phase0;
  o[5] = pow(w0, -5.0 / 3);
  o[6] = 0.5 * (3 + cos(2 * inc));
  o[7] = 2 * cos(inc);
  o[8] = cos(2 * psi);
  o[9] = sin(2 * psi);
  o[10] = 0.5 * (mdp * mdp - ndp * ndp) / (1 + odp);
  o[11] = (mdp * ndp) / (1 + odp);
  o[12] = cosMu;
  o[13] = pd;
  o[14] = o[15] = 0.0;
}

constexpr int CW_TOAS = 128, CW_STAGE = 32;

__global__ void __launch_bounds__(CW_TOAS) cw_catalog_kernel(double* __restrict__ partial, const double* __restrict__ t,
                                                             int64_t n_toa, const double* __restrict__ pre, int64_t n_src,
                                                             int64_t src_per_slice, int mode, int psr_term) {
  __shared__ double sp[CW_STAGE][16];
  const int64_t i = int64_t(blockIdx.x) * CW_TOAS + threadIdx.x;
  const int64_t s_begin = int64_t(blockIdx.y) * src_per_slice;
  const int64_t s_end = min(n_src, s_begin + src_per_slice);
  const double toa = i < n_toa ? t[i] : 0.0;
  double acc = 0.0;
  for (int64_t s0 = s_begin; s0 < s_end; s0 += CW_STAGE) {
    const int64_t rem = s_end - s0;
    const int cnt = rem < CW_STAGE ? static_cast<int>(rem) : CW_STAGE;
    __syncthreads();
    for (int k = threadIdx.x; k < cnt * 16; k += CW_TOAS) sp[k >> 4][k & 15] = pre[(s0 + (k >> 4)) * 16 + (k & 15)];
    __syncthreads();
    for (int k = 0; k < cnt; ++k) {
      const double* q = sp[k];
      const double w0 = q[0], fac1 = q[1], fac2 = q[2], fac3 = q[3], phase0 = q[4], w053 = q[5];
      const double tp = toa - q[13] * (1 - q[12]);
      double omega, omega_p, phase, phase_p;
      if (mode == 0) {
        omega = w0 * pow(1 - fac1 * toa, -3.0 / 8);
        omega_p = w0 * pow(1 - fac1 * tp, -3.0 / 8);
        phase = phase0 + fac2 * (w053 - pow(omega, -5.0 / 3));
        phase_p = phase0 + fac2 * (w053 - pow(omega_p, -5.0 / 3));
      } else if (mode == 1) {
        omega = w0;
        omega_p = w0 * pow(1 + fac1 * q[13] * (1 - q[12]), -3.0 / 8);
        phase = phase0 + omega * toa;
        phase_p = phase0 + fac2 * (w053 - pow(omega_p, -5.0 / 3)) + omega_p * toa;
      } else {
        omega = w0;
        omega_p = omega;
        phase = phase0 + omega * toa;
        phase_p = phase0 + omega * tp;
      }
      double s2, c2, s2p, c2p;
      sincos(2 * phase, &s2, &c2);
      sincos(2 * phase_p, &s2p, &c2p);
      const double At = s2 * q[6], Bt = c2 * q[7], Atp = s2p * q[6], Btp = c2p * q[7];
      const double alpha = fac3 / cbrt(omega), alpha_p = fac3 / cbrt(omega_p);
      const double rplus = alpha * (At * q[8] + Bt * q[9]), rcross = alpha * (-At * q[9] + Bt * q[8]);
      const double rplus_p = alpha_p * (Atp * q[8] + Btp * q[9]), rcross_p = alpha_p * (-Atp * q[9] + Btp * q[8]);
      const double r = psr_term ? q[10] * (rplus_p - rplus) + q[11] * (rcross_p - rcross) : -q[10] * rplus - q[11] * rcross;
      acc += isnan(r) ? 0.0 : r;
    }
  }
  if (i < n_toa) partial[int64_t(blockIdx.y) * n_toa + i] = acc;
}

__global__ void cw_reduce_kernel(double* __restrict__ out, const double* __restrict__ partial, int64_t n_toa, int n_slices,
                                 int accumulate) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_toa) return;
  double a = 0.0;
  for (int k = 0; k < n_slices; ++k) a += partial[int64_t(k) * n_toa + i];
  out[i] = accumulate ? out[i] + a : a;
}

// deterministic.py:771-780: elliptically polarised burst from the two waveform samples of each TOA.  The products are
// rounded one by one (no FMA contraction) so the result is the reference's bit for bit.
__global__ void burst_kernel(double* __restrict__ out, const double* __restrict__ hplus, const double* __restrict__ hcross,
                             double fplus, double fcross, double c2, double s2, int accumulate, int64_t n) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double hp = hplus[i], hx = hcross[i];
  const double rplus = __dsub_rn(__dmul_rn(hp, c2), __dmul_rn(hx, s2));
  const double rcross = __dadd_rn(__dmul_rn(hp, s2), __dmul_rn(hx, c2));
  const double res = __dsub_rn(__dmul_rn(-fplus, rplus), __dmul_rn(fcross, rcross));
  out[i] = accumulate ? out[i] + res : res;
}

// deterministic.py:868-872: burst with memory, a ramp after the burst epoch.
__global__ void memory_kernel(double* __restrict__ out, const double* __restrict__ t, double amp, double t0, int accumulate,
                              int64_t n) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double res = t[i] < t0 ? 0.0 : __dmul_rn(amp, __dsub_rn(t[i], t0));
  out[i] = accumulate ? out[i] + res : res;
}

__global__ void philox_normals_kernel(float* __restrict__ out, int kind, int psr, int64_t realization, int64_t idx0,
                                      int64_t n, uint64_t seed) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= n) return;
  float z[4];
  ptar::normals4(z, static_cast<uint32_t>(idx0 + k), kind, psr, static_cast<uint64_t>(realization) >> 2, ptar::philox_keys(seed));
  out[k] = z[realization & 3];
}

template <int RC, bool INJECT, int WHITE, int DET, int STAGE>
int launch_stage(const ptar_gen_params& p, cudaStream_t st) {
  // STAGE 2 only needs the Cs block; STAGE 0/1 need the GEMM operands
  size_t smem = ptar::gen_smem_bytes(p.J, RC);
  if (STAGE == 2) smem = 16 + sizeof(double) * ptar::EP * ptar::gen_css(RC);
  if (smem > 227 * 1024) return fail(-3, "ptar_generate: J too large for shared memory%s");
  static SmemOptIn optin;  // per instantiation
  if (int rc = opt_in_smem(ptar::gen_kernel<RC, INJECT, WHITE, DET, STAGE>, optin, 227 * 1024, "ptar_generate")) return rc;
  const dim3 grid((p.nreal + RC - 1) / RC, p.n_tiles);
  if (grid.y > 65535) return fail(-3, "ptar_generate: more than 65535 tiles%s");
  if (STAGE != 0 && p.Cbuf) {  // the Cs blocks of this build must fit the scratch the caller gave
    const int64_t need = p.c_rows * int64_t(grid.x) * ptar::gen_css(RC);
    if (p.c_rows <= 0 || p.cbuf_len < need) return fail(-2, "ptar_generate: Cbuf too small for this rc (need c_rows * ceil(nreal/rc) * (3 rc + 2) doubles)%s");
  }
  ptar::gen_kernel<RC, INJECT, WHITE, DET, STAGE><<<grid, ptar::gen_threads(RC), smem, st>>>(p, ptar::philox_keys(p.seed));
  return check_launch("ptar_generate");
}

// Epoch kernel of the two-kernel schedule (RC = 16 blocks of Cbuf, two per CTA).
template <bool INJECT>
int launch_epoch(const ptar_gen_params& p, cudaStream_t st) {
  const size_t smem = ptar::epoch_smem_bytes(p.J);
  if (smem > 227 * 1024) return fail(-3, "ptar_generate: J too large for shared memory%s");
  static SmemOptIn optin;
  if (int rc = opt_in_smem(ptar::epoch_kernel<INJECT>, optin, 227 * 1024, "ptar_generate (epoch kernel)")) return rc;
  const dim3 grid((p.nreal + ptar::EPK_RB - 1) / ptar::EPK_RB, p.n_tiles);
  if (grid.y > 65535) return fail(-3, "ptar_generate: more than 65535 tiles%s");
  {
    const int64_t need = p.c_rows * int64_t((p.nreal + 15) / 16) * ptar::EPK_CSS;
    if (p.c_rows <= 0 || p.cbuf_len < need) return fail(-2, "ptar_generate: Cbuf too small (need c_rows * ceil(nreal/16) * 50 doubles)%s");
  }
  ptar::epoch_kernel<INJECT><<<grid, ptar::EPK_THREADS, smem, st>>>(p, ptar::philox_keys(p.seed));
  return check_launch("ptar_generate (epoch kernel)");
}

thread_local int g_only_stage = 0;  // set by ptar_generate_stage: 1 = epoch kernel only, 2 = TOA kernel only

template <int RC, bool INJECT, int WHITE, int DET>
int launch_gen(const ptar_gen_params& p, cudaStream_t st) {
  const bool has_epoch = (p.flags & (PTAR_F_RED | PTAR_F_ECORR | PTAR_F_GWB)) != 0;
  if (p.Cbuf && has_epoch) {  // two-kernel schedule
    if (g_only_stage != 2) {
      // the epoch kernel does not depend on WHITE / DET
      const int rc = RC == 16 ? launch_epoch<INJECT>(p, st) : launch_stage<RC, INJECT, -1, -1, 1>(p, st);
      if (rc || g_only_stage == 1) return rc;
    }
    return launch_stage<RC, INJECT, WHITE, DET, 2>(p, st);
  }
  if (g_only_stage == 1) return 0;   // nothing to do: no epoch terms (or fused schedule)
  return launch_stage<RC, INJECT, WHITE, DET, 0>(p, st);
}

template <bool INJECT, bool SLICE>
int launch_mix(double* Zm, const double* M, const double* zin, int n_psr, int J, int64_t nreal, uint64_t seed, int64_t real0,
               int8_t* ZS, const double* zinv, int Jpad, int64_t rcap, cudaStream_t st, const char* what) {
  const dim3 grid((J + ptar::MIX_JT - 1) / ptar::MIX_JT,
                  static_cast<unsigned>((nreal + 4 * ptar::MX_GROUPS - 1) / (4 * ptar::MX_GROUPS)));
  if (grid.y > 65535) return fail(-3, "ptar_gwb_mix: more than 262140 realizations per call%s");
  const int KP = (n_psr + 3) & ~3, NP = (n_psr + 7) & ~7;
  // draws [KP][132] + M [NP][KP + 4] doubles: 113 KB at 67 pulsars (2 CTAs per SM), the 227 KB limit is reached at 115 pulsars
  const size_t smem = sizeof(double) * (size_t(KP) * ptar::MX_ZS + size_t(NP) * (KP + 4));
  if (smem > 227 * 1024) return fail(-3, "ptar_gwb_mix: more than 115 pulsars do not fit the shared-memory tile%s");
  static SmemOptIn optin;   // per instantiation
  if (int rc = opt_in_smem(ptar::gwb_mix_dmma_kernel<INJECT, SLICE>, optin, 227 * 1024, what)) return rc;
  ptar::gwb_mix_dmma_kernel<INJECT, SLICE><<<grid, 256, smem, st>>>(Zm, M, zin, n_psr, J, nreal, ptar::philox_keys(seed), real0, ZS,
                                                                     zinv, Jpad, rcap);
  return check_launch(what);
}

}  // namespace

extern "C" {

int ptar_version(void) { return PTAR_VERSION; }
const char* ptar_last_error(void) { return g_err; }

int ptar_cholesky_lower(double* L, const double* A, int n, int batch, int* info, void* stream) {
  if (!L || !A || n <= 0 || n > 1024 || batch <= 0) return fail(-1, "ptar_cholesky_lower: bad argument%s");
  ptar::cholesky_kernel<<<batch, 256, 0, static_cast<cudaStream_t>(stream)>>>(L, A, n, info);
  return check_launch("ptar_cholesky_lower");
}

int ptar_fourier_basis(double* out, const int64_t* row_off, int64_t col_stride, const double* tprime,
                       const int32_t* row_psr, const double* freqs, const double* phase, int K, int convention,
                       int64_t nrows, void* stream) {
  if (!out || !row_off || !tprime || !row_psr || !freqs || K <= 0 || nrows < 0)
    return fail(-1, "ptar_fourier_basis: bad argument%s");
  if (nrows == 0) return 0;
  const int64_t total = nrows * K;
  fourier_basis_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      out, row_off, col_stride, tprime, row_psr, freqs, phase, K, convention, nrows);
  return check_launch("ptar_fourier_basis");
}

int ptar_cgw_delay(double* out, const double* t, const int32_t* psr_of_toa, const double* psr_par, const double* src,
                   int mode, int psr_term, int accumulate, int64_t n, void* stream) {
  if (!out || !t || !psr_of_toa || !psr_par || !src || mode < 0 || mode > 2 || n < 0)
    return fail(-1, "ptar_cgw_delay: bad argument%s");
  if (n == 0) return 0;
  cgw_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      out, t, psr_of_toa, psr_par, src, mode, psr_term, accumulate, n);
  return check_launch("ptar_cgw_delay");
}

int ptar_burst_delay(double* out, const double* hplus, const double* hcross, double fplus, double fcross, double cos2psi,
                     double sin2psi, int accumulate, int64_t n, void* stream) {
  if (!out || !hplus || !hcross || n < 0) return fail(-1, "ptar_burst_delay: bad argument%s");
  if (n == 0) return 0;
  burst_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      out, hplus, hcross, fplus, fcross, cos2psi, sin2psi, accumulate, n);
  return check_launch("ptar_burst_delay");
}

int ptar_memory_delay(double* out, const double* t, double amp, double t0, int accumulate, int64_t n, void* stream) {
  if (!out || !t || n < 0) return fail(-1, "ptar_memory_delay: bad argument%s");
  if (n == 0) return 0;
  memory_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(out, t, amp, t0,
                                                                                                     accumulate, n);
  return check_launch("ptar_memory_delay");
}

int ptar_cw_catalog(double* out, const double* t, int64_t n_toa, const double* phat_host, const double* cat, int64_t n_src,
                    double pdist_kpc, double pphase, int use_pphase, int mode, int psr_term, int accumulate, double* pre,
                    double* partial, int n_slices, void* stream) {
  if (!out || !t || !phat_host || !cat || !pre || !partial || n_toa <= 0 || n_src <= 0 || n_slices <= 0 || mode < 0 || mode > 2)
    return fail(-1, "ptar_cw_catalog: bad argument%s");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cw_prefactor_kernel<<<static_cast<unsigned>((n_src + 255) / 256), 256, 0, st>>>(pre, cat, n_src, phat_host[0], phat_host[1],
                                                                                   phat_host[2], pdist_kpc, pphase, use_pphase);
  if (int rc = check_launch("ptar_cw_catalog (prefactors)")) return rc;
  const int64_t per = (n_src + n_slices - 1) / n_slices;
  const dim3 grid(static_cast<unsigned>((n_toa + CW_TOAS - 1) / CW_TOAS), static_cast<unsigned>(n_slices));
  if (grid.y > 65535) return fail(-3, "ptar_cw_catalog: more than 65535 slices%s");
  cw_catalog_kernel<<<grid, CW_TOAS, 0, st>>>(partial, t, n_toa, pre, n_src, per, mode, psr_term);
  if (int rc = check_launch("ptar_cw_catalog (sources x TOAs)")) return rc;
  cw_reduce_kernel<<<static_cast<unsigned>((n_toa + 255) / 256), 256, 0, st>>>(out, partial, n_toa, n_slices, accumulate);
  return check_launch("ptar_cw_catalog (slice reduction)");
}

int ptar_gwb_mix(double* Zm, const double* M, const double* zin, int n_psr, int J, int64_t nreal, uint64_t seed,
                 int64_t real0, void* stream) {
  if (!Zm || !M || n_psr <= 0 || J <= 0 || nreal <= 0) return fail(-1, "ptar_gwb_mix: bad argument%s");
  if (!zin && (real0 & 3)) return fail(-2, "ptar_gwb_mix: real0 must be a multiple of 4%s");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (zin) return launch_mix<true, false>(Zm, M, zin, n_psr, J, nreal, seed, real0, nullptr, nullptr, 0, 0, st, "ptar_gwb_mix");
  return launch_mix<false, false>(Zm, M, zin, n_psr, J, nreal, seed, real0, nullptr, nullptr, 0, 0, st, "ptar_gwb_mix");
}

int ptar_gwb_mix_i8(int8_t* ZS, const double* M, const double* zinv, int n_psr, int J, int Jpad, int64_t nreal, int64_t rcap,
                    uint64_t seed, int64_t real0, void* stream) {
  if (!ZS || !M || !zinv || n_psr <= 0 || J <= 0 || nreal <= 0) return fail(-1, "ptar_gwb_mix_i8: bad argument%s");
  if (n_psr > 8 * ptar::MX_MAXNT) return fail(-3, "ptar_gwb_mix_i8: more than 72 pulsars (use ptar_gwb_mix + ptar_gwb_slice_i8)%s");
  if (real0 & 3) return fail(-2, "ptar_gwb_mix_i8: real0 must be a multiple of 4%s");
  if (Jpad < J || (Jpad % ptar::I8_BK) || rcap < nreal || (rcap % ptar::I8_BM))
    return fail(-2, "ptar_gwb_mix_i8: need Jpad %% 32 == 0 >= J and rcap %% 128 == 0 >= nreal%s");
  return launch_mix<false, true>(nullptr, M, nullptr, n_psr, J, nreal, seed, real0, ZS, zinv, Jpad, rcap, static_cast<cudaStream_t>(stream),
                                 "ptar_gwb_mix_i8");
}

int ptar_gwb_synth(double* G, int64_t g_ld, int64_t g_ldr, const double* A, int64_t lda, const double* Zm, int J, int64_t nreal,
                   const int32_t* tile_list, int n_tiles, const int32_t* knots, int lower_tri, void* stream) {
  if (!G || !A || !Zm || !tile_list || !knots || J <= 0 || nreal <= 0 || n_tiles <= 0 || g_ld <= 0)
    return fail(-1, "ptar_gwb_synth: bad argument%s");
  if ((J & 3) || (lda & 1) || lda < J || (g_ld & 1) || (g_ldr & 3) || g_ldr < nreal)
    return fail(-2, "ptar_gwb_synth: need J %% 4 == 0, even lda >= J, even g_ld, g_ldr %% 4 == 0 >= nreal%s");
  const int64_t r_blocks = (nreal + ptar::DM_BC - 1) / ptar::DM_BC;
  if (r_blocks > 65535) return fail(-3, "ptar_gwb_synth: too many realizations per call%s");
  static SmemOptIn optin;
  if (int rc = opt_in_smem(ptar::gwb_synth_dmma_kernel, optin, ptar::DM_SMEM, "ptar_gwb_synth")) return rc;
  const dim3 grid(static_cast<unsigned>(n_tiles), static_cast<unsigned>(r_blocks));
  ptar::gwb_synth_dmma_kernel<<<grid, 256, ptar::DM_SMEM, static_cast<cudaStream_t>(stream)>>>(
      G, g_ld, g_ldr, A, lda, Zm, J, nreal, tile_list, knots, lower_tri);
  return check_launch("ptar_gwb_synth");
}

int ptar_gwb_slice_i8(int8_t* ZS, const double* Zm, const double* zinv, int n_psr, int J, int Jpad, int64_t nreal, int64_t rcap,
                      void* stream) {
  if (!ZS || !Zm || !zinv || n_psr <= 0 || J <= 0 || nreal <= 0) return fail(-1, "ptar_gwb_slice_i8: bad argument%s");
  if (Jpad < J || (Jpad % ptar::I8_BK) || rcap < nreal || (rcap % ptar::I8_BM))
    return fail(-2, "ptar_gwb_slice_i8: need Jpad %% 32 == 0 >= J and rcap %% 128 == 0 >= nreal%s");
  const int64_t total = ((nreal + 7) / 8) * (Jpad / 16) * 8 * n_psr;
  ptar::gwb_slice_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      ZS, Zm, zinv, n_psr, J, Jpad, nreal, rcap);
  return check_launch("ptar_gwb_slice_i8");
}

int ptar_gwb_synth_i8(double* G, int64_t g_ld, int64_t g_ldr, const int8_t* AS, const double* colscale, const int8_t* ZS, const double* zscale,
                      int n_psr, int J, int Jpad, int64_t nreal, int64_t rcap, const int32_t* tile_list, int n_tiles,
                      void* stream) {
  if (!G || !AS || !colscale || !ZS || !zscale || !tile_list || n_psr <= 0 || J <= 0 || nreal <= 0 || n_tiles <= 0 || g_ld <= 0)
    return fail(-1, "ptar_gwb_synth_i8: bad argument%s");
  if (Jpad < J || (Jpad % ptar::I8_BK) || rcap < nreal || (rcap % ptar::I8_BM) || (g_ld & 1) || (g_ldr & 3) || g_ldr < nreal)
    return fail(-2, "ptar_gwb_synth_i8: need Jpad %% 32 == 0 >= J, rcap %% 128 == 0 >= nreal, even g_ld, g_ldr %% 4 == 0 >= nreal%s");
  if ((reinterpret_cast<uintptr_t>(AS) | reinterpret_cast<uintptr_t>(ZS)) & 15)
    return fail(-2, "ptar_gwb_synth_i8: AS / ZS must be 16-byte aligned%s");
  if (n_tiles > 65535) return fail(-3, "ptar_gwb_synth_i8: more than 65535 tiles%s");
  static SmemOptIn optin;
  if (int rc = opt_in_smem(ptar::gwb_synth_i8_kernel, optin, ptar::I8_SMEM, "ptar_gwb_synth_i8")) return rc;
  const dim3 grid(static_cast<unsigned>((nreal + ptar::I8_BM - 1) / ptar::I8_BM), static_cast<unsigned>(n_tiles));
  ptar::gwb_synth_i8_kernel<<<grid, ptar::I8_THREADS, ptar::I8_SMEM, static_cast<cudaStream_t>(stream)>>>(
      G, g_ld, g_ldr, AS, colscale, ZS, zscale, n_psr, J, Jpad, nreal, rcap, tile_list);
  return check_launch("ptar_gwb_synth_i8");
}

int ptar_debug_i8_timestamps(void* buf) {
  const cudaError_t e = cudaMemcpyToSymbol(ptar::g_i8_dbg, &buf, sizeof(buf));
  if (e != cudaSuccess) return fail(-100, "ptar_debug_i8_timestamps: %s", cudaGetErrorString(e));
  return 0;
}

int ptar_generate(const ptar_gen_params* pp, void* stream) {
  if (!pp) return fail(-1, "ptar_generate: null params%s");
  const ptar_gen_params& p = *pp;
  if (!p.out || !p.tiles || p.n_tiles <= 0 || p.nreal <= 0 || p.n_psr <= 0)
    return fail(-1, "ptar_generate: bad geometry%s");
  if ((p.ld_out & 3) || (reinterpret_cast<uintptr_t>(p.out) & 31)) return fail(-2, "ptar_generate: out must be 32-byte aligned, ld_out %% 4 == 0%s");
  if ((p.flags & PTAR_F_RED) && (p.J <= 0 || (p.J & 1) || !p.Ftile || !p.rn_scale || !p.rn_omega))
    return fail(-2, "ptar_generate: red noise needs even J, Ftile, rn_scale, rn_omega%s");
  if ((p.flags & PTAR_F_WHITE) && (!p.w1 || (!(p.flags & PTAR_F_WHITE1) && !p.w2)))
    return fail(-2, "ptar_generate: white noise needs w1/w2%s");
  if ((p.flags & PTAR_F_ECORR) && (!p.ep_ecorr || !p.ep_bucket)) return fail(-2, "ptar_generate: ECORR needs ep_ecorr/ep_bucket%s");
  if ((p.flags & (PTAR_F_ECORR | PTAR_F_RED)) && (!p.eloc || !p.dtau)) return fail(-2, "ptar_generate: epoch terms need eloc/dtau%s");
  if ((p.flags & PTAR_F_GWB) && (!p.G || p.npts <= 1 || p.g_ld <= 0 || !p.ep_gidx || !p.ep_gw || !p.ep_ginv || !p.eloc || !p.dtau))
    return fail(-2, "ptar_generate: GWB needs G, g_ld, ep_gidx, ep_gw, ep_ginv, eloc, dtau%s");
  if ((p.flags & PTAR_F_GWB) && ((p.g_ldr & 3) || p.g_ldr < ((p.nreal + 3) & ~3) || (reinterpret_cast<uintptr_t>(p.G) & 31)))
    return fail(-2, "ptar_generate: G must be 32-byte aligned, column-major with g_ldr %% 4 == 0 >= nreal rounded up to 4%s");
  if ((p.flags & PTAR_F_DET) && !p.det) return fail(-2, "ptar_generate: DET needs det%s");
  const bool inject = p.z1 || p.z2 || p.zb || p.zrn;
  if (inject) {
    if ((p.flags & PTAR_F_WHITE) && (!p.z1 || (!(p.flags & PTAR_F_WHITE1) && !p.z2))) return fail(-2, "ptar_generate: injected white draws missing%s");
    if ((p.flags & PTAR_F_ECORR) && (!p.zb || !p.psr_bucket_off)) return fail(-2, "ptar_generate: injected ECORR draws missing%s");
    if ((p.flags & PTAR_F_RED) && !p.zrn) return fail(-2, "ptar_generate: injected red-noise draws missing%s");
  } else if (p.real0 & 3) {
    return fail(-2, "ptar_generate: real0 must be a multiple of 4%s");
  }
  if (p.Cbuf && p.cbuf_len <= 0) return fail(-2, "ptar_generate: Cbuf given without cbuf_len%s");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int rc = p.rc ? p.rc : 16;
  if (rc != 16 && rc != 32) return fail(-2, "ptar_generate: rc must be 16 or 32%s");
  if (inject) return launch_gen<16, true, -1, -1>(p, st);
  const int white = !(p.flags & PTAR_F_WHITE) ? 0 : ((p.flags & PTAR_F_WHITE1) ? 1 : 2);
  const bool det = (p.flags & PTAR_F_DET) != 0;
  if (rc == 32) {
    switch (white * 2 + (det ? 1 : 0)) {
      case 0: return launch_gen<32, false, 0, 0>(p, st);
      case 1: return launch_gen<32, false, 0, 1>(p, st);
      case 2: return launch_gen<32, false, 1, 0>(p, st);
      case 3: return launch_gen<32, false, 1, 1>(p, st);
      case 4: return launch_gen<32, false, 2, 0>(p, st);
      default: return launch_gen<32, false, 2, 1>(p, st);
    }
  }
  switch (white * 2 + (det ? 1 : 0)) {
    case 0: return launch_gen<16, false, 0, 0>(p, st);
    case 1: return launch_gen<16, false, 0, 1>(p, st);
    case 2: return launch_gen<16, false, 1, 0>(p, st);
    case 3: return launch_gen<16, false, 1, 1>(p, st);
    case 4: return launch_gen<16, false, 2, 0>(p, st);
    default: return launch_gen<16, false, 2, 1>(p, st);
  }
}

int ptar_generate_stage(const ptar_gen_params* pp, int stage, void* stream) {
  if (stage < 1 || stage > 2) return fail(-1, "ptar_generate_stage: stage must be 1 (epoch kernel) or 2 (TOA kernel)%s");
  if (!pp || !pp->Cbuf) return fail(-2, "ptar_generate_stage: needs the two-kernel schedule (Cbuf)%s");
  g_only_stage = stage;
  const int rc = ptar_generate(pp, stream);
  g_only_stage = 0;
  return rc;
}

int ptar_philox_normals(float* out, int kind, int psr, int64_t realization, int64_t idx0, int64_t n, uint64_t seed,
                        void* stream) {
  if (!out || n < 0) return fail(-1, "ptar_philox_normals: bad argument%s");
  if (n == 0) return 0;
  philox_normals_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      out, kind, psr, realization, idx0, n, seed);
  return check_launch("ptar_philox_normals");
}

int ptar_peer_export(const void* dev_ptr, void* handle_host, int64_t* offset) {
  if (!dev_ptr || !handle_host || !offset) return fail(-1, "ptar_peer_export: bad argument%s");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  cudaPointerAttributes attr;
  cudaError_t e = cudaPointerGetAttributes(&attr, dev_ptr);
  if (e != cudaSuccess || attr.type != cudaMemoryTypeDevice) return fail(-2, "ptar_peer_export: not a device pointer%s");
  // the handle describes the whole allocation: find its base with the driver's range query (through the runtime's
  // driver entry point, so that libcuda need not be linked)
  typedef int (*range_fn)(unsigned long long*, size_t*, unsigned long long);
  static range_fn get_range = nullptr;
  if (!get_range) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    e = cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr);
    if (e != cudaSuccess || !fn) return fail(-100, "ptar_peer_export: cuMemGetAddressRange unavailable%s");
    get_range = reinterpret_cast<range_fn>(fn);
  }
  unsigned long long base = 0;
  size_t size = 0;
  if (get_range(&base, &size, reinterpret_cast<unsigned long long>(dev_ptr)) != 0)
    return fail(-100, "ptar_peer_export: cuMemGetAddressRange failed%s");
  e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle_host), reinterpret_cast<void*>(base));
  if (e != cudaSuccess) return fail(-100, "ptar_peer_export: cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
  *offset = static_cast<int64_t>(reinterpret_cast<unsigned long long>(dev_ptr) - base);
  return 0;
}

int ptar_peer_open(const void* handle_host, void** base_out) {
  if (!handle_host || !base_out) return fail(-1, "ptar_peer_open: bad argument%s");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_host, sizeof(h));
  const cudaError_t e = cudaIpcOpenMemHandle(base_out, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return fail(-100, "ptar_peer_open: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
  return 0;
}

int ptar_peer_close(void* base) {
  if (!base) return 0;
  const cudaError_t e = cudaIpcCloseMemHandle(base);
  if (e != cudaSuccess) return fail(-100, "ptar_peer_close: %s", cudaGetErrorString(e));
  return 0;
}

int ptar_peer_copy(void* dst, const void* src, int64_t bytes, void* stream) {
  if (!dst || !src || bytes < 0) return fail(-1, "ptar_peer_copy: bad argument%s");
  if (bytes == 0) return 0;
  const cudaError_t e = cudaMemcpyAsync(dst, src, static_cast<size_t>(bytes), cudaMemcpyDefault, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(-100, "ptar_peer_copy: %s", cudaGetErrorString(e));
  return 0;
}

int ptar_run_job(const ptar_job* job, int64_t real0, int32_t nreal, double* out, void* stream) {
  if (!job || !out || nreal <= 0) return fail(-1, "ptar_run_job: bad argument%s");
  ptar_gen_params g = job->gen;
  g.real0 = real0;
  g.nreal = nreal;
  g.out = out;
  const bool inject = g.z1 || g.z2 || g.zb || g.zrn || job->gwb_zin;
  if (g.flags & PTAR_F_GWB) {
    if (!job->M || !job->A || !job->Zm || !job->Gbuf || !job->tile_list || !job->knots)
      return fail(-2, "ptar_run_job: GWB buffers missing%s");
    const bool i8 = job->AS && !inject;
    int rc = 0;
    if (i8 && g.n_psr <= 8 * ptar::MX_MAXNT) {   // mixing emits the digit slices directly
      if (!job->ZS || !job->zinv) return fail(-2, "ptar_run_job: tcgen05 GWB buffers missing%s");
      rc = ptar_gwb_mix_i8(job->ZS, job->M, job->zinv, g.n_psr, job->Jg, job->Jpad, nreal, job->rcap, g.seed, real0, stream);
    } else {
      rc = ptar_gwb_mix(job->Zm, job->M, inject ? job->gwb_zin : nullptr, g.n_psr, job->Jg, nreal, g.seed, real0, stream);
      if (!rc && i8) {
        if (!job->ZS || !job->zinv) return fail(-2, "ptar_run_job: tcgen05 GWB buffers missing%s");
        rc = ptar_gwb_slice_i8(job->ZS, job->Zm, job->zinv, g.n_psr, job->Jg, job->Jpad, nreal, job->rcap, stream);
      }
    }
    if (rc) return rc;
    if (i8) {   // tcgen05 path: exact int8 GEMMs on the digit slices, fp64 fix-up
      if (!job->colscale || !job->zscale || !job->tile_list_i8) return fail(-2, "ptar_run_job: tcgen05 GWB buffers missing%s");
      rc = ptar_gwb_synth_i8(job->Gbuf, g.g_ld, g.g_ldr, job->AS, job->colscale, job->ZS, job->zscale, g.n_psr, job->Jg, job->Jpad, nreal,
                             job->rcap, job->tile_list_i8, job->n_syn_tiles_i8, stream);
    } else {
      rc = ptar_gwb_synth(job->Gbuf, g.g_ld, g.g_ldr, job->A, job->lda, job->Zm, job->Jg, nreal, job->tile_list, job->n_syn_tiles,
                          job->knots, job->lower_tri, stream);
    }
    if (rc) return rc;
    g.G = job->Gbuf;
  }
  return ptar_generate(&g, stream);
}

int ptar_run_job_to_host(const ptar_job* job, int64_t real0, int64_t nreal, int32_t chunk, double* out_host,
                         double* dev_buf0, double* dev_buf1, void* stream0, void* stream1) {
  if (!job || !out_host || !dev_buf0 || !dev_buf1 || chunk <= 0 || (chunk & 3) || nreal <= 0)
    return fail(-1, "ptar_run_job_to_host: bad argument (chunk must be a positive multiple of 4)%s");
  cudaStream_t s0 = static_cast<cudaStream_t>(stream0), s1 = static_cast<cudaStream_t>(stream1);
  cudaEvent_t gen_done[2] = {nullptr, nullptr}, copy_done[2] = {nullptr, nullptr};
  int rc = 0;
  cudaError_t e = cudaSuccess;
  auto ok = [&](cudaError_t r) {  // first runtime error wins; later calls are skipped
    if (e == cudaSuccess && r != cudaSuccess) e = r;
    return e == cudaSuccess;
  };
  for (int i = 0; i < 2; ++i) {
    ok(cudaEventCreateWithFlags(&gen_done[i], cudaEventDisableTiming));
    ok(cudaEventCreateWithFlags(&copy_done[i], cudaEventDisableTiming));
  }
  double* bufs[2] = {dev_buf0, dev_buf1};
  const int64_t ld = job->gen.ld_out;
  int64_t done = 0;
  for (int c = 0; done < nreal && e == cudaSuccess; ++c) {
    const int b = c & 1;
    const int32_t n = static_cast<int32_t>(nreal - done < chunk ? nreal - done : chunk);
    if (c >= 2 && !ok(cudaStreamWaitEvent(s0, copy_done[b], 0))) break;  // buffer free again
    ptar_job j = *job;
    if (j.gen.z1) j.gen.z1 += done * ld;
    if (j.gen.z2) j.gen.z2 += done * ld;
    if (j.gen.zb) j.gen.zb += done * j.gen.n_bucket_total;
    if (j.gen.zrn) j.gen.zrn += done * int64_t(j.gen.n_psr) * j.gen.J;
    if (j.gwb_zin) j.gwb_zin += done * int64_t(j.gen.n_psr) * j.Jg;
    rc = ptar_run_job(&j, real0 + done, n, bufs[b], s0);
    if (rc) break;
    if (!ok(cudaEventRecord(gen_done[b], s0)) || !ok(cudaStreamWaitEvent(s1, gen_done[b], 0))) break;
    if (!ok(cudaMemcpyAsync(out_host + done * ld, bufs[b], sizeof(double) * size_t(n) * ld, cudaMemcpyDeviceToHost, s1))) break;
    if (!ok(cudaEventRecord(copy_done[b], s1))) break;
    done += n;
  }
  // drain both streams even after an error so no copy is still writing out_host when we return
  const cudaError_t d1 = cudaStreamSynchronize(s1), d0 = cudaStreamSynchronize(s0);
  ok(d1);
  ok(d0);
  for (int i = 0; i < 2; ++i) {
    if (gen_done[i]) cudaEventDestroy(gen_done[i]);
    if (copy_done[i]) cudaEventDestroy(copy_done[i]);
  }
  if (rc) return rc;
  ok(cudaGetLastError());
  if (e != cudaSuccess) return fail(-100, "ptar_run_job_to_host: %s", cudaGetErrorString(e));
  return 0;
}

}  // extern "C"
