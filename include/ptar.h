/* ptar.h -- C ABI of the B200 synthetic-PTA residual generator (libptar_b200.so).
 *
 * The reference (bencebecsy/pta_replicator) has no FFI: its boundary is a set of
 * Python functions.  Each entry point below replaces the arithmetic body of one of
 * them and is what a binding on the reference side would call (INTEGRATION.md shows
 * the ctypes stub).  Conventions: plain pointers and sizes, no torch types; every
 * pointer is a DEVICE pointer unless the name ends in _host; the library never
 * allocates or frees user memory; all work is enqueued on the given cudaStream_t
 * (passed as void*); return 0 on success, negative on bad arguments / launch errors
 * (text from ptar_last_error()).  All floating point is IEEE fp64 unless noted.
 *
 * Reference lines replaced (paths relative to /root/reference/pta_replicator/):
 *   ptar_cholesky_lower    np.linalg.cholesky(ORF)                      red_noise.py:235
 *   ptar_fourier_basis     create_fourier_design_matrix_red             red_noise.py:36-103
 *   ptar_gwb_mix           w draws + np.dot(M, w)                       red_noise.py:238-240, :268
 *   ptar_gwb_synth         sqrt(C) scale, Hermitian pack, ifft, crop    red_noise.py:269-285
 *   ptar_gwb_slice_i8 /    the same linear map on the tcgen05 tensor cores (exact int8 digit
 *   ptar_gwb_synth_i8      slices, int32 TMEM accumulators; throughput mode)   red_noise.py:269-285
 *   ptar_cgw_delay         add_cgw arithmetic                           deterministic.py:98-163
 *   ptar_cw_catalog        loop_over_CWs[_parallel] (numba)             deterministic.py:321-561
 *   ptar_burst_delay       add_burst polarisation mix                   deterministic.py:771-780
 *   ptar_memory_delay      add_gw_memory ramp                           deterministic.py:864-873
 *   ptar_generate          efac/equad draw                              white_noise.py:105-109
 *                          U @ (ecorr*z)                                white_noise.py:182
 *                          F @ (sqrt(prior)*z)                          red_noise.py:126-128
 *                          interp1d(ut, Res)(toas)                      red_noise.py:286-287
 *                          (+ CGW / deterministic delays)               deterministic.py:160-165
 *   ptar_philox_normals    np.random.randn (throughput-mode stream)     SURVEY.md 3.6
 */
#ifndef PTAR_H
#define PTAR_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTAR_VERSION 200  /* major*100 + minor; 2.0: column-major grid, g_ldr / c_rows, tcgen05 GWB entry points, pshift phases */

/* Geometry limits of the fused generator kernel. */
#define PTAR_TILE_TOAS   1024  /* TOAs per tile (256 threads x 4)            */
#define PTAR_TILE_EPOCHS 64    /* kernel-epochs per tile                     */
#define PTAR_TOA_ALIGN   4     /* every pulsar segment starts at a multiple  */

/* ptar_gen_params.flags */
#define PTAR_F_WHITE    1u   /* efac*sigma*z1 + w2*z2                         */
#define PTAR_F_ECORR    2u   /* ecorr[epoch] * z[bucket(epoch)]               */
#define PTAR_F_RED      4u   /* Fourier-basis red noise                       */
#define PTAR_F_GWB      8u   /* linear interpolation of the GWB time grid     */
#define PTAR_F_DET      16u  /* precomputed deterministic delay (CGW ...)     */
#define PTAR_F_WHITE1   32u  /* throughput mode only: one N(0, w1^2+w2^2) draw per TOA
                                instead of two (identical distribution)       */

/* Philox stream tags ("kind") -- part of the counter, see ptar_philox_normals. */
#define PTAR_K_WHITE1 1
#define PTAR_K_WHITE2 2
#define PTAR_K_ECORR  3
#define PTAR_K_RED    4
#define PTAR_K_GWB    5

/* One tile of the packed TOA axis: <=1024 consecutive (time-sorted) TOAs of ONE pulsar
 * touching <=64 kernel-epochs. */
typedef struct {
  int32_t toa_start;   /* index into the packed per-TOA arrays (multiple of 4)          */
  int32_t n_toa;       /* real TOAs in this tile (<= PTAR_TILE_TOAS)                    */
  int32_t toa_local0;  /* index of toa_start within its pulsar (Philox counter base)    */
  int32_t ep_start;    /* first kernel-epoch (global index)                             */
  int32_t n_ep;        /* kernel-epochs touched (<= PTAR_TILE_EPOCHS)                   */
  int32_t psr;         /* pulsar index                                                  */
  int32_t nd;          /* Taylor terms used for red noise inside epochs (1..3)          */
  int32_t reserved;    /* c_row0: sum of n_ep over the preceding tiles (rows of the Cbuf blocks)        */
} ptar_tile;

typedef struct {
  /* geometry */
  int32_t n_psr;
  int32_t n_tiles;
  int32_t J;            /* red-noise basis columns (2 * components); 0 if no red noise  */
  int32_t npts;         /* GWB grid length; 0 if no GWB                                  */
  uint32_t flags;       /* PTAR_F_*                                                      */
  int32_t rn_convention;/* 0: (sin,cos) pairs, absolute t;  1: libstempo (cos,sin)      */
  const ptar_tile* tiles;
  /* per-TOA statics, packed axis of length ld_out */
  const double*   w1;    /* efac*sigma [s]                                               */
  const double*   w2;    /* efac*equad (t2equad) or equad (tnequad) [s]                  */
  const double*   dtau;  /* t - t_ref(epoch) [s]                                         */
  const uint16_t* eloc;  /* kernel-epoch index local to the tile                         */
  const double*   det;   /* deterministic delay [s]                                      */
  /* per-epoch statics (an epoch never straddles an ECORR bucket edge or a GWB grid knot) */
  const double*  ep_ecorr;   /* ecorr of the epoch's bucket [s]                          */
  const int32_t* ep_bucket;  /* ECORR bucket id within the pulsar                        */
  const int32_t* ep_gidx;    /* column of knot j in the compact grid G (knot j+1 is the next column) */
  const double*  ep_gw;      /* (t_ref - ut[j]) / (ut[j+1] - ut[j])                      */
  const double*  ep_ginv;    /* 1 / (ut[j+1] - ut[j])  [1/s]                             */
  const int64_t* psr_bucket_off; /* [n_psr]: offset of each pulsar in the injected zb axis */
  const double*  Ftile;      /* [n_tiles][J][64]: basis at epoch reference times          */
  /* per-pulsar red-noise statics */
  const double* rn_scale;    /* [n_psr][J] sqrt(prior)                                   */
  const double* rn_omega;    /* [n_psr][J/2] 2*pi*f_k                                    */
  /* compact GWB grid for this batch of realizations, COLUMN-major: G[q][g_ldr], q < g_ld compact knot columns (only
   * the knots next to some TOA of a pulsar are present, ptar_gwb_synth), row = realization of this call; g_ldr is a
   * multiple of 4 >= nreal rounded up to 4.  The synthesis writes 32 consecutive realizations per store and the
   * generator reads 4 consecutive realizations of a knot as one 32-byte sector. */
  const double* G;
  int64_t g_ld;
  int64_t g_ldr;
  /* injected standard-normal draws (parity mode); all NULL => Philox */
  const double* z1;   /* [nreal][ld_out]                                                 */
  const double* z2;   /* [nreal][ld_out]                                                 */
  const double* zb;   /* [nreal][n_bucket_total]                                         */
  const double* zrn;  /* [nreal][n_psr][J]                                               */
  int64_t n_bucket_total;
  /* RNG (throughput mode) */
  uint64_t seed;
  int64_t  real0;     /* global id of realization 0 of this call (multiple of 4)         */
  /* output: out[r][ld_out] */
  double* out;
  int64_t ld_out;
  int32_t nreal;
  int32_t rc;         /* realizations per CTA: 16 (256 threads) or 32 (512 threads); 0 = default */
  /* optional scratch for the two-kernel schedule (epoch kernel -> TOA kernel): at least
   * c_rows * ceil(nreal/rc) * (3 rc + 2) doubles (rc = 16: 50 per row); NULL selects the fused kernel.
   * The library checks cbuf_len against that size before it launches. */
  double* Cbuf;
  int64_t cbuf_len;   /* doubles available at Cbuf */
  int64_t c_rows;     /* sum of n_ep over all tiles (rows of the epoch blocks)  */
} ptar_gen_params;

int         ptar_version(void);
const char* ptar_last_error(void);

/* Lower Cholesky factor of `batch` SPD matrices A[b][n][n] (row-major) -> L (strict upper
 * part zeroed).  info[b] = 0, or k>0 if the leading minor of order k is not PD.  n <= 1024. */
int ptar_cholesky_lower(double* L, const double* A, int n, int batch, int* info, void* stream);

/* F[row][2k], F[row][2k+1] = trig(2*pi * tprime[row] * freqs[psr(row)][k] (+ phase[psr(row)][k])) in the reference's
 * operation order; convention 0 -> (sin, cos), 1 -> (cos, sin); phase = NULL or the `pshift` random phases
 * (red_noise.py:83-84).  out index: out[row_off[row] + col*col_stride] so the caller chooses row-major or per-tile
 * layouts. */
int ptar_fourier_basis(double* out, const int64_t* row_off, int64_t col_stride,
                       const double* tprime, const int32_t* row_psr, const double* freqs, const double* phase,
                       int K, int convention, int64_t nrows, void* stream);

/* Continuous-wave delay per TOA.  src[16] and psr_par[n_psr][4] = {fplus, fcross, cosMu, pd_sec}
 * are the scalar pre-factors of deterministic.py:51-105; mode 0 evolve, 1 phase_approx, 2 mono.
 * out[i] (+)= delay(t[i]) ; t = mjd*86400 - tref. */
int ptar_cgw_delay(double* out, const double* t, const int32_t* psr_of_toa, const double* psr_par,
                   const double* src, int mode, int psr_term, int accumulate, int64_t n, void* stream);

/* Sum of many continuous-wave sources for ONE pulsar (add_catalog_of_cws, deterministic.py:188-561; SURVEY.md
 * 8f row f1).  cat[8][n_src] = gwtheta, gwphi, mc [Msun], dist [Mpc], fgw [Hz], phase0, psi, inc (device);
 * phat_host[3] = pulsar unit vector (HOST); pdist in kpc or pphase (use_pphase); mode / psr_term as in
 * ptar_cgw_delay; NaN contributions are dropped like the reference does.  Scratch: pre[n_src][16],
 * partial[n_slices][n_toa] (device).  out[i] (+)= sum_s delay_s(t[i]); t = mjd*86400 - tref. */
int ptar_cw_catalog(double* out, const double* t, int64_t n_toa, const double* phat_host, const double* cat,
                    int64_t n_src, double pdist_kpc, double pphase, int use_pphase, int mode, int psr_term,
                    int accumulate, double* pre, double* partial, int n_slices, void* stream);

/* Burst of arbitrary waveform (add_burst, deterministic.py:718-793).  hplus / hcross are the
 * caller's waveform callables sampled at t = mjd*86400 - tref (the callables are Python objects in the reference too);
 * fplus / fcross the antenna pattern (:757-759).  out[i] (+)= -fplus (h+ cos2psi - hx sin2psi) - fcross (h+ sin2psi +
 * hx cos2psi), every product rounded separately like numpy. */
int ptar_burst_delay(double* out, const double* hplus, const double* hcross, double fplus, double fcross, double cos2psi,
                     double sin2psi, int accumulate, int64_t n, void* stream);

/* Burst with memory (add_gw_memory, deterministic.py:822-884): out[i] (+)= t[i] < t0 ? 0 : amp (t[i] - t0), with
 * t = mjd*86400, t0 = t0_mjd*86400 and amp = (cos(2 pol) fplus + sin(2 pol) fcross) * strain. */
int ptar_memory_delay(double* out, const double* t, double amp, double t0, int accumulate, int64_t n, void* stream);

/* Zm[p][r][j] = sum_q M[p][q] z[r][q][j]  (output pulsar-major: [n_psr][nreal][J]).  z is read from
 * zin[r][q][j] (parity) or drawn from Philox (zin == NULL; stream PTAR_K_GWB).  M is n_psr x n_psr lower
 * triangular, row-major. */
int ptar_gwb_mix(double* Zm, const double* M, const double* zin, int n_psr, int J,
                 int64_t nreal, uint64_t seed, int64_t real0, void* stream);

/* Compact grid: G[r][q] = sum_j A[knots[q]][j] * Zm[p(q)][r][j] for every column q of the knot list
 * (knots[g_ld]: per pulsar the sorted rows of A that its TOAs interpolate, each pulsar block starting at an
 * even column and padded to even length with -1).  tile_list[n_tiles][4] = {pulsar, first column, columns
 * (<= 64), k extent} enumerates 64-column blocks, heaviest first; with lower_tri the k loop stops at the
 * tile's k extent (A[n][j] == 0 for j > n). */
int ptar_gwb_synth(double* G, int64_t g_ld, int64_t g_ldr, const double* A, int64_t lda, const double* Zm, int J, int64_t nreal,
                   const int32_t* tile_list, int n_tiles, const int32_t* knots, int lower_tri, void* stream);

/* tcgen05 path of the synthesis (throughput mode; csrc/ptar_gwb_i8.cuh).  Both operands are fixed-point numbers with
 * PTAR_I8_SLICES signed radix-256 digits: value = scale * sum_s digit_s 2^(-8(s+1)), |value / scale| <= 1/4.
 * ptar_gwb_slice_i8: Zm[p][r][J] (fp64) -> ZS, int8 digits in the tensor core's K-major core-matrix tile layout
 *   [slice][pulsar][rcap/128 r-blocks][Jpad/32 k-chunks][16][2][8][16]; zinv[p] = 2^48 / zscale[p]; rcap (multiple of 128)
 *   is the row capacity the buffer was laid out for (>= nreal), Jpad a multiple of 32 (>= J).
 * ptar_gwb_synth_i8: G[q][r] = sum_j A[knot(q)][j] Zm[p(q)][r][j] from the digit slices; AS holds the digits of the
 *   gathered rows of A tile by tile, in tile_list order: [tile][Jpad/32][slice][4][2][8][16]; colscale[q] = (scale of
 *   row knot(q) of A) * 2^-16; tile_list as in ptar_gwb_synth but with 32-column blocks ( A lower triangular: k stops at the
 *    tile's k extent).  Exact int8 x int8 -> int32 products on tcgen05.mma.kind::i8; the slice pairs s + t <= 6 are kept
 *   (dropped weight <= 2^-56 of full scale), the result is rounded once per output in fp64.
 * ptar_gwb_mix_i8: ptar_gwb_mix (Philox draws) with the slicing fused into its epilogue: ZS directly, no fp64 Zm
 *   (n_psr <= 72; larger arrays use ptar_gwb_mix + ptar_gwb_slice_i8). */
#define PTAR_I8_SLICES 6
int ptar_gwb_mix_i8(int8_t* ZS, const double* M, const double* zinv, int n_psr, int J, int Jpad, int64_t nreal, int64_t rcap,
                    uint64_t seed, int64_t real0, void* stream);
int ptar_gwb_slice_i8(int8_t* ZS, const double* Zm, const double* zinv, int n_psr, int J, int Jpad, int64_t nreal,
                      int64_t rcap, void* stream);
int ptar_gwb_synth_i8(double* G, int64_t g_ld, int64_t g_ldr, const int8_t* AS, const double* colscale, const int8_t* ZS,
                      const double* zscale, int n_psr, int J, int Jpad, int64_t nreal, int64_t rcap,
                      const int32_t* tile_list, int n_tiles, void* stream);

/* Diagnostics: buf (device, n_tiles * 8 int64, or NULL to switch off) receives clock64() stamps of the phases of
 * ptar_gwb_synth_i8 for the CTAs of r-block 0 (start, setup, loads issued, first stage landed, MMAs issued,
 * accumulators complete, epilogue done, k-chunks); used by tools/i8_timeline.py. */
int ptar_debug_i8_timestamps(void* buf);

/* The fused generator: out[r][i] = white + ecorr + red + gwb + det for nreal realizations. */
int ptar_generate(const ptar_gen_params* p, void* stream);

/* The two launches of the two-kernel schedule one at a time (p->Cbuf != NULL): stage 1 = epoch kernel
 * (Fourier GEMM, ECORR, GWB grid -> Cbuf), stage 2 = TOA kernel (Cbuf + white noise -> out).  ptar_generate
 * issues both; this entry exists so a caller can time them separately. */
int ptar_generate_stage(const ptar_gen_params* p, int stage, void* stream);

/* Raw throughput-mode normals (fp32 Box-Muller of Philox4x32-10), for tests:
 * out[k] = normal(kind, psr, realization, idx0 + k), k < n.  Counter = (idx, kind | psr << 8,
 * realization >> 2); the four outputs of a counter are realizations 4g .. 4g+3, so a shard of
 * realizations is reproduced bit-for-bit wherever it is generated.  idx = TOA index within the
 * pulsar (white), ECORR bucket, Fourier column (red), grid column (GWB). */
int ptar_philox_normals(float* out, int kind, int psr, int64_t realization, int64_t idx0,
                        int64_t n, uint64_t seed, void* stream);

/* End-to-end job: GWB mix + synth + generate for realizations [real0, real0+nreal), then
 * (ptar_run_job_to_host) copy the residuals to pinned host memory, chunk by chunk, with the
 * copy of chunk c overlapping the generation of chunk c+1 on a second stream. */
typedef struct {
  ptar_gen_params gen;      /* gen.G / gen.out / gen.nreal / gen.real0 are set per chunk  */
  const double* M;          /* [n_psr][n_psr] lower Cholesky factor of the ORF            */
  const double* A;          /* [npts][lda] GWB synthesis matrix                           */
  int64_t lda;
  int32_t Jg;               /* columns of A                                               */
  int32_t lower_tri;
  const int32_t* tile_list; /* [n_syn_tiles][4], see ptar_gwb_synth                       */
  const int32_t* knots;     /* [gen.g_ld]                                                 */
  int32_t n_syn_tiles;
  int32_t reserved;
  double* Zm;               /* scratch [n_psr][chunk][Jg]                                 */
  double* Gbuf;             /* scratch [gen.g_ld][gen.g_ldr]                              */
  const double* gwb_zin;    /* parity mode: [nreal][n_psr][Jg] or NULL                    */
  /* tcgen05 synthesis (throughput mode only; all NULL / 0 selects the fp64 DMMA kernel) */
  const int8_t* AS;         /* digit slices of the gathered rows of A, tile_list_i8 order */
  const double* colscale;   /* [gen.g_ld]                                                 */
  int8_t* ZS;               /* scratch: digit slices of Zm, PTAR_I8_SLICES*n_psr*rcap*Jpad bytes */
  const double* zscale;     /* [n_psr]                                                    */
  const double* zinv;       /* [n_psr] 2^48 / zscale                                      */
  const int32_t* tile_list_i8; /* [n_syn_tiles_i8][4]: 32-column blocks, pulsar-major     */
  int64_t rcap;             /* row capacity of ZS (multiple of 128, >= chunk)             */
  int32_t Jpad;             /* multiple of 32, >= Jg                                      */
  int32_t n_syn_tiles_i8;
} ptar_job;

/* Peer delivery over NVLink without a collective library (multi-GPU row of SURVEY.md 8e; the reference has no
 * distributed code): a rank exports its result buffer (CUDA IPC), every other rank maps it into its own address space
 * and pushes its realization chunks straight into it with the copy engines, overlapped with generation.
 * ptar_peer_export: handle_host[64] <- IPC handle of the allocation that contains dev_ptr, *offset <- dev_ptr - base.
 * ptar_peer_open:   maps an exported allocation into the current device's context (lazy peer access); *base_out is the
 *                   mapped base (add the exporter's offset); ptar_peer_close unmaps it.
 * ptar_peer_copy:   asynchronous device-to-device copy (UVA) between local and mapped peer memory on `stream`. */
int ptar_peer_export(const void* dev_ptr, void* handle_host, int64_t* offset);
int ptar_peer_open(const void* handle_host, void** base_out);
int ptar_peer_close(void* base);
int ptar_peer_copy(void* dst, const void* src, int64_t bytes, void* stream);

int ptar_run_job(const ptar_job* job, int64_t real0, int32_t nreal, double* out, void* stream);
int ptar_run_job_to_host(const ptar_job* job, int64_t real0, int64_t nreal, int32_t chunk,
                         double* out_host, double* dev_buf0, double* dev_buf1, void* stream0, void* stream1);

#ifdef __cplusplus
}
#endif
#endif /* PTAR_H */
