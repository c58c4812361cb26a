// GWB synthesis on the 5th-generation tensor cores (tcgen05 / TMEM), throughput mode.
//
//   G[r][q] = sum_j A[knot(q)][j] * Zm[p(q)][r][j]            (red_noise.py:269-285 as one linear map, see ptar_gwb.cuh)
//
// fp64 has no tcgen05 kind, so the product is evaluated EXACTLY in integers (Ozaki-style error-free splitting):
// both operands are written as fixed-point numbers with six signed radix-256 digits,
//   A[n][j]  = sA[n] * sum_t a_t[n][j] 2^(-8(t+1)),     Zm[p][r][j] = sZ[p] * sum_s z_s[p][r][j] 2^(-8(s+1)),
// (a_t, z_s int8; sA, sZ powers of two chosen so that |value / scale| <= 1/4, i.e. 46-48 significant bits relative
// to the row maximum), every digit-slice product is an int8 x int8 -> int32 GEMM on tcgen05.mma.kind::i8 -- exact --
// and slice pairs of equal weight s + t = d accumulate into the same TMEM columns.  Diagonals d = 0..6 are kept (26 of
// the 36 slice pairs; the dropped ones weigh <= 2^-56 of full scale -- with d <= 5 only, the result was 1.5e-11 of
// the signal rms away from the fp64 kernel, measured; typical operands sit 6-7 bits below full scale).  The epilogue
// reads the seven int32 accumulators back with tcgen05.ld, combines them in fp64 (sum_d D_d 2^(-8 d)) and applies
// the two scales.  Measured against the fp64 DMMA kernel on the same input: tests/test_gpu_gwb_i8.py.
//
// Operand staging: the producers write both operands in the tensor core's canonical K-major no-swizzle layout
// (8 rows x 16 bytes core matrices), tile by tile, so that one k-chunk of one operand is a single contiguous
// bulk-async (TMA) copy:
//   ZS  [slice s][pulsar p][r-block of 128][k-chunk of 32 j][16 row groups][2 x 16-byte k][8 rows][16 bytes]   (4 KB per copy)
//   AS  [tile][k-chunk of 32 j][slice t][4 row groups][2][8][16]                                              (6 KB per copy)
// One instruction multiplies Z slice s (128 realizations x 32 j) with the STACK of A slices t = 0..min(5, 6-s) (N = 32
// per slice, up to 192) and lands in TMEM columns 32 (s + t) + knot: the stacking along N is what makes the 26 slice
// products cost 7 instructions per 32-j step instead of 26.
//
// CTA = one tile of 32 knots x 128 realizations, 256 threads: warp 0 lane 0 issues the bulk copies (3-stage mbarrier
// ring, 30 KB per stage), warp 1 allocates TMEM (256 columns) and its lane 0 issues the MMAs; then all 8 warps run the
// epilogue (warp w reads TMEM lanes 32 (w % 4).. and the knots 16 (w / 4)..).  92 KB of shared memory and half the
// tensor memory: TWO CTAs per SM, so the prologue (set-up, first copies) and the epilogue of one tile run under the
// main loop of the other -- with one 64-knot tile per SM (7 x 64 = 448 accumulator columns) they were 45 % of the
// kernel (measured per phase with tools/i8_timeline.py).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ptar {

constexpr int I8_SLICES = 6;          // digits per operand
constexpr int I8_DIAGS = 7;           // kept weights d = s + t = 0 .. 6
constexpr int I8_BM = 128;            // realizations per CTA (MMA M)
constexpr int I8_BN = 32;             // knots per CTA
constexpr int I8_BK = 32;             // j per pipeline stage (bytes, int8) = the K of one MMA
constexpr int I8_STAGES = 3;
constexpr int I8_A_BYTES = I8_BM * I8_BK;                 // one Z slice of one stage: 4 KB
constexpr int I8_B_BYTES = I8_SLICES * I8_BN * I8_BK;     // all A slices of one stage: 6 KB
constexpr int I8_STAGE_BYTES = I8_SLICES * I8_A_BYTES + I8_B_BYTES;   // 30 KB
constexpr int I8_TMEM_COLS = 256;                         // 7 diagonals x 32 knots = 224 used
constexpr int I8_THREADS = 256;
constexpr size_t I8_SMEM = size_t(I8_STAGES) * I8_STAGE_BYTES + 128;   // + barriers, tmem address

__device__ __forceinline__ uint32_t i8_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void i8_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(i8_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void i8_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(i8_smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void i8_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(i8_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void i8_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(i8_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(i8_smem_u32(bar))
               : "memory");
}
// all previously issued tcgen05.mma of this thread arrive on `bar` when they have completed
__device__ __forceinline__ void i8_umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(i8_smem_u32(bar)) : "memory");
}

// K-major, no swizzle: core matrix = 8 rows x 16 bytes stored contiguously (128 B).  lbo = byte distance between the
// two core matrices that are adjacent in K, sbo = byte distance between consecutive groups of 8 rows (M / N direction).
// Bits: [0,14) address >> 4, [16,30) lbo >> 4, [32,46) sbo >> 4, [46,48) version = 1 (sm_100), [61,64) layout = 0.
__device__ __forceinline__ uint64_t i8_smem_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  return uint64_t((smem_addr >> 4) & 0x3FFF) | (uint64_t((lbo >> 4) & 0x3FFF) << 16) | (uint64_t((sbo >> 4) & 0x3FFF) << 32) |
         (uint64_t(1) << 46);
}
// Instruction descriptor, kind::i8: D = s32 (bits [4,6) = 2), A = B = signed int8 (bits [7,10), [10,13) = 1), both K-major
// (bits 15, 16 = 0), N >> 3 at [17,23), M >> 4 at [24,29).
__device__ __forceinline__ uint32_t i8_instr_desc(int n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | (uint32_t(n >> 3) << 17) | (uint32_t(I8_BM >> 4) << 24);
}
__device__ __forceinline__ void i8_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// Zm (fp64, [p][r][J]) -> six int8 digit slices in the tile layout described above.  zinv[p] = 2^48 / sZ[p].
// One thread = one row x 16 consecutive j (one 16-byte k-piece of every slice).  Consecutive threads take consecutive
// rows of an 8-row core matrix, so a warp writes 4 x 128 contiguous bytes per slice.
__global__ void __launch_bounds__(256) gwb_slice_kernel(int8_t* __restrict__ ZS, const double* __restrict__ Zm,
                                                         const double* __restrict__ zinv, int P, int J, int Jpad, int64_t nreal,
                                                         int64_t rcap) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int npiece = Jpad / 16;
  const int64_t rows8 = (nreal + 7) / 8;                   // groups of 8 rows that hold at least one realization
  const int64_t per_p = rows8 * npiece * 8;
  if (idx >= per_p * P) return;
  const int p = static_cast<int>(idx / per_p);
  int64_t rem = idx % per_p;
  const int r8 = static_cast<int>(rem & 7);
  rem >>= 3;
  const int piece = static_cast<int>(rem % npiece);        // 16-byte piece along j
  const int64_t g8 = rem / npiece;                         // 8-row group
  const int64_t r = g8 * 8 + r8;
  __align__(16) int8_t dig[I8_SLICES][16];
  const double sc = zinv[p];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int j = piece * 16 + k;
    double v = 0.0;
    if (r < nreal && j < J) v = Zm[(size_t(p) * nreal + r) * J + j];
    long long X = __double2ll_rn(v * sc);
#pragma unroll
    for (int s = I8_SLICES - 1; s >= 0; --s) {           // balanced digits, least significant first
      const int d = static_cast<int>(static_cast<int8_t>(X & 0xFF));
      dig[s][k] = static_cast<int8_t>(d);
      X = (X - d) >> 8;
    }
  }
  const int64_t rblk = r / I8_BM;
  const int g = static_cast<int>((r % I8_BM) / 8);
  const int kch = piece / 2, c = piece % 2;
  const int64_t nkch = Jpad / I8_BK;
  const int64_t n_rblk = rcap / I8_BM;
  const size_t tile_off = ((((rblk * nkch + kch) * 16 + g) * 2 + c) * 8 + r8) * 16;
#pragma unroll
  for (int s = 0; s < I8_SLICES; ++s) {
    int8_t* dst = ZS + ((size_t(s) * P + p) * n_rblk) * (nkch * I8_A_BYTES) + tile_off;
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(dig[s]);
  }
}

// Diagnostics (ptar_debug_i8_timestamps): when set, the CTAs of r-block 0 record clock64() at the phase boundaries of
// each role, 8 slots per tile: start, setup done, loads issued, first stage landed, MMAs issued, accumulators complete,
// epilogue done, k-chunks.
__device__ long long* g_i8_dbg = nullptr;

// ---------------------------------------------------------------------------------------------------------------
// grid = (r-blocks, tiles); the r-block index is fastest so the CTAs that share a tile's A slices run together.
// tile_list[tile] = {pulsar, first compact column, columns (<= 32), k extent}; AS tile index = blockIdx.y (the host
// builds AS in tile_list order).  colscale[q] = sA[knot(q)] * 2^-16; zscale[p] = sZ[p].
__global__ void __launch_bounds__(I8_THREADS, 2)
gwb_synth_i8_kernel(double* __restrict__ G, int64_t g_ld, int64_t g_ldr, const int8_t* __restrict__ AS, const double* __restrict__ colscale,
                    const int8_t* __restrict__ ZS, const double* __restrict__ zscale, int P, int J, int Jpad, int64_t nreal,
                    int64_t rcap, const int32_t* __restrict__ tile_list) {
  extern __shared__ __align__(128) unsigned char i8_smem[];
  unsigned char* stage_base = i8_smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(i8_smem + size_t(I8_STAGES) * I8_STAGE_BYTES);
  uint64_t* empty = full + I8_STAGES;
  uint64_t* tmem_full = empty + I8_STAGES;
  uint32_t* tmem_addr_smem = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int32_t* tl = tile_list + size_t(blockIdx.y) * 4;
  const int p = tl[0], kn0 = tl[1], kcnt = tl[2];
  const int kend = min(J, tl[3]);
  const int nk = (kend + I8_BK - 1) / I8_BK;               // k-chunks of 32 with a non-zero A column
  const int64_t rblk = blockIdx.x;
  const int64_t nkch = Jpad / I8_BK;
  const int64_t n_rblk = rcap / I8_BM;

  long long* dbg = (g_i8_dbg != nullptr && blockIdx.x == 0) ? g_i8_dbg + size_t(blockIdx.y) * 8 : nullptr;
  if (dbg && tid == 0) {
    dbg[0] = clock64();
    dbg[7] = nk;
  }
  auto load_stage = [&](int kc) {          // one elected thread: 6 Z-slice tiles + the stacked A-slice tile of k-chunk kc
    const int st = kc % I8_STAGES;
    unsigned char* sb = stage_base + size_t(st) * I8_STAGE_BYTES;
    i8_mbar_expect_tx(full + st, I8_STAGE_BYTES);
#pragma unroll
    for (int s = 0; s < I8_SLICES; ++s) {
      const int8_t* src = ZS + (((size_t(s) * P + p) * n_rblk + rblk) * nkch + kc) * I8_A_BYTES;
      i8_bulk_load(sb + s * I8_A_BYTES, src, I8_A_BYTES, full + st);
    }
    const int8_t* srcb = AS + (size_t(blockIdx.y) * nkch + kc) * I8_B_BYTES;
    i8_bulk_load(sb + I8_SLICES * I8_A_BYTES, srcb, I8_B_BYTES, full + st);
  };
  __shared__ double s_colscale[I8_BN];
  if (tid >= 64 && tid < 64 + I8_BN) {       // scales of this tile's columns (read by the epilogue from shared memory)
    const int n = tid - 64;
    s_colscale[n] = (n < kcnt) ? colscale[kn0 + n] * zscale[p] : 0.0;
  }
  if (tid == 0) {
    for (int s = 0; s < I8_STAGES; ++s) {
      i8_mbar_init(full + s, 1);
      i8_mbar_init(empty + s, 1);
    }
    i8_mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // the first ring of copies leaves before the TMEM allocation and the CTA-wide sync: their latency (~2.5k clk
    // measured) overlaps the set-up
    for (int kc = 0; kc < nk && kc < I8_STAGES; ++kc) load_stage(kc);
  }
  if (warp == 1) {  // one warp allocates the tensor memory; the address lands in shared memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(i8_smem_u32(tmem_addr_smem)),
                 "r"(I8_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_addr_smem;
  if (dbg && tid == 0) dbg[1] = clock64();

  if (tid == 0) {
    // ---- producer (warp 0, lane 0): the remaining k-chunks, each once the MMAs that read its stage have completed
    for (int kc = I8_STAGES; kc < nk; ++kc) {
      i8_mbar_wait(empty + kc % I8_STAGES, ((kc / I8_STAGES) & 1) ^ 1);
      load_stage(kc);
    }
    if (dbg) dbg[2] = clock64();
  } else if (tid == 32) {
    // ---- MMA issuer (warp 1, lane 0)
    for (int kc = 0; kc < nk; ++kc) {
      const int st = kc % I8_STAGES;
      const uint32_t ph = (kc / I8_STAGES) & 1;
      i8_mbar_wait(full + st, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (dbg && kc == 0) dbg[3] = clock64();
      const uint32_t sa = i8_smem_u32(stage_base + size_t(st) * I8_STAGE_BYTES);
      const uint32_t sbm = sa + I8_SLICES * I8_A_BYTES;
      // stage layout of one operand: [row group][2 k-pieces of 16 B][8 rows][16 B]: lbo = 128, sbo = 256; one MMA
      // consumes the 32 bytes of K of a stage
      const uint32_t first = (kc == 0) ? 0u : 1u;
      // (Z slice s, first A slice t0, A slices taken, opens): TMEM columns 32 (s + t0) ..; an instruction either
      // overwrites all its columns (first k-chunk only) or accumulates into all of them, so the slice that first
      // touches diagonal 6 (s = 1, t = 5) is issued on its own
      constexpr int kPlan[7][4] = {{0, 0, 6, 1}, {1, 0, 5, 0}, {1, 5, 1, 1}, {2, 0, 5, 0}, {3, 0, 4, 0}, {4, 0, 3, 0}, {5, 0, 2, 0}};
#pragma unroll
      for (int q = 0; q < 7; ++q) {
        const int s = kPlan[q][0], t0 = kPlan[q][1], take = kPlan[q][2];
        const uint64_t adesc = i8_smem_desc(sa + s * I8_A_BYTES, 128, 256);
        const uint64_t bdesc = i8_smem_desc(sbm + t0 * (I8_BN * I8_BK), 128, 256);
        const uint32_t dcol = tmem_base + uint32_t((s + t0) * I8_BN);
        i8_mma(dcol, adesc, bdesc, i8_instr_desc(take * I8_BN), kPlan[q][3] ? first : 1u);
      }
      i8_umma_commit(empty + st);                        // frees the stage once these MMAs have read it
    }
    i8_umma_commit(tmem_full);                           // accumulators complete
    if (dbg) dbg[4] = clock64();
  }

  // ---- epilogue: all 8 warps.  Warp w reads TMEM lanes 32 (w % 4) .. +31 (= realizations of this r-block) and the
  // knots 16 (w / 4) .. +15; the loads of diagonal d + 1 are in flight while diagonal d is folded in.
  {
    const int quarter = warp & 3, half = warp >> 2;
    const int64_t r = rblk * I8_BM + quarter * 32 + lane;
    i8_mbar_wait(tmem_full, 0);
    __syncwarp();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (dbg && tid == 64) dbg[5] = clock64();
#define I8_TMEM_LD16(v, addr)                                                                                                 \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                                     \
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"                              \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),   \
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])                      \
               : "r"(addr))
    double val[16];
#pragma unroll
    for (int n = 0; n < 16; ++n) val[n] = 0.0;
    if (nk > 0) {
      const uint32_t tbase = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(half * 16);
      uint32_t va[16], vb[16];
      I8_TMEM_LD16(va, tbase);
      double w = 1.0;
#pragma unroll
      for (int d = 0; d < I8_DIAGS; ++d) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (d + 1 < I8_DIAGS) {
          if (d & 1) { I8_TMEM_LD16(va, tbase + uint32_t((d + 1) * I8_BN)); }
          else       { I8_TMEM_LD16(vb, tbase + uint32_t((d + 1) * I8_BN)); }
        }
#pragma unroll
        for (int n = 0; n < 16; ++n) {
          // int32 -> double without the conversion pipe: 2^52 + 2^31 + x sits exactly in the mantissa of {0x43300000, x ^ 2^31}
          const uint32_t vi = (d & 1) ? vb[n] : va[n];
#ifdef I8_EPI_NOMATH
          val[n] = __hiloint2double(0x43300000, static_cast<int>(vi));
#else
          const double x = __hiloint2double(0x43300000, static_cast<int>(vi ^ 0x80000000u)) - 4503601774854144.0;
          val[n] = fma(x, w, val[n]);
#endif
        }
        w *= 0.00390625;                                 // 2^-8 per diagonal
      }
    }
#undef I8_TMEM_LD16
#ifdef I8_EPI_NOSTORE
    if (r < nreal && val[5] == 1.2345e-300) {
#else
    if (r < nreal) {
#endif
      // column-major grid G[kn0 + n][r]: the 32 lanes of a warp are 32 consecutive realizations -> one 256-byte store per
      // knot (row-major stores of a thread's knots were 16 bytes per 32-byte sector and cost 5.9k clk per tile)
      const int kpad = (kcnt + 1) & ~1;
      double* gcol = G + size_t(kn0 + half * 16) * g_ldr + r;
      const double* cs = s_colscale + half * 16;
#pragma unroll
      for (int n = 0; n < 16; ++n) {
        if (half * 16 + n < kpad) gcol[size_t(n) * g_ldr] = val[n] * cs[n];
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (dbg && tid == 64) dbg[6] = clock64();
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(I8_TMEM_COLS));
  }
}

}  // namespace ptar
