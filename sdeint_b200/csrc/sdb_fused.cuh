// sdb_fused.cuh -- the production SRI2/SRS2 kernel for LINEAR systems f = A y, g_k = B_k y with
// on-device Philox + Ikpw/Jkpw noise: the whole time loop of sdeint/integrate.py:494-522 plus
// wiener.py:52 and :67-106 in ONE launch.  One lane per sample path, one warp per 32 paths.
//
// Why it looks the way it does (profiles/r01_microbench.md):
//  * On sm_100a a DFMA cannot take a constant-bank operand, broadcast LDS delivers 8 B/clk/SM
//    and LDCU-fed DFMAs are latency-bound, so "one uniform matrix element per DFMA" caps a
//    thread-per-path kernel below 50 % of the FP64 peak.  All products with the SHARED matrices
//    (G_n = [B_j y], f = A y, and the stage-2 sum) therefore run as FP64 tensor instructions
//    (mma.sync m8n8k4 f64 -> DMMA.8x8x4): the matrix fragment is one register per lane and every
//    instruction does 8 MACs per lane.  DMMA shares the FP64 pipe with DFMA (same 37 TFLOP/s peak),
//    so this is about operand delivery, not extra flops.
//  * Per-path products (the Levy-area series, G_n I) have no shared operand; they stay on DFMA with
//    one lane per path and 4-10 DFMAs per loaded double.
//  * g_k is linear, so the two stage-2 evaluations collapse exactly:
//        (sqrt(h)/2) (g_k(H20 + s_k) - g_k(H20 - s_k)) = B_k (sqrt(h) s_k),   s_k = sum1[:,k]
//    (integrate.py:509-521); sqrt(h) sum1 = G_n I is what the kernel accumulates.  The flop count
//    used for the roofline is reduced accordingly (subtract 2md^2 + 2dm).
//
//  * What bounds it (tools/micro/philox.cu): the FP64 pipe PLUS the 20 IMAD.WIDE of
//    every Philox4x32-10 block, which issue at ~3.7 cycles each and do not overlap with DFMA/DMMA.
//    Per warp-step: 12 560 FP64 cycles + 4070 Philox cycles; the kernel runs at 82 % of that floor,
//    and neither more resident warps (variants 10, 13) nor deeper software ILP move it further.
//
// Per step and per row block b (rows R_b of the state, 4 + 4 + 2 for d = 10):
//   P1  DMMA   [B_j[R_b,:] ; A[R_b,:]] Y  -> shared staging (rows x 32 paths)
//   P2  DFMA   s~[R_b,:] = G_n[R_b,:] I   (Levy areas from shared memory), G_n dW, drift
//              bookkeeping; rows 8.. of the stage-2 sum on DFMA; s~ -> staging
//   P3  DMMA   w[0:8] += [B_k[0:8, R_b]]_k vec(s~[R_b,:])
// Shared memory per warp: staging 48 x 32, Levy areas 45 x 32, dW 10 x 32 doubles (26.4 KB).
#pragma once
#include "sdb_kernels.cuh"
#ifndef SDB_FUSED_AHEAD
#define SDB_FUSED_AHEAD 0  // 1: the random bits of step n + 1 (32-bit integer work) are produced between the
#endif                     //    FP64 instructions of the row blocks of step n -- where issue slots are idle --
                           //    and wait in tensor memory (one 32x32b lane per path, 4 columns per 128-bit
                           //    block); the noise phase then only loads, converts and transforms them.
                           //    Needs n_terms <= SDB_FUSED_AHEAD_TERMS and an even number of Wiener processes.
#ifndef SDB_FUSED_AFFINE
#define SDB_FUSED_AFFINE 0  // 1 (NVRTC builds): g_k = B_k y + c_k -- the constants (after the tables, [j][row]) are
#endif                      //    added to the entries of G_n as P2 reads them; the stage-2 differences cancel them
#ifndef SDB_FUSED_ANTIPHASE
#define SDB_FUSED_ANTIPHASE 0  // 1: the two warps of a sub-partition (w, w + 4) are held half a step apart by a
#endif                         //    named barrier: one generates noise while the other runs its row blocks
#ifndef SDB_FUSED_INPHASE
#define SDB_FUSED_INPHASE 0  // 1: the two warps of a sub-partition enter every phase together (named barrier at the
#endif                       //    top of the step and before the P1 / P3 tensor phases of every row block)
#define SDB_PAIR_BAR() asm volatile("bar.sync %0, 64;" ::"r"(1 + (((int)threadIdx.x >> 5) & 3)) : "memory")
#ifndef SDB_FUSED_AHEAD_TERMS
#define SDB_FUSED_AHEAD_TERMS 5
#endif
#if SDB_FUSED_AHEAD
#include "sdb_tmem.cuh"
#endif

#ifdef __CUDACC__
#define SDB_CX __host__ __device__ constexpr
#else
#define SDB_CX constexpr
#endif

#ifndef SDB_FUSED_THREADS
#define SDB_FUSED_THREADS 256  // 384: three warps per sub-partition (needs <= 168 registers: RB 2, PARK_Y, DW_REGS)
#endif
#define SDB_FUSED_WARPS (SDB_FUSED_THREADS / 32)
#ifndef SDB_FUSED_DW_REGS
#define SDB_FUSED_DW_REGS 0  // 1: dW stays in registers through the row blocks (no dW rows in shared memory)
#endif
#define SDB_FUSED_MAX_TERMS 256
#ifndef SDB_FUSED_KUNROLL
#define SDB_FUSED_KUNROLL 1  // unroll factor of the Levy-series loop over k
#endif
#ifndef SDB_FUSED_VOL_MMA
#define SDB_FUSED_VOL_MMA 1  // 1: asm volatile DMMA (strictly ordered), 0: plain asm (schedulable)
#endif
#ifndef SDB_FUSED_VOL_SMEM
#define SDB_FUSED_VOL_SMEM 1  // 1: volatile shared-memory reads in P2, 0: plain reads between __syncwarp()s
#endif
#ifndef SDB_FUSED_BATCH
#define SDB_FUSED_BATCH 1  // 1: Philox + normal transform of M/2 pairs in lock-step (explicit ILP)
#endif
#ifndef SDB_FUSED_BATCH_ORDER
#define SDB_FUSED_BATCH_ORDER 0  // 1: angle polynomials before the table-dependent log part
#endif
#ifndef SDB_FUSED_SWP
#define SDB_FUSED_SWP 0  // 1: software pipeline of the series loop -- the random bits (32-bit ALU / IMAD work) of
#endif                   //    the NEXT batch are produced next to the FP64 transform of the current one.  Measured
                         //    4.7 % SLOWER (1.877e9 vs 1.969e9 path-steps/s): ptxas interleaves the two streams
                         //    instruction by instruction and the lock-step ILP of the batches is lost
#ifndef SDB_FUSED_BATCH_XZ
#define SDB_FUSED_BATCH_XZ 0  // 1: X_k and Y_k pairs of one series term in ONE lock-step batch of M
#endif
#ifndef SDB_FUSED_RB
#define SDB_FUSED_RB 4  // state rows per block of the P1/P2/P3 pipeline
#endif
#ifndef SDB_FUSED_PARK_Y
#define SDB_FUSED_PARK_Y 0  // 1: the state waits in shared memory while the noise is generated
#endif
#ifndef SDB_FUSED_PARK_DW
#define SDB_FUSED_PARK_DW 0  // 1: dW is re-read from shared memory inside the Levy loop
#endif
#if SDB_FUSED_VOL_SMEM
#define SDB_SMEM_CV const volatile
#else
#define SDB_SMEM_CV const
#endif

// Everything below depends on the tuning macros, so each variant's translation unit gets its own
// namespace (host inline templates with identical names would otherwise be merged by the linker).
#ifndef SDB_FUSED_VARIANT
#define SDB_FUSED_VARIANT 0
#endif
#define SDB_FUSED_NS2(v) sdb_fused_v##v
#define SDB_FUSED_NS1(v) SDB_FUSED_NS2(v)
#define SDB_FUSED_NS SDB_FUSED_NS1(SDB_FUSED_VARIANT)
namespace SDB_FUSED_NS {

#ifndef SDB_FUSED_A2
#define SDB_FUSED_A2 1  // 1: rows of A^2 ride in the P1 product (its padding rows at d = m = 10): the second
#endif                  //    drift evaluation f(Y + f h) h = f h + h^2 A^2 Y needs no DFMA loop of its own
                        // -1: ANY drift with linear diffusion g_k = B_k y (NVRTC builds): no drift rows in the
                        //    tables; the functor SDB_FUSED_DRIFT_T (F::eval(params, y, t, out), sdb_kernels.cuh)
                        //    is evaluated one lane per path at Y_n and at H20 (integrate.py:509, :514)
#ifndef SDB_FUSED_KSPLIT
#define SDB_FUSED_KSPLIT 0  // 1: when d mod 4 is 1 or 2 the last one or two state components do not pad a
#endif                      //    third DMMA k-tile of the P1 product: their contribution is added to the
                            //    accumulator fragments with DFMAs (d = 10: 16 DFMA instead of 4 DMMA per row tile)
// XR: extra row groups of the P1 product after the drift rows (1: the rows of A^2)
template <int D, int M, int XR = 0>
struct SdbFusedShape {
  static constexpr int MM = M * (M - 1) / 2;
  static constexpr int KS = (SDB_FUSED_KSPLIT && D >= 4 && (D % 4 == 1 || D % 4 == 2)) ? D % 4 : 0;
  static constexpr int RB = (SDB_FUSED_RB > 0 && D > SDB_FUSED_RB) ? SDB_FUSED_RB : 4;  // rows per block
  static constexpr int NB = (D + RB - 1) / RB;        // row blocks
  static constexpr int LT = KS ? D / 4 : (D + 3) / 4;  // DMMA k-tiles over the state index l
  static constexpr int IT = D / 8;                    // 8-row tiles of the stage-2 sum done by DMMA
  static constexpr int IL = D - 8 * IT;               // leftover rows done by DFMA
  static SDB_CX int rn(int b) { return b < NB - 1 ? RB : D - RB * (NB - 1); }
  static SDB_CX int r0(int b) { return RB * b; }
  static SDB_CX int nr(int b) { return (M + 1 + XR) * rn(b); }  // P1 rows: m columns + drift (+ A^2)
  static SDB_CX int t1(int b) { return (nr(b) + 7) / 8; }       // P1 row tiles
  static SDB_CX int kt3(int b) { return (rn(b) * M + 3) / 4; }  // P3 k tiles
  static SDB_CX int max2(int a, int b) { return a > b ? a : b; }
  static SDB_CX int grows() {
    int g = max2(8, D);
    for (int b = 0; b < NB; ++b) g = max2(g, max2(8 * t1(b), 4 * kt3(b)));
    return g;
  }
  static SDB_CX int p1_off(int b) {  // offset (in fragments) of block b in the P1 table
    int o = 0;
    for (int q = 0; q < b; ++q) o += t1(q) * LT;
    return o;
  }
  static SDB_CX int p3_off(int b) {
    int o = p1_off(NB);
    for (int q = 0; q < b; ++q) o += kt3(q) * IT;
    return o;
  }
  static SDB_CX int n_frags() { return p3_off(NB); }            // per-lane doubles
  static SDB_CX int left_off(int b) {                            // leftover-row table offsets
    int o = 0;
    for (int q = 0; q < b; ++q) o += IL * rn(q) * M;
    return o;
  }
  static SDB_CX int n_left() { return left_off(NB); }
  static SDB_CX int p1_tiles(int b) { return p1_off(b) / (LT > 0 ? LT : 1); }  // row tiles before block b
  static SDB_CX int n_ks() { return p1_tiles(NB) * KS * 8; }  // columns 4 LT.. of the P1 matrix, by tile row
  static SDB_CX int warp_doubles() { return (grows() + (SDB_FUSED_DW_REGS ? 0 : M)) * 32; }
  static SDB_CX size_t smem_base() {
    return ((size_t)MM * SDB_FUSED_THREADS + (size_t)warp_doubles() * SDB_FUSED_WARPS) * sizeof(double) +
           (SDB_FUSED_MAX_TERMS + 1 + n_left() + 1) * sizeof(double);
  }
  // copies of the log table: 8 (conflict-free, see sdb_normal_batch) when an SM's shared memory has room
  static SDB_CX int tab_copies() { return smem_base() + 8 * 128 * 16 <= 227u * 1024u ? 8 : 1; }
  static SDB_CX size_t smem_bytes() {
    const size_t b = smem_base() + (size_t)tab_copies() * 128 * 16 + (SDB_FUSED_AHEAD ? 16 : 0);
#ifdef SDB_FUSED_MIN_SMEM  // diagnostic builds with fewer threads: keep one block per SM
    return b < (size_t)SDB_FUSED_MIN_SMEM ? (size_t)SDB_FUSED_MIN_SMEM : b;
#else
    return b;
#endif
  }
  // SDB_FUSED_AHEAD: 128-bit blocks per step that wait in tensor memory (dW + the Levy series)
  static constexpr int AHEAD_SLOTS = M / 2 + SDB_FUSED_AHEAD_TERMS * M;
};

#ifndef __CUDACC_RTC__  // host functions are not allowed in NVRTC; sdb_fused_rt.h serves JIT shapes
// Host: lane-major fragment tables.  out[frag * 32 + lane]; then the leftover rows.
template <int D, int M, int XR = 0>
inline void sdb_fused_build_tables(const double* A, const double* B, double* out) {
  typedef SdbFusedShape<D, M, XR> Sh;
  double A2[D * D];
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < D; ++j) {
      double acc = 0.0;
      for (int l = 0; l < D && XR >= 0; ++l) acc += A[i * D + l] * A[l * D + j];
      A2[i * D + j] = acc;
    }
  for (int b = 0; b < Sh::NB; ++b) {
    const int rn = Sh::rn(b), r0 = Sh::r0(b);
    for (int t = 0; t < Sh::t1(b); ++t)
      for (int lt = 0; lt < Sh::LT; ++lt)
        for (int lane = 0; lane < 32; ++lane) {
          const int n = 8 * t + (lane >> 2), l = 4 * lt + (lane & 3);
          double v = 0.0;
          if (l < D) {
            if (n < M * rn) v = B[((n / rn) * D + r0 + n % rn) * D + l];       // B_j[r0+r][l]
            else if (XR < 0) v = 0.0;                                          // no drift rows (SDB_FUSED_A2 -1)
            else if (n < (M + 1) * rn) v = A[(r0 + n - M * rn) * D + l];       // drift rows
            else if (n < Sh::nr(b)) v = A2[(r0 + n - (M + 1) * rn) * D + l];   // rows of A^2 (XR)
          }
          out[(size_t)(Sh::p1_off(b) + t * Sh::LT + lt) * 32 + lane] = v;
        }
    for (int kt = 0; kt < Sh::kt3(b); ++kt)
      for (int it = 0; it < Sh::IT; ++it)
        for (int lane = 0; lane < 32; ++lane) {
          const int i = 8 * it + (lane >> 2), K = 4 * kt + (lane & 3);
          double v = 0.0;
          if (K < rn * M) v = B[((K % M) * D + i) * D + r0 + K / M];           // B_k[i][r0+r]
          out[(size_t)(Sh::p3_off(b) + kt * Sh::IT + it) * 32 + lane] = v;
        }
  }
  double* left = out + (size_t)Sh::n_frags() * 32;
  for (int b = 0; b < Sh::NB; ++b)
    for (int q = 0; q < Sh::IL; ++q)
      for (int K = 0; K < Sh::rn(b) * M; ++K)
        left[Sh::left_off(b) + q * Sh::rn(b) * M + K] =
            B[((K % M) * D + 8 * Sh::IT + q) * D + Sh::r0(b) + K / M];
  double* ks = left + Sh::n_left();  // [row tile][q < KS][row in tile]: column 4 LT + q of the P1 matrix
  for (int b = 0; b < Sh::NB; ++b) {
    const int rn = Sh::rn(b), r0 = Sh::r0(b);
    for (int t = 0; t < Sh::t1(b); ++t)
      for (int q = 0; q < Sh::KS; ++q)
        for (int fr = 0; fr < 8; ++fr) {
          const int n = 8 * t + fr, l = 4 * Sh::LT + q;
          double v = 0.0;
          if (n < M * rn) v = B[((n / rn) * D + r0 + n % rn) * D + l];
          else if (n < (M + 1) * rn) v = A[(r0 + n - M * rn) * D + l];
          else if (n < Sh::nr(b)) v = A2[(r0 + n - (M + 1) * rn) * D + l];
          ks[((Sh::p1_tiles(b) + t) * Sh::KS + q) * 8 + fr] = v;
        }
  }
}
template <int D, int M, int XR = 0>
inline size_t sdb_fused_table_doubles() {
  typedef SdbFusedShape<D, M, XR> Sh;
  return (size_t)Sh::n_frags() * 32 + Sh::n_left() + Sh::n_ks();
}
#endif  // !__CUDACC_RTC__

#ifdef __CUDACC__
SDB_DEV void sdb_dmma(double& c0, double& c1, double a, double b) {
#if SDB_FUSED_VOL_MMA
  asm volatile(
#else
  asm(
#endif
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// column swizzle of the staging buffer: keeps every access pattern used below conflict-free
SDB_DEV constexpr int sdb_swz(int row) { return (row & 1) | (((row >> 1) & 1) << 3); }
// Rows that are written one lane per path and read back as DMMA B-fragments (P0: the state, P3: the
// stage-2 columns): lane (fr, fc) reads row 4 q + fc, column 8 pt + fr, so the four rows of one load
// must land in four different 4-column groups -- sdb_swz maps two of them onto the same banks (2-way
// conflict on every P3 / P0 load; profiles/r02_fused_v0_mx2_ncu_summary.txt).
SDB_DEV constexpr int sdb_swz3(int row) { return (row & 3) << 2; }

template <int M>
SDB_DEV constexpr int sdb_pair_index(int i, int j) {  // (i < j) in K_m row order
  return i * M - i * (i + 1) / 2 + (j - i - 1);
}


// ---- lock-step generation of NB Box-Muller pairs (same arithmetic as sdb_normal_pair, written
// stage by stage over the batch so that NB independent dependency chains are adjacent) ----------
template <int NB>
SDB_DEV void sdb_philox_batch(const SdbPhiloxKey& key, uint32_t p_lo, uint32_t p_hi, uint32_t step,
                              uint32_t slot0, uint32_t (&w)[NB][4]) {
  uint32_t c0[NB], c1[NB], c2[NB], c3[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    c0[b] = p_lo; c1[b] = p_hi; c2[b] = step; c3[b] = slot0 + (uint32_t)b;
  }
#pragma unroll
  for (int r = 0; r < 10; ++r) {
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const uint64_t p0 = (uint64_t)SDB_PHILOX_M0 * c0[b];
      const uint64_t p1 = (uint64_t)SDB_PHILOX_M1 * c2[b];
      const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1[b] ^ key.a[r];
      const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3[b] ^ key.b[r];
      c1[b] = (uint32_t)p1; c3[b] = (uint32_t)p0; c0[b] = n0; c2[b] = n2;
    }
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) { w[b][0] = c0[b]; w[b][1] = c1[b]; w[b][2] = c2[b]; w[b][3] = c3[b]; }
}

#ifndef SDB_FUSED_RNG
#define SDB_FUSED_RNG 1  // 1: slots >= 1 by the MX2 expansion of the step's key block; 0: Philox per slot
#endif
// The random blocks of one (path, step): slot 0 and the expansion key are the two Philox blocks.
struct SdbGen {
  const SdbPhiloxKey& key;
  uint32_t p_lo, p_hi, step;
#if SDB_FUSED_RNG
  uint32_t B0[4], K[4];
#endif
  // expand = false: only slot 0 will be read (m <= 2 Wiener processes, dW only)
  SDB_DEV SdbGen(const SdbPhiloxKey& key_, uint32_t p_lo_, uint32_t p_hi_, uint32_t step_,
                 bool expand = true)
      : key(key_), p_lo(p_lo_), p_hi(p_hi_), step(step_) {
#if SDB_FUSED_RNG
    if (!expand) {
      sdb_philox4x32_10(p_lo, p_hi, step, 0u, key, B0);
      K[0] = K[1] = K[2] = K[3] = 0u;
      return;
    }
    uint32_t c0[2], c1[2], c2[2], c3[2];
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      c0[b] = p_lo; c1[b] = p_hi; c2[b] = step; c3[b] = b == 0 ? 0u : SDB_RNG_KEYSLOT;
    }
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const uint64_t p0 = (uint64_t)SDB_PHILOX_M0 * c0[b];
        const uint64_t p1 = (uint64_t)SDB_PHILOX_M1 * c2[b];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1[b] ^ key.a[r];
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3[b] ^ key.b[r];
        c1[b] = (uint32_t)p1; c3[b] = (uint32_t)p0; c0[b] = n0; c2[b] = n2;
      }
    }
    B0[0] = c0[0]; B0[1] = c1[0]; B0[2] = c2[0]; B0[3] = c3[0];
    K[0] = c0[1]; K[1] = c1[1]; K[2] = c2[1]; K[3] = c3[1];
#endif
  }
  // blocks of slots slot0 .. slot0 + NB - 1; FIRST: slot0 == 0 (compile-time knowledge)
  template <int NB, bool FIRST>
  SDB_DEV void blocks(uint32_t slot0, uint32_t (&w)[NB][4]) const {
#if SDB_FUSED_RNG
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      if (FIRST && b == 0) {
        w[0][0] = B0[0]; w[0][1] = B0[1]; w[0][2] = B0[2]; w[0][3] = B0[3];
      } else {
        sdb_mx2_block(K, slot0 + (uint32_t)b, w[b]);
      }
    }
#else
    sdb_philox_batch<NB>(key, p_lo, p_hi, step, slot0, w);
#endif
  }
};

// z[2b], z[2b+1] = scale * (normal pair b); SCALED = false skips the multiplication.
// TS: stride (in entries) of the log table -- 1 for the plain 128-entry table, 8 when every lane
// position of a quarter-warp has its own copy in its own four banks (tab then points at this lane's
// copy): the random index makes the plain table a 2.5-way bank conflict per LDS.128.
template <int NB, bool SCALED, int TS = 1>
SDB_DEV void sdb_normal_batch(const uint32_t (&w)[NB][4], const double2* __restrict__ tab,
                              double scale, double (&z)[2 * NB]) {
  double x[NB], zz[NB], r[NB], q[NB], L[NB], t[NB], t2[NB], sn[NB], cs[NB], rad[NB];
  double2 tc[NB];
  int kk[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    x[b] = __hiloint2double((int)(0x3FF00000u | (w[b][0] >> 12)), (int)w[b][1]) - SDB_ONE_M53;
    t[b] = __hiloint2double((int)(0x3FF00000u | ((w[b][2] >> 9) & 0x000FFFFFu)), (int)w[b][3]) - 1.0;
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const int xh = __double2hiint(x[b]);
    const int tmp = xh - 0x3FE60000;
    kk[b] = tmp >> 20;
    zz[b] = __hiloint2double(xh - (tmp & (int)0xFFF00000), __double2loint(x[b]));
    tc[b] = tab[((tmp >> 13) & 127) * TS];
  }
#if SDB_FUSED_BATCH_ORDER == 0
#pragma unroll
  for (int b = 0; b < NB; ++b) { r[b] = fma(zz[b], tc[b].x, -1.0); t2[b] = t[b] * t[b]; }
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    q[b] = fma(r[b], 1.0 / 3.0, -0.4);
    L[b] = fma((double)kk[b], SDB_M2LN2, tc[b].y);
    sn[b] = fma(t2[b], SDB_SIN6, SDB_SIN5);
    cs[b] = fma(t2[b], SDB_COS6, SDB_COS5);
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    q[b] = fma(r[b], q[b], 0.5);
    sn[b] = fma(t2[b], sn[b], SDB_SIN4);
    cs[b] = fma(t2[b], cs[b], SDB_COS4);
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    q[b] = fma(r[b], q[b], -2.0 / 3.0);
    sn[b] = fma(t2[b], sn[b], SDB_SIN3);
    cs[b] = fma(t2[b], cs[b], SDB_COS3);
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    q[b] = fma(r[b], q[b], 1.0);
    sn[b] = fma(t2[b], sn[b], SDB_SIN2);
    cs[b] = fma(t2[b], cs[b], SDB_COS2);
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    q[b] = fma(r[b], q[b], -2.0);
    sn[b] = fma(t2[b], sn[b], SDB_SIN1);
    cs[b] = fma(t2[b], cs[b], SDB_COS1);
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) L[b] = fma(r[b], q[b], L[b]);  // -2 ln u1 > 0 (sdb_normal_pair)
#else
  // the angle polynomials first: they do not need the table entry, so they cover its latency
#pragma unroll
  for (int b = 0; b < NB; ++b) t2[b] = t[b] * t[b];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    sn[b] = fma(t2[b], SDB_SIN6, SDB_SIN5);
    cs[b] = fma(t2[b], SDB_COS6, SDB_COS5);
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    sn[b] = fma(t2[b], sn[b], SDB_SIN4);
    cs[b] = fma(t2[b], cs[b], SDB_COS4);
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    sn[b] = fma(t2[b], sn[b], SDB_SIN3);
    cs[b] = fma(t2[b], cs[b], SDB_COS3);
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    sn[b] = fma(t2[b], sn[b], SDB_SIN2);
    cs[b] = fma(t2[b], cs[b], SDB_COS2);
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    sn[b] = fma(t2[b], sn[b], SDB_SIN1);
    cs[b] = fma(t2[b], cs[b], SDB_COS1);
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    r[b] = fma(zz[b], tc[b].x, -1.0);
    L[b] = fma((double)kk[b], SDB_M2LN2, tc[b].y);
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) q[b] = fma(r[b], 1.0 / 3.0, -0.4);
#pragma unroll
  for (int b = 0; b < NB; ++b) q[b] = fma(r[b], q[b], 0.5);
#pragma unroll
  for (int b = 0; b < NB; ++b) q[b] = fma(r[b], q[b], -2.0 / 3.0);
#pragma unroll
  for (int b = 0; b < NB; ++b) q[b] = fma(r[b], q[b], 1.0);
#pragma unroll
  for (int b = 0; b < NB; ++b) q[b] = fma(r[b], q[b], -2.0);
#pragma unroll
  for (int b = 0; b < NB; ++b) L[b] = fma(r[b], q[b], L[b]);  // -2 ln u1 > 0
#endif
  // sqrt(L) by three terms of g / sqrt(1 - e) (see sdb_normal_pair)
  double y[NB], g[NB], ge[NB], e[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    y[b] = sdb_rsqrt_approx(L[b]);
    sn[b] = fma(t2[b], sn[b], SDB_SIN0);
    cs[b] = fma(t2[b], cs[b], SDB_COS0);
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    g[b] = L[b] * y[b];
    sn[b] *= t[b];
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) e[b] = fma(-g[b], y[b], 1.0);
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    ge[b] = g[b] * e[b];
    e[b] = fma(e[b], 0.375, 0.5);
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) rad[b] = fma(ge[b], e[b], g[b]);
  if (SCALED) {
#pragma unroll
    for (int b = 0; b < NB; ++b) rad[b] *= scale;
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const bool swap = (w[b][2] >> 29) & 1u;
    const double a0 = swap ? sn[b] : cs[b];
    const double b0 = swap ? cs[b] : sn[b];
    const uint32_t s0 = (w[b][2] << 1) & 0x80000000u;
    const uint32_t s1 = w[b][2] & 0x80000000u;
    const double p0 = rad[b] * a0, p1 = rad[b] * b0;
    z[2 * b] = __hiloint2double(__double2hiint(p0) ^ (int)s0, __double2loint(p0));
    z[2 * b + 1] = __hiloint2double(__double2hiint(p1) ^ (int)s1, __double2loint(p1));
  }
}

// Compile-time loop and the inverse of sdb_pair_index.
template <int I>
struct SdbInt { static constexpr int value = I; };
// M normals of one (path, step) from normal index 2 slot0 + ODD on: Box-Muller pairs are addressed by
// slot = normal index >> 1 (element = index & 1), so a block of normals that starts at an odd index
// skips the first element of its first pair.  Used for an ODD number of Wiener processes, where the
// blocks dW | X_1 | Y_1 | X_2 | ... of the reference's draw order alternate between even and odd starts.
template <int M, int ODD, bool SCALED, bool FIRST = false, int TS = 1>
SDB_DEV void sdb_normals_m(const SdbGen& gen, uint32_t slot0, const double2* __restrict__ tab, double scale,
                           double (&out)[M]) {
  constexpr int NP = (M + ODD + 1) / 2;
  uint32_t w[NP][4];
  double z[2 * NP];
  gen.template blocks<NP, FIRST>(slot0, w);
  sdb_normal_batch<NP, SCALED, TS>(w, tab, scale, z);
#pragma unroll
  for (int j = 0; j < M; ++j) out[j] = z[j + ODD];
}

template <int I, int N, class F>
SDB_DEV void sdb_static_for(F&& f) {
  if constexpr (I < N) {
    f(SdbInt<I>{});
    sdb_static_for<I + 1, N>(f);
  }
}
template <int M>
SDB_CX int sdb_pair_j(int idx) {
  int j = 0, base = 0;
  while (idx >= base + (M - 1 - j)) {
    base += M - 1 - j;
    ++j;
  }
  return j;
}
template <int M>
SDB_CX int sdb_pair_k(int idx) {
  const int j = sdb_pair_j<M>(idx);
  return idx - (j * M - j * (j + 1) / 2) + j + 1;
}

#if SDB_FUSED_AHEAD
// Blocks slot0 .. slot0 + NB - 1 of the current step from tensor memory (written one step earlier).
template <int NB>
SDB_DEV void sdb_ahead_load(uint32_t tc, uint32_t slot0, uint32_t (&w)[NB][4]) {
  uint32_t r[4 * NB];
  sdb_static_for<0, NB>([&](auto bb) {
    constexpr int b = decltype(bb)::value;
    SdbTm<4>::template ld<4 * b, 4 * NB>(tc + 4u * (slot0 + (uint32_t)b), r);
  });
  sdb_tm_wait_ld();
  sdb_static_for<0, NB>([&](auto bb) {
    constexpr int b = decltype(bb)::value;
    SdbTm<4>::template touch<4 * b, 4 * NB>(r);
  });
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    w[b][0] = r[4 * b]; w[b][1] = r[4 * b + 1]; w[b][2] = r[4 * b + 2]; w[b][3] = r[4 * b + 3];
  }
}
// The share of the next step's blocks that is produced at position (row block rows r0.., column j) of the
// row-block pipeline: proportional to the FP64 work of that position.
template <int D, int M, int NS>
SDB_DEV void sdb_ahead_emit(const uint32_t (&K)[4], uint32_t tc, int ns, int r0, int rn, int j) {
  const int lo = NS * (M * r0 + j * rn) / (M * D), hi = NS * (M * r0 + (j + 1) * rn) / (M * D);
#pragma unroll
  for (int sl = lo; sl < hi; ++sl)
    if (sl >= 1) {  // no test against the run-time count ns: a branch here would fence the blocks off from
      uint32_t w[4];  // the FP64 instructions they are meant to be scheduled between
      sdb_mx2_block(K, (uint32_t)sl, w);
      SdbTm<4>::template st<0, 4>(tc + 4u * (uint32_t)sl, w);
    }
}
#endif

template <int D, int M, int B_>
struct SdbFusedBlock {
  typedef SdbFusedShape<D, M, SDB_FUSED_A2> Sh;
  static constexpr int RN = Sh::rn(B_), R0 = Sh::r0(B_), T1 = Sh::t1(B_), KT3 = Sh::kt3(B_);

  // P1 + P2 + P3 for row block B_.
  static SDB_DEV void run(const double* __restrict__ frag, const double* __restrict__ s_left,
                          double* gbuf, SDB_SMEM_CV double* sA, SDB_SMEM_CV double* sW,
                          const double (&Yf)[Sh::LT][4], const double (&Yx)[Sh::KS > 0 ? Sh::KS : 1][4][2],
                          double h, double dg, int lane,
                          double (&H20)[D], double (&Yacc)[D], double (&w)[4][2],
                          double (&wv)[Sh::IL > 0 ? Sh::IL : 1], const double (&dWr)[M],
                          const uint32_t (&aK)[4], uint32_t tc, int ahead_ns) {
    constexpr int LT = Sh::LT;
    const int fr = lane >> 2, fc = lane & 3;
    // ---- P1: staging[n][p] = sum_l Bmat[n][l] Y[l][p] -----------------------------------------
#if SDB_FUSED_INPHASE
    SDB_PAIR_BAR();
#endif
    const double* f1 = frag + (size_t)Sh::p1_off(B_) * 32 + lane;
#pragma unroll
    for (int t = 0; t < T1; ++t) {
      double a[LT];
#pragma unroll
      for (int lt = 0; lt < LT; ++lt) a[lt] = __ldg(f1 + (t * LT + lt) * 32);
      double c[4][2];
      if constexpr (Sh::KS > 0) {
        // columns 4 LT.. of the matrix: this lane's row (8 t + fr) times the state components 4 LT + q of
        // the eight paths its accumulator fragments belong to
        const double* kst = frag + (size_t)Sh::n_frags() * 32 + Sh::n_left() +
                            (size_t)(Sh::p1_tiles(B_) + t) * Sh::KS * 8 + fr;
        const double k0 = __ldg(kst);
#pragma unroll
        for (int pt = 0; pt < 4; ++pt) {
          c[pt][0] = k0 * Yx[0][pt][0];
          c[pt][1] = k0 * Yx[0][pt][1];
        }
#pragma unroll
        for (int q = 1; q < Sh::KS; ++q) {
          const double kq = __ldg(kst + q * 8);
#pragma unroll
          for (int pt = 0; pt < 4; ++pt) {
            c[pt][0] = fma(kq, Yx[q][pt][0], c[pt][0]);
            c[pt][1] = fma(kq, Yx[q][pt][1], c[pt][1]);
          }
        }
      } else {
#pragma unroll
        for (int pt = 0; pt < 4; ++pt) {
          c[pt][0] = 0.0;
          c[pt][1] = 0.0;
        }
      }
#pragma unroll
      for (int lt = 0; lt < LT; ++lt)  // consecutive DMMAs hit independent accumulators
#pragma unroll
        for (int pt = 0; pt < 4; ++pt) sdb_dmma(c[pt][0], c[pt][1], a[lt], Yf[lt][pt]);
      double* row = gbuf + (8 * t + fr) * 32;
      const int sw = sdb_swz(fr);  // (8t + fr) & 7 == fr
#pragma unroll
      for (int pt = 0; pt < 4; ++pt) {
        row[(8 * pt + 2 * fc) ^ sw] = c[pt][0];
        row[(8 * pt + 2 * fc + 1) ^ sw] = c[pt][1];
      }
    }
    __syncwarp();
    // ---- P2: one lane per path ----------------------------------------------------------------
    {
      constexpr int NT = SDB_FUSED_THREADS;
      SDB_SMEM_CV double* g = gbuf;
      double s[RN][M], GdW[RN];
#pragma unroll
      for (int r = 0; r < RN; ++r) {
        GdW[r] = 0.0;
#pragma unroll
        for (int k = 0; k < M; ++k) s[r][k] = 0.0;
      }
#pragma unroll
      for (int j = 0; j < M; ++j) {
        double gj[RN];
        const double dWj = SDB_FUSED_DW_REGS ? dWr[j] : sW[j * 32];
#pragma unroll
        for (int r = 0; r < RN; ++r) {
          const int n = j * RN + r;
          gj[r] = g[n * 32 + (lane ^ sdb_swz(n))];
#if SDB_FUSED_AFFINE
          gj[r] += __ldg(frag + (size_t)Sh::n_frags() * 32 + Sh::n_left() + Sh::n_ks() + j * D + R0 + r);
#endif
          GdW[r] = fma(gj[r], dWj, GdW[r]);
        }
#pragma unroll
        for (int k = 0; k < M; ++k) {
          double a;
          if (k == j) {
            a = dg;
          } else if (j < k) {
            a = sA[sdb_pair_index<M>(j, k) * NT];
          } else {
            a = -sA[sdb_pair_index<M>(k, j) * NT];
          }
#pragma unroll
          for (int r = 0; r < RN; ++r) s[r][k] = fma(gj[r], a, s[r][k]);
        }
#if SDB_FUSED_AHEAD
        sdb_ahead_emit<D, M, Sh::AHEAD_SLOTS>(aK, tc, ahead_ns, R0, RN, j);
#endif
      }
#pragma unroll
      for (int k = 0; k < M; ++k) {
        const double hd = 0.5 * (SDB_FUSED_DW_REGS ? dWr[k] : sW[k * 32]);
#pragma unroll
        for (int r = 0; r < RN; ++r) s[r][k] = fma(hd, GdW[r], s[r][k]);
      }
#if SDB_FUSED_A2 < 0
      // the drift was evaluated before the row blocks: H20 = Y + f0 h, Yacc = -f0 h / 2 so far
#pragma unroll
      for (int r = 0; r < RN; ++r) Yacc[R0 + r] += GdW[r];
#else
      // drift rows of this block: H20 = Y + f h (integrate.py:509), update without the f1 term
#pragma unroll
      for (int r = 0; r < RN; ++r) {
        const int n = M * RN + r;
        const double f0h = g[n * 32 + (lane ^ sdb_swz(n))] * h;
#if SDB_FUSED_A2
        // f(H20) h = f0 h + h^2 A^2 Y (linear drift): Y_{n+1} - H20 = h^2 A^2 Y / 2 + G dW + stage-2 sum
        const int n2 = (M + 1) * RN + r;
        Yacc[R0 + r] = fma(0.5 * h * h, g[n2 * 32 + (lane ^ sdb_swz(n2))], GdW[r]);
#else
        Yacc[R0 + r] = fma(-0.5, f0h, GdW[r]);
#endif
        H20[R0 + r] += f0h;  // H20 held Y
      }
#endif
      // rows 8.. of the stage-2 sum on the FMA pipe (uniform operands from shared memory)
      if (Sh::IL > 0) {
#pragma unroll
        for (int q = 0; q < Sh::IL; ++q) {
          const double* bl = s_left + Sh::left_off(B_) + q * RN * M;
          double acc = wv[q];
#pragma unroll
          for (int r = 0; r < RN; ++r)
#pragma unroll
            for (int k = 0; k < M; ++k) acc = fma(bl[r * M + k], s[r][k], acc);
          wv[q] = acc;
        }
      }
      __syncwarp();  // every lane has read its G_n entries: the staging rows may be overwritten
#pragma unroll
      for (int r = 0; r < RN; ++r)
#pragma unroll
        for (int k = 0; k < M; ++k) {
          const int K = r * M + k;
          gbuf[K * 32 + (lane ^ sdb_swz3(K))] = s[r][k];
        }
    }
    __syncwarp();
    // ---- P3: w[i][p] += sum_K Bst[i][K] s~[K][p], rows i < 8 IT ------------------------------------
#if SDB_FUSED_INPHASE
    SDB_PAIR_BAR();
#endif
    if (Sh::IT > 0) {
      const double* f3 = frag + (size_t)Sh::p3_off(B_) * 32 + lane;
#pragma unroll
      for (int kt = 0; kt < KT3; ++kt) {
        const double a = __ldg(f3 + kt * Sh::IT * 32);
        const int K = 4 * kt + fc;
        const double* row = gbuf + K * 32;
        const int sw = sdb_swz3(K);
#pragma unroll
        for (int pt = 0; pt < 4; ++pt) sdb_dmma(w[pt][0], w[pt][1], a, row[(8 * pt + fr) ^ sw]);
      }
    }
    __syncwarp();
  }
};

template <int D, int M, int B_>
struct SdbFusedBlocks {
  template <class... Args>
  static SDB_DEV void run(Args&... args) {
    SdbFusedBlocks<D, M, B_ - 1>::run(args...);
    SdbFusedBlock<D, M, B_>::run(args...);
  }
};
template <int D, int M>
struct SdbFusedBlocks<D, M, -1> {
  template <class... Args>
  static SDB_DEV void run(Args&...) {}
};

template <int D, int M, int WIK = 0, class FP = PInline<D * D>>
SDB_DEV void sdb_srk2_fused_linear_body(const SdbLoopArgs& a, const FP& fp,
                                        const double* __restrict__ frag, double* smem) {
  typedef SdbFusedShape<D, M, SDB_FUSED_A2> Sh;
  constexpr int NT = SDB_FUSED_THREADS;
  constexpr int MM = Sh::MM, LT = Sh::LT, GR = Sh::grows();
  static_assert(M % 2 == 0 || (SDB_FUSED_BATCH && !SDB_FUSED_BATCH_XZ && WIK == 0),
                "an odd number of Wiener processes needs the batched generator (and Ikpw/Jkpw)");
  static_assert(Sh::IT <= 1, "stage-2 DMMA tiles: d < 16 supported");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // shared: Levy areas [MM][NT] | per warp { staging [GR][32], dW [M][32] } | tables
  double* warp_base = smem + (size_t)MM * NT + (size_t)warp * Sh::warp_doubles();
  double* gbuf = warp_base;
  double2* tab =
      reinterpret_cast<double2*>(smem + (size_t)MM * NT + (size_t)Sh::warp_doubles() * SDB_FUSED_WARPS);
  constexpr int TS = SDB_FUSED_BATCH ? Sh::tab_copies() : 1;  // (the pair-at-a-time variant reads a plain table)
  double* s_hk = reinterpret_cast<double*>(tab + 128 * TS);
  double* s_left = s_hk + SDB_FUSED_MAX_TERMS + 1;
  for (int i = tid; i < 128 * TS; i += NT) tab[i] = reinterpret_cast<const double2*>(sdb_logtab)[i / TS];
  tab += lane & (TS - 1);  // this lane's copy: entry i at tab[i * TS]
  for (int i = tid; i <= SDB_FUSED_MAX_TERMS; i += NT)
    s_hk[i] = i == 0 ? 0.0 : a.h_global * SDB_INV_2PI / (double)i;
  for (int i = tid; i < Sh::n_left(); i += NT) s_left[i] = frag[(size_t)Sh::n_frags() * 32 + i];
  __syncthreads();
#if SDB_FUSED_AHEAD
  static_assert(M % 2 == 0 && WIK == 0 && SDB_FUSED_BATCH && !SDB_FUSED_SWP && !SDB_FUSED_BATCH_XZ,
                "SDB_FUSED_AHEAD: even m, Ikpw/Jkpw, batched generator");
  static_assert(4 * Sh::AHEAD_SLOTS <= 256 && NT == 256, "two warps share the 512 columns of a TMEM lane quarter");
  // warps w and w + 4 run on the same sub-partition and see the same 32 TMEM lanes: 256 columns each
  uint32_t* tm_slot = reinterpret_cast<uint32_t*>(s_left + Sh::n_left() + 1);
  const uint32_t tm_base = sdb_tm_alloc512(tm_slot);
  const uint32_t tc = tm_base + (warp >= 4 ? 256u : 0u);
  const int ahead_ns = M / 2 + a.n_terms * M;  // blocks per step in use (n_terms <= SDB_FUSED_AHEAD_TERMS)
#else
  const uint32_t tc = 0u;
  const int ahead_ns = 0;
#endif

  const int64_t p = (int64_t)blockIdx.x * NT + tid;
  const bool active = p < a.P;  // idle lanes of the last warp still take part in the DMMAs
  double* sA = smem + tid;
  double* sW = warp_base + GR * 32 + lane;

  double Y[D];
#pragma unroll
  for (int i = 0; i < D; ++i) Y[i] = 0.0;
  if (active) {
    if (a.flags & SDB_FLAG_Y0_BROADCAST) {
#pragma unroll
      for (int i = 0; i < D; ++i) Y[i] = a.y0[i];
    } else {
#pragma unroll
      for (int i = 0; i < D; ++i) Y[i] = a.y0[i * a.y0_sI + p * a.y0_sP];
    }
    sdb_store<D>(a, p, 0, Y);
  }
#if SDB_OBSERVE
  SdbObs<D> obs;
  obs.init(a, Y);
#endif

  const double rh_g = sqrt(a.h_global);
  const double cz = sqrt(2.0 / a.h_global);
  const double dg = (a.flags & SDB_FLAG_STRATONOVICH) ? 0.0 : -0.5 * a.h_global;
  const uint64_t path = (uint64_t)(a.path_id0 + p);
  const SdbPhiloxKey pkey(a.seed);
  const int steps = a.N - 1;
  const int n_terms = a.n_terms;
  const double* __restrict__ hs = a.tspan + a.N;  // per-step h (host-built)
  int next_save = a.save_every;
  int64_t slot = 1;
  const int fr = lane >> 2, fc = lane & 3;

#if SDB_FUSED_AHEAD
  {  // the blocks of the first step
    const SdbGen g0(pkey, (uint32_t)path, (uint32_t)(path >> 32), SDB_STEP(0));
    SdbTm<4>::template st<0, 4>(tc, g0.B0);
#pragma unroll 1
    for (int sl = 1; sl < ahead_ns; ++sl) {
      uint32_t w4[4];
      sdb_mx2_block(g0.K, (uint32_t)sl, w4);
      SdbTm<4>::template st<0, 4>(tc + 4u * (uint32_t)sl, w4);
    }
    sdb_tm_wait_st();
  }
#endif
#pragma unroll 1
  for (int n = 0; n < steps; ++n) {
    const double h = hs[n];
#if SDB_FUSED_ANTIPHASE
    if (warp < 4) asm volatile("bar.sync %0, 64;" ::"r"(1 + (warp & 3)) : "memory");
#endif
#if SDB_FUSED_INPHASE
    SDB_PAIR_BAR();
#endif

#if SDB_FUSED_PARK_Y
    // The state waits in the staging buffer (where P0 needs it anyway) while the noise is generated:
    // 20 registers less across the register-hungry Levy loop.
#pragma unroll
    for (int l = 0; l < D; ++l) gbuf[l * 32 + (lane ^ sdb_swz3(l))] = Y[l];
    __syncwarp();
#endif
    // ---- noise: dW (wiener.py:52), Levy areas (wiener.py:67-74, :102-105), one lane per path ------
    double dW[M];  // lives through the row blocks only with SDB_FUSED_DW_REGS
    {
      const SdbStream st(pkey, path, SDB_STEP(n), tab);
      const SdbGen gen(pkey, st.p_lo, st.p_hi, SDB_STEP(n));  // (unused, hence removed, with SDB_FUSED_AHEAD)
#if SDB_FUSED_AHEAD
      {
        uint32_t w[M / 2][4];
        sdb_ahead_load<M / 2>(tc, 0u, w);
        sdb_normal_batch<M / 2, true, TS>(w, tab, rh_g, dW);
      }
#elif SDB_FUSED_BATCH
      if constexpr (M % 2 == 0) {
        uint32_t w[M / 2][4];
        gen.template blocks<M / 2, true>(0u, w);
        sdb_normal_batch<M / 2, true, TS>(w, tab, rh_g, dW);
      } else {
        sdb_normals_m<M, 0, true, true, TS>(gen, 0u, tab, rh_g, dW);
      }
#else
#pragma unroll
      for (int s = 0; s < M / 2; ++s) {
        double z0, z1;
        st.pair((uint32_t)s, z0, z1);
        dW[2 * s] = rh_g * z0;
        dW[2 * s + 1] = rh_g * z1;
      }
#endif
#if SDB_FUSED_PARK_DW
#pragma unroll
      for (int j = 0; j < M; ++j) sW[j * 32] = dW[j];
#endif
      constexpr int KU = SDB_FUSED_KUNROLL;
      double At[MM];
#pragma unroll
      for (int r = 0; r < MM; ++r) At[r] = 0.0;
#if SDB_FUSED_BATCH && SDB_FUSED_SWP && !SDB_FUSED_BATCH_XZ
      constexpr bool SWP = M % 2 == 0;
      uint32_t wA[SWP ? M / 2 : 1][4];   // bits of the next X batch, one batch ahead of their transform
      if constexpr (SWP) gen.template blocks<M / 2, false>((uint32_t)(M >> 1), wA);
#else
      constexpr bool SWP = false;
#endif
#pragma unroll(KU)
      for (int k = 1; k <= n_terms; ++k) {
        double X[M], Z[M];
        const uint32_t sx = (uint32_t)((M + 2 * (k - 1) * M) >> 1);
        const double hk = s_hk[k];
#if SDB_FUSED_BATCH
#if SDB_FUSED_BATCH_XZ
        {
          uint32_t w[M][4];
          double xz[2 * M];
          gen.template blocks<M, false>(sx, w);
          sdb_normal_batch<M, false, TS>(w, tab, 1.0, xz);
#pragma unroll
          for (int j = 0; j < M; ++j) {
            X[j] = hk * xz[j];
            Z[j] = xz[M + j];
          }
        }
#else
        if constexpr (SWP) {
#if SDB_FUSED_SWP
          uint32_t wB[M / 2][4];
          gen.template blocks<M / 2, false>(sx + M / 2, wB);       // Y_k bits   | integer pipes
          sdb_normal_batch<M / 2, true, TS>(wA, tab, hk, X);        // X_k        | FP64 pipe
          if (k < n_terms) gen.template blocks<M / 2, false>(sx + M, wA);   // X_{k+1} bits
          sdb_normal_batch<M / 2, false, TS>(wB, tab, 1.0, Z);      // Y_k
#endif
        } else if constexpr (M % 2 == 0) {
          uint32_t w[M / 2][4];
#if SDB_FUSED_AHEAD
          sdb_ahead_load<M / 2>(tc, sx, w);
          sdb_normal_batch<M / 2, true, TS>(w, tab, hk, X);
          sdb_ahead_load<M / 2>(tc, sx + M / 2, w);
          sdb_normal_batch<M / 2, false, TS>(w, tab, 1.0, Z);
#else
          gen.template blocks<M / 2, false>(sx, w);
          sdb_normal_batch<M / 2, true, TS>(w, tab, hk, X);
          gen.template blocks<M / 2, false>(sx + M / 2, w);
          sdb_normal_batch<M / 2, false, TS>(w, tab, 1.0, Z);
#endif
        } else {  // X_k starts at the odd normal index M (2k - 1), Y_k at the even one 2 k M
          sdb_normals_m<M, 1, true, false, TS>(gen, sx, tab, hk, X);
          sdb_normals_m<M, 0, false, false, TS>(gen, sx + (M + 1) / 2, tab, 1.0, Z);
        }
#endif
#pragma unroll
        for (int j = 0; j < M; ++j) {
#if SDB_FUSED_PARK_DW
          Z[j] = fma(cz, ((const volatile double*)sW)[j * 32], Z[j]);
#else
          Z[j] = fma(cz, dW[j], Z[j]);
#endif
        }
#else
#pragma unroll
        for (int s = 0; s < M / 2; ++s) {
          double z0, z1;
          st.pair(sx + s, z0, z1);
          X[2 * s] = hk * z0;
          X[2 * s + 1] = hk * z1;
        }
#pragma unroll
        for (int s = 0; s < M / 2; ++s) {
          double z0, z1;
          st.pair(sx + M / 2 + s, z0, z1);
#if SDB_FUSED_PARK_DW
          Z[2 * s] = fma(cz, ((const volatile double*)sW)[2 * s * 32], z0);
          Z[2 * s + 1] = fma(cz, ((const volatile double*)sW)[(2 * s + 1) * 32], z1);
#else
          Z[2 * s] = fma(cz, dW[2 * s], z0);
          Z[2 * s + 1] = fma(cz, dW[2 * s + 1], z1);
#endif
        }
#endif
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
          for (int j = i + 1; j < M; ++j) {
            const int r = sdb_pair_index<M>(i, j);
            At[r] = fma(X[i], Z[j], At[r]);
            At[r] = fma(-Z[i], X[j], At[r]);
          }
      }
      if constexpr (WIK != 0) {
        // Wiktorsson's tail (wiener.py:269-277) in its O(m^2) form (the arithmetic of
        // sdb_x_levy_areas in sdb_fusedx.cuh), the areas staying in registers:
        //   1. A_ij += c1 g_ij,  u_i += g_ij dW_j,  u_j -= g_ij dW_i   (u = G^ dW; g in batches of 5 pairs)
        //   2. A_ij += c2 (dW_j u_i - dW_i u_j)
        double u[M];
        double nrm = 0.0;
#pragma unroll
        for (int j = 0; j < M; ++j) {
          nrm = fma(dW[j], dW[j], nrm);
          u[j] = 0.0;
        }
        const double hgl = a.h_global;
        const double rad = sqrt(1.0 + nrm / hgl);
        const double den = hgl * SDB_INV_2PI * sqrt(sdb_a_tail(n_terms)) / (1.4142135623730951 * (1.0 + rad));
        const double c1 = den * (2.0 + 2.0 * rad);
        const double c2 = den * (2.0 / hgl);
        const uint32_t st0 = (uint32_t)((M + 2 * n_terms * M) >> 1);
        constexpr int NPAIR = (MM + 1) / 2;  // Box-Muller pairs holding the MM tail normals
        sdb_static_for<0, (NPAIR + 4) / 5>([&](auto bb) {
          constexpr int b = decltype(bb)::value;
          constexpr int NBP = NPAIR - 5 * b >= 5 ? 5 : NPAIR - 5 * b;
          uint32_t w[NBP][4];
          double g[2 * NBP];
          gen.template blocks<NBP, false>(st0 + (uint32_t)(5 * b), w);
          sdb_normal_batch<NBP, false, TS>(w, tab, 1.0, g);
          sdb_static_for<0, 2 * NBP>([&](auto ee) {
            constexpr int e = decltype(ee)::value;
            constexpr int r = 10 * b + e;
            if constexpr (r < MM) {
              constexpr int i = sdb_pair_j<M>(r), j = sdb_pair_k<M>(r);
              At[r] = fma(c1, g[e], At[r]);
              u[i] = fma(g[e], dW[j], u[i]);
              u[j] = fma(-g[e], dW[i], u[j]);
            }
          });
        });
        sdb_static_for<0, MM>([&](auto rr) {
          constexpr int r = decltype(rr)::value;
          constexpr int i = sdb_pair_j<M>(r), j = sdb_pair_k<M>(r);
          // column-major unvec puts +Atilde BELOW the diagonal (wiener.py:278-280): the stage-2 code
          // places +A above it (Ikpw, wiener.py:106), hence the sign
          At[r] = -fma(c2, fma(dW[j], u[i], -dW[i] * u[j]), At[r]);
        });
      }
#pragma unroll
      for (int r = 0; r < MM; ++r) sA[r * NT] = At[r];
#if !SDB_FUSED_PARK_DW && !SDB_FUSED_DW_REGS
#pragma unroll
      for (int j = 0; j < M; ++j) sW[j * 32] = dW[j];
#endif
    }

#if SDB_FUSED_ANTIPHASE
    if (warp >= 4) asm volatile("bar.sync %0, 64;" ::"r"(1 + (warp & 3)) : "memory");
#endif
    uint32_t aK[4] = {0u, 0u, 0u, 0u};
#if SDB_FUSED_AHEAD
    {  // the two Philox blocks of the NEXT step: slot 0 goes to tensor memory, the expansion key stays in
       // registers through the row blocks, which produce the other blocks between their FP64 instructions
      const SdbGen g1(pkey, (uint32_t)path, (uint32_t)(path >> 32), SDB_STEP(n + 1));
      SdbTm<4>::template st<0, 4>(tc, g1.B0);
      aK[0] = g1.K[0]; aK[1] = g1.K[1]; aK[2] = g1.K[2]; aK[3] = g1.K[3];
    }
#endif
    // ---- P0: Y -> DMMA B-fragments Yf[lt][pt] = Y[4 lt + fc][path 8 pt + fr] ---------------------
    double Yf[LT][4];
#if SDB_FUSED_PARK_Y
#pragma unroll
    for (int l = 0; l < D; ++l) Y[l] = gbuf[l * 32 + (lane ^ sdb_swz3(l))];
#else
#pragma unroll
    for (int l = 0; l < D; ++l) gbuf[l * 32 + (lane ^ sdb_swz3(l))] = Y[l];
    __syncwarp();
#endif
#pragma unroll
    for (int lt = 0; lt < LT; ++lt) {
      const int l = 4 * lt + fc;
      const bool ok = (4 * lt + 3 < D) || (l < D);
      const int lc = ok ? l : 0;
#pragma unroll
      for (int pt = 0; pt < 4; ++pt) {
        const double v = gbuf[lc * 32 + ((8 * pt + fr) ^ sdb_swz3(lc))];
        Yf[lt][pt] = ok ? v : 0.0;
      }
    }
    double Yx[Sh::KS > 0 ? Sh::KS : 1][4][2];  // components 4 LT.. of the paths 8 pt + 2 fc, + 1 (accumulator layout)
#pragma unroll
    for (int q = 0; q < Sh::KS; ++q) {
      const int l = 4 * LT + q;
#pragma unroll
      for (int pt = 0; pt < 4; ++pt) {
        const double2 v = *reinterpret_cast<const double2*>(gbuf + l * 32 + ((8 * pt + 2 * fc) ^ sdb_swz3(l)));
        Yx[q][pt][0] = v.x;
        Yx[q][pt][1] = v.y;
      }
    }
    __syncwarp();

    // ---- row blocks ----------------------------------------------------------------------------
    double Yacc[D], w[4][2], wv[Sh::IL > 0 ? Sh::IL : 1];
#pragma unroll
    for (int pt = 0; pt < 4; ++pt) {
      w[pt][0] = 0.0;
      w[pt][1] = 0.0;
    }
#pragma unroll
    for (int q = 0; q < (Sh::IL > 0 ? Sh::IL : 1); ++q) wv[q] = 0.0;
#if SDB_FUSED_A2 < 0
    {  // f(Y_n, t_n) (integrate.py:508): the DMMA fragments of Y_n are already in Yf, so Y may become H20 now
      double f0[D];
      SDB_FUSED_DRIFT_T::eval(fp, Y, a.tspan[n], f0);
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const double f0h = f0[i] * h;
        Yacc[i] = -0.5 * f0h;
        Y[i] += f0h;
      }
    }
#endif
    {
      const double* s_left_c = s_left;
      SDB_SMEM_CV double* sAv = sA;
      SDB_SMEM_CV double* sWv = sW;
      SdbFusedBlocks<D, M, Sh::NB - 1>::run(frag, s_left_c, gbuf, sAv, sWv, Yf, Yx, h, dg, lane, Y,
                                            Yacc, w, wv, dW, aK, tc, ahead_ns);
    }
#if SDB_FUSED_AHEAD
    sdb_tm_wait_st();
#endif
    // Y now holds H20 = Y + f(Y) h

    // ---- second drift evaluation f(H20, t_{n+1}) h (integrate.py:514) ----------------------------
#if SDB_FUSED_A2 == 0
#pragma unroll
    for (int i = 0; i < D; ++i) {
      double acc = fp.v[i * D] * Y[0];
#pragma unroll
      for (int l = 1; l < D; ++l) acc = fma(fp.v[i * D + l], Y[l], acc);
      Yacc[i] = fma(0.5 * h, acc, Yacc[i]);
    }
#elif SDB_FUSED_A2 < 0
    {
      double f1[D];
      SDB_FUSED_DRIFT_T::eval(fp, Y, a.tspan[n + 1], f1);
#pragma unroll
      for (int i = 0; i < D; ++i) Yacc[i] = fma(0.5 * h, f1[i], Yacc[i]);
    }
#endif
    // ---- stage-2 sum back to one lane per path ---------------------------------------------------
    if (Sh::IT > 0) {
      double* row = gbuf + fr * 32;
      const int sw = sdb_swz(fr);
#pragma unroll
      for (int pt = 0; pt < 4; ++pt) {
        row[(8 * pt + 2 * fc) ^ sw] = w[pt][0];
        row[(8 * pt + 2 * fc + 1) ^ sw] = w[pt][1];
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) Yacc[i] += gbuf[i * 32 + (lane ^ sdb_swz(i))];
      __syncwarp();
    }
#pragma unroll
    for (int q = 0; q < Sh::IL; ++q) Yacc[8 * Sh::IT + q] += wv[q];
#pragma unroll
    for (int i = 0; i < D; ++i) Y[i] += Yacc[i];  // Y_{n+1} = H20 - f0h/2 + f1h/2 + G dW + sum
#if SDB_OBSERVE
    obs.step(a, n, Y);
#endif

    if (n + 1 == next_save) {
      if (active) sdb_store<D>(a, p, slot, Y);
      ++slot;
      next_save += a.save_every;
    }
  }
  if (active) sdb_finish<D>(a, p, Y);
#if SDB_OBSERVE
  if (active) obs.finish(a, p);
#endif
#if SDB_FUSED_AHEAD
  sdb_tm_free512(tm_base);
#endif
}

}  // namespace SDB_FUSED_NS
using namespace SDB_FUSED_NS;

#define SDB_FUSED_KERNEL_W(NAME, D, M, WIK)                                                   \
  extern "C" __global__ void __launch_bounds__(SDB_FUSED_THREADS, 1)                          \
      NAME(const __grid_constant__ SdbLoopArgs a, const __grid_constant__ PInline<D * D> fp,  \
           const double* __restrict__ frag) {                                                 \
    extern __shared__ __align__(16) double sdb_fused_smem[];                                  \
    sdb_srk2_fused_linear_body<D, M, WIK>(a, fp, frag, sdb_fused_smem);                       \
  }
#define SDB_FUSED_KERNEL(NAME, D, M) SDB_FUSED_KERNEL_W(NAME, D, M, 0)
// SDB_FUSED_A2 -1: the drift functor's parameter struct takes the place of the matrix A
#define SDB_FUSED_KERNEL_UD(NAME, D, M)                                                       \
  extern "C" __global__ void __launch_bounds__(SDB_FUSED_THREADS, 1)                          \
      NAME(const __grid_constant__ SdbLoopArgs a,                                             \
           const __grid_constant__ SDB_FUSED_DRIFT_T::Params fp, const double* __restrict__ frag) { \
    extern __shared__ __align__(16) double sdb_fused_smem[];                                  \
    sdb_srk2_fused_linear_body<D, M, 0, SDB_FUSED_DRIFT_T::Params>(a, fp, frag, sdb_fused_smem); \
  }
#endif  // __CUDACC__
#ifndef __CUDACC__
}  // namespace SDB_FUSED_NS
using namespace SDB_FUSED_NS;
#endif
