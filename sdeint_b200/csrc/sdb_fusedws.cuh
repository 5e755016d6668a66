// sdb_fusedws.cuh -- warp-specialised version of the fused SRI2/SRS2 kernel (sdb_fused.cuh) for
// LINEAR systems with on-device Philox + Ikpw/Jkpw noise.  Same arithmetic, different machine
// mapping: 16 warps per SM instead of 8.
//
// The 8-warp kernel is limited by the register file: a path needs ~250 live registers in both of
// its phases (noise: 45 Levy accumulators + 20 normals + transform temporaries; integration:
// G_n I accumulators + state), so only 2 warps fit on each SM sub-partition and neither the FP64
// pipe (62 % busy) nor the issue slots (45 %) are saturated.  Here
//   * 8 MAIN warps (32 paths each) do the Levy accumulation and the P1/P2/P3 integration pipeline,
//   * 8 HELPER warps do the stateless part of the noise -- Philox4x32-10 and the fp64 Box-Muller
//     transform -- one helper per main warp, lane l of the helper serving path l of its main warp,
//   * the normals travel helper -> main through TENSOR MEMORY used as a per-lane scratchpad
//     (tcgen05.st / tcgen05.ld, shape 32x32b): helper warp 8+i and main warp i sit on the same SM
//     sub-partition, so they address the same TMEM lanes; a 3-slot ring of 2m normals per lane,
//     handed over with mbarriers.  TMEM (330-680 B/clk/SM, tools/micro/tmem.cu) does not touch the
//     LSU, whose wavefronts the 8-warp kernel also came close to exhausting;
//   * TMEM also holds what the 8-warp kernel kept in shared memory or registers between phases: the
//     Levy areas (pair-indexed, streamed 8 at a time in P2), dW, and the per-row pieces of the update
//     (H20, Yacc).  The state itself waits in the staging rows of shared memory between steps;
//   * registers are re-split with setmaxnreg (main 168, helper 88 of the 128 per thread at launch).
// Row blocks are 2 state rows (5 blocks for d = 10): P2 keeps G_n[2][m] and s[2][m] in registers
// and streams the m(m-1)/2 Levy areas from TMEM: for (j<k): s[:,k] += g[:,j] a_jk, s[:,j] -= g[:,k] a_jk.
#pragma once
#ifndef SDB_FUSED_RB
#define SDB_FUSED_RB 2
#endif
#include "sdb_fused.cuh"
#include "sdb_tmem.cuh"

#define SDB_WS_THREADS 512
#define SDB_WS_MAIN 8                      // main warps (= helper warps)
#define SDB_WS_PATHS (SDB_WS_MAIN * 32)    // paths per block
#ifndef SDB_WS_HELPERS
#define SDB_WS_HELPERS 1                   // 0: main warps generate their own noise (debug / A-B)
#endif
#if SDB_WS_HELPERS
#define SDB_WS_MAIN_REGS 168
#define SDB_WS_HELP_REGS 88
#else
#define SDB_WS_MAIN_REGS 232
#define SDB_WS_HELP_REGS 24
#endif

namespace SDB_FUSED_NS {

template <int D, int M>
struct SdbWsShape {
  typedef SdbFusedShape<D, M> Sh;
  static constexpr int MM = Sh::MM;
  // TMEM columns of one main warp (256 available: two main warps per lane quarter): dW of the
  // current step, then the helper -> main ring of normal batches (X_k: 2M columns, Z_k: 2M columns)
  static constexpr int C_DW = 0;
  static constexpr int C_RING = (2 * M + 7) / 8 * 8;
  static constexpr int SLOT = 4 * M;
  static constexpr int NSLOT = (256 - C_RING) / SLOT;     // 5 for m = 10: almost a whole step ahead
  static_assert(NSLOT >= 2 && NSLOT <= 8, "tensor-memory ring of a main warp");
  // shared-memory rows (32 doubles each) of one main warp
  static constexpr int GR = Sh::grows();                  // staging rows (P1 output / s~)
  static constexpr int R_Y = GR;                          // the state between steps
  static constexpr int R_AT = R_Y + D;                    // Levy areas, pair-indexed
  static constexpr int R_ROW = R_AT + MM;                 // per-row pieces of the update: H20[D], Yacc[D]
  static constexpr int WROWS = R_ROW + 2 * D;
  static SDB_CX size_t smem_bytes() {
    return (size_t)WROWS * 32 * 8 * SDB_WS_MAIN + 128 * 16 +
           (SDB_FUSED_MAX_TERMS + 1 + Sh::n_left() + 1) * sizeof(double) +
           SDB_WS_MAIN * 8 * 2 * 8 + 16;
  }
};

#ifdef __CUDACC__
// ---- mbarrier (shared::cta) --------------------------------------------------------------------
SDB_DEV void sdb_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;"
               :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
SDB_DEV void sdb_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];"
               :: "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
SDB_DEV void sdb_mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}

// Ring position shared by producer and consumer code.
struct SdbRing {
  int slot;
  uint32_t phase;
  SDB_DEV SdbRing() : slot(0), phase(0) {}
  template <int NSLOT>
  SDB_DEV void advance() {
    if (++slot == NSLOT) { slot = 0; phase ^= 1u; }
  }
};

// ---- helper warp: Philox + Box-Muller for the 32 paths of its main warp -------------------------
template <int D, int M>
SDB_DEV void sdb_ws_helper(const SdbLoopArgs& a, const double2* __restrict__ tab,
                           const double* __restrict__ s_hk, uint32_t tm, uint64_t* full,
                           uint64_t* empty, int64_t p, int lane) {
  typedef SdbWsShape<D, M> Ws;
  const double rh_g = sqrt(a.h_global);
  const double cz = sqrt(2.0 / a.h_global);
  const uint64_t path = (uint64_t)(a.path_id0 + p);
  const uint32_t p_lo = (uint32_t)path, p_hi = (uint32_t)(path >> 32);
  const SdbPhiloxKey pkey(a.seed);
  const int steps = a.N - 1;
  const int n_terms = a.n_terms;
  SdbRing ring;
#pragma unroll 1
  for (int n = 0; n < steps; ++n) {
    const SdbGen gen(pkey, p_lo, p_hi, SDB_STEP(n));
    double czdW[M];
    {  // batch 0: dW (wiener.py:52)
      double dW[M];
      uint32_t w[M / 2][4];
      gen.template blocks<M / 2, true>(0u, w);
      sdb_normal_batch<M / 2, true>(w, tab, rh_g, dW);
#pragma unroll
      for (int j = 0; j < M; ++j) czdW[j] = cz * dW[j];
      sdb_mbar_wait(&empty[ring.slot], ring.phase ^ 1u);
      sdb_tm_fence_after();
      sdb_tm_st<M>(tm + Ws::C_RING + ring.slot * Ws::SLOT, dW);
      sdb_tm_wait_st();
      sdb_tm_fence_before();
      if (lane == 0) sdb_mbar_arrive(&full[ring.slot]);
      ring.advance<Ws::NSLOT>();
    }
#pragma unroll 1
    for (int k = 1; k <= n_terms; ++k) {  // batch k: X_k scaled by h/(2 pi k), Z_k = Y_k + sqrt(2/h) dW
      const uint32_t sx = (uint32_t)((M + 2 * (k - 1) * M) >> 1);
      const double hk = s_hk[k];
      double X[M], Z[M];
      {
        uint32_t w[M / 2][4];
        gen.template blocks<M / 2, false>(sx, w);
        sdb_normal_batch<M / 2, true>(w, tab, hk, X);
      }
      {
        uint32_t w[M / 2][4];
        gen.template blocks<M / 2, false>(sx + M / 2, w);
        sdb_normal_batch<M / 2, false>(w, tab, 1.0, Z);
      }
#pragma unroll
      for (int j = 0; j < M; ++j) Z[j] += czdW[j];
      sdb_mbar_wait(&empty[ring.slot], ring.phase ^ 1u);
      sdb_tm_fence_after();
      const uint32_t ta = tm + Ws::C_RING + ring.slot * Ws::SLOT;
      sdb_tm_st<M>(ta, X);
      sdb_tm_st<M>(ta + 2 * M, Z);
      sdb_tm_wait_st();
      sdb_tm_fence_before();
      if (lane == 0) sdb_mbar_arrive(&full[ring.slot]);
      ring.advance<Ws::NSLOT>();
    }
  }
}

// ---- one row block of the integration pipeline (main warp) -----------------------------------------
template <int D, int M, int B_>
struct SdbWsBlock {
  typedef SdbFusedShape<D, M> Sh;
  typedef SdbWsShape<D, M> Ws;
  static constexpr int RN = Sh::rn(B_), R0 = Sh::r0(B_), T1 = Sh::t1(B_), KT3 = Sh::kt3(B_);

  static SDB_DEV void run(const double* __restrict__ frag, const double* __restrict__ s_left,
                          double* gbuf, uint32_t tm, double h, double dg, int lane,
                          double (&w)[4][2], double (&wv)[Sh::IL > 0 ? Sh::IL : 1]) {
    constexpr int LT = Sh::LT;
    const int fr = lane >> 2, fc = lane & 3;
    const double* ybuf = gbuf + Ws::R_Y * 32;
    const double* abuf = gbuf + Ws::R_AT * 32 + lane;
    double* rbuf = gbuf + Ws::R_ROW * 32 + lane;
    // ---- P1: staging[n][p] = sum_l [B_j[R_b,:] ; A[R_b,:]][n][l] Y[l][p]  (DMMA) ------------------
    {
      double Yf[LT][4];
#pragma unroll
      for (int lt = 0; lt < LT; ++lt) {
        const int l = 4 * lt + fc;
        const bool ok = (4 * lt + 3 < D) || (l < D);
        const int lc = ok ? l : 0;
#pragma unroll
        for (int pt = 0; pt < 4; ++pt) {
          const double v = ybuf[lc * 32 + ((8 * pt + fr) ^ sdb_swz(lc))];
          Yf[lt][pt] = ok ? v : 0.0;
        }
      }
      const double* f1 = frag + (size_t)Sh::p1_off(B_) * 32 + lane;
#pragma unroll
      for (int t = 0; t < T1; ++t) {
        double af[LT];
#pragma unroll
        for (int lt = 0; lt < LT; ++lt) af[lt] = __ldg(f1 + (t * LT + lt) * 32);
        double c[4][2];
#pragma unroll
        for (int pt = 0; pt < 4; ++pt) {
          c[pt][0] = 0.0;
          c[pt][1] = 0.0;
        }
#pragma unroll
        for (int lt = 0; lt < LT; ++lt)
#pragma unroll
          for (int pt = 0; pt < 4; ++pt) sdb_dmma(c[pt][0], c[pt][1], af[lt], Yf[lt][pt]);
        double* row = gbuf + (8 * t + fr) * 32;
        const int sw = sdb_swz(fr);
#pragma unroll
        for (int pt = 0; pt < 4; ++pt) {
          row[(8 * pt + 2 * fc) ^ sw] = c[pt][0];
          row[(8 * pt + 2 * fc + 1) ^ sw] = c[pt][1];
        }
      }
    }
    __syncwarp();
    // ---- P2: s~[R_b,:] = G_n[R_b,:] I, one lane per path, Levy areas streamed from TMEM -----------
    {
      const volatile double* gv = gbuf;
      double dW[M];
      sdb_tm_ld<M>(tm + Ws::C_DW, dW);
      double g[RN][M], s[RN][M], GdW[RN];
#pragma unroll
      for (int r = 0; r < RN; ++r) GdW[r] = 0.0;
#pragma unroll
      for (int j = 0; j < M; ++j)
#pragma unroll
        for (int r = 0; r < RN; ++r) {
          const int n = j * RN + r;
          g[r][j] = gv[n * 32 + (lane ^ sdb_swz(n))];
          GdW[r] = fma(g[r][j], dW[j], GdW[r]);
        }
      // diagonal of I (Ito: -h/2, Stratonovich: 0) and the symmetric part dW dW^T / 2
#pragma unroll
      for (int k = 0; k < M; ++k) {
        const double hd = 0.5 * dW[k];
#pragma unroll
        for (int r = 0; r < RN; ++r) s[r][k] = fma(hd, GdW[r], dg * g[r][k]);
      }
      // antisymmetric part: pairs (j < k) in K_m row order, own-lane reads of the Levy areas
#pragma unroll
      for (int j = 0; j < M; ++j)
#pragma unroll
        for (int k = j + 1; k < M; ++k) {
          const double av = abuf[sdb_pair_index<M>(j, k) * 32];
#pragma unroll
          for (int r = 0; r < RN; ++r) {
            s[r][k] = fma(g[r][j], av, s[r][k]);
            s[r][j] = fma(-g[r][k], av, s[r][j]);
          }
        }
      // drift rows of this block: H20 = Y + f h (integrate.py:509); update without the f1 term
#pragma unroll
      for (int r = 0; r < RN; ++r) {
        const int n = M * RN + r;
        const double f0h = gv[n * 32 + (lane ^ sdb_swz(n))] * h;
        const double yr = ybuf[(R0 + r) * 32 + (lane ^ sdb_swz(R0 + r))];
        rbuf[(R0 + r) * 32] = yr + f0h;                          // H20
        rbuf[(D + R0 + r) * 32] = fma(-0.5, f0h, GdW[r]);        // Yacc
      }
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
          gbuf[K * 32 + (lane ^ sdb_swz(K))] = s[r][k];
        }
    }
    __syncwarp();
    // ---- P3: w[i][p] += sum_K Bst[i][K] s~[K][p], rows i < 8 IT (DMMA) --------------------------------
    if (Sh::IT > 0) {
      const double* f3 = frag + (size_t)Sh::p3_off(B_) * 32 + lane;
#pragma unroll
      for (int kt = 0; kt < KT3; ++kt) {
        const double af = __ldg(f3 + kt * Sh::IT * 32);
        const int K = 4 * kt + fc;
        const double* row = gbuf + K * 32;
        const int sw = sdb_swz(K);
#pragma unroll
        for (int pt = 0; pt < 4; ++pt) sdb_dmma(w[pt][0], w[pt][1], af, row[(8 * pt + fr) ^ sw]);
      }
    }
    __syncwarp();
  }
};

template <int D, int M, int B_>
struct SdbWsBlocks {
  template <class... Args>
  static SDB_DEV void run(Args&... args) {
    SdbWsBlocks<D, M, B_ - 1>::run(args...);
    SdbWsBlock<D, M, B_>::run(args...);
  }
};
template <int D, int M>
struct SdbWsBlocks<D, M, -1> {
  template <class... Args>
  static SDB_DEV void run(Args&...) {}
};

// ---- main warp ---------------------------------------------------------------------------------------
template <int D, int M>
SDB_DEV void sdb_ws_main(const SdbLoopArgs& a, const PInline<D * D>& fp,
                         const double* __restrict__ frag, double* gbuf,
                         const double2* __restrict__ tab, const double* __restrict__ s_hk,
                         const double* __restrict__ s_left, uint32_t tm, uint64_t* full,
                         uint64_t* empty, int64_t p, int lane) {
  typedef SdbFusedShape<D, M> Sh;
  typedef SdbWsShape<D, M> Ws;
  constexpr int MM = Sh::MM;
  const bool active = p < a.P;  // idle lanes of the last warp still take part in the DMMAs
  const int fr = lane >> 2, fc = lane & 3;
  double* ybuf = gbuf + Ws::R_Y * 32;
  double* abuf = gbuf + Ws::R_AT * 32 + lane;
  const double* rbuf = gbuf + Ws::R_ROW * 32 + lane;
  {
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
#pragma unroll
    for (int l = 0; l < D; ++l) ybuf[l * 32 + (lane ^ sdb_swz(l))] = Y[l];
    __syncwarp();
  }
  const double dg = (a.flags & SDB_FLAG_STRATONOVICH) ? 0.0 : -0.5 * a.h_global;
  const int steps = a.N - 1;
  const int n_terms = a.n_terms;
  const double* __restrict__ hs = a.tspan + a.N;  // per-step h (host-built)
  int next_save = a.save_every;
  int64_t slot = 1;
#if SDB_WS_HELPERS
  SdbRing ring;
#else
  const double rh_g = sqrt(a.h_global);
  const double cz = sqrt(2.0 / a.h_global);
  const uint64_t path = (uint64_t)(a.path_id0 + p);
  const uint32_t p_lo = (uint32_t)path, p_hi = (uint32_t)(path >> 32);
  const SdbPhiloxKey pkey(a.seed);
#endif

#pragma unroll 1
  for (int n = 0; n < steps; ++n) {
    const double h = hs[n];
    // ---- Levy areas (wiener.py:67-74, :102-105) from the normals of this step -----------------------
    {
      double At[MM];
#pragma unroll
      for (int r = 0; r < MM; ++r) At[r] = 0.0;
#if SDB_WS_HELPERS
      {  // batch 0: dW -> private columns (the ring slot is recycled)
        sdb_mbar_wait(&full[ring.slot], ring.phase);
        sdb_tm_fence_after();
        double dW[M];
        sdb_tm_ld<M>(tm + Ws::C_RING + ring.slot * Ws::SLOT, dW);
        sdb_tm_fence_before();
        if (lane == 0) sdb_mbar_arrive(&empty[ring.slot]);
        ring.advance<Ws::NSLOT>();
        sdb_tm_st<M>(tm + Ws::C_DW, dW);
      }
#pragma unroll 1
      for (int k = 1; k <= n_terms; ++k) {
        sdb_mbar_wait(&full[ring.slot], ring.phase);
        sdb_tm_fence_after();
        double XZ[2 * M];
        sdb_tm_ld<2 * M>(tm + Ws::C_RING + ring.slot * Ws::SLOT, XZ);
        sdb_tm_fence_before();
        if (lane == 0) sdb_mbar_arrive(&empty[ring.slot]);
        ring.advance<Ws::NSLOT>();
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
          for (int j = i + 1; j < M; ++j) {
            const int r = sdb_pair_index<M>(i, j);
            At[r] = fma(XZ[i], XZ[M + j], At[r]);
            At[r] = fma(-XZ[M + i], XZ[j], At[r]);
          }
      }
#else
      {
        const SdbGen gen(pkey, p_lo, p_hi, SDB_STEP(n));
        double dW[M];
        {
          uint32_t w[M / 2][4];
          gen.template blocks<M / 2, true>(0u, w);
          sdb_normal_batch<M / 2, true>(w, tab, rh_g, dW);
        }
        sdb_tm_st<M>(tm + Ws::C_DW, dW);
#pragma unroll 1
        for (int k = 1; k <= n_terms; ++k) {
          double X[M], Z[M];
          const uint32_t sx = (uint32_t)((M + 2 * (k - 1) * M) >> 1);
          const double hk = s_hk[k];
          {
            uint32_t w[M / 2][4];
            gen.template blocks<M / 2, false>(sx, w);
            sdb_normal_batch<M / 2, true>(w, tab, hk, X);
            gen.template blocks<M / 2, false>(sx + M / 2, w);
            sdb_normal_batch<M / 2, false>(w, tab, 1.0, Z);
          }
#pragma unroll
          for (int j = 0; j < M; ++j) Z[j] = fma(cz, dW[j], Z[j]);
#pragma unroll
          for (int i = 0; i < M; ++i)
#pragma unroll
            for (int j = i + 1; j < M; ++j) {
              const int r = sdb_pair_index<M>(i, j);
              At[r] = fma(X[i], Z[j], At[r]);
              At[r] = fma(-Z[i], X[j], At[r]);
            }
        }
      }
#endif
#pragma unroll
      for (int r = 0; r < MM; ++r) abuf[r * 32] = At[r];
      sdb_tm_wait_st();  // dW: own store complete before the loads of P2
    }

    // ---- row blocks ----------------------------------------------------------------------------------
    double w[4][2], wv[Sh::IL > 0 ? Sh::IL : 1];
#pragma unroll
    for (int pt = 0; pt < 4; ++pt) {
      w[pt][0] = 0.0;
      w[pt][1] = 0.0;
    }
#pragma unroll
    for (int q = 0; q < (Sh::IL > 0 ? Sh::IL : 1); ++q) wv[q] = 0.0;
    SdbWsBlocks<D, M, Sh::NB - 1>::run(frag, s_left, gbuf, tm, h, dg, lane, w, wv);

    // ---- assemble Y_{n+1} = H20 - f0h/2 + f1h/2 + G dW + stage-2 sum ---------------------------------
    double Y[D];
    {
      double H20[D], Yacc[D];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        H20[i] = rbuf[i * 32];
        Yacc[i] = rbuf[(D + i) * 32];
      }
      // second drift evaluation f(H20, t_{n+1}) h (integrate.py:514)
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double acc = fp.v[i * D] * H20[0];
#pragma unroll
        for (int l = 1; l < D; ++l) acc = fma(fp.v[i * D + l], H20[l], acc);
        Yacc[i] = fma(0.5 * h, acc, Yacc[i]);
      }
      if (Sh::IT > 0) {  // stage-2 sum back to one lane per path
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
      for (int i = 0; i < D; ++i) Y[i] = H20[i] + Yacc[i];
    }
#pragma unroll
    for (int l = 0; l < D; ++l) ybuf[l * 32 + (lane ^ sdb_swz(l))] = Y[l];
    __syncwarp();

    if (n + 1 == next_save) {
      if (active) sdb_store<D>(a, p, slot, Y);
      ++slot;
      next_save += a.save_every;
    }
    if (n + 1 == steps && active) sdb_finish<D>(a, p, Y);
  }
}

template <int D, int M>
SDB_DEV void sdb_srk2_fusedws_body(const SdbLoopArgs& a, const PInline<D * D>& fp,
                                   const double* __restrict__ frag, double* smem) {
  typedef SdbFusedShape<D, M> Sh;
  typedef SdbWsShape<D, M> Ws;
  static_assert(M % 2 == 0, "fused kernel needs an even number of Wiener processes");
  static_assert(Sh::IT <= 1, "stage-2 DMMA tiles: d < 16 supported");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  double2* tab = reinterpret_cast<double2*>(smem + (size_t)Ws::WROWS * 32 * SDB_WS_MAIN);
  double* s_hk = reinterpret_cast<double*>(tab + 128);
  double* s_left = s_hk + SDB_FUSED_MAX_TERMS + 1;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_left + Sh::n_left() + 1);  // [main][full 8 | empty 8]
  uint32_t* tm_slot = reinterpret_cast<uint32_t*>(bars + SDB_WS_MAIN * 16);
  if (tid < 128) tab[tid] = reinterpret_cast<const double2*>(sdb_logtab)[tid];
  for (int i = tid; i <= SDB_FUSED_MAX_TERMS; i += (int)blockDim.x)
    s_hk[i] = i == 0 ? 0.0 : a.h_global * SDB_INV_2PI / (double)i;
  for (int i = tid; i < Sh::n_left(); i += SDB_WS_THREADS) s_left[i] = frag[(size_t)Sh::n_frags() * 32 + i];
  if (tid < SDB_WS_MAIN * 16) sdb_mbar_init(&bars[tid], 1u);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  const uint32_t tm_base = sdb_tm_alloc512(tm_slot);  // contains a __syncthreads()

  const int mw = warp & (SDB_WS_MAIN - 1);  // main warp index this warp is (or serves)
  const int64_t p = (int64_t)blockIdx.x * SDB_WS_PATHS + mw * 32 + lane;
  // TMEM address: lane quarter (warp % 4) in bits 31:16, column in bits 15:0
  const uint32_t tm = tm_base + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)((mw >> 2) * 256);
  uint64_t* full = bars + mw * 16;
  uint64_t* empty = full + 8;
  if (warp < SDB_WS_MAIN) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(SDB_WS_MAIN_REGS));
    double* gbuf = smem + (size_t)warp * Ws::WROWS * 32;
    sdb_ws_main<D, M>(a, fp, frag, gbuf, tab, s_hk, s_left, tm, full, empty, p, lane);
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(SDB_WS_HELP_REGS));
#if SDB_WS_HELPERS
    sdb_ws_helper<D, M>(a, tab, s_hk, tm, full, empty, p, lane);
#endif
  }
  sdb_tm_free512(tm_base);
}
#endif  // __CUDACC__

}  // namespace SDB_FUSED_NS
using namespace SDB_FUSED_NS;

#define SDB_FUSEDWS_KERNEL(NAME, D, M)                                                        \
  extern "C" __global__ void __launch_bounds__(SDB_WS_THREADS, 1)                             \
      NAME(const __grid_constant__ SdbLoopArgs a, const __grid_constant__ PInline<D * D> fp,  \
           const double* __restrict__ frag) {                                                 \
    extern __shared__ __align__(16) double sdb_fused_smem[];                                  \
    sdb_srk2_fusedws_body<D, M>(a, fp, frag, sdb_fused_smem);                                 \
  }
