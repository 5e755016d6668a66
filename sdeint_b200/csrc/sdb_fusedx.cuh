// sdb_fusedx.cuh -- fused SRI2/SRS2 kernel for LARGE linear systems (d = m = 20, BASELINE.json
// configs[3]) and for the Wiktorsson repeated integrals (Iwik/Jwik, sdeint/wiener.py:234-305): the
// whole time loop of sdeint/integrate.py:494-522 with on-device Philox noise in one launch.
//
// At m = 20 a path owns M = 190 Levy areas -- 380 registers, more than a thread has -- and the
// shared memory of an SM holds them for four warps only.  They live in TENSOR MEMORY instead, used as
// a per-lane scratchpad (sdb_tmem.cuh: tcgen05.ld/st, 32x32b shape, ~5x the bandwidth of shared
// memory and off the LSU): one warp per SM sub-partition, 32 paths per warp, each lane owning up to
// 512 columns = 256 doubles of its TMEM lane:
//     [0, 2 MMP)  Levy areas, pair-indexed (K_m row order), in chunks of 16
//     then        dW (m doubles), then the per-row pieces H20[d], Yacc[d] of the update.
// Per step:
//   noise   dW; for k = 1..n: X_k, Z_k = Y_k + sqrt(2/h) dW (lock-step Philox + Box-Muller batches),
//           then the areas are read-modify-written chunk by chunk: A_ij += X_i Z_j - Z_i X_j
//           (X pre-scaled by h / 2 pi k).  TMEM traffic ~15 KB per lane-step: ~4 % of the step.
//   tail    (Wiktorsson) two passes, no storage for the M tail normals g:
//             1. A_ij += c1 g_ij,  u_i += g_ij dW_j,  u_j -= g_ij dW_i      (u = G^ dW)
//             2. A_ij += c2 (dW_j u_i - dW_i u_j)
//           with c1 = (h/2pi) sqrt(a_n) (2 + 2r) / (sqrt2 (1 + r)),  c2 = (h/2pi) sqrt(a_n) (2/h) /
//           (sqrt2 (1 + r)),  r = sqrt(1 + |dW|^2 / h): the O(m^2) form of wiener.py:217-226, :269-277
//           (same arithmetic as sdb_wik_tail in sdb_kernels.cuh).
//   blocks  of 2 state rows (runtime loop, all blocks alike): P1 DMMA [B_j[R,:]; A[R,:]] Y -> staging;
//           P2 DFMA s~ = G_n[R,:] I with the areas streamed from TMEM (for j<k: s[:,k] += g[:,j] a,
//           s[:,j] -= g[:,k] a; sign flipped for Wiktorsson's transposed placement, wiener.py:278-280);
//           P3 DMMA w += [B_k[:, R]] s~ (rows >= 8 floor(d/8) on DFMA with broadcast operands).
//   final   second drift evaluation, Y_{n+1} = H20 - f0h/2 + f1h/2 + G dW + w.
#pragma once
#ifndef SDB_FUSED_RB
#define SDB_FUSED_RB 2
#endif
#include "sdb_fusedws.cuh"  // sdb_fused.cuh, sdb_tmem.cuh, sdb_static_for, sdb_pair_j/k


namespace SDB_FUSED_NS {

template <int D, int M>
struct SdbXShape {
  typedef SdbFusedShape<D, M, SDB_FUSED_A2> Sh;  // rows of A^2 ride in the P1 product (sdb_fused.cuh)
  static constexpr int MM = Sh::MM;
  static constexpr int CH = 16;                              // areas per TMEM chunk (32 columns)
  static constexpr int NCH = (MM + CH - 1) / CH;
  static constexpr int MMP = NCH * CH;
  static constexpr int C_AT = 0;
  static constexpr int C_DW = 2 * MMP;
  static constexpr int C_ROW = C_DW + (2 * M + 7) / 8 * 8;
  static constexpr int C_END = C_ROW + 4 * D;
  static_assert(C_END <= 512, "tensor-memory columns of a lane exceeded");
  static_assert(D % Sh::RB == 0, "uniform row blocks needed (runtime block loop)");
  static constexpr int RN = Sh::RB, NB = D / Sh::RB;
  static constexpr int T1 = Sh::t1(0), KT3 = Sh::kt3(0), LT = Sh::LT, IT = Sh::IT, IL = Sh::IL;
  static constexpr int GR = Sh::grows();
  static constexpr int R_Y = GR;                             // shared-memory rows of one warp
  static constexpr int R_F = R_Y + D;                        // SDB_FUSED_A2 -1: f(Y_n) h of the user's drift
  static constexpr int WROWS = R_Y + D + (SDB_FUSED_A2 < 0 ? D : 0);
  // warps per SM sub-partition: two when two warps' columns fit the 512 of a TMEM lane (and the
  // register file: 2 x 255 registers per lane slot)
#ifdef SDB_X_WPQ
  static constexpr int WPQ = SDB_X_WPQ;                      // tuning override (3: 168 registers/thread)
#else
  static constexpr int WPQ = 2 * C_END <= 512 ? 2 : 1;
#endif
  static constexpr int WARPS = 4 * WPQ, THREADS = 32 * WARPS, PATHS = THREADS;
  static constexpr int C_STRIDE = (512 / WPQ) / 8 * 8;       // TMEM columns per warp
  static_assert(C_STRIDE >= C_END, "tensor-memory columns of a lane exceeded");
  static SDB_CX size_t smem_bytes() {
    size_t b = (size_t)WROWS * 32 * 8 * WARPS + 128 * 16 +
               (SDB_FUSED_MAX_TERMS + 1 + Sh::n_left() + 1) * sizeof(double) + 16;
    return b < 120 * 1024 ? 120 * 1024 : b;  // > half an SM: one block (one TMEM allocation) per SM
  }
};

#ifdef __CUDACC__
// NP Box-Muller pairs (slots slot0 ..) in lock-step batches of at most 5: out[0 .. 2 NP)
// FIRST: slot0 == 0 (slot 0 is the step's plain Philox block, the others its keyed expansion)
template <int NP, int OFF, int TOT, bool SCALED, bool FIRST = false>
SDB_DEV void sdb_gen_pairs(const SdbGen& gen, uint32_t slot0, const double2* __restrict__ tab, double scale,
                           double (&out)[TOT]) {
  if constexpr (NP > 0) {
    constexpr int B = NP >= 5 ? 5 : NP;
    uint32_t w[B][4];
    double z[2 * B];
    gen.template blocks<B, FIRST && OFF == 0>(slot0 + (uint32_t)OFF, w);
    sdb_normal_batch<B, SCALED>(w, tab, scale, z);
#pragma unroll
    for (int i = 0; i < 2 * B; ++i) out[2 * OFF + i] = z[i];
    sdb_gen_pairs<NP - B, OFF + B, TOT, SCALED, FIRST>(gen, slot0, tab, scale, out);
  }
}

// Levy areas of one (path, step) in TENSOR MEMORY: the Kloeden-Platen-Wright series (wiener.py:67-74,
// :102-105 / :264-267) read-modify-written chunk by chunk at columns tm_at (which must hold zeros on
// entry), plus Wiktorsson's tail when WIK.  dW is read from columns tm_dw.  X_k is pre-scaled by
// s_hk[k] = h / (2 pi k).  Used by the fused step kernel below and by the standalone
// Ikpw/Jkpw/Iwik/Jwik kernel (sdb_repintx.cuh).
template <int M, int WIK>
SDB_DEV void sdb_x_levy_areas(const SdbGen& gen,
                              const double2* __restrict__ tab, const double* __restrict__ s_hk,
                              double hg, double cz, double tail_scale, int n_terms, uint32_t tm_at,
                              uint32_t tm_dw) {
  constexpr int MM = M * (M - 1) / 2, CH = 16, NCH = (MM + CH - 1) / CH;
  // ---- Levy-area series (wiener.py:67-74, :102-105 / :264-267) ------------------------------------
#pragma unroll 1
  for (int k = 1; k <= n_terms; ++k) {
    double X[M], Z[M];
    const uint32_t sx = (uint32_t)((M + 2 * (k - 1) * M) >> 1);
    sdb_gen_pairs<M / 2, 0, M, true>(gen, sx, tab, s_hk[k], X);
    sdb_gen_pairs<M / 2, 0, M, false>(gen, sx + M / 2, tab, 1.0, Z);
    {
      double dW[M];
      sdb_tm_ld<M>(tm_dw, dW);
#pragma unroll
      for (int j = 0; j < M; ++j) Z[j] = fma(cz, dW[j], Z[j]);
    }
    SdbTmLoad<CH> ca, cb;
    ca.issue(tm_at);
    ca.wait();
    sdb_static_for<0, NCH>([&](auto cc) {
      constexpr int c = decltype(cc)::value;
      if constexpr (c + 1 < NCH) {
        if constexpr (c & 1) ca.issue(tm_at + 2 * CH * (c + 1));
        else cb.issue(tm_at + 2 * CH * (c + 1));
      }
      double At[CH];
#pragma unroll
      for (int e = 0; e < CH; ++e) At[e] = (c & 1) ? cb.get(e) : ca.get(e);
      sdb_static_for<0, CH>([&](auto ee) {
        constexpr int e = decltype(ee)::value;
        constexpr int idx = CH * c + e;
        if constexpr (idx < MM) {
          constexpr int i = sdb_pair_j<M>(idx), j = sdb_pair_k<M>(idx);
          At[e] = fma(X[i], Z[j], At[e]);
          At[e] = fma(-Z[i], X[j], At[e]);
        }
      });
      sdb_tm_st<CH>(tm_at + 2 * CH * c, At);
      if constexpr (c + 1 < NCH) {
        if constexpr (c & 1) ca.wait();
        else cb.wait();
      }
    });
    sdb_tm_wait_st();
  }
  // ---- Wiktorsson's tail (wiener.py:269-277), O(m^2), no storage for the tail normals ----------------
  if (WIK) {
    double dW[M], u[M];
    sdb_tm_ld<M>(tm_dw, dW);
    double nrm = 0.0;
#pragma unroll
    for (int j = 0; j < M; ++j) {
      nrm = fma(dW[j], dW[j], nrm);
      u[j] = 0.0;
    }
    const double rad = sqrt(1.0 + nrm / hg);
    const double den = tail_scale / (1.4142135623730951 * (1.0 + rad));
    const double c1 = den * (2.0 + 2.0 * rad);
    const double c2 = den * (2.0 / hg);
    const uint32_t st0 = (uint32_t)((M + 2 * n_terms * M) >> 1);
    // pass 1: the generator code exists once (runtime loop); only the index-dependent FMAs of a
    // chunk are selected by a warp-uniform jump (keeps the kernel inside the instruction cache)
#pragma unroll 1
    for (int c = 0; c < NCH; ++c) {
      double g[CH], At[CH];
      sdb_gen_pairs<CH / 2, 0, CH, false>(gen, st0 + (uint32_t)(c * CH / 2),
                                          tab, 1.0, g);
      sdb_tm_ld<CH>(tm_at + 2 * CH * c, At);
      sdb_static_for<0, NCH>([&](auto cc) {
        constexpr int c_ = decltype(cc)::value;
        if (c == c_) {
          sdb_static_for<0, CH>([&](auto ee) {
            constexpr int e = decltype(ee)::value;
            constexpr int idx = CH * c_ + e;
            if constexpr (idx < MM) {
              constexpr int i = sdb_pair_j<M>(idx), j = sdb_pair_k<M>(idx);
              At[e] = fma(c1, g[e], At[e]);
              u[i] = fma(g[e], dW[j], u[i]);
              u[j] = fma(-g[e], dW[i], u[j]);
            }
          });
        }
      });
      sdb_tm_st<CH>(tm_at + 2 * CH * c, At);
    }
    sdb_tm_wait_st();
    sdb_static_for<0, NCH>([&](auto cc) {
      constexpr int c = decltype(cc)::value;
      double At[CH];
      sdb_tm_ld<CH>(tm_at + 2 * CH * c, At);
      sdb_static_for<0, CH>([&](auto ee) {
        constexpr int e = decltype(ee)::value;
        constexpr int idx = CH * c + e;
        if constexpr (idx < MM) {
          constexpr int i = sdb_pair_j<M>(idx), j = sdb_pair_k<M>(idx);
          At[e] = fma(c2, fma(dW[j], u[i], -dW[i] * u[j]), At[e]);
        }
      });
      sdb_tm_st<CH>(tm_at + 2 * CH * c, At);
    });
    sdb_tm_wait_st();
  }
}

template <int D, int M, int WIK, class FP = PInline<D * D>>
SDB_DEV void sdb_srk2_fusedx_body(const SdbLoopArgs& a, const FP& fp,
                                  const double* __restrict__ frag, double* smem) {
  typedef SdbFusedShape<D, M, SDB_FUSED_A2> Sh;  // rows of A^2 ride in the P1 product (sdb_fused.cuh)
  typedef SdbXShape<D, M> Xs;
  constexpr int MM = Xs::MM, CH = Xs::CH, NCH = Xs::NCH, RN = Xs::RN, NB = Xs::NB, LT = Xs::LT;
  constexpr int T1 = Xs::T1, KT3 = Xs::KT3, IT = Xs::IT, IL = Xs::IL;
  static_assert(M % 2 == 0, "fused kernel needs an even number of Wiener processes");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int fr = lane >> 2, fc = lane & 3;

  double2* tab = reinterpret_cast<double2*>(smem + (size_t)Xs::WROWS * 32 * Xs::WARPS);
  double* s_hk = reinterpret_cast<double*>(tab + 128);
  double* s_left = s_hk + SDB_FUSED_MAX_TERMS + 1;
  uint32_t* tm_slot = reinterpret_cast<uint32_t*>(s_left + Sh::n_left() + 1);
  if (tid < 128) tab[tid] = reinterpret_cast<const double2*>(sdb_logtab)[tid];
  for (int i = tid; i <= SDB_FUSED_MAX_TERMS; i += (int)blockDim.x)
    s_hk[i] = i == 0 ? 0.0 : a.h_global * SDB_INV_2PI / (double)i;
  for (int i = tid; i < Sh::n_left(); i += Xs::THREADS) s_left[i] = frag[(size_t)Sh::n_frags() * 32 + i];
  const uint32_t tm_base = sdb_tm_alloc512(tm_slot);  // contains a __syncthreads()
  const uint32_t tm = tm_base + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)((warp >> 2) * Xs::C_STRIDE);

  double* gbuf = smem + (size_t)warp * Xs::WROWS * 32;
  double* ybuf = gbuf + Xs::R_Y * 32;
  const int64_t p = (int64_t)blockIdx.x * Xs::PATHS + tid;
  const bool active = p < a.P;  // idle lanes of the last warp still take part in the DMMAs
#if SDB_OBSERVE
  SdbObs<D> obs;
#endif
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
#if SDB_OBSERVE
    obs.init(a, Y);
#endif
#pragma unroll
    for (int l = 0; l < D; ++l) ybuf[l * 32 + (lane ^ sdb_swz3(l))] = Y[l];
    __syncwarp();
  }

  const double hg = a.h_global;
  const double rh_g = sqrt(hg);
  const double cz = sqrt(2.0 / hg);
  const double dg = (a.flags & SDB_FLAG_STRATONOVICH) ? 0.0 : -0.5 * hg;
  const uint64_t path = (uint64_t)(a.path_id0 + p);
  const uint32_t p_lo = (uint32_t)path, p_hi = (uint32_t)(path >> 32);
  const SdbPhiloxKey pkey(a.seed);
  const int steps = a.N - 1;
  const int n_terms = a.n_terms;
  const double tail_scale = WIK ? hg * SDB_INV_2PI * sqrt(sdb_a_tail(n_terms)) : 0.0;
  const double* __restrict__ hs = a.tspan + a.N;  // per-step h (host-built)
  int next_save = a.save_every;
  int64_t slot = 1;

#pragma unroll 1
  for (int n = 0; n < steps; ++n) {
    const double h = hs[n];
    const SdbGen gen(pkey, p_lo, p_hi, SDB_STEP(n));
    // ---- dW (wiener.py:52) ---------------------------------------------------------------------------
    {
      double dW[M];
      sdb_gen_pairs<M / 2, 0, M, true, true>(gen, 0u, tab, rh_g, dW);
      sdb_tm_st<M>(tm + Xs::C_DW, dW);
      double zero[CH];
#pragma unroll
      for (int e = 0; e < CH; ++e) zero[e] = 0.0;
#pragma unroll
      for (int c = 0; c < NCH; ++c) sdb_tm_st<CH>(tm + Xs::C_AT + 2 * CH * c, zero);
      sdb_tm_wait_st();
    }
    // ---- Levy areas (series + Wiktorsson tail) into tensor memory ------------------------------------
    sdb_x_levy_areas<M, WIK>(gen, tab, s_hk, hg, cz, tail_scale, n_terms,
                             tm + Xs::C_AT, tm + Xs::C_DW);

#if SDB_FUSED_A2 < 0
    {  // any drift, evaluated one lane per path (sdb_fused.cuh): f(Y_n, t_n) h waits in shared memory
      double Yv[D], f0[D];
#pragma unroll
      for (int l = 0; l < D; ++l) Yv[l] = ybuf[l * 32 + (lane ^ sdb_swz3(l))];
      SDB_FUSED_DRIFT_T::eval(fp, Yv, a.tspan[n], f0);
#pragma unroll
      for (int l = 0; l < D; ++l) gbuf[(Xs::R_F + l) * 32 + lane] = f0[l] * h;
      __syncwarp();
    }
#endif
    // ---- row blocks (runtime loop: every block has RN rows) ------------------------------------------
    double w[IT > 0 ? IT : 1][4][2], wv[IL > 0 ? IL : 1];
#pragma unroll
    for (int it = 0; it < (IT > 0 ? IT : 1); ++it)
#pragma unroll
      for (int pt = 0; pt < 4; ++pt) {
        w[it][pt][0] = 0.0;
        w[it][pt][1] = 0.0;
      }
#pragma unroll
    for (int q = 0; q < (IL > 0 ? IL : 1); ++q) wv[q] = 0.0;

#pragma unroll 1
    for (int b = 0; b < NB; ++b) {
      const int R0 = RN * b;
      // P1: staging[n][p] = sum_l [B_j[R_b,:] ; A[R_b,:]][n][l] Y[l][p]  (DMMA)
      {
        double Yf[LT][4];
#pragma unroll
        for (int lt = 0; lt < LT; ++lt) {
          const int l = 4 * lt + fc;
          const bool ok = (4 * lt + 3 < D) || (l < D);
          const int lc = ok ? l : 0;
#pragma unroll
          for (int pt = 0; pt < 4; ++pt) {
            const double v = ybuf[lc * 32 + ((8 * pt + fr) ^ sdb_swz3(lc))];
            Yf[lt][pt] = ok ? v : 0.0;
          }
        }
        const double* f1 = frag + (size_t)(b * T1 * LT) * 32 + lane;
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
      // P2: s~[R_b,:] = G_n[R_b,:] I, one lane per path, Levy areas streamed from TMEM
      {
        const volatile double* gv = gbuf;
        SdbTmLoad<CH> ca, cb;
        ca.issue(tm + Xs::C_AT);
        double g[RN][M], s[RN][M], GdW[RN];
        {
          double dW[M];
          sdb_tm_ld<M>(tm + Xs::C_DW, dW);  // its wait covers the first chunk as well
          ca.depend();
#pragma unroll
          for (int r = 0; r < RN; ++r) GdW[r] = 0.0;
#pragma unroll
          for (int j = 0; j < M; ++j)
#pragma unroll
            for (int r = 0; r < RN; ++r) {
              const int nn = j * RN + r;
              g[r][j] = gv[nn * 32 + (lane ^ sdb_swz(nn))];
              GdW[r] = fma(g[r][j], dW[j], GdW[r]);
            }
          // diagonal of I (Ito: -h/2, Stratonovich: 0) and the symmetric part dW dW^T / 2
#pragma unroll
          for (int k = 0; k < M; ++k) {
            const double hd = 0.5 * dW[k];
#pragma unroll
            for (int r = 0; r < RN; ++r) s[r][k] = fma(hd, GdW[r], dg * g[r][k]);
          }
        }
        double sj[2][RN];
#pragma unroll
        for (int r = 0; r < RN; ++r) {
          sj[0][r] = 0.0;
          sj[1][r] = 0.0;
        }
        sdb_static_for<0, NCH>([&](auto cc) {
          constexpr int c = decltype(cc)::value;
          if constexpr (c + 1 < NCH) {
            if constexpr (c & 1) ca.issue(tm + Xs::C_AT + 2 * CH * (c + 1));
            else cb.issue(tm + Xs::C_AT + 2 * CH * (c + 1));
          }
          sdb_static_for<0, CH>([&](auto ee) {
            constexpr int e = decltype(ee)::value;
            constexpr int idx = CH * c + e;
            if constexpr (idx < MM) {
              constexpr int j = sdb_pair_j<M>(idx), k = sdb_pair_k<M>(idx);
              double av;
              if constexpr (c & 1) av = cb.get(e);
              else av = ca.get(e);
              if (WIK) av = -av;  // column-major unvec: +Atilde goes BELOW the diagonal (wiener.py:278-280)
              // the M-1-j updates of column j form one dependent chain per row: two partial sums
              // (k - j odd / even) halve it -- with one warp per sub-partition latency is exposed
              constexpr int par = (k - j) & 1;
#pragma unroll
              for (int r = 0; r < RN; ++r) {
                s[r][k] = fma(g[r][j], av, s[r][k]);
                sj[par][r] = fma(-g[r][k], av, sj[par][r]);
              }
              if constexpr (k == M - 1) {
#pragma unroll
                for (int r = 0; r < RN; ++r) {
                  s[r][j] += sj[0][r] + sj[1][r];
                  sj[0][r] = 0.0;
                  sj[1][r] = 0.0;
                }
              }
            }
          });
          if constexpr (c + 1 < NCH) {
            if constexpr (c & 1) ca.wait();
            else cb.wait();
          }
        });
        // drift rows of this block: H20 = Y + f h (integrate.py:509); update without the f1 term
        {
          double piece[2 * RN];
#pragma unroll
          for (int r = 0; r < RN; ++r) {
            const int yr = R0 + r;
#if SDB_FUSED_A2 < 0
            const double f0h = gv[(Xs::R_F + yr) * 32 + lane];
#else
            const int nn = M * RN + r;
            const double f0h = gv[nn * 32 + (lane ^ sdb_swz(nn))] * h;
#endif
            const double yv = ybuf[yr * 32 + (lane ^ ((yr & 3) << 2))];
            piece[r] = yv + f0h;                      // H20
#if SDB_FUSED_A2 > 0
            const int n2 = (M + 1) * RN + r;          // f(H20) h = f0 h + h^2 A^2 Y: no second drift loop
            piece[RN + r] = fma(0.5 * h * h, gv[n2 * 32 + (lane ^ sdb_swz(n2))], GdW[r]);
#else
            piece[RN + r] = fma(-0.5, f0h, GdW[r]);   // Yacc
#endif
          }
          sdb_tm_st<2 * RN>(tm + Xs::C_ROW + 4 * R0, piece);
        }
        // rows >= 8 IT of the stage-2 sum on the FMA pipe (uniform operands from shared memory)
        if (IL > 0) {
#pragma unroll
          for (int q = 0; q < IL; ++q) {
            const double* bl = s_left + (b * IL + q) * RN * M;
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
      // P3: w[i][p] += sum_K Bst[i][K] s~[K][p], rows i < 8 IT (DMMA)
      if (IT > 0) {
        const double* f3 = frag + (size_t)(NB * T1 * LT + b * KT3 * IT) * 32 + lane;
#pragma unroll
        for (int kt = 0; kt < KT3; ++kt) {
          const int K = 4 * kt + fc;
          const double* row = gbuf + K * 32;
          const int sw = sdb_swz3(K);
          double bf[4];
#pragma unroll
          for (int pt = 0; pt < 4; ++pt) bf[pt] = row[(8 * pt + fr) ^ sw];
#pragma unroll
          for (int it = 0; it < IT; ++it) {
            const double af = __ldg(f3 + (kt * IT + it) * 32);
#pragma unroll
            for (int pt = 0; pt < 4; ++pt) sdb_dmma(w[it][pt][0], w[it][pt][1], af, bf[pt]);
          }
        }
      }
      __syncwarp();
    }
    sdb_tm_wait_st();  // the per-row pieces

    // ---- assemble Y_{n+1} = H20 - f0h/2 + f1h/2 + G dW + stage-2 sum -----------------------------------
    double Y[D];
    {
      double H20[D], Yacc[D];
      {
        double rows[2 * D];
        sdb_tm_ld<2 * D>(tm + Xs::C_ROW, rows);
#pragma unroll
        for (int i = 0; i < D; ++i) {  // block i / RN stored [H20 rows | Yacc rows]
          H20[i] = rows[2 * RN * (i / RN) + i % RN];
          Yacc[i] = rows[2 * RN * (i / RN) + RN + i % RN];
        }
      }
      // second drift evaluation f(H20, t_{n+1}) h (integrate.py:514): in Yacc already with SDB_FUSED_A2
#if SDB_FUSED_A2 == 0
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double acc = fp.v[i * D] * H20[0];
#pragma unroll
        for (int l = 1; l < D; ++l) acc = fma(fp.v[i * D + l], H20[l], acc);
        Yacc[i] = fma(0.5 * h, acc, Yacc[i]);
      }
#elif SDB_FUSED_A2 < 0
      {
        double f1[D];
        SDB_FUSED_DRIFT_T::eval(fp, H20, a.tspan[n + 1], f1);
#pragma unroll
        for (int i = 0; i < D; ++i) Yacc[i] = fma(0.5 * h, f1[i], Yacc[i]);
      }
#endif
#pragma unroll
      for (int it = 0; it < IT; ++it) {  // stage-2 sum back to one lane per path
        double* row = gbuf + fr * 32;
        const int sw = sdb_swz(fr);
#pragma unroll
        for (int pt = 0; pt < 4; ++pt) {
          row[(8 * pt + 2 * fc) ^ sw] = w[it][pt][0];
          row[(8 * pt + 2 * fc + 1) ^ sw] = w[it][pt][1];
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) Yacc[8 * it + i] += gbuf[i * 32 + (lane ^ sdb_swz(i))];
        __syncwarp();
      }
#pragma unroll
      for (int q = 0; q < IL; ++q) Yacc[8 * IT + q] += wv[q];
#pragma unroll
      for (int i = 0; i < D; ++i) Y[i] = H20[i] + Yacc[i];
    }
#pragma unroll
    for (int l = 0; l < D; ++l) ybuf[l * 32 + (lane ^ sdb_swz3(l))] = Y[l];
    __syncwarp();

#if SDB_OBSERVE
    obs.step(a, n, Y);
#endif
    if (n + 1 == next_save) {
      if (active) sdb_store<D>(a, p, slot, Y);
      ++slot;
      next_save += a.save_every;
    }
    if (n + 1 == steps && active) sdb_finish<D>(a, p, Y);
  }
#if SDB_OBSERVE
  if (active) obs.finish(a, p);
#endif
  sdb_tm_free512(tm_base);
}
#endif  // __CUDACC__

}  // namespace SDB_FUSED_NS
using namespace SDB_FUSED_NS;

#define SDB_FUSEDX_KERNEL_UD(NAME, D, M, WIK)                                                 \
  extern "C" __global__ void __launch_bounds__((SdbXShape<D, M>::THREADS), 1)                 \
      NAME(const __grid_constant__ SdbLoopArgs a,                                             \
           const __grid_constant__ SDB_FUSED_DRIFT_T::Params fp, const double* __restrict__ frag) { \
    extern __shared__ __align__(16) double sdb_fused_smem[];                                  \
    sdb_srk2_fusedx_body<D, M, WIK, SDB_FUSED_DRIFT_T::Params>(a, fp, frag, sdb_fused_smem);  \
  }
#define SDB_FUSEDX_KERNEL(NAME, D, M, WIK)                                                    \
  extern "C" __global__ void __launch_bounds__((SdbXShape<D, M>::THREADS), 1)                 \
      NAME(const __grid_constant__ SdbLoopArgs a, const __grid_constant__ PInline<D * D> fp,  \
           const double* __restrict__ frag) {                                                 \
    extern __shared__ __align__(16) double sdb_fused_smem[];                                  \
    sdb_srk2_fusedx_body<D, M, WIK>(a, fp, frag, sdb_fused_smem);                             \
  }
