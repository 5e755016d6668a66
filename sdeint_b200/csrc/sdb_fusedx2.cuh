// sdb_fusedx2.cuh -- the TMEM-resident SRI2/SRS2 kernel of sdb_fusedx.cuh with TWO LANES PER PATH.
//
// At d = m = 20 the Levy areas of a path fill a whole tensor-memory lane (380 of 512 columns), so
// sdb_fusedx.cuh can keep only one warp per SM sub-partition resident -- and a single warp cannot hide
// its own latencies (ncu: issue slots 30 % busy, FP64 pipe 27 %).  Tensor memory is addressed by lane
// QUADRANT, not by warp: warps w and w + 4 of a block see the same 32 lanes.  Here such a pair of warps
// shares 32 paths and splits the work of every step between its two "roles", lane l of both warps
// working on path l of the pair:
//   noise    role 0 draws dW and X_k, role 1 draws Z_k = Y_k + sqrt(2/h) dW; they swap them through
//            shared memory, and each accumulates HALF of the area chunks (A_ij += X_i Z_j - Z_i X_j)
//            in the common TMEM lane; Wiktorsson's tail likewise by chunk halves (the partial
//            u = G^ dW vectors are summed through shared memory);
//   blocks   role r runs P1 / P2 / P3 of half of the row blocks; P2 streams ALL areas from TMEM;
//   final    the two partial stage-2 sums are added through shared memory, each role evaluates the
//            second drift and the new state for its half of the rows, and the state of the next step
//            is exchanged through the pair's shared-memory copy of Y.
// Two warps per sub-partition with the TMEM footprint of one path per lane.  Synchronisation: named
// barriers of 64 threads (n_terms + 4 per step, one more for Wiktorsson's tail), tcgen05 fences around them for the TMEM hand-over.
// Same Philox slots and the same arithmetic per entry as sdb_fusedx.cuh; only the stage-2 sum is
// added in a different order (two partial sums).
#pragma once
#include "sdb_fusedx.cuh"

namespace SDB_FUSED_NS {

template <int D, int M>
struct SdbX2Shape {
  typedef SdbFusedShape<D, M, SDB_FUSED_A2> Sh;
  typedef SdbXShape<D, M> Xs;
  static constexpr int NB = Xs::NB, NCH = Xs::NCH;
  static constexpr int NBH = NB / 2;     // row blocks of role 0
  static constexpr int NCHH = NCH / 2;   // area chunks of role 0
  static constexpr int DH = Xs::RN * NBH;  // state rows of role 0
  static_assert(NB >= 2 && NCH >= 2, "nothing to split");
  static constexpr int GR = Xs::GR >= 2 * M ? Xs::GR : 2 * M;  // staging rows of one warp (>= the X/Z swap)
  static constexpr int WARPS = 8, THREADS = 256, PATHS = 128;
  static SDB_CX size_t smem_bytes() {
    // (SDB_FUSED_A2 -1, any drift evaluated per lane: D more rows per warp for f(Y_n) h)
    size_t b = ((size_t)GR * 32 * WARPS + (size_t)D * 32 * 4 + (SDB_FUSED_A2 < 0 ? (size_t)D * 32 * WARPS : 0)) * 8 +
               128 * 16 + (SDB_FUSED_MAX_TERMS + 1 + Sh::n_left() + 1) * sizeof(double) + 16;
    return b < 120 * 1024 ? 120 * 1024 : b;  // > half an SM: one block (one TMEM allocation) per SM
  }
};

#ifdef __CUDACC__
// named barrier of one warp pair, with the fences that order tcgen05.st / ld across it
SDB_DEV void sdb_pair_sync(int bar_id) {
  sdb_tm_wait_st();
  sdb_tm_fence_before();
  asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
  sdb_tm_fence_after();
}

// A_ij += X_i Z_j - Z_i X_j for the area chunks [C0, C1) (read-modify-write in tensor memory)
template <int M, int C0, int C1>
SDB_DEV void sdb_x2_series_chunks(const double (&X)[M], const double (&Z)[M], uint32_t tm_at) {
  constexpr int MM = M * (M - 1) / 2, CH = 16;
  SdbTmLoad<CH> ca, cb;
  ca.issue(tm_at + 2 * CH * C0);
  ca.wait();
  sdb_static_for<C0, C1>([&](auto cc) {
    constexpr int c = decltype(cc)::value;
    constexpr bool odd = (c - C0) & 1;
    if constexpr (c + 1 < C1) {
      if constexpr (odd) ca.issue(tm_at + 2 * CH * (c + 1));
      else cb.issue(tm_at + 2 * CH * (c + 1));
    }
    double At[CH];
#pragma unroll
    for (int e = 0; e < CH; ++e) At[e] = odd ? cb.get(e) : ca.get(e);
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
    if constexpr (c + 1 < C1) {
      if constexpr (odd) ca.wait();
      else cb.wait();
    }
  });
  sdb_tm_wait_st();
}

// second pass of Wiktorsson's tail for the chunks [C0, C1): A_ij += c2 (dW_j u_i - dW_i u_j)
template <int M, int C0, int C1>
SDB_DEV void sdb_x2_tail2_chunks(const double (&dW)[M], const double (&u)[M], double c2, uint32_t tm_at) {
  constexpr int MM = M * (M - 1) / 2, CH = 16;
  sdb_static_for<C0, C1>([&](auto cc) {
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

// second drift evaluation and new state for the rows [I0, I1):  Y_i = H20_i + Yacc_i + (h/2) (A H20)_i
template <int D, int I0, int I1, bool A2, class FP>
SDB_DEV void sdb_x2_final_rows(const FP& fp, const double (&H20)[D], const double (&Yacc)[I1 - I0],
                               double h, double t1, double (&Y)[I1 - I0]) {
#if SDB_FUSED_A2 < 0
  double f1[D];  // any drift: f(H20, t_{n+1}) in full (each role of the pair evaluates it), this role's rows kept
  SDB_FUSED_DRIFT_T::eval(fp, H20, t1, f1);
#pragma unroll
  for (int i = I0; i < I1; ++i) Y[i - I0] = H20[i] + fma(0.5 * h, f1[i], Yacc[i - I0]);
  return;
#else
  (void)t1;
#endif
#pragma unroll
  for (int i = I0; i < I1; ++i) {
    if constexpr (A2) {
      Y[i - I0] = H20[i] + Yacc[i - I0];   // (h/2) f(H20) is in Yacc: the rows of A^2 ride in the P1 product
    } else {
      double acc = fp.v[i * D] * H20[0];
#pragma unroll
      for (int l = 1; l < D; ++l) acc = fma(fp.v[i * D + l], H20[l], acc);
      Y[i - I0] = H20[i] + fma(0.5 * h, acc, Yacc[i - I0]);
    }
  }
}

template <int D, int M, int WIK, class FP = PInline<D * D>>
SDB_DEV void sdb_srk2_fusedx2_body(const SdbLoopArgs& a, const FP& fp,
                                   const double* __restrict__ frag, double* smem) {
  typedef SdbFusedShape<D, M, SDB_FUSED_A2> Sh;
  typedef SdbXShape<D, M> Xs;
  typedef SdbX2Shape<D, M> X2;
  // the A^2 rows of the P1 table replace the second drift evaluation; with Wiktorsson's tail compiled in
  // ptxas allocates worse with it (cfg4: 3.10e8 against 3.21e8 path-steps/s), so that build keeps the loop
  constexpr bool USE_A2 = SDB_FUSED_A2 > 0 && WIK == 0;
  constexpr int MM = Xs::MM, CH = Xs::CH, NCH = Xs::NCH, RN = Xs::RN, NB = Xs::NB, LT = Xs::LT;
  constexpr int T1 = Xs::T1, KT3 = Xs::KT3, IT = Xs::IT, IL = Xs::IL;
  constexpr int NBH = X2::NBH, NCHH = X2::NCHH, DH = X2::DH, GR = X2::GR;
  static_assert(M % 2 == 0, "fused kernel needs an even number of Wiener processes");
  static_assert(D - DH <= DH + RN, "row halves");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pair = warp & 3, role = warp >> 2;
  const int fr = lane >> 2, fc = lane & 3;
  const int bar_id = 1 + pair;

  double* ybase = smem + (size_t)GR * 32 * X2::WARPS;
#if SDB_FUSED_A2 < 0
  double* fbuf = ybase + (size_t)D * 32 * 4 + (size_t)warp * D * 32;  // this warp's f(Y_n) h, one column per lane
  double2* tab = reinterpret_cast<double2*>(ybase + (size_t)D * 32 * 4 + (size_t)D * 32 * X2::WARPS);
#else
  double2* tab = reinterpret_cast<double2*>(ybase + (size_t)D * 32 * 4);
#endif
  double* s_hk = reinterpret_cast<double*>(tab + 128);
  double* s_left = s_hk + SDB_FUSED_MAX_TERMS + 1;
  uint32_t* tm_slot = reinterpret_cast<uint32_t*>(s_left + Sh::n_left() + 1);
  if (tid < 128) tab[tid] = reinterpret_cast<const double2*>(sdb_logtab)[tid];
  for (int i = tid; i <= SDB_FUSED_MAX_TERMS; i += (int)blockDim.x)
    s_hk[i] = i == 0 ? 0.0 : a.h_global * SDB_INV_2PI / (double)i;
  for (int i = tid; i < Sh::n_left(); i += X2::THREADS) s_left[i] = frag[(size_t)Sh::n_frags() * 32 + i];
  const uint32_t tm_base = sdb_tm_alloc512(tm_slot);  // contains a __syncthreads()
  const uint32_t tm = tm_base + (((uint32_t)pair * 32u) << 16);  // the pair's 32 lanes, all columns

  double* gbuf = smem + (size_t)warp * GR * 32;                    // this warp's staging rows
  double* gpart = smem + (size_t)(warp ^ 4) * GR * 32;             // the partner's
  double* g0 = smem + (size_t)pair * GR * 32;                      // role 0's, role 1's
  double* g1 = smem + (size_t)(pair + 4) * GR * 32;
  double* ybuf = ybase + (size_t)pair * D * 32;                    // the pair's copy of Y
  const int64_t p = (int64_t)blockIdx.x * X2::PATHS + pair * 32 + lane;
  const bool active = p < a.P;  // idle lanes of the last block still take part in the DMMAs
  const bool lead = role == 0;  // stores, status and observers are the business of role 0
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
      if (lead) sdb_store<D>(a, p, 0, Y);
    }
#if SDB_OBSERVE
    obs.init(a, Y);
#endif
    if (lead) {
#pragma unroll
      for (int l = 0; l < D; ++l) ybuf[l * 32 + (lane ^ sdb_swz3(l))] = Y[l];
    }
  }
  sdb_pair_sync(bar_id);

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
  const int c_lo = lead ? 0 : NCHH, c_hi = lead ? NCHH : NCH;  // this role's area chunks
  const int b_lo = lead ? 0 : NBH, b_hi = lead ? NBH : NB;     // this role's row blocks

#pragma unroll 1
  for (int n = 0; n < steps; ++n) {
    const double h = hs[n];
    const SdbGen gen(pkey, p_lo, p_hi, SDB_STEP(n));
    // ---- dW (wiener.py:52) by role 0; both roles clear their area chunks -----------------------------
    {
      if (lead) {
        double dW[M];
        sdb_gen_pairs<M / 2, 0, M, true, true>(gen, 0u, tab, rh_g, dW);
        sdb_tm_st<M>(tm + Xs::C_DW, dW);
      }
      double zero[CH];
#pragma unroll
      for (int e = 0; e < CH; ++e) zero[e] = 0.0;
#pragma unroll 1
      for (int c = c_lo; c < c_hi; ++c) sdb_tm_st<CH>(tm + Xs::C_AT + 2 * CH * c, zero);
    }
    sdb_pair_sync(bar_id);  // dW is in tensor memory
    // ---- Levy-area series (wiener.py:67-74, :102-105 / :264-267) -------------------------------------
    {
#pragma unroll 1
      for (int k = 1; k <= n_terms; ++k) {
        // role 0: X_k (pre-scaled by h / 2 pi k), role 1: Z_k -- the same generator code, other slots
        const uint32_t sx = (uint32_t)((M + 2 * (k - 1) * M) >> 1) + (lead ? 0u : (uint32_t)(M / 2));
        const int buf = (k & 1) * M;
        {
          double V[M];
          sdb_gen_pairs<M / 2, 0, M, true>(gen, sx, tab, lead ? s_hk[k] : 1.0, V);
          if (!lead) {
            double dW[M];
            sdb_tm_ld<M>(tm + Xs::C_DW, dW);
#pragma unroll
            for (int j = 0; j < M; ++j) V[j] = fma(cz, dW[j], V[j]);
          }
#pragma unroll
          for (int j = 0; j < M; ++j) gbuf[(buf + j) * 32 + lane] = V[j];
        }
        sdb_pair_sync(bar_id);
        double X[M], Z[M];
#pragma unroll
        for (int j = 0; j < M; ++j) {
          X[j] = g0[(buf + j) * 32 + lane];
          Z[j] = g1[(buf + j) * 32 + lane];
        }
        if (lead) sdb_x2_series_chunks<M, 0, NCHH>(X, Z, tm + Xs::C_AT);
        else sdb_x2_series_chunks<M, NCHH, NCH>(X, Z, tm + Xs::C_AT);
      }
      // ---- Wiktorsson's tail (wiener.py:269-277): chunk halves, u = G^ dW summed across the pair --------
      if (WIK) {
        double dW[M], u[M];
        sdb_tm_ld<M>(tm + Xs::C_DW, dW);
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
#pragma unroll 1
        for (int c = c_lo; c < c_hi; ++c) {
          double g[CH], At[CH];
          sdb_gen_pairs<CH / 2, 0, CH, false>(gen, st0 + (uint32_t)(c * CH / 2),
                                              tab, 1.0, g);
          sdb_tm_ld<CH>(tm + Xs::C_AT + 2 * CH * c, At);
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
          sdb_tm_st<CH>(tm + Xs::C_AT + 2 * CH * c, At);
        }
        // partial u through the swap rows the last series term did not use
        const int ub = ((n_terms + 1) & 1) * M;
#pragma unroll
        for (int j = 0; j < M; ++j) gbuf[(ub + j) * 32 + lane] = u[j];
        sdb_pair_sync(bar_id);
#pragma unroll
        for (int j = 0; j < M; ++j) u[j] = g0[(ub + j) * 32 + lane] + g1[(ub + j) * 32 + lane];
        if (lead) sdb_x2_tail2_chunks<M, 0, NCHH>(dW, u, c2, tm + Xs::C_AT);
        else sdb_x2_tail2_chunks<M, NCHH, NCH>(dW, u, c2, tm + Xs::C_AT);
      }
    }
    sdb_pair_sync(bar_id);  // all areas are in tensor memory; the staging rows are free again

#if SDB_FUSED_A2 < 0
    {  // any drift, evaluated one lane per path by BOTH roles (each needs the rows of its own blocks)
      double Yv[D], f0[D];
#pragma unroll
      for (int l = 0; l < D; ++l) Yv[l] = ybuf[l * 32 + (lane ^ sdb_swz3(l))];
      SDB_FUSED_DRIFT_T::eval(fp, Yv, a.tspan[n], f0);
#pragma unroll
      for (int l = 0; l < D; ++l) fbuf[l * 32 + lane] = f0[l] * h;
      __syncwarp();
    }
#endif
    // ---- this role's row blocks (runtime loop: every block has RN rows) ------------------------------
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
    for (int b = b_lo; b < b_hi; ++b) {
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
            const double f0h = ((const volatile double*)fbuf)[yr * 32 + lane];
#else
            const int nn = M * RN + r;
            const double f0h = gv[nn * 32 + (lane ^ sdb_swz(nn))] * h;
#endif
            const double yv = ybuf[yr * 32 + (lane ^ ((yr & 3) << 2))];
            piece[r] = yv + f0h;                      // H20
            if constexpr (USE_A2) {
              const int n2 = (M + 1) * RN + r;
              piece[RN + r] = fma(0.5 * h * h, gv[n2 * 32 + (lane ^ sdb_swz(n2))], GdW[r]);
            } else {
              piece[RN + r] = fma(-0.5, f0h, GdW[r]);   // Yacc
            }
          }
          sdb_tm_st<2 * RN>(tm + Xs::C_ROW + 4 * R0, piece);
        }
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

    // ---- this role's partial stage-2 sum back to one lane per path; the partner's half handed over ------
    double wp[D];
#pragma unroll
    for (int it = 0; it < IT; ++it) {
      double* row = gbuf + fr * 32;
      const int sw = sdb_swz(fr);
#pragma unroll
      for (int pt = 0; pt < 4; ++pt) {
        row[(8 * pt + 2 * fc) ^ sw] = w[it][pt][0];
        row[(8 * pt + 2 * fc + 1) ^ sw] = w[it][pt][1];
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) wp[8 * it + i] = gbuf[i * 32 + (lane ^ sdb_swz(i))];
      __syncwarp();
    }
#pragma unroll
    for (int q = 0; q < IL; ++q) wp[8 * IT + q] = wv[q];
    if (lead) {  // rows [DH, D) belong to role 1
#pragma unroll
      for (int i = DH; i < D; ++i) gbuf[(i - DH) * 32 + lane] = wp[i];
    } else {     // rows [0, DH) to role 0
#pragma unroll
      for (int i = 0; i < DH; ++i) gbuf[i * 32 + lane] = wp[i];
    }
    sdb_pair_sync(bar_id);  // row pieces in tensor memory, partial sums in shared memory

    // ---- Y_{n+1} = H20 - f0h/2 + f1h/2 + G dW + stage-2 sum, each role its half of the rows -------------
    {
      double H20[D];
      double rows[2 * D];
      sdb_tm_ld<2 * D>(tm + Xs::C_ROW, rows);
#pragma unroll
      for (int i = 0; i < D; ++i) H20[i] = rows[2 * RN * (i / RN) + i % RN];  // block i / RN: [H20 | Yacc]
      if (lead) {
        double Yacc[DH], Yn[DH];
#pragma unroll
        for (int i = 0; i < DH; ++i)
          Yacc[i] = rows[2 * RN * (i / RN) + RN + i % RN] + (wp[i] + gpart[i * 32 + lane]);
        sdb_x2_final_rows<D, 0, DH, USE_A2>(fp, H20, Yacc, h, a.tspan[n + 1], Yn);
        // (every reader of the old Y -- P1, the drift rows -- finished before the barrier above)
#pragma unroll
        for (int i = 0; i < DH; ++i) ybuf[i * 32 + (lane ^ sdb_swz3(i))] = Yn[i];
      } else {
        double Yacc[D - DH], Yn[D - DH];
#pragma unroll
        for (int i = DH; i < D; ++i)
          Yacc[i - DH] = rows[2 * RN * (i / RN) + RN + i % RN] + (wp[i] + gpart[(i - DH) * 32 + lane]);
        sdb_x2_final_rows<D, DH, D, USE_A2>(fp, H20, Yacc, h, a.tspan[n + 1], Yn);
#pragma unroll
        for (int i = DH; i < D; ++i) ybuf[i * 32 + (lane ^ sdb_swz3(i))] = Yn[i - DH];
      }
    }
    sdb_pair_sync(bar_id);  // the pair's copy of Y_{n+1} is complete

    const bool save = n + 1 == next_save;
    if (save) {
      ++slot;
      next_save += a.save_every;
    }
    if (lead && (save || n + 1 == steps || SDB_OBSERVE)) {
      double Y[D];
#pragma unroll
      for (int l = 0; l < D; ++l) Y[l] = ybuf[l * 32 + (lane ^ sdb_swz3(l))];
#if SDB_OBSERVE
      obs.step(a, n, Y);
#endif
      if (save && active) sdb_store<D>(a, p, slot - 1, Y);
      if (n + 1 == steps && active) sdb_finish<D>(a, p, Y);
    }
  }
#if SDB_OBSERVE
  if (lead && active) obs.finish(a, p);
#endif
  __syncthreads();
  sdb_tm_free512(tm_base);
}
#endif  // __CUDACC__

}  // namespace SDB_FUSED_NS
using namespace SDB_FUSED_NS;

#define SDB_FUSEDX2_KERNEL_UD(NAME, D, M, WIK)                                                \
  extern "C" __global__ void __launch_bounds__((SdbX2Shape<D, M>::THREADS), 1)                \
      NAME(const __grid_constant__ SdbLoopArgs a,                                             \
           const __grid_constant__ SDB_FUSED_DRIFT_T::Params fp, const double* __restrict__ frag) { \
    extern __shared__ __align__(16) double sdb_fused_smem[];                                  \
    sdb_srk2_fusedx2_body<D, M, WIK, SDB_FUSED_DRIFT_T::Params>(a, fp, frag, sdb_fused_smem); \
  }
#define SDB_FUSEDX2_KERNEL(NAME, D, M, WIK)                                                   \
  extern "C" __global__ void __launch_bounds__((SdbX2Shape<D, M>::THREADS), 1)                \
      NAME(const __grid_constant__ SdbLoopArgs a, const __grid_constant__ PInline<D * D> fp,  \
           const double* __restrict__ frag) {                                                 \
    extern __shared__ __align__(16) double sdb_fused_smem[];                                  \
    sdb_srk2_fusedx2_body<D, M, WIK>(a, fp, frag, sdb_fused_smem);                            \
  }
