// sdb_fused.cuh -- the production SRI2/SRS2 kernel for linear systems f = A y, g_k = B_k y with
// on-device Philox + Ikpw/Jkpw noise: the whole time loop of sdeint/integrate.py:494-522 plus
// wiener.py:52 and :67-106 in ONE launch, one thread per sample path.
//
// FP64-pipe budget per path-step at d = m = 10, n = 5 (DESIGN.md "fused SRK2 kernel"):
//   55 Philox blocks -> 110 normals (32 FP64 ops per pair)            1760
//   Levy-area series, 2 DFMA per (pair, term) + Z = Y + c dW            500
//   drift twice, G_n (m matvecs), G_n I, H2/H3, 2m matvecs, updates    4900
// against 5140 DFMA-equivalents (10 280 flop) of algorithmic work.
//
// Where the data lives.  A and B_k are kernel parameters (constant bank): every loop over
// them is fully unrolled so each element is an immediate constant-bank operand of a DFMA.
// Per-path state that does not fit the 255-register budget at once is staged in shared
// memory, laid out [element][thread] so a warp's access is conflict-free:
//   * the m(m-1)/2 Levy areas (read twice, once per row block of G_n I);
//   * rows [0, d/2) of sum1 = G_n I (d/2 x m), while rows [d/2, d) stay in registers.
// G_n itself is never stored: column j (half of its rows at a time) is produced, folded into
// G_n dW and rank-1-accumulated into sum1, using
//   I[j][k] = dW_j dW_k / 2 + A[j][k] - (h/2) delta_jk   (Ito; no delta term for Stratonovich)
//   => sum1[:,k] sqrt(h) = (dW_k/2) (G_n dW) + sum_j G_n[:,j] (A[j][k] - (h/2) delta_jk).
// 760 B of shared memory per thread at d = m = 10 -> 256 threads (8 warps, 2 per scheduler)
// per SM.
#pragma once
#include "sdb_kernels.cuh"

#define SDB_FUSED_THREADS 256
#define SDB_FUSED_MAX_TERMS 64

template <int D, int M>
struct SdbFusedShape {
  static constexpr int MM = M * (M - 1) / 2;
  static constexpr int D1 = D / 2;       // rows of sum1 staged in shared memory
  static constexpr int D2 = D - D1;      // rows kept in registers
  static constexpr int ROWS = MM + D1 * M;
  static constexpr size_t smem_bytes() {
    return (size_t)ROWS * SDB_FUSED_THREADS * sizeof(double) + 128 * sizeof(double2) +
           (SDB_FUSED_MAX_TERMS + 1) * sizeof(double);
  }
};

// index of pair (i < j) in K_m row order
template <int M>
SDB_DEV constexpr int sdb_pair_index(int i, int j) {
  return i * M - i * (i + 1) / 2 + (j - i - 1);
}

// One row block of   G_n dW   and   sum1 sqrt(h) = G_n I sqrt(h)   (rows R0 .. R0+RN-1).
template <int D, int M, int R0, int RN>
SDB_DEV void sdb_fused_rowblock(const PInline<M * D * D>& gp, const double (&Y)[D],
                                const double (&dW)[M], const volatile double* sA, double dg,
                                double (&Yn)[D], double (&s)[RN][M]) {
  constexpr int NT = SDB_FUSED_THREADS;
  double GdW[RN];
#pragma unroll
  for (int i = 0; i < RN; ++i) {
    GdW[i] = 0.0;
#pragma unroll
    for (int k = 0; k < M; ++k) s[i][k] = 0.0;
  }
#pragma unroll
  for (int j = 0; j < M; ++j) {
    double g[RN];
#pragma unroll
    for (int i = 0; i < RN; ++i) {
      double acc = gp.v[(j * D + R0 + i) * D] * Y[0];
#pragma unroll
      for (int l = 1; l < D; ++l) acc = fma(gp.v[(j * D + R0 + i) * D + l], Y[l], acc);
      g[i] = acc;
      GdW[i] = fma(acc, dW[j], GdW[i]);
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
      for (int i = 0; i < RN; ++i) s[i][k] = fma(g[i], a, s[i][k]);
    }
  }
#pragma unroll
  for (int k = 0; k < M; ++k) {
    const double hd = 0.5 * dW[k];
#pragma unroll
    for (int i = 0; i < RN; ++i) s[i][k] = fma(hd, GdW[i], s[i][k]);
  }
#pragma unroll
  for (int i = 0; i < RN; ++i) Yn[R0 + i] += GdW[i];
}

template <int D, int M>
SDB_DEV void sdb_srk2_fused_linear_body(const SdbLoopArgs& a, const PInline<D * D>& fp,
                                        const PInline<M * D * D>& gp, double* smem) {
  typedef SdbFusedShape<D, M> Sh;
  constexpr int NT = SDB_FUSED_THREADS;
  constexpr int MM = Sh::MM, D1 = Sh::D1, D2 = Sh::D2;
  static_assert(M % 2 == 0, "fused kernel needs an even number of Wiener processes");
  const int tid = threadIdx.x;

  // shared: [ROWS][NT] per-thread staging | log table | h/(2 pi k)
  double2* tab = reinterpret_cast<double2*>(smem + (size_t)Sh::ROWS * NT);
  double* s_hk = reinterpret_cast<double*>(tab + 128);
  if (tid < 128) tab[tid] = reinterpret_cast<const double2*>(sdb_logtab)[tid];
  if (tid <= SDB_FUSED_MAX_TERMS)
    s_hk[tid] = tid == 0 ? 0.0 : a.h_global * SDB_INV_2PI / (double)tid;
  __syncthreads();
  const int64_t p = (int64_t)blockIdx.x * NT + tid;
  if (p >= a.P) return;
  double* sA = smem + tid;
  double* sS = smem + (size_t)MM * NT + tid;

  double Y[D];
  if (a.flags & SDB_FLAG_Y0_BROADCAST) {
#pragma unroll
    for (int i = 0; i < D; ++i) Y[i] = a.y0[i];
  } else {
#pragma unroll
    for (int i = 0; i < D; ++i) Y[i] = a.y0[i * a.y0_sI + p * a.y0_sP];
  }
  sdb_store<D>(a, p, 0, Y);

  const double rh_g = sqrt(a.h_global);
  const double cz = sqrt(2.0 / a.h_global);
  const double dg = (a.flags & SDB_FLAG_STRATONOVICH) ? 0.0 : -0.5 * a.h_global;
  const uint64_t path = (uint64_t)(a.path_id0 + p);
  const int steps = a.N - 1;
  const int n_terms = a.n_terms;
  const double* __restrict__ hs = a.tspan + a.N;  // per-step h, sqrt(h), 1/sqrt(h) (host-built)
  int next_save = a.save_every;
  int64_t slot = 1;

#pragma unroll 1
  for (int n = 0; n < steps; ++n) {
    const double h = hs[n];
    const double rh = hs[steps + n];
    const double inv_rh = hs[2 * steps + n];

    // ---- noise: dW (wiener.py:52), Levy areas (wiener.py:67-74, :102-105) -------------------------
    const SdbStream st(a.seed, path, (uint32_t)n, tab);
    double dW[M];
#pragma unroll
    for (int s = 0; s < M / 2; ++s) {
      double z0, z1;
      st.pair((uint32_t)s, z0, z1);
      dW[2 * s] = rh_g * z0;
      dW[2 * s + 1] = rh_g * z1;
    }
    {
      double At[MM];
#pragma unroll
      for (int r = 0; r < MM; ++r) At[r] = 0.0;
#pragma unroll 1
      for (int k = 1; k <= n_terms; ++k) {
        double X[M], Z[M];
        const uint32_t sx = (uint32_t)((M + 2 * (k - 1) * M) >> 1);
        const double hk = s_hk[k];
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
          Z[2 * s] = fma(cz, dW[2 * s], z0);
          Z[2 * s + 1] = fma(cz, dW[2 * s + 1], z1);
        }
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
          for (int j = i + 1; j < M; ++j) {
            const int r = sdb_pair_index<M>(i, j);
            At[r] = fma(X[i], Z[j], At[r]);
            At[r] = fma(-Z[i], X[j], At[r]);
          }
      }
#pragma unroll
      for (int r = 0; r < MM; ++r) sA[r * NT] = At[r];
    }

    // ---- drift stages (integrate.py:502, :509, :514-515) --------------------------------------------
    double H20[D], Yn[D];
    {
      double f0h[D];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double acc = fp.v[i * D] * Y[0];
#pragma unroll
        for (int l = 1; l < D; ++l) acc = fma(fp.v[i * D + l], Y[l], acc);
        f0h[i] = acc * h;
        H20[i] = Y[i] + f0h[i];
      }
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double acc = fp.v[i * D] * H20[0];
#pragma unroll
        for (int l = 1; l < D; ++l) acc = fma(fp.v[i * D + l], H20[l], acc);
        Yn[i] = Y[i] + 0.5 * fma(acc, h, f0h[i]);
      }
    }

    // ---- G_n dW and sum1 = G_n I (integrate.py:503-508), two row blocks -----------------------------
    {
      double s1[D1][M];
      sdb_fused_rowblock<D, M, 0, D1>(gp, Y, dW, sA, dg, Yn, s1);
#pragma unroll
      for (int i = 0; i < D1; ++i)
#pragma unroll
        for (int k = 0; k < M; ++k) sS[(i * M + k) * NT] = s1[i][k];
    }
    double s2[D2][M];
    sdb_fused_rowblock<D, M, D1, D2>(gp, Y, dW, sA, dg, Yn, s2);

    // ---- stage-2 diffusion evaluations at H2_k, H3_k (integrate.py:509-521) -------------------------
    const double half_rh = 0.5 * rh;
#pragma unroll
    for (int k = 0; k < M; ++k) {
      double H2[D], H3[D];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const double sv = i < D1 ? sS[(i * M + k) * NT] : s2[i < D1 ? 0 : i - D1][k];
        H2[i] = fma(sv, inv_rh, H20[i]);
        H3[i] = fma(-sv, inv_rh, H20[i]);
      }
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double u = gp.v[(k * D + i) * D] * H2[0];
        double v = gp.v[(k * D + i) * D] * H3[0];
#pragma unroll
        for (int l = 1; l < D; ++l) {
          u = fma(gp.v[(k * D + i) * D + l], H2[l], u);
          v = fma(gp.v[(k * D + i) * D + l], H3[l], v);
        }
        Yn[i] = fma(half_rh, u - v, Yn[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < D; ++i) Y[i] = Yn[i];

    if (n + 1 == next_save) {
      sdb_store<D>(a, p, slot, Y);
      ++slot;
      next_save += a.save_every;
    }
  }
  sdb_finish<D>(a, p, Y);
}

#define SDB_FUSED_KERNEL(NAME, D, M)                                                          \
  extern "C" __global__ void __launch_bounds__(SDB_FUSED_THREADS, 1)                          \
      NAME(const __grid_constant__ SdbLoopArgs a, const __grid_constant__ PInline<D * D> fp,  \
           const __grid_constant__ PInline<M * D * D> gp) {                                   \
    extern __shared__ __align__(16) double sdb_fused_smem[];                                  \
    sdb_srk2_fused_linear_body<D, M>(a, fp, gp, sdb_fused_smem);                              \
  }
