// sdb_fusedeh.cuh -- fused itoEuler / stratHeun (sdeint/integrate.py:235-239, :292-302) for LINEAR
// systems f = A y, g_k = B_k y with on-device Philox increments: the whole time loop in one launch.
//
// The generic thread-per-path loop feeds every DFMA of B_k y with a warp-uniform matrix element and
// reaches ~17 % of the FP64 pipe at d = m = 10 (profiles/r01_microbench.md: operand delivery, not
// arithmetic, is the limit).  Here the products with the shared matrices run as FP64 tensor
// instructions exactly as in sdb_fused.cuh (P1: [B_j[R,:]; A[R,:]] Y -> staging rows x 32 paths,
// same fragment tables), and one lane per path contracts its G_n rows with its own dW:
//     Euler   Y += f h + G_n dW
//     Heun    Yb = Y + f h + G_n dW;   Y += (f(Y) + f(Yb)) h/2 + (G(Y) + G(Yb)) dW / 2
// No Levy areas, no stage-2 sum: Euler fits 168 registers, so 12 warps per SM (3 per sub-partition).
#pragma once
#include "sdb_fused.cuh"

#ifndef SDB_EH_XR
#define SDB_EH_XR 0  // -1 (NVRTC builds): ANY drift with linear diffusion g_k = B_k y -- no drift rows in the
#endif               //    tables, the functor SDB_FUSED_DRIFT_T is evaluated one lane per path (sdb_fused.cuh)
// Euler: 12 warps per SM at 168 registers; Heun holds two evaluations: 8 warps at 255 registers
#define SDB_EH_THREADS(HEUN) ((HEUN) ? 256 : 384)

namespace SDB_FUSED_NS {

template <int D, int M>
struct SdbEhShape {
  typedef SdbFusedShape<D, M, SDB_EH_XR> Sh;
  static SDB_CX int grows() {  // staging rows: the widest P1 block, at least the state
    int g = Sh::max2(8, D);
    for (int b = 0; b < Sh::NB; ++b) g = Sh::max2(g, 8 * Sh::t1(b));
    return g;
  }
  static SDB_CX size_t smem_bytes(int heun) {
    return (size_t)grows() * 32 * 8 * (SDB_EH_THREADS(heun) / 32) + 128 * 16;
  }
};

#ifdef __CUDACC__
// G_n dW and f for all rows of the state at Yin: DMMA per row block, contraction per lane.
template <int D, int M>
SDB_DEV void sdb_eh_eval(const double* __restrict__ frag, double* gbuf, const double (&Yin)[D],
                         const double (&dW)[M], int lane, double (&GdW)[D], double (&f)[D]) {
  typedef SdbFusedShape<D, M, SDB_EH_XR> Sh;
  constexpr int LT = Sh::LT;
  const int fr = lane >> 2, fc = lane & 3;
  double Yf[LT][4];
#pragma unroll
  for (int l = 0; l < D; ++l) gbuf[l * 32 + (lane ^ sdb_swz(l))] = Yin[l];
  __syncwarp();
#pragma unroll
  for (int lt = 0; lt < LT; ++lt) {
    const int l = 4 * lt + fc;
    const bool ok = (4 * lt + 3 < D) || (l < D);
    const int lc = ok ? l : 0;
#pragma unroll
    for (int pt = 0; pt < 4; ++pt) {
      const double v = gbuf[lc * 32 + ((8 * pt + fr) ^ sdb_swz(lc))];
      Yf[lt][pt] = ok ? v : 0.0;
    }
  }
  __syncwarp();
  sdb_static_for<0, Sh::NB>([&](auto bb) {
    constexpr int b = decltype(bb)::value;
    constexpr int RN = Sh::rn(b), R0 = Sh::r0(b), T1 = Sh::t1(b);
    const double* f1 = frag + (size_t)Sh::p1_off(b) * 32 + lane;
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
    __syncwarp();
    const volatile double* gv = gbuf;
#pragma unroll
    for (int r = 0; r < RN; ++r) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < M; ++j) {
        const int n = j * RN + r;
        acc = fma(gv[n * 32 + (lane ^ sdb_swz(n))], dW[j], acc);
      }
      GdW[R0 + r] = acc;
#if SDB_EH_XR >= 0
      const int n = M * RN + r;
      f[R0 + r] = gv[n * 32 + (lane ^ sdb_swz(n))];
#endif
    }
    __syncwarp();
  });
}

template <int D, int M, int HEUN, class FP = int>
SDB_DEV void sdb_fusedeh_body(const SdbLoopArgs& a, const double* __restrict__ frag, double* smem,
                              const FP& fp = FP()) {
  typedef SdbEhShape<D, M> Es;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double2* tab = reinterpret_cast<double2*>(smem + (size_t)Es::grows() * 32 * (SDB_EH_THREADS(HEUN) / 32));
  if (tid < 128) tab[tid] = reinterpret_cast<const double2*>(sdb_logtab)[tid];
  __syncthreads();
  double* gbuf = smem + (size_t)warp * Es::grows() * 32;
  const int64_t p = (int64_t)blockIdx.x * SDB_EH_THREADS(HEUN) + tid;
  const bool active = p < a.P;  // idle lanes of the last warp still take part in the DMMAs
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
  const double h = a.h_global;  // Euler / Heun use the global mean step (integrate.py:228, :285)
  const double rh = sqrt(h);
  const uint64_t path = (uint64_t)(a.path_id0 + p);
  const uint32_t p_lo = (uint32_t)path, p_hi = (uint32_t)(path >> 32);
  const SdbPhiloxKey pkey(a.seed);
  const int steps = a.N - 1;
  int next_save = a.save_every;
  int64_t slot = 1;
#pragma unroll 1
  for (int n = 0; n < steps; ++n) {
    double dW[M];
    const SdbGen gen(pkey, p_lo, p_hi, SDB_STEP(n), M > 2);
    if constexpr (M % 2 == 0) {
      uint32_t w[M / 2][4];
      gen.template blocks<M / 2, true>(0u, w);
      sdb_normal_batch<M / 2, true>(w, tab, rh, dW);
    } else {
      sdb_normals_m<M, 0, true, true>(gen, 0u, tab, rh, dW);
    }
    double GdW[D], f0[D];
    sdb_eh_eval<D, M>(frag, gbuf, Y, dW, lane, GdW, f0);
#if SDB_EH_XR < 0
    SDB_FUSED_DRIFT_T::eval(fp, Y, a.tspan[n], f0);  // f(Y_n, t_n): integrate.py:236, :295
#endif
    if (!HEUN) {
#pragma unroll
      for (int i = 0; i < D; ++i) Y[i] = fma(f0[i], h, Y[i]) + GdW[i];  // integrate.py:239
    } else {
      double Yb[D], GdW1[D], f1[D];
#pragma unroll
      for (int i = 0; i < D; ++i) Yb[i] = fma(f0[i], h, Y[i]) + GdW[i];  // integrate.py:299
      sdb_eh_eval<D, M>(frag, gbuf, Yb, dW, lane, GdW1, f1);
#if SDB_EH_XR < 0
      SDB_FUSED_DRIFT_T::eval(fp, Yb, a.tspan[n + 1], f1);  // f(Ybar, t_{n+1}): integrate.py:300
#endif
#pragma unroll
      for (int i = 0; i < D; ++i)  // integrate.py:300-302
        Y[i] = Y[i] + 0.5 * (f0[i] + f1[i]) * h + 0.5 * (GdW[i] + GdW1[i]);
    }
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
}
#endif  // __CUDACC__

}  // namespace SDB_FUSED_NS
using namespace SDB_FUSED_NS;

#define SDB_FUSEDEH_KERNEL_UD(NAME, D, M, HEUN)                                               \
  extern "C" __global__ void __launch_bounds__(SDB_EH_THREADS(HEUN), 1)                       \
      NAME(const __grid_constant__ SdbLoopArgs a,                                             \
           const __grid_constant__ SDB_FUSED_DRIFT_T::Params fp, const double* __restrict__ frag) { \
    extern __shared__ __align__(16) double sdb_fused_smem[];                                  \
    sdb_fusedeh_body<D, M, HEUN, SDB_FUSED_DRIFT_T::Params>(a, frag, sdb_fused_smem, fp);      \
  }
#define SDB_FUSEDEH_KERNEL(NAME, D, M, HEUN)                                                  \
  extern "C" __global__ void __launch_bounds__(SDB_EH_THREADS(HEUN), 1)                       \
      NAME(const __grid_constant__ SdbLoopArgs a, const double* __restrict__ frag) {          \
    extern __shared__ __align__(16) double sdb_fused_smem[];                                  \
    sdb_fusedeh_body<D, M, HEUN>(a, frag, sdb_fused_smem);                                    \
  }
