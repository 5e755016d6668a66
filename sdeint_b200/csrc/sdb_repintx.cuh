// sdb_repintx.cuh -- standalone Ikpw/Jkpw/Iwik/Jwik (sdeint/wiener.py:77-131, :234-305) for LARGE m
// with on-device Philox normals: one lane per (path, step) interval, the m(m-1)/2 Levy areas of the
// interval in TENSOR MEMORY (sdb_x_levy_areas of sdb_fusedx.cuh: series read-modify-written in
// chunks, two-pass O(m^2) Wiktorsson tail), then I (or J) and A (or Atilde) assembled entry by entry
// and stored.  The one-thread-per-interval kernel of sdb_kernels.cuh keeps the areas in a local
// array; at m = 20 that is 190 doubles per thread in local memory and 7-11 % of the HBM rate.
// Host-supplied normals (construction parity with the reference's draws) stay on that kernel.
#pragma once
#ifndef SDB_FUSED_RB
#define SDB_FUSED_RB 2
#endif
#include "sdb_fusedx.cuh"

namespace SDB_FUSED_NS {

template <int M>
struct SdbRxShape {
  static constexpr int MM = M * (M - 1) / 2, CH = 16, NCH = (MM + CH - 1) / CH, MMP = NCH * CH;
  static constexpr int C_AT = 0, C_DW = 2 * MMP, C_END = C_DW + (2 * M + 7) / 8 * 8;
  static_assert(C_END <= 512, "tensor-memory columns of a lane exceeded");
  static constexpr int WPQ = 2 * C_END <= 512 ? 2 : 1;       // warps per SM sub-partition
  static constexpr int WARPS = 4 * WPQ, THREADS = 32 * WARPS, C_STRIDE = 512 / WPQ;
  static SDB_CX size_t smem_bytes() { return 120 * 1024; }  // > half an SM: one TMEM allocation per SM
};

#ifdef __CUDACC__
template <int M, int WIK>
SDB_DEV void sdb_repintx_body(const SdbRepintArgs& a, double* smem) {
  typedef SdbRxShape<M> Rs;
  constexpr int MM = Rs::MM, CH = Rs::CH, NCH = Rs::NCH;
  static_assert(M % 2 == 0 && M >= 2, "even m");
  const int tid = threadIdx.x, warp = tid >> 5;
  double2* tab = reinterpret_cast<double2*>(smem);
  double* s_hk = reinterpret_cast<double*>(tab + 128);
  uint32_t* tm_slot = reinterpret_cast<uint32_t*>(s_hk + SDB_FUSED_MAX_TERMS + 1);
  if (tid < 128) tab[tid] = reinterpret_cast<const double2*>(sdb_logtab)[tid];
  for (int i = tid; i <= SDB_FUSED_MAX_TERMS; i += (int)blockDim.x)
    s_hk[i] = i == 0 ? 0.0 : a.h * SDB_INV_2PI / (double)i;
  const uint32_t tm_base = sdb_tm_alloc512(tm_slot);  // contains a __syncthreads()
  const uint32_t tm = tm_base + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)((warp >> 2) * Rs::C_STRIDE);

  const int64_t total = a.steps * a.paths;
  int64_t idx = (int64_t)blockIdx.x * Rs::THREADS + tid;
  const bool active = idx < total;  // idle lanes of the last block recompute the last interval
  if (!active) idx = total - 1;
  const int64_t n = idx / a.paths;
  const int64_t p = idx - n * a.paths;
  double dW[M];
  {
    const double* w = a.dW + n * a.dW_sN + p * a.dW_sP;
#pragma unroll
    for (int j = 0; j < M; ++j) dW[j] = w[j * a.dW_sJ];
    sdb_tm_st<M>(tm + Rs::C_DW, dW);
    double zero[CH];
#pragma unroll
    for (int e = 0; e < CH; ++e) zero[e] = 0.0;
#pragma unroll
    for (int c = 0; c < NCH; ++c) sdb_tm_st<CH>(tm + Rs::C_AT + 2 * CH * c, zero);
    sdb_tm_wait_st();
  }
  {
    const uint64_t path = (uint64_t)(a.path_id0 + p);
    const SdbPhiloxKey pkey(a.seed);
    const SdbGen gen(pkey, (uint32_t)path, (uint32_t)(path >> 32), (uint32_t)n);
    const double tail_scale = WIK ? a.h * SDB_INV_2PI * sqrt(sdb_a_tail(a.n_terms)) : 0.0;
    sdb_x_levy_areas<M, WIK>(gen, tab, s_hk, a.h,
                             sqrt(2.0 / a.h), tail_scale, a.n_terms, tm + Rs::C_AT, tm + Rs::C_DW);
  }
  // ---- assemble and store: KPW places +A above the diagonal (wiener.py:106), Wiktorsson's
  // column-major unvec below (wiener.py:278-280) -----------------------------------------------------
  double* q = a.IJ + n * a.IJ_sN + p * a.IJ_sP;
  double* o = a.A ? a.A + n * a.A_sN + p * a.A_sP : nullptr;
  const bool strat = a.stratonovich != 0;
  if (active) {
#pragma unroll
    for (int i = 0; i < M; ++i) {
      q[i * a.IJ_sR + i * a.IJ_sC] = strat ? 0.5 * dW[i] * dW[i] : 0.5 * (dW[i] * dW[i] - a.h);
      if (!WIK && o) o[(i * M + i) * a.A_sE] = 0.0;
    }
  }
  sdb_static_for<0, NCH>([&](auto cc) {
    constexpr int c = decltype(cc)::value;
    double At[CH];
    sdb_tm_ld<CH>(tm + Rs::C_AT + 2 * CH * c, At);
    if (active) {
      sdb_static_for<0, CH>([&](auto ee) {
        constexpr int e = decltype(ee)::value;
        constexpr int r = CH * c + e;
        if constexpr (r < MM) {
          constexpr int i = sdb_pair_j<M>(r), j = sdb_pair_k<M>(r);
          const double av = At[e];
          const double sym = 0.5 * dW[i] * dW[j];
          const double sg = WIK ? -av : av;
          q[i * a.IJ_sR + j * a.IJ_sC] = sym + sg;
          q[j * a.IJ_sR + i * a.IJ_sC] = sym - sg;
          if (o) {
            if (WIK) {
              o[r * a.A_sE] = av;
            } else {
              o[(i * M + j) * a.A_sE] = av;
              o[(j * M + i) * a.A_sE] = -av;
            }
          }
        }
      });
    }
  });
  sdb_tm_free512(tm_base);
}
#endif  // __CUDACC__

}  // namespace SDB_FUSED_NS
using namespace SDB_FUSED_NS;

#define SDB_REPINTX_KERNEL(NAME, M, WIK)                                                       \
  extern "C" __global__ void __launch_bounds__((SdbRxShape<M>::THREADS), 1)                    \
      NAME(const __grid_constant__ SdbRepintArgs a) {                                          \
    extern __shared__ __align__(16) double sdb_fused_smem[];                                   \
    sdb_repintx_body<M, WIK>(a, sdb_fused_smem);                                               \
  }
