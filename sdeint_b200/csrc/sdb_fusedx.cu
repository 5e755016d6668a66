// AOT instantiations of the TMEM-resident fused SRI2/SRS2 kernel (sdb_fusedx.cuh): d = m = 20
// (BASELINE.json configs[3]) with Ikpw/Jkpw and Iwik/Jwik, and d = m = 10 with Iwik/Jwik.
#define SDB_FUSED_VARIANT 20
#define SDB_FUSED_RB 2
#define SDB_FUSED_VOL_MMA 0   // one warp per sub-partition: let ptxas interleave the DMMA tiles
#include "sdb_fusedx.cuh"
#include "sdb_host.h"

#include <cuda_runtime.h>

#define SDB_FUSEDX_AOT(D, M, WIK, TAG, VARIANT)                                                       \
  SDB_FUSEDX_KERNEL(sdbk_fusedx_srk2_d##D##_m##M##_##TAG, D, M, WIK)                           \
  static SdbFusedRegistrar sdbreg_fusedx_d##D##_m##M##_##TAG(                                  \
      SdbFusedEntry{D, M, VARIANT, (const void*)sdbk_fusedx_srk2_d##D##_m##M##_##TAG,          \
                    SdbXShape<D, M>::smem_bytes(), sdb_fused_table_doubles<D, M, SDB_FUSED_A2>(),            \
                    &sdb_fused_build_tables<D, M, SDB_FUSED_A2>, SdbXShape<D, M>::THREADS,                   \
                    SdbXShape<D, M>::PATHS,                                                    \
                    WIK ? SDB_NOISE_WIK : SDB_NOISE_KPW, SDB_SCHEME_SRK2});

// d = m = 20: the default is the two-lanes-per-path kernel (sdb_fusedx2.cu); this one is variant 20
SDB_FUSEDX_AOT(20, 20, 0, kpw, 20)
SDB_FUSEDX_AOT(20, 20, 1, wik, 20)
SDB_FUSEDX_AOT(10, 10, 1, wik, 0)

// The run-time mirror used for NVRTC-compiled shapes (sdb_fused_rt.h) must agree with the templates.
#include "sdb_fused_rt.h"
#define SDB_CHECK_XDIMS(D, M)                                                                   \
  static_assert(SdbFusedXDims(D, M).smem_bytes() == SdbXShape<D, M>::smem_bytes() &&            \
                    SdbFusedXDims(D, M).C_END() == SdbXShape<D, M>::C_END &&                    \
                    SdbFusedXDims(D, M).threads() == SdbXShape<D, M>::THREADS &&                \
                    SdbFusedXDims(D, M).sh.n_frags() == SdbFusedShape<D, M, SDB_FUSED_A2>::n_frags() &&       \
                    SdbFusedXDims(D, M).sh.n_left() == SdbFusedShape<D, M, SDB_FUSED_A2>::n_left(),           \
                "sdb_fused_rt.h disagrees with SdbXShape");
SDB_CHECK_XDIMS(20, 20) SDB_CHECK_XDIMS(10, 10) SDB_CHECK_XDIMS(16, 12) SDB_CHECK_XDIMS(4, 6)
SDB_CHECK_XDIMS(12, 14) SDB_CHECK_XDIMS(6, 4)
