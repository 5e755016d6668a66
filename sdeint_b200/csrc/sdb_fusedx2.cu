// AOT instantiations of the two-lanes-per-path TMEM kernel (sdb_fusedx2.cuh): d = m = 20
// (BASELINE.json configs[3]) with Ikpw/Jkpw and Iwik/Jwik.
#define SDB_FUSED_VARIANT 21
#define SDB_FUSED_RB 2
#define SDB_FUSED_VOL_MMA 0
#include "sdb_fusedx2.cuh"
#include "sdb_host.h"

#include <cuda_runtime.h>

#define SDB_FUSEDX2_AOT(D, M, WIK, TAG, VARIANT)                                               \
  SDB_FUSEDX2_KERNEL(sdbk_fusedx2_srk2_d##D##_m##M##_##TAG, D, M, WIK)                         \
  static SdbFusedRegistrar sdbreg_fusedx2_d##D##_m##M##_##TAG(                                 \
      SdbFusedEntry{D, M, VARIANT, (const void*)sdbk_fusedx2_srk2_d##D##_m##M##_##TAG,         \
                    SdbX2Shape<D, M>::smem_bytes(), sdb_fused_table_doubles<D, M, SDB_FUSED_A2>(),           \
                    &sdb_fused_build_tables<D, M, SDB_FUSED_A2>, SdbX2Shape<D, M>::THREADS,                  \
                    SdbX2Shape<D, M>::PATHS,                                                   \
                    WIK ? SDB_NOISE_WIK : SDB_NOISE_KPW, SDB_SCHEME_SRK2});

SDB_FUSEDX2_AOT(20, 20, 0, kpw, 0)
SDB_FUSEDX2_AOT(20, 20, 1, wik, 0)

// The run-time mirror used for NVRTC-compiled shapes (sdb_fused_rt.h) must agree with the templates.
#include "sdb_fused_rt.h"
#define SDB_CHECK_X2DIMS(D, M)                                                                  \
  static_assert(SdbFusedX2Dims(D, M).smem_bytes() == SdbX2Shape<D, M>::smem_bytes() &&          \
                    SdbFusedX2Dims(D, M).supported(),                                           \
                "sdb_fused_rt.h disagrees with SdbX2Shape");
SDB_CHECK_X2DIMS(20, 20) SDB_CHECK_X2DIMS(8, 16) SDB_CHECK_X2DIMS(12, 18) SDB_CHECK_X2DIMS(16, 16)
