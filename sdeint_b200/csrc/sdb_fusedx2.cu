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
                    SdbX2Shape<D, M>::smem_bytes(), sdb_fused_table_doubles<D, M>(),           \
                    &sdb_fused_build_tables<D, M>, SdbX2Shape<D, M>::THREADS,                  \
                    SdbX2Shape<D, M>::PATHS,                                                   \
                    WIK ? SDB_NOISE_WIK : SDB_NOISE_KPW, SDB_SCHEME_SRK2});

SDB_FUSEDX2_AOT(20, 20, 0, kpw, 21)
SDB_FUSEDX2_AOT(20, 20, 1, wik, 21)
