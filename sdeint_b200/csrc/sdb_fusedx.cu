// AOT instantiations of the TMEM-resident fused SRI2/SRS2 kernel (sdb_fusedx.cuh): d = m = 20
// (BASELINE.json configs[3]) with Ikpw/Jkpw and Iwik/Jwik, and d = m = 10 with Iwik/Jwik.
#define SDB_FUSED_VARIANT 20
#define SDB_FUSED_RB 2
#define SDB_FUSED_VOL_MMA 0   // one warp per sub-partition: let ptxas interleave the DMMA tiles
#include "sdb_fusedx.cuh"
#include "sdb_host.h"

#include <cuda_runtime.h>

#define SDB_FUSEDX_AOT(D, M, WIK, TAG)                                                         \
  SDB_FUSEDX_KERNEL(sdbk_fusedx_srk2_d##D##_m##M##_##TAG, D, M, WIK)                           \
  static SdbFusedRegistrar sdbreg_fusedx_d##D##_m##M##_##TAG(                                  \
      SdbFusedEntry{D, M, 0, (const void*)sdbk_fusedx_srk2_d##D##_m##M##_##TAG,                \
                    SdbXShape<D, M>::smem_bytes(), sdb_fused_table_doubles<D, M>(),            \
                    &sdb_fused_build_tables<D, M>, SdbXShape<D, M>::THREADS,                   \
                    SdbXShape<D, M>::PATHS,                                                    \
                    WIK ? SDB_NOISE_WIK : SDB_NOISE_KPW});

SDB_FUSEDX_AOT(20, 20, 0, kpw)
SDB_FUSEDX_AOT(20, 20, 1, wik)
SDB_FUSEDX_AOT(10, 10, 1, wik)
