// AOT instantiations of the fused Euler-Maruyama / Heun kernel for linear systems (sdb_fusedeh.cuh).
#define SDB_FUSED_VARIANT 40
#include "sdb_fusedeh.cuh"
#include "sdb_fused_rt.h"
#include "sdb_host.h"

#include <cuda_runtime.h>

#define SDB_FUSEDEH_AOT(D, M, HEUN, TAG)                                                       \
  SDB_FUSEDEH_KERNEL(sdbk_fusedeh_##TAG##_d##D##_m##M, D, M, HEUN)                             \
  static SdbFusedRegistrar sdbreg_fusedeh_##TAG##_d##D##_m##M(                                 \
      SdbFusedEntry{D, M, 0, (const void*)sdbk_fusedeh_##TAG##_d##D##_m##M,                    \
                    SdbEhShape<D, M>::smem_bytes(HEUN), sdb_fused_table_doubles<D, M>(),       \
                    &sdb_fused_build_tables<D, M>, SDB_EH_THREADS(HEUN), SDB_EH_THREADS(HEUN), \
                    SDB_NOISE_KPW, HEUN ? SDB_SCHEME_HEUN : SDB_SCHEME_EULER});                \
  static_assert(SdbFusedEhDims(D, M, HEUN).smem_bytes() == SdbEhShape<D, M>::smem_bytes(HEUN), \
                "sdb_fused_rt.h disagrees with SdbEhShape");

SDB_FUSEDEH_AOT(10, 10, 0, euler)
SDB_FUSEDEH_AOT(10, 10, 1, heun)
static_assert(SdbFusedEhDims(5, 4, 0).smem_bytes() == SdbEhShape<5, 4>::smem_bytes(0) &&
                  SdbFusedEhDims(12, 6, 1).smem_bytes() == SdbEhShape<12, 6>::smem_bytes(1) &&
                  SdbFusedEhDims(3, 2, 1).threads() == SDB_EH_THREADS(1) &&
                  SdbFusedEhDims(3, 2, 0).threads() == SDB_EH_THREADS(0),
              "sdb_fused_rt.h disagrees with SdbEhShape");
