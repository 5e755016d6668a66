// AOT instantiations of the fused linear SRI2/SRS2 kernel (sdb_fused.cuh).
#include "sdb_fused.cuh"
#include "sdb_host.h"

#include <cuda_runtime.h>

#define SDB_FUSED_AOT(D, M)                                                                   \
  SDB_FUSED_KERNEL(sdbk_fused_srk2_d##D##_m##M, D, M)                                         \
  static SdbAotRegistrar sdbreg_fused_d##D##_m##M(                                            \
      sdb_fused_key(D, M), (const void*)sdbk_fused_srk2_d##D##_m##M,                          \
      SdbFusedShape<D, M>::smem_bytes());

SDB_FUSED_AOT(10, 10)   // BASELINE.json configs[2] and [4]
SDB_FUSED_AOT(4, 4)     // small case for tests (d even, m even)
SDB_FUSED_AOT(5, 4)     // Roessler (7.4) dimensions: odd d exercises D1 != D2
