// AOT instantiation of the warp-specialised fused SRI2/SRS2 kernel (sdb_fusedws.cuh): 8 main +
// 8 helper warps per SM, normals handed over through tensor memory.  Registered as tuning
// variant 10 of the fused kernel (selected with SDB_FUSED_VARIANT=10).
#define SDB_FUSED_VARIANT 10
#define SDB_FUSED_RB 2
#define SDB_WS_HELPERS 1
#include "sdb_fusedws.cuh"
#include "sdb_host.h"

#include <cuda_runtime.h>

#define SDB_FUSEDWS_AOT(D, M)                                                                  \
  SDB_FUSEDWS_KERNEL(sdbk_fusedws_srk2_d##D##_m##M, D, M)                                      \
  static SdbFusedRegistrar sdbreg_fusedws_d##D##_m##M(                                         \
      SdbFusedEntry{D, M, SDB_FUSED_VARIANT, (const void*)sdbk_fusedws_srk2_d##D##_m##M,       \
                    SdbWsShape<D, M>::smem_bytes(), sdb_fused_table_doubles<D, M>(),           \
                    &sdb_fused_build_tables<D, M>, SDB_WS_THREADS, SDB_WS_PATHS, SDB_NOISE_KPW, SDB_SCHEME_SRK2});

SDB_FUSEDWS_AOT(10, 10)
