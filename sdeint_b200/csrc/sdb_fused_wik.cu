// AOT instantiation of the register-resident fused SRI2/SRS2 kernel (sdb_fused.cuh) with Wiktorsson's
// tail: d = m = 10 with Iwik/Jwik.  Registered as variant 30 first (A/B against the TMEM-resident
// sdb_fusedx.cuh instance), see sdb_capi.cu for the default.
#define SDB_FUSED_VARIANT 30
#define SDB_FUSED_KUNROLL 1
#include "sdb_fused.cuh"
#include "sdb_host.h"

#include <cuda_runtime.h>

#define SDB_FUSED_WIK_AOT(D, M, VARIANT)                                                        \
  SDB_FUSED_KERNEL_W(sdbk_fused_srk2_d##D##_m##M##_wik, D, M, 1)                                \
  static SdbFusedRegistrar sdbreg_fused_wik_d##D##_m##M(                                        \
      SdbFusedEntry{D, M, VARIANT, (const void*)sdbk_fused_srk2_d##D##_m##M##_wik,              \
                    SdbFusedShape<D, M, SDB_FUSED_A2>::smem_bytes(),                            \
                    sdb_fused_table_doubles<D, M, SDB_FUSED_A2>(),                              \
                    &sdb_fused_build_tables<D, M, SDB_FUSED_A2>, SDB_FUSED_THREADS, SDB_FUSED_THREADS, \
                    SDB_NOISE_WIK, SDB_SCHEME_SRK2});

SDB_FUSED_WIK_AOT(10, 10, 30)
