// AOT instantiations of the fused commutative-noise SRI2/SRS2 kernel for linear systems (sdb_fusedc.cuh).
#define SDB_FUSED_VARIANT 50
#include "sdb_fusedc.cuh"
#include "sdb_fused_rt.h"
#include "sdb_host.h"

#include <cuda_runtime.h>

#define SDB_FUSEDC_AOT(D, M)                                                                   \
  SDB_FUSEDC_KERNEL(sdbk_fusedc_srk2_d##D##_m##M, D, M)                                        \
  static SdbFusedRegistrar sdbreg_fusedc_d##D##_m##M(                                          \
      SdbFusedEntry{D, M, 0, (const void*)sdbk_fusedc_srk2_d##D##_m##M,                        \
                    SdbFcShape<D, M>::smem_bytes(), SdbFcShape<D, M>::table_doubles(),         \
                    &sdb_fusedc_build_tables<D, M>, SDB_FC_THREADS, SDB_FC_THREADS,            \
                    SDB_NOISE_COMM, SDB_SCHEME_SRK2});                                         \
  static_assert(SdbFusedCDims(D, M).smem_bytes() == SdbFcShape<D, M>::smem_bytes() &&          \
                    SdbFusedCDims(D, M).table_doubles() == SdbFcShape<D, M>::table_doubles() && \
                    SdbFusedCDims(D, M).table1_doubles() == SdbFcShape<D, M>::table1_doubles(), \
                "sdb_fused_rt.h disagrees with SdbFcShape");

SDB_FUSEDC_AOT(10, 10)
static_assert(SdbFusedCDims(5, 4).smem_bytes() == SdbFcShape<5, 4>::smem_bytes() &&
                  SdbFusedCDims(5, 4).table_doubles() == SdbFcShape<5, 4>::table_doubles() &&
                  SdbFusedCDims(12, 6).smem_bytes() == SdbFcShape<12, 6>::smem_bytes() &&
                  SdbFusedCDims(3, 2).table1_doubles() == SdbFcShape<3, 2>::table1_doubles(),
              "sdb_fused_rt.h disagrees with SdbFcShape");
