// AOT instantiations of the fused linear SRI2/SRS2 kernel (sdb_fused.cuh): default variant.
#define SDB_FUSED_VARIANT 0
#define SDB_FUSED_KUNROLL 1
#include "sdb_fused_inst.inc"

std::vector<SdbFusedEntry>& sdb_fused_registry() {
  static std::vector<SdbFusedEntry> reg;
  return reg;
}

SDB_FUSED_AOT(10, 10)   // BASELINE.json configs[2] and [4]
SDB_FUSED_AOT(4, 4)     // small case for tests: no DMMA tile of the stage-2 sum (IT = 0)
SDB_FUSED_AOT(8, 2)     // exactly one DMMA tile, no leftover rows
SDB_FUSED_AOT(5, 4)     // Roessler (7.4) dimensions

// The run-time mirror used for NVRTC-compiled shapes (sdb_fused_rt.h) must agree with the templates.
#include "sdb_fused_rt.h"
#define SDB_CHECK_DIMS(D, M)                                                                   \
  static_assert(SdbFusedDims(D, M).smem_bytes() == SdbFusedShape<D, M>::smem_bytes() &&        \
                    SdbFusedDims(D, M).n_frags() == SdbFusedShape<D, M>::n_frags() &&          \
                    SdbFusedDims(D, M).n_left() == SdbFusedShape<D, M>::n_left() &&            \
                    SdbFusedDims(D, M).grows() == SdbFusedShape<D, M>::grows() &&              \
                    SdbFusedDims(D, M).p3_off(1) == SdbFusedShape<D, M>::p3_off(1) &&          \
                    SdbFusedDims(D, M, 4, 1).smem_bytes() == SdbFusedShape<D, M, 1>::smem_bytes() &&  \
                    SdbFusedDims(D, M, 4, 1).n_frags() == SdbFusedShape<D, M, 1>::n_frags() &&  \
                    SdbFusedDims(D, M, 4, 1).tab_copies() == SdbFusedShape<D, M, 1>::tab_copies() && \
                    SdbFusedDims(D, M, 4, 1).p3_off(1) == SdbFusedShape<D, M, 1>::p3_off(1) &&  \
                    SdbFusedDims(D, M, 4, -1).smem_bytes() == SdbFusedShape<D, M, -1>::smem_bytes() && \
                    SdbFusedDims(D, M, 4, -1).n_frags() == SdbFusedShape<D, M, -1>::n_frags() && \
                    SdbFusedDims(D, M, 4, -1).p3_off(1) == SdbFusedShape<D, M, -1>::p3_off(1), \
                "sdb_fused_rt.h disagrees with SdbFusedShape");
SDB_CHECK_DIMS(10, 10) SDB_CHECK_DIMS(4, 4) SDB_CHECK_DIMS(8, 2) SDB_CHECK_DIMS(5, 4)
SDB_CHECK_DIMS(3, 2) SDB_CHECK_DIMS(12, 6) SDB_CHECK_DIMS(15, 8) SDB_CHECK_DIMS(7, 10)
static_assert(SDB_FUSED_THREADS == SdbFusedDims::THREADS && SDB_FUSED_MAX_TERMS == SdbFusedDims::MAX_TERMS,
              "sdb_fused_rt.h constants");
