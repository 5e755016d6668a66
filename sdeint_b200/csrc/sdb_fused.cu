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
