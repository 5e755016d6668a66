// Tuning variant 2 of the fused kernel (SDB_FUSED_VARIANT=2): SDB_FUSED_BATCH SDB_FUSED_PARK_Y
#define SDB_FUSED_VARIANT 2
#define SDB_FUSED_BATCH 1
#define SDB_FUSED_PARK_Y 1
#include "sdb_fused_inst.inc"
SDB_FUSED_AOT(10, 10)
