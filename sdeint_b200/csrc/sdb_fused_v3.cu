// Tuning variant 3 of the fused kernel (SDB_FUSED_VARIANT=3): SDB_FUSED_BATCH SDB_FUSED_PARK_Y SDB_FUSED_PARK_DW
#define SDB_FUSED_VARIANT 3
#define SDB_FUSED_BATCH 1
#define SDB_FUSED_PARK_Y 1
#define SDB_FUSED_PARK_DW 1
#include "sdb_fused_inst.inc"
SDB_FUSED_AOT(10, 10)
