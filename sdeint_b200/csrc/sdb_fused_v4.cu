// Tuning variant 4 of the fused kernel (SDB_FUSED_VARIANT=4): SDB_FUSED_PARK_Y
#define SDB_FUSED_VARIANT 4
#define SDB_FUSED_PARK_Y 1
#include "sdb_fused_inst.inc"
SDB_FUSED_AOT(10, 10)
