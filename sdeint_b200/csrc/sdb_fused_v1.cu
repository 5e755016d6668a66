// Tuning variant 1 of the fused kernel (SDB_FUSED_VARIANT=1): SDB_FUSED_BATCH
#define SDB_FUSED_VARIANT 1
#define SDB_FUSED_BATCH 1
#include "sdb_fused_inst.inc"
SDB_FUSED_AOT(10, 10)
