// Tuning variant 5 of the fused kernel (SDB_FUSED_VARIANT=5), for A/B measurements (profiles/).
#define SDB_FUSED_VARIANT 5
#define SDB_FUSED_BATCH_XZ 1
#include "sdb_fused_inst.inc"
SDB_FUSED_AOT(10, 10)
