// Tuning variant 7 of the fused kernel (SDB_FUSED_VARIANT=7), for A/B measurements (profiles/).
#define SDB_FUSED_VARIANT 7
#define SDB_FUSED_SWP 1
#include "sdb_fused_inst.inc"
SDB_FUSED_AOT(10, 10)
