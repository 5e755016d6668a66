// Tuning variant 6 of the fused kernel (SDB_FUSED_VARIANT=6), for A/B measurements (profiles/).
#define SDB_FUSED_VARIANT 6
#define SDB_FUSED_KUNROLL 5
#include "sdb_fused_inst.inc"
SDB_FUSED_AOT(10, 10)
