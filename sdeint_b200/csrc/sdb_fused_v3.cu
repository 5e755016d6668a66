// Tuning variant 3 of the fused kernel (SDB_FUSED_VARIANT=3), for A/B measurements (profiles/).
#define SDB_FUSED_VARIANT 3
#define SDB_FUSED_VOL_MMA 0
#include "sdb_fused_inst.inc"
SDB_FUSED_AOT(10, 10)
