// Tuning variant 4 of the fused kernel (SDB_FUSED_VARIANT=4), for A/B measurements (profiles/).
#define SDB_FUSED_VARIANT 4
#define SDB_FUSED_VOL_MMA 0
#define SDB_FUSED_VOL_SMEM 0
#include "sdb_fused_inst.inc"
SDB_FUSED_AOT(10, 10)
