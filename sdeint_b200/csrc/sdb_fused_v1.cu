// Tuning variant 1 of the fused kernel (selected with SDB_FUSED_VARIANT=1): DMMA asm not volatile.
#define SDB_FUSED_VARIANT 1
#define SDB_FUSED_VOL_MMA 0
#include "sdb_fused_inst.inc"
SDB_FUSED_AOT(10, 10)
