// Tuning variant 2 of the fused kernel (selected with SDB_FUSED_VARIANT=2): DMMA asm and the
// shared-memory reads of P2 not volatile (ordered by __syncwarp() only).
#define SDB_FUSED_VARIANT 2
#define SDB_FUSED_VOL_MMA 0
#define SDB_FUSED_VOL_SMEM 0
#include "sdb_fused_inst.inc"
SDB_FUSED_AOT(10, 10)
