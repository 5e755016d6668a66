// Tuning variant 1 of the fused kernel (selected with SDB_FUSED_VARIANT=1): Levy loop unrolled x2.
#define SDB_FUSED_VARIANT 1
#define SDB_FUSED_KUNROLL 2
#include "sdb_fused_inst.inc"
SDB_FUSED_AOT(10, 10)
