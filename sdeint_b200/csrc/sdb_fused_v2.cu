// Tuning variant 2 of the fused kernel (selected with SDB_FUSED_VARIANT=2): Levy loop unrolled x5.
#define SDB_FUSED_VARIANT 2
#define SDB_FUSED_KUNROLL 5
#include "sdb_fused_inst.inc"
SDB_FUSED_AOT(10, 10)
