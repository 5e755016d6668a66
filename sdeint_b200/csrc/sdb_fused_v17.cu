// Tuning variant 17 of the fused kernel (SDB_FUSED_VARIANT=17): the two warps of a sub-partition run half a
// step apart (SDB_FUSED_ANTIPHASE, sdb_fused.cuh).
#define SDB_FUSED_VARIANT 17
#define SDB_FUSED_ANTIPHASE 1
#include "sdb_fused_inst.inc"
SDB_FUSED_AOT(10, 10)
