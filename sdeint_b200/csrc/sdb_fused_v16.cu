// Tuning variant 16 of the fused kernel (SDB_FUSED_VARIANT=16): the random bits of step n + 1 are produced
// inside the row blocks of step n and wait in tensor memory (SDB_FUSED_AHEAD, sdb_fused.cuh).
#define SDB_FUSED_VARIANT 16
#define SDB_FUSED_AHEAD 1
#include "sdb_fused_inst.inc"
SDB_FUSED_AOT(10, 10)
