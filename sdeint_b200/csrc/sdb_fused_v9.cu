// Diagnostic variant 9 of the fused kernel (SDB_FUSED_VARIANT=9): ONE warp per SM sub-partition (128 threads,
// one block per SM): how long a warp-step takes when nothing competes with it.
#define SDB_FUSED_VARIANT 9
#define SDB_FUSED_THREADS 128
#define SDB_FUSED_MIN_SMEM (120 * 1024)
#include "sdb_fused_inst.inc"
SDB_FUSED_AOT(10, 10)
