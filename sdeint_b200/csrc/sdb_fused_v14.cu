// Tuning variant 14 of the fused kernel (SDB_FUSED_VARIANT=14): the state components beyond the last full
// DMMA k-tile (d = 10: components 8, 9) enter the P1 product through DFMAs on the accumulator fragments
// instead of a third, half-empty k-tile.
#define SDB_FUSED_VARIANT 14
#define SDB_FUSED_KSPLIT 1
#include "sdb_fused_inst.inc"
SDB_FUSED_AOT(10, 10)
SDB_FUSED_AOT(5, 4)
