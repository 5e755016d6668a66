// Tuning variant 18 of the fused kernel (SDB_FUSED_VARIANT=18): the two warps of a sub-partition enter every
// phase together (SDB_FUSED_INPHASE, sdb_fused.cuh): a DMMA needs the FP64 pipe drained, so a warp in a
// tensor phase waits on the other warp's DFMA bursts when they drift apart.
#define SDB_FUSED_VARIANT 18
#define SDB_FUSED_INPHASE 1
#include "sdb_fused_inst.inc"
SDB_FUSED_AOT(10, 10)
