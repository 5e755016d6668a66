// Tuning variant 1 of the fused kernel (selected with SDB_FUSED_VARIANT=1): the pair-at-a-time
// noise generation (no lock-step batch), kept for A/B measurements (profiles/).
#define SDB_FUSED_VARIANT 1
#define SDB_FUSED_BATCH 0
#include "sdb_fused_inst.inc"
SDB_FUSED_AOT(10, 10)
