// Tuning variant 8 of the fused kernel (SDB_FUSED_VARIANT=8): THREE warps per SM sub-partition (384
// threads, <= 168 registers) -- 2-row blocks, the state parked in the staging buffer during the noise
// phase, dW in registers through the row blocks (its shared-memory rows would not fit next to the
// Levy areas of 384 paths).
#define SDB_FUSED_VARIANT 8
#define SDB_FUSED_THREADS 384
#define SDB_FUSED_RB 2
#define SDB_FUSED_PARK_Y 1
#define SDB_FUSED_DW_REGS 1
#include "sdb_fused_inst.inc"
SDB_FUSED_AOT(10, 10)
