// Tuning variant 2 of the fused kernel (SDB_FUSED_VARIANT=2): Philox4x32-10 for EVERY slot (the round-1
// generator), kept for A/B measurements against the keyed MX2 expansion.
#define SDB_FUSED_VARIANT 2
#define SDB_FUSED_RNG 0
#include "sdb_fused_inst.inc"
SDB_FUSED_AOT(10, 10)
