// AOT kernels: SYS_LIN(20,20) (BASELINE.json configs[3]); A and B (67 KB) exceed the 32 KB
// kernel-parameter space, so they are read from global memory.
#include "sdb_aot.cuh"

typedef DriftLinear<20, PGlobal> F_lin20;
typedef DiffLinearCols<20, 20, PGlobal> G_lin20;
SDB_AOT_ONE(lin20, 2, 0, 0, 1, 20, 20, false)
SDB_AOT_ONE(lin20, 2, 1, 0, 1, 20, 20, false)
SDB_AOT_ONE(lin20, 2, 2, 0, 1, 20, 20, false)
