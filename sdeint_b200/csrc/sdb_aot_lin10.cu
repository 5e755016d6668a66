// AOT kernels: SYS_LIN(10,10), the headline configuration (BASELINE.json configs[2], [4]).
#include "sdb_aot.cuh"

typedef DriftLinear<10, PInline<100>> F_lin10;
typedef DiffLinearCols<10, 10, PInline<1000>> G_lin10;
SDB_AOT_SYSTEM(lin10, 0, 1, 10, 10, true)
