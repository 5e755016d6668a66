// Tuning variant 13 of the headline kernel (SDB_FUSED_VARIANT=13): the TMEM-resident pipeline of
// sdb_fusedx.cuh for d = m = 10 with Ikpw/Jkpw and 3 warps per SM sub-partition.
#define SDB_FUSED_VARIANT 13
#define SDB_FUSED_RB 2
#define SDB_FUSED_VOL_MMA 0
#define SDB_X_WPQ 3
#include "sdb_fusedx.cuh"
#include "sdb_host.h"

#include <cuda_runtime.h>

SDB_FUSEDX_KERNEL(sdbk_fusedx_t3_srk2_d10_m10_kpw, 10, 10, 0)
static SdbFusedRegistrar sdbreg_fusedx_t3(
    SdbFusedEntry{10, 10, SDB_FUSED_VARIANT, (const void*)sdbk_fusedx_t3_srk2_d10_m10_kpw,
                  SdbXShape<10, 10>::smem_bytes(), sdb_fused_table_doubles<10, 10, SDB_FUSED_A2>(),
                  &sdb_fused_build_tables<10, 10, SDB_FUSED_A2>, SdbXShape<10, 10>::THREADS, SdbXShape<10, 10>::PATHS,
                  SDB_NOISE_KPW, SDB_SCHEME_SRK2});
