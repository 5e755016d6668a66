// AOT instantiations of the TMEM-resident standalone Ikpw/Jkpw/Iwik/Jwik kernel (sdb_repintx.cuh).
#define SDB_FUSED_VARIANT 30
#define SDB_FUSED_RB 2
#define SDB_FUSED_VOL_MMA 0
#include "sdb_repintx.cuh"
#include "sdb_host.h"

#include <cuda_runtime.h>

#define SDB_REPINTX_AOT(M, WIK, TAG)                                                           \
  SDB_REPINTX_KERNEL(sdbk_repintx_m##M##_##TAG, M, WIK)                                        \
  static SdbRepintXRegistrar sdbreg_repintx_m##M##_##TAG(SdbRepintXEntry{                      \
      M, WIK ? SDB_NOISE_WIK : SDB_NOISE_KPW, (const void*)sdbk_repintx_m##M##_##TAG,          \
      SdbRxShape<M>::smem_bytes(), SdbRxShape<M>::THREADS});

std::vector<SdbRepintXEntry>& sdb_repintx_registry() {
  static std::vector<SdbRepintXEntry> reg;
  return reg;
}

SDB_REPINTX_AOT(20, 0, kpw)
SDB_REPINTX_AOT(20, 1, wik)
