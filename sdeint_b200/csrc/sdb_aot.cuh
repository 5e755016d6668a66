// sdb_aot.cuh -- macros that instantiate and register the step-loop kernels of one system.
#pragma once
#include "sdb_host.h"
#include "sdb_kernels.cuh"

#define SDB_AOT_ONE(TAG, SCH, NOI, FK, GK, D, M, INL)                                         \
  SDB_LOOP_KERNEL(sdbk_##TAG##_s##SCH##_n##NOI, SCH, F_##TAG, G_##TAG, NOI)                   \
  static SdbAotRegistrar sdbreg_##TAG##_s##SCH##_n##NOI(                                      \
      sdb_loop_key(SCH, FK, GK, D, M, NOI, INL), (const void*)sdbk_##TAG##_s##SCH##_n##NOI);

// Euler/Heun need only dW (noise id KPW stands for "on-device Philox"); SRK2/KP2iS also need
// the repeated integrals, from the caller (HOST), Ikpw/Jkpw (KPW) or Iwik/Jwik (WIK).
#define SDB_AOT_SYSTEM(TAG, FK, GK, D, M, INL)   \
  SDB_AOT_ONE(TAG, 0, 0, FK, GK, D, M, INL)      \
  SDB_AOT_ONE(TAG, 0, 1, FK, GK, D, M, INL)      \
  SDB_AOT_ONE(TAG, 1, 0, FK, GK, D, M, INL)      \
  SDB_AOT_ONE(TAG, 1, 1, FK, GK, D, M, INL)      \
  SDB_AOT_ONE(TAG, 2, 0, FK, GK, D, M, INL)      \
  SDB_AOT_ONE(TAG, 2, 1, FK, GK, D, M, INL)      \
  SDB_AOT_ONE(TAG, 2, 2, FK, GK, D, M, INL)      \
  SDB_AOT_ONE(TAG, 3, 0, FK, GK, D, M, INL)      \
  SDB_AOT_ONE(TAG, 3, 1, FK, GK, D, M, INL)      \
  SDB_AOT_ONE(TAG, 3, 2, FK, GK, D, M, INL)
