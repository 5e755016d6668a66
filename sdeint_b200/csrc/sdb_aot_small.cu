// AOT kernels: the scalar / small test systems of sdeint/tests/test_integrate.py and README.rst.
#include "sdb_aot.cuh"

// README 2-d linear SDE, A y dt + B dW (README.rst:83-95): BASELINE.json configs[1]
typedef DriftLinear<2, PInline<4>> F_lin2;
typedef DiffConst<2, 2, PInline<4>> G_lin2;
SDB_AOT_SYSTEM(lin2, 0, 0, 2, 2, true)

// 1-d Ornstein-Uhlenbeck (test_integrate.py:37-54)
typedef DriftLinear<1, PInline<1>> F_ou;
typedef DiffConst<1, 1, PInline<1>> G_ou;
SDB_AOT_SYSTEM(ou, 0, 0, 1, 1, true)

// KP (4.46) = README scalar SDE (README.rst:60-69): configs[0]
typedef DriftKP4446<PInline<2>> F_kp4446;
typedef DiffKP4446<PInline<1>> G_kp4446;
SDB_AOT_SYSTEM(kp4446, 1, 2, 1, 1, true)

// KPS (4.5)
typedef DriftKPS445<PInline<1>> F_kps445;
typedef DiffKPS445<PInline<1>> G_kps445;
SDB_AOT_SYSTEM(kps445, 2, 3, 1, 1, true)

// KP (4.59): d=1, m=2 linear multiplicative noise
typedef DriftLinear<1, PInline<1>> F_kp4459;
typedef DiffLinearCols<1, 2, PInline<2>> G_kp4459;
SDB_AOT_SYSTEM(kp4459, 0, 1, 1, 2, true)

// Roessler (7.4): d=5, m=4 linear
typedef DriftLinear<5, PInline<25>> F_r74;
typedef DiffLinearCols<5, 4, PInline<100>> G_r74;
SDB_AOT_SYSTEM(r74, 0, 1, 5, 4, true)
