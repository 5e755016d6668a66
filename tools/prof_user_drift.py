"""ncu driver: one launch of the fused kernel with a user drift (jit:fusedjit|ud), d = m = 10.
    python tools/prof_user_drift.py [paths] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, sdeint_b200 as sdb
P = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
_, G, y0 = sdb.ops.sys_lin(10, 10)
f = sdb.ops.ExprDrift(["-p[0]*y[%d] + p[1]*y[%d] - p[2]*y[%d]*y[%d]*y[%d]" % (i, (i + 1) % 10, i, i, i)
                       for i in range(10)], params=[1.5, 0.3, 0.2])
tspan = np.linspace(0.0, T / 1000.0, T + 1)
y = sdb.itoSRI2(f, G, y0, tspan, paths=P, seed=7, save_every=T, out="torch", commutative=False)
torch.cuda.synchronize()
print(sdb.last_kernel(), y[-1].mean(dim=1)[:2].tolist())
