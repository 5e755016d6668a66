"""d = m = 20: nonlinear drift (source) + linear diffusion, stratSRS2 + Jwik: two lanes per path (x2) against one lane
per path (SDB_DISABLE_FUSEDX2=1) and the generic loop kernel.   python tools/bench_user_drift20.py [paths] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, sdeint_b200 as sdb
P = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
d = m = 20
_, G, y0 = sdb.ops.sys_lin(d, m)
f = sdb.ops.ExprDrift(["-p[0]*y[%d] + p[1]*y[%d] - p[2]*y[%d]*y[%d]*y[%d]" % (i, (i + 1) % d, i, i, i)
                       for i in range(d)], params=[1.5, 0.3, 0.2])
tspan = np.linspace(0.0, T / 1000.0, T + 1)
for label, env, PP in (("two lanes per path", None, P), ("one lane per path", "SDB_DISABLE_FUSEDX2", P),
                       ("generic loop kernel", "SDB_DISABLE_FUSED", P // 5)):
    if env:
        os.environ[env] = "1"
    best = None
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        sdb.stratSRS2(f, G, y0, tspan, sdb.Jwik, paths=PP, seed=7, save_every=T, out="torch", commutative=False)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    print("user drift d=m=20 stratSRS2+Jwik %-20s %.3e path-steps/s  [%s]" % (label, PP * T / best * 1e3, sdb.last_kernel()),
          flush=True)
    if env:
        del os.environ[env]
