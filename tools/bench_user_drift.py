"""Nonlinear drift (source) + linear diffusion g_k = B_k y: the fused kernel with the drift evaluated per lane
(jit:fusedjit|ud) against the generic loop kernel.   python tools/bench_user_drift.py [d] [m] [paths] [steps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdeint_b200 as sdb

d = int(sys.argv[1]) if len(sys.argv) > 1 else 10
m = int(sys.argv[2]) if len(sys.argv) > 2 else 10
P = int(sys.argv[3]) if len(sys.argv) > 3 else 2000000
T = int(sys.argv[4]) if len(sys.argv) > 4 else 200
_, G, y0 = sdb.ops.sys_lin(d, m)
f = sdb.ops.ExprDrift(["-p[0]*y[%d] + p[1]*y[%d] - p[2]*y[%d]*y[%d]*y[%d]" % (i, (i + 1) % d, i, i, i)
                       for i in range(d)], params=[1.5, 0.3, 0.2])
tspan = np.linspace(0.0, T / 1000.0, T + 1)
for label, env in (("fused, drift per lane", None), ("generic loop kernel", "SDB_DISABLE_FUSED")):
    if env:
        os.environ[env] = "1"
    best = None
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        y = sdb.itoSRI2(f, G, y0, tspan, paths=P, seed=7, save_every=T, out="torch", commutative=False)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None or rep == 0 else min(best, ms) if rep > 0 else ms
    print("user drift d=%d m=%d itoSRI2+Ikpw %-22s %d paths x %d steps in %.1f ms -> %.3e path-steps/s  [%s]"
          % (d, m, label, P, T, best, P * T / best * 1e3, sdb.last_kernel()), flush=True)
    if env:
        del os.environ[env]
if d == 10 and m == 10:
    for label, fn, kw in (("itoSRI2+Iwik", sdb.itoSRI2, {"Imethod": sdb.Iwik}), ("stratHeun", sdb.stratHeun, {}),
                          ("itoEuler", sdb.itoEuler, {})):
        for env in (None, "SDB_DISABLE_FUSED"):
            if env:
                os.environ[env] = "1"
            best = None
            for rep in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                if "Imethod" in kw:
                    fn(f, G, y0, tspan, kw["Imethod"], paths=P, seed=7, save_every=T, out="torch", commutative=False)
                else:
                    fn(f, G, y0, tspan, paths=P, seed=7, save_every=T, out="torch")
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                best = ms if best is None else min(best, ms)
            print("user drift d=%d m=%d %-14s %.3e path-steps/s  [%s]" % (d, m, label, P * T / best * 1e3, sdb.last_kernel()),
                  flush=True)
            if env:
                del os.environ[env]
