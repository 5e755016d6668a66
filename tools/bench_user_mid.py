"""Throughput of the generic (thread-per-path, NVRTC-built) kernels on a mid-size NONLINEAR system
given as formulas: d = 7, m = 5, non-commutative, time-dependent (the system of
tests/test_gpu_parity.py::test_mid_size_nonlinear_user_system).
    python tools/bench_user_mid.py [paths] [steps] [d] [m]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdeint_b200 as sdb

P = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 500
d = int(sys.argv[3]) if len(sys.argv) > 3 else 7
m = int(sys.argv[4]) if len(sys.argv) > 4 else 5
drift = ["-1.2*y[%d] + 0.3*sin(y[%d]) + 0.1*y[%d]*y[%d]/(1.0+y[%d]*y[%d]) + 0.05*cos(t)"
         % (i, (i + 1) % d, (i + 2) % d, (i + 3) % d, (i + 2) % d, (i + 2) % d) for i in range(d)]
rows = [["%g*(%s)" % (0.08 + 0.01 * ((3 * i + k) % 5),
                      ["y[%d]" % ((i + k) % d), "sin(y[%d])+0.5" % ((i + 2 * k) % d),
                       "1.0+0.2*y[%d]*y[%d]" % (i, (k + 1) % d), "cos(y[%d]+t)" % ((i + k + 3) % d),
                       "0.3+tanh(y[%d])" % ((2 * i + k) % d)][(i + k) % 5])
         for k in range(m)] for i in range(d)]
f, G = sdb.ops.ExprDrift(drift, []), sdb.ops.ExprDiffusion(rows, [])
y0 = 0.5 + 0.1 * np.arange(d)
tspan = np.linspace(0.0, T * 1e-3, T + 1)
for name, fn, kw in (("itoSRI2+Ikpw", sdb.itoSRI2, {}), ("itoSRI2+Iwik", sdb.itoSRI2, {"Imethod": sdb.Iwik}),
                     ("stratHeun", sdb.stratHeun, {}), ("itoEuler", sdb.itoEuler, {})):
    best = None
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        y = fn(f, G, y0, tspan, paths=P, seed=7, save_every=T, out="torch", **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    print("user d=%d m=%d %-13s %d paths x %d steps in %.1f ms -> %.3e path-steps/s  [%s]"
          % (d, m, name, P, T, best, P * T / best * 1e3, sdb.last_kernel()), flush=True)
