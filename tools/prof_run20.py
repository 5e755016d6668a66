"""Timing / ncu driver for the d = m = 20 kernel: python tools/prof_run20.py [paths] [steps] [reps] [kpw|wik]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdeint_b200 as sdb

P = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128
T = int(sys.argv[2]) if len(sys.argv) > 2 else 50
R = int(sys.argv[3]) if len(sys.argv) > 3 else 2
meth = sdb.Jwik if (len(sys.argv) > 4 and sys.argv[4] == "wik") else sdb.Jkpw
f, G, y0 = sdb.ops.sys_lin(20, 20)
tspan = np.linspace(0.0, T / 1000.0, T + 1)
best = None
for rep in range(R):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    y = sdb.stratSRS2(f, G, y0, tspan, meth, paths=P, seed=20261019, save_every=T, out="torch")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    best = ms if best is None else min(best, ms)
print("best: %d paths x %d steps in %.3f ms -> %.4e path-steps/s" % (P, T, best, P * T / best * 1e3), flush=True)
