"""Small driver for ncu: one launch of the headline kernel (SYS_LIN(10,10) itoSRI2 + Ikpw).
    python tools/prof_run.py [paths] [time_steps]        (SDB_DISABLE_FUSED=1 -> generic loop)
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdeint_b200 as sdb

P = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 256 * 2
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
f, G, y0 = sdb.ops.sys_lin(10, 10)
tspan = np.linspace(0.0, T / 1000.0, T + 1)
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    y = sdb.itoSRI2(f, G, y0, tspan, paths=P, seed=20261019, save_every=T, out="torch")
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("rep %d: %d paths x %d steps in %.3f ms -> %.3e path-steps/s; mean y[-1]=%s"
          % (rep, P, T, dt * 1e3, P * T / dt, y[-1].mean(dim=1)[:3].tolist()))
