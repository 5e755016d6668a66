"""Small driver for timing / ncu: launches of the headline kernel (SYS_LIN(10,10) itoSRI2 + Ikpw).
    python tools/prof_run.py [paths] [time_steps] [reps]      (SDB_DISABLE_FUSED=1 -> generic loop)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdeint_b200 as sdb

P = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 256 * 2
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
R = int(sys.argv[3]) if len(sys.argv) > 3 else 2
SE = int(os.environ.get("SDB_PROF_SAVE_EVERY", T))
f, G, y0 = sdb.ops.sys_lin(10, 10)
tspan = np.linspace(0.0, T / 1000.0, T + 1)
best = None
for rep in range(R):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    y = sdb.itoSRI2(f, G, y0, tspan, paths=P, seed=20261019, save_every=SE, out="torch")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    best = ms if best is None else min(best, ms)
    print("rep %d: %d paths x %d steps in %.3f ms (device) -> %.3e path-steps/s; mean y[-1]=%s"
          % (rep, P, T, ms, P * T / ms * 1e3, y[-1].mean(dim=1)[:2].tolist()), flush=True)
print("best: %.3f ms -> %.4e path-steps/s" % (best, P * T / best * 1e3), flush=True)
