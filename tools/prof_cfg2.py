"""Small driver for timing / ncu of the cfg2 kernels (README 2-d linear SDE, additive noise).
    python tools/prof_cfg2.py [euler|heun] [paths] [time_steps]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdeint_b200 as sdb

which = sys.argv[1] if len(sys.argv) > 1 else "euler"
P = int(sys.argv[2]) if len(sys.argv) > 2 else 10 ** 6
T = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
f, G, y0 = sdb.ops.sys_lin2()
tspan = np.linspace(0.0, T / 1000.0, T + 1)
fn = sdb.itoEuler if which == "euler" else sdb.stratHeun
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    y = fn(f, G, y0, tspan, paths=P, seed=20261019, save_every=10, out="torch")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("rep %d: %s %d paths x %d steps in %.3f ms -> %.3e path-steps/s" % (rep, which, P, T, ms, P * T / ms * 1e3),
          flush=True)
