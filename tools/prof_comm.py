"""Timing / ncu driver for the commutative-noise kernel (sdb_fusedc.cuh) at d = m = 10:
    python tools/prof_comm.py [paths] [steps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdeint_b200 as sdb

P = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 256 * 4
T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
rng = np.random.default_rng(3)
Cm = rng.standard_normal((10, 10)) / np.sqrt(10)
B = np.stack([(0.25 / np.sqrt(10)) * ((1.0 + 0.1 * k) * np.eye(10) + 0.5 * Cm + 0.1 * (k - 5) * Cm @ Cm)
              for k in range(10)])
A = np.full((10, 10), 0.05)
np.fill_diagonal(A, -1.5)
f, G, y0 = sdb.ops.LinearDrift(A), sdb.ops.LinearDiffusion(B), np.ones(10)
tspan = np.linspace(0.0, T / 1000.0, T + 1)
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    y = sdb.itoSRI2(f, G, y0, tspan, paths=P, seed=1, save_every=T, out="torch")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("%s: %d paths x %d steps in %.3f ms -> %.3e path-steps/s" % (sdb.last_kernel(), P, T, ms, P * T / ms * 1e3),
          flush=True)
