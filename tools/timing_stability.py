"""Launch-to-launch timing stability of one generic (NVRTC) kernel: 24 back-to-back itoEuler calls.
    python tools/timing_stability.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdeint_b200 as sdb

d = m = 6
rng = np.random.default_rng(66)
A = -np.eye(d) + 0.2 * rng.standard_normal((d, d)) / np.sqrt(d)
B = 0.3 * rng.standard_normal((m, d, d)) / np.sqrt(d * m)
f, G, y0 = sdb.ops.LinearDrift(A), sdb.ops.LinearDiffusion(B), np.ones(d)
T, P = 1000, 1000000
tspan = np.linspace(0.0, 1.0, T + 1)
ts = []
for rep in range(24):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    y = sdb.itoEuler(f, G, y0, tspan, paths=P, seed=1, save_every=T, out="torch")
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(" ".join("%.1f" % t for t in ts))
print("median %.2f ms, max %.2f ms (after the first call: %.2f)" % (np.median(ts), max(ts), max(ts[1:])))
