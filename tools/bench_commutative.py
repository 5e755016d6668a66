"""Commutative-noise fast path (no Levy areas) against the full Ikpw construction on one GPU.

    python tools/bench_commutative.py [paths]

f = A y, g_k = B_k y with commuting B_k (polynomials in one matrix): itoSRI2 over 1000 steps,
save_every=100, device-timed; `commutative=False` forces the full construction on the same system.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdeint_b200 as sdb

P = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2 * 10 ** 6


def commuting(d, m, seed=3):
    rng = np.random.default_rng(seed)
    Cm = rng.standard_normal((d, d)) / np.sqrt(d)
    B = np.stack([(0.25 / np.sqrt(m)) * ((1.0 + 0.1 * k) * np.eye(d) + 0.5 * Cm + 0.1 * (k - m / 2) * Cm @ Cm)
                  for k in range(m)])
    A = np.full((d, d), 0.5 / d)
    np.fill_diagonal(A, -1.5)
    return sdb.ops.LinearDrift(A), sdb.ops.LinearDiffusion(B), np.ones(d)


def timed(fn, reps=3):
    best = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        y = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
        del y
    return best


tspan = np.linspace(0.0, 1.0, 1001)
for d, m, scale in ((10, 10, 1.0), (5, 4, 1.0), (20, 20, 0.25)):
    f, G, y0 = commuting(d, m)
    assert G.commutative
    n = int(P * scale)
    for comm in (None, False):
        sdb.itoSRI2(f, G, y0, tspan[:21], paths=10 ** 5, seed=1, save_every=20, out="torch", commutative=comm)
        ms = timed(lambda: sdb.itoSRI2(f, G, y0, tspan, paths=n, seed=20261019, save_every=100,
                                       out="torch", commutative=comm))
        print(json.dumps({"d": d, "m": m, "paths": n, "steps": 1000,
                          "noise": "commutative (dW only)" if comm is None else "full Ikpw n=5",
                          "ms": ms, "path_steps_per_s": n * 1000 / ms * 1e3}), flush=True)
