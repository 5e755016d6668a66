"""Cost of the in-loop path functionals (observe=) on the headline shape, one GPU, device-timed.

    python tools/bench_observe.py [paths]

SYS_LIN(10,10) itoSRI2 + Ikpw(n=5), 1000 steps, save_every=100: the ahead-of-time kernel without
observers against the NVRTC-built observing variant with 1 and 4 observers.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdeint_b200 as sdb

P = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2 * 10 ** 6
f, G, y0 = sdb.ops.sys_lin(10, 10)
tspan = np.linspace(0.0, 1.0, 1001)
cases = (("none", None),
         ("1: first_below", [("first_below", 0, 0.5)]),
         ("4: first_below, first_above, max, integral",
          [("first_below", 0, 0.5), ("first_above", 3, 1.2), ("max", 5), ("integral", 9)]))


def timed(fn, reps=3):
    best = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        r = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
        del r
    return best


for name, obs in cases:
    run = lambda n=P: sdb.itoSRI2(f, G, y0, tspan, paths=n, seed=20261019, save_every=100, out="torch",
                                  observe=obs)
    run(10 ** 5)
    ms = timed(run)
    rec = {"observers": name, "paths": P, "steps": 1000, "ms": ms, "path_steps_per_s": P * 1000 / ms * 1e3}
    if obs is not None:
        _, o = run()
        t = o[0]
        rec["P(first passage below 0.5 within T=1)"] = float(torch.isfinite(t).double().mean())
        rec["mean first-passage time (of those)"] = float(t[torch.isfinite(t)].mean())
    print(json.dumps(rec), flush=True)
