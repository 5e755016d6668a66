"""Throughput of the non-headline BASELINE.json configs on one GPU (device-timed, CUDA events).

    python tools/bench_configs.py [cfg2|cfg4|all] [scale]

cfg2: README 2-d linear SDE, itoEuler and stratHeun, 10^6 paths x 10^4 steps, save_every=10.
cfg4: SYS_LIN(20,20), stratSRS2 + Jwik(n=5), 1000 steps, 10^6 paths (scale < 1 shrinks paths).
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdeint_b200 as sdb

which = sys.argv[1] if len(sys.argv) > 1 else "all"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0


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


out = []
# warm-up: clocks, module load, NVRTC caches
_f, _G, _y0 = sdb.ops.sys_lin2()
for _ in range(3):
    sdb.itoEuler(_f, _G, _y0, np.linspace(0.0, 1.0, 2001), paths=10 ** 6, seed=1, save_every=2000, out="torch")
torch.cuda.synchronize()
if which in ("cfg2", "all"):
    f, G, y0 = sdb.ops.sys_lin2()
    P, T, se = int(1e6 * scale), 10000, 10
    tspan = np.linspace(0.0, 10.0, T + 1)
    for name, fn, flops in (("itoEuler", sdb.itoEuler, 18), ("stratHeun", sdb.stratHeun, 40)):
        ms = timed(lambda: fn(f, G, y0, tspan, paths=P, seed=20261019, save_every=se, out="torch"))
        ps = P * T / ms * 1e3
        out.append({"config": "cfg2", "scheme": name, "paths": P, "steps": T, "save_every": se,
                    "ms": ms, "path_steps_per_s": ps, "store_GBps": ps * 16.0 / se / 1e9,
                    "counted_TFLOPs": ps * flops / 1e12})
    # every step stored (save_every=1): the coalesced trajectory store at full rate
    P1, T1 = int(1e6 * scale), 1000
    ts1 = np.linspace(0.0, 1.0, T1 + 1)
    ms = timed(lambda: sdb.itoEuler(f, G, y0, ts1, paths=P1, seed=20261019, save_every=1, out="torch"))
    ps = P1 * T1 / ms * 1e3
    out.append({"config": "cfg2-dense", "scheme": "itoEuler", "paths": P1, "steps": T1, "save_every": 1,
                "ms": ms, "path_steps_per_s": ps, "store_GBps": ps * 16.0 / 1e9})
if which in ("cfg4", "all"):
    f, G, y0 = sdb.ops.sys_lin(20, 20)
    P, T, se = int(1e6 * scale), 1000, 100
    tspan = np.linspace(0.0, 1.0, T + 1)
    sdb.stratSRS2(f, G, y0, np.linspace(0.0, 0.02, 21), sdb.Jwik, paths=10 ** 5, seed=1, save_every=20, out="torch")
    sdb.stratSRS2(f, G, y0, np.linspace(0.0, 0.02, 21), sdb.Jkpw, paths=10 ** 5, seed=1, save_every=20, out="torch")
    for meth, nm, fl in ((sdb.Jwik, "Jwik", 67280 + 8983), (sdb.Jkpw, "Jkpw", 67280 + 5880)):
        ms = timed(lambda: sdb.stratSRS2(f, G, y0, tspan, meth, paths=P, seed=20261019,
                                         save_every=se, out="torch"), reps=1)
        ps = P * T / ms * 1e3
        out.append({"config": "cfg4", "scheme": "stratSRS2+" + nm, "paths": P, "steps": T,
                    "save_every": se, "ms": ms, "path_steps_per_s": ps,
                    "counted_TFLOPs": ps * fl / 1e12})
for o in out:
    print(json.dumps(o), flush=True)
