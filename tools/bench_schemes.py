"""Throughput of stratKP2iS / stratSRS2 / stratHeun / itoEuler on two linear systems: python tools/bench_schemes.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, sdeint_b200 as sdb
for name, mk in (("r74 d=5 m=4", lambda: sdb.ops.sys_r74(True)), ("lin10 d=m=10", lambda: sdb.ops.sys_lin(10, 10))):
    f, G, y0 = mk()
    T, P = 200, 2000000
    tspan = np.linspace(0.0, T / 1000.0, T + 1)
    for sch, fn in (("stratKP2iS", sdb.stratKP2iS), ("stratSRS2", sdb.stratSRS2), ("stratHeun", sdb.stratHeun),
                    ("itoEuler", sdb.itoEuler)):
        best = None
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            y = fn(f, G, y0, tspan, paths=P, seed=1, save_every=T, out="torch")
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1); best = ms if best is None else min(best, ms)
        print(name, sch, "%.3e path-steps/s" % (P * T / best * 1e3))
