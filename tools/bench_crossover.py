"""Fused (DMMA) vs generic (thread-per-path) kernels on mid-size linear systems: the measured crossover
behind the dispatch thresholds of sdb_integrate.  python tools/bench_crossover.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, sdeint_b200 as sdb
for d, m in ((6, 6), (8, 6), (8, 8), (10, 10), (12, 8)):
    rng = np.random.default_rng(d * 10 + m)
    A = -np.eye(d) + 0.2 * rng.standard_normal((d, d)) / np.sqrt(d)
    B = 0.3 * rng.standard_normal((m, d, d)) / np.sqrt(d * m)
    f, G, y0 = sdb.ops.LinearDrift(A), sdb.ops.LinearDiffusion(B), np.ones(d)
    T, P = 1000, 1000000
    tspan = np.linspace(0.0, 1.0, T + 1)
    for sch, fn in (("stratHeun", sdb.stratHeun), ("itoEuler", sdb.itoEuler), ("itoSRI2", sdb.itoSRI2)):
        res = []
        for dis in ("", "1"):
            if dis: os.environ["SDB_DISABLE_FUSED"] = "1"
            else: os.environ.pop("SDB_DISABLE_FUSED", None)
            best = None
            for rep in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); e0.record()
                y = fn(f, G, y0, tspan, paths=P, seed=1, save_every=T, out="torch")
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1); best = ms if best is None else min(best, ms)
            res.append(P * T / best * 1e3)
        print("d=%d m=%d %-10s fused %.3e generic %.3e ratio %.2f" % (d, m, sch, res[0], res[1], res[0] / res[1]))
