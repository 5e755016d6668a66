"""Timing of itoSRI2 d=m=10 with Iwik (TMEM-resident fused kernel) next to Ikpw: python tools/prof_wik10.py"""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, sdeint_b200 as sdb
f, G, y0 = sdb.ops.sys_lin(10, 10)
T, P = 200, 2000000
tspan = np.linspace(0.0, T / 1000.0, T + 1)
for meth, nm in ((sdb.Iwik, "Iwik"), (sdb.Ikpw, "Ikpw")):
    best = None
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        y = sdb.itoSRI2(f, G, y0, tspan, meth, paths=P, seed=1, save_every=T, out="torch")
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1); best = ms if best is None else min(best, ms)
    print(nm, "%.3e path-steps/s" % (P * T / best * 1e3))
