"""A/B at d=m=20: the two-lanes-per-path TMEM kernel (sdb_fusedx2.cuh, default) against the one-lane
kernel (sdb_fusedx.cuh, SDB_FUSED_VARIANT=20): agreement and throughput.  python tools/check_x2.py [paths]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sdeint_b200 as sdb

f, G, y0 = sdb.ops.sys_lin(20, 20)
ts = np.linspace(0.0, 0.05, 51)


def rel(a, b):
    return float(np.max(np.linalg.norm(a - b, axis=-1) / np.linalg.norm(b, axis=-1)))


for meth, name in ((sdb.Jkpw, "Jkpw"), (sdb.Jwik, "Jwik")):
    os.environ.pop("SDB_FUSED_VARIANT", None)
    ref = sdb.stratSRS2(f, G, y0, ts, meth, paths=333, seed=5, path_id0=7, save_every=10)
    os.environ["SDB_FUSED_VARIANT"] = "20"
    new = sdb.stratSRS2(f, G, y0, ts, meth, paths=333, seed=5, path_id0=7, save_every=10)
    print(name, "rel err one-lane (variant 20) vs default:", rel(new, ref), "finite:", np.isfinite(new).all(), flush=True)
tspan = np.linspace(0.0, 1.0, 1001)
P = int(float(sys.argv[1])) if len(sys.argv) > 1 else 500000
for var in (None, "20"):
    if var is None:
        os.environ.pop("SDB_FUSED_VARIANT", None)
    else:
        os.environ["SDB_FUSED_VARIANT"] = var
    for meth, name in ((sdb.Jwik, "Jwik"), (sdb.Jkpw, "Jkpw")):
        best = None
        for _ in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            y = sdb.stratSRS2(f, G, y0, tspan, meth, paths=P, seed=20261019, save_every=100, out="torch")
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1); best = ms if best is None else min(best, ms)
        print("variant", var, name, "%.1f ms -> %.3e path-steps/s" % (best, P * 1000 / best * 1e3), flush=True)
