"""cfg1 (BASELINE.json configs[0]): the README scalar Ito SDE through itoint, one path, 5001 points."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdeint_b200 as sdb
f, G, y0 = sdb.ops.sys_kp4446(False)
tspan = np.linspace(0.0, 5.0, 5001)
sdb.itoint(f, G, y0, tspan, np.random.default_rng(1))
t0 = time.perf_counter()
for _ in range(5):
    y = sdb.itoint(f, G, y0, tspan, np.random.default_rng(1))
print("cfg1 itoint 5001 points, single path: %.4f s per call, y[-1]=%g" % ((time.perf_counter() - t0) / 5, y[-1, 0]))
