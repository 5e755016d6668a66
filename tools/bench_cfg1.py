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
# the same call with the README's own plain Python functions (traced into device code)
a, b = 1.0, 0.8
def f_py(x, t): return -(a + x*b**2)*(1 - x**2)
def g_py(x, t): return b*(1 - x**2)
sdb.itoint(f_py, g_py, 0.1, tspan, np.random.default_rng(1))
t0 = time.perf_counter()
for _ in range(5):
    y2 = sdb.itoint(f_py, g_py, 0.1, tspan, np.random.default_rng(1))
print("cfg1 with plain Python f, g (traced):     %.4f s per call, y[-1]=%g  [%s]"
      % ((time.perf_counter() - t0) / 5, np.asarray(y2).reshape(-1)[-1], sdb.last_kernel()))
