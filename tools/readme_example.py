"""The README example, runnable: python tools/readme_example.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, sdeint_b200 as sdb
a, b = 1.0, 0.8
tspan = np.linspace(0.0, 5.0, 5001)
def f(x, t): return -(a + x*b**2)*(1 - x**2)
def g(x, t): return b*(1 - x**2)
result = sdb.itoint(f, g, 0.1, tspan); print(np.asarray(result).shape, sdb.last_kernel())
paths = sdb.itoint(f, g, 0.1, tspan, paths=10**5, seed=0, save_every=100); print(paths.shape)
f, G, y0 = sdb.ops.sys_lin(10, 10)
tspan = np.linspace(0.0, 1.0, 1001)
y = sdb.itoSRI2(f, G, y0, tspan, paths=10**5, seed=1, save_every=100); print(y.shape)
y1 = sdb.itoint(f, G, y0, tspan); print(y1.shape)
f = sdb.ops.ExprDrift(["p[0]*(y[1]-y[0])", "y[0]*(p[1]-y[2])-y[1]", "y[0]*y[1]-p[2]*y[2]"], [10, 28, 8/3])
G = sdb.ops.ExprDiffusion([["p[0]*y[0]", "0", "0"], ["0", "p[0]*y[1]", "0"], ["0", "0", "p[0]*y[2]"]], [0.1],
                          diagonal=True)
y = sdb.itoint(f, G, [1.0, 1.0, 20.0], tspan, paths=10**5, seed=2, save_every=100); print(y.shape, y[:, -1].mean(axis=0))
y, obs = sdb.itoint(f, G, [1.0, 1.0, 20.0], tspan, paths=10**5, seed=2, save_every=100,
                    observe=[("first_above", 0, 15.0), ("max", 2)])
print(obs.shape, np.isfinite(obs[:, 0]).mean(), obs[:, 1].mean())
_, GL, y0L = sdb.ops.sys_lin(10, 10)
fN = sdb.ops.ExprDrift(["-1.5*y[%d] + 0.3*y[%d] - 0.2*y[%d]*y[%d]*y[%d]" % (i, (i + 1) % 10, i, i, i) for i in range(10)])
yN = sdb.itoSRI2(fN, GL, y0L, tspan, paths=10**5, seed=3, save_every=100); print(yN.shape, sdb.last_kernel())
