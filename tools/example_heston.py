"""Heston stochastic-volatility model written exactly as one would for mattja/sdeint -- plain Python f and G --
integrated as an ensemble on the GPU: European call price by Monte Carlo against the semi-closed form is out of
scope here; the script checks the martingale E[S_T] = S_0 exp(r T) and reports the throughput.
    python tools/example_heston.py [paths] [steps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdeint_b200 as sdb

P = int(sys.argv[1]) if len(sys.argv) > 1 else 4000000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 500
r, kappa, theta, xi, rho = 0.03, 2.0, 0.04, 0.3, -0.7
s0, v0 = 100.0, 0.04


def f(y, t):                      # y = (S, v); full truncation: v+ = max(v, 0)
    return np.array([r * y[0], kappa * (theta - np.maximum(y[1], 0.0))])


def G(y, t):                      # correlated Brownian motions through a Cholesky factor
    sv = np.sqrt(np.maximum(y[1], 0.0))
    return np.array([[sv * y[0], 0.0],
                     [xi * sv * rho, xi * sv * np.sqrt(1.0 - rho ** 2)]])


tspan = np.linspace(0.0, 1.0, T + 1)
y = sdb.itoint(f, G, np.array([s0, v0]), tspan, paths=P, seed=1, save_every=T, out="torch")   # compile + warm up
torch.cuda.synchronize()
t0 = time.perf_counter()
y = sdb.itoint(f, G, np.array([s0, v0]), tspan, paths=P, seed=2, save_every=T, out="torch")
torch.cuda.synchronize()
dt = time.perf_counter() - t0
ST = y[-1, 0]
mean, err = float(ST.mean()), float(ST.std() / np.sqrt(P))
print("Heston, itoint (SRI2 + Ikpw), %d paths x %d steps: %.3f s -> %.3e path-steps/s  [%s]"
      % (P, T, dt, P * T / dt, sdb.last_kernel()))
print("E[S_T] = %.4f +- %.4f   (S_0 e^{rT} = %.4f)   E[v_T] = %.5f" % (mean, err, s0 * np.exp(r), float(y[-1, 1].mean())))
assert abs(mean - s0 * np.exp(r)) < 6 * err + 0.02
