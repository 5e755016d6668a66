"""Throughput of the generic loop kernel on a NONLINEAR user system given as CUDA source (NVRTC):
stochastic Lorenz-63 with multiplicative non-commutative noise, d = m = 3, itoSRI2 + Ikpw(n=5).
    python tools/bench_user.py [paths] [steps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdeint_b200 as sdb

SRC = r"""
__device__ void sde_drift(const double* y, double t, const double* p, double* out) {
  out[0] = p[0] * (y[1] - y[0]);
  out[1] = y[0] * (p[1] - y[2]) - y[1];
  out[2] = y[0] * y[1] - p[2] * y[2];
}
__device__ void sde_diffusion_col(int k, const double* y, double t, const double* p, double* out) {
  // g_0 = s (y0, 0, 0), g_1 = s (0, y1 + 0.1 y0, 0), g_2 = s (0.1 y1, 0, y2): non-commutative
  out[0] = out[1] = out[2] = 0.0;
  if (k == 0) out[0] = p[0] * y[0];
  else if (k == 1) out[1] = p[0] * (y[1] + 0.1 * y[0]);
  else { out[0] = 0.1 * p[0] * y[1]; out[2] = p[0] * y[2]; }
}
"""
P = int(sys.argv[1]) if len(sys.argv) > 1 else 4000000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
f = sdb.ops.CudaDrift(SRC, 3, [10.0, 28.0, 8.0 / 3.0])
G = sdb.ops.CudaDiffusion(SRC, 3, 3, [0.1])
y0 = np.array([1.0, 1.0, 20.0])
tspan = np.linspace(0.0, T * 1e-3, T + 1)
best = None
for rep in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    y = sdb.itoSRI2(f, G, y0, tspan, paths=P, seed=7, save_every=T, out="torch")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    best = ms if best is None else min(best, ms)
print("user Lorenz-63 d=m=3 itoSRI2+Ikpw: %d paths x %d steps in %.1f ms -> %.3e path-steps/s; mean y(T) = %s"
      % (P, T, best, P * T / best * 1e3, y[-1].mean(dim=1).tolist()))
