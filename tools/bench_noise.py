"""Throughput of the standalone noise kernels (deltaW, Ikpw, Iwik) on one GPU, device-timed.
    python tools/bench_noise.py
Units: (path, step) intervals per second; bytes = what the kernel writes (and reads)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdeint_b200 as sdb
from sdeint_b200 import _lib, _device as dev
import ctypes as C


def timed(fn, reps=4):
    best = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return best


lib = _lib.load()
out = []
for m, P, N in ((10, 200000, 100), (2, 1000000, 100)):
    buf = torch.empty((N, m, P), dtype=torch.float64, device="cuda")
    ms = timed(lambda: _lib.check(lib.sdb_delta_w(N, m, 1e-3, P, 0, 7, dev.ptr(buf), m * P, P, 1, dev.stream_ptr())))
    out.append({"kernel": "deltaW", "m": m, "units": P * N, "ms": ms, "units_per_s": P * N / ms * 1e3,
                "GBps_written": 8.0 * m * P * N / ms / 1e6})
for meth, name in ((1, "Ikpw"), (2, "Iwik")):
    for m, P, N in ((10, 100000, 20), (20, 20000, 20)):
        M = m * (m - 1) // 2
        dW = torch.empty((N, m, P), dtype=torch.float64, device="cuda")
        _lib.check(lib.sdb_delta_w(N, m, 1e-3, P, 0, 7, dev.ptr(dW), m * P, P, 1, dev.stream_ptr()))
        IJ = torch.empty((N, m, m, P), dtype=torch.float64, device="cuda")
        nA = m * m if meth == 1 else max(M, 1)
        A = torch.empty((N, nA, P), dtype=torch.float64, device="cuda")
        r = _lib.RepintArgs()
        r.method, r.stratonovich, r.n_terms, r.m = meth, 0, 5, m
        r.steps, r.paths, r.path_id0, r.seed, r.h = N, P, 0, 7, 1e-3
        r.d_dW = dev.ptr(dW); r.dW_stride_step, r.dW_stride_comp, r.dW_stride_path = m * P, P, 1
        r.d_A = dev.ptr(A); r.A_stride_step, r.A_stride_elem, r.A_stride_path = nA * P, P, 1
        r.d_IJ = dev.ptr(IJ)
        r.IJ_stride_step, r.IJ_stride_row, r.IJ_stride_col, r.IJ_stride_path = m * m * P, m * P, P, 1
        r.stream = dev.stream_ptr()
        ms = timed(lambda: _lib.check(lib.sdb_repeated_integrals(C.byref(r))))
        byts = 8.0 * (m * m + nA + m) * P * N
        out.append({"kernel": name, "m": m, "units": P * N, "ms": ms, "units_per_s": P * N / ms * 1e3,
                    "GBps": byts / ms / 1e6})
for o in out:
    print(json.dumps(o), flush=True)
