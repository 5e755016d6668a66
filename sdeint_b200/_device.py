"""Device plumbing for the shim: torch owns memory and streams, libsdb owns the kernels."""
from __future__ import annotations

import ctypes as C
import hashlib
import threading
import weakref

import numpy as np

from . import _lib, ops

_systems = {}            # insertion-ordered: least recently used first
_SYSTEMS_MAX = 64
_systems_lock = threading.Lock()


def torch_mod():
    import torch
    return torch


def require_cuda(device=None):
    """Resolve the CUDA device; fail loudly when there is none (no CPU fallback)."""
    torch = torch_mod()
    if not torch.cuda.is_available():
        raise RuntimeError(
            "sdeint_b200: no CUDA device available.  This package integrates on a B200 "
            "(sm_100a) only; there is no CPU fallback.")
    _lib.load()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    else:
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("sdeint_b200: device must be a CUDA device, got %s" % device)
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
    return device


def stream_ptr():
    return C.c_void_p(torch_mod().cuda.current_stream().cuda_stream)


def to_device(arr, device):
    """Host ndarray (float64) -> device tensor, via pinned memory for large arrays."""
    torch = torch_mod()
    if isinstance(arr, torch.Tensor):
        return arr.to(device=device, dtype=torch.float64)
    a = np.ascontiguousarray(arr, dtype=np.float64)
    t = torch.from_numpy(a)
    if a.nbytes >= (1 << 20):
        t = t.pin_memory()
    return t.to(device, non_blocking=True)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def host_ptr(a):
    return C.c_void_p(a.ctypes.data)


class System:
    """Owner of one sdb_system handle (drift + diffusion on the device)."""

    def __init__(self, f, G, device):
        lib = _lib.load()
        self.d, self.m = G.d, G.m      # (f and G themselves are not kept: the cache is keyed by content)
        fp = np.ascontiguousarray(f.params, dtype=np.float64)
        gp = np.ascontiguousarray(G.params, dtype=np.float64)
        handle = C.c_void_p()
        user = f.kind == ops.DRIFT_USER or G.kind in (ops.DIFF_USER, ops.DIFF_USER_DIAG)
        torch = torch_mod()
        with torch.cuda.device(device):
            if user:
                src = "\n".join(s for s in dict.fromkeys(
                    [getattr(f, "source", None), getattr(G, "source", None)]) if s)
                rc = lib.sdb_system_nvrtc(src.encode(), f.kind, host_ptr(fp), fp.size, G.kind,
                                          host_ptr(gp), gp.size, self.d, self.m, C.byref(handle))
            else:
                rc = lib.sdb_system_builtin(f.kind, host_ptr(fp), fp.size, G.kind, host_ptr(gp),
                                            gp.size, self.d, self.m, C.byref(handle))
        _lib.check(rc)
        self.handle = handle
        self._fin = weakref.finalize(self, lib.sdb_system_destroy, handle)


def _op_key(op):
    p = np.ascontiguousarray(op.params, dtype=np.float64)
    return (int(op.kind), hashlib.sha256(p.tobytes()).hexdigest(), getattr(op, "source", None))


def get_system(f, G, device):
    """Device systems cached by CONTENT -- operator kinds, parameter values, CUDA source, device -- in a
    small LRU: a parameter sweep that builds a new operator per point reuses the handle when the values
    repeat and drops the oldest handles (device buffers, fragment tables, NVRTC modules) otherwise, and
    changing `f.params` / `f.A` between calls is seen (it changes the key)."""
    key = (_op_key(f), _op_key(G), int(G.d), int(G.m), str(device))
    with _systems_lock:
        hit = _systems.pop(key, None)
        if hit is None:
            hit = System(f, G, device)
        _systems[key] = hit                      # most recently used last
        while len(_systems) > _SYSTEMS_MAX:
            _systems.pop(next(iter(_systems)))
        return hit
