"""Device plumbing for the shim: torch owns memory and streams, libsdb owns the kernels."""
from __future__ import annotations

import ctypes as C
import threading
import weakref

import numpy as np

from . import _lib, ops

_systems = {}
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
        self.f, self.G = f, G
        self.d, self.m = G.d, G.m
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


def get_system(f, G, device):
    """Cache systems per (operator objects, device) so repeated calls reuse compiled kernels."""
    key = (id(f), id(G), str(device))
    with _systems_lock:
        hit = _systems.get(key)
        if hit is not None and hit[0]() is f and hit[1]() is G:
            return hit[2]
        sysobj = System(f, G, device)
        _systems[key] = (weakref.ref(f), weakref.ref(G), sysobj)
        if len(_systems) > 256:
            for k in [k for k, v in _systems.items() if v[0]() is None or v[1]() is None]:
                _systems.pop(k, None)
        return sysobj
