"""ctypes binding of libsdb.so (include/sdb.h).  Loading is lazy and LOUD: if the library
is missing or a symbol is not exported, every compute entry point raises -- there is no
CPU fallback on this path."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsdb.so")

c_i64 = C.c_int64
c_u64 = C.c_uint64
c_dp = C.c_void_p   # device / host double* passed as integer addresses


class IntegrateArgs(C.Structure):
    _fields_ = [
        ("scheme", C.c_int), ("noise", C.c_int), ("flags", C.c_int), ("n_terms", C.c_int),
        ("paths", c_i64), ("path_id0", c_i64), ("seed", c_u64),
        ("N", C.c_int), ("save_every", C.c_int),
        ("h_tspan", c_dp),
        ("d_y0", c_dp), ("y0_stride_comp", c_i64), ("y0_stride_path", c_i64),
        ("d_y", c_dp), ("y_stride_save", c_i64), ("y_stride_comp", c_i64), ("y_stride_path", c_i64),
        ("d_dW", c_dp), ("dW_stride_step", c_i64), ("dW_stride_comp", c_i64), ("dW_stride_path", c_i64),
        ("d_IJ", c_dp), ("IJ_stride_step", c_i64), ("IJ_stride_row", c_i64),
        ("IJ_stride_col", c_i64), ("IJ_stride_path", c_i64),
        ("h_kp2is", c_dp), ("rtol", C.c_double),
        ("d_status", c_dp), ("stream", c_dp),
    ]


class RepintArgs(C.Structure):
    _fields_ = [
        ("method", C.c_int), ("stratonovich", C.c_int), ("n_terms", C.c_int), ("m", C.c_int),
        ("steps", c_i64), ("paths", c_i64), ("path_id0", c_i64), ("seed", c_u64),
        ("h", C.c_double),
        ("d_dW", c_dp), ("dW_stride_step", c_i64), ("dW_stride_comp", c_i64), ("dW_stride_path", c_i64),
        ("d_X", c_dp), ("d_Y", c_dp),
        ("XY_stride_term", c_i64), ("XY_stride_step", c_i64), ("XY_stride_comp", c_i64),
        ("XY_stride_path", c_i64),
        ("d_Gt", c_dp), ("Gt_stride_step", c_i64), ("Gt_stride_pair", c_i64), ("Gt_stride_path", c_i64),
        ("d_A", c_dp), ("A_stride_step", c_i64), ("A_stride_elem", c_i64), ("A_stride_path", c_i64),
        ("d_IJ", c_dp), ("IJ_stride_step", c_i64), ("IJ_stride_row", c_i64),
        ("IJ_stride_col", c_i64), ("IJ_stride_path", c_i64),
        ("stream", c_dp),
    ]


# name -> (restype, argtypes); must list every symbol include/sdb.h declares
class Observer(C.Structure):
    """== sdb_observer (include/sdb.h)."""
    _fields_ = [("kind", C.c_int), ("comp", C.c_int), ("level", C.c_double)]


class IntegrateExt(C.Structure):
    """== sdb_integrate_ext (include/sdb.h)."""
    _fields_ = [("step0", c_i64), ("h_noise", C.c_double), ("n_obs", C.c_int),
                ("h_obs", C.POINTER(Observer)), ("d_obs", C.c_void_p),
                ("obs_stride_obs", c_i64), ("obs_stride_path", c_i64)]


SIGNATURES = {
    "sdb_last_error": (C.c_char_p, []),
    "sdb_version": (C.c_int, []),
    "sdb_last_kernel": (C.c_char_p, []),
    "sdb_launch_count": (c_u64, []),
    "sdb_jit_cache_stats": (None, [C.POINTER(C.c_long), C.POINTER(C.c_long)]),
    "sdb_system_builtin": (C.c_int, [C.c_int, c_dp, c_i64, C.c_int, c_dp, c_i64, C.c_int, C.c_int,
                                     C.POINTER(C.c_void_p)]),
    "sdb_system_nvrtc": (C.c_int, [C.c_char_p, C.c_int, c_dp, c_i64, C.c_int, c_dp, c_i64, C.c_int,
                                   C.c_int, C.POINTER(C.c_void_p)]),
    "sdb_system_destroy": (C.c_int, [C.c_void_p]),
    "sdb_integrate": (C.c_int, [C.c_void_p, C.POINTER(IntegrateArgs)]),
    "sdb_integrate_observed": (C.c_int, [C.c_void_p, C.POINTER(IntegrateArgs), C.POINTER(Observer), C.c_int,
                                         c_dp, c_i64, c_i64]),
    "sdb_integrate_segment": (C.c_int, [C.c_void_p, C.POINTER(IntegrateArgs), C.POINTER(IntegrateExt)]),
    "sdb_delta_w": (C.c_int, [c_i64, C.c_int, C.c_double, c_i64, c_i64, c_u64, c_dp, c_i64, c_i64,
                              c_i64, c_dp]),
    "sdb_repeated_integrals": (C.c_int, [C.POINTER(RepintArgs)]),
    "sdb_moments": (C.c_int, [c_dp, c_i64, C.c_int, c_i64, c_dp, c_dp, C.c_int, C.c_double,
                              C.c_double, c_dp]),
    "sdb_moments_welford": (C.c_int, [c_dp, c_i64, C.c_int, c_i64, c_dp, c_dp]),
    "sdb_cross_moments": (C.c_int, [c_dp, c_i64, C.c_int, c_i64, c_dp, c_dp]),
    "sdb_copy2d_d2h_async": (C.c_int, [c_dp, c_i64, c_dp, c_i64, c_i64, c_i64, c_dp]),
    "sdb_fp64_peak": (C.c_int, [C.c_int, C.POINTER(C.c_double), c_dp]),
    "sdb_philox_block": (C.c_int, [c_u64, c_u64, C.c_uint32, C.c_uint32,
                                   C.POINTER(C.c_uint32 * 4)]),
    "sdb_random_block": (C.c_int, [c_u64, c_u64, C.c_uint32, C.c_uint32,
                                   C.POINTER(C.c_uint32 * 4)]),
}

_lib = None


class SdbError(RuntimeError):
    """An error reported by libsdb (message from sdb_last_error())."""


def load():
    """dlopen libsdb.so and bind every entry point; raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "sdeint_b200: %s is missing -- build it with `python __graft_entry__.py` "
            "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name, None)
        if fn is None:
            raise RuntimeError("sdeint_b200: libsdb.so does not export %s" % name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise SdbError(load().sdb_last_error().decode("utf-8", "replace"))


def launch_count():
    return int(load().sdb_launch_count())


def jit_cache_stats():
    """(hits, misses): NVRTC-built kernels of this process loaded from the on-disk cache / compiled."""
    h, m = C.c_long(0), C.c_long(0)
    load().sdb_jit_cache_stats(C.byref(h), C.byref(m))
    return int(h.value), int(m.value)


def last_kernel():
    """Name of the kernel the last integrator call of this thread launched (include/sdb.h:
    sdb_last_kernel): an ahead-of-time kernel's symbol or "jit:<key>"."""
    return load().sdb_last_kernel().decode()
