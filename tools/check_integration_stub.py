"""Runs the ctypes stub printed in INTEGRATION.md against the shipping shim (GPU needed):
    python tools/check_integration_stub.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# sdeint/_sdb.py -- minimal ctypes binding of include/sdb.h
import ctypes as C, numpy as np, torch

_lib = C.CDLL(os.path.join(ROOT, "sdeint_b200", "libsdb.so"))
_lib.sdb_last_error.restype = C.c_char_p

class IntegrateArgs(C.Structure):              # == sdb_integrate_args, field for field
    _fields_ = [("scheme", C.c_int), ("noise", C.c_int), ("flags", C.c_int), ("n_terms", C.c_int),
                ("paths", C.c_int64), ("path_id0", C.c_int64), ("seed", C.c_uint64),
                ("N", C.c_int), ("save_every", C.c_int), ("h_tspan", C.c_void_p),
                ("d_y0", C.c_void_p), ("y0_stride_comp", C.c_int64), ("y0_stride_path", C.c_int64),
                ("d_y", C.c_void_p), ("y_stride_save", C.c_int64), ("y_stride_comp", C.c_int64),
                ("y_stride_path", C.c_int64),
                ("d_dW", C.c_void_p), ("dW_stride_step", C.c_int64), ("dW_stride_comp", C.c_int64),
                ("dW_stride_path", C.c_int64),
                ("d_IJ", C.c_void_p), ("IJ_stride_step", C.c_int64), ("IJ_stride_row", C.c_int64),
                ("IJ_stride_col", C.c_int64), ("IJ_stride_path", C.c_int64),
                ("h_kp2is", C.c_void_p), ("rtol", C.c_double),
                ("d_status", C.c_void_p), ("stream", C.c_void_p)]

def _check(rc):
    if rc:
        raise RuntimeError(_lib.sdb_last_error().decode())

def linear_system(A, Bcols):
    """f = A y, g_k = B_k y  ->  opaque sdb_system*  (SDB_DRIFT_LINEAR=0, SDB_DIFF_LINEAR_COLS=1)."""
    A = np.ascontiguousarray(A, float); B = np.ascontiguousarray(Bcols, float)   # (m, d, d)
    h = C.c_void_p()
    _check(_lib.sdb_system_builtin(0, A.ctypes.data_as(C.c_void_p), C.c_int64(A.size),
                                   1, B.ctypes.data_as(C.c_void_p), C.c_int64(B.size),
                                   A.shape[0], B.shape[0], C.byref(h)))
    return h

def srk2(system, y0, tspan, dW=None, IJ=None, strat=False, paths=1, seed=0, save_every=1):
    """Drop-in for the loop of _Roessler2010_SRK2 (integrate.py:494-522) over `paths` paths.
    dW (N-1, m), IJ (N-1, m, m) as in the reference, or None -> on-device Philox + Ikpw/Jkpw."""
    tspan = np.ascontiguousarray(tspan, float); N = len(tspan); d = len(y0)
    ns = (N - 1) // save_every + 1
    y = torch.empty((ns, d, paths), dtype=torch.float64, device="cuda")    # path-minor
    y0d = torch.as_tensor(np.asarray(y0, float), device="cuda")
    st = torch.empty(4, dtype=torch.int64, device="cuda")
    a = IntegrateArgs(scheme=2, noise=1 if dW is None else 0,
                      flags=1 | (2 if strat else 0), n_terms=5, paths=paths, path_id0=0, seed=seed,
                      N=N, save_every=save_every, h_tspan=tspan.ctypes.data,
                      d_y0=y0d.data_ptr(), d_y=y.data_ptr(),
                      y_stride_save=d * paths, y_stride_comp=paths, y_stride_path=1,
                      d_status=st.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    keep = []
    if dW is not None:                       # the reference's own arrays, no transpose needed:
        m = dW.shape[-1]                     # strides address (N-1, m) / (N-1, m, m) directly
        dWd = torch.as_tensor(np.ascontiguousarray(dW, float), device="cuda"); keep.append(dWd)
        IJd = torch.as_tensor(np.ascontiguousarray(IJ, float), device="cuda"); keep.append(IJd)
        a.d_dW, a.dW_stride_step, a.dW_stride_comp, a.dW_stride_path = dWd.data_ptr(), m, 1, 0
        a.d_IJ, a.IJ_stride_step, a.IJ_stride_row, a.IJ_stride_col, a.IJ_stride_path = \
            IJd.data_ptr(), m * m, m, 1, 0
    _check(_lib.sdb_integrate(system, C.byref(a)))
    return y.permute(2, 0, 1).cpu().numpy()   # (paths, n_saved, d); [0] is the reference's (N, d)


if __name__ == "__main__":
    import sdeint_b200 as sdb
    f, G, y0 = sdb.ops.sys_r74(False)
    tspan = np.linspace(0.0, 0.1, 21)
    system = linear_system(f.A, G.B)
    # host-supplied noise in the reference's layout
    rng = np.random.default_rng(0)
    dW = rng.normal(0.0, np.sqrt(0.005), (20, 4))
    I = rng.normal(0.0, 0.005, (20, 4, 4))
    y_stub = srk2(system, y0, tspan, dW, I)[0]
    y_shim = sdb.itoSRI2(f, G, y0, tspan, dW=dW, I=I)
    assert y_stub.shape == y_shim.shape == (21, 5) and np.array_equal(y_stub, y_shim)
    # on-device noise, ensemble
    y2 = srk2(system, y0, tspan, paths=100, seed=5)
    # (Roessler 7.4 has commuting B_k: the shim would skip the Levy areas; the stub asks for Ikpw)
    y3 = sdb.itoSRI2(f, G, y0, tspan, paths=100, seed=5, commutative=False)
    assert np.array_equal(y2, y3)
    y4 = sdb.itoSRI2(f, G, y0, tspan, paths=100, seed=5)          # area-free dispatch: same paths
    assert np.max(np.abs(y4 - y3)) <= 1e-12 * np.max(np.abs(y3))
    print("INTEGRATION.md stub ok")
