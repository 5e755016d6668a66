"""Wiener increments and repeated stochastic integrals on the GPU.

Same names and signatures as sdeint/wiener.py -- `deltaW(N, m, h, generator=None)`
(wiener.py:36) and `Ikpw / Jkpw / Iwik / Jwik(dW, h, n=5, generator=None)` (wiener.py:77,
:111, :234, :285) returning the 2-tuples `(A, I)`, `(A, J)`, `(Atilde, I)`, `(Atilde, J)` --
plus keyword-only ensemble extensions (`paths=`, `seed=`, `path_id0=`, `normals=`, `out=`).

Random numbers come from an on-device Philox4x32-10 generator keyed by `seed` with counter
(global path id, step, draw slot), NOT from numpy's PCG64: a `numpy.random.Generator` passed
as `generator` is advanced by one draw to derive the Philox key, so a seeded generator gives
reproducible results, but not the reference's numbers.  Parity with the reference's
CONSTRUCTION is exact in the tested sense (<= 1e-12) when the standard normals the reference
would draw are supplied through `normals=(X, Y)` / `normals=(X, Y, Gt)`.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _device as dev
from . import _lib

NOISE_KPW = 1
NOISE_WIK = 2
NOISE_COMM = 3
_U63 = (1 << 63) - 1


def _derive_seed(seed, generator):
    """Philox key: explicit `seed`, else one draw from `generator` (fresh default_rng() if
    None, like wiener.py:50-51)."""
    if seed is not None:
        return int(seed) & 0xFFFFFFFFFFFFFFFF
    if generator is None:
        generator = np.random.default_rng()
    return int(generator.integers(0, _U63, dtype=np.int64))


def deltaW(N, m, h, generator=None, *, paths=None, seed=None, path_id0=0, device=None,
           out="numpy"):
    """Wiener increments for m independent processes over N intervals of length h
    (sdeint/wiener.py:36-52).

    Returns shape (N, m); with `paths=P`, shape (P, N, m) (a view of the path-minor device
    layout [N][m][P]).  `out="torch"` returns the device tensor (N, m, P) instead.
    """
    device = dev.require_cuda(device)
    torch = dev.torch_mod()
    N, m = int(N), int(m)
    P = 1 if paths is None else int(paths)
    if N < 0 or m < 1 or P < 1:
        raise ValueError("deltaW: need N >= 0, m >= 1, paths >= 1")
    key = _derive_seed(seed, generator)
    with torch.cuda.device(device):
        buf = torch.empty((N, m, P), dtype=torch.float64, device=device)
        _lib.check(_lib.load().sdb_delta_w(N, m, float(h), P, int(path_id0), key, dev.ptr(buf),
                                           m * P, P, 1, dev.stream_ptr()))
    if out == "torch":
        return buf
    host = buf.cpu().numpy()
    if paths is None:
        return np.ascontiguousarray(host[:, :, 0])
    return host.transpose(2, 0, 1)


def _prep_dW(dW, ensemble=None):
    """Accept (N, m) or (N, m, 1) like the reference (bare ValueError otherwise,
    wiener.py:98-101, :255-258), or an ensemble (P, N, m); `ensemble=True` resolves the
    ambiguous (P, N, 1).  Returns (array (P,N,m), kind)."""
    torch = dev.torch_mod()
    if isinstance(dW, torch.Tensor):
        raise TypeError("pass device tensors through the ensemble API (out='torch')")
    dW = np.asarray(dW, dtype=np.float64)
    if dW.ndim == 2:
        return dW[None], "single"
    if dW.ndim == 3 and ensemble:
        return dW, "ensemble"
    if dW.ndim == 3 and dW.shape[2] == 1:
        return dW[None, :, :, 0], "single"
    if dW.ndim == 3:
        return dW, "ensemble"
    raise ValueError


def _repint(method, strat, dW, h, n, generator, seed, path_id0, normals, device, ensemble):
    device = dev.require_cuda(device)
    torch = dev.torch_mod()
    arr, kind = _prep_dW(dW, ensemble)
    P, N, m = arr.shape
    n = int(n)
    if n < 1 and method != NOISE_COMM:
        raise ValueError("need at least one series term")
    M = m * (m - 1) // 2
    need_rng = normals is None and method != NOISE_COMM and not (method == NOISE_WIK and m == 1)
    key = _derive_seed(seed, generator) if need_rng else 0
    lib = _lib.load()
    with torch.cuda.device(device):
        d_dW = dev.to_device(arr, device)                      # (P, N, m)
        nA = m * m if method != NOISE_WIK else max(M, 1)
        d_A = torch.empty((P, N, nA), dtype=torch.float64, device=device)
        d_IJ = torch.empty((P, N, m, m), dtype=torch.float64, device=device)
        a = _lib.RepintArgs()
        a.method, a.stratonovich, a.n_terms, a.m = method, int(strat), n, m
        a.steps, a.paths, a.path_id0, a.seed, a.h = N, P, int(path_id0), key, float(h)
        a.d_dW = dev.ptr(d_dW)
        a.dW_stride_step, a.dW_stride_comp, a.dW_stride_path = m, 1, N * m
        keep = []
        if normals is not None and method != NOISE_COMM and not (method == NOISE_WIK and m == 1):
            X = np.asarray(normals[0], dtype=np.float64)
            Y = np.asarray(normals[1], dtype=np.float64)
            if X.ndim == 3:                                    # (n, N, m) single path
                X, Y = X[:, None], Y[:, None]
            if X.shape != (n, P, N, m) or Y.shape != X.shape:
                raise ValueError("normals X, Y must have shape (n, [P,] N, m)")
            d_X, d_Y = dev.to_device(X, device), dev.to_device(Y, device)
            keep += [d_X, d_Y]
            a.d_X, a.d_Y = dev.ptr(d_X), dev.ptr(d_Y)
            a.XY_stride_term, a.XY_stride_path = P * N * m, N * m
            a.XY_stride_step, a.XY_stride_comp = m, 1
            if method == NOISE_WIK:
                if len(normals) < 3:
                    raise ValueError("Wiktorsson needs normals=(X, Y, Gt)")
                Gt = np.asarray(normals[2], dtype=np.float64)
                if Gt.ndim == 2:
                    Gt = Gt[None]
                if Gt.shape != (P, N, M):
                    raise ValueError("normals Gt must have shape ([P,] N, m(m-1)/2)")
                d_G = dev.to_device(Gt, device)
                keep.append(d_G)
                a.d_Gt = dev.ptr(d_G)
                a.Gt_stride_path, a.Gt_stride_step, a.Gt_stride_pair = N * M, M, 1
        a.d_A = dev.ptr(d_A)
        a.A_stride_path, a.A_stride_step, a.A_stride_elem = N * nA, nA, 1
        a.d_IJ = dev.ptr(d_IJ)
        a.IJ_stride_path, a.IJ_stride_step, a.IJ_stride_row, a.IJ_stride_col = N * m * m, m * m, m, 1
        a.stream = dev.stream_ptr()
        _lib.check(lib.sdb_repeated_integrals(C.byref(a)))
        A = d_A.cpu().numpy()
        IJ = d_IJ.cpu().numpy()
    if method != NOISE_WIK:
        A = A.reshape(P, N, m, m)
    else:
        A = A.reshape(P, N, max(M, 1), 1)
    if kind == "single":
        return A[0], IJ[0]
    return A, IJ


def Ikpw(dW, h, n=5, generator=None, *, seed=None, path_id0=0, normals=None, device=None,
         ensemble=None):
    """Repeated Ito integrals by the Kloeden-Platen-Wright series (sdeint/wiener.py:77-108).
    Returns (A, I): Levy areas and I, shapes (N, m, m) (or (P, N, m, m) for an ensemble dW)."""
    return _repint(NOISE_KPW, False, dW, h, n, generator, seed, path_id0, normals, device, ensemble)


def Jkpw(dW, h, n=5, generator=None, *, seed=None, path_id0=0, normals=None, device=None,
         ensemble=None):
    """Stratonovich version, J = I + h/2 * 1 (sdeint/wiener.py:111-131)."""
    return _repint(NOISE_KPW, True, dW, h, n, generator, seed, path_id0, normals, device, ensemble)


def Iwik(dW, h, n=5, generator=None, *, seed=None, path_id0=0, normals=None, device=None,
         ensemble=None):
    """Repeated Ito integrals by Wiktorsson's method (sdeint/wiener.py:234-282), with the
    Sigma_inf tail correction applied in O(m^2) per step instead of the reference's dense
    (M x m^2)(m^2 x m^2)(m^2 x M) products.  Returns (Atilde (N, M, 1), I (N, m, m))."""
    return _repint(NOISE_WIK, False, dW, h, n, generator, seed, path_id0, normals, device, ensemble)


def Jwik(dW, h, n=5, generator=None, *, seed=None, path_id0=0, normals=None, device=None,
         ensemble=None):
    """Stratonovich version of Iwik (sdeint/wiener.py:285-305)."""
    return _repint(NOISE_WIK, True, dW, h, n, generator, seed, path_id0, normals, device, ensemble)


def Icomm(dW, h, n=0, generator=None, *, seed=None, path_id0=0, normals=None, device=None,
          ensemble=None):
    """Repeated Ito integrals for COMMUTATIVE noise: I = (dW dW^T - h 1)/2, no Levy areas and no
    random numbers beyond dW (`n`, `generator`, `seed` are accepted and ignored so that it can stand
    in for Ikpw/Iwik as `Imethod=`).  When L^j g_k = L^k g_j the antisymmetric part of I cancels in
    the order-1.0 term of the schemes -- the fast path the reference plans at
    sdeint/integrate.py:150-151.  Returns (A, I) with A == 0, shapes as Ikpw."""
    return _repint(NOISE_COMM, False, dW, h, 0, None, 0, path_id0, None, device, ensemble)


def Jcomm(dW, h, n=0, generator=None, *, seed=None, path_id0=0, normals=None, device=None,
          ensemble=None):
    """Stratonovich version of Icomm: J = dW dW^T / 2."""
    return _repint(NOISE_COMM, True, dW, h, 0, None, 0, path_id0, None, device, ensemble)


def _a_tail(n):
    """pi^2/6 - sum_{k<=n} 1/k^2 (sdeint/wiener.py:229-231)."""
    return np.pi ** 2 / 6.0 - sum(1.0 / (k * k) for k in range(1, int(n) + 1))


def series_terms(h, m, method="kpw", C=1.0, n_max=100000):
    """Number of series terms n for which the repeated integrals keep the schemes at strong order 1.0
    as h -> 0: the mean-square error of the approximation must not exceed C h^3 per step
    (Kloeden & Platen 1992, sec. 10.3).  The reference fixes n = 5 (wiener.py:77, :234), which is a
    constant relative accuracy of the Levy areas and degrades non-commutative problems to order 1/2
    for small h.

    "kpw"  the truncated series of Ikpw/Jkpw as the reference sums it (no tail correction): each
           Levy area has E|A_ij - A_ij^(n)|^2 = 3 h^2 a(n) / (2 pi^2), a(n) = pi^2/6 - sum_{k<=n} 1/k^2
           ~ 1/n (derived from wiener.py:67-74; checked numerically in tests/test_host.py), so
           n ~ 3 / (2 pi^2 C h).
    "wik"  Wiktorsson's tail-corrected approximation (Iwik/Jwik): the published bound
           sum_{i<j} E|A_ij - A_ij^(n)|^2 <= 5 m^2 (m - 1) h^2 / (24 pi^2 n^2) (Wiktorsson 2001, Thm 4.1),
           so n = ceil(sqrt(5 m^2 (m - 1) / (24 pi^2 C h))) -- O(h^-1/2) instead of O(h^-1).
    Scalar noise (m = 1) has no Levy areas: 1 term."""
    h, m = float(h), int(m)
    if not h > 0.0 or m < 1 or not C > 0.0:
        raise ValueError("series_terms: need h > 0, m >= 1, C > 0")
    if m == 1:
        return 1
    if method in ("kpw", Ikpw, Jkpw):
        target = 2.0 * np.pi ** 2 * C * h / 3.0
        n = max(1, int(3.0 / (2.0 * np.pi ** 2 * C * h)) - 2)
        a = _a_tail(n)
        while a > target and n < n_max:
            n += 1
            a -= 1.0 / (n * n)
        return n
    if method in ("wik", Iwik, Jwik):
        return max(1, int(np.ceil(np.sqrt(5.0 * m * m * (m - 1) / (24.0 * np.pi ** 2 * C * h)))))
    raise ValueError("series_terms: method must be 'kpw' or 'wik'")
