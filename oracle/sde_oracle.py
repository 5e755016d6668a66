"""CPU oracle for the sdeint hot path (TEST INFRASTRUCTURE -- not product code).

This module restates, in plain numpy with explicit loops, the arithmetic of the
reference `mattja/sdeint` (v0.3.1-dev) for the path this repository accelerates:

    integrators  : sdeint/integrate.py:235-239 (itoEuler), :292-302 (stratHeun),
                   :494-522 (_Roessler2010_SRK2 = itoSRI2/stratSRS2),
                   :613-645 (stratKP2iS)
    noise inputs : sdeint/wiener.py:36-52 (deltaW), :67-74 + :102-106 (Ikpw),
                   :130 (Jkpw), :206-282 (Iwik), :304 (Jwik), :229-231 (_a)

Who may use it: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs, and there only as the checker or the
timed CPU baseline.  The product package `sdeint_b200` never imports it.

Pinning: the reference publishes no golden vectors ("parity
unpinned" by the reference's own tests; the single literal constant is
`_a(5) == 0.181322955737115`, sdeint/tests/test_wiener.py:91).  The oracle is
therefore pinned against OUTPUTS OF THE REFERENCE ITSELF: `oracle/make_golden.py`
imports the unmodified reference from /root/reference, runs every function on
seeded inputs and stores inputs+outputs under `tests/golden/`;
`tests/test_oracle.py` checks this restatement against those fixtures (and,
when /root/reference is importable, against the live reference as well).

The repeated-integral constructions are written ELEMENTWISE (no Kronecker /
permutation-matrix products), which is also the form the CUDA kernels use; the
vec/Kronecker helpers of the reference (`_vec,_unvec,_kp,_kp2,_P,_K`) are
restated separately at the bottom so that the index conventions are pinned too.
"""
from __future__ import annotations

import math

import numpy as np

__all__ = [
    "euler", "heun", "srk2", "kp2is",
    "a_tail", "draw_kpw_normals", "draw_wik_normals",
    "levy_kpw", "ikpw_from_normals", "jkpw_from_normals",
    "iwik_from_normals", "jwik_from_normals",
    "ikpw", "jkpw", "iwik", "jwik", "delta_w",
    "vec", "unvec", "kp", "kp2", "perm_P", "sel_K",
    "pair_list",
]


# --------------------------------------------------------------------------------------
# integrators (one sample path per call, like the reference)
# --------------------------------------------------------------------------------------

def _global_h(tspan):
    """h = (t[N-1]-t[0])/(N-1): sdeint/integrate.py:228, :285, :480, :594."""
    N = len(tspan)
    return (tspan[N - 1] - tspan[0]) / (N - 1)


def euler(f, G, y0, tspan, dW):
    """Euler-Maruyama, sdeint/integrate.py:235-239.

    y[n+1] = y[n] + f(y[n],t[n])*h + G(y[n],t[n]) @ dW[n]  with the GLOBAL mean h.
    """
    y0 = np.asarray(y0, dtype=np.float64)
    N = len(tspan)
    h = _global_h(tspan)
    out = np.empty((N, y0.shape[0]))
    out[0] = y0
    for n in range(N - 1):
        yn = out[n]
        tn = tspan[n]
        out[n + 1] = yn + f(yn, tn) * h + G(yn, tn).dot(dW[n])
    return out


def heun(f, G, y0, tspan, dW):
    """Stratonovich Heun, sdeint/integrate.py:292-302 (global h)."""
    y0 = np.asarray(y0, dtype=np.float64)
    N = len(tspan)
    h = _global_h(tspan)
    out = np.empty((N, y0.shape[0]))
    out[0] = y0
    for n in range(N - 1):
        yn = out[n]
        t0, t1 = tspan[n], tspan[n + 1]
        w = dW[n]
        f0 = f(yn, t0)
        G0 = G(yn, t0)
        ypred = yn + f0 * h + G0.dot(w)
        f1 = f(ypred, t1)
        G1 = G(ypred, t1)
        out[n + 1] = yn + 0.5 * (f0 + f1) * h + (0.5 * (G0 + G1)).dot(w)
    return out


def r2(f, G, y0, tspan, dW):
    """Burrage & Burrage (1996, "High strong order explicit Runge-Kutta methods for stochastic
    ordinary differential equations", Appl. Numer. Math. 22) two-stage method R2 -- not in the
    reference, which lists "Burrage and Burrage SRKs" as to-do (README.rst:125-128); restated from the
    published tableau A = B = [[0, 0], [2/3, 0]], alpha = gamma = (1/4, 3/4) for the Stratonovich
    equation dy = f dt + G o dW with the increment of every Wiener process in the stochastic part
    (strong order 1 for one Wiener process or commutative noise):
        Y2 = y + (2/3) (f(y) h + G(y) dW),   y' = y + (1/4)(f(y) h + G(y) dW) + (3/4)(f(Y2) h + G(Y2) dW),
    stage 2 evaluated at t + 2h/3; global h like the reference's stratHeun (integrate.py:285)."""
    y0 = np.asarray(y0, dtype=np.float64)
    N = len(tspan)
    h = _global_h(tspan)
    out = np.empty((N, y0.shape[0]))
    out[0] = y0
    for n in range(N - 1):
        yn = out[n]
        t0 = tspan[n]
        w = dW[n]
        k1 = f(yn, t0) * h + G(yn, t0).dot(w)
        y2 = yn + (2.0 / 3.0) * k1
        tm = t0 + (2.0 / 3.0) * h
        k2 = f(y2, tm) * h + G(y2, tm).dot(w)
        out[n + 1] = yn + 0.25 * k1 + 0.75 * k2
    return out


def srk2(f, G, y0, tspan, dW, IJ):
    """Roessler (2010) SRI2 / SRS2, sdeint/integrate.py:494-522.

    `G` is either one callable returning (d, m) or a sequence of m column
    callables (integrate.py:478, :503-507, :516-521).  The step size is the
    PER-STEP h = t[n+1]-t[n] (integrate.py:497).  `IJ` holds Ito integrals for
    SRI2 and Stratonovich integrals for SRS2 -- nothing else differs.
    """
    y0 = np.asarray(y0, dtype=np.float64)
    d = y0.shape[0]
    N = len(tspan)
    columns = None if callable(G) else tuple(G)
    m = dW.shape[1]
    out = np.empty((N, d))
    out[0] = y0
    for n in range(N - 1):
        t0, t1 = tspan[n], tspan[n + 1]
        h = t1 - t0
        rh = np.sqrt(h)
        yn = out[n]
        f0h = f(yn, t0) * h
        if columns is None:
            Gn = G(yn, t0)
        else:
            Gn = np.empty((d, m))
            for k in range(m):
                Gn[:, k] = columns[k](yn, t0)
        S = np.dot(Gn, IJ[n]) / rh                      # (d, m)  :508
        base = yn + f0h                                 # H20     :509
        f1h = f(base, t1) * h                           # :514
        ynew = yn + 0.5 * (f0h + f1h) + np.dot(Gn, dW[n])   # :515
        for k in range(m):                              # :516-521, in order k=0..m-1
            up = base + S[:, k]
            dn = base - S[:, k]
            if columns is None:
                diff = G(up, t1)[:, k] - G(dn, t1)[:, k]
            else:
                diff = columns[k](up, t1) - columns[k](dn, t1)
            ynew = ynew + 0.5 * rh * diff
        out[n + 1] = ynew
    return out


def kp2is(f, G, y0, tspan, dW, J, gam=None, al1=None, al2=None, rtol=1e-12,
          solver="fsolve"):
    """Kloeden-Platen two-step implicit strong-1.0 scheme, integrate.py:613-645.

    solver="fsolve" follows the reference (MINPACK hybrd through
    scipy.optimize.fsolve with xtol=rtol, integrate.py:638); solver="newton"
    is a dependency-free Newton iteration with a forward-difference Jacobian
    used to cross-check the device solver.  Default rtol=1e-12 because parity
    at 1e-12 needs the implicit equation solved that tightly.
    """
    y0 = np.asarray(y0, dtype=np.float64)
    d = y0.shape[0]
    N = len(tspan)
    m = dW.shape[1]
    gam = np.full(d, 0.5) if gam is None else np.asarray(gam, dtype=np.float64)
    al1 = np.full(d, 0.5) if al1 is None else np.asarray(al1, dtype=np.float64)
    al2 = np.full(d, 0.5) if al2 is None else np.asarray(al2, dtype=np.float64)
    out = np.zeros((N, d))
    out[0] = y0
    f_prev = None
    V_prev = None
    f_cur = None
    V_cur = None
    for n in range(N - 1):
        t0, t1 = tspan[n], tspan[n + 1]
        h = t1 - t0
        rh = np.sqrt(h)
        yn = out[n]
        f_prev, f_cur = f_cur, f(yn, t0)
        Gn = G(yn, t0)
        Ybar = (yn + f_cur * h).reshape(d, 1) + Gn * rh           # :624
        acc = np.zeros(d)
        for j1 in range(m):                                        # :626-627
            acc = acc + np.dot(G(Ybar[:, j1], t0) - Gn, J[n][j1, :])
        V_prev, V_cur = V_cur, np.dot(Gn, dW[n]) + acc / rh        # :629
        if n == 0:
            out[1] = yn + f_cur * h + V_cur                        # :632
            continue
        ym1 = out[n - 1]
        const = ((1 - gam) * yn + gam * ym1
                 + ((gam * al1 + (1 - al2)) * f_cur + gam * (1 - al1) * f_prev) * h
                 + V_cur + gam * V_prev)

        def resid(x):                                              # :603-609
            return const + al2 * f(x, t1) * h - x

        if solver == "fsolve":
            from scipy.optimize import fsolve
            sol, _info, status, msg = fsolve(resid, yn, xtol=rtol, full_output=True)
            if status != 1:
                # hybrd reports "not making good progress" when the start point is
                # already a root to rounding level; the reference raises there
                # (integrate.py:640-645).  The oracle accepts a converged iterate
                # so that tight-xtol runs do not depend on that stopping quirk.
                if np.linalg.norm(resid(sol)) > 1e-13 * max(1.0, np.linalg.norm(sol)):
                    raise RuntimeError("kp2is oracle: fsolve failed at t=%g: %s" % (t0, msg))
        else:
            sol = _newton_fd(resid, yn.copy(), rtol)
        out[n + 1] = sol
    return out


def _newton_fd(resid, x, xtol, maxit=50):
    d = x.shape[0]
    eps = math.sqrt(np.finfo(np.float64).eps)
    for _ in range(maxit):
        r = resid(x)
        Jm = np.empty((d, d))
        for j in range(d):
            step = eps * abs(x[j])
            if step == 0.0:
                step = eps
            xp = x.copy()
            xp[j] += step
            Jm[:, j] = (resid(xp) - r) / step
        dx = np.linalg.solve(Jm, -r)
        x = x + dx
        if np.linalg.norm(dx) <= xtol * np.linalg.norm(x):
            return x
    raise RuntimeError("kp2is oracle: Newton did not converge")


# --------------------------------------------------------------------------------------
# Wiener increments and repeated integrals
# --------------------------------------------------------------------------------------

def delta_w(N, m, h, generator):
    """sdeint/wiener.py:52."""
    return generator.normal(0.0, np.sqrt(h), (N, m))


def a_tail(n):
    """sum_{k>n} 1/k^2, sdeint/wiener.py:229-231; a_tail(5)=0.181322955737115."""
    s = 0.0
    for k in range(1, n + 1):
        s += 1.0 / k ** 2
    return np.pi ** 2 / 6.0 - s


def pair_list(m):
    """Pairs (i<j) in the row order of K_m (wiener.py:193-203): (0,1),(0,2),..,(1,2),.."""
    return [(i, j) for i in range(m) for j in range(i + 1, m)]


def draw_kpw_normals(generator, N, m, n):
    """Replay the reference's draw order for Ikpw (wiener.py:70-71, :102-104):
    for k=1..n: X_k = standard_normal((N,m,1)) then Y_k likewise.
    Returns X, Y of shape (n, N, m)."""
    X = np.empty((n, N, m))
    Y = np.empty((n, N, m))
    for k in range(n):
        X[k] = generator.standard_normal((N, m, 1))[:, :, 0]
        Y[k] = generator.standard_normal((N, m, 1))[:, :, 0]
    return X, Y


def draw_wik_normals(generator, N, m, n):
    """Draw order for Iwik (wiener.py:209-210, :264-266, :275): all X_k, Y_k as
    for Ikpw, THEN one standard_normal((N, M, 1)).  m == 1 draws nothing
    (wiener.py:259-260)."""
    M = m * (m - 1) // 2
    if m == 1:
        return np.empty((n, N, 1)), np.empty((n, N, 1)), np.empty((N, 0))
    X, Y = draw_kpw_normals(generator, N, m, n)
    Gt = generator.standard_normal((N, M, 1))[:, :, 0]
    return X, Y, Gt


def levy_kpw(dW, h, X, Y):
    """Truncated Levy-area series, elementwise form of wiener.py:67-74,:102-105.

    A[s,i,j] = (h/2pi) * sum_{k=1..n} (X_k[i] Z_k[j] - Z_k[i] X_k[j]) / k,
    Z_k = Y_k + sqrt(2/h) dW.  Accumulation order over k matches the reference
    (term for k=1 first, `A += term_k`, then one scaling by h/2pi).
    """
    dW = np.asarray(dW, dtype=np.float64)
    Ns, m = dW.shape
    n = X.shape[0]
    c = np.sqrt(2.0 / h)
    A = np.zeros((Ns, m, m))
    for k in range(1, n + 1):
        Xk = X[k - 1]
        Zk = Y[k - 1] + c * dW
        for i in range(m):
            for j in range(m):
                term = (Xk[:, i] * Zk[:, j] - Zk[:, i] * Xk[:, j]) / k
                if k == 1:
                    A[:, i, j] = term
                else:
                    A[:, i, j] += term
    return (h / (2.0 * np.pi)) * A


def ikpw_from_normals(dW, h, X, Y):
    """(A, I) of sdeint.Ikpw given the standard normals it would have drawn.
    I = 0.5*(dW dW^T - h*1) + A   (wiener.py:106)."""
    dW = np.asarray(dW, dtype=np.float64)
    Ns, m = dW.shape
    A = levy_kpw(dW, h, X, Y)
    I = np.empty_like(A)
    for i in range(m):
        for j in range(m):
            hd = h if i == j else 0.0
            I[:, i, j] = 0.5 * (dW[:, i] * dW[:, j] - hd) + A[:, i, j]
    return A, I


def jkpw_from_normals(dW, h, X, Y):
    """J = I + 0.5*h*1 (wiener.py:130)."""
    A, I = ikpw_from_normals(dW, h, X, Y)
    m = dW.shape[1]
    return A, I + 0.5 * h * np.eye(m).reshape(1, m, m)


def iwik_from_normals(dW, h, X, Y, Gt):
    """(Atilde, I) of sdeint.Iwik given its normals; collapsed O(m^2) form.

    Follows wiener.py:234-282 term by term but applies Sigma_inf (wiener.py:
    217-226) to the vector G analytically instead of materialising
    (M x m^2)(m^2 x m^2)(m^2 x M):
        u = Ghat dW,  Ghat antisymmetric with Ghat[i,j] = g_(i,j) for i<j
        (Sigma_inf g)_(i,j) = 2 g_(i,j) + (2/h) (dW_j u_i - dW_i u_j)
        sqrtSigma g = (Sigma_inf g + 2 r g) / (sqrt2 (1 + r)),  r = sqrt(1+|dW|^2/h)
        Atilde = Atilde_n + (h/2pi) sqrt(a_tail(n)) * sqrtSigma g
    and the column-major unvec placement (wiener.py:278-280):
        I[j,i] = 0.5 dW_i dW_j + Atilde_(i,j),  I[i,j] = 0.5 dW_i dW_j - Atilde_(i,j), i<j
        I[i,i] = 0.5 (dW_i^2 - h).
    """
    dW = np.asarray(dW, dtype=np.float64)
    Ns, m = dW.shape
    if m == 1:
        w = dW.reshape(Ns, 1, 1)
        return np.zeros((Ns, 1, 1)), (w * w - h) / 2.0
    n = X.shape[0]
    pairs = pair_list(m)
    M = len(pairs)
    c = np.sqrt(2.0 / h)
    # truncated series in K-row order; (K(P-1) kp(Z,X))_(i,j) = X_i Z_j - Z_i X_j
    At = np.zeros((Ns, M))
    for k in range(1, n + 1):
        Xk = X[k - 1]
        Zk = Y[k - 1] + c * dW
        for r, (i, j) in enumerate(pairs):
            term = (Xk[:, i] * Zk[:, j] - Zk[:, i] * Xk[:, j]) / k
            if k == 1:
                At[:, r] = term
            else:
                At[:, r] += term
    At *= h / (2.0 * np.pi)
    # tail correction
    rad = np.sqrt(1.0 + np.sum(dW * dW, axis=1) / h)
    u = np.zeros((Ns, m))
    for r, (i, j) in enumerate(pairs):
        u[:, i] += Gt[:, r] * dW[:, j]
        u[:, j] -= Gt[:, r] * dW[:, i]
    scale = h / (2.0 * np.pi) * a_tail(n) ** 0.5
    Atilde = np.empty((Ns, M))
    for r, (i, j) in enumerate(pairs):
        sig = 2.0 * Gt[:, r] + (2.0 / h) * (dW[:, j] * u[:, i] - dW[:, i] * u[:, j])
        root = (sig + 2.0 * rad * Gt[:, r]) / (np.sqrt(2.0) * (1.0 + rad))
        Atilde[:, r] = At[:, r] + scale * root
    I = np.empty((Ns, m, m))
    for i in range(m):
        I[:, i, i] = 0.5 * (dW[:, i] * dW[:, i] - h)
    for r, (i, j) in enumerate(pairs):
        sym = 0.5 * dW[:, i] * dW[:, j]
        I[:, j, i] = sym + Atilde[:, r]
        I[:, i, j] = sym - Atilde[:, r]
    return Atilde.reshape(Ns, M, 1), I


def jwik_from_normals(dW, h, X, Y, Gt):
    At, I = iwik_from_normals(dW, h, X, Y, Gt)
    m = dW.shape[1]
    return At, I + 0.5 * h * np.eye(m).reshape(1, m, m)


def ikpw(dW, h, n=5, generator=None):
    """Drop-in restatement of sdeint.Ikpw (same draws from `generator`)."""
    generator = np.random.default_rng() if generator is None else generator
    dW2 = _as_2d(dW)
    X, Y = draw_kpw_normals(generator, dW2.shape[0], dW2.shape[1], n)
    return ikpw_from_normals(dW2, h, X, Y)


def jkpw(dW, h, n=5, generator=None):
    generator = np.random.default_rng() if generator is None else generator
    dW2 = _as_2d(dW)
    X, Y = draw_kpw_normals(generator, dW2.shape[0], dW2.shape[1], n)
    return jkpw_from_normals(dW2, h, X, Y)


def iwik(dW, h, n=5, generator=None):
    generator = np.random.default_rng() if generator is None else generator
    dW2 = _as_2d(dW)
    X, Y, Gt = draw_wik_normals(generator, dW2.shape[0], dW2.shape[1], n)
    return iwik_from_normals(dW2, h, X, Y, Gt)


def jwik(dW, h, n=5, generator=None):
    generator = np.random.default_rng() if generator is None else generator
    dW2 = _as_2d(dW)
    X, Y, Gt = draw_wik_normals(generator, dW2.shape[0], dW2.shape[1], n)
    return jwik_from_normals(dW2, h, X, Y, Gt)


def _as_2d(dW):
    """Ikpw/Iwik accept (N,m) or (N,m,1); anything else is a bare ValueError
    (wiener.py:98-101, :255-258)."""
    dW = np.asarray(dW)
    if dW.ndim == 2:
        return dW
    if dW.ndim == 3 and dW.shape[2] == 1:
        return dW[:, :, 0]
    raise ValueError


# --------------------------------------------------------------------------------------
# vec / Kronecker helpers (index conventions, sdeint/wiener.py:136-203)
# --------------------------------------------------------------------------------------

def vec(A):
    """Column-major stack: vec(A)[s, c*m + r, 0] = A[s, r, c] (wiener.py:136-147)."""
    Ns, m, n = A.shape
    out = np.empty((Ns, m * n, 1), dtype=A.dtype)
    for c in range(n):
        for r in range(m):
            out[:, c * m + r, 0] = A[:, r, c]
    return out


def unvec(v, m=None):
    """Inverse of vec (wiener.py:150-155)."""
    Ns, L, _ = v.shape
    if m is None:
        m = int(round(math.sqrt(L)))
    n = L // m
    out = np.empty((Ns, m, n), dtype=v.dtype)
    for c in range(n):
        for r in range(m):
            out[:, r, c] = v[:, c * m + r, 0]
    return out


def kp(a, b):
    """kp(a,b)[s, i*m + j, 0] = a[s,i,0] * b[s,j,0] (wiener.py:158-167)."""
    if a.shape != b.shape or a.shape[-1] != 1:
        raise ValueError
    Ns, m, _ = a.shape
    out = np.empty((Ns, m * m, 1))
    for i in range(m):
        for j in range(m):
            out[:, i * m + j, 0] = a[:, i, 0] * b[:, j, 0]
    return out


def kp2(A, B):
    """Batched Kronecker product of matrices (wiener.py:170-179)."""
    Ns = A.shape[0]
    if B.shape[0] != Ns:
        raise ValueError
    _, p, q = A.shape
    _, r, s = B.shape
    out = np.empty((Ns, p * r, q * s))
    for i in range(p):
        for j in range(q):
            out[:, i * r:(i + 1) * r, j * s:(j + 1) * s] = A[:, i:i + 1, j:j + 1] * B
    return out


def perm_P(m):
    """Commutation matrix: P kp(a,b) = kp(b,a) (wiener.py:182-190)."""
    P = np.zeros((m * m, m * m), dtype=np.int64)
    for i in range(m):
        for j in range(m):
            P[i * m + j, j * m + i] = 1
    return P


def sel_K(m):
    """Selector of the strict upper triangle: row r <-> pair_list(m)[r]=(i,j),
    a single 1 at column i*m + j (wiener.py:193-203)."""
    pairs = pair_list(m)
    K = np.zeros((len(pairs), m * m), dtype=np.int64)
    for r, (i, j) in enumerate(pairs):
        K[r, i * m + j] = 1
    return K
