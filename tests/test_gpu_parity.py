"""GPU parity tests (run on the B200 box: `pytest -m gpu`).  Everything here goes through
the public shim -> ctypes -> C ABI (include/sdb.h) -> CUDA kernels, and is compared with
(a) the committed outputs of the unmodified reference (tests/golden) and (b) the pinned
oracle on seeded inputs.  Tolerance: 1e-12 relative (fp64), the north-star bound; the
KP2iS scheme, whose implicit solve the reference delegates to MINPACK hybrd (fixtures generated at
the tightest xtol it completes with, 1e-12 / 1e-11; its termination leaves a few 1e-12 relative
against an exact solve), gets 5e-12 against the reference and 1e-11 against the oracle's Newton solve."""
import os

import numpy as np
import pytest

from conftest import ROOT, SYSTEMS, load_golden, rel_err
from oracle import philox_oracle as po
from oracle import sde_oracle as orc

pytestmark = pytest.mark.gpu

TOL = 1e-12


@pytest.fixture(scope="module")
def sdb():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import sdeint_b200
    sdeint_b200._lib.load()
    return sdeint_b200


# ---- integrators on host-supplied noise vs the reference's golden outputs ------------------

@pytest.mark.parametrize("name", sorted(SYSTEMS))
def test_integrators_vs_reference_golden(sdb, name):
    g = load_golden("integrate_%s.npz" % name)
    f_ito, G, _ = SYSTEMS[name](False)
    f_str, _, _ = SYSTEMS[name](True)
    tspan, dW, I, J, y0 = g["tspan"], g["dW"], g["I"], g["J"], g["y0"]
    before = sdb.launch_count()
    # ensemble form: all paths in one launch, (P, N, d)
    assert rel_err(sdb.itoEuler(f_ito, G, y0, tspan, dW=dW), g["y_itoEuler"]) <= TOL
    assert rel_err(sdb.stratHeun(f_str, G, y0, tspan, dW=dW), g["y_stratHeun"]) <= TOL
    assert rel_err(sdb.itoSRI2(f_ito, G, y0, tspan, dW=dW, I=I), g["y_itoSRI2"]) <= TOL
    assert rel_err(sdb.stratSRS2(f_str, G, y0, tspan, dW=dW, J=J), g["y_stratSRS2"]) <= TOL
    if "y_itoSRI2_listG" in g:
        y = sdb.itoSRI2(f_ito, G.columns(), y0, tspan, dW=dW, I=I)
        assert rel_err(y, g["y_itoSRI2_listG"]) <= TOL
    if "y_stratKP2iS" in g:
        y = sdb.stratKP2iS(f_str, G, y0, tspan, dW=dW, J=J, rtol=1e-12)
        assert rel_err(y, g["y_stratKP2iS"]) <= 5e-12
    assert sdb.launch_count() > before          # kernels really launched
    # single-path form: the reference's exact call and return layout (N, d)
    y1 = sdb.itoSRI2(f_ito, G, y0, tspan, dW=dW[0], I=I[0])
    assert y1.shape == g["y_itoSRI2"][0].shape and y1.dtype == np.float64
    assert np.array_equal(y1[0], y0)
    assert rel_err(y1, g["y_itoSRI2"][0]) <= TOL


def test_scalar_y0_wrapping(sdb):
    """Scalar y0 -> output (N, 1) (integrate.py:54-60); int -> float64."""
    g = load_golden("integrate_kp4446.npz")
    f, G, y0 = SYSTEMS["kp4446"](False)
    assert np.ndim(y0) == 0
    y = sdb.itoSRI2(f, G, y0, g["tspan"], dW=g["dW"][0], I=g["I"][0])
    assert y.shape == (len(g["tspan"]), 1)
    assert rel_err(y, g["y_itoSRI2"][0]) <= TOL
    f2, G2, _ = SYSTEMS["ou"](False)
    y = sdb.itoEuler(f2, G2, 1, np.linspace(0, 1, 11), dW=np.zeros((10, 1)))
    assert y.dtype == np.float64 and y[0, 0] == 1.0


# ---- repeated integrals: construction parity with host-supplied normals --------------------

def test_repeated_integrals_vs_reference_golden(sdb):
    g = load_golden("wiener.npz")
    for ci, (m, n, N, h) in enumerate(g["cases"]):
        m, n, N = int(m), int(n), int(N)
        pre = "c%d_" % ci
        dW, X, Y, Gt = g[pre + "dW"], g[pre + "X"], g[pre + "Y"], g[pre + "Gt"]
        scale = np.max(np.abs(g[pre + "I_kpw"]))
        A, I = sdb.Ikpw(dW, h, n=n, normals=(X, Y))
        assert A.shape == g[pre + "A_kpw"].shape and I.shape == g[pre + "I_kpw"].shape
        assert np.max(np.abs(A - g[pre + "A_kpw"])) <= TOL * scale
        assert np.max(np.abs(I - g[pre + "I_kpw"])) <= TOL * scale
        _, Jk = sdb.Jkpw(dW, h, n=n, normals=(X, Y))
        assert np.max(np.abs(Jk - g[pre + "J_kpw"])) <= TOL * scale
        At, Iw = sdb.Iwik(dW, h, n=n, normals=(X, Y, Gt))
        assert At.shape == g[pre + "At_wik"].shape and Iw.shape == g[pre + "I_wik"].shape
        assert np.max(np.abs(At - g[pre + "At_wik"])) <= TOL * scale
        assert np.max(np.abs(Iw - g[pre + "I_wik"])) <= TOL * scale
        _, Jw = sdb.Jwik(dW.reshape(N, m, 1), h, n=n, normals=(X, Y, Gt))   # (N, m, 1) accepted
        assert np.max(np.abs(Jw - g[pre + "J_wik"])) <= TOL * scale


def test_repeated_integrals_bad_shape(sdb):
    with pytest.raises(ValueError):
        sdb.Ikpw(np.zeros((4,)), 0.1)
    with pytest.raises(ValueError):
        sdb.Iwik(np.zeros((2, 3, 4, 5)), 0.1)


def test_identities_on_device_noise(sdb):
    """Wiktorsson (2.1) identities (test_wiener.py:94-130) on fully on-device output."""
    N, h, m = 4000, 0.002, 8
    dW = sdb.deltaW(N, m, h, seed=5)
    assert dW.shape == (N, m)
    outer = dW[:, :, None] * dW[:, None, :]
    eye = np.eye(m)[None]
    A, I = sdb.Ikpw(dW, h, seed=5)
    assert A.shape == (N, m, m) and I.shape == (N, m, m)
    assert np.allclose(I + I.transpose(0, 2, 1), outer - h * eye, atol=1e-15)
    assert np.allclose(A, -A.transpose(0, 2, 1), atol=0)
    assert np.allclose(2.0 * (I - A), outer - h * eye, atol=1e-15)
    _, J = sdb.Jkpw(dW, h, seed=5)
    assert np.allclose(J + J.transpose(0, 2, 1), outer, atol=1e-15)
    At, Iw = sdb.Iwik(dW, h, seed=5)
    M = m * (m - 1) // 2
    assert At.shape == (N, M, 1)
    assert np.allclose(Iw + Iw.transpose(0, 2, 1), outer - h * eye, atol=1e-15)
    K0, P0 = orc.sel_K(m), orc.perm_P(m)
    Afull = orc.unvec(np.einsum("ab,sbc->sac", (np.eye(m * m) - P0).dot(K0.T), At))
    assert np.allclose(2.0 * (Iw - Afull), outer - h * eye, atol=1e-15)
    # Levy area statistics: E A_ij = 0, Var A_ij (n=5 series) = h^2/(2 pi^2) sum_{k<=5} 1/k^2 * ...
    a01 = A[:, 0, 1]
    assert abs(a01.mean()) < 5 * a01.std() / np.sqrt(N)


# ---- the on-device generator vs its numpy restatement ---------------------------------------

def test_deltaW_matches_philox_oracle(sdb):
    steps, m, h, P, seed = 37, 5, 0.01, 19, 987654321
    dW = sdb.deltaW(steps, m, h, paths=P, seed=seed, path_id0=1000)
    want = po.delta_w(steps, m, h, P, seed, path_id0=1000)
    assert dW.shape == want.shape
    assert np.max(np.abs(dW - want)) <= 1e-14
    # rank invariance: a shard with a path offset reproduces the same paths
    part = sdb.deltaW(steps, m, h, paths=7, seed=seed, path_id0=1005)
    assert np.array_equal(part, dW[5:12])


def test_device_Ikpw_matches_oracle_on_philox_normals(sdb):
    steps, m, h, P, seed, n = 11, 6, 0.004, 5, 4242, 5
    dW = sdb.deltaW(steps, m, h, paths=P, seed=seed)
    X, Y = po.kpw_normals(steps, m, n, P, seed)
    Gt = po.wik_tail_normals(steps, m, n, P, seed)
    A, I = sdb.Ikpw(dW, h, n=n, seed=seed)
    At, Jw = sdb.Jwik(dW, h, n=n, seed=seed)
    for p in range(P):
        A_o, I_o = orc.ikpw_from_normals(dW[p], h, X[:, p], Y[:, p])
        At_o, Jw_o = orc.jwik_from_normals(dW[p], h, X[:, p], Y[:, p], Gt[p])
        s = np.max(np.abs(I_o))
        assert np.max(np.abs(A[p] - A_o)) <= 1e-12 * s
        assert np.max(np.abs(I[p] - I_o)) <= 1e-12 * s
        assert np.max(np.abs(At[p] - At_o)) <= 1e-12 * s
        assert np.max(np.abs(Jw[p] - Jw_o)) <= 1e-12 * s


def test_normals_distribution(sdb):
    """Moments and a KS test of the device normals (2.4e6 draws)."""
    from scipy import stats
    z = sdb.deltaW(300, 8, 1.0, paths=1000, seed=77).reshape(-1)
    n = z.size
    assert abs(z.mean()) < 5 / np.sqrt(n)
    assert abs(z.var() - 1.0) < 5 * np.sqrt(2.0 / n)
    assert abs((z ** 3).mean()) < 5 * np.sqrt(15.0 / n)
    assert abs((z ** 4).mean() - 3.0) < 5 * np.sqrt(96.0 / n)
    assert stats.kstest(z[:200000], "norm").pvalue > 1e-3
    # independence across paths and across steps
    zz = z.reshape(1000, 300, 8)
    assert abs(np.corrcoef(zz[:-1, :, 0].ravel(), zz[1:, :, 0].ravel())[0, 1]) < 0.01
    assert abs(np.corrcoef(zz[:, :-1, 0].ravel(), zz[:, 1:, 0].ravel())[0, 1]) < 0.01
    assert abs(np.corrcoef(zz[:, :, 0].ravel(), zz[:, :, 1].ravel())[0, 1]) < 0.01


# ---- fused on-device noise path vs oracle on the noise the device generated ----------------

@pytest.mark.parametrize("name,strat,method", [
    ("lin10", False, "kpw"), ("lin10", True, "kpw"), ("lin10", False, "wik"),
    ("r74", False, "kpw"), ("r74", True, "wik"), ("kp4459", False, "kpw"),
    ("kp4446", False, "wik"), ("lin2", False, "kpw"),
])
def test_fused_device_noise_matches_oracle(sdb, name, strat, method):
    f, G, y0 = SYSTEMS[name](strat)
    y0v = np.atleast_1d(np.asarray(y0, dtype=float))
    m = G.m
    tspan = np.linspace(0.0, 0.2, 41)
    N = len(tspan)
    h = (tspan[-1] - tspan[0]) / (N - 1)
    P, seed = 70, 31337
    integ = sdb.stratSRS2 if strat else sdb.itoSRI2
    rep = {(False, "kpw"): sdb.Ikpw, (False, "wik"): sdb.Iwik,
           (True, "kpw"): sdb.Jkpw, (True, "wik"): sdb.Jwik}[(strat, method)]
    y = integ(f, G, y0, tspan, rep, paths=P, seed=seed, path_id0=3)
    assert y.shape == (P, N, len(y0v))
    dW = sdb.deltaW(N - 1, m, h, paths=P, seed=seed, path_id0=3)
    _, IJ = rep(dW, h, seed=seed, path_id0=3, ensemble=True)
    for p in (0, 1, 33, P - 1):
        ref = orc.srk2(f, G, y0v, tspan, dW[p], IJ[p])
        assert rel_err(y[p], ref) <= TOL
    # decimated storage returns the same states
    yd = integ(f, G, y0, tspan, rep, paths=P, seed=seed, path_id0=3, save_every=10)
    assert yd.shape == (P, 5, len(y0v))
    assert np.array_equal(yd, y[:, ::10])
    # unfused pipeline through the host-noise kernel agrees too
    y2 = integ(f, G, y0, tspan, rep, dW=dW, **({"J": IJ} if strat else {"I": IJ}))
    assert rel_err(y2, y) <= TOL


@pytest.mark.parametrize("d,m,strat", [(10, 10, False), (6, 4, True), (7, 5, False)])
def test_fused_user_drift_linear_diffusion(sdb, monkeypatch, d, m, strat):
    """Nonlinear drift given as source + linear diffusion g_k = B_k y: the fused kernel with the drift
    evaluated per lane (`jit:fusedjit|ud`), against the oracle on the noise the device generated and
    against the generic loop kernel; time-dependent drift, decimated storage, resumed runs."""
    _, G, y0 = sdb.ops.sys_lin(d, m)
    exprs = ["-p[0]*y[%d] + p[1]*y[%d] - p[2]*y[%d]*y[%d]*y[%d] + 0.1*sin(3*t)" % (i, (i + 1) % d, i, i, i)
             for i in range(d)]
    f = sdb.ops.ExprDrift(exprs, params=[1.5, 0.3, 0.2])
    tspan = np.linspace(0.0, 0.25, 51)
    N = len(tspan)
    h = (tspan[-1] - tspan[0]) / (N - 1)
    P, seed = 70, 4242
    integ = sdb.stratSRS2 if strat else sdb.itoSRI2
    rep = sdb.Jkpw if strat else sdb.Ikpw
    y = integ(f, G, y0, tspan, rep, paths=P, seed=seed, path_id0=5, commutative=False)
    assert sdb.last_kernel() == "jit:fusedjit|ud"
    dW = sdb.deltaW(N - 1, m, h, paths=P, seed=seed, path_id0=5)
    _, IJ = rep(dW, h, seed=seed, path_id0=5, ensemble=True)
    for p in (0, 31, 32, P - 1):
        ref = orc.srk2(f, G, np.asarray(y0, float), tspan, dW[p], IJ[p])
        assert rel_err(y[p], ref) <= TOL
    yd = integ(f, G, y0, tspan, rep, paths=P, seed=seed, path_id0=5, save_every=10, commutative=False)
    assert np.array_equal(yd, y[:, ::10])
    # the generic loop kernel on the same noise
    monkeypatch.setenv("SDB_DISABLE_FUSED", "1")
    yg = integ(f, G, y0, tspan, rep, paths=P, seed=seed, path_id0=5, commutative=False)
    assert sdb.last_kernel().startswith("jit:loop")
    monkeypatch.delenv("SDB_DISABLE_FUSED")
    assert rel_err(yg, y) <= TOL
    # in-loop observers ride in the same kernel
    yo, obs = integ(f, G, y0, tspan, rep, paths=P, seed=seed, path_id0=5, commutative=False,
                    observe=[("max", 0)])
    assert sdb.last_kernel() == "jit:fusedjit|ud|obs"
    assert rel_err(yo, y) <= TOL and obs.shape == (P, 1)
    assert np.array_equal(obs[:, 0], yo[:, :, 0].max(axis=1))


def test_reference_style_python_callables(sdb):
    """The reference's own calling convention: plain Python f(y, t), G(y, t) (README.rst:60-97 and the systems
    of sdeint/tests/test_integrate.py), traced into device operators; every integrator against the oracle
    running the ORIGINAL Python functions on the same host-supplied noise."""
    rng = np.random.default_rng(21)
    # README scalar equation (KP 4.46) through itoint, exactly as the reference's README writes it
    a, b = 1.0, 0.8
    tspan = np.linspace(0.0, 5.0, 501)
    f = lambda x, t: -(a + x * b ** 2) * (1 - x ** 2)
    g = lambda x, t: b * (1 - x ** 2)
    h = 0.01
    dW = rng.normal(0.0, np.sqrt(h), (500, 1))
    y = sdb.itoEuler(f, g, 0.1, tspan, dW=dW)
    assert y.shape == (501,) or y.shape == (501, 1)
    assert rel_err(np.asarray(y).reshape(501, 1), orc.euler(f, g, np.array([0.1]), tspan, dW).reshape(501, 1)) <= TOL
    y = sdb.itoint(f, g, 0.1, tspan, generator=np.random.default_rng(1))
    assert np.all(np.isfinite(y)) and np.asarray(y).reshape(-1).shape == (501,)
    # README 2-d linear system with constant B (additive: no repeated integrals needed)
    A = np.array([[-0.5, -2.0], [2.0, -1.0]])
    B = np.diag([0.5, 0.5])
    fl = lambda x, t: A.dot(x)
    Gl = lambda x, t: B
    ts = np.linspace(0.0, 1.0, 101)
    dW = rng.normal(0.0, 0.1, (100, 2))
    x0 = np.array([3.0, 3.0])
    assert rel_err(sdb.itoEuler(fl, Gl, x0, ts, dW=dW), orc.euler(fl, Gl, x0, ts, dW)) <= TOL
    assert rel_err(sdb.stratHeun(fl, Gl, x0, ts, dW=dW), orc.heun(fl, Gl, x0, ts, dW)) <= TOL
    _, I = sdb.Ikpw(dW, 0.01, generator=np.random.default_rng(5))
    assert rel_err(sdb.itoSRI2(fl, Gl, x0, ts, dW=dW, I=I), orc.srk2(fl, Gl, x0, ts, dW, I)) <= TOL
    # Roessler (7.4)-type system written with np.dot, G as one function and as a list of column functions
    d, m = 5, 4
    A5 = -np.eye(d) + 0.1 * rng.standard_normal((d, d))
    Bm = 0.2 * rng.standard_normal((m, d, d))
    f5 = lambda y, t: np.dot(A5, y)
    G5 = lambda y, t: np.dot(Bm, y).T
    G5_cols = [(lambda y, t, k=k: np.dot(Bm[k], y)) for k in range(m)]
    y0 = np.ones(d)
    dW = rng.normal(0.0, 0.1, (100, m))
    _, J = sdb.Jkpw(dW, 0.01, generator=np.random.default_rng(6))
    ref = orc.srk2(f5, G5, y0, ts, dW, J)
    assert rel_err(sdb.stratSRS2(f5, G5, y0, ts, dW=dW, J=J), ref) <= TOL
    assert rel_err(sdb.stratSRS2(f5, G5_cols, y0, ts, dW=dW, J=J), ref) <= TOL
    ye = sdb.stratSRS2(f5, G5, y0, ts, paths=64, seed=9)                   # ensemble on device noise
    assert ye.shape == (64, 101, d) and sdb.last_kernel().startswith(("sdbk_fused", "jit:fused"))
    # a nonlinear, time-dependent system: every scheme
    fn = lambda y, t: np.array([-y[0] / (1.0 + y[1] ** 2) + 0.2 * np.cos(3 * t), -0.5 * y[1] + 0.1 * np.sin(y[0])])
    Gn = lambda y, t: np.array([[0.3 * (1.0 + 0.5 * np.sin(y[1])), 0.1], [0.05, 0.2 * np.cos(y[0]) * np.exp(-t)]])
    z0 = np.array([0.5, -0.2])
    dW = rng.normal(0.0, 0.1, (100, 2))
    _, I = sdb.Iwik(dW, 0.01, generator=np.random.default_rng(7))
    assert rel_err(sdb.itoEuler(fn, Gn, z0, ts, dW=dW), orc.euler(fn, Gn, z0, ts, dW)) <= TOL
    assert rel_err(sdb.stratHeun(fn, Gn, z0, ts, dW=dW), orc.heun(fn, Gn, z0, ts, dW)) <= TOL
    assert rel_err(sdb.itoSRI2(fn, Gn, z0, ts, dW=dW, I=I), orc.srk2(fn, Gn, z0, ts, dW, I)) <= TOL
    # piecewise formulas: comparisons as masks, np.where, np.sign, np.maximum (device prelude lt/le/gt/ge)
    fp = lambda y, t: np.array([np.where(y > 0.3, -2.0 * y, 0.5 - y)[0] + 0.2 * np.heaviside(t - 0.5, 0.5),
                                -np.sign(y[1]) * np.sqrt(np.maximum(abs(y[1]), 1e-3))])
    Gp = lambda y, t: np.array([[0.3 * (y[0] > 0) * y[0], 0.1], [0.05 * (y[1] <= 0.1), 0.2]])
    assert rel_err(sdb.stratHeun(fp, Gp, z0, ts, dW=dW), orc.heun(fp, Gp, z0, ts, dW)) <= TOL
    assert rel_err(sdb.itoSRI2(fp, Gp, z0, ts, dW=dW, I=I), orc.srk2(fp, Gp, z0, ts, dW, I)) <= TOL
    fe = lambda y, t: np.array([np.arctan(y[1]) - np.log1p(y[0] * y[0]) + np.expm1(-t),
                                np.arctan2(y[0], 1.0 + y[1] * y[1]) - np.cbrt(y[1]) + np.hypot(y[0], 0.5) - 0.5])
    assert rel_err(sdb.stratHeun(fe, Gn, z0, ts, dW=dW), orc.heun(fe, Gn, z0, ts, dW)) <= TOL
    # one independent multiplicative noise per component is recognised as diagonal: no Levy areas are drawn,
    # and the paths equal those of the full construction on the same seed
    sig = np.array([0.2, 0.3, 0.1])
    fd = lambda y, t: -y + 0.1 * np.sin(y)
    Gd = lambda y, t: np.diag(sig * y * (1 + 0.1 * np.cos(t)))
    yd = sdb.itoSRI2(fd, Gd, np.ones(3), ts, paths=50, seed=4)
    from sdeint_b200 import integrate as _integ
    _, _, fo, Go, *_ = _integ._check_args(fd, Gd, np.ones(3), ts)
    assert Go.kind == sdb.ops.DIFF_USER_DIAG
    yfull = sdb.itoSRI2(fo, sdb.ops.ExprDiffusion(Go.exprs, Go.params, diagonal=False), np.ones(3), ts, paths=50, seed=4)
    assert rel_err(yd, yfull) <= TOL
    # what cannot be traced is refused, like any other bad argument
    with pytest.raises(sdb.SDEValueError, match="could not be turned into a device function"):
        sdb.itoint(lambda y, t: -y if y[0] > 0 else y, Gn, z0, ts)


@pytest.mark.parametrize("d,m,linear_drift", [(10, 10, True), (10, 10, False), (6, 4, False)])
def test_affine_diffusion_fused(sdb, monkeypatch, d, m, linear_drift):
    """g_k = B_k y + c_k (multiplicative + additive noise): the fused kernel with the constants added to G_n
    (`jit:fusedjit|aff`, `jit:fusedjit|ud|aff`) against the oracle on replayed device noise and the generic
    kernel; an affine ExprDiffusion is promoted to it."""
    fL, GL, y0 = sdb.ops.sys_lin(d, m)
    rng = np.random.default_rng(8)
    Cc = 0.05 * rng.standard_normal((d, m))
    G = sdb.ops.AffineDiffusion(GL.B, Cc)
    f = fL if linear_drift else sdb.ops.ExprDrift(
        ["-1.2*y[%d] + 0.3*y[%d] - 0.2*y[%d]*y[%d]*y[%d]" % (i, (i + 1) % d, i, i, i) for i in range(d)])
    tspan = np.linspace(0.0, 0.2, 41)
    h = 0.2 / 40
    P, seed = 50, 321
    for strat, integ, rep in ((False, sdb.itoSRI2, sdb.Ikpw), (True, sdb.stratSRS2, sdb.Jkpw)):
        y = integ(f, G, y0, tspan, rep, paths=P, seed=seed)
        assert sdb.last_kernel() == ("jit:fusedjit|aff" if linear_drift else "jit:fusedjit|ud|aff")
        dW = sdb.deltaW(40, m, h, paths=P, seed=seed)
        _, IJ = rep(dW, h, seed=seed, ensemble=True)
        for p in (0, 31, 32, P - 1):
            assert rel_err(y[p], orc.srk2(f, G, np.asarray(y0, float), tspan, dW[p], IJ[p])) <= TOL
    monkeypatch.setenv("SDB_DISABLE_FUSED", "1")
    yg = sdb.stratSRS2(f, G, y0, tspan, sdb.Jkpw, paths=P, seed=seed)
    monkeypatch.delenv("SDB_DISABLE_FUSED")
    assert rel_err(yg, y) <= TOL
    # host-supplied noise and the other schemes go through the generic kernels with the same operator
    assert rel_err(sdb.stratHeun(f, G, y0, tspan, dW=dW[3]), orc.heun(f, G, np.asarray(y0, float), tspan, dW[3])) <= TOL
    if d * m >= 16 and not linear_drift:
        GE = sdb.ops.ExprDiffusion([["+".join("%r*y[%d]" % (float(GL.B[k, i, l]), l) for l in range(d)) +
                                     "+%r" % float(Cc[i, k]) for k in range(m)] for i in range(d)])
        ye = sdb.stratSRS2(f, GE, y0, tspan, sdb.Jkpw, paths=P, seed=seed)
        assert sdb.last_kernel() == "jit:fusedjit|ud|aff" and rel_err(ye, y) <= 1e-13


def test_nvrtc_results_are_cached_on_disk(tmp_path):
    """A second process integrating the same user system loads its kernels from $SDB_CACHE_DIR instead of
    compiling them again (same results, bit for bit); SDB_JIT_CACHE=0 compiles."""
    import subprocess, sys, json
    code = (
        "import sys, json, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "import sdeint_b200 as sdb\n"
        "f = sdb.ops.ExprDrift(['-y[0] + 0.2*y[1]*y[1]', '-0.5*y[1] + 0.1*sin(y[0])'])\n"
        "G = sdb.ops.ExprDiffusion([['0.3*(1.0+0.5*sin(y[1]))', '0.1'], ['0.05', '0.2*cos(y[0])']])\n"
        "y = sdb.itoSRI2(f, G, np.array([0.5, -0.2]), np.linspace(0, 0.1, 11), paths=8, seed=3)\n"
        "h, m = sdb.jit_cache_stats()\n"
        "print(json.dumps({'hits': h, 'misses': m, 'sum': float(y.sum()), 'kernel': sdb.last_kernel()}))\n"
    ) % ROOT
    env = dict(os.environ, SDB_CACHE_DIR=str(tmp_path / "jit"))
    env.pop("SDB_JIT_CACHE", None)
    runs = []
    for extra in ({}, {}, {"SDB_JIT_CACHE": "0"}):
        out = subprocess.run([sys.executable, "-c", code], env=dict(env, **extra), capture_output=True,
                             text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        runs.append(json.loads(out.stdout.strip().splitlines()[-1]))
    first, second, off = runs
    assert first["kernel"].startswith("jit:") and first["misses"] >= 1 and first["hits"] == 0
    assert second["misses"] == 0 and second["hits"] == first["misses"]
    assert off["hits"] == 0 and off["misses"] == first["misses"]
    assert first["sum"] == second["sum"] == off["sum"]
    assert len(list((tmp_path / "jit").glob("*.cubin"))) == first["misses"]


@pytest.mark.parametrize("name", ["heston", "piecewise", "oscillator", "linear_np_dot"])
def test_python_callables_match_the_reference(sdb, name):
    """The SAME Python functions that oracle/make_golden.py handed to the unmodified reference
    (tests/callable_systems.py -> tests/golden/callables.npz), traced into device code here: every scheme
    reproduces the reference's paths to 1e-12, single paths and batched."""
    import callable_systems as cs
    g = load_golden("callables.npz")
    made = cs.SYSTEMS[name]()
    f, G, y0 = made[:3]
    ts, dW, I, J = (g["%s_%s" % (name, k)] for k in ("tspan", "dW", "I", "J"))
    for p in range(dW.shape[0]):
        assert rel_err(sdb.itoEuler(f, G, y0, ts, dW=dW[p]), g[name + "_y_itoEuler"][p]) <= TOL
        assert rel_err(sdb.stratHeun(f, G, y0, ts, dW=dW[p]), g[name + "_y_stratHeun"][p]) <= TOL
        assert rel_err(sdb.itoSRI2(f, G, y0, ts, dW=dW[p], I=I[p]), g[name + "_y_itoSRI2"][p]) <= TOL
        assert rel_err(sdb.stratSRS2(f, G, y0, ts, dW=dW[p], J=J[p]), g[name + "_y_stratSRS2"][p]) <= TOL
    assert rel_err(sdb.itoSRI2(f, G, y0, ts, dW=dW, I=I), g[name + "_y_itoSRI2"]) <= TOL      # batched over paths
    if len(made) > 3:
        assert rel_err(sdb.itoSRI2(f, made[3], y0, ts, dW=dW, I=I), g[name + "_y_itoSRI2_listG"]) <= TOL


def test_user_drift_long_run_parity(sdb):
    """SURVEY 8(d)-style long run for the kernels with the drift evaluated per lane: 64 paths x 1000 steps at
    d = m = 10, SRI2 (Ikpw: 8-warp kernel; Iwik: TMEM kernel) and Heun, device noise replayed through the oracle."""
    d = m = 10
    _, G, y0 = sdb.ops.sys_lin(d, m)
    f = sdb.ops.ExprDrift(["-p[0]*y[%d] + p[1]*y[%d] - p[2]*y[%d]*y[%d]*y[%d] + 0.05*cos(2*t)"
                           % (i, (i + 1) % d, i, i, i) for i in range(d)], params=[1.5, 0.3, 0.2])
    tspan = np.linspace(0.0, 1.0, 1001)
    P, seed, h = 64, 2026, 1e-3
    dW = sdb.deltaW(1000, m, h, paths=P, seed=seed)
    for rep, key in ((sdb.Ikpw, "jit:fusedjit|ud"), (sdb.Iwik, "jit:fusedxjit|wik|ud")):
        y = sdb.itoSRI2(f, G, y0, tspan, rep, paths=P, seed=seed, save_every=100, commutative=False)
        assert sdb.last_kernel() == key
        _, I = rep(dW, h, seed=seed, ensemble=True)
        for p in (0, 63):
            ref = orc.srk2(f, G, np.asarray(y0, float), tspan, dW[p], I[p])[::100]
            assert rel_err(y[p], ref) <= TOL
    yh = sdb.stratHeun(f, G, y0, tspan, paths=P, seed=seed, save_every=100)
    assert sdb.last_kernel() == "jit:fusedehjit|heun|ud"
    for p in (0, 63):
        assert rel_err(yh[p], orc.heun(f, G, np.asarray(y0, float), tspan, dW[p])[::100]) <= TOL


def test_linear_expressions_run_the_linear_kernels(sdb):
    """An SDE written as expressions that happen to be linear gets the kernels of the linear operators
    (ops.as_linear): bit-identical to LinearDrift / LinearDiffusion; with a nonlinear drift, the fused kernel
    that evaluates the drift per lane."""
    d = m = 10
    fL, GL, y0 = sdb.ops.sys_lin(d, m)
    A, B = np.asarray(fL.A), np.asarray(GL.B)
    GE = sdb.ops.ExprDiffusion([["+".join("p[%d]*y[%d]" % ((k * d + i) * d + l, l) for l in range(d))
                                 for k in range(m)] for i in range(d)], B.ravel())
    fE = sdb.ops.ExprDrift(["+".join("p[%d]*y[%d]" % (i * d + l, l) for l in range(d)) for i in range(d)], A.ravel())
    ts = np.linspace(0.0, 0.05, 26)
    kw = dict(paths=40, seed=11)
    y_lin = sdb.itoSRI2(fL, GL, y0, ts, **kw)
    y_exp = sdb.itoSRI2(fE, GE, y0, ts, **kw)
    assert sdb.last_kernel() == "sdbk_fused_srk2_d10_m10_v0"
    assert np.array_equal(y_exp, y_lin)
    fN = sdb.ops.ExprDrift(["-1.5*y[%d] - 0.2*y[%d]*y[%d]*y[%d]" % (i, i, i, i) for i in range(d)])
    y_n = sdb.itoSRI2(fN, GE, y0, ts, **kw)
    assert sdb.last_kernel() == "jit:fusedjit|ud"
    y_n2 = sdb.itoSRI2(fN, GL, y0, ts, **kw)
    assert np.array_equal(y_n, y_n2)


@pytest.mark.parametrize("d,m,wik", [(10, 10, True), (16, 12, False), (8, 6, True), (20, 20, True), (12, 16, False)])
def test_fusedx_user_drift(sdb, monkeypatch, d, m, wik):
    """Nonlinear drift + linear diffusion through the TMEM-resident kernel (Iwik/Jwik, or shapes beyond the
    8-warp kernel): `jit:fusedxjit|...|ud` against the oracle on replayed device noise and the generic kernel."""
    _, G, y0 = sdb.ops.sys_lin(d, m)
    f = sdb.ops.ExprDrift(["-p[0]*y[%d] + p[1]*y[%d] - p[2]*y[%d]*y[%d]*y[%d] + 0.05*sin(5*t)"
                           % (i, (i + 2) % d, i, i, i) for i in range(d)], params=[1.4, 0.2, 0.1])
    tspan = np.linspace(0.0, 0.1, 21)
    N = len(tspan)
    h = (tspan[-1] - tspan[0]) / (N - 1)
    P, seed = 40, 909
    rep = sdb.Iwik if wik else sdb.Ikpw
    y = sdb.itoSRI2(f, G, y0, tspan, rep, paths=P, seed=seed, commutative=False)
    x2 = "|x2" if m >= 16 else ""        # two lanes per path where a path's areas fill its TMEM lane
    assert sdb.last_kernel() == "jit:fusedxjit|%s%s|ud" % ("wik" if wik else "kpw", x2)
    dW = sdb.deltaW(N - 1, m, h, paths=P, seed=seed)
    _, I = rep(dW, h, seed=seed, ensemble=True)
    for p in (0, 33, P - 1):
        assert rel_err(y[p], orc.srk2(f, G, np.asarray(y0, float), tspan, dW[p], I[p])) <= TOL
    monkeypatch.setenv("SDB_DISABLE_FUSED", "1")
    yg = sdb.itoSRI2(f, G, y0, tspan, rep, paths=P, seed=seed, commutative=False)
    assert sdb.last_kernel().startswith("jit:loop")
    monkeypatch.delenv("SDB_DISABLE_FUSED")
    assert rel_err(yg, y) <= TOL


@pytest.mark.parametrize("scheme", ["euler", "heun"])
def test_fused_euler_heun_user_drift(sdb, monkeypatch, scheme):
    """itoEuler / stratHeun with a nonlinear, time-dependent drift from source and linear diffusion: the fused
    kernels with the drift evaluated per lane (`jit:fusedehjit|...|ud`) against the oracle on the device's
    increments and against the generic loop kernel."""
    d, m = 10, 10
    _, G, y0 = sdb.ops.sys_lin(d, m)
    f = sdb.ops.ExprDrift(["-p[0]*y[%d] + p[1]*y[%d] - p[2]*y[%d]*y[%d]*y[%d] + 0.1*cos(2*t)"
                           % (i, (i + 3) % d, i, i, i) for i in range(d)], params=[1.2, 0.25, 0.15])
    tspan = np.linspace(0.0, 0.3, 61)
    h = 0.3 / 60
    P, seed = 45, 77
    integ, ref = (sdb.itoEuler, orc.euler) if scheme == "euler" else (sdb.stratHeun, orc.heun)
    y = integ(f, G, y0, tspan, paths=P, seed=seed)
    assert sdb.last_kernel() == "jit:fusedehjit|%s|ud" % scheme
    dW = sdb.deltaW(60, m, h, paths=P, seed=seed)
    for p in (0, 31, 32, P - 1):
        assert rel_err(y[p], ref(f, G, np.asarray(y0, float), tspan, dW[p])) <= TOL
    monkeypatch.setenv("SDB_DISABLE_FUSED", "1")
    yg = integ(f, G, y0, tspan, paths=P, seed=seed)
    assert sdb.last_kernel().startswith("jit:loop")
    monkeypatch.delenv("SDB_DISABLE_FUSED")
    assert rel_err(yg, y) <= TOL


def test_euler_heun_device_noise(sdb):
    f, G, y0 = SYSTEMS["lin2"](False)
    tspan = np.linspace(0.0, 1.0, 101)
    h = 0.01
    P, seed = 40, 99
    dW = sdb.deltaW(100, 2, h, paths=P, seed=seed)
    ye = sdb.itoEuler(f, G, y0, tspan, paths=P, seed=seed)
    yh = sdb.stratHeun(f, G, y0, tspan, paths=P, seed=seed)
    for p in (0, 17, P - 1):
        assert rel_err(ye[p], orc.euler(f, G, y0, tspan, dW[p])) <= TOL
        assert rel_err(yh[p], orc.heun(f, G, y0, tspan, dW[p])) <= TOL
    # a seeded numpy generator gives reproducible results and is advanced
    g1, g2 = np.random.default_rng(8), np.random.default_rng(8)
    a = sdb.itoEuler(f, G, y0, tspan, generator=g1)
    b = sdb.itoEuler(f, G, y0, tspan, generator=g2)
    assert a.shape == (101, 2) and np.array_equal(a, b)
    assert g1.integers(1 << 30) == g2.integers(1 << 30)
    assert g1.integers(1 << 30) != np.random.default_rng(8).integers(1 << 30)


@pytest.mark.parametrize("d,wik", [(10, False), (20, True)])
def test_long_run_parity_every_integrator(sdb, d, wik):
    """SURVEY 8(d) parity run: P = 64 paths, N - 1 = 1000 steps, EVERY integrator, at the cfg3 shape
    (SYS_LIN(10,10), Ikpw/Jkpw) and the cfg4 shape (SYS_LIN(20,20), Iwik/Jwik), on the noise the
    production kernels generate on the device (replayed through the standalone deltaW / Ikpw.. kernels,
    which share the generator), every saved state against the oracle: max_n |y - y_ref| / |y_ref|
    <= 1e-12 (1e-11 for KP2iS: Newton to 1e-12 on the device, to 1e-13 in the oracle; first 16 paths).
    Error growth over 1000 steps through the DMMA / TMEM kernels is what this pins."""
    m = d
    f, G, y0 = sdb.ops.sys_lin(d, m)
    N = 1001
    tspan = np.linspace(0.0, 1.0, N)
    h = (tspan[-1] - tspan[0]) / (N - 1)
    P, seed = 64, 8675309
    Im, Jm = (sdb.Iwik, sdb.Jwik) if wik else (sdb.Ikpw, sdb.Jkpw)
    dW = sdb.deltaW(N - 1, m, h, paths=P, seed=seed)
    _, I = Im(dW, h, seed=seed, ensemble=True)
    J = I + 0.5 * h * np.eye(m)
    runs = {
        "itoEuler": (sdb.itoEuler(f, G, y0, tspan, paths=P, seed=seed),
                     lambda p: orc.euler(f, G, y0, tspan, dW[p]), 1e-12, P),
        "stratHeun": (sdb.stratHeun(f, G, y0, tspan, paths=P, seed=seed),
                      lambda p: orc.heun(f, G, y0, tspan, dW[p]), 1e-12, P),
        "itoSRI2": (sdb.itoSRI2(f, G, y0, tspan, Im, paths=P, seed=seed),
                    lambda p: orc.srk2(f, G.columns(), y0, tspan, dW[p], I[p]), 1e-12, P),
        "stratSRS2": (sdb.stratSRS2(f, G, y0, tspan, Jm, paths=P, seed=seed),
                      lambda p: orc.srk2(f, G.columns(), y0, tspan, dW[p], J[p]), 1e-12, P),
        "stratKP2iS": (sdb.stratKP2iS(f, G, y0, tspan, Jm, rtol=1e-12, paths=16, seed=seed),
                       lambda p: orc.kp2is(f, G, y0, tspan, dW[p], J[p], rtol=1e-13, solver="newton"),
                       1e-11, 16),
    }
    worst = {}
    for name, (y, ref, tol, count) in runs.items():
        assert y.shape == (count, N, d), name
        worst[name] = max(rel_err(y[p], ref(p)) for p in range(count))
        assert worst[name] <= tol, (name, worst)
    print("long-run parity d=m=%d %s: %s" % (d, "wik" if wik else "kpw",
                                             ", ".join("%s %.1e" % kv for kv in worst.items())))


@pytest.mark.parametrize("name", ["kp4446", "kp4459", "kps445", "r74", "lin2", "lin10"])
def test_new_schemes_match_oracle(sdb, name):
    """The two schemes from the reference's to-do list (README.rst:125-128), which it does not have:
    stratR2 (Burrage & Burrage's two-stage SRK) against the oracle's restatement of the published
    tableau, and itoKP2i (the Ito version of the two-step implicit scheme) against the oracle's KP2iS
    recursion fed with the Ito drift and Ito integrals -- on the golden fixtures' noise."""
    g = load_golden("integrate_%s.npz" % name)
    f_ito, G, y0 = SYSTEMS[name](False)
    f_str, _, _ = SYSTEMS[name](True)
    tspan, dW, I, y0v = g["tspan"], g["dW"], g["I"], g["y0"]
    y = sdb.stratR2(f_str, G, y0, tspan, dW=dW)
    # (KPS445 evaluates sin / cos 1000 times per path: CUDA's and glibc's differ in the last bit, which
    # the 250 steps amplify to 1.5e-12 -- there is no reference implementation of R2 to be closer to)
    tol_r2 = 4e-12 if name == "kps445" else TOL
    for p in range(dW.shape[0]):
        assert rel_err(y[p], orc.r2(f_str, G, y0v, tspan, dW[p])) <= tol_r2
    yk = sdb.itoKP2i(f_ito, G, y0, tspan, dW=dW, I=I, rtol=1e-12)
    for p in range(dW.shape[0]):
        ref = orc.kp2is(f_ito, G, y0v, tspan, dW[p], I[p], rtol=1e-13, solver="newton")
        assert rel_err(yk[p], ref) <= 1e-11
    # on-device noise: same generator slots as stratHeun / stratKP2iS
    P, seed = 5, 77
    N = len(tspan)
    h = (tspan[-1] - tspan[0]) / (N - 1)
    yd = sdb.stratR2(f_str, G, y0, tspan, paths=P, seed=seed)
    dWd = sdb.deltaW(N - 1, G.m, h, paths=P, seed=seed)
    for p in range(P):
        assert rel_err(yd[p], orc.r2(f_str, G, y0v, tspan, dWd[p])) <= tol_r2
    ykd = sdb.itoKP2i(f_ito, G, y0, tspan, sdb.Iwik, rtol=1e-12, paths=P, seed=seed)
    _, Id = sdb.Iwik(dWd, h, seed=seed, ensemble=True)
    for p in range(P):
        ref = orc.kp2is(f_ito, G, y0v, tspan, dWd[p], Id[p], rtol=1e-13, solver="newton")
        assert rel_err(ykd[p], ref) <= 1e-11


def test_callable_repeated_integral_method(sdb):
    """Imethod / Jmethod may be ANY function (dW, h, generator=...) -> (A, I) like in the reference
    (integrate.py:306, :371, :487): it is evaluated on the host and its integrals are handed to the
    kernels.  Here: the oracle's Ikpw (numpy PCG64 draws) as the method, single path and ensemble."""
    f, G, y0 = SYSTEMS["r74"](False)
    tspan = np.linspace(0.0, 0.1, 21)
    h = 0.005
    calls = []

    def my_ikpw(dW, h_, generator=None):
        calls.append(dW.shape)
        return orc.ikpw(dW, h_, n=7, generator=generator)

    dW = sdb.deltaW(20, 4, h, seed=3)
    y = sdb.itoSRI2(f, G, y0, tspan, Imethod=my_ikpw, dW=dW, generator=np.random.default_rng(5))
    _, I = orc.ikpw(dW, h, n=7, generator=np.random.default_rng(5))
    assert calls == [(20, 4)] and rel_err(y, orc.srk2(f, G, y0, tspan, dW, I)) <= TOL
    # ensemble: one host call per path, the generator advancing from path to path
    dWp = sdb.deltaW(20, 4, h, paths=3, seed=4)
    yp = sdb.itoSRI2(f, G, y0, tspan, Imethod=my_ikpw, dW=dWp, generator=np.random.default_rng(6))
    gen = np.random.default_rng(6)
    for p in range(3):
        _, Ip = orc.ikpw(dWp[p], h, n=7, generator=gen)
        assert rel_err(yp[p], orc.srk2(f, G, y0, tspan, dWp[p], Ip)) <= TOL
    # dW drawn on the device when only the method is given; a method with the wrong shape is refused
    y2 = sdb.stratSRS2(SYSTEMS["r74"](True)[0], G, y0, tspan,
                       Jmethod=lambda dW, h_, generator=None: (None, np.zeros((20, 4, 4))), seed=9)
    assert y2.shape == (21, 5)
    with pytest.raises(sdb.SDEValueError):
        sdb.itoSRI2(f, G, y0, tspan, Imethod=lambda dW, h_, generator=None: (None, np.zeros((20, 4))), seed=9)


def test_mixed_noise_inputs(sdb):
    """Only one of dW / I supplied: the other is drawn on the device (integrate.py:481-489)."""
    f, G, y0 = SYSTEMS["r74"](False)
    tspan = np.linspace(0.0, 0.1, 21)
    h = 0.005
    dW = sdb.deltaW(20, 4, h, seed=3)
    y = sdb.itoSRI2(f, G, y0, tspan, dW=dW, seed=11)
    _, I = sdb.Ikpw(dW, h, seed=11)
    assert rel_err(y, orc.srk2(f, G, y0, tspan, dW, I)) <= TOL
    y = sdb.itoSRI2(f, G, y0, tspan, I=I, seed=12)
    dW2 = sdb.deltaW(20, 4, h, seed=12)
    assert rel_err(y, orc.srk2(f, G, y0, tspan, dW2, I)) <= TOL


def test_kp2is_device_noise_and_params(sdb):
    f, G, y0 = SYSTEMS["r74"](True)
    tspan = np.linspace(0.0, 0.2, 41)
    h = 0.005
    P, seed = 6, 5150
    gam, al1, al2 = np.full(5, 0.3), np.full(5, 0.6), np.full(5, 0.8)
    y = sdb.stratKP2iS(f, G, y0, tspan, sdb.Jwik, gam=gam, al1=al1, al2=al2, rtol=1e-12,
                       paths=P, seed=seed)
    dW = sdb.deltaW(40, 4, h, paths=P, seed=seed)
    _, J = sdb.Jwik(dW, h, seed=seed, ensemble=True)
    for p in range(P):
        ref = orc.kp2is(f, G, y0, tspan, dW[p], J[p], gam, al1, al2, rtol=1e-13, solver="newton")
        assert rel_err(y[p], ref) <= 1e-11


# ---- argument validation (integrate.py:47-121 semantics) ---------------------------------------

def test_argument_validation(sdb):
    f, G, y0 = SYSTEMS["r74"](False)
    tspan = np.linspace(0.0, 0.1, 21)
    with pytest.raises(sdb.SDEValueError):                       # unequal spacing
        sdb.itoint(f, G, y0, np.array([0.0, 0.1, 0.3]))
    with pytest.raises(sdb.SDEValueError):                       # y0 / f mismatch (test_mismatched_f)
        sdb.itoint(f, G, np.zeros(3), tspan)
    with pytest.raises(sdb.SDEValueError):                       # wrong dW shape
        sdb.itoEuler(f, G, y0, tspan, dW=np.zeros((20, 3)))
    with pytest.raises(sdb.SDEValueError):                       # nested list has no .shape
        sdb.itoEuler(f, G, y0, tspan, dW=[[0.0] * 4] * 20)
    with pytest.raises(sdb.SDEValueError):                       # wrong I shape
        sdb.itoSRI2(f, G, y0, tspan, dW=np.zeros((20, 4)), I=np.zeros((20, 4, 3)))
    with pytest.raises(sdb.SDEValueError):                       # Python callables that cannot be traced
        sdb.itoint(lambda y, t: -y if y[0] > 0 else y, lambda y, t: np.eye(5), y0, tspan)
    with pytest.raises(TypeError):                               # list-G only for SRI2/SRS2
        sdb.itoEuler(f, G.columns(), y0, tspan)
    with pytest.raises(sdb.SDEValueError):
        sdb.stratKP2iS(f, G.columns(), y0, tspan)
    with pytest.raises(sdb.SDEValueError):                       # fp64 only
        sdb.itoint(f, G, np.ones(5, dtype=np.float32), tspan)
    assert issubclass(sdb.SDEValueError, sdb.Error)


# ---- moments / histogram kernels ----------------------------------------------------------------

def test_moments_and_histogram(sdb):
    import torch
    rng = np.random.default_rng(0)
    S, d, P = 3, 4, 10007
    y = rng.normal(0.5, 1.0, (S, d, P))
    yd = torch.from_numpy(y).cuda()
    sums, counts = sdb.ensemble.moments(yd, hist=(32, -2.0, 3.0))
    sums, counts = sums.cpu().numpy(), counts.cpu().numpy()
    assert np.allclose(sums[..., 0], y.sum(-1), rtol=1e-12)
    assert np.allclose(sums[..., 1], (y * y).sum(-1), rtol=1e-12)
    for s in range(S):
        for i in range(d):
            want, _ = np.histogram(y[s, i], bins=32, range=(-2.0, 3.0))
            inside = (y[s, i] >= -2.0) & (y[s, i] < 3.0)
            assert counts[s, i].sum() == inside.sum()
            assert np.abs(counts[s, i] - want).sum() <= 2     # edge rounding only
    # deterministic: same input, same bits
    sums2, _ = sdb.ensemble.moments(yd)
    assert np.array_equal(sums2.cpu().numpy(), sums)


# ---- NVRTC user operators ------------------------------------------------------------------------

USER_SRC = r"""
__device__ void sde_drift(const double* y, double t, const double* p, double* out) {
  // damped oscillator with a time-dependent forcing: p = (k, c, amp)
  out[0] = y[1];
  out[1] = -p[0] * y[0] - p[1] * y[1] + p[2] * sin(t);
}
__device__ void sde_diffusion_col(int k, const double* y, double t, const double* p, double* out) {
  // column 0: additive on velocity; column 1: multiplicative on both
  if (k == 0) { out[0] = 0.0; out[1] = p[0]; }
  else        { out[0] = p[1] * y[0]; out[1] = p[1] * y[1] * cos(t); }
}
"""


def test_nvrtc_user_system(sdb):
    kf = np.array([4.0, 0.3, 0.5])
    kg = np.array([0.2, 0.1])

    def f_host(y, t):
        return np.array([y[1], -kf[0] * y[0] - kf[1] * y[1] + kf[2] * np.sin(t)])

    def g_host(y, t):
        return np.array([[0.0, kg[1] * y[0]], [kg[0], kg[1] * y[1] * np.cos(t)]])

    f = sdb.ops.CudaDrift(USER_SRC, 2, kf, host=f_host)
    G = sdb.ops.CudaDiffusion(USER_SRC, 2, 2, kg, host=g_host)
    y0 = np.array([1.0, 0.0])
    tspan = np.linspace(0.0, 1.0, 201)
    rng = np.random.default_rng(1)
    h = 0.005
    dW = orc.delta_w(200, 2, h, rng)
    _, I = orc.ikpw(dW, h, generator=rng)
    _, J = orc.jkpw(dW, h, generator=rng)
    assert rel_err(sdb.itoEuler(f, G, y0, tspan, dW=dW), orc.euler(f_host, g_host, y0, tspan, dW)) <= TOL
    assert rel_err(sdb.stratHeun(f, G, y0, tspan, dW=dW), orc.heun(f_host, g_host, y0, tspan, dW)) <= TOL
    assert rel_err(sdb.itoSRI2(f, G, y0, tspan, dW=dW, I=I), orc.srk2(f_host, g_host, y0, tspan, dW, I)) <= TOL
    assert rel_err(sdb.stratSRS2(f, G, y0, tspan, dW=dW, J=J), orc.srk2(f_host, g_host, y0, tspan, dW, J)) <= TOL
    yk = sdb.stratKP2iS(f, G, y0, tspan, dW=dW, J=J, rtol=1e-12)
    assert rel_err(yk, orc.kp2is(f_host, g_host, y0, tspan, dW, J, rtol=1e-13, solver="newton")) <= 1e-11


def test_jit_builtin_dimensions(sdb):
    """A (d, m) with no ahead-of-time instantiation goes through NVRTC with the same header."""
    rng = np.random.default_rng(3)
    d, m = 3, 2
    A = -np.eye(d) + 0.1 * rng.standard_normal((d, d))
    B = 0.2 * rng.standard_normal((m, d, d))
    f, G = sdb.ops.LinearDrift(A), sdb.ops.LinearDiffusion(B)
    y0 = np.ones(d)
    tspan = np.linspace(0, 0.5, 51)
    dW = orc.delta_w(50, m, 0.01, rng)
    _, I = orc.iwik(dW, 0.01, generator=rng)
    assert rel_err(sdb.itoSRI2(f, G, y0, tspan, dW=dW, I=I), orc.srk2(f, G, y0, tspan, dW, I)) <= TOL
    # repeated integrals for an m without an AOT kernel
    dW13 = orc.delta_w(6, 13, 0.01, rng)
    X, Y, Gt = orc.draw_wik_normals(np.random.default_rng(4), 6, 13, 4)
    At, Iw = sdb.Iwik(dW13, 0.01, n=4, normals=(X, Y, Gt))
    At_o, Iw_o = orc.iwik_from_normals(dW13, 0.01, X, Y, Gt)
    assert np.max(np.abs(Iw - Iw_o)) <= 1e-12 * np.max(np.abs(Iw_o))
    assert np.max(np.abs(At - At_o)) <= 1e-12 * np.max(np.abs(Iw_o))


@pytest.mark.parametrize("d,m", [(10, 10), (4, 4), (8, 2), (5, 4),          # ahead-of-time kernels
                                 (6, 4), (4, 6), (12, 6), (9, 10), (15, 8),   # NVRTC-compiled fused kernel
                                 (7, 5), (10, 3), (5, 9)])                     # ... odd numbers of Wiener processes
@pytest.mark.parametrize("strat", [False, True])
def test_fused_dmma_kernel_shapes(sdb, d, m, strat, monkeypatch):
    """The fused linear SRI2/SRS2 kernel (DMMA products, exact linear collapse of the two
    stage-2 evaluations) against the generic two-evaluation loop kernel and the oracle, for
    every instantiated (d, m): DMMA tile / leftover-row / row-block edge cases, a partial last
    warp (P not a multiple of 32) and per-path initial states."""
    rng = np.random.default_rng(100 * d + m)
    A = -1.2 * np.eye(d) + 0.3 * rng.standard_normal((d, d)) / np.sqrt(d)
    B = 0.4 * rng.standard_normal((m, d, d)) / np.sqrt(d * m)
    f, G = sdb.ops.LinearDrift(A), sdb.ops.LinearDiffusion(B)
    y0 = rng.standard_normal(d)
    tspan = np.linspace(0.0, 0.3, 31)
    h = 0.01
    P, seed = 77, 424242
    integ = sdb.stratSRS2 if strat else sdb.itoSRI2
    rep = sdb.Jkpw if strat else sdb.Ikpw
    y0p = y0[None, :] + 0.1 * rng.standard_normal((P, d))
    y_fused = integ(f, G, y0, tspan, paths=P, seed=seed, path_id0=11, y0_paths=y0p)
    monkeypatch.setenv("SDB_DISABLE_FUSED", "1")
    y_generic = integ(f, G, y0, tspan, paths=P, seed=seed, path_id0=11, y0_paths=y0p)
    monkeypatch.delenv("SDB_DISABLE_FUSED")
    assert rel_err(y_fused, y_generic) <= TOL
    dW = sdb.deltaW(30, m, h, paths=P, seed=seed, path_id0=11)
    _, IJ = rep(dW, h, seed=seed, path_id0=11, ensemble=True)
    for p in (0, 31, 32, P - 1):
        ref = orc.srk2(f, G, y0p[p], tspan, dW[p], IJ[p])
        assert rel_err(y_fused[p], ref) <= TOL


@pytest.mark.parametrize("variant", ["1", "10", "11", "13"])
def test_fused_tuning_variants_agree(sdb, variant, monkeypatch):
    """The tuning variants of the headline kernel kept for A/B measurements -- pair-at-a-time
    noise (1), warp-specialised with a tensor-memory ring (10), the same pipeline without helper
    warps (11), the TMEM-resident pipeline with 12 warps per SM at 168 registers (13) -- integrate the same Philox streams: they must agree with the default kernel to
    rounding and with the oracle to 1e-12."""
    f, G, y0 = SYSTEMS["lin10"](False)
    tspan = np.linspace(0.0, 0.25, 51)
    h = 0.005
    P, seed = 300, 777            # two blocks of the 256-path kernels, partial last warp
    base = sdb.itoSRI2(f, G, y0, tspan, paths=P, seed=seed, save_every=5)
    monkeypatch.setenv("SDB_FUSED_VARIANT", variant)
    alt = sdb.itoSRI2(f, G, y0, tspan, paths=P, seed=seed, save_every=5)
    alt_s = sdb.stratSRS2(f, G, y0, tspan, paths=P, seed=seed, save_every=5)
    monkeypatch.delenv("SDB_FUSED_VARIANT")
    base_s = sdb.stratSRS2(f, G, y0, tspan, paths=P, seed=seed, save_every=5)
    assert rel_err(alt, base) <= TOL and rel_err(alt_s, base_s) <= TOL
    dW = sdb.deltaW(50, 10, h, paths=P, seed=seed)
    _, I = sdb.Ikpw(dW, h, seed=seed, ensemble=True)
    for p in (0, 255, 256, P - 1):
        ref = orc.srk2(f, G, y0, tspan, dW[p], I[p])
        assert rel_err(alt[p], ref[::5]) <= TOL


def test_register_resident_wiktorsson_variant(sdb, monkeypatch):
    """Iwik/Jwik at d = m = 10 runs the TMEM-resident kernel; the register-resident kernel with
    Wiktorsson's tail (sdb_fused.cuh WIK, SDB_FUSED_VARIANT=30, kept for A/B: 3 % slower) draws the same
    noise and agrees with it to rounding and with the oracle to 1e-12."""
    f, G, y0 = SYSTEMS["lin10"](False)
    tspan = np.linspace(0.0, 0.1, 21)
    P, seed = 300, 31
    base = sdb.itoSRI2(f, G, y0, tspan, sdb.Iwik, paths=P, seed=seed)
    assert sdb.last_kernel() == "sdbk_fusedx_srk2_d10_m10_wik"
    monkeypatch.setenv("SDB_FUSED_VARIANT", "30")
    alt = sdb.itoSRI2(f, G, y0, tspan, sdb.Iwik, paths=P, seed=seed)
    assert sdb.last_kernel() == "sdbk_fused_srk2_d10_m10_wik"
    alt_s = sdb.stratSRS2(f, G, y0, tspan, sdb.Jwik, paths=P, seed=seed)
    monkeypatch.delenv("SDB_FUSED_VARIANT")
    assert rel_err(alt, base) <= 1e-14
    assert rel_err(alt_s, sdb.stratSRS2(f, G, y0, tspan, sdb.Jwik, paths=P, seed=seed)) <= 1e-14
    dW = sdb.deltaW(20, 10, 0.005, paths=P, seed=seed)
    _, I = sdb.Iwik(dW, 0.005, seed=seed, ensemble=True)
    for p in (0, 256, P - 1):
        assert rel_err(alt[p], orc.srk2(f, G, y0, tspan, dW[p], I[p])) <= TOL


@pytest.mark.parametrize("strat,method", [(False, "kpw"), (True, "kpw"), (False, "wik"), (True, "wik")])
def test_fused_tmem_kernel_d20(sdb, strat, method, monkeypatch):
    """d = m = 20 (BASELINE configs[3]): the TMEM-resident fused kernel (sdb_fusedx.cuh: Levy areas
    read-modify-written in tensor memory, two-pass Wiktorsson tail) against the generic loop kernel
    and against the oracle on the noise the device generated."""
    f, G, y0 = sdb.ops.sys_lin(20, 20)
    tspan = np.linspace(0.0, 0.1, 21)
    h = 0.005
    P, seed = 70, 20262026
    integ = sdb.stratSRS2 if strat else sdb.itoSRI2
    rep = {(False, "kpw"): sdb.Ikpw, (False, "wik"): sdb.Iwik,
           (True, "kpw"): sdb.Jkpw, (True, "wik"): sdb.Jwik}[(strat, method)]
    y = integ(f, G, y0, tspan, rep, paths=P, seed=seed, path_id0=5)
    assert y.shape == (P, 21, 20) and np.all(np.isfinite(y))
    monkeypatch.setenv("SDB_DISABLE_FUSED", "1")
    y_gen = integ(f, G, y0, tspan, rep, paths=P, seed=seed, path_id0=5)
    monkeypatch.delenv("SDB_DISABLE_FUSED")
    assert rel_err(y, y_gen) <= TOL
    dW = sdb.deltaW(20, 20, h, paths=P, seed=seed, path_id0=5)
    _, IJ = rep(dW, h, seed=seed, path_id0=5, ensemble=True)
    for p in (0, 31, 32, P - 1):
        ref = orc.srk2(f, G, y0, tspan, dW[p], IJ[p])
        assert rel_err(y[p], ref) <= TOL
    yd = integ(f, G, y0, tspan, rep, paths=P, seed=seed, path_id0=5, save_every=4)
    assert np.array_equal(yd, y[:, ::4])


def test_fused_tmem_kernel_d20_one_and_two_lanes_per_path(sdb, monkeypatch):
    """d = m = 20 runs two lanes per path by default (sdb_fusedx2.cuh: warp pairs sharing a TMEM lane
    quadrant); the one-lane kernel (SDB_FUSED_VARIANT=20) draws the same noise and agrees to rounding,
    also for a block with idle pairs and for more than one block."""
    f, G, y0 = sdb.ops.sys_lin(20, 20)
    tspan = np.linspace(0.0, 0.05, 11)
    for P in (33, 128 * 3 + 5):
        y2 = sdb.stratSRS2(f, G, y0, tspan, sdb.Jwik, paths=P, seed=11, path_id0=P)
        monkeypatch.setenv("SDB_FUSED_VARIANT", "20")
        y1 = sdb.stratSRS2(f, G, y0, tspan, sdb.Jwik, paths=P, seed=11, path_id0=P)
        monkeypatch.delenv("SDB_FUSED_VARIANT")
        assert rel_err(y2, y1) <= 1e-14 and not np.array_equal(y2, y1)


def test_additive_noise_never_builds_repeated_integrals(sdb):
    """Noise-structure dispatch, first case (integrate.py:150-151): for additive noise the stage-2
    differences of SRI2/SRS2 and the G(Ybar) - G_n of KP2iS vanish identically, so I / J are neither
    generated nor read -- and the results stay those of the reference."""
    g = load_golden("integrate_lin2.npz")
    f, G, y0 = SYSTEMS["lin2"](False)
    tspan, dW, I, J = g["tspan"], g["dW"], g["I"], g["J"]
    with_I = sdb.itoSRI2(f, G, y0, tspan, dW=dW, I=I)
    n0 = sdb.launch_count()
    without_I = sdb.itoSRI2(f, G, y0, tspan, dW=dW)
    assert sdb.launch_count() - n0 == 2            # status init + the step loop: no repint kernel
    assert np.array_equal(with_I, without_I)
    assert rel_err(without_I, g["y_itoSRI2"]) <= TOL
    assert rel_err(sdb.stratSRS2(f, G, y0, tspan, dW=dW), g["y_stratSRS2"]) <= TOL
    if "y_stratKP2iS" in g:
        assert rel_err(sdb.stratKP2iS(f, G, y0, tspan, dW=dW, rtol=1e-12), g["y_stratKP2iS"]) <= 5e-12
    # on-device noise: only dW is drawn; any I gives the same oracle result
    P, seed = 50, 606
    y = sdb.itoSRI2(f, G, y0, tspan, paths=P, seed=seed)
    h = (tspan[-1] - tspan[0]) / (len(tspan) - 1)
    dWd = sdb.deltaW(len(tspan) - 1, 2, h, paths=P, seed=seed)
    for p in (0, P - 1):
        ref = orc.srk2(f, G, y0, tspan, dWd[p], np.zeros((len(tspan) - 1, 2, 2)))
        assert rel_err(y[p], ref) <= TOL


# ---- edge cases of the ensemble interface ---------------------------------------------------------

@pytest.mark.parametrize("name", ["lin10", "r74", "lin2", "kp4446"])
def test_edge_shapes(sdb, name):
    """One step, one path, decimation that does not divide N-1 or exceeds it, path ids and seeds
    beyond 32 bits, non-default series lengths: fused and generic kernels alike."""
    f, G, y0 = SYSTEMS[name](False)
    y0v = np.atleast_1d(np.asarray(y0, dtype=float))
    m, d = G.m, len(y0v)
    big_id, big_seed = (1 << 40) + 12345, (1 << 63) + 987654321
    # N = 2: a single step, a single path
    t2 = np.array([0.0, 0.01])
    y = sdb.itoSRI2(f, G, y0, t2, paths=1, seed=big_seed, path_id0=big_id)
    assert y.shape == (1, 2, d) and np.array_equal(y[0, 0], y0v)
    dW = sdb.deltaW(1, m, 0.01, paths=1, seed=big_seed, path_id0=big_id)
    _, I = sdb.Ikpw(dW, 0.01, seed=big_seed, path_id0=big_id, ensemble=True)
    assert rel_err(y[0], orc.srk2(f, G, y0v, t2, dW[0], I[0])) <= TOL
    # the high word of the path id matters
    y_lo = sdb.itoSRI2(f, G, y0, t2, paths=1, seed=big_seed, path_id0=12345)
    assert not np.array_equal(y_lo, y)
    # save_every that does not divide N-1 / is larger than N-1
    tspan = np.linspace(0.0, 0.17, 18)
    full = sdb.itoSRI2(f, G, y0, tspan, paths=33, seed=5)
    part = sdb.itoSRI2(f, G, y0, tspan, paths=33, seed=5, save_every=5)
    assert part.shape == (33, 4, d) and np.array_equal(part, full[:, ::5])
    only0 = sdb.itoSRI2(f, G, y0, tspan, paths=33, seed=5, save_every=100)
    assert only0.shape == (33, 1, d) and np.array_equal(only0[:, 0], full[:, 0])
    # series length n != 5 (Ikpw(dW, h, n)): the integrators use the default n = 5 like the
    # reference; the standalone constructions take any n >= 1
    h = 0.01
    dW1 = sdb.deltaW(17, m, h, paths=2, seed=9)
    for n in (1, 8):
        A, I = sdb.Ikpw(dW1, h, n=n, seed=9, ensemble=True)
        X, Y = po.kpw_normals(17, m, n, 2, 9)
        _, I_o = orc.ikpw_from_normals(dW1[1], h, X[:, 1], Y[:, 1])
        assert np.max(np.abs(I[1] - I_o)) <= 1e-12 * max(np.max(np.abs(I_o)), 1e-300)


DIAG_SRC = r"""
__device__ void sde_drift(const double* y, double t, const double* p, double* out) {
  for (int i = 0; i < 4; ++i) out[i] = -p[0] * y[i] + p[1] * y[(i + 1) & 3];
}
__device__ void sde_diffusion_col(int k, const double* y, double t, const double* p, double* out) {
  for (int i = 0; i < 4; ++i) out[i] = 0.0;
  out[k] = p[0] * sin(y[k]) + p[1];        // column k: only component k, depends only on y_k
}
"""


def test_diagonal_noise_needs_no_levy_areas(sdb):
    """Noise-structure dispatch, second case: for diagonal noise SRI2/SRS2 depend on I only through
    I_kk = (dW_k^2 - h)/2, so nothing but dW is generated (or read).  Checked against the oracle fed
    with the FULL I of Ikpw (the reference's computation) and with the diagonal only."""
    rng = np.random.default_rng(17)
    d = 6
    A = -np.eye(d) + 0.2 * rng.standard_normal((d, d))
    b = 0.3 + 0.2 * rng.random(d)
    f, G = sdb.ops.LinearDrift(A), sdb.ops.DiagLinearDiffusion(b)
    y0 = 1.0 + rng.random(d)
    tspan = np.linspace(0.0, 0.4, 81)
    h = 0.005
    dW = orc.delta_w(80, d, h, rng)
    _, I_full = orc.ikpw(dW, h, generator=rng)
    _, J_full = orc.jkpw(dW, h, generator=rng)
    n0 = sdb.launch_count()
    y = sdb.itoSRI2(f, G, y0, tspan, dW=dW)                           # no I: diagonal built on device
    assert sdb.launch_count() - n0 == 2                               # no repeated-integral kernel
    assert rel_err(y, orc.srk2(f, G, y0, tspan, dW, I_full)) <= TOL
    assert rel_err(sdb.itoSRI2(f, G, y0, tspan, dW=dW, I=I_full), y) <= TOL
    ys = sdb.stratSRS2(f, G, y0, tspan, dW=dW)
    assert rel_err(ys, orc.srk2(f, G, y0, tspan, dW, J_full)) <= TOL
    # on-device noise: only the m dW normals of each step are drawn
    P, seed = 40, 2468
    yp = sdb.itoSRI2(f, G, y0, tspan, paths=P, seed=seed)
    dWd = sdb.deltaW(80, d, h, paths=P, seed=seed)
    for p in (0, P - 1):
        Idiag = np.zeros((80, d, d))
        Idiag[:, np.arange(d), np.arange(d)] = 0.5 * (dWd[p] ** 2 - h)
        assert rel_err(yp[p], orc.srk2(f, G, y0, tspan, dWd[p], Idiag)) <= TOL
    # KP2iS does use off-diagonal J: it still gets the full construction
    yk = sdb.stratKP2iS(f, G, y0, tspan, dW=dW, J=J_full, rtol=1e-12)
    ref = orc.kp2is(f, G, y0, tspan, dW, J_full, rtol=1e-13, solver="newton")
    assert rel_err(yk, ref) <= 1e-11
    # user CUDA source with the diagonal promise
    kf, kg = np.array([1.2, 0.3]), np.array([0.25, 0.1])

    def f_host(yv, t):
        return np.array([-kf[0] * yv[i] + kf[1] * yv[(i + 1) & 3] for i in range(4)])

    def g_host(yv, t):
        return np.diag(kg[0] * np.sin(yv) + kg[1])

    fu = sdb.ops.CudaDrift(DIAG_SRC, 4, kf, host=f_host)
    Gu = sdb.ops.CudaDiffusion(DIAG_SRC, 4, 4, kg, host=g_host, diagonal=True)
    y0u = np.array([0.5, -0.2, 0.8, 0.1])
    dWu = orc.delta_w(80, 4, h, rng)
    _, Iu = orc.ikpw(dWu, h, generator=rng)
    yu = sdb.itoSRI2(fu, Gu, y0u, tspan, dW=dWu)
    assert rel_err(yu, orc.srk2(f_host, g_host, y0u, tspan, dWu, Iu)) <= TOL


@pytest.mark.parametrize("d,m,method,strat", [(6, 4, "wik", False), (12, 8, "wik", True),
                                              (16, 12, "kpw", False), (14, 14, "wik", False),
                                              (8, 16, "wik", False), (12, 18, "kpw", True)])
def test_fused_tmem_kernel_jit_shapes(sdb, d, m, method, strat, monkeypatch):
    """Shapes / methods without an ahead-of-time kernel: sdb_fusedx.cuh -- or, from m = 16, where only
    one warp per sub-partition fits, the two-lanes-per-path sdb_fusedx2.cuh -- compiled by NVRTC
    (run-time shape mirror in sdb_fused_rt.h) against the generic loop kernel and the oracle."""
    rng = np.random.default_rng(1000 * d + m)
    A = -1.1 * np.eye(d) + 0.3 * rng.standard_normal((d, d)) / np.sqrt(d)
    B = 0.4 * rng.standard_normal((m, d, d)) / np.sqrt(d * m)
    f, G = sdb.ops.LinearDrift(A), sdb.ops.LinearDiffusion(B)
    y0 = rng.standard_normal(d)
    tspan = np.linspace(0.0, 0.12, 25)
    h = 0.005
    P, seed = 45, 90210
    integ = sdb.stratSRS2 if strat else sdb.itoSRI2
    rep = {(False, "kpw"): sdb.Ikpw, (False, "wik"): sdb.Iwik,
           (True, "kpw"): sdb.Jkpw, (True, "wik"): sdb.Jwik}[(strat, method)]
    n0 = sdb.launch_count()
    y = integ(f, G, y0, tspan, rep, paths=P, seed=seed, path_id0=2)
    assert sdb.launch_count() - n0 == 2
    monkeypatch.setenv("SDB_DISABLE_FUSED", "1")
    y_gen = integ(f, G, y0, tspan, rep, paths=P, seed=seed, path_id0=2)
    monkeypatch.delenv("SDB_DISABLE_FUSED")
    assert rel_err(y, y_gen) <= TOL
    dW = sdb.deltaW(24, m, h, paths=P, seed=seed, path_id0=2)
    _, IJ = rep(dW, h, seed=seed, path_id0=2, ensemble=True)
    for p in (0, 32, P - 1):
        assert rel_err(y[p], orc.srk2(f, G, y0, tspan, dW[p], IJ[p])) <= TOL


def test_shared_noise_is_broadcast_not_replicated(sdb):
    """A 2-d dW / 3-d I with paths=P is ONE realisation shared by all paths (path stride 0 at the
    C ABI): with per-path initial states the paths differ, with equal ones they coincide."""
    g = load_golden("integrate_r74.npz")
    f, G, y0 = SYSTEMS["r74"](False)
    tspan, dW, I = g["tspan"], g["dW"][0], g["I"][0]
    one = sdb.itoSRI2(f, G, y0, tspan, dW=dW, I=I)
    many = sdb.itoSRI2(f, G, y0, tspan, dW=dW, I=I, paths=37)
    assert many.shape == (37,) + one.shape
    assert all(np.array_equal(many[p], one) for p in range(37))
    y0p = np.asarray(y0)[None, :] * (1.0 + 0.01 * np.arange(37))[:, None]
    spread = sdb.itoSRI2(f, G, y0, tspan, dW=dW, I=I, paths=37, y0_paths=y0p)
    assert rel_err(spread[0], one) <= TOL and not np.array_equal(spread[5], one)
    assert rel_err(spread[5], orc.srk2(f, G, y0p[5], tspan, dW, I)) <= TOL


# ---- commutative noise: no Levy areas (reference to-do integrate.py:150-151) ----

def _commuting_system(d, m, seed=3):
    """f = A y, g_k = B_k y with B_k polynomials in one matrix C, so that [B_j, B_k] = 0."""
    import sdeint_b200 as sdb
    rng = np.random.default_rng(seed)
    Cm = rng.standard_normal((d, d)) / np.sqrt(d)
    B = np.stack([(0.25 / np.sqrt(m)) * ((1.0 + 0.1 * k) * np.eye(d) + 0.5 * Cm + 0.1 * (k - m / 2) * Cm @ Cm)
                  for k in range(m)])
    A = np.full((d, d), 0.5 / d)
    np.fill_diagonal(A, -1.5)
    return sdb.ops.LinearDrift(A), sdb.ops.LinearDiffusion(B), np.ones(d)


def test_icomm_jcomm_standalone(sdb):
    h, m, N, P = 0.01, 6, 9, 33
    dW = sdb.deltaW(N, m, h, paths=P, seed=12)
    A, I = sdb.Icomm(dW, h, ensemble=True)
    _, J = sdb.Jcomm(dW, h, ensemble=True)
    sym = 0.5 * dW[..., :, None] * dW[..., None, :]
    assert A.shape == (P, N, m, m) and not A.any()
    assert np.allclose(I, sym - 0.5 * h * np.eye(m), rtol=0, atol=1e-16)
    assert np.allclose(J, sym, rtol=0, atol=1e-16)
    A1, I1 = sdb.Icomm(dW[0], h)                     # the reference's (N, m) -> (N, m, m) shapes
    assert A1.shape == (N, m, m) and np.array_equal(I1, I[0])


@pytest.mark.parametrize("d,m,strat", [(5, 4, False), (10, 10, False), (10, 10, True), (20, 20, False),
                                       (3, 2, True), (7, 5, False), (9, 3, True)])
def test_commutative_noise_needs_no_levy_areas(sdb, d, m, strat):
    """For commuting B_k the SRI2/SRS2 result does not depend on the Levy areas: the automatic
    area-free path (only dW drawn), the forced full Ikpw construction on the same seed, and the oracle
    fed the device's dW with I from Icomm and from Ikpw all agree pathwise to rounding."""
    f, G, y0 = _commuting_system(d, m)
    assert G.commutative
    tspan = np.linspace(0.0, 0.25, 51)
    h = 0.005
    P, seed = 70, 777
    integ = sdb.stratSRS2 if strat else sdb.itoSRI2
    y = integ(f, G, y0, tspan, paths=P, seed=seed)                       # dispatches to COMM
    full = integ(f, G, y0, tspan, paths=P, seed=seed, commutative=False)  # Levy areas generated
    assert rel_err(y, full) <= TOL and not np.array_equal(y, full)
    explicit = integ(f, G, y0, tspan, sdb.Jcomm if strat else sdb.Icomm, paths=P, seed=seed,
                     commutative=False)
    assert np.array_equal(explicit, y)
    dW = sdb.deltaW(50, m, h, paths=P, seed=seed)
    _, Ic = (sdb.Jcomm if strat else sdb.Icomm)(dW, h, ensemble=True)
    _, Ik = (sdb.Jkpw if strat else sdb.Ikpw)(dW, h, seed=seed, ensemble=True)
    for p in (0, 31, P - 1):
        assert rel_err(y[p], orc.srk2(f, G, y0, tspan, dW[p], Ic[p])) <= TOL
        assert rel_err(y[p], orc.srk2(f, G, y0, tspan, dW[p], Ik[p])) <= TOL
    # dW supplied, I omitted: built from dW without random numbers
    y1 = integ(f, G, y0, tspan, dW=dW[5])
    assert rel_err(y1, y[5]) <= TOL


def test_commutative_kernels_agree(sdb, monkeypatch):
    """The dedicated commutative-noise kernel (sdb_fusedc.cuh) against the Ikpw kernel run with an empty
    series (SDB_DISABLE_FUSEDC) and against the generic loop (SDB_DISABLE_FUSED): same paths."""
    f, G, y0 = _commuting_system(10, 10)
    tspan = np.linspace(0.0, 0.25, 51)
    kw = dict(paths=300, seed=5, path_id0=11, save_every=5)
    y = sdb.stratSRS2(f, G, y0, tspan, **kw)
    monkeypatch.setenv("SDB_DISABLE_FUSEDC", "1")
    assert rel_err(sdb.stratSRS2(f, G, y0, tspan, **kw), y) <= TOL
    monkeypatch.setenv("SDB_DISABLE_FUSED", "1")
    assert rel_err(sdb.stratSRS2(f, G, y0, tspan, **kw), y) <= TOL


def test_commutative_flag_on_user_operator(sdb):
    """A nonlinear system whose columns are multiples of one vector field (g_k = c_k phi(y), so
    L^j g_k = c_j c_k phi' phi is symmetric): `commutative=True` skips the Levy areas; the result is
    the scheme applied to I = (dW dW^T - h)/2 and stays within O(h) of the full construction."""
    import sdeint_b200 as sdb_
    c = [0.3, 0.2, 0.1]
    fX = sdb_.ops.ExprDrift(["-y[0] + 0.2*y[1]", "-0.5*y[1] + 0.1*sin(y[0])"], [])
    rows = [["%g*(1.0+0.5*sin(y[1]))" % ck for ck in c], ["%g*cos(y[0])" % ck for ck in c]]
    Gc = sdb_.ops.ExprDiffusion(rows, [], commutative=True)
    Gn = sdb_.ops.ExprDiffusion(rows, [])
    y0 = np.array([0.5, -0.2])
    tspan = np.linspace(0.0, 0.5, 101)
    h = 0.005
    P, seed = 50, 99
    y = sdb.itoSRI2(fX, Gc, y0, tspan, paths=P, seed=seed)
    dW = sdb.deltaW(100, 3, h, paths=P, seed=seed)
    _, Ic = sdb.Icomm(dW, h, ensemble=True)
    for p in (0, 17, P - 1):
        assert rel_err(y[p], orc.srk2(fX, Gn, y0, tspan, dW[p], Ic[p])) <= TOL
    assert np.array_equal(sdb.itoSRI2(fX, Gn, y0, tspan, paths=P, seed=seed, commutative=True), y)
    full = sdb.itoSRI2(fX, Gn, y0, tspan, paths=P, seed=seed)             # Ikpw, same dW
    assert not np.array_equal(full, y)
    assert np.sqrt(np.mean((full[:, -1] - y[:, -1]) ** 2)) < 5e-3          # both order-1.0 strong


# ---- path functionals evaluated inside the step loop (first passage, extrema) ----

def _functionals(y, tspan, specs):
    """numpy restatement on a full-resolution trajectory y (N, d)."""
    out = []
    for spec in specs:
        c = y[:, spec[1]]
        if spec[0] in ("first_above", "first_below"):
            hit = np.nonzero(c >= spec[2] if spec[0] == "first_above" else c <= spec[2])[0]
            out.append(tspan[hit[0]] if len(hit) else np.inf)
        elif spec[0] == "max":
            out.append(c.max())
        elif spec[0] == "min":
            out.append(c.min())
        else:
            out.append(np.sum(0.5 * np.diff(tspan) * (c[:-1] + c[1:])))
    return np.array(out)


def _check_functionals(obs, y, tspan, specs):
    """obs (P, n_obs) against the trajectories y (P, N, d) the same launch stored."""
    for p in range(y.shape[0]):
        ref = _functionals(y[p], tspan, specs)
        for o, spec in enumerate(specs):
            if spec[0] == "integral":
                assert abs(obs[p, o] - ref[o]) <= 1e-13 * max(1.0, abs(ref[o]))
            else:   # grid times and extrema are exact
                assert obs[p, o] == ref[o], (p, spec, obs[p, o], ref[o])


@pytest.mark.parametrize("name,scheme,method", [
    ("lin10", "sri2", "kpw"), ("lin10", "sri2", "wik"), ("lin10", "heun", None), ("lin10", "euler", None),
    ("r74", "sri2", "kpw"), ("kp4446", "sri2", "kpw"), ("lin2", "euler", None), ("lin20", "srs2", "wik"),
    ("r74", "kp2is", "kpw"),
])
def test_observers_match_full_trajectory(sdb, name, scheme, method):
    """first passage / max / min / integral from inside the step loop == the same functionals of the
    full-resolution trajectory of the same launch; a decimated store changes nothing; the states are
    those of the non-observing kernels."""
    strat = scheme in ("heun", "srs2", "kp2is")
    f, G, y0 = (sdb.ops.sys_lin(20, 20) if name == "lin20" else SYSTEMS[name](strat))
    y0v = np.atleast_1d(np.asarray(y0, dtype=float))
    d = len(y0v)
    tspan = np.linspace(0.0, 0.3, 61)
    P, seed = 45, 2024
    c1 = d - 1
    up, dn = float(y0v[0]) * 1.02 + 0.01, float(y0v[c1]) * 0.9 - 0.01
    specs = [("first_above", 0, up), ("first_below", c1, dn), ("max", 0), ("integral", c1)]
    kw = dict(paths=P, seed=seed)
    if scheme == "sri2":
        call = lambda **k: sdb.itoSRI2(f, G, y0, tspan, sdb.Ikpw if method == "kpw" else sdb.Iwik, **kw, **k)
    elif scheme == "srs2":
        call = lambda **k: sdb.stratSRS2(f, G, y0, tspan, sdb.Jkpw if method == "kpw" else sdb.Jwik, **kw, **k)
    elif scheme == "kp2is":
        call = lambda **k: sdb.stratKP2iS(f, G, y0, tspan, rtol=1e-10, **kw, **k)
    elif scheme == "heun":
        call = lambda **k: sdb.stratHeun(f, G, y0, tspan, **kw, **k)
    else:
        call = lambda **k: sdb.itoEuler(f, G, y0, tspan, **kw, **k)
    y, obs = call(observe=specs)
    assert y.shape == (P, 61, d) and obs.shape == (P, 4)
    _check_functionals(obs, y, tspan, specs)
    assert np.isfinite(obs[:, 0]).any() or np.isfinite(obs[:, 1]).any()    # thresholds are reachable
    yd, obsd = call(observe=specs, save_every=20)
    assert np.array_equal(obsd, obs) and np.array_equal(yd, y[:, ::20])
    assert rel_err(y, call()) <= (1e-9 if scheme == "kp2is" else TOL)   # Newton to rtol=1e-10
    # "min" and a single observer given as one tuple
    _, mn = call(observe=("min", c1))
    assert mn.shape == (P, 1) and np.array_equal(mn[:, 0], y[:, :, c1].min(axis=1))


def test_observers_user_system_single_path_and_validation(sdb):
    fX = sdb.ops.ExprDrift(["-y[0] + 0.2*y[1]", "-0.5*y[1] + 0.1*sin(y[0])"], [])
    GX = sdb.ops.ExprDiffusion([["0.3*(1.0+0.5*sin(y[1]))", "0.1"], ["0.05", "0.2*cos(y[0])"]], [])
    y0 = np.array([0.5, -0.2])
    tspan = np.linspace(0.0, 1.0, 201)
    specs = [("max", 1), ("first_above", 0, 0.55)]
    y, obs = sdb.itoint(fX, GX, y0, tspan, paths=30, seed=4, observe=specs)
    _check_functionals(obs, y, tspan, specs)
    dW = sdb.deltaW(200, 2, 0.005, paths=30, seed=4)
    _, I = sdb.Ikpw(dW, 0.005, seed=4, ensemble=True)
    for p in (0, 29):                                      # and against the oracle's trajectory
        ref = orc.srk2(fX, GX, y0, tspan, dW[p], I[p])
        assert abs(obs[p, 0] - ref[:, 1].max()) <= 1e-12 * max(1.0, abs(ref[:, 1].max()))
    y1, o1 = sdb.itoint(fX, GX, y0, tspan, generator=np.random.default_rng(3), observe=specs)
    assert y1.shape == (201, 2) and o1.shape == (2,) and o1[0] == y1[:, 1].max()
    # host-supplied noise goes through the same observers
    y2, o2 = sdb.itoSRI2(fX, GX, y0, tspan, dW=dW[3], I=I[3], observe=specs)
    assert np.array_equal(o2, _functionals(y2, tspan, specs))
    for bad in ([("median", 0)], [("max", 2)], [("first_above", 0)], [("max", 0)] * 5, []):
        with pytest.raises(sdb.SDEValueError):
            sdb.itoint(fX, GX, y0, tspan, observe=bad)


def test_first_passage_distribution_ou(sdb):
    """First-passage statistics of an Ornstein-Uhlenbeck process through the in-loop observer and the
    on-device histogram: the CDF equals the empirical one of the returned times, and the probability of
    having hit by T agrees with a fine-step numpy simulation of the same SDE within sampling error."""
    f, G, y0 = sdb.ops.sys_ou(1.0, 0.5)
    tspan = np.linspace(0.0, 2.0, 401)
    P = 200000
    y, t = sdb.itoEuler(f, G, y0, tspan, paths=P, seed=8, save_every=400, out="torch",
                        observe=[("first_above", 0, 0.4)])
    assert t.shape == (1, P)
    edges, cdf = sdb.ensemble.first_passage_cdf(t, 20, 0.0, 2.0 + 1e-9, P)
    th = t.cpu().numpy()[0]
    emp = np.array([(th < e).mean() for e in edges])
    assert np.allclose(cdf[0], emp, atol=1e-12) and 0.2 < cdf[0, -1] < 0.9
    rng = np.random.default_rng(1)                       # independent reference simulation, same grid
    Q = 100000
    x = np.zeros(Q)
    hit = np.zeros(Q, dtype=bool)
    h = tspan[1] - tspan[0]
    for _ in range(400):
        x = x - x * h + 0.5 * np.sqrt(h) * rng.standard_normal(Q)
        hit |= x >= 0.4
    se = np.sqrt(cdf[0, -1] * (1 - cdf[0, -1]) * (1.0 / P + 1.0 / Q))
    assert abs(cdf[0, -1] - hit.mean()) < 5 * se


def test_auto_series_length_and_long_series(sdb):
    """n_terms="auto" picks wiener.series_terms(h, m, method) and the fused kernels take long series
    (up to 256 terms): same result as the explicit n, checked against the oracle on replayed normals."""
    f, G, y0 = SYSTEMS["lin10"](False)
    tspan = np.linspace(0.0, 0.02, 11)
    h = 0.002
    n = sdb.wiener.series_terms(h, 10, "kpw")
    assert 70 <= n <= 80
    kw = dict(paths=40, seed=123, path_id0=4)
    y_auto = sdb.itoSRI2(f, G, y0, tspan, n_terms="auto", **kw)
    assert np.array_equal(y_auto, sdb.itoSRI2(f, G, y0, tspan, n_terms=n, **kw))
    assert not np.array_equal(y_auto, sdb.itoSRI2(f, G, y0, tspan, **kw))
    dW = sdb.deltaW(10, 10, h, paths=40, seed=123, path_id0=4)
    _, I = sdb.Ikpw(dW, h, n=n, seed=123, path_id0=4, ensemble=True)
    for p in (0, 39):
        assert rel_err(y_auto[p], orc.srk2(f, G, y0, tspan, dW[p], I[p])) <= TOL
    nw = sdb.wiener.series_terms(h, 10, "wik")
    yw = sdb.itoSRI2(f, G, y0, tspan, sdb.Iwik, n_terms="auto", **kw)
    _, Iw = sdb.Iwik(dW, h, n=nw, seed=123, path_id0=4, ensemble=True)
    assert rel_err(yw[7], orc.srk2(f, G, y0, tspan, dW[7], Iw[7])) <= TOL
    with pytest.raises(sdb.SDEValueError):
        sdb.itoSRI2(f, G, y0, tspan, n_terms="many")


@pytest.mark.parametrize("name,scheme", [("lin10", "sri2"), ("lin10", "sri2wik"), ("lin10", "heun"),
                                         ("lin20", "srs2"), ("kp4446", "sri2"), ("lin2", "euler"),
                                         ("r74", "sri2")])
def test_resumed_segments_are_bit_identical(sdb, name, scheme):
    """Checkpoint / resume: the on-device noise depends only on (seed, path id, step counter), so a run
    cut after k intervals and continued with `step0=k` from the states it had reached reproduces the
    single call bit for bit (every fused and generic kernel; sdb_integrate_segment)."""
    strat = scheme in ("heun", "srs2")
    f, G, y0 = (sdb.ops.sys_lin(20, 20) if name == "lin20" else SYSTEMS[name](strat))
    tspan = np.linspace(0.0, 0.2, 41)
    h = (tspan[-1] - tspan[0]) / 40
    kw = dict(paths=70, seed=99, path_id0=6)
    if scheme == "sri2":
        call = lambda ts, **k: sdb.itoSRI2(f, G, y0, ts, **kw, **k)
    elif scheme == "sri2wik":
        call = lambda ts, **k: sdb.itoSRI2(f, G, y0, ts, sdb.Iwik, **kw, **k)
    elif scheme == "srs2":
        call = lambda ts, **k: sdb.stratSRS2(f, G, y0, ts, sdb.Jwik, **kw, **k)
    elif scheme == "heun":
        call = lambda ts, **k: sdb.stratHeun(f, G, y0, ts, **kw, **k)
    else:
        call = lambda ts, **k: sdb.itoEuler(f, G, y0, ts, **kw, **k)
    whole = call(tspan)
    cut = 17
    first = call(tspan[:cut + 1], h_noise=h)
    assert np.array_equal(first, whole[:, :cut + 1])
    second = call(tspan[cut:], y0_paths=first[:, -1], step0=cut, h_noise=h)
    assert np.array_equal(second, whole[:, cut:])
    assert not np.array_equal(call(tspan[cut:], y0_paths=first[:, -1], h_noise=h), whole[:, cut:])


def test_checkpointed_run_resumes_bit_for_bit(sdb, tmp_path):
    """ensemble.checkpointed: a run in segments with a checkpoint file after each; killing it and
    calling it again continues from the file -- final states identical to one uninterrupted call."""
    f, G, y0 = SYSTEMS["lin10"](False)
    tspan = np.linspace(0.0, 0.3, 61)
    P, seed = 500, 4321
    whole = sdb.itoSRI2(f, G, y0, tspan, paths=P, seed=seed, save_every=60)[:, -1]
    ck = str(tmp_path / "run.npz")
    seen = []

    class Stop(Exception):
        pass

    def crash_after_two(step, states):
        seen.append(step)
        if len(seen) == 2:
            raise Stop()                      # the checkpoint of segment 2 is NOT written

    with pytest.raises(Stop):
        sdb.ensemble.checkpointed(sdb.itoSRI2, f, G, y0, tspan, paths=P, seed=seed, segment=16,
                                  checkpoint=ck, on_segment=crash_after_two)
    assert seen == [16, 32] and int(np.load(ck)["step"]) == 16
    seen.clear()
    final = sdb.ensemble.checkpointed(sdb.itoSRI2, f, G, y0, tspan, paths=P, seed=seed, segment=16,
                                      checkpoint=ck, on_segment=lambda s_, y_: seen.append(s_))
    assert seen == [32, 48, 60] and np.array_equal(final, whole)
    # a finished run returns its states without integrating again
    n0 = sdb.launch_count()
    again = sdb.ensemble.checkpointed(sdb.itoSRI2, f, G, y0, tspan, paths=P, seed=seed, segment=16, checkpoint=ck)
    assert sdb.launch_count() == n0 and np.array_equal(again, whole)
    # the checkpoint of a DIFFERENT run (seed, integrator, system parameters, y0, options) is refused
    # instead of being resumed from unrelated states
    f2 = sdb.ops.LinearDrift(np.array(f.A) * 1.01)
    for args, kw in (((sdb.itoSRI2, f, G, y0), dict(seed=seed + 1)), ((sdb.stratSRS2, f, G, y0), dict(seed=seed)),
                     ((sdb.itoSRI2, f2, G, y0), dict(seed=seed)), ((sdb.itoSRI2, f, G, y0 * 2.0), dict(seed=seed)),
                     ((sdb.itoSRI2, f, G, y0), dict(seed=seed, n_terms=7))):
        with pytest.raises(ValueError):
            sdb.ensemble.checkpointed(*args, tspan, paths=P, segment=30, checkpoint=ck, **kw)
    import os
    os.remove(ck)
    other = sdb.ensemble.checkpointed(sdb.itoSRI2, f, G, y0, tspan, paths=P, seed=seed + 1, segment=30,
                                      checkpoint=ck)
    assert not np.array_equal(other, whole)


def test_resume_argument_checks(sdb):
    f, G, y0 = SYSTEMS["r74"](True)
    tspan = np.linspace(0.0, 0.1, 21)
    with pytest.raises(sdb.SDEValueError):
        sdb.stratKP2iS(f, G, y0, tspan, step0=3)
    with pytest.raises(sdb.SDEValueError):
        sdb.stratSRS2(f, G, y0, tspan, step0=-1)
    with pytest.raises(sdb.SDEValueError):
        sdb.stratSRS2(f, G, y0, tspan, dW=np.zeros((20, 4)), step0=2)      # J would be drawn from step 0


def test_dispatch_launches_the_documented_kernels(sdb, monkeypatch):
    """sdb_last_kernel(): every route of DESIGN.md section 4 ends in the kernel it names -- no silent
    fall-back to the generic loop (ahead-of-time kernels by symbol, NVRTC-built ones by cache key)."""
    ts = np.linspace(0.0, 0.01, 11)
    kw = dict(paths=64, seed=1)

    def ran(fn, *a, **k):
        fn(*a, **k)
        return sdb.last_kernel()

    f, G, y0 = sdb.ops.sys_lin(10, 10)
    assert ran(sdb.itoSRI2, f, G, y0, ts, **kw) == "sdbk_fused_srk2_d10_m10_v0"
    assert ran(sdb.itoSRI2, f, G, y0, ts, sdb.Iwik, **kw) == "sdbk_fusedx_srk2_d10_m10_wik"
    assert ran(sdb.stratHeun, f, G, y0, ts, **kw) == "sdbk_fusedeh_heun_d10_m10"
    assert ran(sdb.itoEuler, f, G, y0, ts, **kw) == "sdbk_fusedeh_euler_d10_m10"
    assert ran(sdb.itoSRI2, f, G, y0, ts, observe=[("max", 0)], **kw) == "jit:fusedjit|obs"
    assert ran(sdb.itoSRI2, f, G, y0, ts, step0=3, **kw) == "jit:fusedjit|res"
    assert ran(sdb.itoSRI2, f, G, y0, ts, dW=np.zeros((10, 10)), I=np.zeros((10, 10, 10))) == "sdbk_lin10_s2_n0"
    monkeypatch.setenv("SDB_DISABLE_FUSED", "1")
    assert ran(sdb.itoSRI2, f, G, y0, ts, **kw) == "sdbk_lin10_s2_n1"
    monkeypatch.delenv("SDB_DISABLE_FUSED")
    f, G, y0 = sdb.ops.sys_lin(20, 20)
    assert ran(sdb.stratSRS2, f, G, y0, ts, sdb.Jwik, **kw) == "sdbk_fusedx2_srk2_d20_m20_wik"
    assert ran(sdb.stratSRS2, f, G, y0, ts, **kw) == "sdbk_fusedx2_srk2_d20_m20_kpw"
    monkeypatch.setenv("SDB_FUSED_VARIANT", "20")
    assert ran(sdb.stratSRS2, f, G, y0, ts, sdb.Jwik, **kw) == "sdbk_fusedx_srk2_d20_m20_wik"
    monkeypatch.delenv("SDB_FUSED_VARIANT")
    f, G, y0 = SYSTEMS["r74"](False)
    assert ran(sdb.itoSRI2, f, G, y0, ts, **kw) == "jit:fusedcjit"                       # commuting B_k
    assert ran(sdb.itoSRI2, f, G, y0, ts, commutative=False, **kw) == "sdbk_fused_srk2_d5_m4_v0"
    f, G, y0 = _commuting_system(10, 10)
    assert ran(sdb.itoSRI2, f, G, y0, ts, **kw) == "sdbk_fusedc_srk2_d10_m10"
    f, G, y0 = sdb.ops.sys_lin(8, 16)
    assert ran(sdb.itoSRI2, f, G, y0, ts, sdb.Iwik, **kw) == "jit:fusedxjit|wik|x2"      # two lanes per path
    f, G, y0 = sdb.ops.sys_lin(6, 4)
    assert ran(sdb.itoSRI2, f, G, y0, ts, **kw) == "jit:fusedjit"
    f, G, y0 = sdb.ops.sys_lin(7, 5)                                                     # odd m
    assert ran(sdb.itoSRI2, f, G, y0, ts, **kw) == "jit:fusedjit"
    assert ran(sdb.itoSRI2, f, G, y0, ts, sdb.Iwik, **kw).startswith("jit:loop|s2|f0|g1|d7|m5|n2")   # TMEM kernels: even m
    f, G, y0 = SYSTEMS["lin2"](False)
    assert ran(sdb.itoEuler, f, G, y0, ts, **kw) == "sdbk_lin2_s0_n1"
    fX, GX = sdb.ops.ExprDrift(["-y[0]"], []), sdb.ops.ExprDiffusion([["0.1*y[0]"]], [])
    assert ran(sdb.itoint, fX, GX, [1.0], ts, **kw).startswith("jit:loop|s2|f9|g9|d1|m1")


def test_device_resident_noise_tensors(sdb):
    """dW / I handed over as torch tensors already on the GPU are used in place (no host round
    trip) and give the same bits as the numpy arrays of the reference's call."""
    import torch
    g = load_golden("integrate_r74.npz")
    f, G, y0 = SYSTEMS["r74"](False)
    tspan, dW, I = g["tspan"], g["dW"][0], g["I"][0]
    ref = sdb.itoSRI2(f, G, y0, tspan, dW=dW, I=I)
    dWd, Id = torch.as_tensor(dW, device="cuda"), torch.as_tensor(I, device="cuda")
    assert np.array_equal(sdb.itoSRI2(f, G, y0, tspan, dW=dWd, I=Id), ref)
    assert np.array_equal(sdb.itoEuler(f, G, y0, tspan, dW=dWd), sdb.itoEuler(f, G, y0, tspan, dW=dW))
    # a strided view (every path's noise, path axis first) is made contiguous, not misread
    stack = torch.stack([dWd, 2.0 * dWd], dim=1)[:, 0]          # non-contiguous (N-1, m)
    assert not stack.is_contiguous()
    assert np.array_equal(sdb.itoEuler(f, G, y0, tspan, dW=stack), sdb.itoEuler(f, G, y0, tspan, dW=dW))


@pytest.mark.parametrize("m", [12, 16, 20])
def test_standalone_repeated_integrals_large_m(sdb, m):
    """Ikpw/Jkpw/Iwik/Jwik with on-device normals at large m go through the TMEM-resident kernel
    (sdb_repintx.cuh; m = 20 ahead of time, others NVRTC): checked against the oracle on the replayed
    Philox normals, for a partial last block, and against the generic kernel."""
    steps, h, P, seed, n = 3, 0.005, 75, 4242, 5
    dW = sdb.deltaW(steps, m, h, paths=P, seed=seed, path_id0=9)
    X, Y = po.kpw_normals(steps, m, n, P, seed, path_id0=9)
    Gt = po.wik_tail_normals(steps, m, n, P, seed, path_id0=9)
    A, I = sdb.Ikpw(dW, h, n=n, seed=seed, path_id0=9, ensemble=True)
    _, J = sdb.Jkpw(dW, h, n=n, seed=seed, path_id0=9, ensemble=True)
    At, Iw = sdb.Iwik(dW, h, n=n, seed=seed, path_id0=9, ensemble=True)
    _, Jw = sdb.Jwik(dW, h, n=n, seed=seed, path_id0=9, ensemble=True)
    M = m * (m - 1) // 2
    assert A.shape == (P, steps, m, m) and At.shape == (P, steps, M, 1)
    for p in (0, 31, 32, P - 1):
        A_o, I_o = orc.ikpw_from_normals(dW[p], h, X[:, p], Y[:, p])
        At_o, Iw_o = orc.iwik_from_normals(dW[p], h, X[:, p], Y[:, p], Gt[p])
        s = np.max(np.abs(I_o))
        assert np.max(np.abs(A[p] - A_o)) <= 1e-12 * s and np.max(np.abs(I[p] - I_o)) <= 1e-12 * s
        assert np.max(np.abs(At[p] - At_o)) <= 1e-12 * s and np.max(np.abs(Iw[p] - Iw_o)) <= 1e-12 * s
        eye = 0.5 * h * np.eye(m)[None]
        assert np.max(np.abs(J[p] - I_o - eye)) <= 1e-12 * s
        assert np.max(np.abs(Jw[p] - Iw_o - eye)) <= 1e-12 * s
    # single-path call in the reference's layout
    A1, I1 = sdb.Ikpw(dW[0], h, n=n, seed=seed, path_id0=9)
    assert A1.shape == (steps, m, m) and np.array_equal(I1, I[0])


def test_pipelined_host_output_matches_single_launch(sdb):
    """Large ensembles written to a caller-owned pinned host array are integrated in chunks with the
    device -> host copies overlapped (sdb_copy2d_d2h_async): same bits as one launch, because the
    Philox streams are keyed by the global path id."""
    import torch
    P = (1 << 19) + 777
    f, G, y0 = SYSTEMS["lin10"](False)
    tspan = np.linspace(0.0, 0.01, 11)
    ref = sdb.itoSRI2(f, G, y0, tspan, paths=P, seed=3, path_id0=10, save_every=5, out="torch").cpu().numpy()
    host = torch.empty((3, 10, P), dtype=torch.float64).pin_memory()
    n0 = sdb.launch_count()
    y = sdb.itoSRI2(f, G, y0, tspan, paths=P, seed=3, path_id0=10, save_every=5, host_out=host)
    assert sdb.launch_count() - n0 >= 4                       # more than one chunk
    assert y.shape == (P, 3, 10) and np.array_equal(y, ref.transpose(2, 0, 1))
    assert sdb.last_status()[0] == 0
    # a pageable result array is filled through the fixed ring of two pinned staging slots
    import sdeint_b200.integrate as integ
    pageable = torch.zeros((3, 10, P), dtype=torch.float64)
    old_slot = integ._RING_SLOT_BYTES
    integ._RING_SLOT_BYTES = 3 * 10 * 8 * (1 << 16)        # force > 8 chunks of 2^16.. paths
    try:
        n0 = sdb.launch_count()
        yp = sdb.itoSRI2(f, G, y0, tspan, paths=P, seed=3, path_id0=10, save_every=5, host_out=pageable)
        assert sdb.launch_count() - n0 >= 8
    finally:
        integ._RING_SLOT_BYTES = old_slot
    assert not pageable.is_pinned() and np.array_equal(yp, ref.transpose(2, 0, 1))
    f2, G2, y02 = SYSTEMS["lin2"](False)
    host2 = torch.empty((2, 2, P), dtype=torch.float64).pin_memory()
    y2 = sdb.itoEuler(f2, G2, y02, tspan, paths=P, seed=4, save_every=10, host_out=host2)
    ref2 = sdb.itoEuler(f2, G2, y02, tspan, paths=P, seed=4, save_every=10, out="torch").cpu().numpy()
    assert np.array_equal(y2, ref2.transpose(2, 0, 1))
    # observers ride along chunk by chunk, each chunk at its global path offset
    specs = [("max", 0), ("first_below", 1, 2.99)]
    _, oref = sdb.itoEuler(f2, G2, y02, tspan, paths=P, seed=4, save_every=10, observe=specs)
    y3, o3 = sdb.itoEuler(f2, G2, y02, tspan, paths=P, seed=4, save_every=10, host_out=host2, observe=specs)
    assert o3.shape == (P, 2) and np.array_equal(o3, oref) and np.array_equal(y3, y2)


def test_welford_moments_chan_merge(sdb):
    """(count, mean, M2) partials with a fixed-order Chan merge: exact where
    E[y^2] - E[y]^2 has lost every digit (|mean| = 1e8 std), identical bits for any split of the
    ensemble into successive batches of whole 4096-path chunks, and the rank-order merge."""
    import torch
    rng = np.random.default_rng(12)
    S, d, P = 2, 3, 5 * 4096 + 123
    y = rng.standard_normal((S, d, P)) * 0.5 + 5e7
    yd = torch.from_numpy(y).cuda()
    w = sdb.ensemble.welford(yd)
    mean, var = sdb.ensemble.welford_mean_var(w)
    assert np.all(w[..., 0].cpu().numpy() == P)
    assert np.max(np.abs(mean.cpu().numpy() - y.mean(axis=-1))) <= 1e-7
    assert np.max(np.abs(var.cpu().numpy() / y.var(axis=-1) - 1.0)) <= 1e-9
    sums, _ = sdb.ensemble.moments(yd)
    _, naive = sdb.ensemble.mean_var(sums, float(P))
    assert np.max(np.abs(naive.cpu().numpy() / y.var(axis=-1) - 1.0)) > 1e-3     # what it replaces
    # successive batches merged on the device == one call, bit for bit (chunk-aligned split)
    cut = 2 * 4096
    w2 = sdb.ensemble.welford(yd[:, :, :cut].contiguous())
    w2 = sdb.ensemble.welford(yd[:, :, cut:].contiguous(), into=w2)
    assert torch.equal(w2, w)
    # per-"rank" partials merged in rank order on the host side of the collective
    parts = [sdb.ensemble.welford(yd[:, :, a:b].contiguous()) for a, b in ((0, cut), (cut, P))]
    merged = sdb.ensemble.chan_merge(parts)
    assert torch.equal(merged[..., 0], w[..., 0])
    # (a different association of the same merges: the means carry 5e7 x 2^-53 of rounding each)
    assert np.max(np.abs((merged[..., 2] / w[..., 2]).cpu().numpy() - 1.0)) <= 1e-9


def test_second_moments_covariance_quantiles(sdb):
    """Ensemble analysis beyond mean/variance: sum_p y_i y_j per saved point on the
    device (deterministic), covariance, and quantiles from the on-device histograms."""
    import torch
    rng = np.random.default_rng(5)
    S, d, P = 3, 7, 20011
    L = rng.standard_normal((d, d)) * 0.4 + np.eye(d)
    y = np.einsum("ij,sjp->sip", L, rng.standard_normal((S, d, P))) + 0.3
    yd = torch.from_numpy(y).cuda()
    cross = sdb.ensemble.second_moments(yd)
    want = np.einsum("sip,sjp->sij", y, y)
    assert np.allclose(cross.cpu().numpy(), want, rtol=1e-12, atol=1e-9)
    assert np.array_equal(sdb.ensemble.second_moments(yd).cpu().numpy(), cross.cpu().numpy())
    sums, counts = sdb.ensemble.moments(yd, hist=(400, -8.0, 8.0))
    cov = sdb.ensemble.covariance(sums, cross, float(P)).cpu().numpy()
    for s in range(S):
        assert np.allclose(cov[s], np.cov(y[s], bias=True), rtol=1e-9, atol=1e-10)
    qs = sdb.ensemble.quantiles(counts, -8.0, 8.0, [0.1, 0.5, 0.9])
    ref = np.quantile(y, [0.1, 0.5, 0.9], axis=-1).transpose(1, 2, 0)
    assert np.max(np.abs(qs - ref)) < 2 * 16.0 / 400
    with pytest.raises(sdb._lib.SdbError):
        sdb.ensemble.second_moments(torch.zeros((1, 17, 8), dtype=torch.float64, device="cuda"))


@pytest.mark.parametrize("name,n", [("lin10", 1), ("lin10", 8), ("r74", 3), ("lin10", 70)])
def test_series_length_in_integrators(sdb, name, n):
    """`n_terms` (the series length n of Ikpw/Iwik, default 5 like the reference) is honoured by the
    fused and the generic kernels alike; n > 64 leaves the fused kernels for the generic loop."""
    f, G, y0 = SYSTEMS[name](False)
    m = G.m
    tspan = np.linspace(0.0, 0.05, 11)
    h = 0.005
    P, seed = 34, 1212
    y = sdb.itoSRI2(f, G, y0, tspan, paths=P, seed=seed, n_terms=n)
    dW = sdb.deltaW(10, m, h, paths=P, seed=seed)
    X, Y = po.kpw_normals(10, m, n, P, seed)
    for p in (0, P - 1):
        _, I = orc.ikpw_from_normals(dW[p], h, X[:, p], Y[:, p])
        assert rel_err(y[p], orc.srk2(f, G, np.atleast_1d(y0), tspan, dW[p], I)) <= TOL
    yw = sdb.itoSRI2(f, G, y0, tspan, sdb.Iwik, paths=P, seed=seed, n_terms=n)
    Gt = po.wik_tail_normals(10, m, n, P, seed)
    _, Iw = orc.iwik_from_normals(dW[0], h, X[:, 0], Y[:, 0], Gt[0])
    assert rel_err(yw[0], orc.srk2(f, G, np.atleast_1d(y0), tspan, dW[0], Iw)) <= TOL


@pytest.mark.parametrize("d,m", [(10, 10), (8, 6), (12, 8), (12, 9)])
def test_fused_euler_heun_linear(sdb, d, m, monkeypatch):
    """itoEuler / stratHeun for linear multiplicative systems on device noise: the DMMA kernel of
    sdb_fusedeh.cuh ((10,10) ahead of time, other shapes NVRTC) against the generic loop kernel and
    the oracle, with a partial last warp and per-path initial states."""
    rng = np.random.default_rng(77 * d + m)
    A = -1.0 * np.eye(d) + 0.3 * rng.standard_normal((d, d)) / np.sqrt(d)
    B = 0.4 * rng.standard_normal((m, d, d)) / np.sqrt(d * m)
    f, G = sdb.ops.LinearDrift(A), sdb.ops.LinearDiffusion(B)
    y0 = rng.standard_normal(d)
    tspan = np.linspace(0.0, 0.5, 51)
    h = 0.01
    P, seed = 77, 5151
    y0p = y0[None, :] + 0.1 * rng.standard_normal((P, d))
    dW = sdb.deltaW(50, m, h, paths=P, seed=seed, path_id0=4)
    for integ, ofn in ((sdb.itoEuler, orc.euler), (sdb.stratHeun, orc.heun)):
        y = integ(f, G, y0, tspan, paths=P, seed=seed, path_id0=4, y0_paths=y0p)
        monkeypatch.setenv("SDB_DISABLE_FUSED", "1")
        y_gen = integ(f, G, y0, tspan, paths=P, seed=seed, path_id0=4, y0_paths=y0p)
        monkeypatch.delenv("SDB_DISABLE_FUSED")
        assert rel_err(y, y_gen) <= TOL
        for p in (0, 31, 32, P - 1):
            assert rel_err(y[p], ofn(f, G, y0p[p], tspan, dW[p])) <= TOL
        yd = integ(f, G, y0, tspan, paths=P, seed=seed, path_id0=4, y0_paths=y0p, save_every=7)
        assert np.array_equal(yd, y[:, ::7])


def test_mid_size_nonlinear_user_system(sdb):
    """A d = 7, m = 5 nonlinear, non-commutative, time-dependent system from formulas (the generic
    thread-per-path kernels built by NVRTC, G_n and I beyond the register file): every scheme against the
    oracle on the same noise, host-supplied and device-drawn."""
    d, m = 7, 5
    drift = ["-1.2*y[%d] + 0.3*sin(y[%d]) + 0.1*y[%d]*y[%d]/(1.0+y[%d]*y[%d]) + 0.05*cos(t)"
             % (i, (i + 1) % d, (i + 2) % d, (i + 3) % d, (i + 2) % d, (i + 2) % d) for i in range(d)]
    rows = [["%g*(%s)" % (0.08 + 0.01 * ((3 * i + k) % 5),
                          ["y[%d]" % ((i + k) % d), "sin(y[%d])+0.5" % ((i + 2 * k) % d),
                           "1.0+0.2*y[%d]*y[%d]" % (i, (k + 1) % d), "cos(y[%d]+t)" % ((i + k + 3) % d),
                           "0.3+tanh(y[%d])" % ((2 * i + k) % d)][(i + k) % 5])
             for k in range(m)] for i in range(d)]
    f, G = sdb.ops.ExprDrift(drift, []), sdb.ops.ExprDiffusion(rows, [])
    y0 = 0.5 + 0.1 * np.arange(d)
    tspan = np.linspace(0.0, 0.25, 101)
    h = 0.0025
    rng = np.random.default_rng(77)
    dW = orc.delta_w(100, m, h, rng)
    _, I = orc.ikpw(dW, h, generator=rng)
    _, J = orc.jwik(dW, h, generator=rng)
    assert rel_err(sdb.itoEuler(f, G, y0, tspan, dW=dW), orc.euler(f, G, y0, tspan, dW)) <= TOL
    assert rel_err(sdb.stratHeun(f, G, y0, tspan, dW=dW), orc.heun(f, G, y0, tspan, dW)) <= TOL
    assert rel_err(sdb.itoSRI2(f, G, y0, tspan, dW=dW, I=I), orc.srk2(f, G, y0, tspan, dW, I)) <= TOL
    assert rel_err(sdb.stratSRS2(f, G, y0, tspan, dW=dW, J=J), orc.srk2(f, G, y0, tspan, dW, J)) <= TOL
    yk = sdb.stratKP2iS(f, G, y0, tspan, dW=dW, J=J, rtol=1e-12)
    assert rel_err(yk, orc.kp2is(f, G, y0, tspan, dW, J, rtol=1e-13, solver="newton")) <= 1e-11
    P, seed = 40, 15
    for meth, integ, strat in ((sdb.Iwik, sdb.itoSRI2, False), (sdb.Jkpw, sdb.stratSRS2, True)):
        y = integ(f, G, y0, tspan, meth, paths=P, seed=seed)
        assert sdb.last_kernel().startswith("jit:loop|s2|f9|g9|d7|m5")
        dWd = sdb.deltaW(100, m, h, paths=P, seed=seed)
        _, IJd = meth(dWd, h, seed=seed, ensemble=True)
        for p in (0, P - 1):
            assert rel_err(y[p], orc.srk2(f, G, y0, tspan, dWd[p], IJd[p])) <= TOL


def test_expression_operators_all_integrators(sdb):
    """f and G given as formulas (ops.ExprDrift / ExprDiffusion -> NVRTC): every integrator against the
    oracle running the SAME formulas as numpy callables; time-dependent, non-commutative noise."""
    f = sdb.ops.ExprDrift(["p[0]*(y[1]-y[0])", "y[0]*(p[1]-y[2])-y[1]", "y[0]*y[1]-p[2]*y[2]"],
                          [10.0, 28.0, 8.0 / 3.0])
    G = sdb.ops.ExprDiffusion([["p[0]*y[0]", "0", "0.1*p[0]*y[1]"],
                               ["0", "p[0]*(y[1]+y[0]/10)", "0"],
                               ["0", "0", "p[0]*y[2]*cos(2*t)"]], [0.1])
    y0 = np.array([1.0, 1.0, 20.0])
    tspan = np.linspace(0.0, 0.2, 201)
    h = 0.001
    rng = np.random.default_rng(31)
    dW = orc.delta_w(200, 3, h, rng)
    _, I = orc.iwik(dW, h, generator=rng)
    _, J = orc.jkpw(dW, h, generator=rng)
    assert rel_err(sdb.itoEuler(f, G, y0, tspan, dW=dW), orc.euler(f, G, y0, tspan, dW)) <= TOL
    assert rel_err(sdb.stratHeun(f, G, y0, tspan, dW=dW), orc.heun(f, G, y0, tspan, dW)) <= TOL
    assert rel_err(sdb.itoSRI2(f, G, y0, tspan, dW=dW, I=I), orc.srk2(f, G, y0, tspan, dW, I)) <= TOL
    assert rel_err(sdb.itoSRI2(f, G.columns(), y0, tspan, dW=dW, I=I), orc.srk2(f, G, y0, tspan, dW, I)) <= TOL
    assert rel_err(sdb.stratSRS2(f, G, y0, tspan, dW=dW, J=J), orc.srk2(f, G, y0, tspan, dW, J)) <= TOL
    yk = sdb.stratKP2iS(f, G, y0, tspan, dW=dW, J=J, rtol=1e-12)
    assert rel_err(yk, orc.kp2is(f, G, y0, tspan, dW, J, rtol=1e-13, solver="newton")) <= 1e-11
    # ensemble on device noise
    P, seed = 33, 8
    y = sdb.itoSRI2(f, G, y0, tspan, paths=P, seed=seed)
    dWd = sdb.deltaW(200, 3, h, paths=P, seed=seed)
    _, Id = sdb.Ikpw(dWd, h, seed=seed, ensemble=True)
    assert rel_err(y[P - 1], orc.srk2(f, G, y0, tspan, dWd[P - 1], Id[P - 1])) <= TOL
