"""Pin the oracle (oracle/sde_oracle.py) against the reference: committed golden
outputs of the unmodified reference (tests/golden, made by oracle/make_golden.py)
and, in the build container, the live reference.  CPU only."""
import numpy as np
import pytest

from conftest import SYSTEMS, load_golden, rel_err
from oracle import sde_oracle as orc

TOL = 1e-12   # the north-star parity bound (relative, fp64)


@pytest.mark.parametrize("name", sorted(SYSTEMS))
def test_integrators_match_reference_golden(name):
    g = load_golden("integrate_%s.npz" % name)
    f_ito, G, _ = SYSTEMS[name](False)
    f_str, _, _ = SYSTEMS[name](True)
    tspan, dW, I, J, y0 = g["tspan"], g["dW"], g["I"], g["J"], g["y0"]
    P = dW.shape[0]
    for p in range(P):
        assert rel_err(orc.euler(f_ito, G, y0, tspan, dW[p]), g["y_itoEuler"][p]) <= TOL
        assert rel_err(orc.heun(f_str, G, y0, tspan, dW[p]), g["y_stratHeun"][p]) <= TOL
        assert rel_err(orc.srk2(f_ito, G, y0, tspan, dW[p], I[p]), g["y_itoSRI2"][p]) <= TOL
        assert rel_err(orc.srk2(f_str, G, y0, tspan, dW[p], J[p]), g["y_stratSRS2"][p]) <= TOL
        if "y_itoSRI2_listG" in g:
            y = orc.srk2(f_ito, G.columns(), y0, tspan, dW[p], I[p])
            assert rel_err(y, g["y_itoSRI2_listG"][p]) <= TOL
        if "y_stratKP2iS" in g:
            for solver in ("fsolve", "newton"):
                # fsolve at the fixture's own xtol reproduces the reference exactly in structure
                # (same MINPACK calls); the Newton solve is exact to 1e-13 and sits within hybrd's
                # termination noise of it (a few 1e-12)
                y = orc.kp2is(f_str, G, y0, tspan, dW[p], J[p], rtol=float(g["kp2is_rtol"]) if
                              solver == "fsolve" else 1e-13, solver=solver)
                assert rel_err(y, g["y_stratKP2iS"][p]) <= 5e-12, solver


def test_wiener_constants():
    g = load_golden("wiener.npz")
    assert abs(orc.a_tail(5) - 0.181322955737115) < 1e-15   # test_wiener.py:90-91
    assert orc.a_tail(5) == float(g["a5"]) and orc.a_tail(3) == float(g["a3"])
    for m in (2, 3, 5):
        assert np.array_equal(orc.perm_P(m), g["P%d" % m])
        assert np.array_equal(orc.sel_K(m), g["K%d" % m])


def test_repeated_integrals_match_reference_golden():
    g = load_golden("wiener.npz")
    for ci, (m, n, N, h) in enumerate(g["cases"]):
        m, n, N = int(m), int(n), int(N)
        pre = "c%d_" % ci
        dW, X, Y, Gt = g[pre + "dW"], g[pre + "X"], g[pre + "Y"], g[pre + "Gt"]
        A, I = orc.ikpw_from_normals(dW, h, X, Y)
        assert np.array_equal(A, g[pre + "A_kpw"])          # bit-identical
        assert np.array_equal(I, g[pre + "I_kpw"])
        _, Jk = orc.jkpw_from_normals(dW, h, X, Y)
        assert np.array_equal(Jk, g[pre + "J_kpw"])
        At, Iw = orc.iwik_from_normals(dW, h, X, Y, Gt)
        assert At.shape == g[pre + "At_wik"].shape
        scale = np.max(np.abs(g[pre + "I_wik"]))
        assert np.max(np.abs(At - g[pre + "At_wik"])) <= 1e-14 * scale
        assert np.max(np.abs(Iw - g[pre + "I_wik"])) <= 1e-14 * scale
        _, Jw = orc.jwik_from_normals(dW, h, X, Y, Gt)
        assert np.max(np.abs(Jw - g[pre + "J_wik"])) <= 1e-14 * scale


def test_helper_conventions():
    """Index conventions of vec/kp/P/K (test_wiener.py:22-87 re-expressed)."""
    rng = np.random.default_rng(5)
    A = np.arange(10 * 3 * 3, dtype=np.float64).reshape(10, 3, 3)
    v = orc.vec(A)
    assert v.shape == (10, 9, 1) and np.array_equal(orc.unvec(v), A)
    assert np.array_equal(v[:, :, 0], A.transpose(0, 2, 1).reshape(10, 9))
    m = 6
    X = rng.standard_normal((20, m, 1))
    Y = rng.standard_normal((20, m, 1))
    XkY = orc.kp(X, Y)
    P0 = orc.perm_P(m)
    for s in range(20):
        assert np.allclose(np.kron(X[s, :, 0], Y[s, :, 0]), XkY[s, :, 0])
        assert np.allclose(P0.dot(XkY[s, :, 0]), orc.kp(Y, X)[s, :, 0])
    assert np.array_equal(P0, P0.T) and np.array_equal(P0.dot(P0), np.eye(m * m))
    Am = rng.standard_normal((4, 3, 2))
    Bm = rng.standard_normal((4, 5, 4))
    AkB = orc.kp2(Am, Bm)
    for s in range(4):
        assert np.allclose(np.kron(Am[s], Bm[s]), AkB[s])
    for q in range(2, 10):                                   # Wiktorsson eq (4.3)
        P0, K0 = orc.perm_P(q), orc.sel_K(q)
        M = q * (q - 1) // 2
        Iq = np.eye(q * q)
        assert np.array_equal(K0.dot(K0.T), np.eye(M))
        assert np.array_equal(K0.dot(P0).dot(K0.T), np.zeros((M, M)))
        lhs = (Iq - P0).dot(K0.T).dot(K0).dot(Iq - P0)
        assert np.array_equal(lhs, Iq - P0)


def test_identities_of_repeated_integrals():
    """Wiktorsson (2.1) identities, test_wiener.py:94-130, on oracle output."""
    rng = np.random.default_rng(77)
    N, h, m = 2000, 0.002, 8
    dW = orc.delta_w(N, m, h, rng)
    outer = dW[:, :, None] * dW[:, None, :]
    eye = np.eye(m)[None]
    A, I = orc.ikpw(dW, h, generator=rng)
    assert A.shape == (N, m, m) and I.shape == (N, m, m)
    assert np.allclose(I + I.transpose(0, 2, 1), outer - h * eye)
    assert np.allclose(A, -A.transpose(0, 2, 1))
    assert np.allclose(2.0 * (I - A), outer - h * eye)
    _, J = orc.jkpw(dW, h, generator=rng)
    assert np.allclose(J + J.transpose(0, 2, 1), outer)
    At, Iw = orc.iwik(dW, h, generator=rng)
    M = m * (m - 1) // 2
    assert At.shape == (N, M, 1)
    assert np.allclose(Iw + Iw.transpose(0, 2, 1), outer - h * eye)
    K0, P0 = orc.sel_K(m), orc.perm_P(m)
    Afull = orc.unvec(np.einsum("ab,sbc->sac", (np.eye(m * m) - P0).dot(K0.T), At))
    assert np.allclose(Afull, -Afull.transpose(0, 2, 1))
    assert np.allclose(2.0 * (Iw - Afull), outer - h * eye)


def test_oracle_against_live_reference(reference):
    """In the build container: same seeded generator through the reference and
    through the oracle's drop-in wrappers (draw order included)."""
    sdeint = reference
    for m, n in ((1, 5), (4, 5), (7, 3)):
        N, h = 64, 0.003
        dW = sdeint.deltaW(N, m, h, np.random.default_rng(3))
        for ref_fn, orc_fn in ((sdeint.Ikpw, orc.ikpw), (sdeint.Jkpw, orc.jkpw),
                               (sdeint.Iwik, orc.iwik), (sdeint.Jwik, orc.jwik)):
            a = ref_fn(dW, h, n=n, generator=np.random.default_rng(9))
            b = orc_fn(dW, h, n=n, generator=np.random.default_rng(9))
            assert a[0].shape == b[0].shape and a[1].shape == b[1].shape
            assert np.max(np.abs(a[0] - b[0])) <= 1e-16 + 1e-13 * np.max(np.abs(a[1]))
            assert np.max(np.abs(a[1] - b[1])) <= 1e-13 * np.max(np.abs(a[1]))
    # a full default call: deltaW followed by Ikpw with the same generator
    f, G, y0 = SYSTEMS["lin10"](False)
    tspan = np.linspace(0, 0.05, 51)
    y_ref = sdeint.itoSRI2(f, G.columns(), y0, tspan, generator=np.random.default_rng(21))
    g = np.random.default_rng(21)
    hh = (tspan[-1] - tspan[0]) / 50
    dW = orc.delta_w(50, 10, hh, g)
    _, I = orc.ikpw(dW, hh, generator=g)
    assert rel_err(orc.srk2(f, G.columns(), y0, tspan, dW, I), y_ref) <= TOL


def test_oracle_matches_reference_on_python_callables():
    """tests/golden/callables.npz: the unmodified reference run on plain Python f, G (tests/callable_systems.py:
    np.maximum, np.where, masks, preallocated results, np.dot, list-of-columns G); the oracle must reproduce it."""
    import callable_systems as cs
    g = load_golden("callables.npz")
    for name, make in cs.SYSTEMS.items():
        made = make()
        f, G, y0 = made[:3]
        ts, dW, I, J = (g["%s_%s" % (name, k)] for k in ("tspan", "dW", "I", "J"))
        for p in range(dW.shape[0]):
            assert rel_err(orc.euler(f, G, y0, ts, dW[p]), g[name + "_y_itoEuler"][p]) <= 1e-13
            assert rel_err(orc.heun(f, G, y0, ts, dW[p]), g[name + "_y_stratHeun"][p]) <= 1e-13
            assert rel_err(orc.srk2(f, G, y0, ts, dW[p], I[p]), g[name + "_y_itoSRI2"][p]) <= 1e-13
            assert rel_err(orc.srk2(f, G, y0, ts, dW[p], J[p]), g[name + "_y_stratSRS2"][p]) <= 1e-13
            if len(made) > 3:
                assert rel_err(orc.srk2(f, made[3], y0, ts, dW[p], I[p]), g[name + "_y_itoSRI2_listG"][p]) <= 1e-13
