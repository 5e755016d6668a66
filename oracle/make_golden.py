"""Generate tests/golden/*.npz by running the UNMODIFIED reference (mattja/sdeint
at /root/reference) on seeded inputs.  Test infrastructure only.

Run here (the build container, where /root/reference exists):

    python oracle/make_golden.py

The fixtures hold inputs AND the reference's outputs, so the GPU box (which has
no /root/reference) can check both the oracle restatement and the CUDA path
against the real reference.  Fixtures are deliberately small (few paths, few
hundred steps); larger parity cases compare CUDA with the pinned oracle.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import sdeint  # noqa: E402  (the reference)
from sdeint import wiener as ref_wiener  # noqa: E402

from sdeint_b200 import ops  # noqa: E402
from oracle import sde_oracle as orc  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _noise(seed, P, N, m, h, method):
    """Per path: dW = deltaW, (.., I) = Ikpw/Iwik with the same generator, like
    sdeint/integrate.py:483-487 would do inside one call."""
    dW = np.empty((P, N - 1, m))
    I = np.empty((P, N - 1, m, m))
    for p in range(P):
        g = np.random.default_rng([seed, p])
        dW[p] = sdeint.deltaW(N - 1, m, h, g)
        I[p] = method(dW[p], h, generator=g)[1]
    J = I + 0.5 * h * np.eye(m).reshape(1, 1, m, m)
    return dW, I, J


def integrator_fixture(name, make_sys, P, N, T, seed, method=sdeint.Ikpw, list_g=True,
                       kp2is=True):
    f_ito, G, y0 = make_sys(False)
    f_str, _, _ = make_sys(True)
    d, m = G.d, G.m
    tspan = np.linspace(0.0, T, N)
    h = (tspan[-1] - tspan[0]) / (N - 1)
    dW, I, J = _noise(seed, P, N, m, h, method)
    y0v = np.array([y0]) if np.ndim(y0) == 0 else np.asarray(y0, dtype=float)

    def run(fn, *a, **kw):
        return np.stack([fn(*a, **{k: (v[p] if isinstance(v, np.ndarray) and v.ndim >= 3
                                           else v) for k, v in kw.items()})
                         for p in range(P)])

    out = dict(tspan=tspan, dW=dW, I=I, J=J, y0=y0v, d=d, m=m)
    out["y_itoEuler"] = run(sdeint.itoEuler, f_ito, G, y0v, tspan, dW=dW)
    out["y_stratHeun"] = run(sdeint.stratHeun, f_str, G, y0v, tspan, dW=dW)
    out["y_itoSRI2"] = run(sdeint.itoSRI2, f_ito, G, y0v, tspan, dW=dW, I=I)
    out["y_stratSRS2"] = run(sdeint.stratSRS2, f_str, G, y0v, tspan, dW=dW, J=J)
    if list_g:
        out["y_itoSRI2_listG"] = run(sdeint.itoSRI2, f_ito, G.columns(), y0v, tspan,
                                     dW=dW, I=I)
    if kp2is:
        # The tightest MINPACK tolerance the reference completes with on every path: fsolve
        # stalls ("not making good progress" -> RuntimeError, integrate.py:640-645) below
        # ~1e-11 on some steps of some systems.  hybrd's termination leaves the solution of
        # the implicit equation good to a few 1e-12 relative whatever xtol says (measured
        # against a Newton solve to 1e-13), which is the tolerance the parity tests state.
        for rtol in (1e-12, 1e-11, 1e-10, 1e-9):
            try:
                out["y_stratKP2iS"] = run(sdeint.stratKP2iS, f_str, G, y0v, tspan, dW=dW, J=J,
                                          rtol=rtol)
            except RuntimeError:
                continue
            out["kp2is_rtol"] = np.array(rtol)
            break
    np.savez(os.path.join(OUT, "integrate_%s.npz" % name), **out)
    print("wrote integrate_%s.npz  P=%d N=%d d=%d m=%d" % (name, P, N, d, m))


def wiener_fixture():
    out = {}
    out["a5"] = np.array(ref_wiener._a(5))
    out["a3"] = np.array(ref_wiener._a(3))
    for m in (2, 3, 5):
        out["P%d" % m] = ref_wiener._P(m)
        out["K%d" % m] = ref_wiener._K(m)
    cases = [(1, 5, 40, 0.002), (2, 5, 40, 0.002), (3, 3, 30, 0.01), (8, 5, 25, 0.002),
             (10, 5, 12, 0.001)]
    out["cases"] = np.array(cases, dtype=np.float64)
    for ci, (m, n, N, h) in enumerate(cases):
        seed = 4242 + ci
        g = np.random.default_rng(seed)
        dW = sdeint.deltaW(N, m, h, g)
        # reference results, each from an identically seeded generator state
        state = g.bit_generator.state
        A, I = sdeint.Ikpw(dW, h, n=n, generator=g)
        g.bit_generator.state = state
        _, Jk = sdeint.Jkpw(dW, h, n=n, generator=g)
        g.bit_generator.state = state
        At, Iw = sdeint.Iwik(dW, h, n=n, generator=g)
        g.bit_generator.state = state
        _, Jw = sdeint.Jwik(dW, h, n=n, generator=g)
        # the normals those calls consumed, replayed in the reference's draw order
        g.bit_generator.state = state
        X, Y, Gt = orc.draw_wik_normals(g, N, m, n)
        if m == 1:
            g.bit_generator.state = state
            X, Y = orc.draw_kpw_normals(g, N, m, n)
        pre = "c%d_" % ci
        out.update({pre + "dW": dW, pre + "X": X, pre + "Y": Y, pre + "Gt": Gt,
                    pre + "A_kpw": A, pre + "I_kpw": I, pre + "J_kpw": Jk,
                    pre + "At_wik": At, pre + "I_wik": Iw, pre + "J_wik": Jw})
    np.savez(os.path.join(OUT, "wiener.npz"), **out)
    print("wrote wiener.npz")


def callables_fixture():
    """Plain Python callables (tests/callable_systems.py) through the unmodified reference: what sdeint_b200
    must reproduce when it TRACES the same objects into device code."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import callable_systems as cs
    out = {}
    for ci, (name, make) in enumerate(sorted(cs.SYSTEMS.items())):
        made = make()
        f, G, y0 = made[:3]
        m = np.asarray(G(y0, 0.0)).shape[1]
        P, N, T = 2, 201, 0.5
        tspan = np.linspace(0.0, T, N)
        h = T / (N - 1)
        dW, I, J = _noise(900 + ci, P, N, m, h, sdeint.Ikpw)
        pre = name + "_"
        out.update({pre + "tspan": tspan, pre + "dW": dW, pre + "I": I, pre + "J": J, pre + "y0": y0})
        out[pre + "y_itoEuler"] = np.stack([sdeint.itoEuler(f, G, y0, tspan, dW=dW[p]) for p in range(P)])
        out[pre + "y_stratHeun"] = np.stack([sdeint.stratHeun(f, G, y0, tspan, dW=dW[p]) for p in range(P)])
        out[pre + "y_itoSRI2"] = np.stack([sdeint.itoSRI2(f, G, y0, tspan, dW=dW[p], I=I[p]) for p in range(P)])
        out[pre + "y_stratSRS2"] = np.stack([sdeint.stratSRS2(f, G, y0, tspan, dW=dW[p], J=J[p]) for p in range(P)])
        if len(made) > 3:     # the list-of-columns form of G (integrate.py:330-334)
            out[pre + "y_itoSRI2_listG"] = np.stack([sdeint.itoSRI2(f, made[3], y0, tspan, dW=dW[p], I=I[p])
                                                     for p in range(P)])
    np.savez(os.path.join(OUT, "callables.npz"), **out)
    print("wrote callables.npz  (%s)" % ", ".join(sorted(cs.SYSTEMS)))


def main():
    os.makedirs(OUT, exist_ok=True)
    if "--callables-only" in sys.argv:
        callables_fixture()
        return
    wiener_fixture()
    callables_fixture()
    integrator_fixture("kp4446", lambda s: ops.sys_kp4446(s), P=3, N=251, T=1.0, seed=11,
                       method=sdeint.Iwik, list_g=False)
    integrator_fixture("kp4459", lambda s: ops.sys_kp4459(s), P=3, N=251, T=0.5, seed=12,
                       method=sdeint.Iwik)
    integrator_fixture("kps445", lambda s: ops.sys_kps445(s), P=3, N=251, T=1.0, seed=13,
                       method=sdeint.Iwik, list_g=False)
    integrator_fixture("r74", lambda s: ops.sys_r74(s), P=3, N=251, T=1.0, seed=14,
                       method=sdeint.Iwik)
    integrator_fixture("lin2", lambda s: ops.sys_lin2(), P=3, N=251, T=2.5, seed=15)
    integrator_fixture("lin10", lambda s: ops.sys_lin(10, 10), P=2, N=101, T=0.1, seed=16)
    integrator_fixture("ou", lambda s: tuple(x if i < 2 else np.array([0.3])
                                             for i, x in enumerate(ops.sys_ou())),
                       P=2, N=201, T=2.0, seed=17)


if __name__ == "__main__":
    main()
