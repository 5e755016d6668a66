"""Drift / diffusion operators understood by the CUDA integrators.

The reference takes arbitrary Python callables `f(y, t)` and `G(y, t)`
(sdeint/integrate.py:124-146).  A Python callable cannot run inside a CUDA
kernel, so on this path `f` and `G` are OPERATOR OBJECTS: either one of the
built-in families below (linear `A.y`, constant `B`, linear columns `B_k.y`,
and the exactly solvable test equations of sdeint/tests/test_integrate.py) or
CUDA device-function source compiled at run time with NVRTC for sm_100a
(`CudaDrift`, `CudaDiffusion`).

Every operator is ALSO a plain Python callable with the reference's calling
convention, `f(y, t) -> (d,)`, `G(y, t) -> (d, m)`, and a diffusion can be
split into its m column callables with `.columns()` (the "list of g_k" form of
integrate.py:330-334).  That host-side evaluation exists for argument
validation (`_check_args` probes `f(y0, t0)` exactly like integrate.py:71-107)
and so that the very same object can be handed to the reference / the oracle in
tests.  It is never used to integrate: there is no CPU fallback.

Kind ids and parameter layouts are shared with `include/sdb.h`.
"""
from __future__ import annotations

import numbers

import numpy as np

# ---- kind ids (must match include/sdb.h) ---------------------------------------------
DRIFT_LINEAR = 0      # params: A[d][d] row-major                      f = A y
DRIFT_KP4446 = 1      # params: a, c          f = -(a + c*y)(1 - y^2)  (c=b^2 Ito, 0 Strat)
DRIFT_KPS445 = 2      # params: s             f = -sin2y - sin4y/4 + s*2 sin y cos^3 y
DRIFT_USER = 9        # NVRTC source

DIFF_CONST = 0        # params: B[d][m] row-major                      G = B
DIFF_LINEAR_COLS = 1  # params: B[m][d][d]                             g_k = B_k y
DIFF_KP4446 = 2       # params: b                                      G = b (1 - y^2)
DIFF_KPS445 = 3       # params: (none; one dummy)                      G = sqrt2 cos^2 y
DIFF_DIAG_LINEAR = 4  # params: b[d]                                   g_k = b_k y_k e_k (m = d)
DIFF_AFFINE_COLS = 5  # params: B[m][d][d], C[d][m]                    g_k = B_k y + c_k
DIFF_USER = 9
DIFF_USER_DIAG = 10   # NVRTC source, diagonal noise promised by the caller


def _is_scalar(y):
    return isinstance(y, numbers.Number) or np.ndim(y) == 0


class DriftOp:
    """Base class: a drift operator with a device descriptor."""
    kind = -1
    d = 0
    params = np.zeros(0)
    source = None          # CUDA source for DRIFT_USER

    @property
    def __name__(self):    # the reference's scalar wrappers copy fn.__name__ (integrate.py:64)
        return type(self).__name__

    def key(self):
        return ("f", self.kind, self.d, self.source)


class DiffusionOp:
    """Base class: a diffusion operator (d x m) with a device descriptor."""
    kind = -1
    d = 0
    m = 0
    params = np.zeros(0)
    source = None
    # Commutative noise (L^j g_k = L^k g_j for all j, k): the Levy areas cancel in the order-1.0
    # term, so SRI2/SRS2 can run on I = (dW dW^T - h 1)/2 alone.  LinearDiffusion decides it from
    # the commutators of its B_k; user operators may promise it (`commutative=True`).
    commutative = False

    @property
    def __name__(self):
        return type(self).__name__

    def key(self):
        return ("G", self.kind, self.d, self.m, self.source)

    def column(self, k, y, t):
        return np.asarray(self(y, t))[:, k]

    def columns(self):
        """The m column callables g_k(y, t) -> (d,) (integrate.py:330-334)."""
        return [DiffusionColumn(self, k) for k in range(self.m)]


class DiffusionColumn:
    """Column k of a DiffusionOp; a list of these is recognised by the shim as
    the list-of-columns form of the same device operator."""

    def __init__(self, parent, k):
        self.parent = parent
        self.k = int(k)
        self.__name__ = "%s_col%d" % (type(parent).__name__, k)

    def __call__(self, y, t):
        return self.parent.column(self.k, y, t)


# ---- built-ins -----------------------------------------------------------------------

class LinearDrift(DriftOp):
    """f(y, t) = A y."""
    kind = DRIFT_LINEAR

    def __init__(self, A):
        A = np.atleast_2d(np.asarray(A, dtype=np.float64))
        if A.ndim != 2 or A.shape[0] != A.shape[1]:
            raise ValueError("LinearDrift needs a square matrix A")
        self.A = np.ascontiguousarray(A)
        self.d = A.shape[0]
        self.params = self.A.reshape(-1).copy()

    def __call__(self, y, t):
        if _is_scalar(y) and self.d == 1:
            return float(self.A[0, 0]) * y
        return self.A.dot(y)


class ConstDiffusion(DiffusionOp):
    """G(y, t) = B (additive noise)."""
    kind = DIFF_CONST

    def __init__(self, B):
        B = np.asarray(B, dtype=np.float64)
        if B.ndim == 0:
            B = B.reshape(1, 1)
        if B.ndim != 2:
            raise ValueError("ConstDiffusion needs a (d, m) matrix B")
        self.B = np.ascontiguousarray(B)
        self.d, self.m = B.shape
        self.params = self.B.reshape(-1).copy()

    def __call__(self, y, t):
        if _is_scalar(y) and self.d == 1 and self.m == 1:
            return float(self.B[0, 0])
        return self.B.copy()

    def column(self, k, y, t):
        return self.B[:, k].copy()


class LinearDiffusion(DiffusionOp):
    """g_k(y, t) = B_k y for k = 0..m-1, B of shape (m, d, d)."""
    kind = DIFF_LINEAR_COLS

    def __init__(self, B):
        B = np.asarray(B, dtype=np.float64)
        if B.ndim != 3 or B.shape[1] != B.shape[2]:
            raise ValueError("LinearDiffusion needs B of shape (m, d, d)")
        self.B = np.ascontiguousarray(B)
        self.m, self.d = B.shape[0], B.shape[1]
        self.params = self.B.reshape(-1).copy()
        self._commutative = None

    @property
    def commutative(self):
        """True when every pair of B_k commutes EXACTLY in floating point (both products are
        evaluated with the same numpy matmul): then sum_{j,k} B_k B_j y I_jk does not see the
        antisymmetric part of I, and SRI2/SRS2 are unchanged -- to rounding, pathwise for the same
        dW -- if the Levy areas are never generated."""
        if self._commutative is None:
            ok = True
            for j in range(self.m):
                for k in range(j + 1, self.m):
                    c = self.B[j] @ self.B[k] - self.B[k] @ self.B[j]
                    scale = np.abs(self.B[j]).max() * np.abs(self.B[k]).max() * self.d
                    if np.abs(c).max() > 8 * np.finfo(float).eps * scale:
                        ok = False
                        break
                if not ok:
                    break
            self._commutative = ok
        return self._commutative

    def __call__(self, y, t):
        if _is_scalar(y) and self.d == 1 and self.m == 1:
            return float(self.B[0, 0, 0]) * y
        return np.dot(self.B, y).T

    def column(self, k, y, t):
        return self.B[k].dot(y)


class AffineDiffusion(DiffusionOp):
    """g_k(y, t) = B_k y + c_k: multiplicative and additive noise together (B of shape (m, d, d), C of shape
    (d, m) with columns c_k).  The stage-2 differences of SRI2/SRS2 do not see the constants, so the fused
    tensor-instruction kernels apply as for LinearDiffusion (sdb_fused.cuh, SDB_FUSED_AFFINE)."""
    kind = DIFF_AFFINE_COLS

    def __init__(self, B, C):
        B = np.asarray(B, dtype=np.float64)
        C = np.asarray(C, dtype=np.float64)
        if B.ndim != 3 or B.shape[1] != B.shape[2] or C.shape != (B.shape[1], B.shape[0]):
            raise ValueError("AffineDiffusion needs B of shape (m, d, d) and C of shape (d, m)")
        self.B, self.C = np.ascontiguousarray(B), np.ascontiguousarray(C)
        self.m, self.d = B.shape[0], B.shape[1]
        self.params = np.concatenate([self.B.reshape(-1), self.C.reshape(-1)])

    def __call__(self, y, t):
        return np.dot(self.B, y).T + self.C

    def column(self, k, y, t):
        return self.B[k].dot(y) + self.C[:, k]


class DiagLinearDiffusion(DiffusionOp):
    """Diagonal multiplicative noise g_k(y, t) = b_k y_k e_k (m = d).  For diagonal noise
    SRI2/SRS2/KP2iS need only I_kk = (dW_k^2 - h)/2, so no Levy areas are generated: the first
    "fast path" of the reference's to-do list (README.rst:130, integrate.py:150-151)."""
    kind = DIFF_DIAG_LINEAR

    def __init__(self, b):
        self.b = np.asarray(b, dtype=np.float64).reshape(-1).copy()
        self.d = self.m = len(self.b)
        self.params = self.b.copy()

    def __call__(self, y, t):
        return np.diag(self.b * np.asarray(y, dtype=np.float64))


class KP4446Drift(DriftOp):
    """Kloeden & Platen (4.46) drift, sdeint/tests/test_integrate.py:93-99.
    Ito: -(a + y b^2)(1 - y^2); Stratonovich: -a (1 - y^2)."""
    kind = DRIFT_KP4446
    d = 1

    def __init__(self, a=1.0, b=0.8, stratonovich=False):
        self.a, self.b, self.stratonovich = float(a), float(b), bool(stratonovich)
        c = 0.0 if stratonovich else self.b ** 2
        self.params = np.array([self.a, c], dtype=np.float64)

    def __call__(self, y, t):
        a, c = self.params
        return -(a + y * c) * (1 - y ** 2)


class KP4446Diffusion(DiffusionOp):
    """G = b (1 - y^2), test_integrate.py:95-96."""
    kind = DIFF_KP4446
    d = 1
    m = 1

    def __init__(self, b=0.8):
        self.b = float(b)
        self.params = np.array([self.b], dtype=np.float64)

    def __call__(self, y, t):
        if _is_scalar(y):
            return self.b * (1 - y ** 2)
        return (self.b * (1 - np.asarray(y) ** 2)).reshape(1, 1)

    def column(self, k, y, t):
        return self.b * (1 - np.asarray(y) ** 2)


class KPS445Drift(DriftOp):
    """Kloeden, Platen & Schurz (4.5) drift, test_integrate.py:263-269."""
    kind = DRIFT_KPS445
    d = 1

    def __init__(self, stratonovich=False):
        self.stratonovich = bool(stratonovich)
        self.params = np.array([1.0 if stratonovich else 0.0], dtype=np.float64)

    def __call__(self, y, t):
        out = -np.sin(2 * y) - 0.25 * np.sin(4 * y)
        if self.stratonovich:
            out = out + 2.0 * np.sin(y) * np.cos(y) ** 3
        return out


class KPS445Diffusion(DiffusionOp):
    """G = sqrt(2) cos^2 y, test_integrate.py:265-266."""
    kind = DIFF_KPS445
    d = 1
    m = 1

    def __init__(self):
        self.params = np.array([np.sqrt(2.0)], dtype=np.float64)

    def __call__(self, y, t):
        if _is_scalar(y):
            return np.sqrt(2.0) * np.cos(y) ** 2
        return (np.sqrt(2.0) * np.cos(np.asarray(y)) ** 2).reshape(1, 1)

    def column(self, k, y, t):
        return np.sqrt(2.0) * np.cos(np.asarray(y)) ** 2


# ---- user CUDA source (NVRTC) --------------------------------------------------------

class CudaDrift(DriftOp):
    """Drift given as CUDA source.  The source must define

        __device__ void sde_drift(const double* y, double t,
                                  const double* p, double* out);   // out[d]

    `params` are passed as `p`.  `host` is an optional Python callable with the
    reference's signature, used only for argument validation and tests."""
    kind = DRIFT_USER

    def __init__(self, source, d, params=(), host=None):
        self.source = str(source)
        self.d = int(d)
        self.params = np.asarray(params, dtype=np.float64).reshape(-1).copy()
        self.host = host

    def __call__(self, y, t):
        if self.host is None:
            return np.zeros(self.d)        # shape probe only
        return self.host(y, t)


class CudaDiffusion(DiffusionOp):
    """Diffusion given as CUDA source.  The source must define

        __device__ void sde_diffusion_col(int k, const double* y, double t,
                                          const double* p, double* out);  // out[d]

    i.e. column k of G (the integrators only ever need single columns,
    integrate.py:516-521)."""
    kind = DIFF_USER

    def __init__(self, source, d, m, params=(), host=None, diagonal=False, commutative=False):
        self.source = str(source)
        self.d, self.m = int(d), int(m)
        self.params = np.asarray(params, dtype=np.float64).reshape(-1).copy()
        self.host = host
        self.commutative = bool(commutative)   # promise: L^j g_k = L^k g_j (no Levy areas needed)
        if diagonal:   # promise: column k has only component k and depends only on y_k
            if self.d != self.m:
                raise ValueError("diagonal noise needs m == d")
            self.kind = DIFF_USER_DIAG

    def __call__(self, y, t):
        if self.host is None:
            return np.zeros((self.d, self.m))
        return self.host(y, t)


# ---- drift / diffusion from arithmetic expressions -------------------------------------
# The reference takes Python callables; a CUDA kernel cannot.  For the common case of closed-form
# f and G these two classes take the formulas as strings -- one per component, in the state
# y[0..d-1], the time t and the parameters p[0..] with + - * / and sin cos tan exp log sqrt pow
# fabs tanh -- and generate both the CUDA source (compiled by NVRTC) and the equivalent numpy
# callable handed to the oracle in tests, so no CUDA has to be written by hand.

_EXPR_FUNCS = ("sin", "cos", "tan", "exp", "log", "sqrt", "pow", "fabs", "tanh", "sinh", "cosh", "fmax", "fmin",
               "atan", "asin", "acos", "atan2", "hypot", "log1p", "expm1", "log10", "log2", "cbrt",
               "lt", "le", "gt", "ge")   # comparisons as 0.0 / 1.0 (masks for piecewise formulas)
# device definitions of the names that are not C math functions (guarded: drift and diffusion sources of one
# system are compiled as one translation unit)
_EXPR_PRELUDE = (
    "#ifndef SDB_EXPR_PRELUDE\n#define SDB_EXPR_PRELUDE\n"
    "__device__ __forceinline__ double lt(double a, double b) { return a < b ? 1.0 : 0.0; }\n"
    "__device__ __forceinline__ double le(double a, double b) { return a <= b ? 1.0 : 0.0; }\n"
    "__device__ __forceinline__ double gt(double a, double b) { return a > b ? 1.0 : 0.0; }\n"
    "__device__ __forceinline__ double ge(double a, double b) { return a >= b ? 1.0 : 0.0; }\n"
    "#endif\n")
_EXPR_PY = {"fabs": np.abs, "pow": np.power, "atan": np.arctan, "asin": np.arcsin, "acos": np.arccos,
            "atan2": np.arctan2,
            "lt": lambda a, b: float(a < b), "le": lambda a, b: float(a <= b),
            "gt": lambda a, b: float(a > b), "ge": lambda a, b: float(a >= b)}


def _prelude(exprs):
    import re
    used = set(re.findall(r"[A-Za-z_]\w*", " ".join(exprs)))
    return _EXPR_PRELUDE if used & {"lt", "le", "gt", "ge"} else ""


def _check_expr(expr):
    import re
    expr = str(expr)
    # (numeric literals first: the e of 1.5e-06 is not a name; an identifier such as y2 keeps its digits)
    bare = re.sub(r"(?<![\w.])(?:\d+\.?\d*|\.\d+)(?:[eE][+-]?\d+)?", " ", expr)
    for name in re.findall(r"[A-Za-z_]\w*", bare):
        if name not in _EXPR_FUNCS and name not in ("y", "t", "p"):
            raise ValueError("unknown name %r in expression %r (allowed: y[i], t, p[i], %s)"
                             % (name, expr, ", ".join(_EXPR_FUNCS)))
    if ";" in expr or "{" in expr or "}" in expr:
        raise ValueError("expression %r must be a single arithmetic expression" % expr)
    return expr


def _c_expr(expr):
    """The expression as C: bare integer literals become doubles (1/3 must not be integer division);
    subscripts and exponents of floating literals stay as they are."""
    import re
    out, depth, pos = [], 0, 0
    number = re.compile(r"(?:\d+\.?\d*|\.\d+)(?:[eE][+-]?\d+)?")
    while pos < len(expr):
        ch = expr[pos]
        if ch.isalpha() or ch == "_":                 # identifier (may contain digits): copy whole
            mt = re.match(r"\w+", expr[pos:])
            out.append(mt.group(0))
            pos += mt.end()
            continue
        mt = number.match(expr, pos)
        if mt:
            tok = mt.group(0)
            out.append(tok + ".0" if depth == 0 and tok.isdigit() else tok)
            pos = mt.end()
            continue
        depth += ch == "["
        depth -= ch == "]"
        out.append(ch)
        pos += 1
    return "".join(out)


def _expr_eval(expr, y, t, p):
    env = {k: _EXPR_PY[k] if k in _EXPR_PY else getattr(np, k) for k in _EXPR_FUNCS}
    env.update(y=y, t=t, p=p)
    return float(eval(expr, {"__builtins__": {}}, env))      # names were whitelisted by _check_expr


class ExprDrift(CudaDrift):
    """f_i(y, t) = the i-th expression, e.g. ExprDrift(["p[0]*(y[1]-y[0])", "y[0]*(p[1]-y[2])-y[1]",
    "y[0]*y[1]-p[2]*y[2]"], params=[10, 28, 8/3])."""

    def __init__(self, exprs, params=()):
        self.exprs = [_check_expr(e) for e in exprs]
        body = "\n".join("  out[%d] = %s;" % (i, _c_expr(e)) for i, e in enumerate(self.exprs))
        src = (_prelude(self.exprs) +
               "__device__ void sde_drift(const double* y, double t, const double* p, double* out) {\n"
               "%s\n}\n" % body)
        par = np.asarray(params, dtype=np.float64).reshape(-1)
        super().__init__(src, len(self.exprs), par,
                         host=lambda y, t: np.array([_expr_eval(e, np.asarray(y, float), t, par)
                                                     for e in self.exprs]))


class ExprDiffusion(CudaDiffusion):
    """G_ik(y, t) = exprs[i][k] (d rows of m expressions; "0" for structural zeros).
    `diagonal=True` promises diagonal noise, `commutative=True` commutative noise (see CudaDiffusion)."""

    def __init__(self, exprs, params=(), diagonal=False, commutative=False):
        rows = [[_check_expr(e) for e in row] for row in exprs]
        d, m = len(rows), len(rows[0])
        if any(len(r) != m for r in rows):
            raise ValueError("ExprDiffusion needs a d x m table of expressions")
        self.exprs = rows
        cases = []
        for k in range(m):
            lines = "\n".join("      out[%d] = %s;" % (i, _c_expr(rows[i][k])) for i in range(d))
            cases.append("    case %d:\n%s\n      break;" % (k, lines))
        src = (_prelude([e for row in rows for e in row]) +
               "__device__ void sde_diffusion_col(int k, const double* y, double t, const double* p, "
               "double* out) {\n  switch (k) {\n%s\n  }\n}\n" % "\n".join(cases))
        par = np.asarray(params, dtype=np.float64).reshape(-1)
        super().__init__(src, d, m, par, diagonal=diagonal, commutative=commutative,
                         host=lambda y, t: np.array([[_expr_eval(rows[i][k], np.asarray(y, float), t, par)
                                                      for k in range(m)] for i in range(d)]))


def as_linear(op, dense_only=False):
    """An `ExprDrift` / `ExprDiffusion` whose expressions are exactly linear in y and do not depend on t,
    as the equivalent `LinearDrift` / `LinearDiffusion` -- those have tensor-instruction kernels (DESIGN.md
    section 4) -- else the operator itself.  The matrices are read off at the unit vectors; linearity is
    then checked at random points and two times (a nonlinear or affine expression fails it by many orders
    of magnitude more than rounding).  `dense_only`: keep a drift with fewer than d^2 / 2 nonzero
    coefficients as source (evaluated per lane it is cheaper than rows of a dense product).  Diffusions
    flagged `diagonal=True` keep their own fast path.  Cached on the operator; SDB_NO_PROMOTE=1 disables."""
    import os
    if os.environ.get("SDB_NO_PROMOTE") or not isinstance(op, (ExprDrift, ExprDiffusion)):
        return op
    key = "_as_linear_dense" if dense_only else "_as_linear"
    if key in op.__dict__:
        return op.__dict__[key]
    # operators traced from Python callables are new objects on every call: remember the verdict by content
    if isinstance(op, ExprDrift):
        ckey = (key, "f", tuple(op.exprs), op.params.tobytes())
    else:
        ckey = (key, "G", tuple(tuple(r) for r in op.exprs), op.params.tobytes(), op.kind)
    hit = _AS_LINEAR_CACHE.get(ckey)
    if hit is not None:
        res = op if hit is False else hit
        op.__dict__[key] = res
        return res
    d = op.d
    res = op
    try:
        rng = np.random.default_rng(12345)
        eye = np.eye(d)
        if isinstance(op, ExprDrift):
            A = np.stack([np.asarray(op(eye[l], 0.0), dtype=np.float64) for l in range(d)], axis=1)   # (d, d)
            lin = lambda y: A @ y
            scale = max(1.0, float(np.abs(A).max()))
            ok = np.abs(np.asarray(op(np.zeros(d), 0.0), dtype=np.float64)).max() == 0.0
            cand = None
            if ok and not (dense_only and np.count_nonzero(A) < 0.5 * d * d):
                cand = LinearDrift(A)
            ok = ok and cand is not None
        else:
            if op.kind == DIFF_USER_DIAG:
                op.__dict__[key] = op
                return op
            cols = [np.asarray(op(eye[l], 0.0), dtype=np.float64) for l in range(d)]                   # each (d, m)
            B = np.stack([np.stack([cols[l][:, k] for l in range(d)], axis=1) for k in range(op.m)])  # (m, d, d)
            lin = lambda y: np.dot(B, y).T
            scale = max(1.0, float(np.abs(B).max()))
            G0 = np.asarray(op(np.zeros(d), 0.0), dtype=np.float64)                                     # (d, m)
            ok = bool(np.all(np.isfinite(G0)))
            if ok and np.abs(G0).max() == 0.0:
                cand = LinearDiffusion(B)
                if getattr(op, "commutative", False):      # the caller's promise outlives the promotion
                    cand._commutative = True
            elif ok:                                   # B_k y + c_k: read B off G(e_l) - G(0)
                B = B - G0.T[:, :, None]
                lin = lambda y: np.dot(B, y).T + G0
                cand = AffineDiffusion(B, G0)
            else:
                cand = None
        for trial in range(4):
            if not ok:
                break
            y = rng.standard_normal(d) * (1.0 + trial)
            t = (0.0, 0.731, 2.9, -1.3)[trial]
            got = np.asarray(op(y, t), dtype=np.float64)
            ok = bool(np.all(np.isfinite(got))) and \
                float(np.abs(got - lin(y)).max()) <= 64 * np.finfo(float).eps * scale * float(np.abs(y).sum() + 1.0)
        if ok:
            res = cand
    except Exception:
        res = op                                   # anything odd (domain errors at the probes): leave it alone
    op.__dict__[key] = res
    if len(_AS_LINEAR_CACHE) >= 256:
        _AS_LINEAR_CACHE.clear()
    _AS_LINEAR_CACHE[ckey] = False if res is op else res
    return res


_AS_LINEAR_CACHE = {}


# ---- named systems of SURVEY.md 8(d) --------------------------------------------------

def sys_lin(d, m):
    """SYS_LIN(d, m): non-commutative linear system generalising Roessler (7.4).
    A[i,i] = -1.5, A[i,j] = 0.5/d; B_k = (0.5/sqrt m)(1 + 0.5 R_k / sqrt d),
    R_k = default_rng(7000+k).standard_normal((d,d)); y0 = ones(d)."""
    A = np.full((d, d), 0.5 / d)
    np.fill_diagonal(A, -1.5)
    B = np.empty((m, d, d))
    for k in range(m):
        R = np.random.default_rng(7000 + k).standard_normal((d, d))
        B[k] = (0.5 / np.sqrt(m)) * (np.eye(d) + 0.5 * R / np.sqrt(d))
    return LinearDrift(A), LinearDiffusion(B), np.ones(d)


def sys_lin2():
    """README 2-d linear Ito SDE (README.rst:83-95)."""
    A = np.array([[-0.5, -2.0], [2.0, -1.0]])
    B = np.diag([0.5, 0.5])
    return LinearDrift(A), ConstDiffusion(B), np.array([3.0, 3.0])


def sys_kp4446(stratonovich=False, a=1.0, b=0.8):
    """README scalar SDE / KP (4.46) (README.rst:60-69)."""
    return KP4446Drift(a, b, stratonovich), KP4446Diffusion(b), 0.1


def sys_kp4459(stratonovich=False):
    """KP (4.59): d=1, m=2 geometric Brownian motion (test_integrate.py:170-199)."""
    a, b1, b2 = -0.02, 1.0, 2.0
    coef = a - (b1 ** 2 + b2 ** 2) / 2.0 if stratonovich else a
    return (LinearDrift([[coef]]), LinearDiffusion(np.array([[[b1]], [[b2]]])),
            np.array([1.0]))


def sys_kps445(stratonovich=False):
    return KPS445Drift(stratonovich), KPS445Diffusion(), 1.0


def sys_r74(stratonovich=False):
    """Roessler (7.4): d=5, m=4 linear system (test_integrate.py:350-375)."""
    d, m = 5, 4
    A = np.full((d, d), 0.05)
    B0 = np.full((d, d), 0.01)
    np.fill_diagonal(A, -1.5)
    np.fill_diagonal(B0, 0.2)
    if stratonovich:
        S = np.full((d, d), -0.0086)
        np.fill_diagonal(S, -0.0808)
        A = A + S
    return LinearDrift(A), LinearDiffusion(np.stack([B0] * m)), np.ones(d)


def sys_ou(theta=1.0, sigma=0.2):
    """1-d Ornstein-Uhlenbeck, dy = -theta y dt + sigma dW (test_integrate.py:37-54)."""
    return LinearDrift([[-theta]]), ConstDiffusion([[sigma]]), 0.0
