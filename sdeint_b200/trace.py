"""Plain Python callables f(y, t), G(y, t) -- what the reference takes (sdeint/integrate.py:124-146) -- turned
into device operators by TRACING: the function is called once with symbolic stand-ins for y and t that record
the arithmetic done on them, and the recorded expressions become an `ops.ExprDrift` / `ops.ExprDiffusion`
(CUDA source for NVRTC + the numpy callable used by tests).  Expressions that turn out to be constant or
exactly linear are handed on to `ConstDiffusion` / `LinearDrift` / `LinearDiffusion`, i.e. to the additive-noise
path and the tensor-instruction kernels (ops.as_linear).

What can be traced: + - * / **, unary minus, abs, numpy ufuncs sin cos tan exp log sqrt tanh sinh cosh log1p
expm1 log10 log2 cbrt arctan arcsin arccos arctan2 hypot,
np.maximum / np.minimum / np.sign / np.heaviside, comparisons used as 0/1 masks (also on arrays of symbols: `np.sin(y)`, `A.dot(y)`, `y**2`, `np.array([...])`), indexing, closures over
numbers and arrays.  What cannot: anything that needs the VALUE of y or t while tracing -- comparisons and
branches on y, float(y), np.where on symbols, calls into compiled code.  Those raise TraceError
and the caller is pointed to ops.ExprDrift / ops.CudaDrift."""
from __future__ import annotations

import numbers

import numpy as np

_FUNCS = ("sin", "cos", "tan", "exp", "log", "sqrt", "tanh", "sinh", "cosh", "log1p", "expm1", "log10", "log2",
          "cbrt")
_RENAMED = {"arctan": "atan", "arcsin": "asin", "arccos": "acos"}      # numpy name -> C name
_BINARY = {"arctan2": "atan2", "hypot": "hypot"}


class TraceError(Exception):
    pass


def _lit(v):
    """A float as a literal that survives the round trip (repr) and needs no sign handling."""
    v = float(v)
    if not np.isfinite(v):
        raise TraceError("non-finite constant %r in the traced function" % v)
    r = repr(v)
    if "e" not in r and "." not in r:
        r += ".0"
    return "(%s)" % r if v < 0 else r


class Sym:
    """A recorded scalar expression: either a constant (`c` is its value) or a string in y[i], t."""
    __slots__ = ("e", "c")
    __array_priority__ = 10000.0

    def __init__(self, e=None, c=None):
        self.e, self.c = e, c

    # -- helpers ---------------------------------------------------------------------------------------
    @staticmethod
    def of(x):
        if isinstance(x, Sym):
            return x
        if isinstance(x, (numbers.Real, np.floating, np.integer)) and not isinstance(x, bool):
            return Sym(c=float(x))
        if isinstance(x, np.ndarray) and x.ndim == 0:
            return Sym.of(x.item())
        raise TraceError("cannot trace an operation with a %s" % type(x).__name__)

    @property
    def s(self):
        return _lit(self.c) if self.c is not None else self.e

    def _bin(self, other, op, swap=False):
        try:
            o = Sym.of(other)
        except TraceError:
            return NotImplemented
        a, b = (o, self) if swap else (self, o)
        if a.c is not None and b.c is not None:
            with np.errstate(all="ignore"):
                return Sym(c=float({"+": np.add, "-": np.subtract, "*": np.multiply, "/": np.divide}[op](a.c, b.c)))
        if op == "+":
            if a.c == 0.0:
                return b
            if b.c == 0.0:
                return a
        elif op == "-":
            if b.c == 0.0:
                return a
            if a.c == 0.0:
                return -b
        elif op == "*":
            if a.c == 0.0 or b.c == 0.0:
                return Sym(c=0.0)
            if a.c == 1.0:
                return b
            if b.c == 1.0:
                return a
        elif op == "/":
            if a.c == 0.0:
                return Sym(c=0.0)
            if b.c == 1.0:
                return a
        return Sym("(%s%s%s)" % (a.s, op, b.s))

    def __add__(self, o): return self._bin(o, "+")
    def __radd__(self, o): return self._bin(o, "+", True)
    def __sub__(self, o): return self._bin(o, "-")
    def __rsub__(self, o): return self._bin(o, "-", True)
    def __mul__(self, o): return self._bin(o, "*")
    def __rmul__(self, o): return self._bin(o, "*", True)
    def __truediv__(self, o): return self._bin(o, "/")
    def __rtruediv__(self, o): return self._bin(o, "/", True)

    def __neg__(self):
        return Sym(c=-self.c) if self.c is not None else Sym("(-%s)" % self.e)

    def __pos__(self):
        return self

    def __abs__(self):
        return Sym(c=abs(self.c)) if self.c is not None else Sym("fabs(%s)" % self.e)

    def __pow__(self, o):
        try:
            o = Sym.of(o)
        except TraceError:
            return NotImplemented
        if self.c is not None and o.c is not None:
            return Sym(c=float(self.c ** o.c))
        if o.c is not None:
            n = o.c
            if n == 0.0:
                return Sym(c=1.0)
            if n == 0.5:
                return Sym("sqrt(%s)" % self.s)
            if n == int(n) and 1 <= abs(int(n)) <= 4:           # small integer powers: multiplications
                prod = "(" + "*".join([self.s] * abs(int(n))) + ")" if abs(int(n)) > 1 else self.s
                return Sym(prod) if n > 0 else Sym("(1.0/%s)" % prod)
        return Sym("pow(%s,%s)" % (self.s, o.s))

    def __rpow__(self, o):
        return Sym.of(o).__pow__(self)

    # numpy calls these by name for object arrays (np.sin(obj_array) -> element.sin())
    def _fn(self, name):
        if self.c is not None:
            with np.errstate(all="ignore"):
                return Sym(c=float(getattr(np, name)(self.c)))
        return Sym("%s(%s)" % (_RENAMED.get(name, name), self.e))

    def _fn2(self, other, name):
        a, b = self, Sym.of(other)
        if a.c is not None and b.c is not None:
            return Sym(c=float(getattr(np, name)(a.c, b.c)))
        return Sym("%s(%s,%s)" % (_BINARY[name], a.s, b.s))

    def sin(self): return self._fn("sin")
    def cos(self): return self._fn("cos")
    def tan(self): return self._fn("tan")
    def exp(self): return self._fn("exp")
    def log(self): return self._fn("log")
    def sqrt(self): return self._fn("sqrt")
    def tanh(self): return self._fn("tanh")
    def sinh(self): return self._fn("sinh")
    def cosh(self): return self._fn("cosh")
    def log1p(self): return self._fn("log1p")
    def expm1(self): return self._fn("expm1")
    def log10(self): return self._fn("log10")
    def log2(self): return self._fn("log2")
    def cbrt(self): return self._fn("cbrt")
    def arctan(self): return self._fn("arctan")
    def arcsin(self): return self._fn("arcsin")
    def arccos(self): return self._fn("arccos")
    def arctan2(self, o): return self._fn2(o, "arctan2")
    def hypot(self, o): return self._fn2(o, "hypot")

    def conjugate(self):
        return self

    # np.sin(sym), np.float64(2) * sym, np.add(arr, sym), ... arrive here
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__" or kwargs.get("out") is not None:
            return NotImplemented
        name = ufunc.__name__
        if any(isinstance(x, np.ndarray) and x.ndim > 0 for x in inputs):   # array (op) symbol: elementwise
            arrs = [x if isinstance(x, np.ndarray) and x.ndim > 0 else None for x in inputs]
            shape = np.broadcast(*[a for a in arrs if a is not None]).shape
            out = np.empty(shape, dtype=object)
            its = [np.broadcast_to(np.asarray(a, dtype=object), shape) if a is not None else None for a in arrs]
            for idx in np.ndindex(*shape):
                args = [it[idx] if it is not None else x for it, x in zip(its, inputs)]
                out[idx] = Sym._apply(name, args)
            return out
        return Sym._apply(name, list(inputs))

    @staticmethod
    def _apply(name, args):
        a = [Sym.of(x) for x in args]
        two = {"add": "+", "subtract": "-", "multiply": "*", "true_divide": "/", "divide": "/"}
        if name in two and len(a) == 2:
            return a[0]._bin(a[1], two[name])
        if name == "power" and len(a) == 2:
            return a[0] ** a[1]
        if name == "negative":
            return -a[0]
        if name == "positive":
            return a[0]
        if name in ("absolute", "fabs"):
            return abs(a[0])
        if name == "square":
            return a[0] ** 2
        if name == "reciprocal":
            return 1.0 / a[0]
        if name in ("maximum", "fmax", "minimum", "fmin") and len(a) == 2:   # e.g. sqrt(max(v, 0)) of CIR / Heston
            fn = "fmax" if name in ("maximum", "fmax") else "fmin"
            if a[0].c is not None and a[1].c is not None:
                return Sym(c=float(getattr(np, fn)(a[0].c, a[1].c)))
            return Sym("%s(%s,%s)" % (fn, a[0].s, a[1].s))
        cmp = {"less": "lt", "less_equal": "le", "greater": "gt", "greater_equal": "ge"}
        if name in cmp and len(a) == 2:
            return a[0]._cmp(a[1], cmp[name], False)
        if name == "sign" and len(a) == 1:
            return a[0]._cmp(0.0, "gt", False) - a[0]._cmp(0.0, "lt", False)
        if name == "heaviside" and len(a) == 2:          # 0 for x < 0, h0 at 0, 1 for x > 0
            x, h0 = a
            return x._cmp(0.0, "gt", False) + h0 * (x._cmp(0.0, "ge", False) - x._cmp(0.0, "gt", False))
        if name == "clip" and len(a) == 3:
            return Sym._apply("minimum", [Sym._apply("maximum", [a[0], a[1]]), a[2]])
        if (name in _FUNCS or name in _RENAMED) and len(a) == 1:
            return a[0]._fn(name)
        if name in _BINARY and len(a) == 2:
            return a[0]._fn2(a[1], name)
        raise TraceError("numpy.%s cannot be traced (supported: arithmetic, abs, %s)" % (name, ", ".join(_FUNCS)))

    # comparisons are recorded as 0.0 / 1.0 masks ((y > 0) * a + (y <= 0) * b); using one as the condition of a
    # Python `if` still needs its value and is refused by __bool__
    def _cmp(self, o, fn, swapped):
        try:
            o = Sym.of(o)
        except TraceError:
            return NotImplemented
        a, b = (o, self) if swapped else (self, o)
        if a.c is not None and b.c is not None:
            return Sym(c=float({"lt": a.c < b.c, "le": a.c <= b.c, "gt": a.c > b.c, "ge": a.c >= b.c}[fn]))
        return Sym("%s(%s,%s)" % (fn, a.s, b.s))

    def __lt__(self, o): return self._cmp(o, "lt", False)
    def __le__(self, o): return self._cmp(o, "le", False)
    def __gt__(self, o): return self._cmp(o, "gt", False)
    def __ge__(self, o): return self._cmp(o, "ge", False)

    # anything that needs the value
    def _no_value(self, *a, **k):
        raise TraceError("the function needs the value of y or t while it is traced (a comparison or branch, "
                         "float() / math.* on y, or storing into an array created with an explicit float dtype); only "
                         "arithmetic and numpy's elementary functions can be traced")

    __bool__ = __float__ = __int__ = __index__ = _no_value
    __eq__ = __ne__ = _no_value
    __hash__ = None

    def __repr__(self):
        return "Sym(%s)" % self.s


import operator as _op

_CMP = {name: np.frompyfunc(getattr(_op, name), 2, 1) for name in ("lt", "le", "gt", "ge")}


class SymArray(np.ndarray):
    """The object array of symbols handed to the traced function as y: numpy's own comparison of object arrays
    converts every result with bool(); here `y > 0` stays an array of recorded 0/1 masks, and np.where on such a
    mask becomes mask * a + (1 - mask) * b."""

    def __lt__(self, o): return _CMP["lt"](np.asarray(self), np.asarray(o, dtype=object)).view(SymArray)
    def __le__(self, o): return _CMP["le"](np.asarray(self), np.asarray(o, dtype=object)).view(SymArray)
    def __gt__(self, o): return _CMP["gt"](np.asarray(self), np.asarray(o, dtype=object)).view(SymArray)
    def __ge__(self, o): return _CMP["ge"](np.asarray(self), np.asarray(o, dtype=object)).view(SymArray)

    def __array_function__(self, func, types, args, kwargs):
        def plain(x):
            if isinstance(x, SymArray):
                return np.asarray(x)
            if isinstance(x, (list, tuple)):
                return type(x)(plain(v) for v in x)
            return x
        if func is np.where and len(args) == 3 and not kwargs:
            c, a, b = (np.asarray(plain(v), dtype=object) for v in args)
            return (c * a + (1.0 - c) * b).view(SymArray)
        out = func(*plain(args), **{k: plain(v) for k, v in kwargs.items()})
        return out.view(SymArray) if isinstance(out, np.ndarray) and out.dtype == object else out


import contextlib
import threading

_PATCH_LOCK = threading.RLock()   # re-entrant: a traced function may itself trace (nested calls)


@contextlib.contextmanager
def _object_allocators():
    """While a function is traced, np.zeros / ones / empty / full without an explicit dtype return OBJECT arrays, so
    the common `out = np.zeros(d); out[0] = ...; return out` can hold symbols (a float array cannot).  The
    originals are restored on exit; one trace at a time (the patch is process-wide)."""
    names = ("zeros", "ones", "empty", "full")
    orig = {n: getattr(np, n) for n in names}

    def wrap(fn):
        def alloc(*args, **kwargs):
            npos = 3 if fn is orig["full"] else 2           # position of dtype among the positional arguments
            if "dtype" not in kwargs and len(args) < npos:
                kwargs["dtype"] = object
            return fn(*args, **kwargs)
        return alloc

    with _PATCH_LOCK:
        if getattr(np.zeros, "_sdb_traced", False):            # nested trace: already patched
            yield
            return
        try:
            for n in names:
                w = wrap(orig[n])
                w._sdb_traced = True
                setattr(np, n, w)
            yield
        finally:
            for n in names:
                setattr(np, n, orig[n])


def _strings(res, shape, what):
    """The traced return value as an object array of expression strings of the given shape."""
    if isinstance(res, (list, tuple)):
        res = np.array([np.asarray(r, dtype=object) if isinstance(r, (list, tuple, np.ndarray)) else r
                        for r in res], dtype=object)
    arr = np.asarray(res, dtype=object)
    try:
        arr = arr.reshape(shape)
    except ValueError:
        raise TraceError("%s returned an array of shape %s, expected %s" % (what, arr.shape, shape))
    out = np.empty(shape, dtype=object)
    for idx in np.ndindex(*shape):
        out[idx] = Sym.of(arr[idx]).s
    return out


_KEEP = (0.0, 0.5, 1.0, 2.0, 3.0, 4.0)


def _lift(exprs):
    """Numeric constants of the recorded expressions -> parameters p[k]: a sweep over the numbers a function
    closes over then produces the SAME source every time (one NVRTC compilation, found again in the on-disk
    cache by later processes) and only the parameter vector changes.  Structural constants (0, 1/2, 1 ... 4)
    stay literals.  Returns (expressions, parameter vector)."""
    import re
    values, index = [], {}

    def repl(mt):
        v = float(mt.group(0))
        if v in _KEEP:
            return mt.group(0)
        if v not in index:
            index[v] = len(values)
            values.append(v)
        return "p[%d]" % index[v]

    # every literal written by _lit has a '.' or an exponent; subscripts (y[3]) are bare integers
    number = re.compile(r"(?<![\w\[.])(?:\d+\.\d*(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)")
    return [number.sub(repl, e) for e in exprs], np.array(values, dtype=np.float64)


def _symbols(d, scalar):
    t = Sym("t")
    if scalar:
        return Sym("y[0]"), t
    y = np.empty(d, dtype=object)
    for i in range(d):
        y[i] = Sym("y[%d]" % i)
    return y.view(SymArray), t


def trace_drift(f, d, scalar=False):
    """The drift operator of a Python callable f(y, t) -> (d,) (a scalar for scalar y0)."""
    from . import ops
    y, t = _symbols(d, scalar)
    with _object_allocators():
        res = f(y, t)
    exprs, params = _lift(list(_strings(res, (d,), "f(y, t)")))
    op = ops.ExprDrift(exprs, params)
    op.traced_from = f
    return op


def trace_diffusion(G, d, scalar=False):
    """The diffusion operator of a callable G(y, t) -> (d, m) (a scalar for scalar y0), or of the list
    [g_1, ..., g_m] of column callables g_k(y, t) -> (d,) (the reference's alternative form,
    sdeint/integrate.py:330-334).  Constant results become ops.ConstDiffusion (additive noise)."""
    from . import ops
    y, t = _symbols(d, scalar)
    if isinstance(G, (list, tuple)):
        with _object_allocators():
            raw = [g(y, t) for g in G]
        cols = [_strings(r, (d,), "g_k(y, t)") for r in raw]
        table = np.stack(cols, axis=1)
    else:
        with _object_allocators():
            res = G(y, t)
        arr = np.asarray(res, dtype=object)
        if d == 1 and arr.ndim == 0:                           # scalar equation, scalar G (m = 1)
            arr = arr.reshape(1, 1)
        if arr.ndim != 2 or arr.shape[0] != d:
            raise TraceError("G(y, t) must return an array of shape (%d, m), got %s" % (d, arr.shape))
        table = _strings(arr, arr.shape, "G(y, t)")
    try:                                                       # no y, no t anywhere: additive noise
        B = np.array([[float(eval(e, {"__builtins__": {}}, {})) for e in row] for row in table])
        op = ops.ConstDiffusion(B)
    except Exception:
        # diagonal noise -- G_ik = 0 off the diagonal and G_kk a function of y_k and t only, the common case of
        # one independent noise per component -- needs no Levy areas (SRI2/SRS2 use I_kk alone): read off the
        # recorded expressions, so it is exact
        import re
        diag = table.shape[0] == table.shape[1] and all(
            (table[i, k] in ("0.0", "(-0.0)") if i != k
             else set(re.findall(r"y\[(\d+)\]", table[i, k])) <= {str(k)})
            for i in range(table.shape[0]) for k in range(table.shape[1]))
        flat, params = _lift([e for row in table for e in row])
        mcols = table.shape[1]
        op = ops.ExprDiffusion([flat[i * mcols:(i + 1) * mcols] for i in range(table.shape[0])], params,
                               diagonal=bool(diag))
    op.traced_from = G
    return op
