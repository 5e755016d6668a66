"""The integrators of sdeint/integrate.py with the same names, argument order and keyword
names, running on the GPU.

    itoint(f, G, y0, tspan, generator=None)                          integrate.py:124
    stratint(f, G, y0, tspan, generator=None)                        integrate.py:157
    itoEuler(f, G, y0, tspan, dW=None, generator=None)               integrate.py:190
    stratHeun(f, G, y0, tspan, dW=None, generator=None)              integrate.py:243
    itoSRI2(f, G, y0, tspan, Imethod=Ikpw, dW=None, I=None, generator=None)   :306
    stratSRS2(f, G, y0, tspan, Jmethod=Jkpw, dW=None, J=None, generator=None) :371
    stratKP2iS(f, G, y0, tspan, Jmethod=Jkpw, gam=None, al1=None, al2=None,
               rtol=1e-4, dW=None, J=None, generator=None)           integrate.py:526

Differences from the reference, all forced by "the loop runs in a CUDA kernel":
  * `f` and `G` are operator objects from `sdeint_b200.ops` (built-in families or NVRTC
    source), or plain Python callables f(y, t), G(y, t) / [g_1, ..., g_m] made of arithmetic and
    elementary functions, which are traced into such operators (trace.py); a callable that
    branches on the value of y cannot be traced and raises SDEValueError; `G` may also be the
    list `G.columns()`.
  * fp64 only (the north-star scope); other dtypes raise SDEValueError.
  * noise that the reference would draw from numpy's generator is drawn by an on-device
    Philox generator (see sdeint_b200/wiener.py); results on host-supplied `dW`, `I`/`J`
    match the reference to <= 1e-12 relative.
Noise structure is exploited where the result is unchanged: additive noise (ops.ConstDiffusion)
never builds I/J, diagonal noise (ops.DiagLinearDiffusion, CudaDiffusion(diagonal=True)) builds only
I_kk = (dW_k^2 - h)/2 for SRI2/SRS2, scalar noise draws nothing beyond dW with Iwik, and commutative
noise (LinearDiffusion with commuting B_k -- detected --, or `commutative=True` on a user operator or
on the call, or Imethod=Icomm / Jmethod=Jcomm) runs SRI2/SRS2 on I = (dW dW^T - h 1)/2 without Levy
areas; `commutative=False` on the call forces the full construction -- the first items of the
reference's own to-do list (integrate.py:150-151, README.rst:130).
Keyword-only ENSEMBLE extensions: `paths=P` integrates P independent sample paths in one
launch (returns (P, n_saved, d)); `seed`, `path_id0` select the Philox streams; `save_every`
decimates the stored trajectory; `out="torch"` leaves the result on the device in the
path-minor layout (n_saved, d, P); `host_out=` (a pinned torch tensor (n_saved, d, P)) receives
large results chunk by chunk while the next chunk of paths is integrated.  `dW` / `I` / `J` may
then carry a leading path axis (without it, one realisation is shared by all paths).
`step0=k` (with `y0_paths=` the states reached and, for bit-identical noise, `h_noise=` the h of the
whole run) continues a run after k intervals: the on-device noise depends only on (seed, path id, step
counter), so a run cut into segments -- checkpoint / resume, full-resolution output produced piecewise --
gives the same bits as one call.
`n_terms=` sets the number of series terms of Ikpw/Iwik (the reference's n = 5 by default);
`n_terms="auto"` picks the number that keeps strong order 1.0 at this h (wiener.series_terms).
`observe=[("first_above", comp, level), ("first_below", comp, level), ("max", comp), ("min", comp),
("integral", comp)]` (up to 4) evaluates path functionals at EVERY grid point inside the step loop --
what a decimated store cannot give back -- and makes the call return `(y, obs)`, obs of shape
(P, n_obs): first-passage times are grid times (inf if never), the integral is trapezoidal.
"""
from __future__ import annotations

import ctypes as C
import numbers

import numpy as np

from . import _device as dev
from . import _lib, ops, trace
from . import wiener
from .wiener import Icomm, Ikpw, Iwik, Jcomm, Jkpw, Jwik, deltaW  # noqa: F401  (re-exported like the reference)

SCHEME_EULER, SCHEME_HEUN, SCHEME_SRK2, SCHEME_KP2IS, SCHEME_R2 = 0, 1, 2, 3, 4
NOISE_HOST, NOISE_KPW, NOISE_WIK, NOISE_COMM = 0, 1, 2, 3
FLAG_Y0_BROADCAST, FLAG_STRATONOVICH = 1, 2


class Error(Exception):
    pass


class SDEValueError(Error):
    """Thrown if integration arguments fail some basic sanity checks"""
    pass


_last_status = None


def last_status():
    """(non-finite paths, first such path id, unconverged implicit solves, first such step)
    of the most recent integration, or None."""
    return _last_status


def _check_args(f, G, y0, tspan, dW=None, IJ=None):
    """Validation common to all algorithms; finds d and m.  Mirrors the checks of
    sdeint/integrate.py:47-121 (equal spacing, scalar promotion, int -> float64, shape of
    f(y0), G as one (d, m) operator or a list of m columns, shapes of dW and I/J), with an
    optional leading path axis on dW / IJ.  Returns (d, m, f, Gop, y0, tspan, dW, IJ)."""
    # C-contiguous: the C ABI reads h_tspan[i] with unit stride (a view such as t[::2] is copied)
    tspan = np.ascontiguousarray(tspan, dtype=np.float64)
    if tspan.ndim != 1 or tspan.shape[0] < 2:
        raise SDEValueError('tspan must be a 1-d sequence of at least two time points.')
    steps = np.diff(tspan)
    if not np.isclose(steps.min(), steps.max()):
        raise SDEValueError('Currently time steps must be equally spaced.')
    scalar_y0 = isinstance(y0, numbers.Number)
    if isinstance(y0, numbers.Number):
        if isinstance(y0, numbers.Complex) and not isinstance(y0, numbers.Real):
            raise SDEValueError('sdeint_b200 integrates real float64 systems only.')
        if isinstance(y0, np.floating) and y0.dtype != np.float64:
            raise SDEValueError('sdeint_b200 integrates float64 systems only (got %s).' % y0.dtype)
        y0 = np.array([float(y0)], dtype=np.float64)
    else:
        y0 = np.array(y0)
        if y0.dtype.kind in 'iub':
            y0 = y0.astype(np.float64)
        if y0.dtype != np.float64:
            raise SDEValueError('sdeint_b200 integrates float64 systems only (got %s).' % y0.dtype)
        if y0.ndim != 1:
            raise SDEValueError('y0 must be a scalar or a 1-d array of length d.')
    d = len(y0)
    if not isinstance(f, ops.DriftOp):
        # a plain Python callable, as the reference takes it (integrate.py:124-146): traced once with
        # symbolic y, t into an expression operator (trace.py); what cannot be traced is refused
        if not callable(f):
            raise SDEValueError(
                'f must be a callable f(y, t) or a drift operator from sdeint_b200.ops (LinearDrift, '
                'KP4446Drift, ..., ExprDrift, or CudaDrift with NVRTC source).')
        try:
            f = trace.trace_drift(f, d, scalar_y0)
        except trace.TraceError as exc:
            raise SDEValueError(
                'f is a Python callable that could not be turned into a device function: %s.  Give it as '
                'sdeint_b200.ops.ExprDrift / CudaDrift instead (a Python callable cannot run inside a CUDA '
                'kernel and this package has no CPU fallback).' % exc)
        except SDEValueError:
            raise
        except Exception as exc:
            raise SDEValueError('y0 and f have incompatible shapes, or f could not be traced (%s: %s).'
                                % (type(exc).__name__, exc))
    if f.d != d:
        raise SDEValueError('y0 and f have incompatible shapes.')
    message = ("y0 has length %d. So G must either be a single diffusion operator of shape "
               "(%d, m), or else the list of its m columns (G.columns()), each of shape (%d,)"
               % (d, d, d))
    if isinstance(G, ops.DiffusionOp):
        Gop = G
    elif isinstance(G, (list, tuple)) and len(G) > 0 and any(isinstance(c, ops.DiffusionColumn) for c in G):
        cols = tuple(G)
        if not all(isinstance(c, ops.DiffusionColumn) for c in cols):
            raise SDEValueError(message)
        Gop = cols[0].parent
        if (len(cols) != Gop.m or any(c.parent is not Gop for c in cols)
                or [c.k for c in cols] != list(range(Gop.m))):
            raise SDEValueError('a list G must be exactly G.columns() of one diffusion operator.')
    elif callable(G) or (isinstance(G, (list, tuple)) and len(G) > 0 and all(callable(g) for g in G)):
        try:
            Gop = trace.trace_diffusion(G, d, scalar_y0)   # G(y, t) -> (d, m), or the list of its columns
        except trace.TraceError as exc:
            raise SDEValueError(
                'G is a Python callable that could not be turned into a device function: %s.  Give it as '
                'sdeint_b200.ops.ExprDiffusion / CudaDiffusion instead (a Python callable cannot run inside '
                'a CUDA kernel).' % exc)
        except Exception as exc:
            raise SDEValueError('%s (%s: %s)' % (message, type(exc).__name__, exc))
    else:
        raise SDEValueError(
            'G must be a callable G(y, t), the list of its column callables, or a diffusion operator from '
            'sdeint_b200.ops (or its .columns() list).')
    if Gop.d != d:
        raise SDEValueError(message)
    # expressions that are exactly linear get the kernels of the linear operators (ops.as_linear); only where
    # those exist (d m >= 16: below that the unrolled source, whose structural zeros vanish, is the faster one)
    if d * Gop.m >= 16:
        Gop = ops.as_linear(Gop)
        if Gop.kind in (ops.DIFF_LINEAR_COLS, ops.DIFF_AFFINE_COLS):
            f = ops.as_linear(f, dense_only=True)
    m = Gop.m
    N = len(tspan)
    if dW is not None:
        if not hasattr(dW, 'shape') or tuple(dW.shape[-2:]) != (N - 1, m) or len(dW.shape) > 3:
            raise SDEValueError(
                "From function G, it seems m==%d. If present, the optional parameter dW must "
                "be an array of shape (len(tspan)-1, m) [or (paths, len(tspan)-1, m)] giving m "
                "independent Wiener increments for each time interval." % m)
    if IJ is not None:
        if not hasattr(IJ, 'shape') or tuple(IJ.shape[-3:]) != (N - 1, m, m) or len(IJ.shape) > 4:
            raise SDEValueError(
                "From function G, it seems m==%d. If present, the optional parameter I or J "
                "must be an array of shape (len(tspan)-1, m, m) [or (paths, len(tspan)-1, m, m)] "
                "giving an m x m matrix of repeated integral values for each time interval." % m)
    return (d, m, f, Gop, y0, tspan, dW, IJ)


def _method_id(method, strat):
    table = ({Jkpw: NOISE_KPW, Jwik: NOISE_WIK, Jcomm: NOISE_COMM} if strat
             else {Ikpw: NOISE_KPW, Iwik: NOISE_WIK, Icomm: NOISE_COMM})
    if method in table:
        return table[method]
    names = "Jkpw, Jwik or Jcomm" if strat else "Ikpw, Iwik or Icomm"
    raise SDEValueError("the repeated-integral method must be sdeint_b200.%s" % names)


def _resolve_paths(paths, dW, IJ):
    """Ensemble size: explicit `paths`, else the leading axis of a 3-d dW / 4-d IJ, else one
    path returned in the reference's (N, d) layout."""
    lead = []
    if dW is not None and len(dW.shape) == 3:
        lead.append(int(dW.shape[0]))
    if IJ is not None and len(IJ.shape) == 4:
        lead.append(int(IJ.shape[0]))
    if paths is not None:
        P = int(paths)
        if P < 1:
            raise SDEValueError('paths must be >= 1')
        if any(p != P for p in lead):
            raise SDEValueError('dW / I / J path axis does not match paths=%d' % P)
        return P, True
    if lead:
        if any(p != lead[0] for p in lead):
            raise SDEValueError('dW and I / J disagree on the number of paths')
        return lead[0], True
    return 1, False


def _run(scheme, f, G, y0, tspan, dW, IJ, strat, method, generator, kp2is, paths, seed,
         path_id0, save_every, n_terms, device, out, y0_paths=None, host_out=None,
         commutative=None, observe=None, step0=0, h_noise=None):
    global _last_status
    (d, m, f, Gop, y0, tspan, dW, IJ) = _check_args(f, G, y0, tspan, dW, IJ)
    P, ensemble = _resolve_paths(paths, dW, IJ)
    # additive noise (G independent of y): SRI2/SRS2/KP2iS never use the repeated integrals, so
    # they are neither generated nor uploaded -- results are bit-identical (ops.ConstDiffusion)
    need_ij = scheme in (SCHEME_SRK2, SCHEME_KP2IS) and Gop.kind != ops.DIFF_CONST
    # diagonal noise: only I_kk = (dW_k^2 - h)/2 enters; unless the caller supplies I/J the kernel
    # builds that diagonal from dW and no Levy areas are generated at all
    diag = Gop.kind in (ops.DIFF_DIAG_LINEAR, ops.DIFF_USER_DIAG)
    if diag and IJ is None and scheme == SCHEME_SRK2:   # (KP2iS does use off-diagonal J)
        need_ij = False
    N = len(tspan)
    step0 = int(step0)
    if step0 < 0 or (h_noise is not None and not h_noise > 0.0):
        raise SDEValueError('step0 must be >= 0 and h_noise > 0')
    if step0 and scheme == SCHEME_KP2IS:
        raise SDEValueError('stratKP2iS is a two-step scheme and cannot be resumed (step0 != 0)')
    save_every = int(save_every)
    if save_every < 1:
        raise SDEValueError('save_every must be >= 1')
    n_saved = (N - 1) // save_every + 1
    h = (tspan[N - 1] - tspan[0]) / (N - 1)
    if h_noise:
        h = float(h_noise)      # the h of the whole run when this call is one segment of it
    if (need_ij and IJ is None and method is not None and callable(method)
            and method not in (Ikpw, Iwik, Icomm, Jkpw, Jwik, Jcomm)):
        # Any function (dW, h, generator=...) -> (A, I or J), like the Imethod / Jmethod argument of the
        # reference (integrate.py:306, :371, called at :487): evaluated on the HOST, one path at a time,
        # and fed to the kernels as caller-supplied integrals.  Meant for single paths / small ensembles
        # (the noise of the production path is generated on the device).
        gen = generator if generator is not None else np.random.default_rng()
        if dW is None:
            dW = wiener.deltaW(N - 1, m, h, generator=gen, seed=seed, path_id0=path_id0,
                               paths=(P if ensemble else None), device=device)
        dWh = np.asarray(dW, dtype=np.float64)
        if dWh.ndim == 3:
            IJ = np.stack([np.asarray(method(dWh[p], h, generator=gen)[1], dtype=np.float64)
                           for p in range(dWh.shape[0])])
        else:
            IJ = np.asarray(method(dWh, h, generator=gen)[1], dtype=np.float64)
        if tuple(IJ.shape[-3:]) != (N - 1, m, m):
            raise SDEValueError('the repeated-integral method returned an array of shape %s, expected '
                                '(len(tspan)-1, m, m) = %s' % (IJ.shape, (N - 1, m, m)))
    method_id = _method_id(method, strat) if (need_ij and IJ is None) else NOISE_KPW
    if isinstance(n_terms, str):
        if n_terms != "auto":
            raise SDEValueError('n_terms must be an integer or "auto"')
        # enough series terms for strong order 1.0 at this step size (wiener.series_terms)
        n_terms = wiener.series_terms(h, m, "wik" if method_id == NOISE_WIK else "kpw") if need_ij else 1
    # commutative noise: the antisymmetric part of I/J (the Levy areas) cancels in the order-1.0 term
    # of SRI2/SRS2 -- exactly for linear g_k = B_k y with commuting B_k, which LinearDiffusion detects --
    # so unless the caller supplies I/J or says commutative=False the areas are never generated
    if (need_ij and IJ is None and scheme == SCHEME_SRK2 and commutative is not False
            and (commutative or Gop.commutative)):
        method_id = NOISE_COMM
    device = dev.require_cuda(device)
    torch = dev.torch_mod()
    lib = _lib.load()
    have_dW, have_IJ = dW is not None, (IJ is not None or not need_ij)
    key = 0
    if not (have_dW and have_IJ):
        key = wiener._derive_seed(seed, generator)
        if step0 and (have_dW or (need_ij and IJ is not None)):
            raise SDEValueError('a resumed segment (step0 != 0) needs dW and I/J both supplied or both drawn')

    with torch.cuda.device(device):
        system = dev.get_system(f, Gop, device)
        a = _lib.IntegrateArgs()
        keep = []
        # --- noise --------------------------------------------------------------------
        if not have_dW and not (need_ij and IJ is not None):
            a.noise = method_id                         # everything drawn on the device, fused
        else:
            a.noise = NOISE_HOST
            if have_dW:
                # a torch tensor (host or device) is used as it is; anything else goes through numpy
                dW_h = dW if isinstance(dW, torch.Tensor) else np.asarray(dW, dtype=np.float64)
                # (N-1, m): one realisation shared by every path -> path stride 0, no replication
                sW = (m, 1, (N - 1) * m if dW_h.ndim == 3 else 0)
                d_dW = dev.to_device(dW_h, device).contiguous()   # (P, N-1, m) or (N-1, m)
            else:                                        # I/J given, dW drawn (integrate.py:481-483)
                d_dW = torch.empty((N - 1, m, P), dtype=torch.float64, device=device)
                _lib.check(lib.sdb_delta_w(N - 1, m, float(h), P, int(path_id0), key,
                                           dev.ptr(d_dW), m * P, P, 1, dev.stream_ptr()))
                sW = (m * P, P, 1)
            keep.append(d_dW)
            a.d_dW = dev.ptr(d_dW)
            a.dW_stride_step, a.dW_stride_comp, a.dW_stride_path = sW
            if need_ij:
                if IJ is not None:
                    IJ_h = IJ if isinstance(IJ, torch.Tensor) else np.asarray(IJ, dtype=np.float64)
                    ij_path = (N - 1) * m * m if IJ_h.ndim == 4 else 0
                    d_IJ = dev.to_device(IJ_h, device).contiguous()  # (P, N-1, m, m) or (N-1, m, m)
                else:                                    # dW given, I/J drawn from it (:484-487)
                    ij_path = (N - 1) * m * m
                    d_IJ = torch.empty((P, N - 1, m, m), dtype=torch.float64, device=device)
                    r = _lib.RepintArgs()
                    r.method, r.stratonovich, r.n_terms, r.m = method_id, int(strat), n_terms, m
                    r.steps, r.paths, r.path_id0, r.seed, r.h = N - 1, P, int(path_id0), key, float(h)
                    r.d_dW = a.d_dW
                    r.dW_stride_step, r.dW_stride_comp, r.dW_stride_path = sW
                    r.d_IJ = dev.ptr(d_IJ)
                    r.IJ_stride_path, r.IJ_stride_step = (N - 1) * m * m, m * m
                    r.IJ_stride_row, r.IJ_stride_col = m, 1
                    r.stream = dev.stream_ptr()
                    _lib.check(lib.sdb_repeated_integrals(C.byref(r)))
                keep.append(d_IJ)
                a.d_IJ = dev.ptr(d_IJ)
                a.IJ_stride_path, a.IJ_stride_step = ij_path, m * m
                a.IJ_stride_row, a.IJ_stride_col = m, 1
        # --- state --------------------------------------------------------------------
        if isinstance(y0_paths, torch.Tensor):
            # a device tensor is used in place through its strides: `y[-1].T` of an out="torch" result
            # (path-minor storage) continues a run without a copy
            if tuple(y0_paths.shape) != (P, d) or y0_paths.dtype != torch.float64:
                raise SDEValueError('y0_paths must be float64 of shape (paths, d)')
            d_y0 = y0_paths.to(device)
            a.flags = 0
            a.y0_stride_path, a.y0_stride_comp = d_y0.stride(0), d_y0.stride(1)
        elif y0_paths is not None:
            y0p = np.asarray(y0_paths, dtype=np.float64)
            if y0p.shape != (P, d):
                raise SDEValueError('y0_paths must have shape (paths, d)')
            d_y0 = dev.to_device(y0p, device)
            a.flags = 0
            a.y0_stride_comp, a.y0_stride_path = 1, d
        else:
            d_y0 = dev.to_device(y0, device)
            a.flags = FLAG_Y0_BROADCAST
        if strat:
            a.flags |= FLAG_STRATONOVICH
        a.scheme, a.n_terms = scheme, int(n_terms)
        a.N, a.save_every, a.seed = N, save_every, key
        a.h_tspan = dev.host_ptr(tspan)
        a.d_y0 = dev.ptr(d_y0)
        if kp2is is not None:
            kp = np.ascontiguousarray(np.concatenate(kp2is[:3]), dtype=np.float64)
            a.h_kp2is = dev.host_ptr(kp)
            a.rtol = float(kp2is[3])
        obs_arr = _observers(observe, d) if observe is not None else None
        if (host_out is not None and out != "torch" and a.noise != NOISE_HOST and y0_paths is None
                and P >= _PIPELINE_MIN_PATHS):
            # Large ensemble, on-device noise, caller-owned host result: integrate the paths in chunks
            # and copy chunk c to the host while chunk c+1 is being integrated (two device buffers).
            if tuple(host_out.shape) != (n_saved, d, P) or host_out.dtype != torch.float64 \
                    or not host_out.is_contiguous():
                raise SDEValueError('host_out must be a contiguous float64 tensor of shape %s'
                                    % ((n_saved, d, P),))
            d_obs = None if obs_arr is None else \
                torch.empty((len(obs_arr), P), dtype=torch.float64, device=device)
            status = _run_pipelined(lib, system, a, host_out, P, int(path_id0), n_saved, d, device,
                                    obs_arr, d_obs, step0, h_noise)
            _last_status = status
            if scheme == SCHEME_KP2IS and status[2] > 0:
                raise RuntimeError(
                    "At time t_n = %g Failed to solve for Y_{n+1}: the implicit equation did not "
                    "converge to rtol=%g on %d path(s)." % (tspan[status[3]], kp2is[3], status[2]))
            host = host_out.numpy()
            if d_obs is not None:
                return host.transpose(2, 0, 1), np.ascontiguousarray(d_obs.cpu().numpy().T)
            if not ensemble:
                return np.ascontiguousarray(host[:, :, 0])
            return host.transpose(2, 0, 1)
        d_y = torch.empty((n_saved, d, P), dtype=torch.float64, device=device)
        d_status = torch.empty(4, dtype=torch.int64, device=device)
        a.paths, a.path_id0 = P, int(path_id0)
        a.d_y = dev.ptr(d_y)
        a.y_stride_save, a.y_stride_comp, a.y_stride_path = d * P, P, 1
        a.d_status = dev.ptr(d_status)
        a.stream = dev.stream_ptr()
        d_obs = None
        if obs_arr is not None:   # path functionals evaluated at every grid point inside the step loop
            d_obs = torch.empty((len(obs_arr), P), dtype=torch.float64, device=device)
        _launch(lib, system, a, obs_arr, dev.ptr(d_obs) if d_obs is not None else None, P, step0, h_noise)
        status = tuple(int(v) for v in d_status.cpu().tolist())
        _last_status = status
        if scheme == SCHEME_KP2IS and status[2] > 0:
            raise RuntimeError(
                "At time t_n = %g Failed to solve for Y_{n+1}: the implicit equation did not "
                "converge to rtol=%g on %d path(s)." % (tspan[status[3]], kp2is[3], status[2]))
        if out == "torch":
            return d_y if d_obs is None else (d_y, d_obs)
        if host_out is not None:
            # caller-owned (ideally pinned) host tensor in the device layout (n_saved, d, P)
            if tuple(host_out.shape) != tuple(d_y.shape) or host_out.dtype != d_y.dtype:
                raise SDEValueError('host_out must be a float64 tensor of shape %s'
                                    % (tuple(d_y.shape),))
            host_out.copy_(d_y, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            host = host_out.numpy()
        else:
            host = d_y.cpu().numpy()
    if d_obs is not None:                                # (y, observers): (P, n_obs), or (n_obs,) for one path
        ob = d_obs.cpu().numpy()
        if not ensemble:
            return np.ascontiguousarray(host[:, :, 0]), np.ascontiguousarray(ob[:, 0])
        return host.transpose(2, 0, 1), np.ascontiguousarray(ob.T)
    if not ensemble:
        return np.ascontiguousarray(host[:, :, 0])      # (N, d), y[0] == y0
    return host.transpose(2, 0, 1)                       # (P, n_saved, d)


_PIPELINE_MIN_PATHS = 1 << 19
_PIPELINE_CHUNKS = 8


def _launch(lib, system, a, obs_arr, d_obs_ptr, obs_stride, step0, h_noise):
    """sdb_integrate, or sdb_integrate_segment when observers / a resumed segment are asked for."""
    if obs_arr is None and not step0 and not h_noise:
        _lib.check(lib.sdb_integrate(system.handle, C.byref(a)))
        return
    x = _lib.IntegrateExt()
    x.step0, x.h_noise = int(step0), float(h_noise or 0.0)
    if obs_arr is not None:
        x.n_obs, x.h_obs, x.d_obs = len(obs_arr), obs_arr, d_obs_ptr
        x.obs_stride_obs, x.obs_stride_path = obs_stride, 1
    _lib.check(lib.sdb_integrate_segment(system.handle, C.byref(a), C.byref(x)))


_RING_SLOT_BYTES = 512 << 20
_ring = {}


def _pinned_ring(device, nbytes):
    """Two pinned staging slots of `nbytes` each, kept for the life of the process (pinning is slow)."""
    torch = dev.torch_mod()
    key = str(device)
    ring = _ring.get(key)
    if ring is None or ring[0].numel() * 8 < nbytes:
        ring = [torch.empty((nbytes + 7) // 8, dtype=torch.float64).pin_memory() for _ in range(2)]
        _ring[key] = ring
    return ring


def _run_pipelined(lib, system, a, host_out, P, path_id0, n_saved, d, device, obs_arr=None, d_obs=None,
                   step0=0, h_noise=None):
    """Chunked integration with overlapped device -> host copies; returns the merged status.

    A pinned `host_out` receives the chunks directly.  A pageable one is filled through a fixed ring
    of two pinned slots (<= 512 MB each): chunk c is integrated while chunk c-1 is copied into its
    slot and the host drains that slot into the caller's array, so the pinned footprint does not grow
    with the ensemble (8 ranks x 8.8 GB results on one host)."""
    torch = dev.torch_mod()
    # chunk = a whole number of waves of the fused kernels (one 256-path block per SM), so that only
    # the last chunk has a partial wave -- exactly like the single launch it replaces
    wave = 256 * torch.cuda.get_device_properties(device).multi_processor_count
    direct = host_out.is_pinned()
    chunks = min(_PIPELINE_CHUNKS, max(2, P >> 18))
    if not direct:
        chunks = max(chunks, -(-(8 * n_saved * d * P) // _RING_SLOT_BYTES))
    Pc = max(wave, (P // chunks) // wave * wave)
    chunks = -(-P // Pc)
    compute = torch.cuda.current_stream()
    copier = torch.cuda.Stream(device=device)
    bufs = [torch.empty((n_saved, d, Pc), dtype=torch.float64, device=device) for _ in range(2)]
    ring = None if direct else _pinned_ring(device, 8 * n_saved * d * Pc)
    host_np = None if direct else host_out.numpy()
    stats = torch.empty((chunks, 4), dtype=torch.int64, device=device)
    copied = [None, None]
    base = host_out.data_ptr()

    def drain(c):   # pinned slot of chunk c -> the caller's pageable array (host memcpy)
        p0 = c * Pc
        n = min(Pc, P - p0)
        copied[c & 1].synchronize()
        host_np[:, :, p0:p0 + n] = ring[c & 1].numpy()[:n_saved * d * n].reshape(n_saved, d, n)

    for c in range(chunks):
        p0 = c * Pc
        n = min(Pc, P - p0)
        buf = bufs[c & 1]
        if copied[c & 1] is not None:
            compute.wait_event(copied[c & 1])     # the copy of chunk c-2 has left this buffer
        a.paths, a.path_id0 = n, path_id0 + p0
        a.d_y = dev.ptr(buf)
        a.y_stride_save, a.y_stride_comp, a.y_stride_path = d * Pc, Pc, 1
        a.d_status = C.c_void_p(stats[c].data_ptr())
        a.stream = C.c_void_p(compute.cuda_stream)
        # observers of chunk c land at their global path offset
        _launch(lib, system, a, obs_arr, C.c_void_p(d_obs.data_ptr() + 8 * p0) if d_obs is not None else None,
                P, step0, h_noise)
        done = torch.cuda.Event()
        done.record(compute)
        if not direct and c >= 1:
            drain(c - 1)                          # host copy overlaps the kernel of chunk c
        copier.wait_event(done)
        if direct:
            _lib.check(lib.sdb_copy2d_d2h_async(C.c_void_p(base + 8 * p0), 8 * P, dev.ptr(buf), 8 * Pc,
                                                8 * n, n_saved * d, C.c_void_p(copier.cuda_stream)))
        else:   # rows of n doubles, packed, into slot c & 1 (drained of chunk c-2 one iteration ago)
            _lib.check(lib.sdb_copy2d_d2h_async(C.c_void_p(ring[c & 1].data_ptr()), 8 * n, dev.ptr(buf),
                                                8 * Pc, 8 * n, n_saved * d,
                                                C.c_void_p(copier.cuda_stream)))
        ev = torch.cuda.Event()
        ev.record(copier)
        copied[c & 1] = ev
    if not direct:
        drain(chunks - 1)
    copier.synchronize()
    compute.wait_stream(copier)
    st = stats.cpu().numpy()
    return (int(st[:, 0].sum()), int(st[:, 1].min()), int(st[:, 2].sum()), int(st[:, 3].min()))


_ENS = dict(paths=None, seed=None, path_id0=0, save_every=1, n_terms=5, device=None, out="numpy",
            y0_paths=None, host_out=None, commutative=None, observe=None, step0=0, h_noise=None)

_OBS_KINDS = {"first_above": 0, "first_below": 1, "max": 2, "min": 3, "integral": 4, "lagprod": 5}
MAX_LAG = 32
MAX_OBSERVERS = 4


def _observers(observe, d):
    """`observe=[("first_above", comp, level), ("max", comp), ...]` -> ctypes array of sdb_observer."""
    specs = [observe] if (isinstance(observe, tuple) and observe and isinstance(observe[0], str)) \
        else list(observe)
    if not 1 <= len(specs) <= MAX_OBSERVERS:
        raise SDEValueError('observe takes 1 to %d observers' % MAX_OBSERVERS)
    arr = (_lib.Observer * len(specs))()
    for i, spec in enumerate(specs):
        kind = spec[0]
        if kind not in _OBS_KINDS:
            raise SDEValueError('unknown observer %r (one of %s)' % (kind, ", ".join(sorted(_OBS_KINDS))))
        need_level = kind.startswith("first_") or kind == "lagprod"
        if len(spec) != (3 if need_level else 2):
            raise SDEValueError('observer %r takes %s' % (kind, "(kind, comp, level)" if need_level
                                                          else "(kind, comp)"))
        comp = int(spec[1])
        if not 0 <= comp < d:
            raise SDEValueError('observer component %d out of range for d=%d' % (comp, d))
        if kind == "lagprod" and not (float(spec[2]) == int(spec[2]) and 0 <= int(spec[2]) <= MAX_LAG):
            raise SDEValueError('the lag of a "lagprod" observer must be an integer number of grid steps '
                                'in [0, %d]' % MAX_LAG)
        arr[i].kind, arr[i].comp = _OBS_KINDS[kind], comp
        arr[i].level = float(spec[2]) if need_level else 0.0
    return arr


def itoEuler(f, G, y0, tspan, dW=None, generator=None, **ens):
    """Euler-Maruyama for the Ito equation dy = f dt + G dW (integrate.py:190-240)."""
    kw = _ens(ens)
    if not isinstance(G, ops.DiffusionOp) and isinstance(G, (list, tuple)):
        raise TypeError("'tuple' object is not callable")   # the reference's behaviour, App. B
    return _run(SCHEME_EULER, f, G, y0, tspan, dW, None, False, None, generator, None, **kw)


def stratHeun(f, G, y0, tspan, dW=None, generator=None, **ens):
    """Stratonovich Heun (integrate.py:243-303)."""
    kw = _ens(ens)
    if not isinstance(G, ops.DiffusionOp) and isinstance(G, (list, tuple)):
        raise TypeError("'tuple' object is not callable")
    return _run(SCHEME_HEUN, f, G, y0, tspan, dW, None, True, None, generator, None, **kw)


def itoSRI2(f, G, y0, tspan, Imethod=Ikpw, dW=None, I=None, generator=None, **ens):
    """Roessler (2010) order 1.0 strong SRK scheme SRI2 for Ito equations
    (integrate.py:306-368, loop at :494-522)."""
    kw = _ens(ens)
    return _run(SCHEME_SRK2, f, G, y0, tspan, dW, I, False, Imethod, generator, None, **kw)


def stratSRS2(f, G, y0, tspan, Jmethod=Jkpw, dW=None, J=None, generator=None, **ens):
    """Roessler (2010) SRS2 for Stratonovich equations (integrate.py:371-433)."""
    kw = _ens(ens)
    return _run(SCHEME_SRK2, f, G, y0, tspan, dW, J, True, Jmethod, generator, None, **kw)


def stratKP2iS(f, G, y0, tspan, Jmethod=Jkpw, gam=None, al1=None, al2=None, rtol=1e-4, dW=None,
               J=None, generator=None, **ens):
    """Kloeden-Platen two-step implicit order 1.0 strong scheme (integrate.py:526-646).  The
    implicit equation is solved per path by Newton iteration with a forward-difference
    Jacobian to relative step tolerance `rtol` (the reference uses MINPACK hybrd with
    xtol=rtol); RuntimeError if it does not converge."""
    kw = _ens(ens)
    if not isinstance(G, ops.DiffusionOp) and isinstance(G, (list, tuple)):
        raise SDEValueError('G should be a function returning a d x m matrix.')
    y0a = np.asarray(y0)
    if np.iscomplexobj(y0a):
        raise SDEValueError("stratKP2iS() can't yet handle complex variables.")
    d = 1 if y0a.ndim == 0 else len(y0a)
    pars = []
    for v in (gam, al1, al2):
        v = np.full(d, 0.5) if v is None else np.asarray(v, dtype=np.float64).reshape(-1)
        if v.shape != (d,):
            raise SDEValueError('gam, al1, al2 must have shape (d,)')
        pars.append(v)
    return _run(SCHEME_KP2IS, f, G, y0, tspan, dW, J, True, Jmethod, generator,
                (pars[0], pars[1], pars[2], float(rtol)), **kw)


def itoKP2i(f, G, y0, tspan, Imethod=Ikpw, gam=None, al1=None, al2=None, rtol=1e-4, dW=None,
            I=None, generator=None, **ens):
    """The Ito version of the Kloeden-Platen two-step implicit order 1.0 strong scheme -- on the
    reference's to-do list (README.rst:125-128), not in it.  Kloeden & Platen (1992) sec. 12.4: the
    same recursion as stratKP2iS (integrate.py:603-645) with the Ito drift and the Ito integrals,
        V_n = G_n dW + (1/sqrt h) sum_{j1} (G(Ybar_j1) - G_n) I[j1, :],  Ybar_j1 = Y_n + f h + G_n[:, j1] sqrt h.
    Same arguments as stratKP2iS with Imethod / I in place of Jmethod / J."""
    kw = _ens(ens)
    if not isinstance(G, ops.DiffusionOp) and isinstance(G, (list, tuple)):
        raise SDEValueError('G should be a function returning a d x m matrix.')
    y0a = np.asarray(y0)
    if np.iscomplexobj(y0a):
        raise SDEValueError("itoKP2i() can't handle complex variables.")
    d = 1 if y0a.ndim == 0 else len(y0a)
    pars = []
    for v in (gam, al1, al2):
        v = np.full(d, 0.5) if v is None else np.asarray(v, dtype=np.float64).reshape(-1)
        if v.shape != (d,):
            raise SDEValueError('gam, al1, al2 must have shape (d,)')
        pars.append(v)
    return _run(SCHEME_KP2IS, f, G, y0, tspan, dW, I, False, Imethod, generator,
                (pars[0], pars[1], pars[2], float(rtol)), **kw)


def stratR2(f, G, y0, tspan, dW=None, generator=None, **ens):
    """Burrage & Burrage (1996) two-stage Runge-Kutta method R2 for Stratonovich equations -- the
    reference lists "Burrage and Burrage SRKs" as to-do (README.rst:125-128).  Strong order 1.0 for a
    single Wiener process or commutative noise with only dW (no repeated integrals): stage 2 at
    y + (2/3)(f h + G dW), weights (1/4, 3/4) for drift and diffusion alike.  Signature of stratHeun."""
    kw = _ens(ens)
    if not isinstance(G, ops.DiffusionOp) and isinstance(G, (list, tuple)):
        raise TypeError("'tuple' object is not callable")
    return _run(SCHEME_R2, f, G, y0, tspan, dW, None, True, None, generator, None, **kw)


def itoint(f, G, y0, tspan, generator=None, **ens):
    """Integrate the Ito equation dy = f(y,t)dt + G(y,t)dW (integrate.py:124-154): currently
    always itoSRI2, like the reference."""
    _check_args(f, G, y0, tspan, None, None)
    return itoSRI2(f, G, y0, tspan, generator=generator, **ens)


def stratint(f, G, y0, tspan, generator=None, **ens):
    """Integrate the Stratonovich equation dy = f dt + G o dW (integrate.py:157-187): always
    stratSRS2, like the reference."""
    _check_args(f, G, y0, tspan, None, None)
    return stratSRS2(f, G, y0, tspan, generator=generator, **ens)


def _ens(ens):
    bad = set(ens) - set(_ENS)
    if bad:
        raise TypeError("unexpected keyword argument(s): %s" % ", ".join(sorted(bad)))
    kw = dict(_ENS)
    kw.update(ens)
    return kw
