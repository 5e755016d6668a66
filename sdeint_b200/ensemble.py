"""Ensemble statistics and multi-GPU sharding.

Paths are independent (sdeint/integrate.py:494-522 reads only a path's own state and noise),
so an ensemble shards across ranks with no exchange during stepping: rank r of R integrates
global path ids [r*P/R, (r+1)*P/R) and Philox counters are keyed by the GLOBAL path id, so the
trajectory of path p is identical for any R.  The only collective is one all-reduce (NCCL
over NVLink through torch.distributed) of the moment sums and histogram counts."""
from __future__ import annotations

import numpy as np

from . import _device as dev
from . import _lib


def shard_paths(paths, rank, world_size):
    """Contiguous block partition of global path ids: returns (path_id0, count)."""
    paths, rank, world_size = int(paths), int(rank), int(world_size)
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(paths, world_size)
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def moments(y_dev, hist=None):
    """Per (saved point, component): sum y, sum y^2 (and optional histogram counts) of a
    device trajectory tensor in path-minor layout (n_saved, d, P).

    hist = (bins, lo, hi).  Returns (sums (n_saved, d, 2) float64 tensor,
    counts (n_saved, d, bins) int64 tensor or None), both on the device."""
    torch = dev.torch_mod()
    if y_dev.dim() != 3 or y_dev.dtype != torch.float64 or not y_dev.is_contiguous():
        raise ValueError("moments: need a contiguous float64 tensor (n_saved, d, P)")
    S, d, P = y_dev.shape
    with torch.cuda.device(y_dev.device):
        sums = torch.zeros((S, d, 2), dtype=torch.float64, device=y_dev.device)
        counts = None
        bins, lo, hi = 0, 0.0, 1.0
        if hist is not None:
            bins, lo, hi = int(hist[0]), float(hist[1]), float(hist[2])
            counts = torch.zeros((S, d, bins), dtype=torch.int64, device=y_dev.device)
        _lib.check(_lib.load().sdb_moments(dev.ptr(y_dev), S, d, P, dev.ptr(sums),
                                           dev.ptr(counts), bins, lo, hi, dev.stream_ptr()))
    return sums, counts


def welford(y_dev, into=None):
    """(count, mean, M2 = sum (y - mean)^2) per (saved point, component) of a device trajectory tensor
    (n_saved, d, P): float64 device tensor (n_saved, d, 3).  Unlike E[y^2] - E[y]^2 this does not
    cancel when |mean| >> std: every chunk of 4096 paths is centred on its own mean and chunks are
    merged in a fixed order with Chan's update (sdb_moments_welford).  `into` = a previous result to
    merge this tensor's paths into (successive batches of one ensemble)."""
    torch = dev.torch_mod()
    if y_dev.dim() != 3 or y_dev.dtype != torch.float64 or not y_dev.is_contiguous():
        raise ValueError("welford: need a contiguous float64 tensor (n_saved, d, P)")
    S, d, P = y_dev.shape
    with torch.cuda.device(y_dev.device):
        out = into if into is not None else torch.zeros((S, d, 3), dtype=torch.float64, device=y_dev.device)
        _lib.check(_lib.load().sdb_moments_welford(dev.ptr(y_dev), S, d, P, dev.ptr(out), dev.stream_ptr()))
    return out


def chan_merge(parts):
    """Fixed-order Chan merge of (count, mean, M2) partials (arrays or tensors (..., 3), e.g. one per
    rank in rank order): the bit-reproducible combination SURVEY 8e asks for."""
    acc = None
    for part in parts:
        if acc is None:
            acc = part.clone() if hasattr(part, "clone") else np.array(part, dtype=np.float64)
            continue
        n, mean, m2 = acc[..., 0], acc[..., 1], acc[..., 2]
        nb, mb, qb = part[..., 0], part[..., 1], part[..., 2]
        tot = n + nb
        safe = tot + (tot == 0)
        dlt = mb - mean
        new_mean = mean + dlt * (nb / safe)
        new_m2 = m2 + qb + dlt * dlt * (n * nb / safe)
        acc = acc.clone() if hasattr(acc, "clone") else acc.copy()
        acc[..., 0], acc[..., 1], acc[..., 2] = tot, new_mean, new_m2
    return acc


def allgather_welford(w, group=None):
    """All ranks' (count, mean, M2) partials gathered (torch.distributed; NCCL on GPUs) and merged in
    rank order: every rank gets the same bits."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        torch = dev.torch_mod()
        parts = [torch.empty_like(w) for _ in range(dist.get_world_size(group))]
        dist.all_gather(parts, w.contiguous(), group=group)
        return chan_merge(parts)
    return w


def welford_mean_var(w, ddof=0):
    """Ensemble mean and variance from (count, mean, M2)."""
    return w[..., 1], w[..., 2] / (w[..., 0] - ddof)


def second_moments(y_dev):
    """sum_p y_i y_j at every saved point of a device trajectory tensor (n_saved, d, P), d <= 16:
    returns a float64 device tensor (n_saved, d, d) (deterministic; all-reduce it like the sums)."""
    torch = dev.torch_mod()
    if y_dev.dim() != 3 or y_dev.dtype != torch.float64 or not y_dev.is_contiguous():
        raise ValueError("second_moments: need a contiguous float64 tensor (n_saved, d, P)")
    S, d, P = y_dev.shape
    with torch.cuda.device(y_dev.device):
        out = torch.zeros((S, d, d), dtype=torch.float64, device=y_dev.device)
        _lib.check(_lib.load().sdb_cross_moments(dev.ptr(y_dev), S, d, P, dev.ptr(out), dev.stream_ptr()))
    return out


def covariance(sums, cross, paths):
    """Ensemble covariance matrices (n_saved, d, d) from the moment sums and the second moments."""
    mean = sums[..., 0] / paths
    return cross / paths - mean[..., :, None] * mean[..., None, :]


def quantiles(counts, lo, hi, q):
    """Quantiles from the on-device histograms (n_saved, d, bins) on [lo, hi): linear interpolation of
    the empirical CDF inside the bin (resolution = the bin width).  q: scalar or sequence in (0, 1)."""
    c = np.asarray(counts.cpu() if hasattr(counts, "cpu") else counts, dtype=np.float64)
    q = np.atleast_1d(np.asarray(q, dtype=np.float64))
    bins = c.shape[-1]
    width = (hi - lo) / bins
    cdf = np.cumsum(c, axis=-1)
    total = cdf[..., -1:]
    out = np.empty(c.shape[:-1] + (q.size,))
    for k, qq in enumerate(q):
        target = qq * total[..., 0]
        idx = np.minimum((cdf < target[..., None]).sum(axis=-1), bins - 1)
        below = np.where(idx > 0, np.take_along_axis(cdf, np.maximum(idx - 1, 0)[..., None], -1)[..., 0], 0.0)
        inbin = np.take_along_axis(c, idx[..., None], -1)[..., 0]
        frac = np.where(inbin > 0, (target - below) / np.where(inbin > 0, inbin, 1.0), 0.5)
        out[..., k] = lo + (idx + np.clip(frac, 0.0, 1.0)) * width
    return out


def first_passage_cdf(times_dev, bins, t0, t1, paths, group=None):
    """Distribution of first-passage times from the in-loop observers (`observe=("first_above", ...)`
    with out="torch": a device tensor (n_obs, P), +inf where the level was never reached): the
    fraction of ALL paths that have hit by each of `bins` equal sub-intervals of [t0, t1), histogrammed
    on the device (sdb_moments) and summed over the ranks of `group`.  `paths` is the global ensemble
    size.  Returns (right bin edges (bins,), cdf (n_obs, bins)) as numpy arrays; 1 - cdf is the
    survival function."""
    torch = dev.torch_mod()
    if times_dev.dim() != 2 or times_dev.dtype != torch.float64:
        raise ValueError("first_passage_cdf: need a float64 device tensor (n_obs, P)")
    n_obs, P = times_dev.shape
    _, counts = moments(times_dev.contiguous().view(n_obs, 1, P), hist=(int(bins), float(t0), float(t1)))
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    c = counts[:, 0, :].cpu().numpy().astype(np.float64)
    edges = t0 + (t1 - t0) * np.arange(1, int(bins) + 1) / int(bins)
    return edges, np.cumsum(c, axis=1) / float(paths)


def stationary_autocovariance(lag_sums, lags, n_points, mean=0.0, group=None):
    """Autocovariance of a stationary component at FULL time resolution from the in-loop
    lagged-product observers (`observe=[("lagprod", comp, L), ...]`: per path sum_{n >= L} y_n y_{n-L}
    over the n_points grid points): C(L) = E[sum] / (n_points - L) - mean^2, the expectation over the
    paths (and over the ranks of `group`).  lag_sums: (P, n_lags) array or device tensor (n_lags, P)."""
    torch = dev.torch_mod()
    if isinstance(lag_sums, torch.Tensor):
        tot = lag_sums.sum(dim=1).to(torch.float64)
        cnt = torch.tensor([float(lag_sums.shape[1])], dtype=torch.float64, device=lag_sums.device)
    else:
        arr = np.asarray(lag_sums, dtype=np.float64)
        tot = torch.from_numpy(arr.sum(axis=0))
        cnt = torch.tensor([float(arr.shape[0])], dtype=torch.float64)
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
    lags = np.asarray(lags, dtype=np.float64)
    return tot.cpu().numpy() / float(cnt.item()) / (float(n_points) - lags) - float(mean) ** 2


def autocovariance(y_dev, comp, paths=None, group=None):
    """Ensemble autocovariance of one component across the saved times of a device trajectory tensor
    (n_saved, d, P):  C[s, t] = E[(y_s - E y_s)(y_t - E y_t)], the expectation over paths.
    A plain fp64 library GEMM (X X^T through torch/cuBLAS -- analysis plumbing, not the integration
    path) on the raw sums, so that ranks can be combined exactly like the moments: the (n_saved,
    n_saved) Gram matrix and the (n_saved,) sums are all-reduced over `group` before centring.
    `paths` is the global ensemble size (default: this tensor's P).  Returns a numpy array."""
    torch = dev.torch_mod()
    if y_dev.dim() != 3 or y_dev.dtype != torch.float64:
        raise ValueError("autocovariance: need a float64 device tensor (n_saved, d, P)")
    X = y_dev[:, int(comp), :]
    gram = X @ X.T
    tot = X.sum(dim=1)
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(gram, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=group)
    n = float(paths if paths is not None else y_dev.shape[2])
    mean = tot / n
    return (gram / n - mean[:, None] * mean[None, :]).cpu().numpy()


def periodogram(y_dev, comp, dt, first=0, paths=None, group=None):
    """Ensemble-averaged periodogram of one component over the saved times `first:` (uniform spacing
    `dt`), an estimate of the two-sided power spectral density at the non-negative frequencies:
        S(f_k) = dt / n * E |sum_j (y_j - mean) exp(-2 pi i j k / n)|^2,   f_k = k / (n dt).
    The spectral check the reference leaves as a to-do (sdeint/tests/test_integrate.py:455-456,
    :479-480).  torch.fft along the time axis of the path-minor tensor (library plumbing); the sum
    over paths is all-reduced over `group`.  Returns (freqs (n//2+1,), S (n//2+1,)) numpy arrays."""
    torch = dev.torch_mod()
    X = y_dev[int(first):, int(comp), :]
    n = X.shape[0]
    X = X - X.mean(dim=0, keepdim=True)                 # per-path mean removed (stationary segment)
    power = (torch.fft.rfft(X, dim=0).abs() ** 2).sum(dim=1)
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(power, op=dist.ReduceOp.SUM, group=group)
    total = float(paths if paths is not None else y_dev.shape[2])
    S = (power * (float(dt) / n / total)).cpu().numpy()
    return np.arange(n // 2 + 1) / (n * float(dt)), S


def checkpointed(integrator, f, G, y0, tspan, *, paths, seed, segment, checkpoint, on_segment=None,
                 **kw):
    """Integrate a long run in time segments of `segment` intervals with a checkpoint after each, and
    resume from the checkpoint file if there is one -- same final states, bit for bit, as one call
    (the noise depends only on (seed, path id, step counter); see `step0=` in sdeint_b200.integrate).

    integrator   itoEuler / stratHeun / itoSRI2 / stratSRS2 / itoint / stratint (not the two-step KP2iS)
    checkpoint   path of an .npz file holding (step, states, seed, paths, grid fingerprint); written
                 atomically after every segment, read on entry when it exists and matches
    on_segment   optional callback(step_done, states_dev (paths, d) device view) after each segment,
                 e.g. to accumulate statistics (`moments`) without storing trajectories
    Returns the final states as a numpy array (paths, d)."""
    import os
    torch = dev.torch_mod()
    tspan = np.asarray(tspan, dtype=np.float64)
    n_int = len(tspan) - 1
    segment = int(segment)
    if segment < 1:
        raise ValueError("checkpointed: segment must be >= 1")
    h = (tspan[-1] - tspan[0]) / n_int
    finger = np.array([tspan[0], tspan[-1], float(n_int)])
    # what the run IS: integrator, operators (kinds, parameter values, source), y0, grid and the options
    # that change the numbers -- a checkpoint of anything else must not be resumed
    import hashlib
    hh = hashlib.sha256()
    hh.update(getattr(integrator, "__name__", repr(integrator)).encode())
    for op in (f, G):
        hh.update(repr((int(op.kind), getattr(op, "source", None))).encode())
        hh.update(np.ascontiguousarray(op.params, dtype=np.float64).tobytes())
    hh.update(np.ascontiguousarray(np.atleast_1d(np.asarray(y0, dtype=np.float64))).tobytes())
    hh.update(tspan.tobytes())
    hh.update(repr(sorted((k, getattr(v, "__name__", v)) for k, v in kw.items() if k != "device")).encode())
    run_id = hh.hexdigest()
    step, states = 0, None
    if os.path.exists(checkpoint):
        with np.load(checkpoint) as ck:
            same = (int(ck["seed"]) == int(seed) and int(ck["paths"]) == int(paths)
                    and np.array_equal(ck["grid"], finger)
                    and "run_id" in ck.files and str(ck["run_id"]) == run_id)
            if not same:
                raise ValueError("checkpointed: %s is the checkpoint of a different run (system, integrator, "
                                 "y0, grid, seed, paths or options differ); remove it to start over"
                                 % checkpoint)
            step, states = int(ck["step"]), ck["states"]
    managed = {"save_every", "step0", "y0_paths", "out", "h_noise", "host_out", "observe", "dW", "I", "J"}
    if managed & set(kw):
        raise ValueError("checkpointed manages %s itself" % ", ".join(sorted(managed & set(kw))))
    opts = dict(kw)
    opts.update(paths=paths, seed=seed, out="torch", h_noise=h)
    cur = None if states is None else dev.to_device(np.ascontiguousarray(states), dev.require_cuda(kw.get("device")))
    while step < n_int:
        stop = min(step + segment, n_int)
        y = integrator(f, G, y0, tspan[step:stop + 1], save_every=stop - step, step0=step,
                       y0_paths=cur, **opts)                     # (2, d, P) on the device
        cur = y[-1].T                                            # (P, d) view, no copy
        step = stop
        if on_segment is not None:
            on_segment(step, cur)
        tmp = checkpoint + ".tmp.npz"
        with open(tmp, "wb") as fh:
            np.savez(fh, step=step, states=cur.contiguous().cpu().numpy(), seed=int(seed),
                     paths=int(paths), grid=finger, run_id=np.array(run_id))
            fh.flush()
            os.fsync(fh.fileno())
        os.replace(tmp, checkpoint)
    return cur.contiguous().cpu().numpy() if cur is not None else np.asarray(states)


def allreduce_moments(sums, counts=None, group=None, deterministic=False):
    """Sum the per-rank partials over all ranks (torch.distributed; NCCL on GPUs), in place.

    The floating-point sums of an all-reduce depend on the reduction order the collective picks.
    `deterministic=True` all-gathers the per-rank partials instead and adds them in rank order, so
    the result is bit-reproducible for a given number of ranks; the integer histogram
    counts are exact either way."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        if deterministic:
            torch = dev.torch_mod()
            parts = [torch.empty_like(sums) for _ in range(dist.get_world_size(group))]
            dist.all_gather(parts, sums.contiguous(), group=group)
            total = parts[0].clone()
            for part in parts[1:]:
                total += part
            sums.copy_(total)
        else:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
        if counts is not None:
            dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return sums, counts


def mean_var(sums, paths):
    """Ensemble mean and (population) variance from (sum, sum of squares)."""
    s = sums[..., 0] / paths
    q = sums[..., 1] / paths
    return s, q - s * s


def merge_moments_host(partials):
    """Fixed-order merge of per-rank (count, sums) partials on the host (numpy): the
    reference semantics of the all-reduce, used by the CPU tests."""
    total = 0
    acc = None
    for count, sums in partials:
        total += int(count)
        acc = np.array(sums, dtype=np.float64) if acc is None else acc + sums
    return total, acc
