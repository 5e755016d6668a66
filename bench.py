#!/usr/bin/env python
"""bench.py -- SRI2 path-steps/s (d=m=10, fp64) on N B200s, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--paths P] [--impl reference]
                    [--no-configs]

Workload (BASELINE.json configs[2]; SURVEY.md 8d "cfg3"): SYS_LIN(10,10), itoSRI2, Ikpw n=5,
tspan = linspace(0, 1, 1001) (1000 time steps), P = 10^7 paths PER GPU (weak scaling),
on-device noise seed 20261019, save_every = 100 (11 saved points).  One bench "step" = one
pass of the hot path over that batch: P x 1000 path-steps, followed by the ensemble
moments + 64-bin histograms of the saved states and (N > 1) one NCCL all-reduce of them
(configs[4]).  Rank r integrates global path ids [r P, (r+1) P).

`value`  : path-steps/s, device-timed (CUDA events, max over ranks), nothing crossing PCIe.
`e2e`    : the same metric through the public call a user of the reference would make,
           sdeint_b200.itoSRI2(f, G, y0, tspan, paths=P, ...) -> host array: operator
           matrices, y0, tspan go host->device and the decimated trajectories (8.8 GB per
           10^7 paths) come back device->host inside the timed region (N = 1: straight into a
           pinned result array; N > 1: into a pageable one through the shim's fixed ring of two
           pinned 512 MB slots, so 10^7 paths per rank flow on every rank).
`configs`: (N = 1) the other single-GPU configurations of BASELINE.json, same warm-up / timing rule,
           each with its own roofline and e2e: configs[1] README 2-d linear SDE with itoEuler and
           stratHeun, 10^6 paths x 10^4 steps; configs[3] stratSRS2 + Jwik, d=m=20, 10^6 paths.
`--impl reference`: the UNMODIFIED reference (baseline/_ref, installed by build() from
           /root/reference) through its public API -- sdeint.itoSRI2(f, [g_1..g_m], y0, tspan,
           generator=default_rng(seed)), one path per call -- on all host cores; when that directory
           is absent, the oracle port (kind "port") and the line says so.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# NVRTC results of the benchmark run (the extension config) go to a scratch directory
import tempfile  # noqa: E402
os.environ.setdefault("SDB_CACHE_DIR", os.path.join(tempfile.gettempdir(), "sdeint_b200_jit_bench"))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")

D = M = 10
N_TERMS = 5
N_POINTS = 1001
SAVE_EVERY = 100
SEED = 20261019
HIST = (64, -1.0, 3.0)
FP64_PEAK_MICROBENCH = 37.1   # TFLOP/s, best register-resident DFMA rate (profiles/r01_microbench.md)


def flops_srk2(d, m):       # SURVEY 8d, integrate.py:494-522 operation by operation
    return 6 * m * d * d + 2 * d * m * m + 4 * d * d + 4 * d * m + 4 * d


def flops_ikpw(m, n):
    return m * (n + 4) + (m * (m - 1) // 2) * (5 * n + 5)


def flops_iwik(m, n):       # collapsed O(m^2) form, SURVEY App. A.4
    return m * (n + 2 * m + 5) + (m * (m - 1) // 2) * (5 * n + 17) + 3


def collapse(d, m):         # linear stage-2 collapse: subtract 2md^2 + 2dm
    return 2 * m * d * d + 2 * d * m


FLOPS_REFERENCE_FORM = flops_srk2(D, M) + flops_ikpw(M, N_TERMS)              # 10280
FLOPS_PER_PATH_STEP = FLOPS_REFERENCE_FORM - collapse(D, M)                   # 8080 executed
BYTES_PER_PATH_STEP = 8.0 * D / SAVE_EVERY + 8.0 * D / (N_POINTS - 1)         # 0.88 B (stores)
CPU_PATHS, CPU_STEPS = 16, 1000   # CPU sample per worker: ~1.2 s each, ~20 core-seconds on 16 cores
CPU_SAMPLES = 3                   # median of this many samples
METRIC = "SRI2 path-steps/s (d=m=10, fp64)"
UNIT = "path-steps/s"


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm
# ------------------------------------------------------------------------------------------------
def reference_available():
    return os.path.isfile(os.path.join(REF_DIR, "sdeint", "integrate.py"))


def _cpu_worker(job):
    seed, n_paths, n_steps, use_ref = job
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"
    from sdeint_b200 import ops
    fop, Gop, y0 = ops.sys_lin(D, M)
    A, B = np.array(fop.A), np.array(Gop.B)
    tspan = np.linspace(0.0, n_steps / 1000.0, n_steps + 1)
    if use_ref:
        # the unmodified reference, list-of-columns G (its faster form, integrate.py:330-334)
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        import sdeint

        def f(y, t):
            return A.dot(y)
        cols = [(lambda y, t, Bk=B[k]: Bk.dot(y)) for k in range(M)]
        t0 = time.perf_counter()
        for p in range(n_paths):
            sdeint.itoSRI2(f, cols, y0, tspan, generator=np.random.default_rng([seed, p]))
        return time.perf_counter() - t0
    from oracle import sde_oracle as orc
    cols = Gop.columns()
    h = (tspan[-1] - tspan[0]) / n_steps
    t0 = time.perf_counter()
    for p in range(n_paths):
        g = np.random.default_rng([seed, p])
        dW = orc.delta_w(n_steps, M, h, g)
        _, I = orc.ikpw(dW, h, n=N_TERMS, generator=g)
        orc.srk2(fop, cols, y0, tspan, dW, I)
    return time.perf_counter() - t0


class CpuArm:
    """All host cores, one process each (BLAS threads = 1), one path per call."""

    def __init__(self, cores=None):
        import multiprocessing as mp
        self.cores = cores or os.cpu_count() or 1
        self.use_ref = reference_available()
        self.kind = "reference" if self.use_ref else "port"
        self.pool = mp.get_context("spawn").Pool(self.cores)
        self.pool.map(_cpu_worker, [(1, 1, 20, self.use_ref)] * self.cores)      # imports, warm-up
        self.sample_text = ("%d workers x %d paths x %d steps of SYS_LIN(10,10) itoSRI2+Ikpw(n=5), %s"
                            % (self.cores, CPU_PATHS, CPU_STEPS,
                               "unmodified sdeint from baseline/_ref" if self.use_ref else
                               "oracle port (baseline/_ref absent: run __graft_entry__.build() "
                               "where /root/reference exists)"))

    def sample(self, tag=0):
        t0 = time.perf_counter()
        self.pool.map(_cpu_worker, [(100 + 1000 * tag + w, CPU_PATHS, CPU_STEPS, self.use_ref)
                                    for w in range(self.cores)])
        wall = time.perf_counter() - t0
        return self.cores * CPU_PATHS * CPU_STEPS / wall

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_baseline_block(samples=CPU_SAMPLES):
    arm = CpuArm()
    vals = [arm.sample(i) for i in range(samples)]
    arm.close()
    return {"value": float(statistics.median(vals)), "unit": UNIT, "cores": arm.cores,
            "kind": arm.kind, "sample": arm.sample_text + "; median of %d samples" % samples,
            "samples": [float(v) for v in vals]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    arm = CpuArm()
    vals = []
    for i in range(args.warmup + args.steps):
        v = arm.sample(i)
        if i >= args.warmup:
            vals.append(v)
    arm.close()
    value = float(statistics.median(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * arm.cores * CPU_PATHS * CPU_STEPS / value, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": arm.kind,
                         "sample": arm.sample_text + " per step; value = median over the timed steps",
                         "samples": [float(v) for v in vals]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def workload_config(args):
    return {
        "workload": "cfg3: SYS_LIN(10,10) itoSRI2 + on-device noise/Ikpw(n=5), "
                    "tspan=linspace(0,1,1001), %d paths per GPU, save_every=100, "
                    "moments+64-bin histograms of 11 saved points (+NCCL all-reduce for N>1)"
                    % args.paths,
        "paths_per_gpu": args.paths, "time_steps": N_POINTS - 1, "d": D, "m": M,
        "n_terms": N_TERMS, "save_every": SAVE_EVERY, "seed": SEED,
        "l2": "no resident inputs (noise is generated in registers); the %.1f GB of trajectory "
              "stores per step exceed the 126 MB L2" % (8e-9 * D * 11 * args.paths),
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx),
                "reasons": sorted(reasons), "samples": len(sm)}


def fp64_roofline(flops_exec, flops_ref_form, path_steps, kernel_ms, peak, peak_sus, symbol, note):
    achieved = flops_exec * path_steps / (kernel_ms * 1e-3) / 1e12
    return {
        "bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
        "frac": achieved / peak,
        "peak_source": "measured live: sdb_fp64_peak, register-resident DFMA chains on all SMs, "
                       "burst (4 ms launches); MEASURED_PEAKS.json has no fp64 entry, nominal 37.2",
        "peak_sustained": peak_sus,
        "peak_microbench": FP64_PEAK_MICROBENCH,
        "frac_vs_microbench_peak": achieved / FP64_PEAK_MICROBENCH,
        "kernel_symbol": symbol, "kernel_ms": kernel_ms,
        "flops_per_path_step": flops_exec, "flops_note": note,
        "frac_reference_form": flops_ref_form * path_steps / (kernel_ms * 1e-3) / 1e12 / peak,
    }


def run_side_config(sdb, torch, name, call, host_shape, path_steps, steps, warmup):
    """One of the other BASELINE.json configurations on this GPU: W warm-up + K timed device
    steps (CUDA events around the public call with out="torch"), then the e2e leg (host_out)."""
    for _ in range(warmup):
        y = call(out="torch")
        del y
    torch.cuda.synchronize()
    n0 = sdb.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(steps)]
    for a, b in ev:
        a.record()
        y = call(out="torch")
        b.record()
        del y
    torch.cuda.synchronize()
    symbol = sdb.last_kernel()
    launches = sdb.launch_count() - n0
    ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    host = torch.empty(host_shape, dtype=torch.float64, pin_memory=True)
    call(host_out=host)
    torch.cuda.synchronize()
    e2e_steps = max(1, min(steps, 2))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        yh = call(host_out=host)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / e2e_steps
    res = {"value": path_steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
           "warmup": warmup, "kernel_symbol": symbol, "gpu_launches": int(launches),
           "e2e": {"value": path_steps / wall, "unit": UNIT, "steps": e2e_steps,
                   "d2h_bytes_per_step": int(host.numel() * 8),
                   "checksum": float(np.asarray(yh[0, -1]).sum())}}
    del host, yh
    return name, res


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import sdeint_b200 as sdb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this benchmark has no CPU fallback; "
                         "use --impl reference for the CPU arm)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    P = args.paths
    f, G, y0 = sdb.ops.sys_lin(D, M)
    tspan = np.linspace(0.0, 1.0, N_POINTS)
    path_id0 = rank * P
    n_saved = (N_POINTS - 1) // SAVE_EVERY + 1

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step_device():
        y = sdb.itoSRI2(f, G, y0, tspan, paths=P, seed=SEED, path_id0=path_id0,
                        save_every=SAVE_EVERY, out="torch")
        sums, counts = sdb.ensemble.moments(y, hist=HIST)
        sdb.ensemble.allreduce_moments(sums, counts)
        return y, sums, counts

    # -- warm-up, then K timed steps bracketed by barrier + synchronize ---------------------
    for _ in range(args.warmup):
        y, sums, counts = step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = sdb.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    k1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    e0.record()
    for i in range(args.steps):
        k0[i].record()
        y = sdb.itoSRI2(f, G, y0, tspan, paths=P, seed=SEED, path_id0=path_id0,
                        save_every=SAVE_EVERY, out="torch")
        k1[i].record()                       # brackets the integrator launch (dominant kernel)
        kernel_symbol = sdb.last_kernel()    # what sdb_integrate dispatched to (include/sdb.h)
        sums, counts = sdb.ensemble.moments(y, hist=HIST)
        sdb.ensemble.allreduce_moments(sums, counts)
    e1.record()
    barrier()
    launches = sdb.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    ms_kernel = sum(a.elapsed_time(b) for a, b in zip(k0, k1)) / args.steps
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    path_steps = float(P) * (N_POINTS - 1)
    value = world * path_steps * args.steps / (ms_total * 1e-3)

    # sanity of the physics: E y(1) = expm(A) y0 for the Ito system
    mean, _var = sdb.ensemble.mean_var(sums, float(P * world))
    from scipy.linalg import expm
    want = expm(f.A * 1.0).dot(y0)
    mean_err = float(np.max(np.abs(mean[-1].cpu().numpy() - want)))
    # and E y_i(1)^2 from dP/dt = A P + P A^T + sum_k B_k P B_k^T, P(0) = y0 y0^T
    eye = np.eye(D)
    Lop = np.kron(f.A, eye) + np.kron(eye, f.A) + sum(np.kron(Bk, Bk) for Bk in G.B)
    P2 = expm(Lop).dot(np.outer(y0, y0).reshape(-1)).reshape(D, D)
    m2 = (sums[-1, :, 1] / float(P * world)).cpu().numpy()
    m2_err = float(np.max(np.abs(m2 - np.diag(P2)) / np.diag(P2)))
    del y

    # -- e2e: the public host-array call, copies inside the timed region ---------------------
    # Every rank integrates its full 10^7 paths.  One rank: the result array is pinned and the
    # chunks land in it directly.  Several ranks on one host: a pageable array per rank, filled
    # through the shim's ring of two pinned 512 MB slots (sdeint_b200/integrate.py:_run_pipelined).
    Pe = P
    host = torch.empty((n_saved, D, Pe), dtype=torch.float64, pin_memory=(world == 1))
    if world > 1:
        host.zero_()                                                         # touch the pages once
    e2e_steps = max(1, min(args.steps, 3))
    sdb.itoSRI2(f, G, y0, tspan, paths=Pe, seed=SEED, path_id0=path_id0, save_every=SAVE_EVERY,
                host_out=host)                                           # warm the path
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        yh = sdb.itoSRI2(f, G, y0, tspan, paths=Pe, seed=SEED, path_id0=path_id0,
                         save_every=SAVE_EVERY, host_out=host)
    barrier()
    wall = time.perf_counter() - t0
    tw = torch.tensor([wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
    e2e_val = world * float(Pe) * (N_POINTS - 1) * e2e_steps / float(tw.item())
    h2d = int(8 * (f.params.size + G.params.size + y0.size + tspan.size))
    d2h = int(host.numel() * 8 + 32)
    e2e = {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "steps": e2e_steps, "paths_per_call": Pe,
           "host_array": "pinned" if world == 1 else "pageable via 2 x 512 MB pinned ring",
           "checksum": float(np.asarray(yh[0, -1]).sum())}
    del host, yh

    if rank == 0:
        import ctypes as C
        peak, peak_sus = C.c_double(), C.c_double()
        sdb._lib.check(sdb._lib.load().sdb_fp64_peak(8000, C.byref(peak), None))       # ~4 ms launches
        sdb._lib.check(sdb._lib.load().sdb_fp64_peak(400000, C.byref(peak_sus), None))  # ~0.2 s launches
        peak, peak_sus = float(peak.value), float(peak_sus.value)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        roofline = fp64_roofline(
            FLOPS_PER_PATH_STEP, FLOPS_REFERENCE_FORM, path_steps, ms_kernel, peak, peak_sus,
            kernel_symbol,
            "executed flop: SURVEY 8d count 10280 minus 2md^2+2dm=2200 for the linear stage-2 "
            "collapse; normal transforms and random-bit generation not counted")
        roofline["kernel"] = "sdb fused SRI2 step loop (one launch per step)"
        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from one `ncu --set full`
        # capture of a 10^6-path launch with the bench's tspan / save_every: 880.4 MB against
        # 880.0 MB of algorithmic stores (profiles/r02_fused_v0_final_ncu_summary.txt); the traffic is
        # the decimated store and nothing else, so it scales with the paths.
        roofline["traffic"] = 880.38e6 * (P / 1e6)
        roofline["traffic_unit"] = "B per launch"
        roofline["traffic_source"] = ("ncu --set full at 1e6 paths (38.2 MB read + 842.2 MB "
                                      "written), scaled linearly to this launch")
        roofline["hbm"] = {"achieved": BYTES_PER_PATH_STEP * path_steps / (ms_kernel * 1e-3) / 1e9,
                           "peak": hbm_peak, "unit": "GB/s",
                           "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}
        roofline["hbm"]["frac"] = roofline["hbm"]["achieved"] / hbm_peak
        # Issue model: per warp-step (32 path-steps) of a sub-partition the SASS
        # executes 3709 DFMA-class x 2 + 280 DMMA x 16 = 11898 FP64-pipe cycles (ncu:
        # sm__pipe_shared_cycles_active 65.8 % = fp64 41.0 % + dmma 24.8 %); the two Philox blocks left
        # per path-step (2 x 74 cycles) are the only IMAD.WIDE work on that pipe.  An outer-product DFMA
        # (one multiplicand in the reuse cache, two fresh 64-bit operands) measures 2.42 cycles, not 2
        # (tools/micro/dfma_rf.cu): 3709 x 2.42 + 4480 = 13456 cycles of pipe + operand ports.
        sm_hz = 1e6 * float((clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0))
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        floor = sms * 4 * 32 * sm_hz / (11898.0 + 148.0)
        floor_rf = sms * 4 * 32 * sm_hz / (13456.0 + 148.0)
        roofline["issue_model"] = {
            "fp64_pipe_cycles_per_warp_step": 11898, "philox_cycles_per_warp_step": 148,
            "floor_path_steps_per_s": floor,
            "frac_of_floor": (path_steps / (ms_kernel * 1e-3)) / floor,
            "fp64_pipe_and_operand_port_cycles_per_warp_step": 13456,
            "frac_of_operand_port_floor": (path_steps / (ms_kernel * 1e-3)) / floor_rf,
            "ncu_pipe_shared_cycles_active_pct": 65.8,
            "dmma_note": "the 4480 DMMA cycles do not co-issue with other instructions "
                         "(tools/micro/dmma_coissue.cu)",
            "source": "SASS instruction counts (profiles/r02_fused_v0_final_ncu_summary.txt) "
                      "x measured issue costs (tools/micro/rng.cu, dmma_coissue.cu, dfma_rf.cu, philox.cu, coissue.cu)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args), "clocks": clocks,
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "check": {"max_abs_err_of_ensemble_mean_vs_expm": mean_err,
                      "max_rel_err_of_second_moment_vs_moment_ode": m2_err},
        }
        if world == 1 and not args.no_configs:
            line["configs"] = side_configs(sdb, torch, args, peak, peak_sus, hbm_peak)
        if world == 1:
            line["cpu_baseline"] = cpu_baseline_block()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def side_configs(sdb, torch, args, peak, peak_sus, hbm_peak):
    """BASELINE.json configs[1] and configs[3] on this GPU (cfg2, cfg4)."""
    out = {}
    K, W = max(1, min(args.steps, 3)), 3
    # cfg2: README 2-d linear Ito SDE (README.rst:80-97), 10^6 paths x 10^4 steps, save_every=10
    f2, G2, y02 = sdb.ops.sys_lin2()
    P2, T2, se2 = 10 ** 6, 10000, 10
    ts2 = np.linspace(0.0, 10.0, T2 + 1)
    ns2 = T2 // se2 + 1
    for tag, fn, flops in (("cfg2_itoEuler", sdb.itoEuler, 18), ("cfg2_stratHeun", sdb.stratHeun, 40)):
        name, res = run_side_config(
            sdb, torch, tag,
            lambda fn=fn, **kw: fn(f2, G2, y02, ts2, paths=P2, seed=SEED, save_every=se2, **kw),
            (ns2, 2, P2), float(P2) * T2, K, W)
        res["workload"] = ("README 2-d linear SDE (A y dt + B dW, additive), %s, %d paths x %d steps, "
                           "save_every=%d (%.0f GB of stores per step > L2)"
                           % (tag.split("_")[1], P2, T2, se2, 8e-9 * ns2 * 2 * P2))
        ps = res["value"]
        res["roofline"] = {
            "bound": "generator (1 Philox block + 34 FP64 instructions of Box-Muller per path-step; "
                     "FP64 and HBM fractions reported)",
            "fp64": {"achieved": ps * flops / 1e12, "peak": peak, "unit": "TFLOP/s",
                     "frac": ps * flops / 1e12 / peak, "flops_per_path_step": flops},
            "hbm": {"achieved": ps * 16.0 / se2 / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": ps * 16.0 / se2 / 1e9 / hbm_peak, "bytes_per_path_step": 16.0 / se2},
            "kernel_symbol": res["kernel_symbol"], "kernel_ms": res["ms_per_step"]}
        out[name] = res
    # cfg4: SYS_LIN(20,20), stratSRS2 + Jwik n=5 (wiener.py:234-305), 10^6 paths x 1000 steps
    f4, G4, y04 = sdb.ops.sys_lin(20, 20)
    P4, T4, se4 = 10 ** 6, 1000, 100
    ts4 = np.linspace(0.0, 1.0, T4 + 1)
    name, res = run_side_config(
        sdb, torch, "cfg4_stratSRS2_Jwik",
        lambda **kw: sdb.stratSRS2(f4, G4, y04, ts4, sdb.Jwik, paths=P4, seed=SEED, save_every=se4, **kw),
        (T4 // se4 + 1, 20, P4), float(P4) * T4, K, W)
    ref_form = flops_srk2(20, 20) + flops_iwik(20, N_TERMS)                        # 76263
    res["workload"] = "SYS_LIN(20,20) stratSRS2 + Jwik(n=5), %d paths x %d steps, save_every=%d" % (P4, T4, se4)
    res["roofline"] = fp64_roofline(
        ref_form - collapse(20, 20), ref_form, float(P4) * T4, res["ms_per_step"], peak, peak_sus,
        res["kernel_symbol"],
        "executed flop: SURVEY 8d count 76263 minus 2md^2+2dm=16800 for the linear stage-2 collapse")
    out[name] = res
    # Extension beyond BASELINE.json's configs: a NONLINEAR drift given as source
    # (NVRTC) with the linear diffusion of cfg3 -- the fused kernel with the drift evaluated per lane.
    fL, GL, y0L = sdb.ops.sys_lin(D, M)
    fU = sdb.ops.ExprDrift(["-p[0]*y[%d] + p[1]*y[%d] - p[2]*y[%d]*y[%d]*y[%d]" % (i, (i + 1) % D, i, i, i)
                            for i in range(D)], params=[1.5, 0.3, 0.2])
    PU, TU, seU = 10 ** 6, 1000, 100
    tsU = np.linspace(0.0, 1.0, TU + 1)
    name, res = run_side_config(
        sdb, torch, "ext_user_drift_itoSRI2",
        lambda **kw: sdb.itoSRI2(fU, GL, y0L, tsU, paths=PU, seed=SEED, save_every=seU, commutative=False, **kw),
        (TU // seU + 1, D, PU), float(PU) * TU, K, W)
    res["workload"] = ("cubic drift f_i = -1.5 y_i + 0.3 y_(i+1) - 0.2 y_i^3 (ExprDrift -> NVRTC) + g_k = B_k y of "
                       "SYS_LIN(10,10), itoSRI2 + Ikpw(n=5), %d paths x %d steps, save_every=%d" % (PU, TU, seU))
    f_user = 7 * D                                      # flop of one evaluation of this drift
    ref_form = flops_srk2(D, M) + flops_ikpw(M, N_TERMS) - 4 * D * D + 2 * f_user
    res["roofline"] = fp64_roofline(
        ref_form - collapse(D, M), ref_form, float(PU) * TU, res["ms_per_step"], peak, peak_sus,
        res["kernel_symbol"],
        "executed flop: SURVEY 8d count 10280 with the two dense drift products (4 d^2) replaced by two "
        "evaluations of the user's drift (7 d flop each), minus 2md^2+2dm for the linear stage-2 collapse")
    out[name] = res
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--paths", type=int, default=10 ** 7, help="paths per GPU")
    ap.add_argument("--impl", default="sdb", choices=["sdb", "reference"])
    ap.add_argument("--no-configs", action="store_true",
                    help="skip the cfg2 / cfg4 sub-objects (N = 1 only)")
    args = ap.parse_args()
    if args.impl != "reference" and args.warmup < 3:
        # timing rule: at least three untimed steps before the timed region (clocks, caches, module load)
        print("bench.py: --warmup %d raised to 3" % args.warmup, file=sys.stderr)
        args.warmup = 3
    if args.steps < 1:
        ap.error("--steps must be >= 1")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
