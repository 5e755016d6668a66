#!/usr/bin/env python
"""bench.py -- SRI2 path-steps/s (d=m=10, fp64) on N B200s, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--paths P] [--impl reference]

Workload (BASELINE.json configs[2]; SURVEY.md 8d "cfg3"): SYS_LIN(10,10), itoSRI2, Ikpw n=5,
tspan = linspace(0, 1, 1001) (1000 time steps), P = 10^7 paths PER GPU (weak scaling),
on-device Philox seed 20261019, save_every = 100 (11 saved points).  One bench "step" = one
pass of the hot path over that batch: P x 1000 path-steps, followed by the ensemble
moments + 64-bin histograms of the saved states and (N > 1) one NCCL all-reduce of them
(configs[4]).  Rank r integrates global path ids [r P, (r+1) P).

`value`  : path-steps/s, device-timed (CUDA events, max over ranks), nothing crossing PCIe.
`e2e`    : the same metric through the public call a user of the reference would make,
           sdeint_b200.itoSRI2(f, G, y0, tspan, paths=P, ...) -> host array: operator
           matrices, y0, tspan go host->device and the decimated trajectories (8.8 GB per
           10^7 paths) come back device->host inside the timed region.
`--impl reference`: the oracle port of the reference's CPU path (one path per call, Python
           per-step loop, numpy; oracle/sde_oracle.py) on all host cores, bounded sample.
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

D = M = 10
N_TERMS = 5
N_POINTS = 1001
SAVE_EVERY = 100
SEED = 20261019
HIST = (64, -1.0, 3.0)
FLOPS_SRK2 = 6 * M * D * D + 2 * D * M * M + 4 * D * D + 4 * D * M + 4 * D   # 8840
FLOPS_IKPW = M * (N_TERMS + 4) + (M * (M - 1) // 2) * (5 * N_TERMS + 5)      # 1440
FLOPS_REFERENCE_FORM = FLOPS_SRK2 + FLOPS_IKPW                                # 10280
# The fused kernel collapses the two stage-2 evaluations of the LINEAR g_k into one product
# B_k (sqrt(h) s_k): SURVEY 8d says to subtract 2md^2 + 2dm from the numerator in that case.
FLOPS_PER_PATH_STEP = FLOPS_REFERENCE_FORM - (2 * M * D * D + 2 * D * M)      # 8080 executed
BYTES_PER_PATH_STEP = 8.0 * D / SAVE_EVERY + 8.0 * D / (N_POINTS - 1)         # 0.88 B (stores)
CPU_PATHS, CPU_STEPS = 16, 1000   # CPU sample per worker: ~1.2 s each, ~20 core-seconds on 16 cores
METRIC = "SRI2 path-steps/s (d=m=10, fp64)"
UNIT = "path-steps/s"


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port, one path per call like the reference
# ------------------------------------------------------------------------------------------------
def _cpu_worker(job):
    seed, n_paths, n_steps = job
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"
    from oracle import sde_oracle as orc
    from sdeint_b200 import ops
    f, G, y0 = ops.sys_lin(D, M)
    cols = G.columns()
    tspan = np.linspace(0.0, n_steps / 1000.0, n_steps + 1)
    h = (tspan[-1] - tspan[0]) / n_steps
    t0 = time.perf_counter()
    for p in range(n_paths):
        g = np.random.default_rng([seed, p])
        dW = orc.delta_w(n_steps, M, h, g)
        _, I = orc.ikpw(dW, h, n=N_TERMS, generator=g)
        orc.srk2(f, cols, y0, tspan, dW, I)
    return time.perf_counter() - t0


def cpu_path_steps_per_s(paths_per_worker, n_steps, cores=None):
    """All host cores, one process each (BLAS threads = 1): list-of-columns G, the
    reference's fastest form (integrate.py:330-334).  Returns (path-steps/s, cores, sample)."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(1, 1, 20)] * cores)                       # warm-up / imports
        t0 = time.perf_counter()
        pool.map(_cpu_worker, [(100 + w, paths_per_worker, n_steps) for w in range(cores)])
        wall = time.perf_counter() - t0
    total = cores * paths_per_worker * n_steps
    sample = "%d workers x %d paths x %d steps of SYS_LIN(10,10) itoSRI2+Ikpw(n=5)" % (
        cores, paths_per_worker, n_steps)
    return total / wall, cores, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    cores = os.cpu_count() or 1
    sample = ""
    for i in range(args.warmup + args.steps):
        v, cores, sample = cpu_path_steps_per_s(CPU_PATHS, CPU_STEPS, cores)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * cores * CPU_PATHS * CPU_STEPS / value, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample + " per step"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def workload_config(args):
    return {
        "workload": "cfg3: SYS_LIN(10,10) itoSRI2 + on-device Philox/Ikpw(n=5), "
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

    # -- e2e: the public host-array call, copies inside the timed region ---------------------
    # With several ranks on one host the pinned result buffers add up (8.8 GB per 10^7 paths and
    # rank), so each call then integrates at most 2*10^6 paths; the rate is what is reported.
    Pe = P if world == 1 else min(P, 2 * 10 ** 6)
    host = torch.empty((n_saved, D, Pe), dtype=torch.float64).pin_memory()
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
           "checksum": float(np.asarray(yh[0, -1]).sum())}

    if rank == 0:
        import ctypes as C
        peak, peak_sus = C.c_double(), C.c_double()
        sdb._lib.check(sdb._lib.load().sdb_fp64_peak(8000, C.byref(peak), None))       # ~4 ms launches
        sdb._lib.check(sdb._lib.load().sdb_fp64_peak(400000, C.byref(peak_sus), None))  # ~0.2 s launches
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        achieved_tf = FLOPS_PER_PATH_STEP * path_steps / (ms_kernel * 1e-3) / 1e12
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        roofline = {
            "bound": "fp64",
            "achieved": achieved_tf, "peak": float(peak.value), "unit": "TFLOP/s",
            "frac": achieved_tf / float(peak.value),
            "peak_source": "measured live: sdb_fp64_peak, register-resident DFMA chains on all "
                           "SMs, burst (4 ms launches); MEASURED_PEAKS.json has no fp64 entry, "
                           "nominal 37.2",
            "peak_sustained": float(peak_sus.value),
            "kernel": "sdb fused SRI2 step loop (one launch per step)",
            "kernel_symbol": kernel_symbol,
            "kernel_ms": ms_kernel,
            "flops_per_path_step": FLOPS_PER_PATH_STEP,
            "flops_note": "executed flop: SURVEY 8d count 10280 minus 2md^2+2dm=2200 for the "
                          "linear stage-2 collapse; normal transforms and Philox not counted",
            "frac_reference_form": (FLOPS_REFERENCE_FORM * path_steps / (ms_kernel * 1e-3) / 1e12
                                    / float(peak.value)),
            # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from one `ncu --set full`
            # capture of a 10^6-path launch with the bench's tspan / save_every: 880.8 MB against
            # 880.0 MB of algorithmic stores (profiles/r01_fused_v0_bench1e6_se100_ncu_summary.txt);
            # the traffic is the decimated store and nothing else, so it scales with the paths.
            "traffic": 880.82e6 * (P / 1e6),
            "traffic_unit": "B per launch",
            "traffic_source": "ncu --set full at 1e6 paths (37.7 MB read + 843.1 MB written), "
                              "scaled linearly to this launch",
            "hbm": {"achieved": BYTES_PER_PATH_STEP * path_steps / (ms_kernel * 1e-3) / 1e9,
                    "peak": hbm_peak, "unit": "GB/s",
                    "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
        }
        roofline["hbm"]["frac"] = roofline["hbm"]["achieved"] / hbm_peak
        # What the FP64 number does not see (profiles/r01_microbench.md): the 20
        # IMAD.WIDE of each Philox4x32-10 block issue at ~3.7 cycles each and do not overlap with
        # DFMA/DMMA.  Per warp-step (32 path-steps) of a sub-partition: 4040 DFMA-class x 2
        # + 280 DMMA x 16 = 12560 FP64-pipe cycles, + 55 Philox blocks x 74 = 4070 cycles.
        sm_hz = 1e6 * float((clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0))
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        floor = sms * 4 * 32 * sm_hz / (12560.0 + 4070.0)
        roofline["issue_model"] = {
            "fp64_pipe_cycles_per_warp_step": 12560, "philox_cycles_per_warp_step": 4070,
            "floor_path_steps_per_s": floor,
            "frac_of_floor": (path_steps / (ms_kernel * 1e-3)) / floor,
            "source": "SASS instruction counts (profiles/r01_fused_v0_bench1e6_ncu_summary.txt) "
                      "x measured issue costs (tools/micro/philox.cu, coissue.cu)"}
        cpu_v, cores, sample = (cpu_path_steps_per_s(CPU_PATHS, CPU_STEPS) if world == 1
                                else (None, None, None))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args), "clocks": clocks,
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "check": {"max_abs_err_of_ensemble_mean_vs_expm": mean_err,
                      "max_rel_err_of_second_moment_vs_moment_ode": m2_err},
        }
        if cpu_v is not None:
            line["cpu_baseline"] = {"value": cpu_v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--paths", type=int, default=10 ** 7, help="paths per GPU")
    ap.add_argument("--impl", default="sdb", choices=["sdb", "reference"])
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
