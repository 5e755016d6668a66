"""Compact summary of an .ncu-rep: python tools/ncu_summary.py file.ncu-rep [kernel-index]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
EXACT = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg.per_second",
    "smsp__inst_executed.sum", "smsp__inst_executed.avg.per_cycle_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "idc__request_hit_rate.pct", "idc__request_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__pcsamp_sample_buffer_full", "sm__icc_requests_lookup_miss.sum", "sm__icc_requests.sum",
    "gcc__requests_lookup_miss.sum", "gcc__requests.sum",
]
for ki, vals in enumerate(rows[2:]):
    if len(sys.argv) > 2 and int(sys.argv[2]) != ki:
        continue
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    print("== kernel", ki, d.get("Kernel Name", ""))
    for k in EXACT:
        if k in d:
            print("  %-72s %s %s" % (k, d[k], u[k]))
    stalls = [(float(v.replace(",", "")), k) for k, v in d.items()
              if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")
              and v not in ("", "n/a")]
    for v, k in sorted(stalls, reverse=True)[:9]:
        print("  stall %-50s %.3f" % (k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v))
