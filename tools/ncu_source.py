"""Aggregate the ncu SASS source page by instruction class / address range.
    python tools/ncu_source.py file.ncu-rep [nbuckets]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 24
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("instructions:", len(data), "samples:", tot)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
# by opcode
byop = collections.Counter()
byop_n = collections.Counter()
for r in data:
    op = r[ix["Source"]].split()[0]
    if op.startswith("@"):
        op = r[ix["Source"]].split()[1]
    op = op.split(".")[0]
    byop[op] += int(r[ix["# Samples"]] or 0)
    byop_n[op] += int(r[ix["Instructions Executed"]] or 0)
print("-- samples by opcode (share, executed warp-instr)")
for op, s in byop.most_common(14):
    print("  %-8s %5.1f%%  %d" % (op, 100.0 * s / tot, byop_n[op]))
# by address bucket
print("-- by position in the kernel: bucket, share of samples, executed, top stalls")
n = len(data)
for b in range(nb):
    seg = data[b * n // nb:(b + 1) * n // nb]
    s = sum(int(r[ix["# Samples"]] or 0) for r in seg)
    ex = sum(int(r[ix["Instructions Executed"]] or 0) for r in seg)
    st = collections.Counter()
    for r in seg:
        for c in stall_cols:
            v = r[ix[c]]
            if v:
                st[c[6:]] += int(v)
    top = ", ".join("%s %.0f%%" % (k, 100.0 * v / max(1, sum(st.values()))) for k, v in st.most_common(4))
    ops = collections.Counter(r[ix["Source"]].split()[0].split(".")[0] for r in seg)
    print("  %2d %5.1f%% exec %10d | %s | %s" % (b, 100.0 * s / tot, ex, top,
                                                 " ".join("%s:%d" % kv for kv in ops.most_common(4))))
