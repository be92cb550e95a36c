"""per-kernel summary of an ncu --set full report: duration, issue utilisation, instructions per 32-row item, stall mix,
opcode histogram.   python scripts/ncu_phase.py report.ncu-rep rows_per_launch [kernel-regex]"""
import csv
import subprocess
import sys
from collections import Counter

rep, rows = sys.argv[1], float(sys.argv[2])
rx = sys.argv[3] if len(sys.argv) > 3 else "."
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
hdr, units = r[0], r[1]
want = ["gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread"]
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for row in r[2:]:
    name = row[hdr.index("Kernel Name")]
    print("==", name[:70])
    for w in want:
        if w in hdr:
            print(f"   {w:62s} {row[hdr.index(w)]:>14s} {units[hdr.index(w)]}")
    if "smsp__inst_executed.sum" in hdr:
        print(f"   instructions per 32-row item: {float(row[hdr.index('smsp__inst_executed.sum')]) / (rows / 32):.1f}")
    st = sorted(((float(row[hdr.index(s)]), s[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]) for s in stalls), reverse=True)
    print("   stalls:", ", ".join(f"{n} {v:.2f}" for v, n in st[:7]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
cur = None
kern = []
for row in csv.reader(src.splitlines()):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}
        kern.append(cur)
    elif row and row[0] == "Address":
        cur["hdr"] = row
    elif cur is not None and row:
        cur["rows"].append(row)
for k in kern:
    h = k["hdr"]
    ia, isamp, isrc = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
    c, cs = Counter(), Counter()
    for row in k["rows"]:
        t = row[isrc].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        c[op] += int(row[ia])
        cs[op] += int(row[isamp])
    tot, ts = sum(c.values()), max(1, sum(cs.values()))
    print("== opcodes", k["name"][:60], f"{tot / (rows / 32):.1f} per item")
    print("   " + ", ".join(f"{op} {n / (rows / 32):.1f} ({100 * cs[op] / ts:.0f}%)" for op, n in c.most_common(16)))
