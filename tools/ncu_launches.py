"""Summarises an ncu --csv launch list (one metric per row) per kernel: mean of every metric over the launches.
usage: python tools/ncu_launches.py gpurun_out/x.csv [first_launch_index]"""
import csv
import sys
from collections import OrderedDict, defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
acc = OrderedDict()
for r in rows[1:]:
    if int(r[ii]) < skip:
        continue
    d = acc.setdefault(r[ki].split("(")[0], defaultdict(list))
    d[r[mi]].append(float(r[vi].replace(",", "")))
for k, d in acc.items():
    n = len(d["gpu__time_duration.sum"])
    ms = sum(d["gpu__time_duration.sum"]) / n / 1e6
    out = {"n": n, "ms": round(ms, 3)}
    if "dram__bytes_read.sum" in d:
        gb = (sum(d["dram__bytes_read.sum"]) + sum(d["dram__bytes_write.sum"])) / n / 1e9
        out["dram_GB"] = round(gb, 2)
        out["TBps"] = round(gb / ms, 2)
    for m, short in (("smsp__inst_executed.sum", "inst"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
                     ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
                     ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%")):
        if m in d:
            out[short] = round(sum(d[m]) / n, 2) if short != "inst" else int(sum(d[m]) / n)
    print(k, out)
