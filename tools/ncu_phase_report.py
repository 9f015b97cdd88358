#!/usr/bin/env python
"""Summarises an ncu report of the fused kernel: headline metrics + per-phase (barrier-delimited) stalls."""
import collections
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active"]
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h:85s} {units[i]:10s} {data[0][i]}")
for i, h in enumerate(hdr):
    if "smsp__average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio"):
        try:
            v = float(data[0][i])
        except ValueError:
            continue
        if v > 0.03:
            print(f"  stall {h.split('stalled_')[1].split('_per_issue')[0]:25s} {v:.3f}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
data = []
for d in rows[2:]:
    if d and d[0] == "Kernel Name":
        break
    if len(d) >= len(hdr) - 2:
        data.append(d)
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")


def I(x):
    try:
        return int(x)
    except ValueError:
        return 0


bars = [i for i, d in enumerate(data) if "BAR.SYNC" in d[isrc]]
pts = [0] + bars + [len(data)]
st = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(I(d[isamp]) for d in data)
print("static instructions", len(data), "samples", tot, "barriers at", bars)
for a, b in zip(pts[:-1], pts[1:]):
    c = collections.Counter()
    for d in data[a:b]:
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", d[isrc].strip())
        c[m.group(2).split(".")[0] if m else "?"] += I(d[iex])
    ex = sum(c.values())
    sm = sum(I(d[isamp]) for d in data[a:b])
    pipe = 2 * (c["FADD2"] + c["FMUL2"]) + c["FADD"]
    ss = {hdr[i][6:]: sum(I(d[i]) for d in data[a:b]) for i in st}
    t = max(1, sum(ss.values()))
    print(f"[{a:5d},{b:5d}) samples {sm / tot:6.3f}  executed {ex:11d}  fma-pipe cycles {pipe:11d} ({pipe / max(1, ex):.2f}/inst)")
    print("      ops   ", {k: round(v / ex, 3) for k, v in c.most_common(7)})
    print("      stalls", {k: round(v / t, 3) for k, v in ss.items() if v / t > 0.02})
