#!/usr/bin/env python
"""Turns the raw artefacts of tools/gpu_round.sh (gpurun_out/) into the committed summaries under profiles/."""
import collections
import csv
import json
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = os.environ.get("ROUND", "r02")   # file-name prefix of the round being summarised
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
shutil.copy(os.path.join(G, "launches.csv"), os.path.join(P, R + "_ncu_launches_bench_c3.csv"))
shutil.copy(os.path.join(G, "bench.json"), os.path.join(P, R + "_bench_c3_1gpu.json"))
shutil.copy(os.path.join(G, "bench_reference.json"), os.path.join(P, R + "_bench_reference_c3.json"))
with open(os.path.join(P, R + "_ncu_kernelB_summary.txt"), "w") as f:
    subprocess.run(["python", os.path.join(ROOT, "tools", "ncu_phase_report.py"), os.path.join(G, "prof_kernelB.ncu-rep")], stdout=f)
for rep, out in (("prof_screen.ncu-rep", R + "_ncu_screen_summary.txt"), ("prof_kernelB_screened.ncu-rep", R + "_ncu_kernelB_screened_summary.txt")):
    with open(os.path.join(P, out), "w") as f:
        subprocess.run(["python", os.path.join(ROOT, "tools", "ncu_phase_report.py"), os.path.join(G, rep)], stdout=f)
for rep, out in (("prof_kernelB.ncu-rep", R + "_ncu_kernelB_details.txt"), ("prof_others.ncu-rep", R + "_ncu_other_kernels_details.txt"),
                 ("prof_screen.ncu-rep", R + "_ncu_screen_details.txt")):
    with open(os.path.join(P, out), "w") as f:
        subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "details"], stdout=f, stderr=subprocess.DEVNULL)

lines = [l for l in open(os.path.join(P, R + "_ncu_launches_bench_c3.csv")) if not l.startswith("==")]
agg = collections.defaultdict(list)
for row in csv.DictReader(lines):
    if row.get("Metric Name") == "gpu__time_duration.sum":
        v, u = float(row["Metric Value"].replace(",", "")), row["Metric Unit"]
        agg[row["Kernel Name"].split("(")[0][-40:]].append(v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v))
tot = sum(sum(v) for v in agg.values())
out = ["ncu --metrics gpu__time_duration.sum --clock-control none -c 400, command: python bench.py --steps 2 --warmup 3 --no-extras --repeats 1",
       "(per-launch times are cold-cache and serialised: compare SHARES with bench.py's live roofline.share_of_step)", ""]
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    out.append(f"{k:42s} launches={len(v):4d} mean={sum(v) / len(v):9.1f} us  share={sum(v) / tot:.4f}")
open(os.path.join(P, R + "_ncu_launch_shares.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))

raw = subprocess.run(["ncu", "-i", os.path.join(G, "prof_others.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
idx = [hdr.index(w) for w in want]
txt = ["ncu --set full --clock-control none, one launch each = 8 frames of C3 (1920x1080, D=128, K=2); python tools/profile_run.py 3", ""]
for d in rows[2:]:
    txt.append(d[idx[0]].split("(")[0])
    for i in idx[1:]:
        txt.append(f"    {hdr[i]:70s} {d[i]} {units[i]}")
open(os.path.join(P, R + "_ncu_other_kernels_summary.txt"), "w").write("\n".join(txt) + "\n")

raw = subprocess.run(["ncu", "-i", os.path.join(G, "prof_kernelB.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, d = rows[0], rows[1], rows[2]


def val(name):
    i = hdr.index(name)
    return float(d[i]) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}[units[i]]


t = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
ts = 0.0
screen_ncu = {}
for rep in ("prof_screen.ncu-rep", "prof_kernelB_screened.ncu-rep"):
    raw = subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, d = rows[0], rows[1], rows[2]
    ts += val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    if rep == "prof_screen.ncu-rep":
        for key, name in (("shared_wavefronts_pct_of_peak", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
                          ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                          ("fma_pipe_active_pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active")):
            screen_ncu[key] = float(d[hdr.index(name)])
json.dump({"C3_per_frame": t / 8, "C3_screened_per_frame": ts / 8, "screen_kernel_ncu": screen_ncu,
           "note": "(dram__bytes_read.sum + dram__bytes_write.sum) / 8 of one launch over 8 frames of C3 (ncu --set full): C3_per_frame = "
                   "mbm_wta_fast_kernel evaluating all levels (profiles/<round>_ncu_kernelB_summary.txt); C3_screened_per_frame = mbm_screen_kernel + "
                   "mbm_wta_fast_kernel behind the screen (<round>_ncu_screen_summary.txt, <round>_ncu_kernelB_screened_summary.txt); bench.py scales it by "
                   "the frames per launch"},
          open(os.path.join(P, "kernelB_traffic.json"), "w"))
print("kernel B DRAM bytes per frame:", t / 8, "screened:", ts / 8)
