#!/usr/bin/env python
"""Turns the gpurun_out/<tag>_* captures of tools/make_profiles.sh into the tracked files under profiles/:
<tag>_launches.csv (ncu launch list), <tag>_launch_summary.txt (per-kernel totals and shares),
<tag>_ncu_full_summary.txt (key metrics of the --set full captures) and traffic.json (DRAM bytes per bench step
of each hot kernel = per-launch dram__bytes_read+write x the launches of one step)."""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
samples = int(sys.argv[2]) if len(sys.argv) > 2 else 32
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

# launch list
src = os.path.join(G, f"{tag}_launches.csv")
shutil.copy(src, os.path.join(P, f"{tag}_launches.csv"))
lines = [l for l in open(src) if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for row in r:
    v = float(row[vi].replace(",", ""))
    u = row[ui]
    v = v / 1e6 if u == "ns" else v / 1e3 if u == "us" else v * 1e3 if u == "s" else v
    name = re.sub(r"\(.*", "", row[ki]).replace("<unnamed>::", "").replace("void ", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in agg.values())
with open(os.path.join(P, f"{tag}_launch_summary.txt"), "w") as f:
    f.write(f"ncu launch list of `python bench.py --steps 1 --warmup 0 --no-cpu-baseline` (first 500 launches: the resident step + the e2e warm-up step)\n")
    f.write("gpu__time_duration.sum per launch is cold-cache and serialised: compare SHARES, not absolutes\n\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{v[1]:10.3f} ms {100*v[1]/tot:5.1f}%  n={v[0]:4d}  avg={1e3*v[1]/v[0]:9.1f} us  {k}\n")
    f.write(f"{tot:10.3f} ms total\n")

# full captures
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
traffic = {}
with open(os.path.join(P, f"{tag}_ncu_full_summary.txt"), "w") as f:
    for part, per_step in (("stage1", samples), ("merge", 1)):
        rep = os.path.join(G, f"{tag}_{part}.ncu-rep")
        if not os.path.exists(rep):
            continue
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        h, units = rows[0], rows[1]
        kn = h.index("Kernel Name")
        for row in rows[2:]:
            name = re.sub(r"\(.*", "", row[kn]).replace("<unnamed>::", "").replace("void ", "")
            f.write(f"== {name}   (ncu --set full --clock-control none, one launch)\n")
            b = 0.0
            for w in WANT:
                if w in h:
                    i = h.index(w)
                    f.write(f"   {w} = {row[i]} {units[i]}\n")
                    if w.startswith("dram__bytes"):
                        b += float(row[i].replace(",", "")) * UNIT.get(units[i], 1)
            key = {"k_extract<0, 0, 1>": "k_extract<scatter>", "k_extract<0, 0, 0>": "k_extract<scatter>", "k_extract16<0>": "k_extract<scatter>",
                   "k_partition2<1>": "k_partition2<count>", "k_partition2<0>": "k_partition2<scatter>", "k_bin_count": "k_bin_count",
                   "k_merge32<1, 1>": "k_merge", "k_merge_sm<1, 1, 0>": "k_merge", "k_merge": "k_merge"}.get(name)
            if key:
                traffic[key] = b * per_step
    f.write("\ntraffic.json = dram bytes of the launch above x launches per bench step (32 samples; merge: 1)\n")
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
print(open(os.path.join(P, f"{tag}_launch_summary.txt")).read())
print(json.dumps(traffic, indent=1))
