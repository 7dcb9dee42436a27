#!/usr/bin/env python
"""Builds the tracked round-2 evidence under profiles/ from the gpurun_out/ captures of tools/r02_call*.sh:

  r02_launches.csv / r02_launch_summary.txt   ncu launch list of one default bench step (shares, not absolutes)
  r02_ncu_full_summary.txt                    key metrics of the --set full captures (default kernels, record path, SimkaMin, merge)
  traffic.json                                DRAM bytes per input k-mer of every default-path kernel (bench.py scales them)
  r02_sass_evidence.txt                       registers / spills per kernel and the Blackwell-specific SASS (UBLKCP, SYNCS, ATOMS.CAS.128)
  r02_*.json                                  the bench lines (configs 2-5, 2 / 8 GPUs, reference arm, record-path A/B, k = 41)

Usage (build container, after the calls): python tools/r02_profiles.py"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
LIB = os.path.join(ROOT, "simka_b200", "lib", "libsimka_b200.so")
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def short(name):
    return re.sub(r"\(.*", "", name).replace("<unnamed>::", "").replace("void ", "").strip()


def first(*names):
    for n in names:
        p = os.path.join(G, n)
        if os.path.exists(p):
            return p
    return None


def launch_summary():
    src = first("r02h_launches.csv")
    if not src:
        return
    shutil.copy(src, os.path.join(P, "r02_launches.csv"))
    lines = [l for l in open(src) if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for row in r:
        v = float(row[vi].replace(",", ""))
        u = row[ui]
        v = v / 1e6 if u == "ns" else v / 1e3 if u == "us" else v * 1e3 if u == "s" else v
        a = agg.setdefault(short(row[ki]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(P, "r02_launch_summary.txt"), "w") as f:
        f.write("ncu launch list of `python bench.py --steps 1 --warmup 1 --no-cpu-baseline` (first 600 launches of the shipped default path)\n")
        f.write("gpu__time_duration.sum per launch is cold-cache and serialised: compare SHARES, not absolutes\n\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{v[1]:10.3f} ms {100*v[1]/tot:5.1f}%  n={v[0]:4d}  avg={1e3*v[1]/v[0]:9.1f} us  {k}\n")
        f.write(f"{tot:10.3f} ms total\n")


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def full_summary():
    traffic = {}
    parts = [("default stage 1 (bench --samples 2: one launch = one 1M-read sample = 120 M k-mers)", first("r02h_stage1.ncu-rep")),
             ("merge, N = 32 (whole step)", first("r02i_merge.ncu-rep", "r02_merge.ncu-rep")),
             ("SimkaMin distance kernel (config 5 shape, 48 sketches)", first("r02h_simkamin.ncu-rep")),
             ("tile merge, 64 < N <= 512 (bench --samples 128 --reads 200000: first batch)", first("r02k_tiles.ncu-rep")),
             ("tile merge, 512 samples x 150 k reads, k = 21 (sparse: few pairs per record and tile)", first("r02s_tiles512.ncu-rep")),
             ("record path (SIMKA_B200_SUPERK=2, not the default): scan / scatter / regrouping", first("r02b_stage1.ncu-rep")),
             ("record path: counting kernel after the convergent-insert rewrite", first("r02c_count.ncu-rep"))]
    with open(os.path.join(P, "r02_ncu_full_summary.txt"), "w") as f:
        f.write("ncu --set full --clock-control none captures of round 2 (one entry per kernel: the first captured launch)\n\n")
        for title, rep in parts:
            if not rep:
                continue
            h, units, rows = raw_rows(rep)
            ki = h.index("Kernel Name")
            f.write(f"#### {title}   [{os.path.basename(rep)}]\n")
            seen = set()
            for row in rows:
                name = short(row[ki])
                if name in seen:
                    continue
                seen.add(name)
                f.write(f"== {name}\n")
                vals = {}
                for w in WANT:
                    if w in h:
                        i = h.index(w)
                        f.write(f"   {w} = {row[i]} {units[i]}\n")
                        vals[w] = (row[i], units[i])
                try:
                    rd = float(vals["dram__bytes_read.sum"][0].replace(",", "")) * UNIT[vals["dram__bytes_read.sum"][1]]
                    wr = float(vals["dram__bytes_write.sum"][0].replace(",", "")) * UNIT[vals["dram__bytes_write.sum"][1]]
                    if "default stage 1" in title:
                        key = {"k_extract16<0>": "k_extract<scatter>", "k_partition2<1, 4>": "k_partition2<count>", "k_partition2<0, 4>": "k_partition2<scatter>",
                               "k_bin_count": "k_bin_count"}.get(name)
                        if key:
                            traffic[key] = (rd + wr) / 120e6
                    if title.startswith("merge"):
                        traffic["k_merge"] = (rd + wr) / 3.84e9
                except Exception:
                    pass
            f.write("\n")
    if traffic:
        json.dump({"bytes_per_kmer": traffic,
                   "source": "dram__bytes_read.sum + dram__bytes_write.sum of the round-2 ncu --set full captures of the shipped default path "
                             "(profiles/r02_ncu_full_summary.txt), divided by the input k-mers of the captured launch; bench.py multiplies by the k-mers of its step"},
                  open(os.path.join(P, "traffic.json"), "w"), indent=1)


def sass_evidence():
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    with open(os.path.join(P, "r02_sass_evidence.txt"), "w") as f:
        f.write("cuobjdump of simka_b200/lib/libsimka_b200.so (sm_100a), shipped build\n\n== registers / stack (spill) / static shared memory per kernel\n")
        fn = None
        for ln in res.splitlines():
            m = re.search(r"Function (\S+):", ln)
            if m:
                fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
                fn = short(fn.replace("(anonymous namespace)::", ""))
            elif "REG:" in ln and fn:
                f.write(f"{fn:45s} {ln.strip()}\n")
                fn = None
        f.write("\n== Blackwell / Hopper-and-later SASS in the kernels (count per kernel)\n")
        f.write("   UBLKCP = cp.async.bulk (1-D bulk copy, the TMA engine's non-tensor form), SYNCS.* = mbarrier operations, ATOMS.CAS.128 = 16-byte shared-memory CAS\n\n")
        cur, counts, samples = None, collections.OrderedDict(), {}
        for ln in sass.splitlines():
            m = re.search(r"Function : (\S+)", ln)
            if m:
                cur = short(subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().replace("(anonymous namespace)::", ""))
                continue
            for mn in ("UBLKCP", "SYNCS.ARRIVE.TRANS64", "SYNCS.PHASECHK.TRANS64.TRYWAIT", "SYNCS.EXCH.64", "ATOMS.CAS.128", "ATOMS.CAS.64", "ATOMS.ADD", "ATOMS.POPC.INC"):
                if re.search(r"\b" + re.escape(mn) + r"\b", ln) and cur:
                    counts.setdefault(cur, collections.Counter())[mn] += 1
                    if mn in ("UBLKCP", "ATOMS.CAS.128", "SYNCS.ARRIVE.TRANS64") and (cur, mn) not in samples:
                        samples[(cur, mn)] = ln.strip()
        for k, c in counts.items():
            if any(m in c for m in ("UBLKCP", "ATOMS.CAS.128")) or k.startswith(("k_bin_count", "k_merge_sm", "k_extract16", "k_partition2")):
                f.write(f"{k:45s} " + "  ".join(f"{m}={n}" for m, n in c.items()) + "\n")
        f.write("\n== one instruction of each kind, as disassembled\n")
        for (k, mn), ln in samples.items():
            f.write(f"{k:45s} {ln}\n")


def bench_lines():
    pairs = [("r02i_c2.json", "r02_bench_config2_1gpu.json"), ("r02h_c2.json", "r02_bench_config2_1gpu_verified.json"), ("r02g_bench2.json", "r02_bench_config2_2gpu.json"),
             ("r02j_c2_8gpu.json", "r02_bench_config2_8gpu.json"), ("r02f_c2_native.json", "r02_bench_config2_8gpu_before_overlap.json"),
             ("r02f_c2_torch.json", "r02_bench_config2_8gpu_torch_exchange.json"), ("r02f_config3.json", "r02_config3_before_overlap.json"),
             ("r02_config4.json", "r02_config4_generic_merge.json"), ("r02k_config4.json", "r02_config4.json"), ("r02k_n128_tiles.json", "r02_bench_n128_tile_merge.json"),
             ("r02k_n128_generic.json", "r02_bench_n128_generic_merge.json"), ("r02m_config3.json", "r02_config3.json"), ("r02j_config3.json", "r02_config3_generic_merge.json"),
             ("r02q_c2.json", "r02_bench_config2_1gpu_final.json"), ("r02q_config5.json", "r02_config5.json"), ("r02q_refarm.json", "r02_bench_reference_arm.json"), ("r02t_c2_2gpu.json", "r02_bench_config2_2gpu_final.json"), ("r02g_config5_ept6.json", "r02_config5.json"), ("r02h_k41.json", "r02_bench_k41_1gpu.json"),
             ("r02j_c2_4gpu.json", "r02_bench_config2_4gpu.json")]
    for src, dst in pairs:
        p = os.path.join(G, src)
        if os.path.exists(p):
            shutil.copy(p, os.path.join(P, dst))
    for log, dst in (("r02c_rec_default.log", "r02_bench_record_path.json"), ("r02_refarm.log", "r02_bench_reference_arm_first_sizing.json")):
        p = os.path.join(G, log)
        if os.path.exists(p):
            last = [l for l in open(p) if l.startswith("{")]
            if last:
                open(os.path.join(P, dst), "w").write(last[-1])


def test_logs():
    """last lines of the GPU test runs on the shipped HEAD"""
    with open(os.path.join(P, "r02_pytest_gpu_final.txt"), "w") as f:
        for log, what in (("r02q_pytest.log", "python -m pytest tests -m gpu -q -x   (one B200, tools/r02_call17.sh)"),
                          ("r02t_pytest2.log", "python -m pytest tests/test_zz_nccl_gpu.py -m gpu -q -x   (two B200s, tools/r02_call20.sh)"),
                          ("r02m_nccl_check8.log", "torchrun --nproc-per-node 8 tools/nccl_check.py   (eight B200s, tools/r02_call13.sh)")):
            p = os.path.join(G, log)
            if os.path.exists(p):
                lines = [l.rstrip() for l in open(p) if l.strip()]
                f.write(f"$ {what}\n" + "\n".join(lines[-5:]) + "\n\n")


if __name__ == "__main__":
    os.makedirs(P, exist_ok=True)
    launch_summary()
    full_summary()
    sass_evidence()
    bench_lines()
    test_logs()
    print("\n".join(sorted(f for f in os.listdir(P) if f.startswith("r02") or f == "traffic.json")))
