#!/usr/bin/env python
"""SimkaMin sketch stage on one GPU: wall time of simkamin_gpu_sketch on one bench-like sample with and
without the bottom-s pre-filter (SIMKA_B200_SKETCH_PREFILTER), the two results compared bit for bit, the oracle's time on a
bounded sample beside them.  Prints one JSON line.  Wall clock around the C-ABI call: host buffers in, sketch out (includes the
context creation, the H2D copy of the packed reads and the D2H copy + host sort of the counted hashes)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import simka_b200  # noqa: E402
from simka_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=1_000_000)
ap.add_argument("--s", type=int, default=1_000_000)
ap.add_argument("--kmer", type=int, default=21)
ap.add_argument("--cpu-reads", type=int, default=100_000)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--filter", action="store_true")
a = ap.parse_args()
genomes = synth.make_genomes(256, 50_000)
seqs, off = synth.make_sample(genomes, 0, a.reads, 150, 0.01)
res = {}
for tag, env in (("prefilter", None), ("whole_range", "0")):
    if env is None:
        os.environ.pop("SIMKA_B200_SKETCH_PREFILTER", None)
    else:
        os.environ["SIMKA_B200_SKETCH_PREFILTER"] = env
    best = None
    for _ in range(a.reps + 1):
        t0 = time.perf_counter()
        h, c, n = simka_b200.simkamin_sketch(seqs, off, a.kmer, a.s, seed=100, filter=a.filter)
        w = time.perf_counter() - t0
        best = w if best is None else min(best, w)
    res[tag] = {"wall_ms": 1e3 * best, **simka_b200.simkamin_last_sketch_info(), "n_selected": n, "out": (h, c)}
same = bool(np.array_equal(res["prefilter"]["out"][0], res["whole_range"]["out"][0]) and np.array_equal(res["prefilter"]["out"][1], res["whole_range"]["out"][1]))
kmers = a.reads * (150 - a.kmer + 1)
line = {"workload": f"SimkaMin sketch, 1 sample x {a.reads} reads x 150 bp, k={a.kmer}, s={a.s}, filter={a.filter}", "identical": same}
for tag in res:
    r = res[tag]
    line[tag] = {k: v for k, v in r.items() if k != "out"}
    line[tag]["kmers_per_s"] = kmers / (r["wall_ms"] * 1e-3)
from oracle import oracle  # noqa: E402
m = min(a.cpu_reads, a.reads)
cs, co = seqs[:off[m]], off[:m + 1]
t0 = time.perf_counter()
oh, oc, on = oracle.simkamin_sketch(cs, co, a.kmer, a.s, seed=100, filter=a.filter)
dt = time.perf_counter() - t0
gh, gc, gn = simka_b200.simkamin_sketch(cs, co, a.kmer, a.s, seed=100, filter=a.filter)
line["cpu_port"] = {"kmers_per_s": m * (150 - a.kmer + 1) / dt, "sample": f"{m} reads, {dt:.1f} s",
                    "parity": "bit-exact on the sample" if (gn == on and np.array_equal(gh, oh) and np.array_equal(gc, oc)) else "MISMATCH"}
print(json.dumps(line))
