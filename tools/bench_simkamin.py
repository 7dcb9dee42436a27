#!/usr/bin/env python
"""BASELINE config 5: SimkaMin sketch x sketch Jaccard / Bray-Curtis, 128 sketches x 1M k-mers, one GPU.
Prints one JSON line (kernel time from CUDA events inside simkamin_gpu_distance, CPU oracle time beside it)."""
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
ap.add_argument("--n", type=int, default=128)
ap.add_argument("--s", type=int, default=1_000_000)
ap.add_argument("--cpu-n", type=int, default=32, help="sketches of the bounded CPU sample")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
hashes, counts = synth.make_sketches(a.n, a.s)
best, wall = None, None
for _ in range(a.reps + 1):
    t0 = time.perf_counter()
    jac, bc, ms = simka_b200.simkamin_distance(hashes, counts)
    w = time.perf_counter() - t0
    best = ms if best is None else min(best, ms)
    wall = w if wall is None else min(wall, w)
pairs = a.n * (a.n - 1) // 2
consumed = pairs * 2 * a.s  # upper bound on the elements the two-pointer merges consume
line = {"workload": f"SimkaMin distance, {a.n} sketches x {a.s} k-mers", "pairs": pairs, "kernel_ms": best, "e2e_ms": 1e3 * wall,
        "pairs_per_s": pairs / (best * 1e-3), "streamed_GBs_upper": consumed * 12 / (best * 1e-3) / 1e9}
from oracle import oracle  # noqa: E402
m = min(a.cpu_n, a.n)
t0 = time.perf_counter()
oj, ob = oracle.simkamin_distance(hashes[:m], counts[:m])
dt = time.perf_counter() - t0
np.testing.assert_array_equal(jac[:m, :m], oj)
np.testing.assert_array_equal(bc[:m, :m], ob)
line["cpu_port"] = {"pairs_per_s": (m * (m - 1) // 2) / dt, "cores": os.cpu_count(), "sample": f"{m} sketches, {dt:.1f} s", "parity": "bit-exact on the sample"}
print(json.dumps(line))
