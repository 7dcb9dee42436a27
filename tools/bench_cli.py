#!/usr/bin/env python
"""End-to-end timing of the `simka` command line of this build on synthetic FASTA files (first byte of input read ->
matrices written), the measurement SURVEY 8(d) defines for the reference binary.  Files go to /dev/shm."""
import argparse
import json
import os
import shutil
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simka_b200 import synth  # noqa: E402
from simka_b200.build import SIMKA_BIN  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--samples", type=int, default=32)
ap.add_argument("--reads", type=int, default=250_000)
ap.add_argument("--kmer", type=int, default=31)
ap.add_argument("--dir", default="/dev/shm/simka_cli_bench")
ap.add_argument("--nb-cores", type=int, default=0)
ap.add_argument("--with-reference", action="store_true", help="also run the reference `simka` (oracle/_ref/bin) on the same files and compare every matrix byte for byte")
a = ap.parse_args()
shutil.rmtree(a.dir, ignore_errors=True)
os.makedirs(a.dir)
genomes = synth.make_genomes(256 if a.samples <= 32 else 256 * ((a.samples + 31) // 32), 50_000)
lines = []
nbytes = 0
for s in range(a.samples):
    seqs, off = synth.make_sample(genomes, s, a.reads, 150)
    rec = np.empty((a.reads, 153), np.uint8)
    rec[:, 0] = ord(">")
    rec[:, 1] = ord("r")
    rec[:, 2] = ord("\n")
    rec[:, 3:153] = seqs.reshape(a.reads, 150)
    out = np.concatenate([rec, np.full((a.reads, 1), ord("\n"), np.uint8)], axis=1)
    path = os.path.join(a.dir, f"S{s}.fasta")
    out.tofile(path)
    nbytes += out.size
    lines.append(f"S{s}: {path}")
open(os.path.join(a.dir, "input.txt"), "w").write("\n".join(lines) + "\n")
cmd = [str(SIMKA_BIN), "-in", os.path.join(a.dir, "input.txt"), "-out", os.path.join(a.dir, "out"), "-out-tmp", os.path.join(a.dir, "tmp"),
       "-kmer-size", str(a.kmer), "-abundance-min", "2", "-simple-dist", "-complex-dist", "-nb-cores", str(a.nb_cores)]
best = None
for rep in range(2):
    t0 = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    assert r.returncode == 0, r.stdout + r.stderr
    best = dt if best is None else min(best, dt)
kmers = a.samples * a.reads * (150 - a.kmer + 1)
tail = [l for l in r.stdout.splitlines() if l.startswith("GPU time")]
line = {"workload": f"simka CLI, {a.samples} samples x {a.reads} reads x 150 bp FASTA in /dev/shm ({nbytes / 1e9:.2f} GB), k={a.kmer}, all distances",
        "wall_s": best, "kmers_per_s": kmers / best, "fasta_GBs": nbytes / 1e9 / best, "cores": os.cpu_count(), "cli": tail[-1] if tail else ""}
if a.with_reference:
    import gzip
    ref = os.path.join(ROOT, "oracle", "_ref", "bin", "simka")
    assert os.path.exists(ref), "build the reference first: oracle/build_ref.sh"
    cores = a.nb_cores or os.cpu_count()
    rcmd = [ref, "-in", os.path.join(a.dir, "input.txt"), "-out", os.path.join(a.dir, "out_ref"), "-out-tmp", os.path.join(a.dir, "tmp_ref"),
            "-kmer-size", str(a.kmer), "-abundance-min", "2", "-simple-dist", "-complex-dist", "-nb-cores", str(cores), "-max-memory", "20000"]
    t0 = time.perf_counter()
    rr = subprocess.run(rcmd, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    assert rr.returncode == 0, rr.stdout[-2000:] + rr.stderr[-2000:]
    names = sorted(f for f in os.listdir(os.path.join(a.dir, "out_ref")) if f.endswith(".csv.gz"))
    differ = [f for f in names if gzip.open(os.path.join(a.dir, "out_ref", f)).read() != gzip.open(os.path.join(a.dir, "out", f)).read()]
    line["reference"] = {"wall_s": dt, "kmers_per_s": kmers / dt, "cores": cores, "matrices": len(names), "matrices_differing": differ,
                         "speedup_wall": dt / best}
print(json.dumps(line))
shutil.rmtree(a.dir, ignore_errors=True)
