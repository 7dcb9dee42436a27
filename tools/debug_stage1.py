"""debug: stage 1 only, per-sample dump vs oracle"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import simka_b200
from helpers import oracle_run, synthetic_samples
k, amin = int(sys.argv[1]), int(sys.argv[2])
samples = synthetic_samples(6, int(sys.argv[3]) if len(sys.argv) > 3 else 3000)
run = oracle_run(samples, k, amin)
ctx = simka_b200.SimkaGpu(len(samples), k, amin, simple_dist=True, complex_dist=True)
for s, (seqs, off) in enumerate(samples):
    p, b, nw = simka_b200.pack_reads(seqs, off)
    ctx.push_reads(s, p, b, nw, len(off) - 1)
    tot = ctx.finish_sample(s)
    km, ct = ctx.dump_sample_counts(s)
    okm, oct_ = run.sample(s)
    print(s, "totals", tot, "oracle", run.all_totals()[s], "n", len(km), len(okm),
          "dups", len(km) - len(np.unique(km)), "eq", np.array_equal(km, okm) and np.array_equal(ct, oct_))
    if not np.array_equal(km, okm):
        a = set(km.tolist()); o = set(okm.tolist())
        print("   missing", len(o - a), "extra", len(a - o))
