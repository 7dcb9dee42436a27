#!/usr/bin/env python
"""Randomised differential test of the kernels (stepped on the CPU by tests/emu) against the oracle: random k (2..31), abundance
bounds, distance flags, 2..8 samples of reads of awkward lengths (0, 1, k-1, k, k+1, 31..33, 64, ...) cut from random and
low-complexity sequences with lowercase / N / junk characters, packed or ASCII input, default or super-k-mer stage 1.
Developer tool (CPU only):  tools/emu_fuzz.py [seed] [seconds]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import simka_b200
from emu.emu_lib import use_emulator
use_emulator()
from helpers import reads_to_arrays, oracle_run, assert_stats_equal
from oracle import oracle
oracle.build()
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
alpha = np.frombuffer(b"ACGTacgtNn-", np.uint8)
pw = np.array([30, 30, 30, 30, 2, 2, 2, 2, 1, 1, 1], float); pw /= pw.sum()
def genome(n, lowc=False):
    if lowc:
        unit = alpha[rng.integers(0, 4, int(rng.integers(1, 7)))].tobytes()
        return (unit * (n // len(unit) + 1))[:n]
    return alpha[rng.choice(len(alpha), size=n, p=pw)].tobytes()
t0 = time.time(); cases = 0
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 60
while time.time() - t0 < budget:
    k = int(rng.integers(2, 32))
    amin = int(rng.integers(0, 4)); amax = int(rng.choice([999999999, 999999999, 5, 40]))
    simple, complex_ = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
    g = [genome(int(rng.integers(100, 2500)), lowc=rng.random() < 0.25) for _ in range(3)]
    samples = []
    for s in range(int(rng.integers(2, 9))):
        reads = []
        for _ in range(int(rng.integers(0, 90))):
            src = g[int(rng.integers(0, 3))]
            L = int(rng.choice([0, 1, k - 1, k, k + 1, 31, 32, 33, 64, 100, 150, 400]))
            a = int(rng.integers(0, max(1, len(src) - L)))
            r = src[a:a + L]
            if rng.random() < 0.5:
                r = r.translate(bytes.maketrans(b"ACGTacgt", b"TGCAtgca"))[::-1]
            reads.append(r)
        samples.append(reads_to_arrays(reads))
    os.environ["SIMKA_B200_SUPERK"] = str(int(rng.integers(0, 2)))
    ascii_path = bool(rng.integers(0, 2))
    run = oracle_run(samples, k, amin, amax, simple=simple, complex_=complex_)
    try:
        st, ctx = simka_b200.run_samples(samples, k, abundance_min=amin, abundance_max=amax, simple_dist=simple, complex_dist=complex_, ascii_path=ascii_path, keep_ctx=True)
        for s in range(len(samples)):
            km, ct = ctx.dump_sample_counts(s)
            okm, oct_ = run.sample(s)
            assert np.array_equal(km, okm) and np.array_equal(ct, oct_), ("counts", s)
        ctx.close()
        assert_stats_equal(st, run)
    except Exception as e:
        print("FAIL", dict(k=k, amin=amin, amax=amax, simple=simple, complex_=complex_, n=len(samples), ascii=ascii_path, superk=os.environ["SIMKA_B200_SUPERK"], case=cases), repr(e)[:300])
        raise
    cases += 1
print("cases", cases, "OK")
