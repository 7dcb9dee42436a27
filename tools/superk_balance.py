#!/usr/bin/env python
"""Bin balance of the super-k-mer front end (simka_b200/csrc/kernels/superk.cuh), simulated in numpy with the kernel's own
hash functions on one synthetic sample of the bench workload: keys per minimizer bin, distinct keys per bin against the
capacity of the counting table, and records per k-mer.  CPU only; usage: tools/superk_balance.py [reads] [k] [m] [bin_bits]"""
import sys

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from simka_b200 import synth  # noqa: E402

reads = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 31
M = int(sys.argv[3]) if len(sys.argv) > 3 else 13
L = 150
npos = reads * (L + 1)
BIN_TARGET = 768
bin_bits = int(sys.argv[4]) if len(sys.argv) > 4 else min(20, int(np.floor(np.log2(max(1, npos // BIN_TARGET)))))
W = K - M + 1
NMAX = min(32, 51 - K + 1)

genomes = synth.make_genomes(256, 50_000)
seqs, off = synth.make_sample(genomes, 0, reads, L)
codes = ((seqs.reshape(reads, L) >> 1) & 3).astype(np.uint32)


def sk_hash(x):
    x = (x * np.uint32(0x9E3779B1)).astype(np.uint32)
    x ^= x >> np.uint32(15)
    x = (x * np.uint32(0x85EBCA77)).astype(np.uint32)
    return x


hist = np.zeros(1 << bin_bits, np.int64)
nrec = 0
nk = 0
sample_bins = {}
CH = 20000
for c0 in range(0, reads, CH):
    c = codes[c0:c0 + CH]
    R = len(c)
    fwd = np.zeros((R, L - M + 1), np.uint32)
    rev = np.zeros((R, L - M + 1), np.uint32)
    for j in range(M):
        fwd = (fwd << np.uint32(2)) | c[:, j:j + L - M + 1]
        rev = rev | ((c[:, j:j + L - M + 1] ^ np.uint32(2)) << np.uint32(2 * j))
    hv = sk_hash(np.minimum(fwd, rev))
    mv = hv[:, :L - K + 1].copy()
    for j in range(1, W):
        np.minimum(mv, hv[:, j:j + L - K + 1], out=mv)
    b = ((mv * np.uint32(0xC2B2AE3D)).astype(np.uint32) >> np.uint32(32 - bin_bits)).astype(np.int64)
    hist += np.bincount(b.reshape(-1), minlength=1 << bin_bits)
    # records: a new run where the bin changes, at the start of a read, at a 32-position word boundary of the packed
    # stream (position of k-mer j of read r = r * (L + 1) + j) and at the record capacity (ignored: runs are shorter)
    pos = (np.arange(c0, c0 + R, dtype=np.int64)[:, None] * (L + 1) + np.arange(L - K + 1)[None, :])
    start = np.ones(b.shape, bool)
    start[:, 1:] = (b[:, 1:] != b[:, :-1]) | ((pos[:, 1:] & 31) == 0)
    nrec += int(start.sum())
    nk += b.size
print(f"reads {reads}  k {K}  m {M}  window {W}  bins 2^{bin_bits}  k-mers {nk}  records {nrec}  k-mers/record {nk / nrec:.2f}  "
      f"bytes/k-mer through the record list {16 * nrec / nk:.2f}")
mean = hist.mean()
print(f"keys per bin: mean {mean:.0f}  sd {hist.std():.0f} (binomial would be {np.sqrt(mean):.0f})  max {hist.max()}  "
      f"p50 {np.percentile(hist, 50):.0f}  p99 {np.percentile(hist, 99):.0f}  p99.9 {np.percentile(hist, 99.9):.0f}")
for thr in (1536, 2048, 3072, 4096, 8192):
    print(f"  bins with more than {thr} keys: {(hist > thr).mean() * 100:.3f} %  holding {(hist[hist > thr].sum() / hist.sum()) * 100:.2f} % of the keys")

# balanced bins = runs of consecutive micro-bins whose start falls into the same multiple of TARGET keys (k_superk_bins)
if len(sys.argv) > 5:
    T = int(sys.argv[5])
    prefix = np.concatenate([[0], np.cumsum(hist)])
    binid = prefix[:-1] // T
    nb = int(prefix[-1] // T) + 1
    sizes = np.bincount(binid[hist > 0], weights=hist[hist > 0], minlength=nb)
    nz = sizes[sizes > 0]
    print(f"balanced: target {T}  bins {nb} ({(sizes == 0).sum()} empty)  non-empty: mean {nz.mean():.0f}  sd {nz.std():.0f}  p99 {np.percentile(nz, 99):.0f}  max {nz.max():.0f}")
    for thr in (1536, 2048, 4096):
        print(f"  bins with more than {thr} keys: {(nz > thr).mean() * 100:.3f} %  holding {(nz[nz > thr].sum() / nz.sum()) * 100:.2f} % of the keys")
