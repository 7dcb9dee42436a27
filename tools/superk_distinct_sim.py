#!/usr/bin/env python
"""Design simulation for the super-k-mer front end (numpy, CPU only): keys and DISTINCT keys per balanced counting bin on one bench sample, against the capacity of the counting table."""
import sys, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np
from simka_b200 import synth
reads = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 896
K, M, L = 31, 13, 150
W = K - M + 1
MB = 21
genomes = synth.make_genomes(256, 50_000)
seqs, off = synth.make_sample(genomes, 0, reads, L)
codes = ((seqs.reshape(reads, L) >> 1) & 3).astype(np.uint32)
def sk_hash(x):
    x = (x * np.uint32(0x9E3779B1)).astype(np.uint32); x ^= x >> np.uint32(15)
    x = (x * np.uint32(0x85EBCA77)).astype(np.uint32); return x
nkm = L - K + 1
micro_all = np.empty((reads, nkm), np.uint32)
kh_all = np.empty((reads, nkm), np.uint64)
CH = 20000
mask = np.uint64((1 << 62) - 1)
for c0 in range(0, reads, CH):
    c = codes[c0:c0 + CH]; R = len(c)
    fwd = np.zeros((R, L - M + 1), np.uint32); rev = np.zeros((R, L - M + 1), np.uint32)
    for j in range(M):
        fwd = (fwd << np.uint32(2)) | c[:, j:j + L - M + 1]
        rev = rev | ((c[:, j:j + L - M + 1] ^ np.uint32(2)) << np.uint32(2 * j))
    hv = sk_hash(np.minimum(fwd, rev))
    mv = hv[:, :nkm].copy()
    for j in range(1, W): np.minimum(mv, hv[:, j:j + nkm], out=mv)
    micro_all[c0:c0 + R] = (mv * np.uint32(0xC2B2AE3D)).astype(np.uint32) >> np.uint32(32 - MB)
    # canonical 31-mers as 62-bit integers
    kf = np.zeros((R, nkm), np.uint64); kr = np.zeros((R, nkm), np.uint64)
    c64 = c.astype(np.uint64)
    for j in range(K):
        kf = (kf << np.uint64(2)) | c64[:, j:j + nkm]
        kr = kr | ((c64[:, j:j + nkm] ^ np.uint64(2)) << np.uint64(2 * j))
    kh_all[c0:c0 + R] = np.minimum(kf, kr)
micro = micro_all.reshape(-1).astype(np.int64)
hist = np.bincount(micro, minlength=1 << MB)
prefix = np.concatenate([[0], np.cumsum(hist)])
bin_of_micro = prefix[:-1] // T
b = bin_of_micro[micro]
nb = int(b.max()) + 1
keys_per_bin = np.bincount(b, minlength=nb)
# distinct k-mers per bin
order = np.lexsort((kh_all.reshape(-1), b))
bs = b[order]; ks = kh_all.reshape(-1)[order]
newk = np.ones(len(bs), bool); newk[1:] = (bs[1:] != bs[:-1]) | (ks[1:] != ks[:-1])
distinct_per_bin = np.bincount(bs[newk], minlength=nb)
nz = keys_per_bin > 0
print(f"T={T}: bins {nz.sum()}  keys/bin mean {keys_per_bin[nz].mean():.0f} p99 {np.percentile(keys_per_bin[nz],99):.0f} max {keys_per_bin.max()}  "
      f"distinct/bin mean {distinct_per_bin[nz].mean():.0f} p99 {np.percentile(distinct_per_bin[nz],99):.0f} max {distinct_per_bin.max()}")
for cap in (1536, 3072):
    over = distinct_per_bin > cap
    print(f"  bins with more than {cap} distinct keys: {over.sum()} ({over.mean()*100:.3f} %), holding {keys_per_bin[over].sum()/keys_per_bin.sum()*100:.2f} % of the keys")
