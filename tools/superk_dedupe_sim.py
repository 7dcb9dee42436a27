#!/usr/bin/env python
"""Design simulation for the super-k-mer front end (numpy, CPU only): how many k-mer inserts remain if identical content-defined super-k-mers are counted first."""
import sys, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np
from simka_b200 import synth
reads = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
sample = int(sys.argv[2]) if len(sys.argv) > 2 else 0
K, M, L = 31, 13, 150
W = K - M + 1
genomes = synth.make_genomes(256, 50_000)
seqs, off = synth.make_sample(genomes, sample, reads, L)
codes = ((seqs.reshape(reads, L) >> 1) & 3).astype(np.uint32)
def sk_hash(x):
    x = (x * np.uint32(0x9E3779B1)).astype(np.uint32); x ^= x >> np.uint32(15)
    x = (x * np.uint32(0x85EBCA77)).astype(np.uint32); return x
P = np.uint64(0x9E3779B97F4A7C15)
powers = np.ones(L + 1, np.uint64)
for i in range(1, L + 1): powers[i] = powers[i - 1] * P
rec_keys = []
nk = 0; nrec = 0
t0 = time.time()
CH = 20000
for c0 in range(0, reads, CH):
    c = codes[c0:c0 + CH]; R = len(c)
    fwd = np.zeros((R, L - M + 1), np.uint32); rev = np.zeros((R, L - M + 1), np.uint32)
    for j in range(M):
        fwd = (fwd << np.uint32(2)) | c[:, j:j + L - M + 1]
        rev = rev | ((c[:, j:j + L - M + 1] ^ np.uint32(2)) << np.uint32(2 * j))
    hv = sk_hash(np.minimum(fwd, rev))
    nkm = L - K + 1
    mv = hv[:, :nkm].copy()
    for j in range(1, W): np.minimum(mv, hv[:, j:j + nkm], out=mv)
    start = np.ones((R, nkm), bool)
    start[:, 1:] = (mv[:, 1:] != mv[:, :-1]) | (hv[:, W:W + nkm - 1] == mv[:, 1:])
    # prefix hashes of the read and of its reverse complement (base codes + 1)
    def prefix(cc):
        H = np.zeros((R, L + 1), np.uint64)
        for i in range(L): H[:, i + 1] = H[:, i] * P + (cc[:, i].astype(np.uint64) + np.uint64(1))
        return H
    Hf = prefix(c)
    rc = (c[:, ::-1] ^ np.uint32(2))
    Hr = prefix(rc)
    r_idx, p_idx = np.nonzero(start)
    # run length: distance to next start in same read or end
    flat = r_idx * nkm + p_idx
    nxt = np.empty_like(flat); nxt[:-1] = flat[1:]; nxt[-1] = R * nkm
    end_of_read = (r_idx + 1) * nkm
    n = np.minimum(nxt, end_of_read) - flat
    s = p_idx; e = p_idx + n + K - 1  # base range [s, e)
    hf = Hf[r_idx, e] - Hf[r_idx, s] * powers[e - s]
    s2 = L - e; e2 = L - s
    hr = Hr[r_idx, e2] - Hr[r_idx, s2] * powers[e2 - s2]
    key = np.minimum(hf, hr) ^ (n.astype(np.uint64) << np.uint64(58))
    rec_keys.append(np.stack([key, n.astype(np.uint64)], 1))
    nk += R * nkm; nrec += len(n)
allk = np.concatenate(rec_keys)
u, idx = np.unique(allk[:, 0], return_index=True)
distinct_kmers_in_records = int(allk[idx, 1].sum())
print(f"reads {reads}: k-mers {nk}, records {nrec} ({nk/nrec:.2f} k-mers/record, max {int(allk[:,1].max())}), distinct records {len(u)} "
      f"({len(u)/nrec*100:.1f} %), k-mer inserts after record dedupe {distinct_kmers_in_records} = {distinct_kmers_in_records/nk*100:.1f} % of today's", round(time.time()-t0), "s")
