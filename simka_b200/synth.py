"""Seeded synthetic read sets (BASELINE.md section 3 / SURVEY.md 8d generator).

default_rng(12345) -> G "genomes" of `genome_len` bases, uniform over ACGT.  Sample s uses
default_rng(1000+s): abundance profile w ~ Dirichlet(0.3); every read picks a genome ~ w, a uniform
start, exactly `read_len` bases, substitution errors at rate `err` (number of errors ~ Binomial,
positions uniform), reverse-complemented with p = 0.5.
"""
from __future__ import annotations

import numpy as np

_ACGT = np.frombuffer(b"ACGT", np.uint8)
_COMP = np.zeros(256, np.uint8)
for a, b in zip(b"ACGT", b"TGCA"):
    _COMP[a] = b


def make_genomes(n_genomes: int = 256, genome_len: int = 50_000, seed: int = 12345) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return _ACGT[rng.integers(0, 4, size=(n_genomes, genome_len), dtype=np.uint8)]


def make_sample(genomes: np.ndarray, sample: int, n_reads: int, read_len: int = 150, err: float = 0.01):
    """-> (seqs uint8[n_reads*read_len], read_off uint64[n_reads+1])"""
    rng = np.random.default_rng(1000 + sample)
    G, glen = genomes.shape
    w = rng.dirichlet(np.full(G, 0.3))
    g = rng.choice(G, size=n_reads, p=w)
    start = rng.integers(0, glen - read_len + 1, size=n_reads)
    flat = genomes.reshape(-1)
    first = g.astype(np.int64) * glen + start
    idx_t = np.int32 if flat.size < 2**31 else np.int64
    ar = np.arange(read_len, dtype=idx_t)
    reads = np.empty((n_reads, read_len), np.uint8)  # (n_reads, read_len) ASCII, gathered block-wise (same values, bounded temporaries)
    for b0 in range(0, n_reads, 262144):
        b1 = min(n_reads, b0 + 262144)
        reads[b0:b1] = flat[first[b0:b1].astype(idx_t)[:, None] + ar[None, :]]
    total = n_reads * read_len
    n_err = int(rng.binomial(total, err)) if err > 0 else 0
    if n_err:
        pos = rng.integers(0, total, size=n_err)
        r = reads.reshape(-1)
        # substitute by a different base: rotate within ACGT by 1..3
        code = np.searchsorted(_ACGT, r[pos])
        r[pos] = _ACGT[(code + rng.integers(1, 4, size=n_err)) & 3]
    flip = rng.random(n_reads) < 0.5
    reads[flip] = _COMP[reads[flip][:, ::-1]]
    off = np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(read_len)
    return np.ascontiguousarray(reads).reshape(-1), off


def make_sketches(n: int, s: int, seed: int = 7):
    """SimkaMin config-5 sketches: pool = unique(rng(seed).integers(1,2^63,4s)), sketch i = sorted
    default_rng(100+i).choice(pool, s, replace=False), counts uniform in [1,19]."""
    pool = np.unique(np.random.default_rng(seed).integers(1, 2**63, size=4 * s, dtype=np.uint64))
    hashes = np.zeros((n, s), np.uint64)
    counts = np.zeros((n, s), np.uint32)
    for i in range(n):
        rng = np.random.default_rng(100 + i)
        hashes[i] = np.sort(rng.choice(pool, size=s, replace=False))
        counts[i] = rng.integers(1, 20, size=s, dtype=np.uint32)
    return hashes, counts
