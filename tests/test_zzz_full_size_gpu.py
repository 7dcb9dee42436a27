"""Parity at BASELINE sizes: the CUDA path against the CPU oracle on the FULL configuration 2 and on the shape of
configuration 4 (512 samples), every accumulator compared.  Runs last (file name): a minute of host time each."""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import simka_b200
from helpers import assert_stats_equal
from oracle import oracle
from simka_b200 import synth

pytestmark = pytest.mark.gpu


def _run_both(n_samples, n_reads, k, n_genomes, read_len=150, amin=2):
    genomes = synth.make_genomes(n_genomes, 50_000)
    with ThreadPoolExecutor(min(8, os.cpu_count() or 1)) as ex:
        samples = list(ex.map(lambda s: synth.make_sample(genomes, s, n_reads, read_len), range(n_samples)))
    ctx = simka_b200.SimkaGpu(n_samples, k, amin, simple_dist=True, complex_dist=True, expected_kmers_per_sample=n_reads * (read_len - k + 1))
    for s, (seqs, off) in enumerate(samples):
        packed, bad, nw = simka_b200.pack_reads(seqs, off, kmer_size=k)
        ctx.push_reads(s, packed, bad, nw, n_reads)
        ctx.finish_sample(s)
    ctx.merge_and_accumulate()
    st = ctx.stats()
    t = ctx.timings()
    ctx.close()
    seqs = np.concatenate([s[0] for s in samples])
    del samples
    ro = np.arange(n_samples * n_reads + 1, dtype=np.uint64) * np.uint64(read_len)
    so = np.arange(n_samples + 1, dtype=np.uint64) * np.uint64(n_reads)
    run = oracle.Run(seqs, ro, so, k, abundance_min=amin, simple=True, complex_=True, nthreads=os.cpu_count() or 1)
    run.merge()
    return st, run, t


def test_full_config2_all_accumulators_equal_the_oracle():
    """BASELINE.json configs[1] at its stated size: 32 samples x 1M reads x 150 bp, k = 31, abundance-min 2, all distances.
    Integers bit-exact, Jensen-Shannon sums within 1e-10 relative, the 21 matrices equal as floats."""
    st, run, t = _run_both(32, 1_000_000, 31, 256)
    assert int(t["kmers_in"]) == 32 * 1_000_000 * 120
    assert_stats_equal(st, run, kl_rtol=1e-10)
    run.close()


def test_config4_shape_512_samples_equal_the_oracle():
    """The shape of BASELINE.json configs[3] (512 samples, k = 21; 20 000 reads per sample so that the oracle finishes in seconds):
    the N > 64 merge path (global 64-bit atomics, Jensen-Shannon terms added in floating point) against the oracle at 1e-10."""
    st, run, _ = _run_both(512, 20_000, 21, 256 * 16)
    assert_stats_equal(st, run, kl_rtol=1e-10)
    run.close()


def test_merge_sub_pass_split_when_everybody_shares_the_same_kmers(monkeypatch):
    """600 samples holding the same k-mers, lists not regrouped (one big merge bin): the hash-selected sub-passes of k_merge
    hold whole groups of 600 entries, some of them more than fit -> those sub-passes are split in two before anything is
    accumulated (merge.cuh).  Against the oracle."""
    from helpers import oracle_run
    monkeypatch.setenv("SIMKA_B200_REFINE", "0")
    rng = np.random.default_rng(11)
    read = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 150)]
    one = (read.copy(), np.array([0, 150], np.uint64))
    samples = [one] * 600
    st = simka_b200.run_samples(samples, 31, abundance_min=1, simple_dist=True, complex_dist=True)
    assert_stats_equal(st, oracle_run(samples, 31, 1))


@pytest.mark.parametrize("n,reads,batch", [(130, 600, None), (260, 300, "5000"), (512, 200, None)])
def test_merge_by_sample_tiles_gpu(monkeypatch, n, reads, batch):
    """64 < N <= 512: the tile merge (merge_tiles.cuh) against the oracle, several batches forced for one case, and against the
    generic global-atomics kernel"""
    from helpers import oracle_run, synthetic_samples
    if batch:
        monkeypatch.setenv("SIMKA_B200_TILE_BATCH", batch)
    samples = synthetic_samples(n, reads, read_len=100, n_genomes=12, genome_len=3000)
    st = simka_b200.run_samples(samples, 21, abundance_min=2)
    assert_stats_equal(st, oracle_run(samples, 21, 2))
    monkeypatch.setenv("SIMKA_B200_MERGE", "generic")
    st2 = simka_b200.run_samples(samples, 21, abundance_min=2)
    for name in st.U64_SYM + st.U64_SQ:
        np.testing.assert_array_equal(getattr(st, name), getattr(st2, name), err_msg=name)


def test_merge_tiles_huge_abundances_gpu(monkeypatch):
    """the tile merge's shared-memory words and their rare sweeps (merge_tiles.cuh, SIMKA_TM_SWEEP_ENTRIES) with abundances around every
    threshold: 200 samples, eight of them heavy with hundreds of thousands of k-mers, against the generic kernel (64-bit global atomics)"""
    rng = np.random.default_rng(29)
    n, nk = 200, 400_000
    pool = rng.integers(0, 4**31, size=nk, dtype=np.uint64)
    special = np.array([1, 2, (1 << 20) - 1, 1 << 20, (1 << 20) + 1, 1 << 26, (1 << 31) - 1, 1 << 31, (1 << 32) - 1], np.uint32)
    lists = []
    for s in range(n):
        heavy = s % 25 == 0
        km = np.unique(pool[rng.random(nk) < (0.9 if heavy else 0.01)])
        if heavy:
            ct = special[rng.integers(0, len(special), size=len(km))]
            ct[rng.random(len(km)) < 0.75] = (1 << 20) - 1
        else:
            ct = rng.integers(1, 5, size=len(km)).astype(np.uint32)
        lists.append((km, ct))
    results = {}
    for variant in ("", "generic"):
        if variant:
            monkeypatch.setenv("SIMKA_B200_MERGE", variant)
        ctx = simka_b200.SimkaGpu(n, 31, 1, simple_dist=True, complex_dist=True)
        for s, (km, ct) in enumerate(lists):
            ctx.import_counts(s, km, ct, nb_reads=10)
        ctx.merge_and_accumulate()
        results[variant] = ctx.stats()
        ctx.close()
    a, b = results[""], results["generic"]
    for name in a.U64_SYM + a.U64_SQ:
        np.testing.assert_array_equal(getattr(a, name), getattr(b, name), err_msg=name)
    np.testing.assert_allclose(a.kullback_leibler, b.kullback_leibler, rtol=1e-10, atol=0)
