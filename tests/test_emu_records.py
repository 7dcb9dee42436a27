"""CPU tests of the super-k-mer RECORD path of stage 1 (kernels/superk2.cuh: records scattered by minimizer bin, counted straight
from the records that arrive in shared memory by bulk copies) on the kernel-logic emulator, against the oracle."""
import os

import numpy as np
import pytest

import simka_b200
from helpers import assert_stats_equal, nasty_samples, oracle_run, synthetic_samples
from test_emu_kernels import run_superk, small_samples, _emulated_library  # noqa: F401  (autouse fixture: the emulated library)

MODE = "2"


@pytest.mark.parametrize("k,amin", [(31, 2), (30, 1), (26, 2), (21, 1), (18, 2), (17, 0)])
def test_records_per_sample_counts_and_stats(monkeypatch, k, amin):
    monkeypatch.setenv("SIMKA_B200_SUPERK", MODE)
    samples = small_samples(4, 250)
    run = oracle_run(samples, k, amin)
    st, ctx = run_superk(samples, k, amin, expect_fallbacks=0)
    for s in range(len(samples)):
        km, ct = ctx.dump_sample_counts(s)
        okm, oct_ = run.sample(s)
        np.testing.assert_array_equal(km, okm)
        np.testing.assert_array_equal(ct, oct_)
    ctx.close()
    assert_stats_equal(st, run)


@pytest.mark.parametrize("ascii_path", [False, True])
def test_records_nasty_inputs(monkeypatch, ascii_path):
    monkeypatch.setenv("SIMKA_B200_SUPERK", MODE)
    samples = nasty_samples()
    for k, amin, amax in [(31, 0, 999999999), (21, 1, 30), (17, 2, 999999999)]:
        run = oracle_run(samples, k, amin, amax)
        st, ctx = run_superk(samples, k, amin, amax, ascii_path=ascii_path)
        ctx.close()
        assert_stats_equal(st, run)
        assert st.matrix("mat_abundance_braycurtis")[1, 4] == 0.0


def test_records_example_csv_byte_identical(monkeypatch, golden_dir):
    monkeypatch.setenv("SIMKA_B200_SUPERK", MODE)
    from test_gpu_parity import load_example
    ids, samples = load_example(golden_dir)
    for k, t in ((21, 0), (31, 2)):
        st, ctx = run_superk(samples, k, t, expect_fallbacks=0)
        ctx.close()
        tdir = os.path.join(golden_dir, "truth", f"results_k{k}_t{t}")
        for m, name in enumerate(st.matrix_names()):
            path = os.path.join(tdir, name + ".csv")
            if os.path.exists(path):
                assert st.csv(m, ids) == open(path).read(), name


def test_records_record_list_overflow_falls_back(monkeypatch):
    monkeypatch.setenv("SIMKA_B200_SUPERK", MODE)
    monkeypatch.setenv("SIMKA_B200_SUPERK_CAP", "50")
    samples = small_samples(3, 200)
    run = oracle_run(samples, 31, 2)
    st, ctx = run_superk(samples, 31, 2, expect_fallbacks=3)
    ctx.close()
    assert_stats_equal(st, run)


def test_records_large_bins_multi_chunk_and_table_overflow(monkeypatch):
    """bins of many records (several shared-memory chunks per bin) and bins with more distinct k-mers than one table pass holds
    (the hash-selected sub-passes): few micro-bins for a large sample"""
    monkeypatch.setenv("SIMKA_B200_SUPERK", MODE)
    samples = small_samples(3, 1500, n_genomes=24, genome_len=4000)
    run = oracle_run(samples, 31, 1)
    ctx = simka_b200.SimkaGpu(len(samples), 31, 1, simple_dist=True, complex_dist=True, expected_kmers_per_sample=3000)
    for s, (seqs, off) in enumerate(samples):
        p, b, nw = simka_b200.pack_reads(seqs, off)
        ctx.push_reads(s, p, b, nw, len(off) - 1)
        ctx.finish_sample(s)
    ctx.merge_and_accumulate()
    st = ctx.stats()
    assert ctx.timings()["superk_samples"] == 3
    ctx.close()
    assert_stats_equal(st, run)
    # one heavy minimizer: low-complexity reads put thousands of distinct k-mers into a single micro-bin
    rng = np.random.default_rng(5)
    reads = []
    for _ in range(600):
        r = bytearray(b"A" * 150)
        for p in rng.integers(20, 150, 12):
            r[int(p)] = b"ACGT"[int(rng.integers(0, 4))]
        reads.append(bytes(r))
    from helpers import reads_to_arrays
    samples = [reads_to_arrays(reads), small_samples(1, 300)[0]]
    st, ctx = run_superk(samples, 31, 1, expect_fallbacks=0)
    ctx.close()
    assert_stats_equal(st, oracle_run(samples, 31, 1))


def test_records_many_samples(monkeypatch):
    monkeypatch.setenv("SIMKA_B200_SUPERK", MODE)
    samples = synthetic_samples(70, 40, read_len=80, n_genomes=3, genome_len=600)
    st, ctx = run_superk(samples, 21, 2, expect_fallbacks=0)
    ctx.close()
    assert_stats_equal(st, oracle_run(samples, 21, 2))


def test_records_shards_sum_to_whole(monkeypatch):
    """bucket-sharded contexts on the record path (several GPUs, or several passes of one GPU over the reads): the shards'
    statistics sum to the whole"""
    monkeypatch.setenv("SIMKA_B200_SUPERK", MODE)
    samples = small_samples(4, 300)
    run = oracle_run(samples, 31, 2)
    S = 3
    ctxs = [simka_b200.SimkaGpu(len(samples), 31, 2, simple_dist=True, complex_dist=True, shard_rank=r, shard_count=S) for r in range(S)]
    for c in ctxs:
        for s, (seqs, off) in enumerate(samples):
            p, b, nw = simka_b200.pack_reads(seqs, off)
            c.push_reads(s, p, b, nw, len(off) - 1)
            c.finish_sample(s)
        assert c.timings()["superk_samples"] == len(samples)
    tot = sum(c.get_totals() for c in ctxs)
    tot[:, 0] = ctxs[0].get_totals()[:, 0]
    total = None
    for c in ctxs:
        c.set_totals(tot)
        c.merge_and_accumulate()
        st = c.stats()
        total = st if total is None else total.__iadd__(st)
        c.close()
    assert_stats_equal(total, run)
