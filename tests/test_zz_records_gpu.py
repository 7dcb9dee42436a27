"""GPU parity of the super-k-mer RECORD path of stage 1 (SIMKA_B200_SUPERK=2, simka_b200/csrc/kernels/superk2.cuh: records
scattered by minimizer bin and counted straight from shared memory, where they arrive by cp.async.bulk copies).  Same oracle /
golden checks as the k-mer path; tests/test_emu_records.py steps the same kernels on the CPU."""
import os

import numpy as np
import pytest

import simka_b200
from helpers import assert_stats_equal, nasty_samples, oracle_run, synthetic_samples

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _superk(monkeypatch):
    monkeypatch.setenv("SIMKA_B200_SUPERK", "2")


def counted_by_superk(ctx, samples):
    t = ctx.timings()
    return t["superk_samples"] == sum(1 for s, o in samples if len(s) > 0) and t["superk_fallbacks"] == 0


@pytest.mark.parametrize("k,amin", [(31, 2), (27, 1), (21, 2), (17, 0)])
def test_records_synthetic_per_sample_counts_and_stats(k, amin):
    samples = synthetic_samples(6, 3000)
    run = oracle_run(samples, k, amin)
    st, ctx = simka_b200.run_samples(samples, k, abundance_min=amin, keep_ctx=True)
    assert counted_by_superk(ctx, samples)
    for s in range(len(samples)):
        km, ct = ctx.dump_sample_counts(s)
        okm, oct_ = run.sample(s)
        np.testing.assert_array_equal(km, okm)
        np.testing.assert_array_equal(ct, oct_)
    ctx.close()
    assert_stats_equal(st, run)


@pytest.mark.parametrize("k,t", [(21, 0), (31, 2)])
def test_records_example_csv_byte_identical(golden_dir, k, t):
    from test_gpu_parity import load_example
    ids, samples = load_example(golden_dir)
    st, ctx = simka_b200.run_samples(samples, k, abundance_min=t, keep_ctx=True)
    assert counted_by_superk(ctx, samples)
    ctx.close()
    tdir = os.path.join(golden_dir, "truth", f"results_k{k}_t{t}")
    checked = 0
    for m, name in enumerate(st.matrix_names()):
        path = os.path.join(tdir, name + ".csv")
        if os.path.exists(path):
            assert st.csv(m, ids) == open(path).read(), name
            checked += 1
    assert checked == 20


@pytest.mark.parametrize("ascii_path", [False, True])
def test_records_nasty_inputs(ascii_path):
    samples = nasty_samples()
    for k, amin, amax in [(31, 0, 999999999), (21, 2, 999999999), (21, 1, 30), (17, 2, 999999999)]:
        run = oracle_run(samples, k, amin, amax)
        st = simka_b200.run_samples(samples, k, abundance_min=amin, abundance_max=amax, ascii_path=ascii_path)
        assert_stats_equal(st, run)
        assert st.matrix("mat_abundance_braycurtis")[1, 4] == 0.0


def test_records_medium_size_and_many_samples():
    """thousands of minimizer bins with multi-thousand-key bins among them; then 70 samples (regrouped lists, N > 64 merge)"""
    samples = synthetic_samples(5, 40000, n_genomes=32, genome_len=20000)
    st, ctx = simka_b200.run_samples(samples, 31, abundance_min=2, keep_ctx=True)
    assert counted_by_superk(ctx, samples)
    ctx.close()
    assert_stats_equal(st, oracle_run(samples, 31, 2))
    samples = synthetic_samples(70, 600, read_len=100, n_genomes=8, genome_len=3000)
    st = simka_b200.run_samples(samples, 21, abundance_min=2)
    assert_stats_equal(st, oracle_run(samples, 21, 2))


def test_records_heavy_minimizer_sub_passes():
    """thousands of distinct k-mers under ONE minimizer (low-complexity reads): more than a table pass holds -> hash-selected
    sub-passes of the counting kernel; and bins of several shared-memory chunks"""
    from helpers import reads_to_arrays
    rng = np.random.default_rng(5)
    reads = []
    for _ in range(3000):
        r = bytearray(b"A" * 150)
        for p in rng.integers(20, 150, 12):
            r[int(p)] = b"ACGT"[int(rng.integers(0, 4))]
        reads.append(bytes(r))
    samples = [reads_to_arrays(reads)] + synthetic_samples(2, 3000)
    st, ctx = simka_b200.run_samples(samples, 31, abundance_min=1, keep_ctx=True)
    assert counted_by_superk(ctx, samples)
    ctx.close()
    assert_stats_equal(st, oracle_run(samples, 31, 1))


def test_records_record_list_overflow_falls_back(monkeypatch):
    monkeypatch.setenv("SIMKA_B200_SUPERK_CAP", "1000")
    samples = synthetic_samples(3, 2000)
    st, ctx = simka_b200.run_samples(samples, 31, abundance_min=2, keep_ctx=True)
    t = ctx.timings()
    ctx.close()
    assert t["superk_fallbacks"] == 3 and t["superk_samples"] == 0
    assert_stats_equal(st, oracle_run(samples, 31, 2))
