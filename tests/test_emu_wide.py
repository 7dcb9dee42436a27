"""CPU tests of the 128-bit k-mer path (32 <= k <= 63, kernels/wide.cuh) on the kernel-logic emulator: against the oracle
(per-sample k-mer counts and every accumulator) and against truth produced by the reference binaries (tests/golden/truth_ref)."""
import json
import os

import numpy as np
import pytest

import simka_b200
from helpers import assert_stats_equal, nasty_samples, oracle_run
from test_emu_kernels import small_samples, _emulated_library  # noqa: F401  (autouse fixture: the emulated library)


@pytest.mark.parametrize("k,amin", [(32, 1), (41, 2), (47, 1), (63, 0)])
def test_wide_per_sample_counts_and_stats(k, amin):
    samples = small_samples(4, 250)
    run = oracle_run(samples, k, amin)
    st, ctx = simka_b200.run_samples(samples, k, abundance_min=amin, keep_ctx=True)
    for s in range(len(samples)):
        km, ct = ctx.dump_sample_counts(s)
        okm, oct_ = run.sample(s)
        np.testing.assert_array_equal(km, okm)
        np.testing.assert_array_equal(ct, oct_)
    ctx.close()
    assert_stats_equal(st, run)


@pytest.mark.parametrize("ascii_path", [False, True])
def test_wide_nasty_inputs(ascii_path):
    samples = nasty_samples()
    for k, amin, amax in [(33, 0, 999999999), (50, 1, 30), (63, 2, 999999999)]:
        run = oracle_run(samples, k, amin, amax)
        st = simka_b200.run_samples(samples, k, abundance_min=amin, abundance_max=amax, ascii_path=ascii_path)
        assert_stats_equal(st, run)
        assert st.matrix("mat_abundance_braycurtis")[1, 4] == 0.0


@pytest.mark.parametrize("inp,k,t", [("example", 32, 0), ("example", 41, 2), ("example", 63, 2), ("example_n", 41, 0)])
def test_wide_reference_run_truth_byte_identical(golden_dir, inp, k, t):
    from test_gpu_parity import load_example
    ids, samples = load_example(golden_dir, inp)
    st = simka_b200.run_samples(samples, k, abundance_min=t, simple_dist=True, complex_dist=True)
    tdir = os.path.join(golden_dir, "truth_ref", ("" if inp == "example" else inp + "_") + f"results_k{k}_t{t}")
    for m, name in enumerate(st.matrix_names()):
        assert st.csv(m, ids) == open(os.path.join(tdir, name + ".csv")).read(), name
    sync = json.load(open(os.path.join(tdir, "count_synchro.json")))
    for i, ident in enumerate(ids):
        assert [int(x) for x in st.totals[i]] == sync[ident], ident


def test_wide_many_samples_and_large_bins():
    """70 samples (merge sub-passes over one big bin: the lists of 128-bit k-mers are not regrouped) and bins with more distinct
    k-mers than one table pass holds"""
    from helpers import synthetic_samples
    samples = synthetic_samples(70, 40, read_len=100, n_genomes=3, genome_len=600)
    st = simka_b200.run_samples(samples, 41, abundance_min=2)
    assert_stats_equal(st, oracle_run(samples, 41, 2))
    samples = small_samples(2, 1500, n_genomes=24, genome_len=4000)
    ctx = simka_b200.SimkaGpu(len(samples), 35, 1, simple_dist=True, complex_dist=True, expected_kmers_per_sample=3000)
    for s, (seqs, off) in enumerate(samples):
        p, b, nw = simka_b200.pack_reads(seqs, off, kmer_size=35)
        ctx.push_reads(s, p, b, nw, len(off) - 1)
        ctx.finish_sample(s)
    ctx.merge_and_accumulate()
    st = ctx.stats()
    ctx.close()
    assert_stats_equal(st, oracle_run(samples, 35, 1))


def test_wide_cli_reproduces_reference_run_truth(monkeypatch, tmp_path, golden_dir):
    from emu.emu_lib import cli_env
    import test_gpu_parity as G
    monkeypatch.setenv("LD_LIBRARY_PATH", cli_env()["LD_LIBRARY_PATH"])
    G.test_simka_cli_reproduces_reference_run_truth(tmp_path, golden_dir, "example_n", 41, 0)
