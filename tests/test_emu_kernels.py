"""Kernel LOGIC on the CPU: the CUDA sources of simka_b200/csrc compiled with g++ against tests/emu/cuda_runtime.h (CUDA threads
stepped as cooperative fibers) and driven through the same C ABI and the same ctypes front end as the GPU tests, against
the oracle and the reference's golden files, at sizes a CPU finishes in seconds.

This is test infrastructure like the oracle -- not a CPU path of the product: the emulated library is built under
tests/emu/_build, is loaded by this module only, and proves nothing about speed, memory ordering or races.  What it does
catch without a GPU: index and bit arithmetic, bin geometry, the table protocols, the overflow / retry paths and their
edge cases -- in every kernel the GPU tests launch.  The full `-m gpu` suite runs against it with `SIMKA_B200_EMU=1`
(developer switch, tests/conftest.py); the tests here are its fast subset."""
import os

import numpy as np
import pytest

import simka_b200
from emu.emu_lib import emulated
from helpers import assert_stats_equal, nasty_samples, oracle_run, reads_to_arrays, synthetic_samples
from oracle import oracle


@pytest.fixture(autouse=True)
def _emulated_library():
    with emulated():
        yield


def small_samples(n, reads, **kw):
    kw.setdefault("n_genomes", 6)
    kw.setdefault("genome_len", 1500)
    return synthetic_samples(n, reads, **kw)


def test_emulated_library_is_not_the_product_library():
    from simka_b200 import _lib
    assert "tests/emu/_build" in _lib.load()._name
    assert b"sm_100a" in _lib.load().simka_b200_version()  # same sources, same version string


@pytest.mark.parametrize("k,t", [(21, 2), (31, 0)])
def test_example_csv_byte_identical(golden_dir, k, t):
    """the reference's end-to-end golden test (tests/simple_test.py:93-119) through the emulated kernels"""
    from test_gpu_parity import load_example
    ids, samples = load_example(golden_dir)
    st = simka_b200.run_samples(samples, k, abundance_min=t, simple_dist=True, complex_dist=True)
    tdir = os.path.join(golden_dir, "truth", f"results_k{k}_t{t}")
    checked = 0
    for m, name in enumerate(st.matrix_names()):
        path = os.path.join(tdir, name + ".csv")
        if os.path.exists(path):
            assert st.csv(m, ids) == open(path).read(), name
            checked += 1
    assert checked == 20


@pytest.mark.parametrize("k,amin", [(31, 2), (21, 1), (17, 2), (16, 1), (15, 2), (9, 0)])
def test_per_sample_counts_and_stats(k, amin):
    """k = 17 / 16 straddle the switch between the two extraction kernels, k <= 15 is the MiniKC rule"""
    samples = small_samples(4, 250)
    run = oracle_run(samples, k, amin)
    st, ctx = simka_b200.run_samples(samples, k, abundance_min=amin, keep_ctx=True)
    for s in range(len(samples)):
        km, ct = ctx.dump_sample_counts(s)
        okm, oct_ = run.sample(s)
        np.testing.assert_array_equal(km, okm)
        np.testing.assert_array_equal(ct, oct_)
    assert ctx.timings()["kernel_launches"] > 0
    ctx.close()
    assert_stats_equal(st, run)


@pytest.mark.parametrize("ascii_path", [False, True])
def test_nasty_inputs(ascii_path):
    samples = nasty_samples()
    for k, amin, amax in [(31, 0, 999999999), (21, 1, 30), (15, 1, 999999999)]:
        run = oracle_run(samples, k, amin, amax)
        st = simka_b200.run_samples(samples, k, abundance_min=amin, abundance_max=amax, ascii_path=ascii_path)
        assert_stats_equal(st, run)
        assert st.matrix("mat_abundance_braycurtis")[1, 4] == 0.0


@pytest.mark.parametrize("simple,complex_", [(False, False), (True, False), (False, True)])
def test_distance_family_flags(simple, complex_):
    samples = small_samples(3, 200)
    run = oracle_run(samples, 25, 2, simple=simple, complex_=complex_)
    st = simka_b200.run_samples(samples, 25, abundance_min=2, simple_dist=simple, complex_dist=complex_)
    assert_stats_equal(st, run)


def test_forced_exact_count_path(monkeypatch):
    monkeypatch.setenv("SIMKA_B200_EXACT_COUNT", "1")
    samples = small_samples(3, 300)
    run = oracle_run(samples, 31, 2)
    st = simka_b200.run_samples(samples, 31, abundance_min=2)
    assert_stats_equal(st, run)


def test_256_thread_extraction_kernel(monkeypatch):
    monkeypatch.setenv("SIMKA_B200_K1", "256")
    samples = small_samples(3, 300)
    run = oracle_run(samples, 31, 2)
    st = simka_b200.run_samples(samples, 31, abundance_min=2)
    assert_stats_equal(st, run)


def test_heavy_hitter_falls_back_to_exact_count():
    """one k-mer far beyond a level-1 bucket's capacity: the single-pass extraction overflows and the sample is redone"""
    rng = np.random.default_rng(11)
    alpha = np.frombuffer(b"ACGT", np.uint8)
    g = alpha[rng.integers(0, 4, 3000)].tobytes()
    normal = [g[i:i + 100] for i in rng.integers(0, 2900, 150)]
    samples = [reads_to_arrays([b"A" * 150] * 1500 + normal), reads_to_arrays(normal + [b"ACGT" * 30] * 10)]
    run = oracle_run(samples, 31, 2)
    st, ctx = simka_b200.run_samples(samples, 31, abundance_min=2, keep_ctx=True)
    t = ctx.timings()
    ctx.close()
    assert t["exact_retries"] >= 1
    assert_stats_equal(st, run)


def test_bin_with_more_distinct_keys_than_the_table(monkeypatch):
    """expected_kmers_per_sample far below the truth makes few, large bins: the counting kernel's hash-selected sub-passes and
    the merge's regrouping of the lists both run"""
    samples = small_samples(3, 1500, n_genomes=24, genome_len=4000)
    run = oracle_run(samples, 31, 1)
    ctx = simka_b200.SimkaGpu(len(samples), 31, 1, simple_dist=True, complex_dist=True, expected_kmers_per_sample=3000)
    for s, (seqs, off) in enumerate(samples):
        p, b, nw = simka_b200.pack_reads(seqs, off)
        ctx.push_reads(s, p, b, nw, len(off) - 1)
        ctx.finish_sample(s)
    ctx.merge_and_accumulate()
    st = ctx.stats()
    ctx.close()
    assert_stats_equal(st, run)


@pytest.mark.parametrize("n,variant", [(70, None), (70, "tiled"), (40, "generic"), (40, None)])
def test_many_samples_merge_paths(monkeypatch, n, variant):
    """N > 64 takes the global-atomics merge (or the tiled one), N <= 64 the shared-memory-accumulator merge"""
    if variant:
        monkeypatch.setenv("SIMKA_B200_MERGE", variant)
    samples = synthetic_samples(n, 40, read_len=80, n_genomes=3, genome_len=600)
    run = oracle_run(samples, 21, 2)
    st = simka_b200.run_samples(samples, 21, abundance_min=2)
    assert_stats_equal(st, run)


def test_shards_sum_to_whole():
    samples = small_samples(4, 300)
    run = oracle_run(samples, 31, 2)
    S = 3
    ctxs = [simka_b200.SimkaGpu(len(samples), 31, 2, simple_dist=True, complex_dist=True, shard_rank=r, shard_count=S) for r in range(S)]
    for c in ctxs:
        for s, (seqs, off) in enumerate(samples):
            p, b, nw = simka_b200.pack_reads(seqs, off)
            c.push_reads(s, p, b, nw, len(off) - 1)
            c.finish_sample(s)
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


def test_import_counts_equals_counting():
    samples = small_samples(3, 300)
    run = oracle_run(samples, 31, 2)
    ctx = simka_b200.SimkaGpu(len(samples), 31, 2, simple_dist=True, complex_dist=True)
    for s in range(len(samples)):
        km, ct = run.sample(s)
        perm = np.random.default_rng(s).permutation(len(km))
        ctx.import_counts(s, km[perm], ct[perm], nb_reads=300)
    ctx.merge_and_accumulate()
    st = ctx.stats()
    ctx.close()
    assert_stats_equal(st, run)


def test_dsk_known_answers(golden_dir):
    """gatb-core's DSK known answers (test/unit/src/kmer/TestDSK.cpp:147-240) through the emulated counting kernels"""
    from test_oracle_golden import dsk_known_answers
    for seqs, k, amin, want in dsk_known_answers(golden_dir):
        sample = reads_to_arrays([s.encode() for s in seqs])
        st, ctx = simka_b200.run_samples([sample, sample], k, abundance_min=amin, keep_ctx=True)
        for s in (0, 1):
            km, ct = ctx.dump_sample_counts(s)
            assert len(km) == want, (k, amin, s)
            assert (ct >= amin).all() and (km[1:] > km[:-1]).all()
        ctx.close()
        assert int(st.distinct_shared[1]) == want


def test_simkamin_distance():
    from simka_b200 import synth
    hashes, counts = synth.make_sketches(6, 4000)
    hashes[3, :700] = 0  # a short sketch: zero padding at the front
    counts[3, :700] = 0
    jac, bc, _ = simka_b200.simkamin_distance(hashes, counts)
    ojac, obc = oracle.simkamin_distance(hashes, counts)
    np.testing.assert_array_equal(jac, ojac)
    np.testing.assert_array_equal(bc, obc)
    # the device buffers are kept between calls: a smaller call, the first one again, then released and once more
    h2, c2 = synth.make_sketches(3, 900, seed=5)
    j2, b2, _ = simka_b200.simkamin_distance(h2, c2)
    o2 = oracle.simkamin_distance(h2, c2)
    np.testing.assert_array_equal(j2, o2[0])
    np.testing.assert_array_equal(b2, o2[1])
    for release in (False, True):
        if release:
            simka_b200.simkamin_release()
        jac, bc, _ = simka_b200.simkamin_distance(hashes, counts)
        np.testing.assert_array_equal(jac, ojac)
        np.testing.assert_array_equal(bc, obc)


@pytest.mark.parametrize("k,max_reads,nb_kmers,filt", [(21, 0, 100, False), (31, 100, 1000, True)])
def test_simkamin_sketch_reproduces_reference_truth(golden_dir, k, max_reads, nb_kmers, filt):
    from test_oracle_golden import simkamin_csv, simkamin_example_sketches
    ids, gh, gc = simkamin_example_sketches(golden_dir, k, max_reads, nb_kmers,
                                            lambda s, o, kk, n: simka_b200.simkamin_sketch(s, o, kk, n, seed=100, filter=filt))
    _, oh, oc = simkamin_example_sketches(golden_dir, k, max_reads, nb_kmers,
                                          lambda s, o, kk, n: oracle.simkamin_sketch(s, o, kk, n, seed=100, filter=filt))
    np.testing.assert_array_equal(gh, oh)
    np.testing.assert_array_equal(gc, oc)
    jac, bc, _ = simka_b200.simkamin_distance(gh, gc)
    tdir = os.path.join(golden_dir, "truth_simkaMin", f"k{k}_{'filter' if filt else ''}_{max_reads}-{nb_kmers}_n1")
    assert simkamin_csv(ids, jac) == open(os.path.join(tdir, "mat_presenceAbsence_jaccard.csv")).read()
    assert simkamin_csv(ids, bc) == open(os.path.join(tdir, "mat_abundance_braycurtis.csv")).read()


@pytest.mark.parametrize("filt", [False, True])
def test_simkamin_sketch_prefilter_is_exact(monkeypatch, filt):
    """bottom-s pre-filter (only hashes below a bound are extracted; the bound widens 8-fold until the sketch fills): same sketch
    as the oracle and as the whole-range path, with fewer k-mers counted; high multiplicity forces the widening"""
    samples = (small_samples(1, 800, n_genomes=2, genome_len=600, err=0.0) + small_samples(1, 300, n_genomes=40, genome_len=1500) +
               nasty_samples()[:1])  # the first one: error-free reads at 100x coverage, a k-mer occurs ~90 times
    widened = 0
    for seqs, off in samples:
        for k, s in [(21, 50), (31, 200), (21, 1000)]:
            oh, oc, on = oracle.simkamin_sketch(seqs, off, k, s, seed=11, filter=filt)
            monkeypatch.setenv("SIMKA_B200_SKETCH_PREFILTER", "0")
            fh, fc, fn = simka_b200.simkamin_sketch(seqs, off, k, s, seed=11, filter=filt)
            full = simka_b200.simkamin_last_sketch_info()
            monkeypatch.delenv("SIMKA_B200_SKETCH_PREFILTER")
            gh, gc, gn = simka_b200.simkamin_sketch(seqs, off, k, s, seed=11, filter=filt)
            info = simka_b200.simkamin_last_sketch_info()
            assert gn == on == fn
            np.testing.assert_array_equal(gh, oh)
            np.testing.assert_array_equal(gc, oc)
            np.testing.assert_array_equal(fh, oh)
            np.testing.assert_array_equal(fc, oc)
            assert full["attempts"] == 1
            n_pos = 32 * ((len(seqs) + len(off) - 1 + 31) // 32)
            if n_pos > 64 * s + 64 and info["attempts"] == 1 and gn == s:
                assert info["keys_counted"] < full["keys_counted"] and info["distinct_dumped"] <= full["distinct_dumped"]
            widened += info["attempts"] > 1
    assert widened > 0


def test_simkamin_sketch_context_reuse(monkeypatch):
    """the sketch entries keep their stage-1 context between calls (same device / k / seed / filter): a larger sample after a
    smaller one, a change of parameters and a released cache give the sketches a fresh context gives"""
    from simka_b200 import _lib
    small, big = small_samples(1, 60)[0], small_samples(2, 700)[1]
    calls = [(small, 21, 50, False), (big, 21, 300, False), (small, 21, 50, False), (big, 31, 300, False), (big, 31, 300, True), (small, 31, 40, True)]
    monkeypatch.setenv("SIMKA_B200_SKETCH_CTX_CACHE", "0")
    fresh = [simka_b200.simkamin_sketch(sm[0], sm[1], k, s, seed=5, filter=f) for sm, k, s, f in calls]
    monkeypatch.delenv("SIMKA_B200_SKETCH_CTX_CACHE")
    for rnd in range(2):
        for (sm, k, s, f), (fh, fc, fn) in zip(calls, fresh):
            h, c, n = simka_b200.simkamin_sketch(sm[0], sm[1], k, s, seed=5, filter=f)
            assert n == fn
            np.testing.assert_array_equal(h, fh)
            np.testing.assert_array_equal(c, fc)
        _lib.load().simkamin_gpu_release_cache()
        _lib.load().simkamin_gpu_release_cache()


# ---------------------------------------------------------------------------------------------------------------
# experimental super-k-mer front end of stage 1 (SIMKA_B200_SUPERK=1, simka_b200/csrc/kernels/superk.cuh)
# ---------------------------------------------------------------------------------------------------------------
def run_superk(samples, k, amin, amax=999999999, expect_fallbacks=None, **kw):
    st, ctx = simka_b200.run_samples(samples, k, abundance_min=amin, abundance_max=amax, keep_ctx=True, **kw)
    t = ctx.timings()
    assert t["superk_samples"] + t["superk_fallbacks"] == sum(1 for s, o in samples if len(s) > 0)
    if expect_fallbacks is not None:
        assert t["superk_fallbacks"] == expect_fallbacks
    return st, ctx


@pytest.mark.parametrize("k,amin", [(31, 2), (30, 1), (26, 2), (21, 1), (18, 2), (17, 0)])
def test_superk_per_sample_counts_and_stats(monkeypatch, k, amin):
    monkeypatch.setenv("SIMKA_B200_SUPERK", "1")
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


def test_superk_is_off_by_default_and_below_k17(monkeypatch):
    samples = small_samples(2, 100)
    st, ctx = simka_b200.run_samples(samples, 31, abundance_min=1, keep_ctx=True)
    assert ctx.timings()["superk_samples"] == 0
    ctx.close()
    monkeypatch.setenv("SIMKA_B200_SUPERK", "1")
    st, ctx = simka_b200.run_samples(samples, 16, abundance_min=1, keep_ctx=True)
    assert ctx.timings()["superk_samples"] == 0
    ctx.close()
    assert_stats_equal(st, oracle_run(samples, 16, 1))


@pytest.mark.parametrize("ascii_path", [False, True])
def test_superk_nasty_inputs(monkeypatch, ascii_path):
    """N runs, reads shorter than k, lowercase, empty reads and an empty sample, low-complexity repeats (runs at the record
    capacity), a duplicated sample"""
    monkeypatch.setenv("SIMKA_B200_SUPERK", "1")
    samples = nasty_samples()
    for k, amin, amax in [(31, 0, 999999999), (21, 1, 30), (17, 2, 999999999)]:
        run = oracle_run(samples, k, amin, amax)
        st, ctx = run_superk(samples, k, amin, amax, ascii_path=ascii_path)
        ctx.close()
        assert_stats_equal(st, run)
        assert st.matrix("mat_abundance_braycurtis")[1, 4] == 0.0


def test_superk_example_csv_byte_identical(monkeypatch, golden_dir):
    monkeypatch.setenv("SIMKA_B200_SUPERK", "1")
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


def test_superk_expand_one_thread_per_record(monkeypatch):
    """the alternative expansion kernel (SIMKA_B200_SUPERK_EXPAND=thread; the default lays the k-mers out flat)"""
    monkeypatch.setenv("SIMKA_B200_SUPERK", "1")
    monkeypatch.setenv("SIMKA_B200_SUPERK_EXPAND", "thread")
    samples = nasty_samples()[:5] + small_samples(2, 300)
    for k in (31, 22):
        st, ctx = run_superk(samples, k, 1, expect_fallbacks=0)
        ctx.close()
        assert_stats_equal(st, oracle_run(samples, k, 1))


def test_superk_record_list_overflow_falls_back(monkeypatch):
    monkeypatch.setenv("SIMKA_B200_SUPERK", "1")
    monkeypatch.setenv("SIMKA_B200_SUPERK_CAP", "50")
    samples = small_samples(3, 200)
    run = oracle_run(samples, 31, 2)
    st, ctx = run_superk(samples, 31, 2, expect_fallbacks=3)
    ctx.close()
    assert_stats_equal(st, run)


def test_superk_large_bins_and_many_samples(monkeypatch):
    """few, large minimizer bins (sub-passes of the counting kernel) and the N > 64 merge behind the regrouped lists"""
    monkeypatch.setenv("SIMKA_B200_SUPERK", "1")
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
    samples = synthetic_samples(70, 40, read_len=80, n_genomes=3, genome_len=600)
    st, ctx = run_superk(samples, 21, 2, expect_fallbacks=0)
    ctx.close()
    assert_stats_equal(st, oracle_run(samples, 21, 2))


# ---------------------------------------------------------------------------------------------------------------
# multi-GPU data path (sample-sharded counting, export -> all-to-all -> import, range-sharded merge) on virtual ranks
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("world,superk,n", [(2, False, 5), (4, False, 5), (2, True, 5), (2, False, 70)])
def test_sample_sharded_exchange_equals_whole(monkeypatch, world, superk, n):
    """simka_b200.dist.exchange_and_merge with the real export / import entry points and merge kernels (emulated), one thread
    per virtual rank: the summed accumulators equal the oracle's"""
    import threading
    import torch
    from simka_b200 import dist as sdist
    from test_gpu_parity import _FakeComm
    if superk:
        monkeypatch.setenv("SIMKA_B200_SUPERK", "1")
    # (n = 70: the merge context of a rank holds imported lists of > 64 samples and merges them by sample tiles)
    samples = small_samples(5, 300) if n == 5 else small_samples(n, 14, read_len=80, n_genomes=2, genome_len=200)
    run = oracle_run(samples, 31, 2)
    N = len(samples)
    comm = _FakeComm(world)
    dev = torch.device("cpu")  # the emulator's "device memory" is host memory
    results, errors = [None] * world, []

    def worker(rank):
        try:
            cc = simka_b200.SimkaGpu(N, 31, 2, simple_dist=True, complex_dist=True, expected_kmers_per_sample=300 * 120)
            mc = simka_b200.SimkaGpu(N, 31, 2, simple_dist=True, complex_dist=True)
            for s, (seqs, off) in enumerate(samples):
                if sdist.sample_owner(s, world) != rank:
                    continue
                p, b, nw = simka_b200.pack_reads(seqs, off)
                cc.push_reads(s, p, b, nw, len(off) - 1)
                cc.finish_sample(s)
            results[rank] = sdist.exchange_and_merge(cc, mc, N, rank, world, dev, all_to_all=comm.all_to_all(rank),
                                                     all_reduce_sum=comm.all_reduce_sum(rank))
            cc.close()
            mc.close()
        except Exception as e:  # noqa: BLE001
            errors.append(e)
            comm.bar.abort()

    th = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errors, errors
    total = results[0]
    for st in results[1:]:
        total += st
    assert_stats_equal(total, run)


# ---------------------------------------------------------------------------------------------------------------
# the command-line binaries (simka, simkaCount / simkaMerge, simkaMinCore, the SimkaMin pipelines), linked against the
# library through a RUNPATH: LD_LIBRARY_PATH puts the emulated library first for the processes these tests start
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture
def emulated_cli(monkeypatch):
    from emu.emu_lib import cli_env
    monkeypatch.setenv("LD_LIBRARY_PATH", cli_env()["LD_LIBRARY_PATH"])


@pytest.mark.parametrize("pageable", [False, True])
def test_cli_simka_reproduces_truth(emulated_cli, monkeypatch, tmp_path, golden_dir, pageable):
    """pageable: the chunks the parser threads fill before the GPU context exists are plain memory (simka_main.cpp); forced for all"""
    import test_gpu_parity as G
    if pageable:
        monkeypatch.setenv("SIMKA_B200_PAGEABLE_INPUT", "1")
    for i, (k, t) in enumerate([(21, 0), (31, 2)]):
        d = tmp_path / str(i)
        d.mkdir()
        G.test_simka_cli_reproduces_truth(d, golden_dir, k, t)
    d = tmp_path / "filters"
    d.mkdir()
    G.test_simka_cli_max_reads_and_filters(d, golden_dir)


@pytest.mark.parametrize("inp,k,t", [("example_n", 15, 0), ("example_n", 12, 2), ("example_n", 31, 2), ("example_fmt", 21, 0), ("example_fmt", 15, 2)])
def test_reference_run_truth(emulated_cli, tmp_path, golden_dir, inp, k, t):
    """truth produced by the reference binaries (k <= 15 MiniKC quirks, N runs, short reads): library and command line"""
    import test_gpu_parity as G
    G.test_reference_run_truth_byte_identical(golden_dir, inp, k, t)
    G.test_simka_cli_reproduces_reference_run_truth(tmp_path, golden_dir, inp, k, t)


@pytest.mark.parametrize("tag,k,t", [("maxreads0", 31, 2), ("minsize90_shannon1.5", 21, 1), ("maxreads40_minsize85", 31, 2), ("amax3", 21, 3)])  # (all 13 in the GPU suite)
def test_cli_options_reproduce_reference_runs(emulated_cli, tmp_path, golden_dir, tag, k, t):
    """read limits and filters of the command line against reference runs with the same arguments"""
    import test_gpu_parity as G
    G.test_simka_cli_options_reproduce_reference_runs(tmp_path, golden_dir, k, t, tag)


@pytest.mark.parametrize("count,merge,k,t", [("ours", "ours", 31, 2), ("ours", "ref", 21, 0), ("ref", "ours", 21, 0)])
def test_cross_run_with_reference_binaries(emulated_cli, tmp_path, golden_dir, count, merge, k, t):
    import test_gpu_parity as G
    if not os.access(os.path.join(G.REF_BIN, "simka"), os.X_OK):
        pytest.skip("oracle/_ref/bin not built")
    G.test_cross_run_with_reference_binaries(tmp_path, golden_dir, count, merge, k, t)


@pytest.mark.parametrize("count,merge,k,t,tag", [("ours", "ref", 15, 0, ""), ("ref", "ours", 12, 2, ""), ("ours", "ours", 21, 1, "maxreads0")])  # (all seven in the GPU suite)
def test_cross_run_with_reference_binaries_on_awkward_input(emulated_cli, tmp_path, golden_dir, count, merge, k, t, tag):
    import test_gpu_parity as G
    if not os.access(os.path.join(G.REF_BIN, "simka"), os.X_OK):
        pytest.skip("oracle/_ref/bin not built")
    G.test_cross_run_with_reference_binaries(tmp_path, golden_dir, count, merge, k, t, inp="example_n", tag=tag)


@pytest.mark.parametrize("inp,k,t,mine", [("example", 21, 2, ("B", "D")), ("example_n", 15, 0, ("S2", "S4", "S5"))])
def test_mixed_per_sample_counting_with_the_reference(emulated_cli, tmp_path, golden_dir, inp, k, t, mine):
    import test_gpu_parity as G
    if not os.access(os.path.join(G.REF_BIN, "simka"), os.X_OK):
        pytest.skip("oracle/_ref/bin not built")
    G.test_mixed_per_sample_counting_with_the_reference(tmp_path, golden_dir, inp, k, t, mine)


def test_cli_stage_binaries_reference_temp_formats(emulated_cli, tmp_path, golden_dir):
    import test_gpu_parity as G
    G.test_stage_binaries_reference_temp_formats(tmp_path, golden_dir)


def test_cli_simkamincore_and_pipelines_reproduce_truth(emulated_cli, tmp_path, golden_dir):
    import test_gpu_parity as G
    import test_simkamin_pipeline as P
    for i, (k, mr, nk, filt) in enumerate([(21, 0, 100, False), (31, 0, 1000, True)]):
        d = tmp_path / f"core{i}"
        d.mkdir()
        G.test_simkamincore_cli_reproduces_reference_truth(d, golden_dir, k, mr, nk, filt)
    d = tmp_path / "py"
    d.mkdir()
    P.test_simkamin_py_reproduces_reference_truth(d, golden_dir, 31, 100, 100, True)
    d = tmp_path / "update"
    d.mkdir()
    P.test_simkamin_update_equals_one_run(d, golden_dir)


@pytest.mark.parametrize("case", [1, 5])  # (all seven in the GPU suite)
def test_cli_simkamincore_against_reference_binary(emulated_cli, tmp_path, golden_dir, case):
    import test_gpu_parity as G
    G.test_simkamincore_against_reference_binary(tmp_path, golden_dir, case)


def test_single_sample_run():
    """N = 1: no pair exists, the per-sample totals and the (1 x 1) matrices still come out like the oracle's"""
    samples = small_samples(1, 200)
    for k in (31, 15):
        run = oracle_run(samples, k, 2)
        st = simka_b200.run_samples(samples, k, abundance_min=2)
        assert_stats_equal(st, run)


def test_merge_huge_abundances_shared_memory_words_against_global_atomics(monkeypatch):
    """k_merge_sm keeps 32-bit accumulator words in shared memory: increments below 2^20 are added there, a word that reaches 2^31
    is swept to the global accumulator, anything larger goes to it directly (merge.cuh, MS_NACC).  Solid lists with abundances
    around every threshold (imported, no reads needed), against the generic merge kernel that only uses 64-bit global atomics."""
    rng = np.random.default_rng(17)
    n, nk = 6, 9000
    pool = rng.integers(0, 4**31, size=nk, dtype=np.uint64)
    special = np.array([1, 2, (1 << 20) - 1, 1 << 20, (1 << 20) + 1, 1 << 26, (1 << 31) - 1, 1 << 31, (1 << 32) - 1], np.uint32)
    lists = []
    for s in range(n):
        keep = rng.random(nk) < 0.7
        km = np.unique(pool[keep])
        ct = special[rng.integers(0, len(special), size=len(km))]
        ct[rng.random(len(km)) < 0.5] = (1 << 20) - 1  # many adds just below the fast-path bound: words cross 2^31 and are swept
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


def test_merge_tiles_huge_abundances_and_rare_sweeps(monkeypatch):
    """k_merge_tiles sweeps its shared-memory words only when 2047 entries have passed since the last sweep (merge_tiles.cuh): 70 samples
    of which three hold thousands of k-mers with abundances around every threshold (so that words cross 2^31 between sweeps and
    single adds bypass the words), the others a sprinkling of them -- against the generic kernel (64-bit global atomics only)."""
    rng = np.random.default_rng(23)
    n, nk = 70, 12000
    pool = rng.integers(0, 4**31, size=nk, dtype=np.uint64)
    special = np.array([1, 2, (1 << 20) - 1, 1 << 20, (1 << 20) + 1, 1 << 26, (1 << 31) - 1, 1 << 31, (1 << 32) - 1], np.uint32)
    lists = []
    for s in range(n):
        heavy = s in (0, 1, 40)
        km = np.unique(pool[rng.random(nk) < (0.9 if heavy else 0.02)])
        if heavy:
            ct = special[rng.integers(0, len(special), size=len(km))]
            ct[rng.random(len(km)) < 0.75] = (1 << 20) - 1  # > 4096 such adds per word and pair: a word would wrap without the sweeps
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


@pytest.mark.parametrize("n,batch", [(100, None), (96, "1500")])
def test_merge_by_sample_tiles(monkeypatch, n, batch):
    """64 < N <= 512: records of the shared k-mers (sample bitmap + abundances in sample order) written once per merge bin, pairs
    accumulated per tile of 32 x 32 samples in shared memory (merge_tiles.cuh) -- several sample blocks incl. a partial one,
    batches forced small, against the oracle; and the same accumulators as the generic kernel"""
    if batch:
        monkeypatch.setenv("SIMKA_B200_TILE_BATCH", batch)
    samples = synthetic_samples(n, 12, read_len=70, n_genomes=4, genome_len=400)
    st = simka_b200.run_samples(samples, 21, abundance_min=1)
    assert_stats_equal(st, oracle_run(samples, 21, 1))


@pytest.mark.parametrize("nopart", [False, True])
def test_merge_by_sample_tiles_large_bins_partitioned(monkeypatch, nopart):
    """merge bins of several thousand entries (lists not regrouped: SIMKA_B200_REFINE=0): k_group_bins partitions such a bin by sub-pass
    into its scratch once (merge_tiles.cuh) -- against the oracle, with and without the partition"""
    monkeypatch.setenv("SIMKA_B200_REFINE", "0")
    if nopart:
        monkeypatch.setenv("SIMKA_B200_TILE_NOPART", "1")
    samples = synthetic_samples(80, 30, read_len=80, n_genomes=5, genome_len=700)
    st = simka_b200.run_samples(samples, 21, abundance_min=1, simple_dist=True, complex_dist=True)
    assert_stats_equal(st, oracle_run(samples, 21, 1))


def test_merge_by_sample_tiles_everybody_shares(monkeypatch):
    """150 samples holding the same k-mers: groups of 150 members (sub-passes split until a group's bitmap fits), every tile busy"""
    monkeypatch.setenv("SIMKA_B200_REFINE", "0")
    rng = np.random.default_rng(11)
    read = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 120)]
    one = (read.copy(), np.array([0, 120], np.uint64))
    two = (np.concatenate([read, read[::-1].copy()]), np.array([0, 120, 240], np.uint64))
    samples = [one if i % 3 else two for i in range(150)]
    st = simka_b200.run_samples(samples, 31, abundance_min=1, simple_dist=True, complex_dist=True)
    assert_stats_equal(st, oracle_run(samples, 31, 1))
