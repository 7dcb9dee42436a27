"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle."""
import json
import os
import signal

import numpy as np
import pytest

import simka_b200
from helpers import assert_stats_equal, nasty_samples, oracle_run, synthetic_samples
from oracle import oracle

pytestmark = pytest.mark.gpu


def load_example(golden_dir, inp="example"):
    spec = oracle.parse_simka_input(os.path.join(golden_dir, inp, "simka_input.txt"))
    ids, samples = [], []
    for ident, groups in spec:
        ids.append(ident)
        datas, offs, base = [], [np.zeros(1, np.uint64)], 0
        for grp in groups:
            for f in grp:
                d, o = oracle.read_fastx(f)
                datas.append(d)
                offs.append(o[1:] + np.uint64(base))
                base += len(d)
        samples.append((np.concatenate(datas), np.concatenate(offs)))
    return ids, samples


@pytest.mark.parametrize("k", [21, 31])
@pytest.mark.parametrize("t", [0, 2])
def test_example_csv_byte_identical(golden_dir, k, t):
    """The reference's own end-to-end golden test (tests/simple_test.py:93-119) through the GPU path."""
    ids, samples = load_example(golden_dir)
    st = simka_b200.run_samples(samples, k, abundance_min=t, simple_dist=True, complex_dist=True)
    tdir = os.path.join(golden_dir, "truth", f"results_k{k}_t{t}")
    checked = 0
    for m, name in enumerate(st.matrix_names()):
        path = os.path.join(tdir, name + ".csv")
        if not os.path.exists(path):
            continue  # mat_abundance_jaccard has no truth file in the reference either
        assert st.csv(m, ids) == open(path).read(), name
        checked += 1
    assert checked == 20


@pytest.mark.parametrize("inp,k", [("example", 9), ("example", 15), ("example", 16), ("example_n", 12), ("example_n", 15), ("example_n", 21), ("example_n", 31),
                                   ("example", 32), ("example", 41), ("example", 63), ("example_n", 41),  # k > 31: 128-bit k-mers (kernels/wide.cuh)
                                   ("example_fmt", 15), ("example_fmt", 31)])  # file-format corner cases
@pytest.mark.parametrize("t", [0, 2])
def test_reference_run_truth_byte_identical(golden_dir, inp, k, t):
    """Truth produced by the reference binaries themselves (tools/make_ref_golden.py) for what its own truth files do not
    cover: k <= 15 (MiniKC: N windows counted, short reads replay the previous read) and an input with N runs, lower case,
    short reads, multi-line FASTA, FASTQ.gz and a paired sample -- through the library."""
    import json
    ids, samples = load_example(golden_dir, inp)
    st = simka_b200.run_samples(samples, k, abundance_min=t, simple_dist=True, complex_dist=True)
    tdir = os.path.join(golden_dir, "truth_ref", ("" if inp == "example" else inp + "_") + f"results_k{k}_t{t}")
    for m, name in enumerate(st.matrix_names()):
        assert st.csv(m, ids) == open(os.path.join(tdir, name + ".csv")).read(), name
    sync = json.load(open(os.path.join(tdir, "count_synchro.json")))
    for i, ident in enumerate(ids):
        assert [int(x) for x in st.totals[i]] == sync[ident], ident


@pytest.mark.parametrize("inp,k,t", [("example_n", 15, 0), ("example_n", 12, 2), ("example_n", 31, 2), ("example", 16, 0), ("example", 41, 2), ("example_n", 41, 0),
                                     ("example", 63, 0), ("example_fmt", 21, 0), ("example_fmt", 31, 2), ("example_fmt", 15, 2)])
def test_simka_cli_reproduces_reference_run_truth(tmp_path, golden_dir, inp, k, t):
    """same, through the `simka` command line of this build (its own FASTA / FASTQ / gz readers); example_fmt: FASTQ whose quality
    lines start with '@' / '+' / '>', CR-LF, blank lines, header-only records, blanks inside sequence lines, no final newline,
    IUPAC codes, a two-member gz file, long headers"""
    import gzip
    import subprocess
    from simka_b200.build import SIMKA_BIN
    out = tmp_path / "results"
    cmd = [str(SIMKA_BIN), "-in", os.path.join(golden_dir, inp, "simka_input.txt"), "-out", str(out), "-out-tmp", str(tmp_path / "tmp"),
           "-simple-dist", "-complex-dist", "-kmer-size", str(k), "-abundance-min", str(t), "-verbose", "0"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    tdir = os.path.join(golden_dir, "truth_ref", ("" if inp == "example" else inp + "_") + f"results_k{k}_t{t}")
    checked = 0
    for f in sorted(os.listdir(tdir)):
        if f.endswith(".csv"):
            assert gzip.open(out / (f + ".gz"), "rt").read() == open(os.path.join(tdir, f)).read(), f
            checked += 1
    assert checked == 21


OPTION_TAGS = ["maxreads50", "maxreads0", "minsize90_shannon1.5", "maxreads40_minsize85", "amax3", "kshannon1.5"]
OPTION_CASES = [(k, t, tag) for tag in OPTION_TAGS for k, t in ((21, 1), (31, 2))] + [(21, 3, "amax3")]


@pytest.mark.parametrize("k,t,tag", OPTION_CASES)
def test_simka_cli_options_reproduce_reference_runs(tmp_path, golden_dir, k, t, tag):
    """-max-reads (explicit and automatic, with a paired and a concatenated sample), -min-read-size, -min-shannon-index,
    -abundance-max and -kmer-shannon-index (accepted, without effect) through the `simka` command line: per-sample totals and all
    21 matrices equal to what the REFERENCE `simka` produced with the same arguments
    (tests/golden/truth_ref/example_n_results_k*_t*_<tag>, tools/make_ref_golden.py)"""
    import gzip
    import json
    import subprocess
    from simka_b200.build import SIMKA_BIN
    tdir = os.path.join(golden_dir, "truth_ref", f"example_n_results_k{k}_t{t}_{tag}")
    extra = json.load(open(os.path.join(tdir, "options.json")))
    out = tmp_path / "results"
    cmd = [str(SIMKA_BIN), "-in", os.path.join(golden_dir, "example_n", "simka_input.txt"), "-out", str(out), "-out-tmp", str(tmp_path / "tmp"),
           "-simple-dist", "-complex-dist", "-kmer-size", str(k), "-abundance-min", str(t), "-verbose", "0", "-keep-tmp"] + extra
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    sync = json.load(open(os.path.join(tdir, "count_synchro.json")))
    sdir = tmp_path / "tmp" / "simka_output_temp" / "count_synchro"
    for ident, want in sync.items():
        assert [int(x) for x in open(sdir / (ident + ".ok")).read().split()] == want, ident
    checked = 0
    for f in sorted(os.listdir(tdir)):
        if f.endswith(".csv"):
            assert gzip.open(out / (f + ".gz"), "rt").read() == open(os.path.join(tdir, f)).read(), f
            checked += 1
    assert checked == 21


@pytest.mark.parametrize("k,amin", [(31, 2), (21, 1), (15, 2), (9, 0), (32, 2), (45, 1), (63, 2)])
def test_synthetic_per_sample_counts_and_stats(k, amin):
    samples = synthetic_samples(6, 3000)
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
def test_nasty_inputs(ascii_path):
    samples = nasty_samples()
    # k <= 15: the reference's MiniKC path, which also counts windows containing non-ACGT characters
    for k, amin, amax in [(31, 0, 999999999), (21, 2, 999999999), (21, 1, 30), (15, 1, 999999999), (9, 2, 999999999)]:
        run = oracle_run(samples, k, amin, amax)
        st = simka_b200.run_samples(samples, k, abundance_min=amin, abundance_max=amax, ascii_path=ascii_path)
        assert_stats_equal(st, run)
        # duplicated sample -> distance exactly 0 in Bray-Curtis
        assert st.matrix("mat_abundance_braycurtis")[1, 4] == 0.0


@pytest.mark.parametrize("simple,complex_", [(False, False), (True, False), (False, True)])
def test_distance_family_flags(simple, complex_):
    samples = synthetic_samples(4, 1500)
    run = oracle_run(samples, 25, 2, simple=simple, complex_=complex_)
    st = simka_b200.run_samples(samples, 25, abundance_min=2, simple_dist=simple, complex_dist=complex_)
    assert_stats_equal(st, run)


def test_medium_size_larger_bins():
    """Enough k-mers per sample for multi-tile buckets and thousands of bins."""
    samples = synthetic_samples(5, 40000, n_genomes=32, genome_len=20000)
    run = oracle_run(samples, 31, 2)
    st = simka_b200.run_samples(samples, 31, abundance_min=2)
    assert_stats_equal(st, run)


def test_many_samples():
    samples = synthetic_samples(40, 600, read_len=100, n_genomes=8, genome_len=3000)
    run = oracle_run(samples, 21, 2)
    st = simka_b200.run_samples(samples, 21, abundance_min=2)
    assert_stats_equal(st, run)


def test_shards_sum_to_whole():
    """Multi-GPU semantics on one device: the accumulators of S shard contexts sum to the unsharded
    run (the reference's per-partition stats sum, src/SimkaPotara.hpp:1161-1174)."""
    samples = synthetic_samples(5, 2500)
    run = oracle_run(samples, 31, 2)
    S = 3
    ctxs = [simka_b200.SimkaGpu(len(samples), 31, 2, simple_dist=True, complex_dist=True, shard_rank=r, shard_count=S)
            for r in range(S)]
    for c in ctxs:
        for s, (seqs, off) in enumerate(samples):
            p, b, nw = simka_b200.pack_reads(seqs, off)
            c.push_reads(s, p, b, nw, len(off) - 1)
            c.finish_sample(s)
    tot = sum(c.get_totals() for c in ctxs)
    tot[:, 0] = ctxs[0].get_totals()[:, 0]  # nb_reads is per sample, not summed
    total = None
    for c in ctxs:
        c.set_totals(tot)
        c.merge_and_accumulate()
        st = c.stats()
        if total is None:
            total = st
        else:
            total += st
        c.close()
    assert_stats_equal(total, run)


def test_simkamin_distance():
    from simka_b200 import synth
    hashes, counts = synth.make_sketches(12, 5000)
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


def test_heavy_hitter_falls_back_to_exact_count():
    """One k-mer repeated far beyond a bucket's capacity: the optimistic single-pass K1 overflows, the
    sample is redone with the exact count pass, and the result is still bit-exact."""
    from helpers import reads_to_arrays
    rng = np.random.default_rng(11)
    alpha = np.frombuffer(b"ACGT", np.uint8)
    g = alpha[rng.integers(0, 4, 5000)].tobytes()
    normal = [g[i:i + 100] for i in rng.integers(0, 4900, 600)]
    s0 = reads_to_arrays([b"A" * 150] * 3000 + normal)
    s1 = reads_to_arrays(normal + [b"ACGT" * 30] * 10)
    s2 = reads_to_arrays([b"T" * 120] * 50 + normal[:300])
    samples = [s0, s1, s2]
    run = oracle_run(samples, 31, 2)
    st, ctx = simka_b200.run_samples(samples, 31, abundance_min=2, keep_ctx=True)
    t = ctx.timings()
    ctx.close()
    assert t["exact_retries"] >= 1
    assert_stats_equal(st, run)


def test_forced_exact_count_path(monkeypatch):
    monkeypatch.setenv("SIMKA_B200_EXACT_COUNT", "1")
    samples = synthetic_samples(4, 2000)
    run = oracle_run(samples, 31, 2)
    st = simka_b200.run_samples(samples, 31, abundance_min=2)
    assert_stats_equal(st, run)


class _FakeComm:
    """in-process stand-in for torch.distributed over `world` threads (one virtual rank each)"""

    def __init__(self, world):
        import threading
        self.world = world
        self.bar = threading.Barrier(world)
        self.slots = [None] * world

    def all_reduce_sum(self, rank):
        def f(t):
            self.slots[rank] = t.clone()
            self.bar.wait()
            out = sum(self.slots[1:], self.slots[0].clone())
            self.bar.wait()
            return out
        return f

    def all_to_all(self, rank):
        import torch

        def f(send, send_splits, recv_splits):
            self.slots[rank] = torch.split(send, list(map(int, send_splits)))
            self.bar.wait()
            parts = [self.slots[o][rank] for o in range(self.world)]
            assert [len(p) for p in parts] == list(map(int, recv_splits))
            out = torch.cat(parts) if parts else send[:0]
            self.bar.wait()
            return out
        return f


@pytest.mark.parametrize("world", [2, 4])
def test_sample_sharded_exchange_equals_whole(world):
    """The multi-GPU path on one device: `world` virtual ranks each count their own samples, exchange the
    solid k-mer lists by hash range and merge their block; the summed accumulators equal the oracle's."""
    import threading
    import torch
    from simka_b200 import dist as sdist
    samples = synthetic_samples(6, 2500)
    run = oracle_run(samples, 31, 2)
    N = len(samples)
    comm = _FakeComm(world)
    dev = torch.device("cuda", 0)
    results, errors = [None] * world, []

    def worker(rank):
        try:
            cc = simka_b200.SimkaGpu(N, 31, 2, simple_dist=True, complex_dist=True, expected_kmers_per_sample=2500 * 120)
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


@pytest.mark.parametrize("k,t", [(31, 2), (21, 0)])
def test_simka_cli_reproduces_truth(tmp_path, golden_dir, k, t):
    """The reference's end-to-end test (tests/simple_test.py:25-66,93-119) against the `simka` binary of this build:
    same command line, gunzipped matrices byte-identical to the truth files."""
    import gzip
    import subprocess
    from simka_b200.build import SIMKA_BIN
    out = tmp_path / "results"
    cmd = [str(SIMKA_BIN), "-in", os.path.join(golden_dir, "example", "simka_input.txt"), "-out", str(out), "-out-tmp", str(tmp_path / "tmp"),
           "-simple-dist", "-complex-dist", "-kmer-size", str(k), "-abundance-min", str(t), "-verbose", "0"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    tdir = os.path.join(golden_dir, "truth", f"results_k{k}_t{t}")
    checked = 0
    for f in sorted(os.listdir(tdir)):
        got = gzip.open(out / (f + ".gz"), "rt").read()
        assert got == open(os.path.join(tdir, f)).read(), f
        checked += 1
    assert checked == 20
    assert not (tmp_path / "tmp" / "simka_output_temp").exists()  # temp dir removed unless -keep-tmp


def test_simka_cli_max_reads_and_filters(tmp_path, golden_dir):
    """-max-reads / -min-read-size / -min-shannon-index through the CLI == the same selection done in Python + the oracle"""
    import gzip
    import subprocess
    from simka_b200.build import SIMKA_BIN
    out = tmp_path / "results"
    inp = os.path.join(golden_dir, "example", "simka_input.txt")
    cmd = [str(SIMKA_BIN), "-in", inp, "-out", str(out), "-out-tmp", str(tmp_path / "tmp"), "-kmer-size", "21", "-abundance-min", "1",
           "-max-reads", "50", "-min-read-size", "90", "-min-shannon-index", "1.5", "-keep-tmp", "-verbose", "0"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    L = simka_b200._lib.load()
    samples, ids = [], []
    for ident, groups in oracle.parse_simka_input(inp):
        ids.append(ident)
        from helpers import select_reads
        reads = select_reads(groups, 50, keep=lambda rd: len(rd) >= 90 and L.simka_host_shannon_index(rd, len(rd)) >= 1.5)
        from helpers import reads_to_arrays
        samples.append(reads_to_arrays(reads))
    run = oracle_run(samples, 21, 1, simple=False, complex_=False)
    names = run.matrix_names()
    for m, name in enumerate(names):
        got = gzip.open(out / (name + ".csv.gz"), "rt").read()
        assert got == run.csv(m, ids), name
    ok = (tmp_path / "tmp" / "simka_output_temp" / "count_synchro" / "A.ok").read_text().split()
    assert [int(x) for x in ok] == run.all_totals()[0].tolist()


def test_medium_scale_oracle_parity():
    """8 samples x 150k reads x 150 bp, k=31 (1.4e8 k-mers): thousands of full bins per sample, merge bins with
    sub-pass pressure; still seconds for the oracle."""
    samples = synthetic_samples(8, 150_000, n_genomes=64, genome_len=30_000)
    run = oracle_run(samples, 31, 2)
    st = simka_b200.run_samples(samples, 31, abundance_min=2)
    assert_stats_equal(st, run)


def test_full_size_sample_properties():
    """One sample of BASELINE config 2 (1M reads x 150 bp, k=31) through size-independent properties:
    with abundance-min 1 every valid k-mer is counted exactly once (sum of counts == M*(L-k+1) for N-free reads),
    chunked pushes == one push, two runs are bit-identical, and the (k-mer,count) dump is strictly increasing."""
    from simka_b200 import synth
    genomes = synth.make_genomes(256, 50_000)
    seqs, off = synth.make_sample(genomes, 0, 1_000_000, 150)
    M, L, k = 1_000_000, 150, 31

    def run(chunks):
        ctx = simka_b200.SimkaGpu(1, k, 1, simple_dist=True, complex_dist=True)
        bounds = np.linspace(0, M, chunks + 1).astype(int)
        for a, b in zip(bounds[:-1], bounds[1:]):
            p, bd, nw = simka_b200.pack_reads(seqs[int(off[a]):int(off[b])], off[a:b + 1] - off[a])
            ctx.push_reads(0, p, bd, nw, b - a)
        tot = ctx.finish_sample(0)
        km, ct = ctx.dump_sample_counts(0)
        ctx.close()
        return tot, km, ct

    tot1, km1, ct1 = run(1)
    tot3, km3, ct3 = run(3)
    assert int(tot1[0]) == M and int(tot1[2]) == M * (L - k + 1)
    assert int(tot1[1]) == len(km1) and int(ct1.sum()) == int(tot1[2])
    assert int((ct1.astype(np.uint64) ** 2).sum()) == int(tot1[3])
    assert (np.diff(km1.astype(np.int64)) > 0).all()
    np.testing.assert_array_equal(tot1, tot3)
    np.testing.assert_array_equal(km1, km3)
    np.testing.assert_array_equal(ct1, ct3)
    # canonical: no k-mer is larger than its reverse complement
    sub = km1[:: max(1, len(km1) // 50_000)]
    rc = np.array([oracle.lib().orc_revcomp(int(v), k) for v in sub[:20_000]], np.uint64)
    assert (sub[:20_000] <= rc).all()


@pytest.mark.parametrize("n,variant", [(64, ""), (70, ""), (130, ""), (70, "tiled"), (130, "tiled")])
def test_many_samples_merge_paths(monkeypatch, n, variant):
    """N = 64: largest shared-memory-accumulator merge; N > 64: global-atomics merge (default) and the 32 x 32 sample-tile
    variant (diagonal + off-diagonal tiles, a ragged last block, count-only pass for the distinct / shared totals)."""
    monkeypatch.setenv("SIMKA_B200_MERGE", variant)
    samples = synthetic_samples(n, 400, read_len=100, n_genomes=12, genome_len=2500)
    run = oracle_run(samples, 21, 2)
    st = simka_b200.run_samples(samples, 21, abundance_min=2)
    assert_stats_equal(st, run)


@pytest.mark.parametrize("n", [40, 96])
def test_merge_without_list_refinement(monkeypatch, n):
    """many small samples: the merge regroups the per-sample lists by more key bits than stage 1 could (k_refine_*, the default,
    exercised by the tests above); SIMKA_B200_REFINE=0 keeps the stage-1 granularity and re-reads each wide bin in sub-passes"""
    monkeypatch.setenv("SIMKA_B200_REFINE", "0")
    samples = synthetic_samples(n, 500, read_len=100, n_genomes=12, genome_len=2500)
    run = oracle_run(samples, 21, 2)
    st = simka_b200.run_samples(samples, 21, abundance_min=2)
    assert_stats_equal(st, run)


def test_merge_variants_agree(monkeypatch):
    """the two merge kernels (shared-memory accumulators, global atomics) on the same input"""
    samples = synthetic_samples(20, 1500, n_genomes=10, genome_len=3000)
    run = oracle_run(samples, 31, 2)
    for variant in ("", "generic"):
        monkeypatch.setenv("SIMKA_B200_MERGE", variant)
        st = simka_b200.run_samples(samples, 31, abundance_min=2)
        assert_stats_equal(st, run)


def test_simkamin_distance_edge_cases():
    """identical sketches, disjoint sketches, heavy padding, sizes around the tile length of the tiled kernel"""
    from simka_b200 import synth
    rng = np.random.default_rng(3)
    for s in (1, 7, 1535, 1536, 1537, 4000):
        n = 6
        pool = np.unique(rng.integers(1, 2**63, size=3 * s + 10, dtype=np.uint64))
        hashes = np.zeros((n, s), np.uint64)
        counts = np.zeros((n, s), np.uint32)
        for i in range(n):
            hashes[i] = np.sort(rng.choice(pool, size=s, replace=False))
            counts[i] = rng.integers(1, 1000, size=s, dtype=np.uint32)
        hashes[1] = hashes[0]                                  # identical hashes, different counts
        lo = np.sort(pool)[:s]
        hashes[2] = lo                                         # the s smallest of the pool
        if s > 3:
            hashes[3, : s // 2] = 0                            # half padding
            counts[3, : s // 2] = 0
            hashes[4, : s - 1] = 0                             # a single element
            counts[4, : s - 1] = 0
        jac, bc, _ = simka_b200.simkamin_distance(hashes, counts)
        ojac, obc = oracle.simkamin_distance(hashes, counts)
        np.testing.assert_array_equal(jac, ojac, err_msg=f"s={s}")
        np.testing.assert_array_equal(bc, obc, err_msg=f"s={s}")


def test_simkamin_distance_unbalanced_sketches():
    """sketches whose hashes are consumed at very different rates (the tile loader sizes its requests from the previous tile's
    split): dense low ranges against sparse ones, interleaved blocks, one sketch entirely below the other"""
    rng = np.random.default_rng(9)
    s = 20000
    dense = np.sort(rng.choice(np.arange(1, 40 * s, dtype=np.uint64), size=s, replace=False))            # small values, close together
    sparse = np.unique(rng.integers(1, 2**40, size=2 * s, dtype=np.uint64))[:s]                          # 2^25 times sparser
    blocks = np.sort(np.concatenate([np.arange(1 + b * 10**9, 1 + b * 10**9 + 2000, dtype=np.uint64) for b in range(10)]))  # runs of 2000, far apart
    below = np.arange(1, s + 1, dtype=np.uint64)
    above = np.arange(10**12, 10**12 + s, dtype=np.uint64)
    mixed = np.sort(np.unique(np.concatenate([dense[::2], sparse[::2]]))[:s])
    hashes = np.stack([dense, sparse, blocks, below, above, mixed])
    counts = rng.integers(1, 50, size=hashes.shape, dtype=np.uint32)
    jac, bc, _ = simka_b200.simkamin_distance(hashes, counts)
    ojac, obc = oracle.simkamin_distance(hashes, counts)
    np.testing.assert_array_equal(jac, ojac)
    np.testing.assert_array_equal(bc, obc)


def test_stage_binaries_reference_temp_formats(tmp_path, golden_dir):
    """simkaCount (one process per sample) + simkaMerge (one process per partition) of this build, exchanging the
    reference's temp files: solid/part_P/__p__I.gz records, count_synchro/*.ok, stats/part_P.gz
    long-double streams.  The partition files decode to the oracle's per-sample counts and the per-partition
    statistics sum to the oracle's accumulators."""
    import gzip
    import struct
    import subprocess
    from simka_b200.build import SIMKA_BIN
    bindir = os.path.dirname(str(SIMKA_BIN))
    inp = os.path.join(golden_dir, "example", "simka_input.txt")
    k, t, P = 31, 2, 3
    # the orchestrator lays out input/<id> albums and datasetIds (-keep-tmp)
    r = subprocess.run([str(SIMKA_BIN), "-in", inp, "-out", str(tmp_path / "res"), "-out-tmp", str(tmp_path), "-kmer-size", str(k),
                        "-abundance-min", str(t), "-keep-tmp", "-verbose", "0"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    d = str(tmp_path / "simka_output_temp")
    import shutil
    shutil.rmtree(os.path.join(d, "count_synchro"))
    spec = oracle.parse_simka_input(inp)
    ids = [s[0] for s in spec]
    for i, (ident, groups) in enumerate(spec):
        cmd = [os.path.join(bindir, "simkaCountProcess"), os.path.join(bindir, "simkaCount"), "-kmer-size", str(k), "-out-tmp-simka", d, "-out-tmp",
               os.path.join(d, "temp", ident), "-bank-name", ident, "-bank-index", str(i), "-nb-datasets", str(len(groups)), "-max-memory", "2000",
               "-nb-cores", "2", "-in", "dummy", "-abundance-min", str(t), "-abundance-max", "999999999", "-min-read-size", "0",
               "-min-shannon-index", "0.000000", "-max-reads", "0", "-nb-partitions", str(P)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
    for p in range(P):
        r = subprocess.run([os.path.join(bindir, "simkaMerge"), "-kmer-size", str(k), "-in", inp, "-out-tmp-simka", d, "-partition-id", str(p),
                            "-max-memory", "2000", "-nb-cores", "1", "-simple-dist", "-complex-dist"], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        assert os.path.exists(os.path.join(d, "merge_synchro", f"{p}.ok"))
    # oracle on the same reads
    _, samples = load_example(golden_dir)
    run = oracle_run(samples, k, t)
    N = len(ids)
    for i, ident in enumerate(ids):
        ok = [int(x) for x in open(os.path.join(d, "count_synchro", ident + ".ok")).read().split()]
        assert ok == run.all_totals()[i].tolist()
        recs = []
        for p in range(P):
            raw = gzip.open(os.path.join(d, "solid", f"part_{p}", f"__p__{i}.gz"), "rb").read()
            assert len(raw) % 24 == 0
            part = [struct.unpack_from("<QIIQ", raw, o) for o in range(0, len(raw), 24)]
            assert all(b == i for _, b, _, _ in part)
            assert [x[0] for x in part] == sorted(x[0] for x in part)  # ascending k-mers inside a partition file
            recs += [(x[0], x[3]) for x in part]
        recs.sort()
        okm, oct_ = run.sample(i)
        assert [x[0] for x in recs] == okm.tolist() and [x[1] for x in recs] == oct_.tolist()
    # per-partition statistics: x87 long double stream in the order of SimkaStatistics::save
    ns, nn = N * (N + 1) // 2, N * N
    tot = None
    for p in range(P):
        v = np.frombuffer(gzip.open(os.path.join(d, "stats", f"part_{p}.gz"), "rb").read(), np.longdouble)
        assert len(v) == 7 + 3 * N + nn + 2 * ns + (N + 3 * nn) + 3 * nn
        assert v[0] == 1 and v[1] == 1
        tot = v.copy() if tot is None else tot + v
    o = 7
    assert int(tot[4]) == run.nb_distinct_kmers and int(tot[6]) == run.nb_shared_kmers
    assert [int(x) for x in v[o:o + N]] == run.all_totals()[:, 1].tolist()             # global, not summed
    assert [int(x) for x in v[o + 2 * N:o + 3 * N]] == run.all_totals()[:, 2].tolist()
    o += 3 * N
    np.testing.assert_array_equal(tot[o:o + nn].astype(np.uint64).reshape(N, N), run.u64("nb_shared"))
    o += nn
    np.testing.assert_array_equal(tot[o:o + 2 * ns:2].astype(np.uint64), run.u64("distinct_shared"))
    np.testing.assert_array_equal(tot[o + 1:o + 2 * ns:2].astype(np.uint64), run.u64("braycurtis_num"))
    o += 2 * ns + N  # + chord_sqrt_N2
    rows = tot[o:o + 3 * nn].reshape(N, 3, N)
    np.testing.assert_array_equal(rows[:, 0].astype(np.uint64), run.u64("chord_ninj"))
    np.testing.assert_array_equal(rows[:, 1].astype(np.uint64), run.u64("hellinger"))
    np.testing.assert_array_equal(rows[:, 2].astype(np.uint64), run.u64("kulczynski"))
    o += 3 * nn
    rows = tot[o:o + 3 * nn].reshape(N, 3, N)
    np.testing.assert_array_equal(rows[:, 0].astype(np.uint64), run.u64("canberra"))
    np.testing.assert_array_equal(rows[:, 1].astype(np.uint64), run.u64("whittaker"))
    np.testing.assert_allclose(rows[:, 2].astype(np.float64), run.f64("kl"), rtol=1e-10, atol=0)


REF_BIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "bin")


@pytest.mark.skipif(not os.access(os.path.join(REF_BIN, "simka"), os.X_OK), reason="oracle/_ref/bin (reference binaries, oracle/build_ref.sh) not built")
@pytest.mark.parametrize("count,merge,k,t", [("ours", "ours", 31, 2), ("ours", "ref", 21, 0), ("ref", "ours", 21, 0), ("ref", "ours", 31, 2)])
def test_cross_run_with_reference_binaries(tmp_path, golden_dir, count, merge, k, t, inp="example", tag=""):
    """The REFERENCE's `simka` orchestrator (oracle/_ref/bin/simka, unmodified) drives the stage binaries it finds next to
    itself (src/SimkaPotara.hpp:218-228): this build's simkaCount / simkaMerge, or one stage of each -- the temp files
    (solid/part_P/__p__I.gz, count_synchro/*.ok, stats/part_P.gz, merge_synchro) are what crosses between them.  Every mix
    must reproduce the reference's truth matrices byte for byte."""
    import gzip
    import shutil
    import subprocess
    from simka_b200.build import SIMKA_BIN
    ours = os.path.dirname(str(SIMKA_BIN))
    b = tmp_path / "bin"
    b.mkdir()
    shutil.copy(os.path.join(REF_BIN, "simka"), b / "simka")
    for exe, side in (("simkaCount", count), ("simkaCountProcess", count), ("simkaMerge", merge)):
        os.symlink(os.path.join(ours if side == "ours" else REF_BIN, exe), b / exe)
    ex = tmp_path / "ex"
    shutil.copytree(os.path.join(golden_dir, inp), ex)
    out = tmp_path / "out"
    tdir = os.path.join(golden_dir, "truth", f"results_k{k}_t{t}") if inp == "example" else \
        os.path.join(golden_dir, "truth_ref", f"{inp}_results_k{k}_t{t}" + ("_" + tag if tag else ""))
    extra = json.load(open(os.path.join(tdir, "options.json"))) if tag else []
    # (own process group and a time limit: the reference orchestrator waits for ever for the token of a stage process that died)
    proc = subprocess.Popen([str(b / "simka"), "-in", str(ex / "simka_input.txt"), "-out", str(out), "-out-tmp", str(tmp_path / "tmp"), "-kmer-size", str(k),
                             "-abundance-min", str(t), "-simple-dist", "-complex-dist", "-nb-cores", "4", "-max-memory", "4000", "-keep-tmp"] + extra,
                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, start_new_session=True)
    logs = tmp_path / "tmp" / "simka_output_temp" / "log"
    try:
        stdout, _ = proc.communicate(timeout=300)
    except subprocess.TimeoutExpired:
        os.killpg(proc.pid, signal.SIGKILL)
        stdout, _ = proc.communicate()
        tails = "".join(f"--- {f}\n" + open(logs / f, errors="replace").read()[-600:] for f in sorted(os.listdir(logs))[:10]) if logs.is_dir() else ""
        pytest.fail("the orchestrator did not finish\n" + stdout[-1500:] + tails)
    assert proc.returncode == 0, stdout[-2000:]
    first = "A" if inp == "example" else "S1"
    assert str(b / "simkaCount") in open(logs / f"count_{first}.txt").read()  # the orchestrator did start the binaries of this mix
    checked = 0
    for f in sorted(os.listdir(tdir)):
        if f.endswith(".csv"):
            assert gzip.open(out / (f + ".gz"), "rt").read() == open(os.path.join(tdir, f)).read(), f
            checked += 1
    assert checked == (20 if inp == "example" else 21)


@pytest.mark.skipif(not os.access(os.path.join(REF_BIN, "simka"), os.X_OK), reason="oracle/_ref/bin (reference binaries, oracle/build_ref.sh) not built")
@pytest.mark.parametrize("count,merge,k,t,tag", [("ours", "ours", 21, 2, ""), ("ours", "ref", 15, 0, ""), ("ref", "ours", 12, 2, ""), ("ours", "ours", 31, 2, "maxreads50"),
                                                 ("ours", "ref", 21, 1, "minsize90_shannon1.5"), ("ours", "ours", 21, 1, "maxreads0"), ("ref", "ours", 21, 3, "amax3")])
def test_cross_run_with_reference_binaries_on_awkward_input(tmp_path, golden_dir, count, merge, k, t, tag):
    """the same on tests/golden/example_n (N runs, lower case, reads shorter than k, multi-line FASTA, FASTQ.gz, a paired and a
    concatenated sample), with k <= 15 (the MiniKC stage of the reference's simkaCount) and with read limits / filters passed
    down by the orchestrator: truth = what the reference binaries alone produced (tests/golden/truth_ref)"""
    test_cross_run_with_reference_binaries(tmp_path, golden_dir, count, merge, k, t, inp="example_n", tag=tag)


def test_import_counts_equals_counting():
    """a sample registered from its (k-mer, count) list behaves exactly like a counted one"""
    samples = synthetic_samples(5, 2500)
    run = oracle_run(samples, 31, 2)
    ctx = simka_b200.SimkaGpu(len(samples), 31, 2, simple_dist=True, complex_dist=True)
    for s in range(len(samples)):
        km, ct = run.sample(s)
        perm = np.random.default_rng(s).permutation(len(km))
        ctx.import_counts(s, km[perm], ct[perm], nb_reads=2500)
    ctx.merge_and_accumulate()
    st = ctx.stats()
    ctx.close()
    assert_stats_equal(st, run)


@pytest.mark.parametrize("k,max_reads,nb_kmers", [(21, 0, 100), (21, 100, 1000), (31, 0, 1000), (31, 100, 100)])
def test_simkamin_gpu_sketch_and_distance_reproduce_reference_truth(golden_dir, k, max_reads, nb_kmers):
    """SimkaMin end to end on the GPU (sketch kernel path + distance kernel) against the reference's golden matrices
    (tests/simkaMin/truth_simkaMin_symetrical of the reference), and sketch-for-sketch against the oracle."""
    from test_oracle_golden import simkamin_csv, simkamin_example_sketches
    ids, gh, gc = simkamin_example_sketches(golden_dir, k, max_reads, nb_kmers,
                                            lambda s, o, kk, n: simka_b200.simkamin_sketch(s, o, kk, n, seed=100))
    _, oh, oc = simkamin_example_sketches(golden_dir, k, max_reads, nb_kmers, lambda s, o, kk, n: oracle.simkamin_sketch(s, o, kk, n, seed=100))
    np.testing.assert_array_equal(gh, oh)
    np.testing.assert_array_equal(gc, oc)
    jac, bc, _ = simka_b200.simkamin_distance(gh, gc)
    tdir = os.path.join(golden_dir, "truth_simkaMin", f"k{k}__{max_reads}-{nb_kmers}_n1")
    assert simkamin_csv(ids, jac) == open(os.path.join(tdir, "mat_presenceAbsence_jaccard.csv")).read()
    assert simkamin_csv(ids, bc) == open(os.path.join(tdir, "mat_abundance_braycurtis.csv")).read()


@pytest.mark.parametrize("k,max_reads,nb_kmers", [(21, 0, 100), (21, 100, 1000), (31, 0, 1000), (31, 100, 100)])
def test_simkamin_gpu_filter_sketch_reproduces_reference_truth(golden_dir, k, max_reads, nb_kmers):
    """the -filter option end to end on the GPU against the reference's k*_filter_* golden matrices, sketch for sketch against the oracle"""
    from test_oracle_golden import simkamin_csv, simkamin_example_sketches
    ids, gh, gc = simkamin_example_sketches(golden_dir, k, max_reads, nb_kmers,
                                            lambda s, o, kk, n: simka_b200.simkamin_sketch(s, o, kk, n, seed=100, filter=True))
    _, oh, oc = simkamin_example_sketches(golden_dir, k, max_reads, nb_kmers,
                                          lambda s, o, kk, n: oracle.simkamin_sketch(s, o, kk, n, seed=100, filter=True))
    np.testing.assert_array_equal(gh, oh)
    np.testing.assert_array_equal(gc, oc)
    jac, bc, _ = simka_b200.simkamin_distance(gh, gc)
    tdir = os.path.join(golden_dir, "truth_simkaMin", f"k{k}_filter_{max_reads}-{nb_kmers}_n1")
    assert simkamin_csv(ids, jac) == open(os.path.join(tdir, "mat_presenceAbsence_jaccard.csv")).read()
    assert simkamin_csv(ids, bc) == open(os.path.join(tdir, "mat_abundance_braycurtis.csv")).read()


def test_simkamin_gpu_filter_sketch_synthetic():
    """-filter on larger inputs: sketch smaller / larger than the number of repeated k-mers, duplicates, N runs"""
    samples = synthetic_samples(2, 4000) + nasty_samples()[:3]
    for seqs, off in samples:
        for k, s in [(21, 300), (31, 5000), (31, 10**6), (15, 64)]:
            gh, gc, gn = simka_b200.simkamin_sketch(seqs, off, k, s, seed=7, filter=True)
            oh, oc, on = oracle.simkamin_sketch(seqs, off, k, s, seed=7, filter=True)
            assert gn == on
            np.testing.assert_array_equal(gh, oh)
            np.testing.assert_array_equal(gc, oc)


def test_simkamin_gpu_sketch_synthetic():
    """larger inputs: sketch smaller than, equal to and larger than the number of distinct k-mers; N runs; duplicates"""
    samples = synthetic_samples(2, 4000) + nasty_samples()[:3]
    for seqs, off in samples:
        for k, s in [(21, 500), (31, 20000), (31, 10**6), (15, 64)]:
            gh, gc, gn = simka_b200.simkamin_sketch(seqs, off, k, s, seed=7)
            oh, oc, on = oracle.simkamin_sketch(seqs, off, k, s, seed=7)
            assert gn == on
            np.testing.assert_array_equal(gh, oh)
            np.testing.assert_array_equal(gc, oc)


@pytest.mark.gpu
@pytest.mark.parametrize("k,max_reads,nb_kmers,filt", [(21, 0, 100, False), (31, 100, 1000, False), (21, 100, 100, True), (31, 0, 1000, True)])
def test_simkamincore_cli_reproduces_reference_truth(tmp_path, golden_dir, k, max_reads, nb_kmers, filt):
    """the reference's SimkaMin test (tests/simkaMin/test_simkaMin.py:141-154) against the `simkaMinCore` binary of this
    build, driven as simkaMin.py drives it (sketch, distance, export: simkaMin/simkaMin.py:101-197): the csv files are
    byte-identical to the truth; then the same matrix computed in rectangular -start-i/-start-j blocks."""
    import gzip
    import subprocess
    from simka_b200.build import SIMKA_BIN
    exe = os.path.join(os.path.dirname(str(SIMKA_BIN)), "simkaMinCore")
    sk = str(tmp_path / "sketch.bin")
    r = subprocess.run([exe, "sketch", "-in", os.path.join(golden_dir, "example", "simka_input.txt"), "-out", sk, "-seed", "100", "-kmer-size", str(k),
                        "-nb-kmers", str(nb_kmers), "-max-reads", str(max_reads), "-min-read-size", "0", "-min-shannon-index", "0",
                        "-nb-cores", "0", "-max-memory", "8000"] + (["-filter"] if filt else []), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert os.path.getsize(sk) == 13 + 5 * nb_kmers * 16 + sum(1 + len(i) for i in "ABCDE")
    dist = str(tmp_path / "distance")
    r = subprocess.run([exe, "distance", "-in1", sk, "-in2", sk, "-out", dist, "-nb-cores", "0", "-start-i", "0", "-start-j", "0", "-n-i", "5", "-n-j", "5"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run([exe, "export", "-in", dist, "-in1", sk, "-in2", sk, "-out", str(tmp_path / "res"), "-nb-cores", "0"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    tdir = os.path.join(golden_dir, "truth_simkaMin", f"k{k}_{'filter' if filt else ''}_{max_reads}-{nb_kmers}_n1")
    for f in ("mat_presenceAbsence_jaccard", "mat_abundance_braycurtis"):
        assert gzip.open(tmp_path / "res" / (f + ".csv.gz"), "rt").read() == open(os.path.join(tdir, f + ".csv")).read(), f
    # blocks of 2 + 3 datasets as simkaMin.py would issue them with MAX_DATASETS_PROCESS = 2..3: (0,0) (0,1) (1,1)
    dist2 = str(tmp_path / "distance_blocks")
    for si, sj, ni, nj in ((0, 0, 2, 2), (0, 2, 2, 3), (2, 2, 3, 3)):
        r = subprocess.run([exe, "distance", "-in1", sk, "-in2", sk, "-out", dist2, "-start-i", str(si), "-start-j", str(sj), "-n-i", str(ni), "-n-j", str(nj)],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
    for f in ("mat_presenceAbsence_jaccard.bin", "mat_abundance_braycurtis.bin"):
        whole = np.fromfile(os.path.join(dist, f), dtype=np.float32).reshape(5, 5)
        blocks = np.fromfile(os.path.join(dist2, f), dtype=np.float32).reshape(5, 5)
        np.testing.assert_array_equal(np.triu(blocks), np.triu(whole))  # the reference does not mirror rectangular blocks
        np.testing.assert_array_equal(blocks[2:, 2:], whole[2:, 2:])
        np.testing.assert_array_equal(blocks[:2, :2], whole[:2, :2])


SIMKAMIN_REF_CASES = [["-kmer-size", "21", "-nb-kmers", "500"], ["-kmer-size", "31", "-nb-kmers", "300", "-max-reads", "30"],
                      ["-kmer-size", "21", "-nb-kmers", "400", "-filter"], ["-kmer-size", "25", "-nb-kmers", "5000"],
                      ["-kmer-size", "21", "-nb-kmers", "200", "-min-read-size", "95", "-min-shannon-index", "1.8"],
                      ["-kmer-size", "15", "-nb-kmers", "300"], ["-kmer-size", "31", "-nb-kmers", "100", "-seed", "7"]]


@pytest.mark.parametrize("case", range(len(SIMKAMIN_REF_CASES)))
def test_simkamincore_against_reference_binary(tmp_path, golden_dir, case):
    """`simkaMinCore sketch | distance | export` of this build next to the REFERENCE simkaMinCore (oracle/_ref/bin) on the same input --
    multi-line FASTA with lower case and N runs, reads shorter than k, FASTQ.gz, a paired and a concatenated sample, the file-format
    corner cases of example_fmt -- and the same options: the sketch
    files are equal byte for byte outside the 4 padding bytes of every 16-byte record (the reference writes them uninitialised),
    the exported matrices byte for byte; and each build's `distance` reads the other's sketch file."""
    import gzip
    import subprocess
    from simka_b200.build import SIMKA_BIN
    SIMKAMIN_BIN = SIMKA_BIN.parent / "simkaMinCore"
    ref = os.path.join(REF_BIN, "simkaMinCore")
    if not os.access(ref, os.X_OK):
        pytest.skip("oracle/_ref/bin not built")
    # S1 and F2 hold reads shorter than k / header-only records: SelectKmersCommand then counts the previous read again
    # (simkamin_short_read_replay; deterministic in the reference with one core, which is what both are run with)
    ex, fm = os.path.join(golden_dir, "example_n"), os.path.join(golden_dir, "example_fmt")
    inp = tmp_path / "in.txt"
    inp.write_text(f"S1: {ex}/S1.fasta\nS2: {ex}/S2.fasta\nS3: {ex}/S3.fastq.gz\nS4: {ex}/S4_1.fasta ; {ex}/S4_2.fasta\nS6: {ex}/S2.fasta , {ex}/S4_1.fasta\n"
                   f"F0: {fm}/nastyq.fastq\nF2: {fm}/odd.fasta\nF3: {fm}/iupac.fasta\nF4: {fm}/multi.fasta.gz\n")
    opts = SIMKAMIN_REF_CASES[case]

    def run(exe, tag, sketch=None):
        d = tmp_path / tag
        (d / "dist").mkdir(parents=True)
        sk = sketch or str(d / "sketch.bin")
        if not sketch:
            r = subprocess.run([exe, "sketch", "-in", str(inp), "-out", sk, "-nb-cores", "1", "-max-memory", "4000"] + opts, capture_output=True, text=True, timeout=300)
            assert r.returncode == 0, r.stdout + r.stderr
        for cmd in (["distance", "-in1", sk, "-in2", sk, "-out", str(d / "dist"), "-nb-cores", "1"],
                    ["export", "-in", str(d / "dist"), "-in1", sk, "-in2", sk, "-out", str(d), "-nb-cores", "1"]):
            r = subprocess.run([exe] + cmd, capture_output=True, text=True, timeout=300)
            assert r.returncode == 0, r.stdout + r.stderr
        return sk, {f: gzip.open(d / f).read() for f in sorted(os.listdir(d)) if f.startswith("mat_") and f.endswith(".csv.gz")}

    sk_ref, m_ref = run(ref, "ref")
    sk_our, m_our = run(str(SIMKAMIN_BIN), "our")
    assert len(m_ref) == 2 and m_our == m_ref
    a, b = np.fromfile(sk_ref, np.uint8), np.fromfile(sk_our, np.uint8)
    assert a.size == b.size
    differing = np.nonzero(a != b)[0]
    # 13 header bytes (k, sketch size, seed, number of datasets), then the records {u64 hash, u32 count, 4 padding bytes}, then the identifiers
    n_rec = 9 * int(opts[opts.index("-nb-kmers") + 1])
    rec0 = 13
    assert differing.size == 0 or (differing.min() >= rec0 and differing.max() < rec0 + 16 * n_rec and
                                   set(((differing - rec0) % 16).tolist()) <= {12, 13, 14, 15})
    assert run(ref, "ref_on_ours", sketch=sk_our)[1] == m_ref
    assert run(str(SIMKAMIN_BIN), "ours_on_ref", sketch=sk_ref)[1] == m_ref


def test_simkamincore_file_tools_against_reference_binary(tmp_path, golden_dir):
    """`simkaMinCore info | distance (blocks into an existing result directory) | matrix-update | append | export` next to the REFERENCE
    binary on the same two sketch files: same texts, binary matrices, appended sketch file (outside the reference's uninitialised
    padding bytes) and exported matrices"""
    import gzip
    import shutil
    import subprocess
    from simka_b200.build import SIMKA_BIN
    ours = str(SIMKA_BIN.parent / "simkaMinCore")
    ref = os.path.join(REF_BIN, "simkaMinCore")
    if not os.access(ref, os.X_OK):
        pytest.skip("oracle/_ref/bin not built")
    ex = os.path.join(golden_dir, "example_n")
    (tmp_path / "a.txt").write_text(f"S2: {ex}/S2.fasta\nS3: {ex}/S3.fastq.gz\n")
    (tmp_path / "b.txt").write_text(f"S4: {ex}/S4_1.fasta ; {ex}/S4_2.fasta\nS6: {ex}/S2.fasta , {ex}/S4_1.fasta\nS1: {ex}/S1.fasta\n")

    def masked(path):  # 13 header bytes, n x s records {u64, u32, 4 padding bytes}, identifiers
        raw = np.fromfile(path, np.uint8).copy()
        s, n = int(raw[1:5].view(np.uint32)[0]), int(raw[9:13].view(np.uint32)[0])
        raw[13:13 + 16 * n * s].reshape(-1, 16)[:, 12:] = 0
        return raw.tobytes()

    res = {}
    for tag, exe in (("ref", ref), ("our", ours)):
        d = tmp_path / tag
        (d / "dist").mkdir(parents=True)

        def run(*a):
            r = subprocess.run([exe] + [str(x) for x in a], capture_output=True, text=True, timeout=300)
            assert r.returncode == 0, (tag, a, r.stdout + r.stderr)
            return r.stdout.replace(str(d), "<dir>")

        o = {}
        for name in ("a", "b"):
            run("sketch", "-in", tmp_path / f"{name}.txt", "-out", d / f"{name}.bin", "-kmer-size", "21", "-nb-kmers", "300", "-nb-cores", "1", "-max-memory", "4000")
        o["info"] = run("info", "-in", d / "a.bin")
        run("distance", "-in1", d / "a.bin", "-in2", d / "a.bin", "-out", d / "dist", "-nb-cores", "1")
        run("distance", "-in1", d / "a.bin", "-in2", d / "b.bin", "-out", d / "dist", "-nb-cores", "1")
        run("distance", "-in1", d / "b.bin", "-in2", d / "b.bin", "-out", d / "dist", "-nb-cores", "1")
        o["update"] = run("matrix-update", "-in", d / "dist", "-in1", d / "a.bin", "-in2", d / "b.bin")
        o["append"] = run("append", "-in1", d / "a.bin", "-in2", d / "b.bin")
        o["info2"] = run("info", "-in", d / "a.bin")
        run("export", "-in", d / "dist", "-in1", d / "a.bin", "-in2", d / "a.bin", "-out", d, "-nb-cores", "1")
        o["sketch"] = masked(d / "a.bin")
        o["bins"] = {f: open(d / "dist" / f, "rb").read() for f in sorted(os.listdir(d / "dist"))}
        o["csv"] = {f: gzip.open(d / f).read() for f in sorted(os.listdir(d)) if f.endswith(".csv.gz")}
        res[tag] = o
    assert len(res["ref"]["bins"]) == 2 and len(res["ref"]["csv"]) == 2 and "Nb Datasets: 5" in res["ref"]["info2"]
    for key in ("info", "update", "append", "info2", "sketch", "bins", "csv"):
        assert res["our"][key] == res["ref"][key], key


@pytest.mark.skipif(not os.access(os.path.join(REF_BIN, "simka"), os.X_OK), reason="oracle/_ref/bin (reference binaries, oracle/build_ref.sh) not built")
@pytest.mark.parametrize("inp,k,t,mine", [("example", 21, 2, ("B", "D")), ("example_n", 15, 0, ("S2", "S4", "S5")), ("example", 31, 0, ("A", "B", "C", "D", "E"))])
def test_mixed_per_sample_counting_with_the_reference(tmp_path, golden_dir, inp, k, t, mine):
    """ONE run of the reference orchestrator in which some samples are counted by this build's simkaCount and the others by the
    reference's: this build reads the run's config.h5 (host/repartitor.hpp: minimal HDF5 reader + Repartitor table) and sends every
    solid k-mer to the partition the reference would -- the matrices equal the truth, and the partition files written here hold
    exactly the records the reference's simkaCount writes for the same sample."""
    import gzip
    import shutil
    import subprocess
    from simka_b200.build import SIMKA_BIN
    ours = os.path.dirname(str(SIMKA_BIN))

    def run(tag, mine_ids):
        b = tmp_path / ("bin_" + tag)
        b.mkdir()
        shutil.copy(os.path.join(REF_BIN, "simka"), b / "simka")
        os.symlink(os.path.join(REF_BIN, "simkaCountProcess"), b / "simkaCountProcess")
        os.symlink(os.path.join(REF_BIN, "simkaMerge"), b / "simkaMerge")
        pick = "|".join(mine_ids) if mine_ids else "__none__"
        (b / "simkaCount").write_text("#!/bin/bash\nname=\"\"; prev=\"\"\nfor a in \"$@\"; do if [ \"$prev\" = \"-bank-name\" ]; then name=\"$a\"; fi; prev=\"$a\"; done\n"
                                      f"case \"$name\" in {pick}) exec {ours}/simkaCount \"$@\";; *) exec {REF_BIN}/simkaCount \"$@\";; esac\n")
        (b / "simkaCount").chmod(0o755)
        ex = tmp_path / ("ex_" + tag)
        shutil.copytree(os.path.join(golden_dir, inp), ex)
        proc = subprocess.Popen([str(b / "simka"), "-in", str(ex / "simka_input.txt"), "-out", str(tmp_path / ("out_" + tag)), "-out-tmp", str(tmp_path / ("tmp_" + tag)),
                                 "-kmer-size", str(k), "-abundance-min", str(t), "-simple-dist", "-complex-dist", "-nb-cores", "4", "-max-memory", "4000", "-keep-tmp"],
                                stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, start_new_session=True)
        try:
            stdout, _ = proc.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            os.killpg(proc.pid, signal.SIGKILL)
            stdout, _ = proc.communicate()
            pytest.fail("the orchestrator did not finish\n" + stdout[-1500:])
        assert proc.returncode == 0, stdout[-2000:]
        d = tmp_path / ("tmp_" + tag) / "simka_output_temp"
        parts = {}
        for p in sorted(os.listdir(d / "solid")):
            for f in sorted(os.listdir(d / "solid" / p)):
                parts[(p, f)] = gzip.open(d / "solid" / p / f).read()
        counts = {f: open(d / "kmercount_per_partition" / f).read() for f in sorted(os.listdir(d / "kmercount_per_partition"))}
        mats = {f: gzip.open(tmp_path / ("out_" + tag) / f).read() for f in sorted(os.listdir(tmp_path / ("out_" + tag)))}
        logs = {f: open(d / "log" / f, errors="replace").read() for f in os.listdir(d / "log") if f.startswith("count_")}
        return parts, counts, mats, logs

    parts, counts, mats, logs = run("mixed", mine)
    first = mine[0]
    assert "simkaCount" in logs[f"count_{first}.txt"] and "DSK" not in logs[f"count_{first}.txt"]  # (this build's binary did count that sample)
    tdir = os.path.join(golden_dir, "truth", f"results_k{k}_t{t}") if inp == "example" else os.path.join(golden_dir, "truth_ref", f"{inp}_results_k{k}_t{t}")
    checked = 0
    for f in sorted(os.listdir(tdir)):
        if f.endswith(".csv"):
            assert mats[f + ".gz"].decode() == open(os.path.join(tdir, f)).read(), f
            checked += 1
    assert checked >= 20
    ref_parts, ref_counts, _, _ = run("ref", ())
    assert sorted(parts) == sorted(ref_parts) and counts == ref_counts
    rec = np.dtype([("k", "<u8"), ("b", "<u4"), ("pad", "<u4"), ("c", "<u8")])
    for key in parts:  # the 4 padding bytes of a record are not compared
        a, b = np.frombuffer(parts[key], rec), np.frombuffer(ref_parts[key], rec)
        assert len(a) == len(b) and all(np.array_equal(a[x], b[x]) for x in ("k", "b", "c")), key
