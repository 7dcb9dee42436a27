"""Shared helpers for the parity tests: seeded inputs and oracle <-> product comparison."""
import numpy as np

from oracle import oracle
from simka_b200 import synth


def reads_to_arrays(reads):
    """list of bytes -> (seqs uint8, read_off uint64)"""
    lens = np.array([len(r) for r in reads], np.uint64)
    off = np.zeros(len(reads) + 1, np.uint64)
    np.cumsum(lens, out=off[1:])
    seqs = np.frombuffer(b"".join(reads), np.uint8) if reads else np.zeros(0, np.uint8)
    return seqs, off


def synthetic_samples(n_samples, n_reads, read_len=150, n_genomes=16, genome_len=4000, err=0.01):
    genomes = synth.make_genomes(n_genomes, genome_len)
    return [synth.make_sample(genomes, s, n_reads, read_len, err) for s in range(n_samples)]


def nasty_samples(seed=3):
    """The parity fixture of SURVEY.md 8(d): N runs, reads shorter than k, lowercase, empty reads,
    an empty sample, an empty-intersection pair and a duplicated sample."""
    rng = np.random.default_rng(seed)
    alpha = np.frombuffer(b"ACGT", np.uint8)

    def rnd(n):
        return alpha[rng.integers(0, 4, n)].tobytes()

    g1, g2, g3 = rnd(3000), rnd(3000), rnd(2000)

    def reads_from(g, n, L=100, p_n=0.0, lower=False):
        out = []
        for _ in range(n):
            s = int(rng.integers(0, len(g) - L))
            r = bytearray(g[s:s + L])
            if p_n and rng.random() < p_n:
                q = int(rng.integers(0, L - 5))
                r[q:q + int(rng.integers(1, 5))] = b"N" * 4
                r = r[:L]
            if lower and rng.random() < 0.5:
                r = bytearray(bytes(r).lower())
            if rng.random() < 0.5:
                comp = bytes.maketrans(b"ACGTacgtN", b"TGCAtgcaN")
                r = bytearray(bytes(r).translate(comp)[::-1])
            out.append(bytes(r))
        return out

    s0 = reads_from(g1, 400, p_n=0.3) + [b"ACGT", b"", b"N" * 50, rnd(30), rnd(31), rnd(32)]
    s1 = reads_from(g1, 300, lower=True) + reads_from(g2, 200)
    s2 = reads_from(g2, 500, L=150, p_n=0.1)
    s3 = reads_from(g3, 300)           # shares nothing with the others (w.h.p.)
    s4 = list(s1)                      # duplicate of s1 -> distance 0
    s5 = []                            # empty sample
    s6 = [b"A" * 200, b"T" * 200, b"AC" * 100] * 20 + reads_from(g1, 50)  # low complexity, heavy duplicates
    return [reads_to_arrays(s) for s in (s0, s1, s2, s3, s4, s5, s6)]


def oracle_run(samples, k, amin, amax=999999999, simple=True, complex_=True):
    seqs = np.concatenate([s[0] for s in samples]) if samples else np.zeros(0, np.uint8)
    offs = [np.zeros(1, np.uint64)]
    sample_off = [0]
    base = 0
    nreads = 0
    for sq, off in samples:
        offs.append(off[1:] - off[0] + np.uint64(base))
        base += int(off[-1] - off[0])
        nreads += len(off) - 1
        sample_off.append(nreads)
    run = oracle.Run(seqs, np.concatenate(offs), np.array(sample_off, np.uint64), k, abundance_min=amin,
                     abundance_max=amax, simple=simple, complex_=complex_)
    run.merge()
    return run


def assert_stats_equal(st, run, kl_rtol=1e-10):
    """Integer accumulators bit-exact; Jensen-Shannon accumulator within kl_rtol relative."""
    n = run.n
    np.testing.assert_array_equal(st.totals, run.all_totals())
    assert st.nb_distinct_kmers == run.nb_distinct_kmers
    assert st.nb_shared_kmers == run.nb_shared_kmers
    np.testing.assert_array_equal(st.distinct_shared, run.u64("distinct_shared"))
    np.testing.assert_array_equal(st.braycurtis_num, run.u64("braycurtis_num"))
    np.testing.assert_array_equal(st.nb_shared, run.u64("nb_shared"))
    if run.simple:
        np.testing.assert_array_equal(st.chord_ninj, run.u64("chord_ninj"))
        np.testing.assert_array_equal(st.hellinger, run.u64("hellinger"))
        np.testing.assert_array_equal(st.kulczynski, run.u64("kulczynski"))
    if run.complex_:
        np.testing.assert_array_equal(st.canberra, run.u64("canberra"))
        np.testing.assert_array_equal(st.whittaker, run.u64("whittaker"))
        np.testing.assert_allclose(st.kullback_leibler, run.f64("kl"), rtol=kl_rtol, atol=0)
    # the reference stores distances as float and prints %f: matrices must be identical floats,
    # except Jensen-Shannon which inherits the 1e-10 tolerance of its accumulator
    names = run.matrix_names()
    assert st.matrix_names() == names
    for m, name in enumerate(names):
        a, b = st.matrix(m), run.matrix(m)
        if name == "mat_abundance_jensenshannon":
            np.testing.assert_allclose(a, b, rtol=1e-6, atol=0, err_msg=name)
        else:
            np.testing.assert_array_equal(a, b, err_msg=name)


def select_reads(groups, max_reads, keep=lambda rd: True):
    """Read selection of the reference's SimkaInputIterator (src/core/SimkaCommons.hpp:150-314): per ';' group the counter
    skips the first read of every file and the read that makes it reach max_reads is dropped (0 = all reads)."""
    reads = []
    for grp in groups:
        count, stop = 0, False
        for f in grp:
            if stop:
                break
            d, o = oracle.read_fastx(f)
            first = True
            for i in range(len(o) - 1):
                rd = bytes(d[int(o[i]):int(o[i + 1])])
                if not keep(rd):
                    continue
                if first:
                    first = False
                else:
                    count += 1
                    if max_reads and count >= max_reads:
                        stop = True
                        break
                reads.append(rd)
    return reads
