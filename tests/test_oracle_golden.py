"""CPU tests: the oracle against every golden vector the reference's own tests hold for this path
(SURVEY.md 8c).  These pin the oracle; the GPU tests then compare the CUDA path with it."""
import os

import numpy as np
import pytest

from oracle import oracle

# count_synchro/<id>.ok known answers of a reference run: reads distinct sum sum^2
KNOWN_TOTALS = {
    (31, 2): {"A": (99, 3840, 7920, 16560), "B": (102, 880, 1840, 3920), "C": (98, 240, 560, 1360),
              "D": (146, 880, 2080, 6080), "E": (402, 7200, 32160, 188480)},
    (31, 0): {"B": (102, 7200, 8160, 10240), "C": (98, 7520, 7840, 8640), "D": (146, 10480, 11680, 15680)},
    (21, 2): {"A": (99, 4320, 8910, 18630), "B": (102, 990, 2070, 4410), "C": (98, 270, 630, 1530),
              "D": (146, 990, 2340, 6840), "E": (402, 8100, 36180, 212040)},
    (21, 0): {"B": (102, 8100, 9180, 11520), "C": (98, 8460, 8820, 9720), "D": (146, 11790, 13140, 17640)},
}


def test_kmer_model_known_answers():
    """gatb-core test/unit/src/kmer/TestKmer.cpp:141-147,178-191"""
    seq = b"CATTGATAGTGG"
    km, va = oracle.direct_kmers(seq, 3)
    assert km.tolist() == [18, 10, 43, 44, 50, 8, 35, 14, 59, 47] and va.all()
    km, va = oracle.canonical_kmers(seq, 3)
    assert km.tolist() == [11, 2, 16, 36, 9, 8, 24, 6, 17, 20] and va.all()


BADCHAR_SEQ = b"ACGNCNTGCTAGCTATTTAGCTTTAGANAGTAGATGACGCNC"
BADCHAR_VALID = [False] * 6 + [True] * 11 + [False] * 11 + [True, True, False, False]


def kmer_to_string(v, k):
    """Model::toString: first base most significant, A=0 C=1 T=2 G=3"""
    return "".join("ACTG"[(int(v) >> (2 * (k - 1 - i))) & 3] for i in range(k))


def test_kmer_badchar_known_answers():
    """gatb-core test/unit/src/kmer/TestKmer.cpp:509-569 (kmer_badchar): ModelDirect, k = 11, validity of the 32 windows and the
    VALUE of every window, valid or not -- a bad nucleotide reads as G (its (c>>1)&3 code)"""
    km, va = oracle.direct_kmers(BADCHAR_SEQ, 11)
    assert va.astype(bool).tolist() == BADCHAR_VALID
    for i in range(32):
        assert kmer_to_string(km[i], 11) == BADCHAR_SEQ[i:i + 11].decode().replace("N", "G")


def test_kmer_model_bad_chars():
    """windows containing a non-ACGT char are invalid (Model.hpp:754-758); lower case is valid"""
    km, va = oracle.canonical_kmers(b"ACGTNACGTAC", 4)
    assert va.tolist() == [True, False, False, False, False, True, True, True]
    a, _ = oracle.canonical_kmers(b"acgtacgt", 4)
    b, _ = oracle.canonical_kmers(b"ACGTACGT", 4)
    assert a.tolist() == b.tolist()
    assert len(oracle.canonical_kmers(b"ACG", 4)[0]) == 0


def test_revcomp_and_minimizer():
    L = oracle.lib()
    # AAAC -> GTTT ; code A=0 C=1 T=2 G=3
    assert L.orc_revcomp(0b00000001, 4) == 0b11101010
    # minimizer never exceeds the default 4^m-1 and is strand-independent
    rng = np.random.default_rng(1)
    for _ in range(200):
        x = int(rng.integers(0, 2**62))
        rc = L.orc_revcomp(x, 31)
        m1, m2 = L.orc_minimizer(x, 31, 7), L.orc_minimizer(rc, 31, 7)
        assert m1 == m2 and m1 <= 4**7 - 1


@pytest.mark.parametrize("k", [21, 31])
@pytest.mark.parametrize("t", [0, 2])
def test_truth_matrices_byte_identical(golden_dir, k, t):
    """reference tests/simple_test.py:93-119: 20 matrices per (k, abundance-min), byte compare"""
    run, ids = oracle.run_from_input_file(os.path.join(golden_dir, "example", "simka_input.txt"), k, abundance_min=t)
    run.merge()
    assert ids == ["A", "B", "C", "D", "E"]
    tdir = os.path.join(golden_dir, "truth", f"results_k{k}_t{t}")
    names = run.matrix_names()
    assert len(names) == 21
    checked = 0
    for m, name in enumerate(names):
        path = os.path.join(tdir, name + ".csv")
        if not os.path.exists(path):
            assert name == "mat_abundance_jaccard"  # produced by the reference but it ships no truth file
            continue
        assert run.csv(m, ids) == open(path).read(), name
        checked += 1
    assert checked == 20
    tot = run.all_totals()
    for ident, exp in KNOWN_TOTALS[(k, t)].items():
        assert tuple(int(x) for x in tot[ids.index(ident)]) == exp
    # consequence of the u64 Canberra accumulator
    np.testing.assert_array_equal(run.matrix("mat_abundance_canberra"), run.matrix("mat_presenceAbsence_jaccard"))


def test_partition_independence(golden_dir):
    """reference tests/simple_test.py:122-129: results must not depend on the resources / split"""
    path = os.path.join(golden_dir, "example", "simka_input.txt")
    a, ids = oracle.run_from_input_file(path, 21, abundance_min=0, nthreads=1)
    b, _ = oracle.run_from_input_file(path, 21, abundance_min=0, nthreads=4)
    a.merge(), b.merge()
    for m in range(len(a.matrix_names())):
        assert a.csv(m, ids) == b.csv(m, ids)


def test_fastx_reader_variants(tmp_path):
    """multi-line FASTA, FASTQ, gz, CRLF, empty lines: kseq rules of BankFasta.cpp:485-572"""
    import gzip
    fa = tmp_path / "a.fa"
    fa.write_bytes(b">r1 comment\nACGT\nACGT\n\n>r2\r\nTTTT\r\n>r3\n")
    d, o = oracle.read_fastx(str(fa))
    assert [bytes(d[o[i]:o[i + 1]]) for i in range(len(o) - 1)] == [b"ACGTACGT", b"TTTT", b""]
    fq = tmp_path / "a.fq.gz"
    with gzip.open(fq, "wb") as f:
        f.write(b"@q1\nACGTN\n+\n@IIII\n@q2\nGG\nCC\n+q2\nII\nII\n")
    d, o = oracle.read_fastx(str(fq))
    assert [bytes(d[o[i]:o[i + 1]]) for i in range(len(o) - 1)] == [b"ACGTN", b"GGCC"]


# ---- SimkaMin: sketch + distance of the oracle against the reference's own golden matrices ----------------
# (tests/simkaMin/test_simkaMin.py:141-154 of the reference; truth_simkaMin_symetrical/k{21,31}__{0,100}-{100,1000}_n1,
#  the unfiltered half of its grid -- the -filter half uses an approximate Bloom filter)
def simkamin_example_sketches(golden_dir, k, max_reads, nb_kmers, sketch_fn):
    """the reads simkaMin selects for every sample of the example (<= max_reads per ';' group, 0 = all), sketched"""
    import numpy as np
    from helpers import reads_to_arrays, select_reads
    spec = oracle.parse_simka_input(os.path.join(golden_dir, "example", "simka_input.txt"))
    ids, hashes, counts = [], [], []
    for ident, groups in spec:
        reads = select_reads(groups, max_reads)
        seqs, off = reads_to_arrays(reads)
        h, c, _ = sketch_fn(seqs, off, k, nb_kmers)
        ids.append(ident)
        hashes.append(h)
        counts.append(c)
    return ids, np.stack(hashes), np.stack(counts)


def simkamin_csv(ids, mat):
    lines = [";" + ";".join(ids)]
    for i, ident in enumerate(ids):
        lines.append(ident + ";" + ";".join("%f" % float(v) for v in mat[i]))
    return "\n".join(lines) + "\n"


@pytest.mark.parametrize("k", [21, 31])
@pytest.mark.parametrize("max_reads", [0, 100])
@pytest.mark.parametrize("nb_kmers", [100, 1000])
def test_simkamin_oracle_reproduces_reference_truth(golden_dir, k, max_reads, nb_kmers):
    ids, hashes, counts = simkamin_example_sketches(golden_dir, k, max_reads, nb_kmers,
                                                    lambda s, o, kk, n: oracle.simkamin_sketch(s, o, kk, n, seed=100))
    jac, bc = oracle.simkamin_distance(hashes, counts)
    tdir = os.path.join(golden_dir, "truth_simkaMin", f"k{k}__{max_reads}-{nb_kmers}_n1")
    assert simkamin_csv(ids, jac) == open(os.path.join(tdir, "mat_presenceAbsence_jaccard.csv")).read()
    assert simkamin_csv(ids, bc) == open(os.path.join(tdir, "mat_abundance_braycurtis.csv")).read()


@pytest.mark.parametrize("k", [21, 31])
@pytest.mark.parametrize("max_reads", [0, 100])
@pytest.mark.parametrize("nb_kmers", [100, 1000])
def test_simkamin_oracle_filter_reproduces_reference_truth(golden_dir, k, max_reads, nb_kmers):
    """the -filter half of the reference's SimkaMin grid (truth_simkaMin_symetrical/k*_filter_*): the oracle remembers first
    occurrences exactly where the reference asks a 6.4e10-bit Bloom filter, and reproduces the golden matrices byte for byte"""
    ids, hashes, counts = simkamin_example_sketches(golden_dir, k, max_reads, nb_kmers,
                                                    lambda s, o, kk, n: oracle.simkamin_sketch(s, o, kk, n, seed=100, filter=True))
    jac, bc = oracle.simkamin_distance(hashes, counts)
    tdir = os.path.join(golden_dir, "truth_simkaMin", f"k{k}_filter_{max_reads}-{nb_kmers}_n1")
    assert simkamin_csv(ids, jac) == open(os.path.join(tdir, "mat_presenceAbsence_jaccard.csv")).read()
    assert simkamin_csv(ids, bc) == open(os.path.join(tdir, "mat_abundance_braycurtis.csv")).read()


def test_murmur3_known_answer():
    """MurmurHash3_x64_128 of 8 zero bytes with seed 0 is 0 for h1 only if the mix is wrong: pin two self-consistent facts
    instead -- different seeds and keys give different values, and the value is stable (regression pin)."""
    L = oracle.lib()
    a = L.orc_murmur3_x64_128_u64(0, 100)
    b = L.orc_murmur3_x64_128_u64(1, 100)
    c = L.orc_murmur3_x64_128_u64(0, 101)
    assert len({a, b, c}) == 3


def test_murmur3_against_the_reference_source():
    """the oracle's MurmurHash3 restatement == the reference's own MurmurHash3.cpp (compiled unmodified into oracle/_ref by
    oracle/build_ref.sh) on edge keys and 20000 random 64-bit keys, for the default seed and two others"""
    import numpy as np
    if oracle.build_ref() is None:
        pytest.skip("oracle/_ref/libmurmur_ref.so not built (no /root/reference here and no prebuilt file)")
    rng = np.random.default_rng(5)
    keys = np.concatenate([np.array([0, 1, 2**62 - 1, 2**63, 2**64 - 1], dtype=np.uint64), rng.integers(0, 2**64, 20000, dtype=np.uint64)])
    L = oracle.lib()
    for seed in (100, 0, 4294967295):
        ref = oracle.ref_murmur3_x64_128_u64(keys, seed)
        got = np.array([L.orc_murmur3_x64_128_u64(int(x), seed) for x in keys], dtype=np.uint64)
        np.testing.assert_array_equal(got, ref)


# GT/unit/src/kmer/TestKmer.cpp:443-464 (kmer_minimizer3, ModelMinimizer<ModelCanonical>, k=15, m=7): canonical k-mer -> minimizer
TESTKMER_MINIMIZER3_SEQ = b"ATGTCTGAAGTGACCTAACATTGCAGTGTGTT"
TESTKMER_MINIMIZER3 = [
    ("ATGTCTGAAGTGACC", "AAGTGAC"), ("AGGTCACTTCAGACA", "AAGTGAC"), ("TAGGTCACTTCAGAC", "AAGTGAC"), ("TCTGAAGTGACCTAA", "AAGTGAC"),
    ("CTGAAGTGACCTAAC", "AAGTGAC"), ("TGAAGTGACCTAACA", "AAGTGAC"), ("ATGTTAGGTCACTTC", "AAGTGAC"), ("AATGTTAGGTCACTT", "AATGTTA"),
    ("AGTGACCTAACATTG", "AACATTG"), ("GCAATGTTAGGTCAC", "AACATTG"), ("TGACCTAACATTGCA", "AACATTG"), ("CTGCAATGTTAGGTC", "AACATTG"),
    ("ACCTAACATTGCAGT", "AACATTG"), ("CACTGCAATGTTAGG", "AACATTG"), ("ACACTGCAATGTTAG", "AACATTG"), ("CACACTGCAATGTTA", "AACATTG"),
    ("AACATTGCAGTGTGT", "AACATTG"), ("AACACACTGCAATGT", "AACACAC"),
]


def kmer_string(v, k):
    return "".join("ACTG"[(int(v) >> (2 * (k - 1 - i))) & 3] for i in range(k))


def test_minimizer_known_answers_testkmer():
    """the reference's minimizer known-answer table: the canonical 15-mers of the sequence in order, and for each the minimizer
    (smallest canonical 7-mer without an inner "AA") -- pins orc_minimizer and, through test_minimizer_matches_oracle, the host model"""
    km, valid = oracle.canonical_kmers(TESTKMER_MINIMIZER3_SEQ, 15)
    assert valid.all() and len(km) == len(TESTKMER_MINIMIZER3)
    L = oracle.lib()
    for v, (kmer, mini) in zip(km.tolist(), TESTKMER_MINIMIZER3):
        assert kmer_string(v, 15) == kmer
        assert kmer_string(L.orc_minimizer(v, 15, 7), 7) == mini


def dsk_known_answers(golden_dir):
    import json
    d = json.load(open(os.path.join(golden_dir, "testdsk", "check1.json")))
    return [(d["sequences"][r["seqs"]], r["k"], r["abundance_min"], r["nb_solid"]) for r in d["rows"]]


def test_dsk_solid_counts_known_answers(golden_dir):
    """gatb-core's DSK known answers (test/unit/src/kmer/TestDSK.cpp:147-240, 33 rows "computed with the original minia"):
    number of distinct canonical k-mers with count >= abundance-min over a bank of literal sequences"""
    for seqs, k, amin, want in dsk_known_answers(golden_dir):
        run = oracle.Run.from_reads([[s.encode() for s in seqs]], k, abundance_min=amin, simple=False, complex_=False).count()
        km, ct = run.sample(0)
        assert len(km) == want == int(run.totals(0)[1]), (k, amin)
        assert (ct >= amin).all()
        run.close()


# (input fixture, k): truth produced by running the REFERENCE BINARIES themselves on the fixture (tools/make_ref_golden.py ->
# tests/golden/truth_ref): the k values the reference's own truth files do not cover.  k <= 15 is the MiniKC path
# (src/minikc/MiniKC.hpp:195-259: windows with non-ACGT characters are counted, and a read shorter than k replays the previous
# read's k-mers); example_n holds N runs, lower case, short reads, multi-line FASTA, a FASTQ.gz and a paired sample.
REF_TRUTH_CASES = [("example", 9), ("example", 15), ("example", 16), ("example_n", 12), ("example_n", 15), ("example_n", 21), ("example_n", 31),
                   # k > 31: Kmer<64> / LargeInt<2>, 128-bit k-mers
                   ("example", 32), ("example", 41), ("example", 63), ("example_n", 41),
                   # file-format corner cases (tools/make_ref_golden.py, make_example_fmt)
                   ("example_fmt", 15), ("example_fmt", 21), ("example_fmt", 31)]


def ref_truth_dir(golden_dir, inp, k, t):
    return os.path.join(golden_dir, "truth_ref", ("" if inp == "example" else inp + "_") + f"results_k{k}_t{t}")


@pytest.mark.parametrize("inp,k", REF_TRUTH_CASES)
@pytest.mark.parametrize("t", [0, 2])
def test_reference_run_truth_byte_identical(golden_dir, inp, k, t):
    import json
    run, ids = oracle.run_from_input_file(os.path.join(golden_dir, inp, "simka_input.txt"), k, abundance_min=t)
    run.merge()
    tdir = ref_truth_dir(golden_dir, inp, k, t)
    names = run.matrix_names()
    assert len(names) == 21
    for m, name in enumerate(names):
        assert run.csv(m, ids) == open(os.path.join(tdir, name + ".csv")).read(), name
    sync = json.load(open(os.path.join(tdir, "count_synchro.json")))
    tot = run.all_totals()
    for i, ident in enumerate(ids):
        assert [int(x) for x in tot[i]] == sync[ident], ident
