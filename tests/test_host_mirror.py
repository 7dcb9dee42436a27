"""CPU tests of the host-side mirror of the reference's input surface (simka_b200/csrc/host/*.hpp): the
Kmer<32>::Model* iteration, the minimizer, the Shannon filter, Bank::open / Iterator<Sequence>, and the
argument handling of the `simka` command line.  Checked against the oracle and the reference's own
known-answer vectors."""
import ctypes as C
import gzip
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle
from simka_b200 import _lib
from simka_b200.build import SIMKA_BIN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def host_kmers(seq: bytes, k: int, canonical=True):
    L = _lib.load()
    n = max(1, len(seq))
    out = np.zeros(n, np.uint64)
    valid = np.zeros(n, np.uint8)
    m = L.simka_host_kmers(seq, len(seq), k, int(canonical), out.ctypes.data_as(_lib.u64p), valid.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out[:m], valid[:m].astype(bool)


def test_testkmer_known_answers():
    """GT/unit/src/kmer/TestKmer.cpp:141-147,178-191"""
    seq = b"CATTGATAGTGG"
    d, v = host_kmers(seq, 3, canonical=False)
    assert d.tolist() == [18, 10, 43, 44, 50, 8, 35, 14, 59, 47] and v.all()
    c, v = host_kmers(seq, 3, canonical=True)
    assert c.tolist() == [11, 2, 16, 36, 9, 8, 24, 6, 17, 20] and v.all()


def test_testkmer_badchar_known_answers():
    """GT/unit/src/kmer/TestKmer.cpp:509-569: validity and value (bad nucleotide -> G) of every window, ModelDirect k = 11"""
    from test_oracle_golden import BADCHAR_SEQ, BADCHAR_VALID, kmer_to_string
    d, v = host_kmers(BADCHAR_SEQ, 11, canonical=False)
    assert v.tolist() == BADCHAR_VALID
    assert [kmer_to_string(x, 11) for x in d] == [BADCHAR_SEQ[i:i + 11].decode().replace("N", "G") for i in range(32)]


@pytest.mark.parametrize("k", [3, 15, 21, 31])
def test_model_matches_oracle(k):
    rng = np.random.default_rng(k)
    alpha = np.frombuffer(b"ACGTacgtNnX-", np.uint8)
    p = np.array([20, 20, 20, 20, 3, 3, 3, 3, 1, 1, 1, 1], float)
    for n in [0, k - 1, k, k + 1, 200, 1000]:
        seq = alpha[rng.choice(len(alpha), size=n, p=p / p.sum())].tobytes()
        for canonical in (True, False):
            a, av = host_kmers(seq, k, canonical)
            if n < k:
                assert len(a) == 0
                continue
            b, bv = (oracle.canonical_kmers if canonical else oracle.direct_kmers)(seq, k)
            np.testing.assert_array_equal(av, bv.astype(bool))
            np.testing.assert_array_equal(a[av], b[av])  # values of invalid windows are unspecified in the reference too


def test_minimizer_matches_oracle():
    L, O = _lib.load(), oracle.lib()
    rng = np.random.default_rng(5)
    for k, m in [(31, 7), (21, 7), (15, 5), (11, 8)]:
        for v in rng.integers(0, 1 << (2 * k), size=300, dtype=np.uint64).tolist() + [0, (1 << (2 * k)) - 1]:
            assert L.simka_host_minimizer(v, k, m) == O.orc_minimizer(v, k, m)


def test_shannon_index():
    L = _lib.load()
    assert L.simka_host_shannon_index(b"AAAAAAAA", 8) == 0.0
    assert abs(L.simka_host_shannon_index(b"ACGTACGT", 8) - 2.0) < 1e-6
    assert abs(L.simka_host_shannon_index(b"AACC", 4) - 1.0) < 1e-6
    # lower case counts as 'A', as in the reference's table (src/core/SimkaCommons.hpp:400-411)
    assert L.simka_host_shannon_index(b"acgt", 4) == 0.0


def read_bank(uri):
    L = _lib.load()
    ps, po = C.c_void_p(), C.c_void_p()
    n = L.simka_host_read_bank(os.fsencode(uri), C.byref(ps), C.byref(po))
    assert n >= 0, uri
    off = np.ctypeslib.as_array((C.c_uint64 * (n + 1)).from_address(po.value)).copy()
    total = int(off[-1])
    seqs = np.ctypeslib.as_array((C.c_uint8 * max(total, 1)).from_address(ps.value))[:total].copy()
    L.simka_host_free(ps)
    L.simka_host_free(po)
    return seqs, off


def test_bank_reads_like_the_oracle(tmp_path, golden_dir):
    ex = os.path.join(golden_dir, "example")
    files = [os.path.join(ex, f) for f in ("A.fasta", "B.fasta", "C.fasta", "D_paired_1.fasta", "D_paired_2.fasta")]
    for f in files:
        s, o = read_bank(f)
        es, eo = oracle.read_fastx(f)
        np.testing.assert_array_equal(s, es)
        np.testing.assert_array_equal(o, eo)
    # gz, multi-line fasta, fastq (quality line starting with '@'), CRLF, a last record without newline
    multi = tmp_path / "m.fa"
    multi.write_bytes(b">r1 c\nACGT\nAC\n\nGG\n>r2\n\n>r3\nTTTT\r\nAA\r\n>r4\nNNNN")
    fq = tmp_path / "q.fastq"
    fq.write_bytes(b"@q1\nACGTAC\n+\n@IIII+\n@q2\nGGGTTT\n+q2\nIIIIII\n")
    gz = tmp_path / "g.fa.gz"
    with gzip.open(gz, "wb") as f:
        f.write(open(files[0], "rb").read())
    for f in (multi, fq, gz):
        s, o = read_bank(str(f))
        es, eo = oracle.read_fastx(str(f))
        np.testing.assert_array_equal(s, es)
        np.testing.assert_array_equal(o, eo)
    s, o = read_bank(str(fq))
    assert bytes(s) == b"ACGTACGGGTTT" and o.tolist() == [0, 6, 12]
    # album: file of file names, relative to the album
    album = tmp_path / "album.txt"
    album.write_text("m.fa\n" + str(fq) + "\n")
    s, o = read_bank(str(album))
    s1, o1 = read_bank(str(multi))
    s2, o2 = read_bank(str(fq))
    assert bytes(s) == bytes(s1) + bytes(s2) and len(o) == len(o1) + len(o2) - 1


def test_testbank_known_answers(golden_dir):
    """The sample banks of the reference's own bank tests (GT/unit/src/bank/TestBank.cpp:98-250, data files of gatb-core test/db
    copied to tests/golden/bankdb): bank_checkSample1 -- 20 sequences of 20 letters, sequence i = "ARNDCQEGHILKMFPSTWYV" rotated
    left by i, plain and gzip; bank_checkSample2 -- 20 records with a comment and no data; bank_checkSample3 -- 7 non-empty
    fastq records, plain and gzip.  Host reader and oracle reader."""
    db = os.path.join(golden_dir, "bankdb")
    text = b"ARNDCQEGHILKMFPSTWYV"
    for reader in (read_bank, oracle.read_fastx):
        for name in ("sample1.fa", "sample1.fa.gz"):
            s, o = reader(os.path.join(db, name))
            assert o.tolist() == [20 * i for i in range(21)], name
            for i in range(20):
                assert bytes(s[20 * i:20 * i + 20]) == text[i:] + text[:i]
        s, o = reader(os.path.join(db, "sample2.fa"))
        assert o.tolist() == [0] * 21 and len(s) == 0
        for name in ("sample.fastq", "sample.fastq.gz"):
            s, o = reader(os.path.join(db, name))
            assert len(o) == 8 and (np.diff(o.astype(np.int64)) > 0).all(), name
    a, ao = read_bank(os.path.join(db, "sample.fastq"))
    b, bo = read_bank(os.path.join(db, "sample.fastq.gz"))
    assert bytes(a) == bytes(b) and ao.tolist() == bo.tolist()


def test_cli_argument_handling(tmp_path):
    """same conventions as the reference's main: a failing run prints EXCEPTION: <msg> on stdout and ends with EXIT_FAILURE
    (src/SimkaPotara.cpp:147-163); a mistake on the command line prints ERROR: lines and the option list and ends with
    EXIT_SUCCESS (GATB Tool / OptionFailure) -- checked against the reference binary where it is built"""
    exe = str(SIMKA_BIN)
    assert os.path.exists(exe)
    r = subprocess.run([exe, "-help"], capture_output=True, text=True)
    assert r.returncode == 0 and "-abundance-min" in r.stdout and "-complex-dist" in r.stdout
    ref = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "bin", "simka")
    for args, errors in ((["-in", "x"], ["ERROR: Option '-out-tmp' is mandatory"]),
                         (["-in", "x", "-out-tmp", str(tmp_path), "-bogus"], ["ERROR: Unknown parameter '-bogus'"]),
                         ([], ["ERROR: Option '-in' is mandatory", "ERROR: Option '-out-tmp' is mandatory"]),
                         (["-out-tmp", str(tmp_path), "-in"], ["ERROR: Too few arguments for the -in option..."])):
        r = subprocess.run([exe] + args, capture_output=True, text=True)
        assert r.returncode == 0 and "-abundance-min" in r.stdout
        for e in errors:
            assert e in r.stdout.splitlines(), (args, r.stdout[:300])
        if os.access(ref, os.X_OK):
            rr = subprocess.run([ref] + args, capture_output=True, text=True)
            assert rr.returncode == 0
            for e in errors:
                assert e in rr.stdout.splitlines(), (args, rr.stdout[:300])
    r = subprocess.run([exe, "-in", "/nonexistent", "-out-tmp", str(tmp_path), "-out", str(tmp_path / "o")], capture_output=True, text=True)
    assert r.returncode == 1 and "does not exist" in r.stdout
    r = subprocess.run([exe, "-in", "x", "-out-tmp", str(tmp_path), "-max-memory", "100"], capture_output=True, text=True)
    assert r.returncode == 1 and "memory" in r.stdout
    r = subprocess.run([exe, "-in", "x", "-out-tmp", str(tmp_path), "-kmer-size", "64"], capture_output=True, text=True)
    assert r.returncode == 1 and "kmer-size" in r.stdout


def test_stage_binaries_argument_handling(tmp_path):
    bindir = os.path.dirname(str(SIMKA_BIN))
    for exe in ("simkaCount", "simkaMerge", "simkaCountProcess"):
        assert os.path.exists(os.path.join(bindir, exe))
    r = subprocess.run([os.path.join(bindir, "simkaCount"), "-kmer-size", "31"], capture_output=True, text=True)
    assert r.returncode == 1 and r.stdout.startswith("EXCEPTION: ") and "mandatory" in r.stdout
    r = subprocess.run([os.path.join(bindir, "simkaMerge"), "-kmer-size", "31", "-out-tmp-simka", str(tmp_path), "-partition-id", "0"],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "datasetIds" in r.stdout
    # the retry wrapper: gives up after 5 failures, succeeds at once on exit 0
    marker = tmp_path / "n"
    script = tmp_path / "fail.sh"
    script.write_text(f"#!/bin/sh\necho x >> {marker}\nexit 1\n")
    script.chmod(0o755)
    r = subprocess.run([os.path.join(bindir, "simkaCountProcess"), str(script)])
    assert r.returncode == 1 and len(marker.read_text().split()) == 5
    r = subprocess.run([os.path.join(bindir, "simkaCountProcess"), "true"])
    assert r.returncode == 0


def _write_sketch_file(path, k, s, seed, ids, hashes, counts):
    """the reference's sketch file (SimkaMinCommons.hpp:28,42-55,80-93): 13-byte header, n x s 16-byte entries, ids"""
    import struct
    with open(path, "wb") as f:
        f.write(struct.pack("<B", k) + struct.pack("<III", s, seed, len(ids)))
        rec = np.zeros((len(ids), s), dtype=[("h", "<u8"), ("c", "<u4"), ("p", "<u4")])
        rec["h"], rec["c"] = hashes, counts
        f.write(rec.tobytes())
        for i in ids:
            f.write(bytes([len(i)]) + i.encode())


def test_simkamincore_argument_handling_and_export(tmp_path):
    """simkaMinCore: error conventions of the reference's main (src/simkaMin/SimkaMin.cpp:30-110), `distance` refuses
    sketches of different k or seed before any device work (SimkaMinDistance.hpp:990-998), `export` turns the float32
    .bin matrices into csv.gz exactly as writeMatrixASCII (SimkaMinDistanceMatrixExporter.hpp:384-404)"""
    import gzip
    exe = os.path.join(os.path.dirname(str(SIMKA_BIN)), "simkaMinCore")
    assert os.path.exists(exe)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "sketch" in r.stdout
    r = subprocess.run([exe, "bogus"], capture_output=True, text=True)
    assert r.returncode == 1 and r.stdout.startswith("EXCEPTION: ")
    r = subprocess.run([exe, "sketch", "-out", str(tmp_path / "s.bin")], capture_output=True, text=True)
    assert r.returncode == 1 and "mandatory" in r.stdout
    r = subprocess.run([exe, "sketch", "-in", "/nonexistent/input.txt", "-filter"], capture_output=True, text=True)
    assert r.returncode == 1 and r.stdout.startswith("EXCEPTION: ")
    rng = np.random.default_rng(3)
    ids = ["A", "B2", "sample_C"]
    h = np.sort(rng.integers(1, 2**63, (3, 8), dtype=np.uint64), axis=1)
    c = rng.integers(1, 9, (3, 8), dtype=np.uint32)
    _write_sketch_file(tmp_path / "a.bin", 21, 8, 100, ids, h, c)
    _write_sketch_file(tmp_path / "b.bin", 31, 8, 100, ids, h, c)
    _write_sketch_file(tmp_path / "c.bin", 21, 8, 7, ids, h, c)
    r = subprocess.run([exe, "distance", "-in1", str(tmp_path / "a.bin"), "-in2", str(tmp_path / "b.bin"), "-out", str(tmp_path / "d")],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "different kmer sizes" in r.stderr
    r = subprocess.run([exe, "distance", "-in1", str(tmp_path / "a.bin"), "-in2", str(tmp_path / "c.bin"), "-out", str(tmp_path / "d")],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "different seeds" in r.stderr
    # export
    d = tmp_path / "dist"
    d.mkdir()
    m = rng.random((3, 3)).astype(np.float32)
    m.tofile(d / "mat_presenceAbsence_jaccard.bin")
    (1 - m).astype(np.float32).tofile(d / "mat_abundance_braycurtis.bin")
    r = subprocess.run([exe, "export", "-in", str(d), "-in1", str(tmp_path / "a.bin"), "-in2", str(tmp_path / "a.bin"), "-out", str(tmp_path / "res")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    want = ";" + ";".join(ids) + "\n" + "".join(ids[i] + ";" + ";".join("%f" % float(v) for v in m[i]) + "\n" for i in range(3))
    assert gzip.open(tmp_path / "res" / "mat_presenceAbsence_jaccard.csv.gz", "rt").read() == want
    assert (tmp_path / "res" / "mat_abundance_braycurtis.csv.gz").exists()


def test_simkamincore_append_and_info(tmp_path):
    """append (src/simkaMin/SimkaMinAppend.hpp:60-160): -in2 is appended to -in1 in place, header count updated, ids of both
    files at the end; refuses different k / sketch size / seed.  info (SimkaMinInfos.hpp:58-100) lists header + ids."""
    exe = os.path.join(os.path.dirname(str(SIMKA_BIN)), "simkaMinCore")
    rng = np.random.default_rng(11)
    def sk(n, s=6):
        return np.sort(rng.integers(1, 2**63, (n, s), dtype=np.uint64), axis=1), rng.integers(1, 9, (n, s), dtype=np.uint32)
    h1, c1 = sk(2)
    h2, c2 = sk(3)
    a, b, want = tmp_path / "a.bin", tmp_path / "b.bin", tmp_path / "want.bin"
    _write_sketch_file(a, 21, 6, 100, ["A", "Bb"], h1, c1)
    _write_sketch_file(b, 21, 6, 100, ["C", "", "Eee"], h2, c2)
    _write_sketch_file(want, 21, 6, 100, ["A", "Bb", "C", "", "Eee"], np.concatenate([h1, h2]), np.concatenate([c1, c2]))
    r = subprocess.run([exe, "append", "-in1", str(a), "-in2", str(b)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert a.read_bytes() == want.read_bytes()
    r = subprocess.run([exe, "info", "-in", str(a)], capture_output=True, text=True)
    assert r.returncode == 0
    assert r.stdout == f"Sketch info: {a}\n\nk-mer size  : 21\nSketch size : 6\nSeed        : 100\n\nNb Datasets: 5\nA\nBb\nC\n\nEee\n\n"
    for k, s, seed, msg in ((31, 6, 100, "kmer sizes"), (21, 7, 100, "sketch sizes"), (21, 6, 5, "seeds")):
        hh, cc = sk(1, s)
        _write_sketch_file(tmp_path / "c.bin", k, s, seed, ["X"], hh, cc)
        r = subprocess.run([exe, "append", "-in1", str(a), "-in2", str(tmp_path / "c.bin")], capture_output=True, text=True)
        assert r.returncode == 1 and msg in r.stderr
    assert a.read_bytes() == want.read_bytes()
    r = subprocess.run([exe, "info", "-in", str(tmp_path / "missing.bin")], capture_output=True, text=True)
    assert r.returncode == 1 and "does not exist" in r.stderr


def test_simkamincore_matrix_update(tmp_path):
    """matrix-update (hidden sub-command driven by simkaMin_update.py:149-152; SimkaMinDistanceMatrixMerger.hpp:87-127,
    SimkaMinDistanceMatrixExporter.hpp:44-121): every mat_*.bin of -in grows from E x E to (E+N) x (E+N) with the blocks of
    existingVsNew/ (E x N, also transposed into the bottom-left corner) and newVsNew/ (N x N); .temp files are skipped."""
    exe = os.path.join(os.path.dirname(str(SIMKA_BIN)), "simkaMinCore")
    rng = np.random.default_rng(5)
    E, N = 3, 2
    def sk(n, s=4):
        return np.sort(rng.integers(1, 2**63, (n, s), dtype=np.uint64), axis=1), rng.integers(1, 9, (n, s), dtype=np.uint32)
    _write_sketch_file(tmp_path / "old.bin", 21, 4, 100, ["a", "b", "c"], *sk(E))
    _write_sketch_file(tmp_path / "new.bin", 21, 4, 100, ["d", "e"], *sk(N))
    d = tmp_path / "dist"
    (d / "existingVsNew").mkdir(parents=True)
    (d / "newVsNew").mkdir()
    want = {}
    for name in ("mat_presenceAbsence_jaccard.bin", "mat_abundance_braycurtis.bin"):
        ee, en, nn = (rng.random(sh).astype(np.float32) for sh in ((E, E), (E, N), (N, N)))
        ee.tofile(d / name)
        en.tofile(d / "existingVsNew" / name)
        nn.tofile(d / "newVsNew" / name)
        want[name] = np.block([[ee, en], [en.T, nn]])
    (d / "mat_old.bin.temp").write_bytes(b"left alone")
    (d / "notes.txt").write_text("left alone")
    r = subprocess.run([exe, "matrix-update", "-in", str(d), "-in1", str(tmp_path / "old.bin"), "-in2", str(tmp_path / "new.bin")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert sorted(r.stdout.split()) == sorted(want)
    for name, m in want.items():
        got = np.fromfile(d / name, dtype=np.float32).reshape(E + N, E + N)
        assert np.array_equal(got, m)
        assert not (d / (name + ".temp")).exists()
    assert (d / "mat_old.bin.temp").read_bytes() == b"left alone"
    r = subprocess.run([exe, "matrix-update", "-in", str(d), "-in1", str(tmp_path / "old.bin")], capture_output=True, text=True)
    assert r.returncode == 1 and "mandatory" in r.stdout


def test_minimizer_known_answers_testkmer():
    """GT/unit/src/kmer/TestKmer.cpp:443-464 on the host model (ModelMinimizer<ModelCanonical>::getMinimizerValue)"""
    from test_oracle_golden import TESTKMER_MINIMIZER3, TESTKMER_MINIMIZER3_SEQ, kmer_string
    L = _lib.load()
    km, valid = host_kmers(TESTKMER_MINIMIZER3_SEQ, 15)
    assert valid.all() and len(km) == len(TESTKMER_MINIMIZER3)
    for v, (kmer, mini) in zip(km.tolist(), TESTKMER_MINIMIZER3):
        assert kmer_string(v, 15) == kmer
        assert kmer_string(L.simka_host_minimizer(v, 15, 7), 7) == mini


def test_fastx_parser_randomised_against_the_oracle(tmp_path):
    """FastxParser (raw read / zlib, memchr line handling) against the oracle's reader on generated FASTA / FASTQ files: multi-line
    records, CRLF, blank lines, long headers, '+' lines with text, quality lines that begin with '@' or '>', multi-line quality,
    missing final newline, records longer than the 2 MB parse buffer, gz"""
    rng = np.random.default_rng(21)
    alpha = np.frombuffer(b"ACGTacgtN", np.uint8)

    def seq(n):
        return alpha[rng.integers(0, len(alpha), n)].tobytes()

    def wrap(s, width, eol):
        if width <= 0 or len(s) == 0:
            return s + eol
        return b"".join(s[i:i + width] + eol for i in range(0, len(s), width))

    for case in range(12):
        eol = b"\r\n" if case % 3 == 1 else b"\n"
        fastq = case % 2 == 1
        recs = []
        for r in range(int(rng.integers(1, 40))):
            n = int(rng.integers(0, 400)) if case != 10 else int(rng.integers(2_000_000, 2_300_000))
            s = seq(n)
            hdr = (b"@" if fastq else b">") + b"r%d " % r + b"x" * int(rng.integers(0, 120))
            width = int(rng.integers(0, 4)) * 37  # 0 = one line
            if fastq:
                q = bytes(rng.integers(33, 74, n, dtype=np.uint8))
                if n and rng.random() < 0.3:
                    q = (b"@" if rng.random() < 0.5 else b">") + q[1:]
                recs.append(hdr + eol + wrap(s, width, eol) + b"+" + (b"r%d" % r if rng.random() < 0.5 else b"") + eol + wrap(q, width, eol))
            else:
                recs.append(hdr + eol + wrap(s, width, eol) + (eol if rng.random() < 0.2 else b""))
            if case == 10 and r >= 1:
                break
        data = b"".join(recs)
        if case % 4 == 3 and data.endswith(eol):
            data = data[:-len(eol)]  # no final newline
        path = tmp_path / (f"c{case}." + ("fq" if fastq else "fa") + (".gz" if case % 5 == 4 else ""))
        if case % 5 == 4:
            with gzip.open(path, "wb") as f:
                f.write(data)
        else:
            path.write_bytes(data)
        s, o = read_bank(str(path))
        es, eo = oracle.read_fastx(str(path))
        np.testing.assert_array_equal(o, eo, err_msg=str(path))
        np.testing.assert_array_equal(s, es, err_msg=str(path))


def test_estimate_nb_items_small_files(golden_dir):
    """IBank::estimateNbItems (BankFasta.cpp:728-775): exact for a plain file below the threshold, 4 x compressed size / uncompressed
    size x the reads for a gz file -- 34 for the 100 reads of example_n/S3.fastq.gz, the value behind the reference's own
    `-max-reads 0` run (tests/golden/truth_ref/example_n_results_k*_maxreads0: limit 75 = (34 + 116) / 2)"""
    L = _lib.load()
    ex = os.path.join(golden_dir, "example_n")
    want = {"S1.fasta": 123, "S2.fasta": 120, "S3.fastq.gz": 34, "S4_1.fasta": 60, "S4_2.fasta": 60}
    for f, n in want.items():
        assert L.simka_host_estimate_nb_items(os.fsencode(os.path.join(ex, f))) == n, f
    assert L.simka_host_estimate_nb_items(os.fsencode(os.path.join(ex, "no_such_file"))) == -1


def test_estimate_nb_items_matches_reference_binary_beyond_the_threshold(tmp_path):
    """files with more than 50000 sequences (the estimate stops there and extrapolates from the reader's position, which advances
    in buffers of 256 KB), plain / gz / multi-line / FASTQ, a paired and a concatenated sample: the numbers the REFERENCE `simka
    -max-reads 0` prints (smallest, largest, mean, limit) against the ones derived from simka_host_estimate_nb_items"""
    import gzip
    import re
    import subprocess
    ref = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "bin", "simka")
    if not os.access(ref, os.X_OK):
        pytest.skip("oracle/_ref/bin not built")
    rng = np.random.default_rng(5)
    alpha = np.frombuffer(b"ACGT", np.uint8)

    def write(name, n, L_, fastq=False, gz=False, multiline=False):
        lens = L_ + rng.integers(0, 30, n)
        bases = alpha[rng.integers(0, 4, int(lens.sum()))].tobytes()
        out, p = [], 0
        for i in range(n):
            r = bases[p:p + int(lens[i])]
            p += int(lens[i])
            if fastq:
                out.append(b"@q%d\n" % i + r + b"\n+\n" + b"I" * len(r) + b"\n")
            elif multiline:
                out.append(b">r%d c\n" % i + r[:40] + b"\n" + r[40:] + b"\n")
            else:
                out.append(b">r%d\n" % i + r + b"\n")
        with (gzip.open if gz else open)(tmp_path / name, "wb") as f:
            f.write(b"".join(out))

    write("A.fasta", 130000, 60)
    write("B.fastq.gz", 90000, 70, fastq=True, gz=True)
    write("C.fasta", 70000, 80, multiline=True)
    write("D.fastq", 52000, 50, fastq=True)
    write("E.fasta.gz", 3000, 90, gz=True)
    samples = {"A": [["A.fasta"]], "B": [["B.fastq.gz"]], "C": [["C.fasta"], ["D.fastq"]], "D": [["D.fastq"]], "E": [["E.fasta.gz", "A.fasta"]]}
    (tmp_path / "in.txt").write_text("".join(f"{k}: " + " ; ".join(" , ".join(g) for g in v) + "\n" for k, v in samples.items()))
    # the reference prints its estimates before it starts counting; the run itself is not needed (and is stopped: own process group)
    import signal
    import time
    log = open(tmp_path / "ref.log", "w")
    proc = subprocess.Popen([ref, "-in", str(tmp_path / "in.txt"), "-out", str(tmp_path / "out"), "-out-tmp", str(tmp_path / "tmp"), "-kmer-size", "21",
                             "-max-reads", "0", "-nb-cores", "2", "-max-memory", "4000"], stdout=log, stderr=subprocess.STDOUT, cwd=tmp_path,
                            start_new_session=True)
    text = ""
    try:
        for _ in range(600):
            time.sleep(0.1)
            text = open(tmp_path / "ref.log").read()
            if "Reads per sample used up to:" in text or proc.poll() is not None:
                break
    finally:
        try:
            os.killpg(proc.pid, signal.SIGKILL)
        except ProcessLookupError:
            pass
        proc.wait()
        log.close()
    assert "Reads per sample used up to:" in text, text[-1500:]
    got = {k: int(re.search(pat + r"\s*(\d+)", text).group(1)) for k, pat in
           (("min", "Smaller sample contains:"), ("max", "Larger sample contains:"), ("mean", "mean of:"), ("limit", "Reads per sample used up to:"))}
    L = _lib.load()
    per = []
    for k, groups in samples.items():
        n = sum(L.simka_host_estimate_nb_items(os.fsencode(str(tmp_path / f))) for g in groups for f in g)
        per.append(n // len(groups))
    mine = {"min": min(per), "max": max(per), "mean": sum(per) // len(per), "limit": (min(per) + sum(per) // len(per)) // 2}
    assert mine == got, (per, text[-1500:])


def test_repartitor_from_reference_config_h5(tmp_path, golden_dir):
    """host/repartitor.hpp: the minimal HDF5 reader + Repartitor table on the config.h5 of a reference run (tests/golden/config_h5,
    tools/make_ref_golden.py): every sampled solid k-mer goes to the partition the reference's simkaCount wrote it to; damaged or
    foreign files are refused"""
    import json
    L = _lib.load()
    fx = os.path.join(golden_dir, "config_h5")
    want = json.load(open(os.path.join(fx, "kmer_partition.json")))
    km = np.array([p[0] for p in want["kmer_partition"]], np.uint64)
    part = np.array([p[1] for p in want["kmer_partition"]], np.uint16)
    out = np.zeros(len(km), np.uint16)
    nb = L.simka_host_repartition(os.fsencode(os.path.join(fx, "config.h5")), km.ctypes.data_as(C.c_void_p), len(km), want["kmer_size"], out.ctypes.data_as(C.c_void_p))
    assert nb == want["nb_partitions"] and len(km) > 1000 and len(set(part.tolist())) == nb
    np.testing.assert_array_equal(out, part)
    raw = open(os.path.join(fx, "config.h5"), "rb").read()
    for name, data in (("truncated.h5", raw[:3000]), ("not_hdf5.h5", b"x" * 5000), ("empty.h5", b""), ("bad_magic.h5", raw.replace(b"\x78\x56\x34\x12", b"\x00\x00\x00\x00"))):
        (tmp_path / name).write_bytes(data)
        assert L.simka_host_repartition(os.fsencode(str(tmp_path / name)), km.ctypes.data_as(C.c_void_p), 4, 21, out.ctypes.data_as(C.c_void_p)) == -1, name
    assert L.simka_host_repartition(os.fsencode(str(tmp_path / "missing.h5")), km.ctypes.data_as(C.c_void_p), 4, 21, out.ctypes.data_as(C.c_void_p)) == -1
