"""CPU tests of the product's host side: the C ABI loads and exports what the header declares, the
read packer, and the distance finaliser / CSV writer (checked against the oracle and the truth)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import simka_b200
from helpers import nasty_samples, oracle_run, synthetic_samples
from oracle import oracle
from simka_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "simka_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(simka\w*|simkamin\w*)\s*\(", header))
    declared -= {"simka_ctx"}
    assert len(declared) >= 20
    L = C.CDLL(str(_lib.LIB_PATH))
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/simka_b200.h but not exported"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert _lib.load().simka_b200_version().startswith(b"simka_b200")


def test_no_gpu_fails_loudly():
    """no CPU fallback: creating a context without a CUDA device is an error, not a silent slow path"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(simka_b200.SimkaError):
        simka_b200.SimkaGpu(2, 21)


def decode_windows(packed, bad, n_words, k):
    """numpy restatement of what K1 sees: forward k-mer + validity per start position"""
    bits = np.unpackbits(packed[:n_words + 1].byteswap().view(np.uint8))  # MSB-first stream
    codes = bits[0::2] * 2 + bits[1::2]
    badbits = np.unpackbits(bad[:n_words + 1].byteswap().view(np.uint8))
    npos = n_words * 32
    fwd = np.zeros(npos, np.uint64)
    for i in range(k):
        fwd = (fwd << np.uint64(2)) | codes[i:i + npos].astype(np.uint64)
    csum = np.concatenate([[0], np.cumsum(badbits)])
    valid = (csum[k:k + npos] - csum[:npos]) == 0
    return fwd, valid


@pytest.mark.parametrize("k", [5, 21, 31])
def test_pack_reads_layout(k):
    samples = nasty_samples()
    L = oracle.lib()
    for seqs, off in samples:
        packed, bad, nw = simka_b200.pack_reads(seqs, off, nthreads=3)
        assert packed[nw] == 0 and bad[nw] == 0xFFFFFFFF
        fwd, valid = decode_windows(packed, bad, nw, k)
        rc = np.array([L.orc_revcomp(int(x), k) for x in fwd[valid]], dtype=np.uint64)
        got = np.sort(np.minimum(fwd[valid], rc))
        exp = []
        raw = seqs.tobytes()
        for r in range(len(off) - 1):
            km, va = oracle.canonical_kmers(raw[int(off[r]):int(off[r + 1])], k)
            exp.append(km[va])
        exp = np.sort(np.concatenate(exp)) if exp else np.zeros(0, np.uint64)
        np.testing.assert_array_equal(got, exp)


PACK_DIFF_SCRIPT = r"""
import hashlib, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
import numpy as np
import simka_b200
from helpers import nasty_samples, synthetic_samples
h = hashlib.sha256()
rng = np.random.default_rng(1)
cases = nasty_samples() + synthetic_samples(2, 2000)
alpha = np.frombuffer(b"ACGTacgtNnX-\x00\xff", np.uint8)
for t in range(40):  # ragged reads (lengths 0..200, empty ones, odd characters), buffers that end right after the last base
    lens = rng.integers(0, 200, size=rng.integers(1, 400))
    off = np.zeros(len(lens) + 1, np.uint64); off[1:] = np.cumsum(lens)
    cases.append((alpha[rng.integers(0, len(alpha), int(off[-1]))], off))
for seqs, off in cases:
    for k in (31, 15):
        for nt in (1, 3):
            p, b, nw = simka_b200.pack_reads(seqs, off, kmer_size=k, nthreads=nt)
            h.update(np.asarray(p)[:nw + 1].tobytes()); h.update(np.asarray(b)[:nw + 1].tobytes())
print(h.hexdigest())
"""


def test_pack_reads_simd_equals_scalar():
    """simka_pack_reads packs 32 characters at a time with AVX2 + BMI2 where the CPU has them (run-time dispatch, simka_host.cpp);
    SIMKA_B200_PACK_SCALAR=1 keeps the portable loop: same bytes out of both on ragged / dirty inputs (the layout itself is
    pinned by test_pack_reads_layout, which runs on whichever path this machine takes)"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for scalar in (False, True):
        env = dict(os.environ)
        env.pop("SIMKA_B200_PACK_SCALAR", None)
        if scalar:
            env["SIMKA_B200_PACK_SCALAR"] = "1"
        r = subprocess.run([sys.executable, "-c", PACK_DIFF_SCRIPT, root], capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr
        outs.append(r.stdout.strip())
    assert outs[0] == outs[1] and len(outs[0]) == 64


def stats_from_oracle(run):
    st = simka_b200.SimkaStatistics(run.n, run.simple, run.complex_)
    st.totals[:] = run.all_totals()
    st.nb_distinct_kmers, st.nb_shared_kmers = run.nb_distinct_kmers, run.nb_shared_kmers
    for name in ("distinct_shared", "braycurtis_num", "nb_shared"):
        getattr(st, name)[...] = run.u64(name)
    if run.simple:
        for name in ("chord_ninj", "hellinger", "kulczynski"):
            getattr(st, name)[...] = run.u64(name)
    if run.complex_:
        for name in ("canberra", "whittaker"):
            getattr(st, name)[...] = run.u64(name)
        st.kullback_leibler[...] = run.f64("kl")
    return st


@pytest.mark.parametrize("k,t", [(21, 0), (21, 2), (31, 0), (31, 2)])
def test_host_finaliser_reproduces_truth_csv(golden_dir, k, t):
    """SimkaDistance closed forms + dumpMatrix of the product (host C++), fed with oracle accumulators"""
    run, ids = oracle.run_from_input_file(os.path.join(golden_dir, "example", "simka_input.txt"), k, abundance_min=t)
    run.merge()
    st = stats_from_oracle(run)
    assert st.matrix_names() == run.matrix_names()
    tdir = os.path.join(golden_dir, "truth", f"results_k{k}_t{t}")
    for m, name in enumerate(st.matrix_names()):
        np.testing.assert_array_equal(st.matrix(m), run.matrix(m), err_msg=name)
        path = os.path.join(tdir, name + ".csv")
        if os.path.exists(path):
            assert st.csv(m, ids) == open(path).read(), name


def test_output_matrix_files(tmp_path, golden_dir):
    import gzip
    run, ids = oracle.run_from_input_file(os.path.join(golden_dir, "example", "simka_input.txt"), 31, abundance_min=2)
    run.merge()
    st = stats_from_oracle(run)
    st.output_matrix(str(tmp_path / "out"), ids)
    files = sorted(os.listdir(tmp_path / "out"))
    assert len(files) == 21 and all(f.endswith(".csv.gz") for f in files)
    got = gzip.open(tmp_path / "out" / "mat_abundance_braycurtis.csv.gz", "rt").read()
    assert got == open(os.path.join(golden_dir, "truth", "results_k31_t2", "mat_abundance_braycurtis.csv")).read()


def test_finaliser_flag_subsets():
    samples = synthetic_samples(4, 300, n_genomes=4, genome_len=1500)
    for simple, complex_ in [(False, False), (True, False), (False, True)]:
        run = oracle_run(samples, 21, 2, simple=simple, complex_=complex_)
        st = stats_from_oracle(run)
        assert len(st.matrix_names()) == 15 + 3 * simple + 3 * complex_
        for m in range(len(st.matrix_names())):
            np.testing.assert_array_equal(st.matrix(m), run.matrix(m))


def test_statistics_iadd():
    a = simka_b200.SimkaStatistics(3, True, True)
    b = simka_b200.SimkaStatistics(3, True, True)
    a.nb_shared[:] = 5
    b.nb_shared[:] = 7
    a.nb_distinct_kmers, b.nb_distinct_kmers = 2, 3
    b.kullback_leibler[0, 1] = 0.25
    a += b
    assert (a.nb_shared == 12).all() and a.nb_distinct_kmers == 5 and a.kullback_leibler[0, 1] == 0.25


def test_pack_reads_minikc_rule():
    """k <= 15: only the read separators are bad (the reference's MiniKC counts windows with non-ACGT characters), and a read
    shorter than k is replaced by a copy of the previous long-enough read (its k-mer iterator keeps the previous read's k-mers:
    simka_minikc_replay, pinned on reference runs in tests/golden/truth_ref/example_n_*)"""
    from helpers import reads_to_arrays
    seqs, off = reads_to_arrays([b"ACGTNNACGT", b"acgtxACG"])
    p31, b31, nw = simka_b200.pack_reads(seqs, off, kmer_size=31)
    p9, b9, nw9 = simka_b200.pack_reads(seqs, off, kmer_size=8)   # both reads hold a k-mer: nothing is replayed
    assert nw == nw9 == 1
    np.testing.assert_array_equal(p31, p9)  # same codes, (c>>1)&3 for every character
    bits31 = format(int(b31[0]), "032b")
    bits9 = format(int(b9[0]), "032b")
    assert bits31[:20] == "00001100001" + "000010001"   # NN, separator, x, separator
    assert bits9[:20] == "00000000001" + "000000001"    # separators only
    assert bits31[20:] == bits9[20:] == "1" * 12         # padding
    # k = 9: the second read (8 bases) is shorter than k -> the first read again
    seqs2, off2 = simka_b200.api.minikc_replay(seqs, off, 9)
    assert bytes(seqs2) == b"ACGTNNACGT" * 2 and off2.tolist() == [0, 10, 20]
    # a short read with nothing before it stays empty; k > 15 never replays
    seqs3, off3 = simka_b200.api.minikc_replay(*reads_to_arrays([b"ACG", b"ACGTACGTAC", b"AC"]), 9)
    assert bytes(seqs3) == b"ACGTACGTAC" * 2 and off3.tolist() == [0, 0, 10, 20]
    s4, o4 = reads_to_arrays([b"ACG", b"ACGTACGTACGTACGTACGT"])
    s5, o5 = simka_b200.api.minikc_replay(s4, o4, 16)
    assert s5 is s4 and o5 is o4
