"""The SimkaMin wrapper pipelines (simka_b200/simkamin_pipeline.py = the reference's simkaMin/simkaMin.py and
simkaMin/simkaMin_update.py).  CPU part: the sequence of simkaMinCore sub-commands, with a recording stand-in for the
binary.  GPU part: the real binary end to end against the reference's golden SimkaMin matrices."""
import gzip
import json
import os
import stat
import struct
import subprocess
import sys

import numpy as np
import pytest

from simka_b200 import simkamin_pipeline as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPTS = os.path.join(ROOT, "simka_b200", "scripts")

STUB = r'''#!{py}
import json, os, struct, sys
log = os.environ["STUB_LOG"]
args = sys.argv[1:]
open(log, "a").write(json.dumps(args) + "\n")
kv = {{args[i]: (args[i + 1] if i + 1 < len(args) else "") for i in range(1, len(args))}}
if args[0] == "sketch":      # a header is all the pipeline reads back
    n = int(os.environ.get("STUB_N", "5"))
    open(kv["-out"], "wb").write(struct.pack("<BIII", int(kv["-kmer-size"]), int(kv["-nb-kmers"]), int(kv["-seed"]), n))
if args[0] == "distance" and os.environ.get("STUB_FAIL") == "distance":
    sys.exit(3)
'''


@pytest.fixture
def stub_core(tmp_path, monkeypatch):
    exe = tmp_path / "simkaMinCore"
    exe.write_text(STUB.format(py=sys.executable))
    exe.chmod(exe.stat().st_mode | stat.S_IXUSR)
    log = tmp_path / "calls.jsonl"
    monkeypatch.setenv("STUB_LOG", str(log))
    return str(exe), log


def calls(log):
    return [json.loads(l) for l in open(log)]


def test_distance_blocks_order():
    """simkaMin/simkaMin.py:176-190: ceil(n/100) steps, blocks (i, j >= i) row by row, ragged last block"""
    assert list(sp.distance_blocks(5)) == [(0, 0, 5, 5)]
    assert list(sp.distance_blocks(250)) == [(0, 0, 100, 100), (0, 100, 100, 100), (0, 200, 100, 50),
                                             (100, 100, 100, 100), (100, 200, 100, 50), (200, 200, 50, 50)]
    assert list(sp.distance_blocks(200, 100)) == [(0, 0, 100, 100), (0, 100, 100, 100), (100, 100, 100, 100)]
    assert list(sp.distance_blocks(0)) == []


def test_simkamin_pipeline_commands(tmp_path, stub_core, monkeypatch):
    exe, log = stub_core
    monkeypatch.setenv("STUB_N", "230")
    out = tmp_path / "res"
    r = subprocess.run([sys.executable, os.path.join(SCRIPTS, "simkaMin.py"), "-in", "input.txt", "-out", str(out), "-bin", exe,
                        "-kmer-size", "31", "-nb-kmers", "5000", "-filter", "-max-reads", "77"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    c = calls(log)
    sk = str(out / "simkamin" / "sketch" / "sketch.bin")
    dd = str(out / "simkamin" / "distance")
    assert c[0] == ["sketch", "-in", "input.txt", "-out", sk, "-seed", "100", "-kmer-size", "31", "-nb-kmers", "5000", "-filter",
                    "-max-reads", "77", "-min-read-size", "0", "-min-shannon-index", "0", "-nb-cores", "0", "-max-memory", "8000"]
    dist = [x for x in c if x[0] == "distance"]
    assert [(x[x.index("-start-i") + 1], x[x.index("-start-j") + 1], x[x.index("-n-i") + 1], x[x.index("-n-j") + 1]) for x in dist] == \
        [("0", "0", "100", "100"), ("0", "100", "100", "100"), ("0", "200", "100", "30"), ("100", "100", "100", "100"),
         ("100", "200", "100", "30"), ("200", "200", "30", "30")]
    assert all(x[1:7] == ["-in1", sk, "-in2", sk, "-out", dd] for x in dist)
    assert c[-1] == ["export", "-in", dd, "-in1", sk, "-in2", sk, "-out", str(out), "-nb-cores", "0"]
    assert len(c) == 1 + 6 + 1
    # the binary matrices are created empty before the (possibly concurrent) distance jobs; one log per block
    assert os.path.getsize(os.path.join(dd, "mat_presenceAbsence_jaccard.bin")) == 0
    assert sorted(os.listdir(out / "simkamin" / "logs")) == ["log_distance_0-0", "log_distance_0-1", "log_distance_0-2",
                                                              "log_distance_1-1", "log_distance_1-2", "log_distance_2-2"]
    assert "Result dir: " + str(out) in r.stdout


def test_simkamin_pipeline_stops_on_failure(tmp_path, stub_core, monkeypatch):
    exe, log = stub_core
    monkeypatch.setenv("STUB_FAIL", "distance")
    r = subprocess.run([sys.executable, os.path.join(SCRIPTS, "simkaMin.py"), "-in", "input.txt", "-out", str(tmp_path / "res"), "-bin", exe],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "ERROR" in r.stdout
    assert [x[0] for x in calls(log)] == ["sketch", "distance"]
    r = subprocess.run([sys.executable, os.path.join(SCRIPTS, "simkaMin.py"), "-in", "input.txt", "-bin", str(tmp_path / "nope")],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "not found or not executable" in r.stdout
    r = subprocess.run([sys.executable, os.path.join(SCRIPTS, "simkaMin.py"), "-out", "x"], capture_output=True, text=True)
    assert r.returncode == 2  # -in is required


def test_simkamin_update_commands(tmp_path, stub_core):
    exe, log = stub_core
    ex = tmp_path / "res" / "simkamin"
    (ex / "sketch").mkdir(parents=True)
    (ex / "distance").mkdir()
    (ex / "sketch" / "sketch.bin").write_bytes(struct.pack("<BIII", 25, 4000, 42, 7))
    r = subprocess.run([sys.executable, os.path.join(SCRIPTS, "simkaMin_update.py"), "-in", "new.txt", "-in-to-update", str(ex), "-bin", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    c = calls(log)
    old, new, dd = str(ex / "sketch" / "sketch.bin"), str(ex / "sketch" / "sketch_new.bin"), str(ex / "distance")
    assert [x[0] for x in c] == ["sketch", "distance", "distance", "matrix-update", "append", "export"]
    # the new sketch inherits seed / k / sketch size from the existing one (simkaMin_update.py:119-125)
    assert c[0][:11] == ["sketch", "-in", "new.txt", "-out", new, "-seed", "42", "-kmer-size", "25", "-nb-kmers", "4000"]
    assert c[1] == ["distance", "-in1", old, "-in2", new, "-out", os.path.join(dd, "existingVsNew"), "-nb-cores", "0"]
    assert c[2] == ["distance", "-in1", new, "-in2", new, "-out", os.path.join(dd, "newVsNew"), "-nb-cores", "0"]
    assert c[3] == ["matrix-update", "-in", dd, "-in1", old, "-in2", new]
    assert c[4] == ["append", "-in1", old, "-in2", new]
    assert c[5] == ["export", "-in", dd, "-in1", old, "-in2", old, "-out", str(ex), "-nb-cores", "0"]
    assert not os.path.exists(new) and os.listdir(dd) == []  # temporaries removed


def test_read_sketch_header(tmp_path):
    f = tmp_path / "s.bin"
    f.write_bytes(struct.pack("<BIII", 21, 1000000, 100, 128) + b"rest")
    assert sp.read_sketch_header(str(f)) == {"kmerSize": 21, "sketchSize": 1000000, "seed": 100, "nbDatasets": 128}
    f.write_bytes(b"short")
    with pytest.raises(ValueError):
        sp.read_sketch_header(str(f))


# ---------------------------------------------------------------------------------------------------------------
# GPU: the real binary
# ---------------------------------------------------------------------------------------------------------------
def _csv(path):
    return gzip.open(path, "rt").read()


@pytest.mark.gpu
@pytest.mark.parametrize("k,max_reads,nb_kmers,filt", [(21, 0, 1000, False), (31, 100, 100, True)])
def test_simkamin_py_reproduces_reference_truth(tmp_path, golden_dir, k, max_reads, nb_kmers, filt):
    """the reference's own SimkaMin test (tests/simkaMin/test_simkaMin.py:141-154) through the wrapper script"""
    out = tmp_path / "res"
    r = subprocess.run([sys.executable, os.path.join(SCRIPTS, "simkaMin.py"), "-in", os.path.join(golden_dir, "example", "simka_input.txt"),
                        "-out", str(out), "-kmer-size", str(k), "-nb-kmers", str(nb_kmers), "-max-reads", str(max_reads)]
                       + (["-filter"] if filt else []), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    tdir = os.path.join(golden_dir, "truth_simkaMin", f"k{k}_{'filter' if filt else ''}_{max_reads}-{nb_kmers}_n1")
    for f in ("mat_presenceAbsence_jaccard", "mat_abundance_braycurtis"):
        assert _csv(out / (f + ".csv.gz")) == open(os.path.join(tdir, f + ".csv")).read(), f


@pytest.mark.gpu
def test_simkamin_update_equals_one_run(tmp_path, golden_dir):
    """samples A,B,C first, then D,E added with simkaMin_update.py: same csv files and same sketch file as one run over
    all five (and hence the reference's truth)"""
    ex = os.path.join(golden_dir, "example")
    lines = [l for l in open(os.path.join(ex, "simka_input.txt")).read().splitlines() if l.strip()]
    assert len(lines) == 5
    # the input grammar resolves relative paths against the input file's directory: keep the split files next to the data
    first, second = tmp_path / "first.txt", tmp_path / "second.txt"
    def absolutise(line):
        ident, files = line.split(":", 1)
        out, tok = "", ""
        for ch in files + "\n":
            if ch in ",;\n":
                t = tok.strip()
                out += (os.path.join(ex, t) if t and not os.path.isabs(t) else t) + (ch if ch != "\n" else "")
                tok = ""
            else:
                tok += ch
        return ident + ": " + out
    first.write_text("\n".join(absolutise(l) for l in lines[:3]) + "\n")
    second.write_text("\n".join(absolutise(l) for l in lines[3:]) + "\n")
    args = ["-kmer-size", "21", "-nb-kmers", "1000", "-max-reads", "0"]
    whole, inc = tmp_path / "whole", tmp_path / "inc"
    for infile, out in ((os.path.join(ex, "simka_input.txt"), whole), (str(first), inc)):
        r = subprocess.run([sys.executable, os.path.join(SCRIPTS, "simkaMin.py"), "-in", infile, "-out", str(out)] + args, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run([sys.executable, os.path.join(SCRIPTS, "simkaMin_update.py"), "-in", str(second), "-in-to-update", str(inc / "simkamin")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert (inc / "simkamin" / "sketch" / "sketch.bin").read_bytes() == (whole / "simkamin" / "sketch" / "sketch.bin").read_bytes()
    tdir = os.path.join(golden_dir, "truth_simkaMin", "k21__0-1000_n1")
    for f in ("mat_presenceAbsence_jaccard", "mat_abundance_braycurtis"):
        assert _csv(inc / "simkamin" / (f + ".csv.gz")) == _csv(whole / (f + ".csv.gz")) == open(os.path.join(tdir, f + ".csv")).read(), f
        a = np.fromfile(inc / "simkamin" / "distance" / (f + ".bin"), dtype=np.float32)
        b = np.fromfile(whole / "simkamin" / "distance" / (f + ".bin"), dtype=np.float32)
        np.testing.assert_array_equal(a, b)
