#!/usr/bin/env python
"""Golden vectors produced by the REFERENCE BINARIES themselves (oracle/_ref/bin, built by oracle/build_ref.sh).

The reference's own tests hold truth matrices for k = 21 and 31 only.  This script runs the reference `simka` on its
example (tests/golden/example = /root/reference/example) for the k values its tests do not cover -- k <= 15 (the MiniKC
path, src/minikc/MiniKC.hpp:195-259) and k > 31 (LargeInt<2>, span 64) -- and commits what came out as fixtures:

  tests/golden/truth_ref/[example_n_]results_k<K>_t<AMIN>/mat_*.csv      the 21 matrices (gunzipped)
  tests/golden/truth_ref/results_k<K>_t<AMIN>/count_synchro.json   {sample: [nbReads, nbDistinct, sum c, sum c^2]}

Run from the repo root in the build container (needs oracle/_ref/bin/simka):  python tools/make_ref_golden.py
"""
import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIMKA = os.path.join(ROOT, "oracle", "_ref", "bin", "simka")
OUT = os.path.join(ROOT, "tests", "golden", "truth_ref")
CASES = [("example", k, a) for k in (9, 15, 16, 32, 41, 63) for a in (0, 2)]
# a second input with what the example lacks: N runs, lower case, reads shorter than k, a FASTQ and a gz sample, multi-line FASTA
CASES += [("example_n", k, a) for k in (12, 15, 21, 31, 41) for a in (0, 2)]


# file-format corner cases (make_example_fmt)
CASES += [("example_fmt", k, a) for k in (15, 21, 31) for a in (0, 2)]
# option cases: the read filters and limits of the command line, run by the reference itself (tag -> extra arguments); input example_n
OPTION_CASES = {
    "maxreads50": ["-max-reads", "50"],
    "maxreads0": ["-max-reads", "0"],  # automatic limit: (min + mean) / 2 of the per-group read counts (S4 is a paired sample)
    "minsize90_shannon1.5": ["-min-read-size", "90", "-min-shannon-index", "1.5"],
    "maxreads40_minsize85": ["-max-reads", "40", "-min-read-size", "85"],
    "amax3": ["-abundance-max", "3"],
    "kshannon1.5": ["-kmer-shannon-index", "1.5"],  # accepted by the reference, without effect
}
# (not pinnable: a filter that leaves a sample without reads, e.g. -min-read-size 110 here -- the reference's simkaCount stops with
#  "EXCEPTION: Empty property 'seq_number'" and its `simka` then waits for the sample's token for ever)
OPTION_KT = {"amax3": ((21, 1), (31, 2), (21, 3))}  # k, abundance-min per tag (default: (21, 1) and (31, 2))


def make_example_n(dst):
    """tests/golden/example_n: seeded; written once and committed (the reference ran on exactly these files)"""
    import numpy as np
    rng = np.random.default_rng(20261017)
    alpha = np.frombuffer(b"ACGT", np.uint8)
    g = [alpha[rng.integers(0, 4, 1500)].tobytes() for _ in range(3)]

    def reads(genome, n, L, p_n=0.0, lower=False):
        out = []
        for _ in range(n):
            s = int(rng.integers(0, len(genome) - L))
            r = bytearray(genome[s:s + L])
            if p_n and rng.random() < p_n:
                q = int(rng.integers(0, L - 6))
                ln = int(rng.integers(1, 6))
                r[q:q + ln] = b"N" * ln
            if lower and rng.random() < 0.5:
                r = bytearray(bytes(r).lower())
            if rng.random() < 0.5:
                r = bytearray(bytes(r).translate(bytes.maketrans(b"ACGTacgtN", b"TGCAtgcaN"))[::-1])
            out.append(bytes(r))
        return out

    os.makedirs(dst, exist_ok=True)
    s1 = reads(g[0], 120, 100, p_n=0.4) + [b"ACGT", b"N" * 40, b"ACGTACGTACGTAC"]
    s2 = reads(g[0], 60, 100, lower=True) + reads(g[1], 60, 80, p_n=0.2)
    s3 = reads(g[1], 100, 120, p_n=0.3)
    s4 = reads(g[2], 80, 100) + reads(g[0], 40, 100, p_n=0.5, lower=True)
    with open(os.path.join(dst, "S1.fasta"), "wb") as f:
        for i, r in enumerate(s1):
            f.write(b">r%d\n" % i + r + b"\n")
    with open(os.path.join(dst, "S2.fasta"), "wb") as f:  # multi-line records
        for i, r in enumerate(s2):
            f.write(b">r%d some comment\n" % i + r[:37] + b"\n" + r[37:] + b"\n")
    with gzip.open(os.path.join(dst, "S3.fastq.gz"), "wb") as f:
        for i, r in enumerate(s3):
            f.write(b"@q%d\n" % i + r + b"\n+\n" + b"I" * len(r) + b"\n")
    with open(os.path.join(dst, "S4_1.fasta"), "wb") as f:
        for i, r in enumerate(s4[:60]):
            f.write(b">a%d\n" % i + r + b"\n")
    with open(os.path.join(dst, "S4_2.fasta"), "wb") as f:
        for i, r in enumerate(s4[60:]):
            f.write(b">b%d\n" % i + r + b"\n")
    with open(os.path.join(dst, "simka_input.txt"), "w") as f:
        f.write("S1: S1.fasta\nS2: S2.fasta\nS3: S3.fastq.gz\nS4: S4_1.fasta ; S4_2.fasta\nS5: S1.fasta , S2.fasta\n")


def make_example_fmt(dst):
    """tests/golden/example_fmt: file-format corner cases, seeded, written once and committed: FASTQ whose quality lines start with
    '@', '+' or '>' (some records multi-line), CR-LF line ends, blank lines / header-only records / blanks inside sequence lines /
    no final newline, IUPAC codes with dots, dashes and lower case, a gz file of two concatenated members, long headers with '>'
    and '@' inside"""
    import numpy as np
    rng = np.random.default_rng(11)
    alpha = np.frombuffer(b"ACGT", np.uint8)
    g = alpha[rng.integers(0, 4, 3000)].tobytes()
    R = []
    for _ in range(80):
        s0 = int(rng.integers(0, len(g) - 90))
        R.append(g[s0:s0 + 90])
    files = {}
    fq = b""
    for i, r in enumerate(R):
        q = bytes([b"@+>I"[(i + j) % 4] for j in range(len(r))])
        if i % 3 == 0:
            fq += b"@r%d\n" % i + r[:45] + b"\n" + r[45:] + b"\n+r%d\n" % i + q[:45] + b"\n" + q[45:] + b"\n"
        else:
            fq += b"@r%d desc\n" % i + r + b"\n+\n" + q + b"\n"
    files["nastyq.fastq"] = fq
    files["crlf.fasta"] = b"".join(b">r%d\r\n" % i + r + b"\r\n" for i, r in enumerate(R))
    fa = b""
    for i, r in enumerate(R):
        if i % 10 == 0:
            fa += b">empty%d\n" % i
        if i % 7 == 0:
            fa += b"\n"
        fa += b">r%d\n" % i + r[:30] + b" " + r[30:60] + b"\t" + r[60:] + b"\n"
    files["odd.fasta"] = fa.rstrip(b"\n")
    fa = b""
    for i, r in enumerate(R):
        rr = bytearray(r)
        for ch, pos in ((b"R", 10), (b"y", 25), (b"K", 40), (b"-", 55), (b".", 70), (b"n", 80), (b"U", 5)):
            if (i + pos) % 3 == 0:
                rr[pos] = ch[0]
        if i % 2:
            rr = bytearray(bytes(rr).lower())
        fa += b">r%d\n" % i + bytes(rr) + b"\n"
    files["iupac.fasta"] = fa
    half = len(R) // 2
    files["multi.fasta.gz"] = gzip.compress(b"".join(b">a%d\n" % i + r + b"\n" for i, r in enumerate(R[:half])), mtime=0) + \
        gzip.compress(b"".join(b">b%d\n" % i + r + b"\n" for i, r in enumerate(R[half:])), mtime=0)
    files["hdr.fasta"] = b"".join(b">r%d some>thing @odd " % i + b"x" * 300 + b"\n" + r + b"\n" for i, r in enumerate(R))
    os.makedirs(dst, exist_ok=True)
    for n, c in files.items():
        open(os.path.join(dst, n), "wb").write(c)
    open(os.path.join(dst, "simka_input.txt"), "w").write("".join(f"F{i}: {n}\n" for i, n in enumerate(files)))


def make_config_fixture(dst):
    """tests/golden/config_h5: the config.h5 of a reference run on the example (k = 21) and, for 60 solid k-mers of every partition
    file, the partition the reference's simkaCount put them in -- what host/repartitor.hpp must reproduce"""
    import numpy as np
    with tempfile.TemporaryDirectory() as td:
        shutil.copytree(os.path.join(ROOT, "tests", "golden", "example"), os.path.join(td, "ex"))
        subprocess.run([SIMKA, "-in", os.path.join(td, "ex", "simka_input.txt"), "-out", os.path.join(td, "out"), "-out-tmp", os.path.join(td, "tmp"), "-kmer-size", "21",
                        "-abundance-min", "2", "-nb-cores", "4", "-max-memory", "4000", "-keep-tmp"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, check=True, cwd=td)
        d = os.path.join(td, "tmp", "simka_output_temp")
        os.makedirs(dst, exist_ok=True)
        shutil.copy(os.path.join(d, "config.h5"), os.path.join(dst, "config.h5"))
        pairs = []
        rec = np.dtype([("k", "<u8"), ("b", "<u4"), ("pad", "<u4"), ("c", "<u8")])
        for p in sorted(os.listdir(os.path.join(d, "solid"))):
            for f in sorted(os.listdir(os.path.join(d, "solid", p))):
                r = np.frombuffer(gzip.open(os.path.join(d, "solid", p, f)).read(), rec)
                step = max(1, len(r) // 60)
                pairs += [[int(x), int(p.split("_")[1])] for x in r["k"][::step][:60]]
        json.dump({"kmer_size": 21, "nb_partitions": len(os.listdir(os.path.join(d, "solid"))), "kmer_partition": pairs}, open(os.path.join(dst, "kmer_partition.json"), "w"))
        print(f"config fixture: {len(pairs)} (k-mer, partition) pairs")


def main():
    if not os.path.exists(SIMKA):
        sys.exit("oracle/_ref/bin/simka missing: run oracle/build_ref.sh first")
    exn = os.path.join(ROOT, "tests", "golden", "example_n")
    if not os.path.exists(os.path.join(exn, "simka_input.txt")):
        make_example_n(exn)
    cfg = os.path.join(ROOT, "tests", "golden", "config_h5")
    if not os.path.exists(os.path.join(cfg, "config.h5")):
        make_config_fixture(cfg)
    exf = os.path.join(ROOT, "tests", "golden", "example_fmt")
    if not os.path.exists(os.path.join(exf, "simka_input.txt")):
        make_example_fmt(exf)
    cases = [(inp, k, amin, "", []) for inp, k, amin in CASES]
    cases += [("example_n", k, amin, tag, extra) for tag, extra in OPTION_CASES.items() for k, amin in OPTION_KT.get(tag, ((21, 1), (31, 2)))]
    for inp, k, amin, tag, extra in cases:
        with tempfile.TemporaryDirectory() as td:
            shutil.copytree(os.path.join(ROOT, "tests", "golden", inp), os.path.join(td, "ex"))
            cmd = [SIMKA, "-in", os.path.join(td, "ex", "simka_input.txt"), "-out", os.path.join(td, "out"), "-out-tmp", os.path.join(td, "tmp"),
                   "-kmer-size", str(k), "-abundance-min", str(amin), "-simple-dist", "-complex-dist", "-nb-cores", "4", "-max-memory", "4000",
                   "-keep-tmp"] + extra
            # own process group + time limit: when one of its stage processes dies the reference waits for its token for ever
            proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, cwd=td, start_new_session=True)
            try:
                out, _ = proc.communicate(timeout=120)
            except subprocess.TimeoutExpired:
                import signal
                os.killpg(proc.pid, signal.SIGKILL)
                out, _ = proc.communicate()
                logs = ""
                ldir = os.path.join(td, "tmp", "simka_output_temp", "log")
                if os.path.isdir(ldir):
                    for f in sorted(os.listdir(ldir))[:12]:
                        logs += f"--- {f}\n" + open(os.path.join(ldir, f), errors="replace").read()[-400:] + "\n"
                print(f"{inp} k={k} amin={amin} {tag}: the reference did not finish within 120 s (not pinned)\n{logs[-1500:]}")
                continue
            r = subprocess.CompletedProcess(cmd, proc.returncode, out)
            if r.returncode != 0:
                print(f"{inp} k={k} amin={amin}: reference failed rc={r.returncode}\n{r.stdout[-800:]}")
                continue
            dst = os.path.join(OUT, ("" if inp == "example" else inp + "_") + f"results_k{k}_t{amin}" + ("_" + tag if tag else ""))
            if os.path.exists(os.path.join(dst, "count_synchro.json")) and "--force" not in sys.argv:
                continue
            os.makedirs(dst, exist_ok=True)
            n = 0
            for f in sorted(os.listdir(os.path.join(td, "out"))):
                if f.endswith(".csv.gz"):
                    with gzip.open(os.path.join(td, "out", f), "rb") as g, open(os.path.join(dst, f[:-3]), "wb") as o:
                        o.write(g.read())
                    n += 1
            sync = {}
            sdir = os.path.join(td, "tmp", "simka_output_temp", "count_synchro")
            for f in sorted(os.listdir(sdir)):
                if f.endswith(".ok"):
                    sync[f[:-3]] = [int(x) for x in open(os.path.join(sdir, f)).read().split()]
            json.dump(sync, open(os.path.join(dst, "count_synchro.json"), "w"), indent=1, sort_keys=True)
            if extra:
                json.dump(extra, open(os.path.join(dst, "options.json"), "w"))
            print(f"{inp} k={k} amin={amin} {tag}: {n} matrices, totals {sync}")


if __name__ == "__main__":
    main()
