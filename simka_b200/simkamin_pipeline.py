"""SimkaMin pipelines over the `simkaMinCore` binary of this build -- the two wrapper scripts of the reference,
`simkaMin/simkaMin.py` (sketch -> distance -> export, :44-61,84-197) and `simkaMin/simkaMin_update.py` (add datasets to
existing results: sketch the new ones, existing-vs-new and new-vs-new distances, matrix-update, append, export, :44-58,
84-222), with the same options, directory layout (`<out>/simkamin/{sketch/sketch.bin,distance/mat_*.bin,logs/}`) and
result files.  All computation happens in `simkaMinCore` (GPU); this module only sequences the sub-commands.

Distances are issued in blocks of `-block-size` datasets (100, the reference's MAX_DATASETS_PROCESS): like the
reference, a rectangular off-diagonal block is written at (i, j) only, so with more datasets than one block the
exported matrix is upper-block-triangular; give a `-block-size` >= the number of datasets for a fully mirrored matrix
(one kernel launch covers all pairs).
"""
from __future__ import annotations

import argparse
import os
import shutil
import struct
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_CORE = os.path.join(_HERE, "bin", "simkaMinCore")
MAX_DATASETS_PROCESS = 100  # simkaMin/simkaMin.py:159


def read_sketch_header(path: str) -> dict:
    """u8 kmerSize, u32 sketchSize, u32 seed, u32 nbDatasets (simkaMin/simkaMin_utils.py:95-108)"""
    with open(path, "rb") as f:
        raw = f.read(13)
    if len(raw) != 13:
        raise ValueError(f"{path}: not a sketch file")
    k, s, seed, n = struct.unpack("<BIII", raw)
    return {"kmerSize": k, "sketchSize": s, "seed": seed, "nbDatasets": n}


def _is_executable(p: str) -> bool:
    return bool(p) and (shutil.which(p) is not None or (os.path.isfile(p) and os.access(p, os.X_OK)))


def find_core(user_bin: str | None) -> str:
    """-bin, else `simkaMinCore` on PATH, else the binary built in-tree (simkaMin.py:66-83)"""
    if user_bin is not None:
        if not _is_executable(user_bin):
            print("Error: " + user_bin + " not found or not executable, should be <simka_b200>/bin/simkaMinCore")
            raise SystemExit(1)
        return user_bin
    if _is_executable("simkaMinCore"):
        return "simkaMinCore"
    if _is_executable(DEFAULT_CORE):
        return DEFAULT_CORE
    print("Error: simkaMinCore executable not found, please give the executable path with option -bin "
          "(python -m simka_b200.build puts it in simka_b200/bin/simkaMinCore)")
    raise SystemExit(1)


def _run(cmd, title=None, log=None):
    if title:
        print("\n\n#-----------------------------\n# " + title)
        print(" ".join(cmd) + "\n#-----------------------------\n", flush=True)
    if log:
        with open(log, "wb") as lf:
            ret = subprocess.run(cmd, stdout=lf, stderr=subprocess.STDOUT).returncode
    else:
        ret = subprocess.run(cmd).returncode
    if ret != 0:
        print("ERROR")
        raise SystemExit(1)


def _read_options(p, kmer_group):
    kmer_group.add_argument("-filter", action="store_true", dest="filter", help="filter out k-mer seen one time (potentially erroneous)")
    r = p.add_argument_group("[read options]")
    r.add_argument("-max-reads", dest="max_reads", default="0", help="maximum number of reads per sample to process")
    r.add_argument("-min-read-size", dest="min_read_size", default="0", help="minimal size a read should have to be kept")
    r.add_argument("-min-shannon-index", dest="min_shannon_index", default="0", help="minimal Shannon index a read should have to be kept. Float in [0,2]")
    c = p.add_argument_group("[core options]")
    c.add_argument("-nb-cores", dest="nb_cores", default="0", help="number of host cores")
    c.add_argument("-max-memory", dest="max_memory", default="8000", help="max memory (MB)")
    c.add_argument("-block-size", dest="block_size", default=str(MAX_DATASETS_PROCESS), help="datasets per distance block")


def _sketch_cmd(core, a, infile, out, seed, k, nb_kmers):
    cmd = [core, "sketch", "-in", infile, "-out", out, "-seed", str(seed), "-kmer-size", str(k), "-nb-kmers", str(nb_kmers)]
    if a.filter:
        cmd.append("-filter")
    return cmd + ["-max-reads", a.max_reads, "-min-read-size", a.min_read_size, "-min-shannon-index", a.min_shannon_index,
                  "-nb-cores", a.nb_cores, "-max-memory", a.max_memory]


def distance_blocks(n: int, block: int = MAX_DATASETS_PROCESS):
    """(start_i, start_j, n_i, n_j) of the upper-triangular block grid, in the order simkaMin.py:181-190 issues it"""
    steps = -(-n // block)
    for i in range(steps):
        n1 = min(block, n - i * block)
        for j in range(i, steps):
            yield i * block, j * block, n1, min(block, n - j * block)


def main_simkamin(argv=None) -> int:
    p = argparse.ArgumentParser(prog="simkaMin.py", allow_abbrev=False)
    m = p.add_argument_group("[main options]")
    m.add_argument("-in", dest="input_filename", required=True, help="input file of datasets. One sample per line: id1: filename1...")
    m.add_argument("-out", dest="out", default="./simka_results", help="output directory for result files (distance matrices)")
    m.add_argument("-seed", dest="seed", default="100", help="seed used for random k-mer selection")
    m.add_argument("-bin", dest="bin", help="path to the simkaMinCore program")
    k = p.add_argument_group("[k-mer options]")
    k.add_argument("-kmer-size", dest="kmer_size", default="21", help="size of a kmer")
    k.add_argument("-nb-kmers", dest="nb_kmers", default="1000000", help="number of kmers used to compute distances")
    _read_options(p, k)
    a = p.parse_args(argv)
    core = find_core(a.bin)

    out_dir = os.path.join(a.out, "simkamin")
    sketch_dir, dist_dir, logs_dir = (os.path.join(out_dir, d) for d in ("sketch", "distance", "logs"))
    for d in (a.out, out_dir, sketch_dir, dist_dir, logs_dir):
        os.makedirs(d, exist_ok=True)
    sketch = os.path.join(sketch_dir, "sketch.bin")

    _run(_sketch_cmd(core, a, a.input_filename, sketch, a.seed, a.kmer_size, a.nb_kmers), "Sketching")

    print("\n\n#-----------------------------\n# Computing distances\n#-----------------------------\n", flush=True)
    for name in ("mat_presenceAbsence_jaccard.bin", "mat_abundance_braycurtis.bin"):  # fresh, pre-created (simkaMin.py:152-155)
        open(os.path.join(dist_dir, name), "wb").close()
    n = read_sketch_header(sketch)["nbDatasets"]
    block = max(1, int(a.block_size))
    for si, sj, ni, nj in distance_blocks(n, block):
        _run([core, "distance", "-in1", sketch, "-in2", sketch, "-out", dist_dir, "-nb-cores", a.nb_cores,
              "-start-i", str(si), "-start-j", str(sj), "-n-i", str(ni), "-n-j", str(nj)],
             log=os.path.join(logs_dir, f"log_distance_{si // block}-{sj // block}"))

    print("\n\nExporting distance matrices in csv.gz format...", flush=True)
    _run([core, "export", "-in", dist_dir, "-in1", sketch, "-in2", sketch, "-out", a.out, "-nb-cores", a.nb_cores])
    print("\n\n\nResult dir: " + a.out)
    return 0


def main_update(argv=None) -> int:
    p = argparse.ArgumentParser(prog="simkaMin_update.py", allow_abbrev=False)
    m = p.add_argument_group("[main options]")
    m.add_argument("-in", dest="input_filename", required=True, help="input file of datasets (datasets to add to existing simka results)")
    m.add_argument("-in-to-update", dest="input_existingResults", required=True,
                   help="path to existing simka results to update (the <out>/simkamin directory; existing results will be overwritten)")
    m.add_argument("-bin", dest="bin", help="path to the simkaMinCore program")
    _read_options(p, p.add_argument_group("[k-mer options]"))
    a = p.parse_args(argv)
    core = find_core(a.bin)

    existing = a.input_existingResults
    sketch_old = os.path.join(existing, "sketch", "sketch.bin")
    sketch_new = os.path.join(existing, "sketch", "sketch_new.bin")
    dist_dir = os.path.join(existing, "distance")
    d_en, d_nn = os.path.join(dist_dir, "existingVsNew"), os.path.join(dist_dir, "newVsNew")
    os.makedirs(d_en, exist_ok=True)
    os.makedirs(d_nn, exist_ok=True)
    h = read_sketch_header(sketch_old)
    print(h)

    _run(_sketch_cmd(core, a, a.input_filename, sketch_new, h["seed"], h["kmerSize"], h["sketchSize"]), "Sketching new datasets")
    _run([core, "distance", "-in1", sketch_old, "-in2", sketch_new, "-out", d_en, "-nb-cores", a.nb_cores],
         "Computing distances between existing datasets and new datasets")
    _run([core, "distance", "-in1", sketch_new, "-in2", sketch_new, "-out", d_nn, "-nb-cores", a.nb_cores],
         "Computing distances between new datasets")
    _run([core, "matrix-update", "-in", dist_dir, "-in1", sketch_old, "-in2", sketch_new], "Update existing distance matrices")
    _run([core, "append", "-in1", sketch_old, "-in2", sketch_new], "Append new sketch to existing sketch")
    _run([core, "export", "-in", dist_dir, "-in1", sketch_old, "-in2", sketch_old, "-out", existing, "-nb-cores", a.nb_cores],
         "Exporting distances")
    shutil.rmtree(d_en)
    shutil.rmtree(d_nn)
    os.remove(sketch_new)
    print("\n\n\nResult dir: " + existing)
    return 0


if __name__ == "__main__":
    sys.exit(main_update(sys.argv[2:]) if len(sys.argv) > 1 and sys.argv[1] == "update" else main_simkamin(sys.argv[1:]))
