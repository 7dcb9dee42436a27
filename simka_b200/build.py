"""In-tree build of libsimka_b200.so (hand-written CUDA for sm_100a + the C++ host side).

`python -m simka_b200.build` or simka_b200.build.build().  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
LIBDIR = ROOT / "lib"
LIB = LIBDIR / "libsimka_b200.so"
SIMKA_BIN = ROOT / "bin" / "simka"

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-diag-suppress", "128",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (no CPU fallback exists)")


def _stale(target: Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources)


def build(force: bool = False, verbose: bool = False) -> Path:
    srcs = [CSRC / "simka_gpu.cu", CSRC / "simka_host.cpp"]
    deps = srcs + [CSRC / "kmer_device.cuh", ROOT.parent / "include" / "simka_b200.h", *sorted((CSRC / "host").glob("*.hpp")), *sorted((CSRC / "kernels").glob("*.cuh"))]
    LIBDIR.mkdir(exist_ok=True)
    if force or _stale(LIB, deps):
        extra = os.environ.get("SIMKA_B200_NVCC_EXTRA", "").split()  # experiments: -DSIMKA_SK2_SLOTS=... etc.
        cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-shared", "-o", str(LIB), *map(str, srcs), "-lz"]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    BIN = SIMKA_BIN.parent
    BIN.mkdir(exist_ok=True)
    cuda_lib = str(Path(_nvcc()).resolve().parent.parent / "lib64")
    # the command lines of the reference (g++, linked against the library): simka, simkaCount, simkaMerge, simkaCountProcess, simkaMinCore
    for exe, src, link in (("simka", "simka_main.cpp", True), ("simkaCount", "simka_count_main.cpp", True),
                           ("simkaMerge", "simka_merge_main.cpp", True), ("simkaCountProcess", "simka_count_process_main.cpp", False),
                           ("simkaMinCore", "simkamin_main.cpp", True)):
        cli, out = CSRC / src, BIN / exe
        if not cli.exists():
            continue
        if force or _stale(out, deps + [cli]) or _stale(out, [LIB]):
            cmd = ["g++", "-O2", "-std=c++17", "-o", str(out), str(cli)]
            if link:
                cmd += [f"-L{LIBDIR}", "-lsimka_b200", "-Wl,-rpath,$ORIGIN/../lib", f"-Wl,-rpath-link,{cuda_lib}", "-lz", "-lpthread"]
            subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
