"""Builds tests/emu/_build/libsimka_b200_emu.so: the kernel sources of simka_b200/csrc compiled with g++ against the
functional CUDA stand-in tests/emu/cuda_runtime.h (CUDA threads = cooperative fibers on one OS thread).

TEST INFRASTRUCTURE ONLY (see the header of cuda_runtime.h): exercises kernel logic in the CPU test suite; loaded by
tests/test_emu_kernels.py alone.

The sources are used as they are except for two syntactic rewrites g++ cannot parse:
  kernel<<<grid, block[, smem[, stream]]>>>(args);   ->  emu::launch(grid, block[, smem], [&]{ kernel(args); });
  extern __shared__ [__align__(n)] T name[];         ->  T *name = reinterpret_cast<T *>(emu::dyn_smem());
"""
from __future__ import annotations

import re
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
CSRC = ROOT / "simka_b200" / "csrc"
OUT = HERE / "_build"
LIB = OUT / "libsimka_b200_emu.so"


def _match_back(text: str, end: int) -> int:
    """index where the kernel expression ending at `end` (exclusive) starts: identifier [<...>] [(...)]"""
    i = end
    if text[i - 1] == ")":  # a call returning the kernel pointer
        depth = 0
        while True:
            i -= 1
            depth += text[i] == ")"
            depth -= text[i] == "("
            if depth == 0:
                break
    elif text[i - 1] == ">":  # explicit template arguments
        depth = 0
        while True:
            i -= 1
            depth += text[i] == ">"
            depth -= text[i] == "<"
            if depth == 0:
                break
    while i > 0 and (text[i - 1].isalnum() or text[i - 1] in "_:"):
        i -= 1
    return i


def _split_top(s: str):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def rewrite_launches(text: str) -> tuple[str, int]:
    out, pos, n = "", 0, 0
    while True:
        a = text.find("<<<", pos)
        if a < 0:
            break
        b = text.index(">>>", a)
        start = _match_back(text, a)
        kern = text[start:a]
        cfg = _split_top(text[a + 3:b])
        assert 2 <= len(cfg) <= 4, cfg
        i = b + 3
        assert text[i] == "(", text[a:a + 80]
        depth, j = 0, i
        while True:
            depth += text[j] == "("
            depth -= text[j] == ")"
            j += 1
            if depth == 0:
                break
        args = text[i + 1:j - 1]
        assert text[j] == ";", text[a:j + 1]
        head = ", ".join(cfg[:3])
        out += text[pos:start] + f"emu::launch({head}, [&]{{ {kern}({args}); }});"
        pos = j + 1
        n += 1
    return out + text[pos:], n


DYN_SMEM = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w ]+?)\s+(\w+)\[\];")


def transform(text: str) -> tuple[str, int]:
    text = DYN_SMEM.sub(lambda m: f"{m.group(1)} *{m.group(2)} = reinterpret_cast<{m.group(1)} *>(emu::dyn_smem());", text)
    return rewrite_launches(text)


def build(force: bool = False, asan: bool = False) -> Path:
    """asan=True: a second library compiled with -fsanitize=address (out-of-bounds accesses of "device" and "shared" memory by
    the kernels abort with a report); the process loading it needs LD_PRELOAD=$(gcc -print-file-name=libasan.so)."""
    srcs = [CSRC / "simka_gpu.cu", CSRC / "kmer_device.cuh", *sorted((CSRC / "kernels").glob("*.cuh"))]
    deps = srcs + [CSRC / "simka_host.cpp", *sorted((CSRC / "host").glob("*.hpp")), ROOT / "include" / "simka_b200.h",
                   HERE / "cuda_runtime.h", Path(__file__)]
    lib = OUT / "libsimka_b200_emu_asan.so" if asan else LIB
    if not force and lib.exists() and all(d.stat().st_mtime <= lib.stat().st_mtime for d in deps):
        return lib
    gen = OUT / "csrc"
    (gen / "kernels").mkdir(parents=True, exist_ok=True)
    launches = 0
    for s in srcs:
        text, n = transform(s.read_text())
        launches += n
        text = text.replace('#include "../../include/simka_b200.h"', f'#include "{ROOT / "include" / "simka_b200.h"}"')
        dst = gen / s.relative_to(CSRC)
        if dst.suffix == ".cu":
            dst = dst.with_name(dst.stem + "_emu.cpp")
        dst.write_text(text)
    assert launches >= 30, f"only {launches} kernel launches rewritten"
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-w", f"-I{HERE}", "-DSIMKA_CUDA_EMU=1", "-DSIMKA_NO_NCCL=1",
           *(["-fsanitize=address", "-fno-omit-frame-pointer"] if asan else []),
           "-o", str(lib), str(gen / "simka_gpu_emu.cpp"), str(CSRC / "simka_host.cpp"), "-lz", "-lpthread"]
    subprocess.run(cmd, check=True)
    return lib


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, asan="--asan" in sys.argv))
