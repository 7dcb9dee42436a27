#!/usr/bin/env python
"""Memory check of the kernels on the CPU: the emulated library (tests/emu) compiled with AddressSanitizer, every kernel family
run once on awkward inputs (default, super-k-mer and record-path stage 1, k = 31 / 21 / 15 / 41, the tile merge, SimkaMin sketch with and without -filter,
SimkaMin distance).  An out-of-bounds access of "device" or "shared" memory aborts with ASan's report.  Developer tool:
    tools/emu_asan.py            (re-executes itself with libasan preloaded)"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

if os.environ.get("SIMKA_EMU_ASAN_CHILD") != "1":
    from emu import build_emu
    build_emu.build(asan=True)
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True, check=True).stdout.strip()
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0", SIMKA_EMU_ASAN_CHILD="1")
    sys.exit(subprocess.run([sys.executable, os.path.abspath(__file__)], env=env).returncode)

import simka_b200  # noqa: E402
from emu import build_emu  # noqa: E402
from helpers import nasty_samples, synthetic_samples  # noqa: E402
from simka_b200 import _lib, synth  # noqa: E402

L = C.CDLL(str(build_emu.OUT / "libsimka_b200_emu_asan.so"))
for name, (res, args) in _lib.SYMBOLS.items():
    fn = getattr(L, name)
    fn.restype, fn.argtypes = res, args
_lib._lib = L
for superk in ("0", "1"):
    os.environ["SIMKA_B200_SUPERK"] = superk
    for k in (31, 21, 15):
        samples = synthetic_samples(3, 200) + nasty_samples()
        st, ctx = simka_b200.run_samples(samples, k, abundance_min=1, keep_ctx=True)
        for s in range(len(samples)):
            ctx.dump_sample_counts(s)
        ctx.close()
        print("ok: stage 1 + merge, superk", superk, "k", k, flush=True)
os.environ["SIMKA_B200_SUPERK"] = "2"  # record path (superk2.cuh)
samples = synthetic_samples(3, 200) + nasty_samples()
simka_b200.run_samples(samples, 31, abundance_min=1)
print("ok: record path", flush=True)
os.environ.pop("SIMKA_B200_SUPERK")
for n, reads, order in ((70, 12, ""), (150, 6, "reverse"), (100, 10, "random:3")):  # tile merge (merge_tiles.cuh), several thread orders
    os.environ["SIMKA_EMU_SCHED"] = order
    samples = synthetic_samples(n, reads, read_len=70, n_genomes=4, genome_len=400)
    simka_b200.run_samples(samples, 21, abundance_min=1, simple_dist=True, complex_dist=True)
    same = [samples[0]] * n  # every k-mer in every sample: full tiles, chunks shrink to the member capacity
    simka_b200.run_samples(same, 31, abundance_min=1, simple_dist=True, complex_dist=True)
    print("ok: tile merge, N =", n, order or "ascending", flush=True)
os.environ.pop("SIMKA_EMU_SCHED")
simka_b200.run_samples(synthetic_samples(3, 100) + nasty_samples(), 41, abundance_min=1)  # 128-bit k-mers (wide.cuh)
print("ok: k = 41", flush=True)
h, c = synth.make_sketches(5, 3000)
simka_b200.simkamin_distance(h, c)
seqs, off = synthetic_samples(1, 300)[0]
simka_b200.simkamin_sketch(seqs, off, 21, 500, seed=100)
simka_b200.simkamin_sketch(seqs, off, 21, 500, seed=100, filter=True)
print("ok: SimkaMin sketch / distance\nno AddressSanitizer report")
