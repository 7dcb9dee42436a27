#!/usr/bin/env python
"""Developer check: host -> device copy rate of pageable and pinned host memory on this box, through torch and through the library's
cudaMemcpy (simkamin_gpu_distance copies its sketches with a plain cudaMemcpy)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import simka_b200  # noqa: E402
from simka_b200 import synth  # noqa: E402

torch.cuda.init()
n = 1 << 27  # 1 GiB of u64
pag = torch.from_numpy(np.arange(n, dtype=np.uint64))
t0 = time.perf_counter(); pin = pag.pin_memory(); print(f"pin_memory of 1 GiB: {time.perf_counter() - t0:.3f} s")
dev = torch.empty(n, dtype=torch.uint64, device="cuda")
for name, src in (("pageable", pag), ("pinned", pin)):
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); dev.copy_(src); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"torch copy_ {name}: {dt * 1e3:.1f} ms = {n * 8 / dt / 1e9:.1f} GB/s")
h, c = synth.make_sketches(64, 1_000_000)
hp, cp = torch.from_numpy(h).pin_memory().numpy(), torch.from_numpy(c).pin_memory().numpy()
for name, (a, b) in (("pageable", (h, c)), ("pinned", (hp, cp)), ("pageable", (h, c)), ("pinned", (hp, cp))):
    t0 = time.perf_counter(); _, _, ms = simka_b200.simkamin_distance(a, b); dt = time.perf_counter() - t0
    print(f"simkamin_distance 64 sketches {name}: wall {dt * 1e3:.1f} ms, kernel {ms:.1f} ms")
