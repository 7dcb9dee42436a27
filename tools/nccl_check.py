#!/usr/bin/env python
"""Multi-GPU parity check of simka_gpu_exchange_and_merge (NCCL inside the library), run under torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/nccl_check.py
Every rank counts its samples (s % world == rank), calls exchange_and_merge, and compares the GLOBAL statistics it gets back
with the CPU oracle on all samples (integers bit-exact, Jensen-Shannon sums 1e-10).  Also checks the torch-level exchange of
simka_b200/dist.py against the same oracle.  Exit code 0 = all ranks agree with the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import simka_b200  # noqa: E402
from helpers import assert_stats_equal, oracle_run, synthetic_samples  # noqa: E402
from simka_b200 import dist as sdist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    for n_samples, n_reads, k, amin in ((7, 4000, 31, 2), (3, 2000, 21, 1), (1, 500, 31, 1), (70, 300, 21, 1)):
        samples = synthetic_samples(n_samples, n_reads)
        run = oracle_run(samples, k, amin)
        exp = n_reads * (150 - k + 1)
        ctx = simka_b200.SimkaGpu(n_samples, k, amin, simple_dist=True, complex_dist=True, device=local, expected_kmers_per_sample=exp)
        sdist.native_comm_init(ctx, rank, world, dev)
        for rep in range(2):  # twice: the second run reuses the merge context and its buffers
            ctx.reset()
            for s in range(rank, n_samples, world):
                p, b, nw = simka_b200.pack_reads(*samples[s], kmer_size=k)
                ctx.push_reads(s, p, b, nw, len(samples[s][1]) - 1)
                ctx.finish_sample(s)
            st = sdist.exchange_and_merge_native(ctx)
            assert_stats_equal(st, run)
        t = ctx.timings()
        assert t["pairs"] > 0 or n_samples == 1
        ctx.close()
        if rank == 0:
            print(f"native exchange ok: {n_samples} samples x {n_reads} reads, k={k}, world={world}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
