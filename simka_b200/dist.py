"""Multi-GPU plumbing: one process per GPU, hash buckets sharded across ranks, two tiny all-reduces.

The reference combines its per-partition results only by element-wise sums
(src/SimkaPotara.hpp:1161-1174 -> src/core/SimkaDistance.cpp:156-213) after a barrier that makes the
per-sample totals global (SimkaDistance.cpp:116-151).  Here the same two reductions run over
torch.distributed (NCCL over NVLink on GPUs, gloo on CPU for the tests); no k-mer crosses a link.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .api import SimkaStatistics


def _device():
    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def _allreduce_u64(arr: np.ndarray) -> np.ndarray:
    """sum modulo 2^64 (two's complement int64 addition is the same bit pattern)."""
    t = torch.from_numpy(np.ascontiguousarray(arr).view(np.int64).copy()).to(_device())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy().view(np.uint64).reshape(arr.shape)


def allreduce_totals(local_totals: np.ndarray) -> np.ndarray:
    """C1: global per-sample totals.  Column 0 (nb_reads) is identical on all ranks and not summed."""
    out = _allreduce_u64(local_totals)
    out[:, 0] = local_totals[:, 0]
    return out


def allreduce_stats(st: SimkaStatistics) -> SimkaStatistics:
    """C2: one integer all-reduce (all u64 accumulators packed in one buffer) + one f64 all-reduce."""
    names = SimkaStatistics.U64_SYM + SimkaStatistics.U64_SQ
    parts = [getattr(st, n).reshape(-1) for n in names]
    parts.append(np.array([st.nb_distinct_kmers, st.nb_shared_kmers], np.uint64))
    flat = _allreduce_u64(np.concatenate(parts))
    out = SimkaStatistics(st.nb_samples, st.simple, st.complex_)
    out.totals[:] = st.totals
    pos = 0
    for n in names:
        a = getattr(out, n)
        a.reshape(-1)[:] = flat[pos:pos + a.size]
        pos += a.size
    out.nb_distinct_kmers, out.nb_shared_kmers = int(flat[pos]), int(flat[pos + 1])
    kl = torch.from_numpy(st.kullback_leibler.copy()).to(_device())
    dist.all_reduce(kl, op=dist.ReduceOp.SUM)
    out.kullback_leibler[:] = kl.cpu().numpy()
    return out
