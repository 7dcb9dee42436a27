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


# ------------------------------------------------------------------------------------------------
# Sample-sharded mode: every rank counts its own samples (the reference's one simkaCount process per
# sample), then the solid (k-mer, count) lists are exchanged so that every rank merges a contiguous
# block of hash ranges over ALL samples (the reference's one simkaMerge process per partition reading
# the solid/part_P files of every sample).  Only compacted solid k-mers cross NVLink (12 B per
# sample-distinct solid k-mer), never reads or raw k-mer occurrences.
# ------------------------------------------------------------------------------------------------
class _DevArray:
    """zero-copy view of raw device memory for torch.as_tensor"""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def _dev_tensor(ptr: int, n: int, typestr: str, device) -> torch.Tensor:
    dt = {"<i8": torch.int64, "<i4": torch.int32}[typestr]
    if n == 0:
        return torch.empty(0, dtype=dt, device=device)
    if torch.device(device).type == "cpu":  # host memory (gloo tests of the exchange logic with stand-in contexts)
        import ctypes
        ct = {"<i8": ctypes.c_int64, "<i4": ctypes.c_int32}[typestr]
        return torch.from_numpy(np.ctypeslib.as_array((ct * n).from_address(ptr)))
    return torch.as_tensor(_DevArray(ptr, n, typestr), device=device)


def sample_owner(sample: int, world: int) -> int:
    return sample % world


def slice_plan(fine_bits: int, world: int):
    """-> (parts, local_bits): the 2^fine_bits fine ranges are cut into `parts` equal blocks (a power of two
    <= world); rank r < parts merges block r, the other ranks merge nothing."""
    parts = 1
    while parts * 2 <= world and parts * 2 <= (1 << fine_bits):
        parts *= 2
    local_bits = fine_bits - (parts.bit_length() - 1)
    return parts, local_bits


def exchange_and_merge(count_ctx, merge_ctx, nb_samples: int, rank: int, world: int, device, all_to_all=None, all_reduce_sum=None):
    """Runs the barrier + exchange + stage 2 for one rank and returns this rank's SimkaStatistics
    (NOT yet summed over ranks: call allreduce_stats).  `count_ctx` has finished the samples this rank owns
    (sample_owner); `merge_ctx` is a fresh context of the same parameters.
    all_to_all(send, send_splits, recv_splits) -> recv and all_reduce_sum(int64 tensor) default to
    torch.distributed; tests pass in-process stand-ins."""
    if all_to_all is None:
        def all_to_all(send, send_splits, recv_splits):
            recv = torch.empty(int(sum(recv_splits)), dtype=send.dtype, device=send.device)
            dist.all_to_all_single(recv, send, list(map(int, recv_splits)), list(map(int, send_splits)))
            return recv
    if all_reduce_sum is None:
        def all_reduce_sum(t):
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return t
    N = nb_samples
    mine = [s for s in range(N) if sample_owner(s, world) == rank]
    import os
    import time
    timing = os.environ.get("SIMKA_DIST_TIMING")
    marks = [("start", time.perf_counter())]

    def mark(name):
        if timing:
            if torch.device(device).type == "cuda":
                torch.cuda.synchronize(device)
            marks.append((name, time.perf_counter()))
    # C1: global per-sample totals (only the owner has non-zero rows)
    tot = count_ctx.get_totals()
    own = np.zeros_like(tot)
    own[mine] = tot[mine]
    g_tot = all_reduce_sum(torch.from_numpy(own.view(np.int64).copy()).to(device)).cpu().numpy().view(np.uint64)
    # common fine-range geometry
    fb = count_ctx.fine_bits() if mine else -1
    fbt = all_reduce_sum(torch.tensor([max(fb, 0), 1 if fb >= 0 else 0, max(fb, 0) ** 2], dtype=torch.int64, device=device)).cpu().numpy()
    if fbt[1] == 0:
        raise RuntimeError("no rank counted a sample")
    fine_bits = int(fbt[0] // fbt[1])
    if fbt[1] * fbt[2] != fbt[0] * fbt[0]:
        raise RuntimeError("ranks disagree on fine_bits: pass the same expected_kmers_per_sample to every rank")
    parts, lb = slice_plan(fine_bits, world)
    per = 1 << lb
    fgn = per + 1
    # split points of my samples, sizes of every (sample, destination)
    cnt = np.zeros((N, world), np.int64)
    exports = {}
    offs = []
    for s in mine:
        pk, pc, po, n = count_ctx.export_sample(s)
        off = _dev_tensor(po, (1 << fine_bits) + 1, "<i4", device)  # (uint32 offsets viewed as int32: masked back below)
        offs.append(off[:: per][: parts + 1])
        exports[s] = [pk, pc, off, n, None]
    if mine:  # one device->host copy for the split points of all my samples
        all_cuts = (torch.stack(offs).to(torch.int64) & 0xFFFFFFFF).cpu().numpy()  # a sample may hold >= 2^31 solid k-mers
        for i, s in enumerate(mine):
            exports[s][4] = all_cuts[i]
            cnt[s, :parts] = np.diff(all_cuts[i])
    mark("totals+export")
    cnt = all_reduce_sum(torch.from_numpy(cnt).to(device)).cpu().numpy()
    mark("sizes")
    # send buffers, destination-major then sample-major
    se, sc, so = [], [], []
    for r in range(world):
        for s in mine:
            pk, pc, off, n, cuts = exports[s]
            if r < parts:
                a, b = int(cuts[r]), int(cuts[r + 1])
                se.append(_dev_tensor(pk, n, "<i8", device)[a:b])
                sc.append(_dev_tensor(pc, n, "<i4", device)[a:b])
                so.append(off[r * per: r * per + fgn] - off[r * per])
            else:
                so.append(torch.zeros(fgn, dtype=torch.int32, device=device))
    send_e = torch.cat(se) if se else torch.empty(0, dtype=torch.int64, device=device)
    send_c = torch.cat(sc) if sc else torch.empty(0, dtype=torch.int32, device=device)
    send_o = torch.cat(so) if so else torch.empty(0, dtype=torch.int32, device=device)
    mark("pack")
    owners = [[s for s in range(N) if sample_owner(s, world) == o] for o in range(world)]
    send_splits = [int(cnt[mine, r].sum()) if mine else 0 for r in range(world)]
    recv_splits = [int(cnt[owners[o], rank].sum()) if owners[o] else 0 for o in range(world)]
    recv_e = all_to_all(send_e, send_splits, recv_splits)
    recv_c = all_to_all(send_c, send_splits, recv_splits)
    recv_o = all_to_all(send_o, [len(mine) * fgn] * world, [len(owners[o]) * fgn for o in range(world)])
    if torch.device(device).type == "cuda":
        torch.cuda.current_stream(device).synchronize()  # the merge context runs on its own stream
    mark("all_to_all")
    # stage 2 on my block of fine ranges over all samples
    pos_e, pos_o = 0, 0
    for o in range(world):
        for s in owners[o]:
            n = int(cnt[s, rank])
            merge_ctx.import_sample(s, recv_e.data_ptr() + 8 * pos_e, recv_c.data_ptr() + 4 * pos_e, n,
                                    recv_o.data_ptr() + 4 * pos_o, lb, int(g_tot[s, 0]))
            pos_e += n
            pos_o += fgn
    merge_ctx._keep.append((recv_e, recv_c, recv_o))
    merge_ctx.set_totals(g_tot)
    mark("import")
    merge_ctx.merge_and_accumulate()
    mark("merge")
    st = merge_ctx.stats()
    mark("stats")
    if timing and rank == 0:
        print("[exchange_and_merge ms] " + " ".join(f"{n}={1e3 * (t - marks[i][1]):.2f}" for i, (n, t) in enumerate(marks[1:])), flush=True)
    st.totals[:] = g_tot
    return st


# ------------------------------------------------------------------------------------------------
# The same, inside the library: simka_gpu_exchange_and_merge (csrc/simka_gpu.cu) does the totals all-reduce, the list
# exchange (ncclSend / ncclRecv straight from the per-sample lists), the merge and the accumulator all-reduce in C++ on its
# own NCCL communicator.  torch.distributed only carries the 128-byte NCCL unique id to the ranks, once per context.
# ------------------------------------------------------------------------------------------------
def native_comm_init(ctx, rank: int, world: int, device):
    """creates the library's NCCL communicator for this context (collective: every rank calls it)"""
    from .api import comm_unique_id
    ident = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        ident = torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8).to(device)
    dist.broadcast(ident, src=0)
    ctx.comm_init(bytes(ident.cpu().numpy().tobytes()), rank, world)


def exchange_and_merge_native(ctx):
    """-> the GLOBAL SimkaStatistics (already summed over the ranks) after simka_gpu_exchange_and_merge"""
    ctx.exchange_and_merge()
    return ctx.stats()
