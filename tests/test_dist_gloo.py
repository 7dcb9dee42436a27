"""world_size-2 gloo test of the multi-GPU host logic (the two reductions of simka_b200.dist)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    import simka_b200
    from simka_b200 import dist as sdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 4
    rng = np.random.default_rng(10 + rank)
    totals = rng.integers(1, 1000, size=(n, 4)).astype(np.uint64)
    totals[:, 0] = [11, 12, 13, 14]  # nb_reads: identical on every rank
    g = sdist.allreduce_totals(totals)
    st = simka_b200.SimkaStatistics(n, True, True)
    for name in st.U64_SYM + st.U64_SQ:
        a = getattr(st, name)
        a[...] = rng.integers(0, 2**62, size=a.shape, dtype=np.uint64)
    st.whittaker[0, 1] = np.uint64(2**64 - 5 - rank)  # wrap-around must survive the reduction
    st.kullback_leibler[...] = rng.random((n, n))
    st.nb_distinct_kmers, st.nb_shared_kmers = 100 + rank, 10 + rank
    out = sdist.allreduce_stats(st)
    q.put((rank, totals, g, {k: getattr(st, k).copy() for k in st.U64_SYM + st.U64_SQ}, st.kullback_leibler.copy(),
           {k: getattr(out, k).copy() for k in out.U64_SYM + out.U64_SQ}, out.kullback_leibler.copy(),
           out.nb_distinct_kmers, out.nb_shared_kmers))
    dist.barrier()
    dist.destroy_process_group()


def test_allreduce_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (_, t0, g0, u0, k0, o0, ok0, nd0, ns0), (_, t1, g1, u1, k1, o1, ok1, nd1, ns1) = res
    exp = t0 + t1
    exp[:, 0] = t0[:, 0]
    np.testing.assert_array_equal(g0, exp)
    np.testing.assert_array_equal(g1, exp)
    for name in u0:
        with np.errstate(over="ignore"):
            np.testing.assert_array_equal(o0[name], u0[name] + u1[name])
        np.testing.assert_array_equal(o0[name], o1[name])
    np.testing.assert_allclose(ok0, k0 + k1, rtol=1e-15)
    assert (nd0, ns0) == (201, 21) == (nd1, ns1)


# ---- sample-sharded exchange (simka_b200.dist.exchange_and_merge) over gloo, with stand-in contexts ----
class _FakeCountCtx:
    """holds per-sample (hashed k-mer, count) lists in host memory, grouped by the top `fine_bits` of 16-bit keys"""

    def __init__(self, n_samples, fine_bits, rank, world):
        from simka_b200 import dist as sdist
        self.fb = fine_bits
        self.data = {}
        self.tot = np.zeros((n_samples, 4), np.uint64)
        for s in range(n_samples):
            if sdist.sample_owner(s, world) != rank:
                continue
            rng = np.random.default_rng(100 + s)
            keys = np.sort(rng.choice(1 << 16, size=500 + 37 * s, replace=False)).astype(np.int64)
            cnt = rng.integers(1, 50, size=len(keys)).astype(np.int32)
            off = np.searchsorted(keys >> (16 - fine_bits), np.arange((1 << fine_bits) + 1)).astype(np.int32)
            self.data[s] = (keys, cnt, off)
            self.tot[s] = [1000 + s, len(keys), int(cnt.sum()), int((cnt.astype(np.int64) ** 2).sum())]

    def get_totals(self):
        return self.tot

    def fine_bits(self):
        return self.fb

    def export_sample(self, s):
        k, c, o = self.data[s]
        return k.ctypes.data, c.ctypes.data, o.ctypes.data, len(k)


class _FakeMergeCtx:
    def __init__(self):
        self.imported = {}
        self._keep = []
        self.totals = None

    def import_sample(self, s, pk, pc, n, po, fine_bits, nb_reads):
        import ctypes
        k = np.ctypeslib.as_array((ctypes.c_int64 * max(n, 1)).from_address(pk))[:n].copy() if n else np.zeros(0, np.int64)
        c = np.ctypeslib.as_array((ctypes.c_int32 * max(n, 1)).from_address(pc))[:n].copy() if n else np.zeros(0, np.int32)
        o = np.ctypeslib.as_array((ctypes.c_int32 * ((1 << fine_bits) + 1)).from_address(po)).copy()
        self.imported[s] = (k, c, o, fine_bits, nb_reads)

    def set_totals(self, t):
        self.totals = t.copy()

    def merge_and_accumulate(self):
        pass

    def stats(self):
        import simka_b200
        return simka_b200.SimkaStatistics(len(self.imported), False, False)


def _exchange_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from simka_b200 import dist as sdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_samples, fb = 5, 6
    cc, mc = _FakeCountCtx(n_samples, fb, rank, world), _FakeMergeCtx()
    sdist.exchange_and_merge(cc, mc, n_samples, rank, world, torch.device("cpu"))
    q.put((rank, mc.imported, mc.totals))
    dist.barrier()
    dist.destroy_process_group()


def test_sample_sharded_exchange_two_ranks():
    from simka_b200 import dist as sdist
    assert sdist.slice_plan(17, 8) == (8, 14) and sdist.slice_plan(2, 8) == (4, 0) and sdist.slice_plan(10, 1) == (1, 10)
    assert sdist.slice_plan(5, 6) == (4, 3)  # non power-of-two world: 4 merging ranks
    world, n_samples, fb = 2, 5, 6
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_exchange_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    parts, lb = sdist.slice_plan(fb, world)
    ref = {}
    for r in range(world):
        ref.update(_FakeCountCtx(n_samples, fb, r, world).data)
    exp_tot = sum(_FakeCountCtx(n_samples, fb, r, world).tot for r in range(world))
    for rank, imported, totals in res:
        np.testing.assert_array_equal(totals, exp_tot)
        assert sorted(imported) == list(range(n_samples))
        for s, (k, c, o, fbl, nb_reads) in imported.items():
            keys, cnt, off = ref[s]
            lo, hi = off[rank << lb], off[(rank + 1) << lb]
            np.testing.assert_array_equal(k, keys[lo:hi])
            np.testing.assert_array_equal(c, cnt[lo:hi])
            np.testing.assert_array_equal(o, off[rank << lb: ((rank + 1) << lb) + 1] - lo)
            assert fbl == lb and nb_reads == 1000 + s
