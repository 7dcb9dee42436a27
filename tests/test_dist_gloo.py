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
