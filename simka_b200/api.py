"""Host-side mirror of the reference interface for the hot path, on top of the C ABI.

Names follow the reference: SimkaStatistics (src/core/SimkaDistance.hpp:68-139), its `+=`
(SimkaDistance.cpp:156-213) and outputMatrix (:603-649); the two stages are the in-process
replacements of the simkaCount / simkaMerge executables.  Everything that computes
runs in libsimka_b200.so; nothing here falls back to numpy/torch when the library is missing.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import SampleTotals, SimkaParams, StatsView, Timings

ERRORS = {-1: "invalid argument", -2: "CUDA error", -3: "on-chip table overflow", -4: "out of memory", -5: "I/O error"}


class SimkaError(RuntimeError):
    pass


def _vp(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def minikc_replay(seqs: np.ndarray, read_off: np.ndarray, kmer_size: int, any_k: bool = False):
    """k <= 15: every read shorter than k replaced by a copy of the previous long-enough read (the MiniKC iterator quirk,
    simka_minikc_replay in include/simka_b200.h).  any_k: the same for every k (SimkaMin, simkamin_short_read_replay).
    Returns the inputs unchanged when nothing is to be replaced."""
    if (kmer_size > 15 and not any_k) or len(read_off) < 2:
        return seqs, read_off
    L = _lib.load()
    seqs = np.ascontiguousarray(seqs, np.uint8)
    read_off = np.ascontiguousarray(read_off, np.uint64)
    ps, po = C.c_void_p(), C.c_void_p()
    src = seqs if seqs.size else np.zeros(1, np.uint8)
    fn = L.simkamin_short_read_replay if any_k else L.simka_minikc_replay
    rc = fn(_vp(src), _vp(read_off), len(read_off) - 1, kmer_size, C.byref(ps), C.byref(po))
    if rc != 0:
        raise SimkaError(f"simka_minikc_replay failed ({rc})")
    if not ps.value:
        return seqs, read_off
    try:
        off = np.ctypeslib.as_array(C.cast(po, C.POINTER(C.c_uint64)), shape=(len(read_off),)).copy()
        n = int(off[-1])
        out = np.ctypeslib.as_array(C.cast(ps, C.POINTER(C.c_uint8)), shape=(max(n, 1),))[:n].copy()
    finally:
        L.simka_host_free(ps)
        L.simka_host_free(po)
    return out, off


def pack_reads(seqs: np.ndarray, read_off: np.ndarray, nthreads: int = 0, pinned: bool = False, kmer_size: int = 31):
    """2-bit pack + bad mask in the layout simka_gpu_push_reads expects (kmer_size <= 15: the MiniKC rule, only read
    separators are bad, see simka_pack_reads_k).
    -> (packed uint64[n_words+1], bad uint32[n_words+1], n_words)"""
    L = _lib.load()
    seqs, read_off = minikc_replay(seqs, read_off, kmer_size)
    seqs = np.ascontiguousarray(seqs, np.uint8)
    read_off = np.ascontiguousarray(read_off, np.uint64)
    n_reads = len(read_off) - 1
    total = int(read_off[-1] - read_off[0]) if n_reads > 0 else 0
    nw = int(L.simka_pack_words(total, n_reads))
    if pinned:
        packed = pinned_empty(nw + 1, np.uint64)
        bad = pinned_empty(nw + 1, np.uint32)
    else:
        packed = np.empty(nw + 1, np.uint64)
        bad = np.empty(nw + 1, np.uint32)
    if seqs.size == 0:
        seqs = np.zeros(1, np.uint8)
    used = L.simka_pack_reads_k(_vp(seqs), _vp(read_off), n_reads, _vp(packed), _vp(bad), nthreads, kmer_size)
    assert used == nw
    return packed, bad, nw


class _PinnedOwner:
    """frees the cudaHostAlloc'ed block when the last array viewing it is gone"""

    def __init__(self, ptr):
        self.ptr = ptr

    def __del__(self):
        try:
            _lib.load().simka_gpu_host_free(self.ptr)
        except Exception:
            pass


class PinnedArray(np.ndarray):
    """ndarray over pinned host memory; views and slices keep the owner alive"""

    def __array_finalize__(self, obj):
        self._owner = getattr(obj, "_owner", None)


def comm_unique_id() -> bytes:
    """rank 0: the NCCL unique id every rank passes to SimkaGpu.comm_init (simka_comm_unique_id)"""
    buf = (C.c_uint8 * 128)()
    rc = _lib.load().simka_comm_unique_id(buf)
    if rc != 0:
        raise SimkaError(f"simka_comm_unique_id failed ({ERRORS.get(rc, rc)}): libnccl.so.2 not loadable?")
    return bytes(buf)


def pinned_empty(n: int, dtype) -> np.ndarray:
    """numpy array over cudaHostAlloc'd memory (for asynchronous H2D copies)."""
    L = _lib.load()
    dt = np.dtype(dtype)
    ptr = C.c_void_p()
    rc = L.simka_gpu_host_alloc(C.byref(ptr), max(1, n) * dt.itemsize)
    if rc != 0:
        raise SimkaError("cudaHostAlloc failed")
    buf = (C.c_uint8 * (max(1, n) * dt.itemsize)).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=dt, count=n).view(PinnedArray)
    arr._owner = _PinnedOwner(ptr)
    return arr


class SimkaStatistics:
    """Accumulators of one run / one shard; mirrors the reference class of the same name."""

    U64_SYM = ("distinct_shared", "braycurtis_num")
    U64_SQ = ("nb_shared", "chord_ninj", "hellinger", "kulczynski", "canberra", "whittaker")

    def __init__(self, nb_samples: int, simple: bool, complex_: bool):
        n = nb_samples
        self.nb_samples, self.simple, self.complex_ = n, bool(simple), bool(complex_)
        self.nb_distinct_kmers = 0
        self.nb_shared_kmers = 0
        self.totals = np.zeros((n, 4), np.uint64)  # nb_reads, nb_distinct, nb_kmers, chord_n2
        for name in self.U64_SYM:
            setattr(self, name, np.zeros(n * (n + 1) // 2, np.uint64))
        for name in self.U64_SQ:
            setattr(self, name, np.zeros((n, n), np.uint64))
        self.kullback_leibler = np.zeros((n, n), np.float64)

    def view(self) -> StatsView:
        v = StatsView()
        v.nb_samples, v.simple_dist, v.complex_dist = self.nb_samples, int(self.simple), int(self.complex_)
        v.nb_distinct_kmers, v.nb_shared_kmers = self.nb_distinct_kmers, self.nb_shared_kmers
        v.totals = self.totals.ctypes.data_as(C.POINTER(SampleTotals))
        for name in self.U64_SYM + self.U64_SQ:
            setattr(v, name, getattr(self, name).ctypes.data_as(_lib.u64p))
        v.kullback_leibler = self.kullback_leibler.ctypes.data_as(_lib.f64p)
        return v

    def __iadd__(self, other: "SimkaStatistics"):
        """Element-wise sum over partitions / shards (SimkaDistance.cpp:156-213); per-sample totals
        are NOT summed there (they are global already)."""
        self.nb_distinct_kmers += other.nb_distinct_kmers
        self.nb_shared_kmers += other.nb_shared_kmers
        for name in self.U64_SYM + self.U64_SQ:
            a = getattr(self, name)
            a += getattr(other, name)
        self.kullback_leibler += other.kullback_leibler
        return self

    # ---- distances -----------------------------------------------------------------
    def matrix_names(self):
        L = _lib.load()
        v = self.view()
        return [L.simka_matrix_name(C.byref(v), m).decode() for m in range(L.simka_nb_matrices(C.byref(v)))]

    def matrix(self, m) -> np.ndarray:
        L = _lib.load()
        if isinstance(m, str):
            m = self.matrix_names().index(m)
        out = np.zeros((self.nb_samples, self.nb_samples), np.float32)
        v = self.view()
        rc = L.simka_distance_matrix(C.byref(v), m, out.ctypes.data_as(_lib.f32p))
        if rc != 0:
            raise SimkaError(f"simka_distance_matrix: {ERRORS.get(rc, rc)}")
        return out

    def csv(self, m, ids) -> str:
        L = _lib.load()
        if isinstance(m, str):
            m = self.matrix_names().index(m)
        arr = (C.c_char_p * self.nb_samples)(*[i.encode() for i in ids])
        v = self.view()
        need = L.simka_matrix_csv(C.byref(v), m, arr, None, 0)
        buf = C.create_string_buffer(need + 1)
        L.simka_matrix_csv(C.byref(v), m, arr, buf, need + 1)
        return buf.raw[:need].decode()

    def output_matrix(self, out_dir: str, ids):
        """outputMatrix (SimkaDistance.cpp:603-649): writes <out_dir>/mat_*.csv.gz"""
        L = _lib.load()
        os.makedirs(out_dir, exist_ok=True)
        arr = (C.c_char_p * self.nb_samples)(*[i.encode() for i in ids])
        v = self.view()
        rc = L.simka_write_matrices(C.byref(v), arr, os.fsencode(out_dir))
        if rc != 0:
            raise SimkaError(f"simka_write_matrices: {ERRORS.get(rc, rc)}")


class SimkaGpu:
    """One GPU context: stage 1 (simkaCount) per sample, then stage 2 (simkaMerge) once."""

    def __init__(self, nb_samples: int, kmer_size: int = 21, abundance_min: int = 2,
                 abundance_max: int = 999999999, simple_dist: bool = False, complex_dist: bool = False,
                 device: int = 0, shard_rank: int = 0, shard_count: int = 1,
                 expected_kmers_per_sample: int = 0):
        self._L = _lib.load()
        self.nb_samples = nb_samples
        self.kmer_size = kmer_size
        self.simple, self.complex_ = bool(simple_dist), bool(complex_dist)
        p = SimkaParams(kmer_size, nb_samples, abundance_min, abundance_max, int(simple_dist), int(complex_dist),
                        device, shard_rank, shard_count, 0, expected_kmers_per_sample)
        h = C.c_void_p()
        rc = self._L.simka_gpu_create(C.byref(h), C.byref(p))
        if rc != 0:
            raise SimkaError(f"simka_gpu_create failed: {ERRORS.get(rc, rc)} (a CUDA device is required)")
        self._h = h
        self._keep = []

    def _check(self, rc, what):
        if rc != 0:
            msg = self._L.simka_gpu_last_error(self._h)
            raise SimkaError(f"{what}: {ERRORS.get(rc, rc)}: {msg.decode() if msg else ''}")

    def push_reads(self, sample: int, packed: np.ndarray, bad: np.ndarray, n_words: int, n_reads: int):
        assert packed.dtype == np.uint64 and bad.dtype == np.uint32
        assert len(packed) >= n_words + 1 and len(bad) >= n_words + 1
        self._keep.append((packed, bad))
        self._check(self._L.simka_gpu_push_reads(self._h, sample, _vp(packed), _vp(bad), n_words, n_reads), "push_reads")

    def push_reads_device(self, sample: int, d_packed: int, d_bad: int, n_words: int, n_reads: int):
        """device-resident input (raw device pointers, e.g. torch.Tensor.data_ptr()); zero copy"""
        self._check(self._L.simka_gpu_push_reads_device(self._h, sample, C.c_void_p(d_packed), C.c_void_p(d_bad), n_words, n_reads),
                    "push_reads_device")

    def reset(self):
        self._check(self._L.simka_gpu_reset(self._h), "reset")
        self._keep.clear()

    def push_ascii(self, sample: int, seqs: np.ndarray, read_off: np.ndarray):
        seqs, read_off = minikc_replay(seqs, read_off, self.kmer_size)  # k <= 15 only
        seqs = np.ascontiguousarray(seqs, np.uint8)
        read_off = np.ascontiguousarray(read_off, np.uint64)
        if seqs.size == 0:
            seqs = np.zeros(1, np.uint8)
        self._keep.append((seqs, read_off))
        self._check(self._L.simka_gpu_push_ascii(self._h, sample, _vp(seqs), _vp(read_off), len(read_off) - 1), "push_ascii")

    def finish_sample(self, sample: int) -> np.ndarray:
        t = SampleTotals()
        self._check(self._L.simka_gpu_finish_sample(self._h, sample, C.byref(t)), "finish_sample")
        if len(self._keep) > 4:
            del self._keep[:-4]
        return np.array([t.nb_reads, t.nb_distinct, t.nb_kmers, t.chord_n2], np.uint64)

    def dump_sample_counts(self, sample: int):
        if self.kmer_size > 31:  # 128-bit k-mers: (n, 2) array of [low word, high word]
            n = C.c_uint64(0)
            self._check(self._L.simka_gpu_dump_sample_counts128(self._h, sample, None, None, None, C.byref(n)), "dump_sample_counts128")
            lo, hi, ct = np.zeros(max(1, n.value), np.uint64), np.zeros(max(1, n.value), np.uint64), np.zeros(max(1, n.value), np.uint32)
            self._check(self._L.simka_gpu_dump_sample_counts128(self._h, sample, lo.ctypes.data_as(_lib.u64p), hi.ctypes.data_as(_lib.u64p),
                                                                ct.ctypes.data_as(_lib.u32p), C.byref(n)), "dump_sample_counts128")
            return np.stack([lo[:n.value], hi[:n.value]], axis=1), ct[:n.value]
        n = C.c_uint64(0)
        self._check(self._L.simka_gpu_dump_sample_counts(self._h, sample, None, None, C.byref(n)), "dump_sample_counts")
        km = np.zeros(max(1, n.value), np.uint64)
        ct = np.zeros(max(1, n.value), np.uint32)
        self._check(self._L.simka_gpu_dump_sample_counts(self._h, sample, km.ctypes.data_as(_lib.u64p),
                                                         ct.ctypes.data_as(_lib.u32p), C.byref(n)), "dump_sample_counts")
        return km[:n.value], ct[:n.value]

    # ---- multi-GPU inside the library (NCCL): simka_gpu_comm_init / simka_gpu_exchange_and_merge ----
    def comm_init(self, comm_id: bytes, rank: int, world: int):
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(comm_id).ljust(128, b"\0")[:128])
        self._check(self._L.simka_gpu_comm_init(self._h, buf, rank, world), "comm_init")
        self.comm_rank, self.comm_world = rank, world

    def exchange_and_merge(self):
        """after this rank's samples (s % world == rank) are finished: global totals, list exchange, merge of this rank's block
        of hash ranges, accumulators summed over the ranks.  stats() then returns the GLOBAL statistics."""
        self._check(self._L.simka_gpu_exchange_and_merge(self._h), "exchange_and_merge")

    def exchange_timings(self):
        ms = (C.c_float * 4)()
        self._L.simka_gpu_exchange_timings(self._h, ms)
        return {"totals_sizes_ms": ms[0], "exchange_ms": ms[1], "merge_ms": ms[2], "allreduce_ms": ms[3]}

    # ---- sample-sharded multi-GPU: the exchange between the stages (see include/simka_b200.h) ----
    def fine_bits(self) -> int:
        return int(self._L.simka_gpu_fine_bits(self._h))

    def export_sample(self, sample: int):
        """-> (d_kmers, d_counts, d_off, n): raw device pointers owned by this context"""
        pk, pc, po, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_uint64(0)
        self._check(self._L.simka_gpu_export_sample(self._h, sample, C.byref(pk), C.byref(pc), C.byref(po), C.byref(n)), "export_sample")
        return pk.value or 0, pc.value or 0, po.value or 0, int(n.value)

    def import_sample(self, sample: int, d_kmers: int, d_counts: int, n: int, d_off: int, fine_bits: int, nb_reads: int):
        self._check(self._L.simka_gpu_import_sample(self._h, sample, C.c_void_p(d_kmers), C.c_void_p(d_counts), n, C.c_void_p(d_off),
                                                    fine_bits, nb_reads), "import_sample")

    def import_counts(self, sample: int, kmers: np.ndarray, counts: np.ndarray, nb_reads: int = 0):
        """registers a sample from its solid (canonical k-mer, count) list (host arrays) instead of counting reads"""
        kmers = np.ascontiguousarray(kmers, np.uint64)
        counts = np.ascontiguousarray(counts, np.uint32)
        assert len(kmers) == len(counts)
        self._check(self._L.simka_gpu_import_counts(self._h, sample, _vp(kmers) if len(kmers) else None, _vp(counts) if len(counts) else None,
                                                    len(kmers), nb_reads), "import_counts")

    def get_totals(self) -> np.ndarray:
        t = np.zeros((self.nb_samples, 4), np.uint64)
        self._check(self._L.simka_gpu_get_totals(self._h, t.ctypes.data_as(C.POINTER(SampleTotals))), "get_totals")
        return t

    def set_totals(self, totals: np.ndarray):
        t = np.ascontiguousarray(totals, np.uint64)
        assert t.shape == (self.nb_samples, 4)
        self._check(self._L.simka_gpu_set_totals(self._h, t.ctypes.data_as(C.POINTER(SampleTotals))), "set_totals")

    def merge_and_accumulate(self):
        self._check(self._L.simka_gpu_merge_and_accumulate(self._h), "merge_and_accumulate")

    def stats(self) -> SimkaStatistics:
        st = SimkaStatistics(self.nb_samples, self.simple, self.complex_)
        v = st.view()
        self._check(self._L.simka_gpu_get_stats(self._h, C.byref(v)), "get_stats")
        st.nb_distinct_kmers, st.nb_shared_kmers = int(v.nb_distinct_kmers), int(v.nb_shared_kmers)
        return st

    def timings(self) -> dict:
        t = Timings()
        self._check(self._L.simka_gpu_get_timings(self._h, C.byref(t)), "get_timings")
        return {f: getattr(t, f) for f, _ in Timings._fields_}

    @property
    def stream(self) -> int:
        return int(self._L.simka_gpu_stream(self._h) or 0)

    def close(self):
        if getattr(self, "_h", None):
            self._L.simka_gpu_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run_samples(samples, kmer_size: int, abundance_min: int = 2, abundance_max: int = 999999999,
                simple_dist: bool = True, complex_dist: bool = True, device: int = 0, ascii_path: bool = False,
                keep_ctx: bool = False):
    """samples: list of (seqs uint8, read_off uint64) -> SimkaStatistics (+ the context if keep_ctx)."""
    ctx = SimkaGpu(len(samples), kmer_size, abundance_min, abundance_max, simple_dist, complex_dist, device)
    for s, (seqs, off) in enumerate(samples):
        if ascii_path:
            ctx.push_ascii(s, seqs, off)
        else:
            packed, bad, nw = pack_reads(seqs, off, kmer_size=kmer_size)
            ctx.push_reads(s, packed, bad, nw, len(off) - 1)
        ctx.finish_sample(s)
    ctx.merge_and_accumulate()
    st = ctx.stats()
    if keep_ctx:
        return st, ctx
    ctx.close()
    return st


def simkamin_sketch(seqs: np.ndarray, read_off: np.ndarray, kmer_size: int, sketch_size: int, seed: int = 100, device: int = 0, filter: bool = False):
    """SimkaMin sketch of one sample on the GPU (src/simkaMin/SimkaMinCount.hpp:217-338; filter=True: the -filter option, :341-380).
    -> (hashes uint64[s] ascending, zero padding at the front; counts uint32[s]; n_selected)"""
    L = _lib.load()
    seqs, read_off = minikc_replay(seqs, read_off, kmer_size, any_k=True)  # a read shorter than k counts the previous read again
    packed, bad, nw = pack_reads(seqs, read_off)
    h = np.zeros(sketch_size, np.uint64)
    c = np.zeros(sketch_size, np.uint32)
    n = C.c_uint32(0)
    fn = L.simkamin_gpu_sketch_filter if filter else L.simkamin_gpu_sketch
    rc = fn(device, kmer_size, sketch_size, seed, _vp(packed), _vp(bad), nw, len(read_off) - 1,
            h.ctypes.data_as(_lib.u64p), c.ctypes.data_as(_lib.u32p), C.byref(n))
    if rc != 0:
        raise SimkaError(f"simkamin_gpu_sketch: {ERRORS.get(rc, rc)}")
    return h, c, int(n.value)


def simkamin_last_sketch_info() -> dict:
    """How the last simkamin_sketch call of this thread got its sketch (bottom-s pre-filter, include/simka_b200.h):
    counting passes run, k-mer occurrences that passed the hash bound, distinct hashes brought back to the host."""
    L = _lib.load()
    info = _lib.SketchInfo()
    rc = L.simkamin_gpu_last_sketch_info(C.byref(info))
    if rc != 0:
        raise SimkaError(f"simkamin_gpu_last_sketch_info: {ERRORS.get(rc, rc)}")
    return {"attempts": int(info.attempts), "keys_counted": int(info.keys_counted), "distinct_dumped": int(info.distinct_dumped)}


def simkamin_release():
    """frees the device buffers the SimkaMin distance entry points keep between calls"""
    _lib.load().simkamin_gpu_release()


def simkamin_distance(hashes: np.ndarray, counts: np.ndarray, device: int = 0):
    """All-pairs SimkaMin Jaccard / Bray-Curtis (src/simkaMin/SimkaMinDistance.hpp:190-303).
    -> (jaccard float32[n,n], braycurtis float32[n,n], kernel_ms)"""
    L = _lib.load()
    hashes = np.ascontiguousarray(hashes, np.uint64)
    counts = np.ascontiguousarray(counts, np.uint32)
    n, s = hashes.shape
    jac = np.zeros((n, n), np.float32)
    bc = np.zeros((n, n), np.float32)
    ms = C.c_float(0)
    rc = L.simkamin_gpu_distance(device, _vp(hashes), _vp(counts), n, s, jac.ctypes.data_as(_lib.f32p),
                                 bc.ctypes.data_as(_lib.f32p), C.byref(ms))
    if rc != 0:
        raise SimkaError(f"simkamin_gpu_distance: {ERRORS.get(rc, rc)}")
    return jac, bc, ms.value
