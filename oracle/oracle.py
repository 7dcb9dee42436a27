"""ctypes front-end of the CPU oracle (oracle/simka_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package (simka_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libsimka_oracle.so"
_lib = None


def build(force: bool = False) -> Path:
    """Compile the oracle with the Makefile next to this file (gcc only)."""
    src_m = max((_HERE / "simka_oracle.c").stat().st_mtime, (_HERE / "simka_oracle.h").stat().st_mtime)
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src_m:
        subprocess.run(["make", "-C", str(_HERE), "-s"], check=True)
    return _LIB_PATH


REF_MURMUR = _HERE / "_ref" / "libmurmur_ref.so"


def build_ref() -> Path | None:
    """oracle/_ref: the reference's own MurmurHash3.cpp compiled where it lies (build_ref.sh); only possible where
    /root/reference exists (the build container).  Returns the library path, or None when it could not be built."""
    if os.path.exists("/root/reference/src/simkaMin/MurmurHash3.cpp"):
        subprocess.run(["make", "-C", str(_HERE), "-s", "ref"], check=True)
    return REF_MURMUR if REF_MURMUR.exists() else None


def ref_murmur3_x64_128_u64(keys: np.ndarray, seed: int) -> np.ndarray:
    """out[0] of the REFERENCE's MurmurHash3_x64_128(&key, 8, seed, out) for every key (oracle/_ref/libmurmur_ref.so)"""
    L = C.CDLL(str(REF_MURMUR))
    f = getattr(L, "_Z19MurmurHash3_x64_128PKvijPv")  # void MurmurHash3_x64_128(const void*, int, uint32_t, void*)
    f.restype = None
    f.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_void_p]
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    out = np.zeros(2, dtype=np.uint64)
    res = np.empty(len(keys), dtype=np.uint64)
    for i in range(len(keys)):
        f(keys[i:i + 1].ctypes.data, 8, seed, out.ctypes.data)
        res[i] = out[0]
    return res


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_LIB_PATH))
        u64p, u32p, u8p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)
        L.orc_canonical_kmers.restype = C.c_size_t
        L.orc_canonical_kmers.argtypes = [C.c_char_p, C.c_size_t, C.c_int, u64p, u8p]
        L.orc_direct_kmers.restype = C.c_size_t
        L.orc_direct_kmers.argtypes = [C.c_char_p, C.c_size_t, C.c_int, u64p, u8p]
        L.orc_revcomp.restype = C.c_uint64
        L.orc_revcomp.argtypes = [C.c_uint64, C.c_int]
        L.orc_minimizer.restype = C.c_uint32
        L.orc_minimizer.argtypes = [C.c_uint64, C.c_int, C.c_int]
        L.orc_run_create.restype = C.c_void_p
        L.orc_run_create.argtypes = [C.c_void_p, u64p, u64p, C.c_int, C.c_int, C.c_int64, C.c_int64,
                                     C.c_int, C.c_int, C.c_int]
        for f in ("orc_run_count", "orc_run_merge", "orc_run_destroy"):
            getattr(L, f).restype = None
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_sample_size.restype = C.c_size_t
        L.orc_sample_size.argtypes = [C.c_void_p, C.c_int]
        L.orc_sample_kmers.restype = u64p
        L.orc_sample_kmers.argtypes = [C.c_void_p, C.c_int]
        L.orc_sample_kmers_hi.restype = u64p
        L.orc_sample_kmers_hi.argtypes = [C.c_void_p, C.c_int]
        L.orc_sample_counts.restype = u32p
        L.orc_sample_counts.argtypes = [C.c_void_p, C.c_int]
        L.orc_sample_totals.restype = None
        L.orc_sample_totals.argtypes = [C.c_void_p, C.c_int, u64p]
        L.orc_get_u64.restype = C.c_int
        L.orc_get_u64.argtypes = [C.c_void_p, C.c_char_p, u64p]
        L.orc_get_f64.restype = C.c_int
        L.orc_get_f64.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_double)]
        L.orc_nb_distinct_kmers.restype = C.c_uint64
        L.orc_nb_distinct_kmers.argtypes = [C.c_void_p]
        L.orc_nb_shared_kmers.restype = C.c_uint64
        L.orc_nb_shared_kmers.argtypes = [C.c_void_p]
        L.orc_nb_matrices.restype = C.c_int
        L.orc_nb_matrices.argtypes = [C.c_void_p]
        L.orc_matrix_name.restype = C.c_char_p
        L.orc_matrix_name.argtypes = [C.c_void_p, C.c_int]
        L.orc_matrix.restype = None
        L.orc_matrix.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float)]
        L.orc_matrix_csv.restype = C.c_size_t
        L.orc_matrix_csv.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.c_char_p, C.c_size_t]
        L.orc_simkamin_distance.restype = None
        L.orc_simkamin_distance.argtypes = [u64p, u32p, C.c_int, C.c_uint32, C.POINTER(C.c_float),
                                            C.POINTER(C.c_float), C.c_int]
        L.orc_murmur3_x64_128_u64.restype = C.c_uint64
        L.orc_murmur3_x64_128_u64.argtypes = [C.c_uint64, C.c_uint32]
        for f in (L.orc_simkamin_sketch, L.orc_simkamin_sketch_filtered):
            f.restype = C.c_uint32
            f.argtypes = [C.c_void_p, u64p, C.c_uint64, C.c_int, C.c_uint32, C.c_uint32, u64p, u32p]
        L.orc_read_fastx.restype = C.c_int64
        L.orc_read_fastx.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), u64p, u64p,
                                     C.POINTER(C.c_void_p), u64p, u64p]
        _lib = L
    return _lib


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


def canonical_kmers(seq: bytes, k: int):
    n = max(len(seq) - k + 1, 0)
    km = np.zeros(max(n, 1), np.uint64)
    va = np.zeros(max(n, 1), np.uint8)
    m = lib().orc_canonical_kmers(seq, len(seq), k, _p(km, C.c_uint64), _p(va, C.c_uint8))
    return km[:m], va[:m].astype(bool)


def direct_kmers(seq: bytes, k: int):
    n = max(len(seq) - k + 1, 0)
    km = np.zeros(max(n, 1), np.uint64)
    va = np.zeros(max(n, 1), np.uint8)
    m = lib().orc_direct_kmers(seq, len(seq), k, _p(km, C.c_uint64), _p(va, C.c_uint8))
    return km[:m], va[:m].astype(bool)


def read_fastx(path: str):
    """-> (bytes-array uint8, offsets uint64[nreads+1]) with the kseq rules of BankFasta."""
    L = lib()
    buf, off = C.c_void_p(None), C.c_void_p(None)
    blen, bcap, noff, ocap = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
    n = L.orc_read_fastx(os.fsencode(path), C.byref(buf), C.byref(blen), C.byref(bcap),
                         C.byref(off), C.byref(noff), C.byref(ocap))
    if n < 0:
        raise FileNotFoundError(path)
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    data = np.frombuffer((C.c_uint8 * blen.value).from_address(buf.value), np.uint8).copy() if blen.value else np.zeros(0, np.uint8)
    offs = np.frombuffer((C.c_uint64 * noff.value).from_address(off.value), np.uint64).copy() if noff.value else np.zeros(1, np.uint64)
    libc.free(buf)
    libc.free(off)
    return data, offs


class Run:
    """One oracle run over N samples given as lists of read byte-strings or a packed layout."""

    def __init__(self, seqs: np.ndarray, read_off: np.ndarray, sample_read_off: np.ndarray, k: int,
                 abundance_min: int = 2, abundance_max: int = 999999999, simple: bool = True,
                 complex_: bool = True, nthreads: int = 0):
        self._seqs = np.ascontiguousarray(seqs, np.uint8)
        if self._seqs.size == 0:
            self._seqs = np.zeros(1, np.uint8)
        self._ro = np.ascontiguousarray(read_off, np.uint64)
        self._so = np.ascontiguousarray(sample_read_off, np.uint64)
        self.n = len(self._so) - 1
        self.k = k
        self.simple, self.complex_ = bool(simple), bool(complex_)
        L = lib()
        self._h = L.orc_run_create(self._seqs.ctypes.data_as(C.c_void_p), _p(self._ro, C.c_uint64),
                                   _p(self._so, C.c_uint64), self.n, k, abundance_min, abundance_max,
                                   int(simple), int(complex_), nthreads)
        if not self._h:
            raise ValueError("orc_run_create failed (k must be 1..63)")

    @classmethod
    def from_reads(cls, samples, k, **kw):
        """samples: list (per sample) of lists of bytes reads."""
        flat = [r for s in samples for r in s]
        lens = np.array([len(r) for r in flat], np.uint64)
        ro = np.zeros(len(flat) + 1, np.uint64)
        np.cumsum(lens, out=ro[1:])
        so = np.zeros(len(samples) + 1, np.uint64)
        np.cumsum([len(s) for s in samples], out=so[1:])
        seqs = np.frombuffer(b"".join(flat), np.uint8)
        return cls(seqs, ro, so, k, **kw)

    def count(self):
        lib().orc_run_count(self._h)
        return self

    def merge(self):
        lib().orc_run_merge(self._h)
        return self

    def sample(self, s):
        L = lib()
        n = L.orc_sample_size(self._h, s)
        if n == 0:
            return np.zeros(0, np.uint64), np.zeros(0, np.uint32)
        km = np.ctypeslib.as_array(L.orc_sample_kmers(self._h, s), (n,)).copy()
        ct = np.ctypeslib.as_array(L.orc_sample_counts(self._h, s), (n,)).copy()
        if self.k > 31:  # 128-bit k-mers: (n, 2) array of [low word, high word], ascending as 128-bit integers
            hi = np.ctypeslib.as_array(L.orc_sample_kmers_hi(self._h, s), (n,)).copy()
            return np.stack([km, hi], axis=1), ct
        return km, ct

    def totals(self, s):
        out = np.zeros(4, np.uint64)
        lib().orc_sample_totals(self._h, s, _p(out, C.c_uint64))
        return out

    def all_totals(self):
        return np.stack([self.totals(s) for s in range(self.n)])

    def u64(self, name):
        n = self.n
        packed = name in ("distinct_shared", "braycurtis_num")  # upper triangle with diagonal, N (N + 1) / 2 entries
        out = np.zeros(n * (n + 1) // 2 if packed else n * n, np.uint64)
        if lib().orc_get_u64(self._h, name.encode(), _p(out, C.c_uint64)) != 0:
            raise KeyError(name)
        return out if packed else out.reshape(n, n)

    def f64(self, name):
        n = self.n
        out = np.zeros(n * n, np.float64)
        if lib().orc_get_f64(self._h, name.encode(), _p(out, C.c_double)) != 0:
            raise KeyError(name)
        return out.reshape(n, n)

    @property
    def nb_distinct_kmers(self):
        return int(lib().orc_nb_distinct_kmers(self._h))

    @property
    def nb_shared_kmers(self):
        return int(lib().orc_nb_shared_kmers(self._h))

    def matrix_names(self):
        L = lib()
        return [L.orc_matrix_name(self._h, m).decode() for m in range(L.orc_nb_matrices(self._h))]

    def matrix(self, m):
        if isinstance(m, str):
            m = self.matrix_names().index(m)
        out = np.zeros((self.n, self.n), np.float32)
        lib().orc_matrix(self._h, m, _p(out, C.c_float))
        return out

    def csv(self, m, ids):
        if isinstance(m, str):
            m = self.matrix_names().index(m)
        arr = (C.c_char_p * self.n)(*[i.encode() for i in ids])
        need = lib().orc_matrix_csv(self._h, m, arr, None, 0)
        buf = C.create_string_buffer(need + 1)
        lib().orc_matrix_csv(self._h, m, arr, buf, need + 1)
        return buf.raw[:need].decode()

    def close(self):
        if self._h:
            lib().orc_run_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def simkamin_sketch(seqs: np.ndarray, read_off: np.ndarray, k: int, sketch_size: int, seed: int = 100, filter: bool = False):
    """-> (hashes uint64[s] ascending with zero padding at the front, counts uint32[s], n_selected); filter = the -filter option"""
    L = lib()
    seqs = np.ascontiguousarray(seqs, np.uint8)
    read_off = np.ascontiguousarray(read_off, np.uint64)
    if seqs.size == 0:
        seqs = np.zeros(1, np.uint8)
    h = np.zeros(sketch_size, np.uint64)
    c = np.zeros(sketch_size, np.uint32)
    fn = L.orc_simkamin_sketch_filtered if filter else L.orc_simkamin_sketch
    n = fn(seqs.ctypes.data_as(C.c_void_p), _p(read_off, C.c_uint64), len(read_off) - 1, k, sketch_size, seed, _p(h, C.c_uint64), _p(c, C.c_uint32))
    return h, c, int(n)


def simkamin_distance(hashes: np.ndarray, counts: np.ndarray, nthreads: int = 0):
    hashes = np.ascontiguousarray(hashes, np.uint64)
    counts = np.ascontiguousarray(counts, np.uint32)
    n, s = hashes.shape
    jac = np.zeros((n, n), np.float32)
    bc = np.zeros((n, n), np.float32)
    lib().orc_simkamin_distance(_p(hashes, C.c_uint64), _p(counts, C.c_uint32), n, s,
                                _p(jac, C.c_float), _p(bc, C.c_float), nthreads)
    return jac, bc


def parse_simka_input(path: str):
    """Input grammar of src/core/SimkaAlgorithm.cpp:267-335: `ID: f1 , f2 ; g1 , g2`.
    -> list of (id, [[files of paired-group 0], [files of group 1], ...])."""
    base = os.path.dirname(os.path.abspath(path))
    out = []
    with open(path) as fh:
        for line in fh:
            line = line.strip().replace(" ", "")
            if not line:
                continue
            ident, rest = line.split(":", 1)
            groups = []
            for grp in rest.split(";"):
                files = [f if os.path.isabs(f) else os.path.join(base, f) for f in grp.split(",") if f]
                groups.append(files)
            out.append((ident, groups))
    return out


def run_from_input_file(path: str, k: int, **kw):
    """All reads of all files of a sample, in file order (-max-reads -1, no read filter)."""
    spec = parse_simka_input(path)
    datas, offs, sample_off, ids = [], [np.zeros(1, np.uint64)], [0], []
    base = 0
    nreads = 0
    for ident, groups in spec:
        ids.append(ident)
        for grp in groups:
            for f in grp:
                d, o = read_fastx(f)
                datas.append(d)
                offs.append(o[1:] + np.uint64(base))
                base += len(d)
                nreads += len(o) - 1
        sample_off.append(nreads)
    seqs = np.concatenate(datas) if datas else np.zeros(0, np.uint8)
    ro = np.concatenate(offs)
    return Run(seqs, ro, np.array(sample_off, np.uint64), k, **kw), ids
