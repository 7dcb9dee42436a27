"""ctypes binding of libsimka_b200.so (the C ABI declared in include/simka_b200.h)."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "lib" / "libsimka_b200.so"

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
f32p = C.POINTER(C.c_float)
f64p = C.POINTER(C.c_double)


class SimkaParams(C.Structure):
    _fields_ = [("kmer_size", C.c_int32), ("nb_samples", C.c_int32), ("abundance_min", C.c_int64),
                ("abundance_max", C.c_int64), ("simple_dist", C.c_int32), ("complex_dist", C.c_int32),
                ("device", C.c_int32), ("shard_rank", C.c_int32), ("shard_count", C.c_int32),
                ("reserved", C.c_int32), ("expected_kmers_per_sample", C.c_uint64), ("sketch_mode", C.c_int32),
                ("sketch_seed", C.c_uint32)]


class SampleTotals(C.Structure):
    _fields_ = [("nb_reads", C.c_uint64), ("nb_distinct", C.c_uint64), ("nb_kmers", C.c_uint64),
                ("chord_n2", C.c_uint64)]


class StatsView(C.Structure):
    _fields_ = [("nb_samples", C.c_int32), ("simple_dist", C.c_int32), ("complex_dist", C.c_int32),
                ("reserved", C.c_int32), ("nb_distinct_kmers", C.c_uint64), ("nb_shared_kmers", C.c_uint64),
                ("totals", C.POINTER(SampleTotals)), ("distinct_shared", u64p), ("braycurtis_num", u64p),
                ("nb_shared", u64p), ("chord_ninj", u64p), ("hellinger", u64p), ("kulczynski", u64p),
                ("canberra", u64p), ("whittaker", u64p), ("kullback_leibler", f64p)]


class Timings(C.Structure):
    _fields_ = [("extract_count_ms", C.c_float), ("extract_scatter_ms", C.c_float), ("part_count_ms", C.c_float),
                ("part_scatter_ms", C.c_float), ("bin_count_ms", C.c_float), ("compact_ms", C.c_float),
                ("merge_ms", C.c_float), ("samples_ms", C.c_float), ("total_ms", C.c_float),
                ("kmers_in", C.c_uint64), ("entries", C.c_uint64), ("distinct_kmers", C.c_uint64),
                ("pairs", C.c_uint64), ("kernel_launches", C.c_uint64), ("exact_retries", C.c_uint64),
                ("superk_samples", C.c_uint64), ("superk_fallbacks", C.c_uint64)]


class SketchInfo(C.Structure):
    _fields_ = [("attempts", C.c_uint64), ("keys_counted", C.c_uint64), ("distinct_dumped", C.c_uint64)]


# every symbol include/simka_b200.h declares: (restype, argtypes)
SYMBOLS = {
    "simka_gpu_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(SimkaParams)]),
    "simka_gpu_destroy": (None, [C.c_void_p]),
    "simka_gpu_reset": (C.c_int, [C.c_void_p]),
    "simka_gpu_last_error": (C.c_char_p, [C.c_void_p]),
    "simka_gpu_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "simka_gpu_host_free": (C.c_int, [C.c_void_p]),
    "simka_gpu_push_reads": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64]),
    "simka_gpu_push_reads_device": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64]),
    "simka_gpu_push_ascii": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64]),
    "simka_gpu_finish_sample": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(SampleTotals)]),
    "simka_gpu_dump_sample_counts": (C.c_int, [C.c_void_p, C.c_int, u64p, u32p, u64p]),
    "simka_gpu_dump_sample_counts128": (C.c_int, [C.c_void_p, C.c_int, u64p, u64p, u32p, u64p]),
    "simka_gpu_fine_bits": (C.c_int, [C.c_void_p]),
    "simka_gpu_export_sample": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), u64p]),
    "simka_gpu_import_sample": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_uint64]),
    "simka_comm_unique_id": (C.c_int, [C.c_void_p]),
    "simka_gpu_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "simka_gpu_exchange_and_merge": (C.c_int, [C.c_void_p]),
    "simka_gpu_exchange_timings": (C.c_int, [C.c_void_p, f32p]),
    "simka_gpu_import_counts": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64]),
    "simka_gpu_get_totals": (C.c_int, [C.c_void_p, C.POINTER(SampleTotals)]),
    "simka_gpu_set_totals": (C.c_int, [C.c_void_p, C.POINTER(SampleTotals)]),
    "simka_gpu_merge_and_accumulate": (C.c_int, [C.c_void_p]),
    "simka_gpu_get_stats": (C.c_int, [C.c_void_p, C.POINTER(StatsView)]),
    "simka_gpu_get_timings": (C.c_int, [C.c_void_p, C.POINTER(Timings)]),
    "simka_gpu_stream": (C.c_void_p, [C.c_void_p]),
    "simka_pack_words": (C.c_uint64, [C.c_uint64, C.c_uint64]),
    "simka_pack_reads": (C.c_uint64, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]),
    "simka_pack_reads_k": (C.c_uint64, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "simka_nb_matrices": (C.c_int, [C.POINTER(StatsView)]),
    "simka_matrix_name": (C.c_char_p, [C.POINTER(StatsView), C.c_int]),
    "simka_distance_matrix": (C.c_int, [C.POINTER(StatsView), C.c_int, f32p]),
    "simka_matrix_csv": (C.c_size_t, [C.POINTER(StatsView), C.c_int, C.POINTER(C.c_char_p), C.c_char_p, C.c_size_t]),
    "simka_write_matrices": (C.c_int, [C.POINTER(StatsView), C.POINTER(C.c_char_p), C.c_char_p]),
    "simkamin_gpu_sketch": (C.c_int, [C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, u64p, u32p,
                            C.POINTER(C.c_uint32)]),
    "simkamin_gpu_sketch_filter": (C.c_int, [C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, u64p, u32p,
                            C.POINTER(C.c_uint32)]),
    "simkamin_gpu_last_sketch_info": (C.c_int, [C.POINTER(SketchInfo)]),
    "simkamin_gpu_release_cache": (None, []),
    "simkamin_gpu_distance": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, f32p, f32p, f32p]),
    "simka_host_kmers": (C.c_uint64, [C.c_char_p, C.c_uint64, C.c_int, C.c_int, u64p, C.POINTER(C.c_uint8)]),
    "simka_host_minimizer": (C.c_uint64, [C.c_uint64, C.c_int, C.c_int]),
    "simka_host_shannon_index": (C.c_float, [C.c_char_p, C.c_uint64]),
    "simka_host_repartition": (C.c_int, [C.c_char_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]),
    "simka_host_estimate_nb_items": (C.c_int64, [C.c_char_p]),
    "simka_host_read_bank": (C.c_int64, [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "simka_host_free": (None, [C.c_void_p]),
    "simka_minikc_replay": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "simkamin_short_read_replay": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "simkamin_gpu_release": (C.c_int, []),
    "simkamin_gpu_distance_rect": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, f32p, f32p, f32p]),
    "simka_b200_version": (C.c_char_p, []),
}

_lib = None


def load() -> C.CDLL:
    """Loads the CUDA extension.  There is no fallback: a missing library is a hard error."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m simka_b200.build` "
                "(simka_b200 has no CPU or PyTorch fallback)")
        L = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
