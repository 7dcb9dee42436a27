"""simka_b200 -- B200-native (sm_100a) implementation of the Simka hot path.

reads -> canonical k-mer extraction + bucket partition -> per-bin exact counting -> multi-sample
merge -> N x N ecological-distance accumulators, behind the C ABI of include/simka_b200.h.
The compute lives in lib/libsimka_b200.so (hand-written CUDA); this package is the ctypes front end
and fails loudly if that library is missing -- there is no CPU path.
"""
from .api import (SimkaError, SimkaGpu, SimkaStatistics, comm_unique_id, pack_reads, pinned_empty, run_samples,  # noqa: F401
                  simkamin_distance, simkamin_last_sketch_info, simkamin_release, simkamin_sketch)

__all__ = ["SimkaError", "SimkaGpu", "SimkaStatistics", "comm_unique_id", "pack_reads", "pinned_empty", "run_samples",
           "simkamin_distance", "simkamin_last_sketch_info", "simkamin_release", "simkamin_sketch"]
