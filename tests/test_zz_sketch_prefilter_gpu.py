"""GPU parity of the bottom-s pre-filter of the SimkaMin sketches (simkamin_sketch_impl, simka_b200/csrc/simka_gpu.cu): only
hashes below a bound are extracted and counted, the bound widening until the sketch fills.  Added after the round's GPU budget
was spent: bit-exact under the kernel emulator (tests/test_emu_kernels.py::test_simkamin_sketch_prefilter_is_exact); the
other sketch tests of tests/test_gpu_parity.py run through it as well, since it is the default."""
import numpy as np
import pytest

import simka_b200
from helpers import nasty_samples, synthetic_samples
from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("filt", [False, True])
def test_sketch_prefilter_equals_whole_range_and_oracle(monkeypatch, filt):
    samples = (synthetic_samples(1, 20000, n_genomes=2, genome_len=2000, err=0.0) +  # ~750 occurrences per k-mer: the bound widens
               synthetic_samples(1, 20000, n_genomes=64, genome_len=20000) + nasty_samples()[:2])
    widened = saved = 0
    for seqs, off in samples:
        for k, s in [(21, 100), (31, 3000), (21, 100000)]:
            oh, oc, on = oracle.simkamin_sketch(seqs, off, k, s, seed=100, filter=filt)
            monkeypatch.setenv("SIMKA_B200_SKETCH_PREFILTER", "0")
            fh, fc, fn = simka_b200.simkamin_sketch(seqs, off, k, s, seed=100, filter=filt)
            full = simka_b200.simkamin_last_sketch_info()
            monkeypatch.delenv("SIMKA_B200_SKETCH_PREFILTER")
            gh, gc, gn = simka_b200.simkamin_sketch(seqs, off, k, s, seed=100, filter=filt)
            info = simka_b200.simkamin_last_sketch_info()
            assert gn == on == fn and full["attempts"] == 1
            for a, b in ((gh, oh), (gc, oc), (fh, oh), (fc, oc)):
                np.testing.assert_array_equal(a, b)
            widened += info["attempts"] > 1
            saved += info["keys_counted"] * 4 < full["keys_counted"]
    assert widened > 0 and saved > 0
