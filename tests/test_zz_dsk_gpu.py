"""gatb-core's DSK known answers (test/unit/src/kmer/TestDSK.cpp:147-240) through the CUDA counting path."""
import numpy as np
import pytest

import simka_b200
from helpers import reads_to_arrays
from test_oracle_golden import dsk_known_answers

pytestmark = pytest.mark.gpu


def test_dsk_solid_counts_known_answers_gpu(golden_dir):
    """stage 1 (extract -> partition -> bin count -> abundance filter) on the literal banks of the reference's DSK test: the
    number of solid k-mers per row is the reference's; the bank is given twice (two samples) so that the merge runs too and
    reports every solid k-mer as shared"""
    for seqs, k, amin, want in dsk_known_answers(golden_dir):
        sample = reads_to_arrays([s.encode() for s in seqs])
        st, ctx = simka_b200.run_samples([sample, sample], k, abundance_min=amin, keep_ctx=True)
        for s in (0, 1):
            km, ct = ctx.dump_sample_counts(s)
            assert len(km) == want, (k, amin, s)
            assert (ct >= amin).all()
            assert (km[1:] > km[:-1]).all()  # ascending, distinct
        ctx.close()
        assert int(st.distinct_shared[1]) == want  # pair (0, 1) of the packed upper triangle: j + (N-1) i - i (i-1) / 2
        if want:
            assert st.matrix("mat_abundance_braycurtis")[0, 1] == 0.0
