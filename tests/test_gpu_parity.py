"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _same(gpu_sketch, ora_sketch, p):
    comps = ora_sketch.components(p)
    assert len(comps) == len(gpu_sketch.codes)
    for c, (codes, counts) in enumerate(comps):
        assert gpu_sketch.codes[c].size == codes.size, "component %d: %d vs %d codes" % (c, gpu_sketch.codes[c].size, codes.size)
        assert np.array_equal(gpu_sketch.codes[c], codes), "component %d codes/order differ" % c
        if counts is not None:
            assert np.array_equal(gpu_sketch.counts[c], counts), "component %d counts differ" % c


@pytest.mark.parametrize("k,subk,L,nreads", [(11, 6, 3, 20000), (11, 5, 2, 20000), (10, 6, 3, 8000), (9, 4, 1, 3000)])
def test_fastq_koc_matches_oracle(oracle, lib_built, shuf, k, subk, L, nreads):
    sid, perm = shuf(1234 + k * 100 + subk, k, subk, L)
    p = oracle.params(k, subk, L)
    S = oracle.synth(42, 20, 200000, 150)
    fq = S.fastq(0, nreads)
    want = oracle.fastq_koc(p, perm, fq)
    assert want.status == 0
    with lib_built.Sketcher(perm, k, subk, L) as sk:
        got = sk.fastq_koc_host(fq)
    _same(got, want, p)
