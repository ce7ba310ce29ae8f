"""CPU: the oracle's restatements of fastq2co() (`dist -Q q -n m` on FASTQ without -A) and uniq_fasta2co()
(`dist -u`) against the vectors the unmodified reference binary produced at -p 1
(tests/golden/reference_vectors_r2.npz, tests/golden/make_golden_r2.py).  Bit-exact, on-disk order."""
import os

import numpy as np
import pytest

import golden_cases as G

GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_vectors_r2.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _same(gold, name, p, sketch):
    comps = sketch.components(p)
    assert int(gold[name + "/comp_num"][0]) == len(comps)
    for c, (codes, _) in enumerate(comps):
        assert np.array_equal(codes, gold["%s/combco.%d" % (name, c)]), "%s component %d" % (name, c)


def test_fastq_co_golden(oracle, gold):
    n = 0
    for name, (k, subk, L, seed), text, Q, M in G.fastq_co_cases():
        sid, perm = oracle.make_shuf(seed, k, subk, L)
        p = oracle.params(k, subk, L)
        sk = oracle.fastq_co(p, perm, text, Q, M)
        assert sk.status == 0
        _same(gold, name, p, sk)
        n += 1
    assert n >= 17


def test_fasta_uniq_golden(oracle, gold):
    n = 0
    for name, (k, subk, L, seed), text in G.uniq_cases():
        sid, perm = oracle.make_shuf(seed, k, subk, L)
        p = oracle.params(k, subk, L)
        sk = oracle.fasta_co_uniq(p, perm, text)
        _same(gold, name, p, sk)
        if "plain" not in name:     # the duplicated part really removes codes
            assert sk.codes.size < oracle.fasta_co(p, perm, text).codes.size
        n += 1
    assert n >= 5
