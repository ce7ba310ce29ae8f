"""CPU: the oracle (oracle/kssd_oracle.c) against the golden vectors produced by the unmodified
reference binary at -p 1 (tests/golden/reference_vectors.npz, see make_golden.py).  Bit-exact:
codes, counts, on-disk order, per-component split, MarkerDB, species coverage lines."""
import os

import numpy as np
import pytest

import golden_cases as G

GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _check(gold, name, p, sketch):
    comps = sketch.components(p)
    assert int(gold[name + "/comp_num"][0]) == len(comps) == p.component_num
    kmerlen, dim_rd_len, all_ct, ct0 = [int(x) for x in gold[name + "/hdr"]]
    assert kmerlen == 2 * p.k and dim_rd_len == 2 * p.drlevel
    assert all_ct == ct0 == sketch.codes.size
    for c, (codes, counts) in enumerate(comps):
        assert np.array_equal(codes, gold["%s/combco.%d" % (name, c)]), "%s component %d codes/order" % (name, c)
        assert np.array_equal(counts, gold["%s/abund.%d" % (name, c)]), "%s component %d counts" % (name, c)


def test_fastq_koc_golden(oracle, gold):
    n = 0
    for name, (k, subk, L, seed), text in G.fastq_cases():
        sid, perm = oracle.make_shuf(seed, k, subk, L)
        p = oracle.params(k, subk, L)
        sk = oracle.fastq_koc(p, perm, text)
        assert sk.status == 0
        _check(gold, name, p, sk)
        n += 1
    assert n >= 17


def test_fasta_markerdb_composite_golden(oracle, gold):
    from helpers import markerdb_from_sketches
    k, subk, L, seed = G.MDB_PARAMS
    sid, perm = oracle.make_shuf(seed, k, subk, L)
    p = oracle.params(k, subk, L)
    S = G.mdb_synth()
    sketches = []
    for s in range(G.MDB_SPECIES):
        sk = oracle.fasta_co(p, perm, G.mdb_fasta(S, s))
        codes = sk.components(p)[0][0]
        assert np.array_equal(codes, gold["fasta/sp%d" % s]), "genome sketch %d" % s
        sketches.append(codes)
    # the reference's grouping_genomes() re-hashes every species' codes into its own table, so the
    # MarkerDB order inside a species differs from the genome sketch order: compare as sets per
    # species (composite does not depend on that order) ...
    ref_codes, ref_index = markerdb_from_sketches(sketches)
    for s in range(G.MDB_SPECIES):
        mine = ref_codes[int(ref_index[s]):int(ref_index[s + 1])]
        assert np.array_equal(np.sort(mine), np.sort(gold["markerdb/sp%d" % s])), "markers of species %d" % s
    # ... and the species coverage lines byte for byte, with species in the MarkerDB's own order
    order = [int(x) for x in gold["markerdb/order"]]
    codes = np.concatenate([gold["markerdb/sp%d" % s] for s in order])
    index = np.zeros(len(order) + 1, dtype=np.uint64)
    index[1:] = np.cumsum([gold["markerdb/sp%d" % s].size for s in order])
    names = ["%d_sp%d" % (s + 1, s) for s in order]
    q = oracle.fastq_koc(p, perm, G.mdb_reads(S))
    tsv = oracle.composite([(codes, index)], names, [q.components(p)[0]], "Q")
    got = ["\t".join(l.split("\t")[1:]) for l in tsv.splitlines()]
    assert got == [str(x) for x in gold["composite/lines"]] and len(got) >= 5
