"""GPU: the CUDA path against the committed golden vectors of the reference binary (no oracle in
between): FASTQ -A sketches for every golden case, genome sketches, species coverage lines."""
import os

import numpy as np
import pytest

import golden_cases as G

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz")


def test_fastq_koc_against_reference_vectors(lib_built, shuf):
    gold = np.load(GOLD)
    ctxs = {}
    for name, (k, subk, L, seed), text in G.fastq_cases():
        key = (k, subk, L, seed)
        if key not in ctxs:
            sid, perm = shuf(seed, k, subk, L)
            ctxs[key] = lib_built.Sketcher(perm, k, subk, L)
        got = ctxs[key].fastq_koc_host(np.ascontiguousarray(text))
        assert len(got.codes) == int(gold[name + "/comp_num"][0])
        for c in range(len(got.codes)):
            assert np.array_equal(got.codes[c], gold["%s/combco.%d" % (name, c)]), "%s comp %d codes" % (name, c)
            assert np.array_equal(got.counts[c], gold["%s/abund.%d" % (name, c)]), "%s comp %d counts" % (name, c)
    for s in ctxs.values():
        s.close()


def test_fasta_and_composite_against_reference_vectors(lib_built, shuf):
    gold = np.load(GOLD)
    k, subk, L, seed = G.MDB_PARAMS
    sid, perm = shuf(seed, k, subk, L)
    S = G.mdb_synth()
    with lib_built.Sketcher(perm, k, subk, L) as sk:
        got = sk.fasta_co_host([G.mdb_fasta(S, s) for s in range(G.MDB_SPECIES)])
        for s in range(G.MDB_SPECIES):
            assert np.array_equal(got[s].codes[0], gold["fasta/sp%d" % s]), "genome sketch %d" % s
        order = [int(x) for x in gold["markerdb/order"]]
        codes = np.concatenate([gold["markerdb/sp%d" % s] for s in order])
        index = np.zeros(len(order) + 1, dtype=np.uint64)
        index[1:] = np.cumsum([gold["markerdb/sp%d" % s].size for s in order])
        names = ["%d_sp%d" % (s + 1, s) for s in order]
        q = sk.fastq_koc_host(G.mdb_reads(S))
        stats = sk.composite([(codes, index)], [(q.codes[0], q.counts[0])])
        tsv = lib_built.composite_tsv("Q", names, stats)
        lines = ["\t".join(l.split("\t")[1:]) for l in tsv.splitlines()]
        assert lines == [str(x) for x in gold["composite/lines"]]
