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


def _set_inputs(oracle, gold):
    (k, subk, L, seed), genomes, groups = G.set_case()
    sid, perm = oracle.make_shuf(seed, k, subk, L)
    p = oracle.params(k, subk, L)
    order = [int(x) for x in gold["set/gsk_order"]]          # the reference lists genomes in a time-seeded order
    sk = [oracle.fasta_co(p, perm, genomes[g]).components(p)[0][0] for g in order]
    codes = np.concatenate(sk)
    index = np.zeros(len(sk) + 1, np.uint64)
    index[1:] = np.cumsum([s.size for s in sk])
    taxids = [int(groups[g].split("\t")[0]) for g in order]
    names = {int(groups[g].split("\t")[0]): groups[g].split("\t")[1] for g in order}
    return codes, index, taxids, names


def test_set_pipeline_golden(oracle, gold):
    """organize_taxf() order, grouping_genomes(), uniq_sketch_union(), sketch_operate() against the reference's
    pan / union_sp / markerdb directories"""
    codes, index, taxids, names = _set_inputs(oracle, gold)
    taxon_of, ids = oracle.organize_taxa(taxids)
    assert ["%d_%s" % (t, names[t]) for t in ids] == [str(x) for x in gold["set/pan/names"]]
    pc, pi = oracle.set_group(codes, index, taxon_of, len(ids))
    assert np.array_equal(pc, gold["set/pan/combco.0"]) and np.array_equal(pi, gold["set/pan/index.0"])
    u = oracle.set_uniq_union(pc)
    assert np.array_equal(u, gold["set/uniq_pan.0"]) and 0 < u.size < pc.size
    mc, mi = oracle.set_operate(u, pc, pi, True)
    assert np.array_equal(mc, gold["set/markerdb/combco.0"]) and np.array_equal(mi, gold["set/markerdb/index.0"])
    sc, si = oracle.set_operate(u, pc, pi, False)
    assert sc.size + mc.size == pc.size


# ---- `dist -r`: shared k-mer counts and the distance table (tests/golden/reference_vectors_r2b.npz) ----------------
GOLD_B = os.path.join(os.path.dirname(__file__), "golden", "reference_vectors_r2b.npz")


def _opts(flags):
    o = dict(metric=0, outfields=2, correction=0, n_max=0, max_dist=1.0)
    key = {"-M": "metric", "-O": "outfields", "-N": "n_max", "--correction": "correction"}
    for a, b in zip(flags[::2], flags[1::2]):
        if a == "-D":
            o["max_dist"] = float(b)
        else:
            o[key[a]] = int(b)
    return o


def test_dist_search_golden(oracle):
    """ko_shared_counts / ko_distance_out against `metakssd dist -r ref qry` of the reference binary: the count matrix
    it keeps with --keepskf and distance.out for eight option sets, byte for byte."""
    from helpers import dist_search_world
    gold = np.load(GOLD_B)
    p, perm, ref_names, ref, ref_ct, qry_names, qry, qry_ct = dist_search_world(oracle, gold)
    assert np.array_equal(ref_ct, gold["ref/ctx_ct"]) and np.array_equal(qry_ct, gold["qry/ctx_ct"])
    counts = oracle.shared_counts([ref], [qry], qry_ct)
    assert np.array_equal(counts, gold["sharedk_ct"])
    assert counts.max() > 30 and (counts > 0).sum() > 100
    for name, flags in G.DIST_SEARCH_OPTIONS.items():
        txt = oracle.distance_out(counts, ref_ct, qry_ct, ref_names, qry_names, 2 * p.k, 2 * p.drlevel, **_opts(flags))
        assert txt == str(gold["out/" + name]), name
    # a read sample sketched with -A as the query (counts ignored, codes as they are on disk)
    (k, subk, L, seed), _, _, _, reads = G.dist_search_case()
    sk = oracle.fastq_koc(p, perm, np.frombuffer(reads, np.uint8)).components(p)[0][0]
    assert np.array_equal(sk, gold["qryA/combco.0"])
    qa = (sk.astype(np.uint32), np.array([0, sk.size], dtype=np.uint64))
    ca = oracle.shared_counts([ref], [qa], gold["qryA/ctx_ct"])
    assert np.array_equal(ca, gold["sharedk_ct_A"])
    txt = oracle.distance_out(ca, ref_ct, gold["qryA/ctx_ct"], ref_names, ["reads.fq"], 2 * p.k, 2 * p.drlevel)
    assert txt == str(gold["outA/default"])


def test_dist_search_golden_empty_sketches(oracle):
    """Sketches without a code on either side: 0/0 and x/0 go through the same libm calls and printf formats as in the
    reference (`-nan`, `-NAN`, `inf`), and such lines are printed (NaN > max_dist is false)."""
    from helpers import dist_search_edge_world
    gold = np.load(GOLD_B)
    p, perm, ref_names, ref, ref_ct, qry_names, qry, qry_ct = dist_search_edge_world(oracle, gold)
    assert np.array_equal(ref_ct, gold["edge/ref/ctx_ct"]) and np.array_equal(qry_ct, gold["edge/qry/ctx_ct"])
    assert 0 in ref_ct and 0 in qry_ct
    counts = oracle.shared_counts([ref], [qry], qry_ct)
    assert np.array_equal(counts, gold["edge/sharedk_ct"])
    for name in ("default", "containment", "nearest3", "corrected"):
        txt = oracle.distance_out(counts, ref_ct, qry_ct, ref_names, qry_names, 2 * p.k, 2 * p.drlevel, **_opts(G.DIST_SEARCH_OPTIONS[name]))
        assert txt == str(gold["edge/out/" + name]), name
        assert name == "nearest3" or "nan" in txt          # (the N-nearest selection never picks a NaN metric)
