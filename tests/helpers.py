"""Shared helpers for the test-suite (test infrastructure; may use oracle/)."""
import numpy as np


def same_sketch(gpu_sketch, ora_sketch, p):
    comps = ora_sketch.components(p)
    assert len(comps) == len(gpu_sketch.codes)
    for c, (codes, counts) in enumerate(comps):
        assert gpu_sketch.codes[c].size == codes.size, \
            "component %d: %d vs %d codes" % (c, gpu_sketch.codes[c].size, codes.size)
        assert np.array_equal(gpu_sketch.codes[c], codes), "component %d codes/order differ" % c
        if counts is not None:
            assert np.array_equal(gpu_sketch.counts[c], counts), "component %d counts differ" % c


def markerdb_from_sketches(species_codes):
    """The `set -g / -q / -i` result for one genome per species (command_set.c:831,427,322):
    every species keeps the codes no other species has.  species_codes: list of uint32 arrays
    (one component).  Returns (codes uint32[], index uint64[S+1])."""
    allc = np.concatenate(species_codes) if species_codes else np.empty(0, np.uint32)
    uniq, cnt = np.unique(allc, return_counts=True)
    solo = set(uniq[cnt == 1].tolist())
    out, index = [], [0]
    for codes in species_codes:
        keep = np.array([c for c in codes.tolist() if c in solo], dtype=np.uint32)
        out.append(keep)
        index.append(index[-1] + keep.size)
    return (np.concatenate(out) if out else np.empty(0, np.uint32)), np.array(index, dtype=np.uint64)


def dist_search_world(oracle, gold):
    """The sketches of golden_cases.dist_search_case() in the file order the reference gave its directories
    (tests/golden/reference_vectors_r2b.npz): (p, perm, ref names, ref (codes, index), ref ctx_ct, qry names, qry (codes, index),
    qry ctx_ct), one component (L3K11)."""
    import golden_cases as G
    import numpy as np
    (k, subk, L, seed), named, _, _, _ = G.dist_search_case()
    p = oracle.params(k, subk, L)
    _, perm = oracle.make_shuf(seed, k, subk, L)

    def side(names):
        sk = [oracle.fasta_co(p, perm, named[n]).components(p)[0][0] for n in names]
        index = np.zeros(len(sk) + 1, dtype=np.uint64)
        index[1:] = np.cumsum([s.size for s in sk])
        return (np.concatenate(sk).astype(np.uint32), index), np.array([s.size for s in sk], dtype=np.uint32)

    ref_names = [str(n) for n in gold["ref/names"]]
    qry_names = [str(n) for n in gold["qry/names"]]
    ref, ref_ct = side(ref_names)
    qry, qry_ct = side(qry_names)
    return p, perm, ref_names, ref, ref_ct, qry_names, qry, qry_ct


def dist_search_edge_world(oracle, gold):
    """The edge case of the `dist -r` vectors: a reference and a query sketch without a single code."""
    import golden_cases as G
    import numpy as np
    (k, subk, L, seed), named, _, _, _ = G.dist_search_case()
    named = dict(named)
    named["tiny_a.fasta"] = b">t\nACGTACGTAC\n"
    named["tiny_b.fasta"] = b">t\nTTTTGGGGCCCCAAAA\n"
    p = oracle.params(k, subk, L)
    _, perm = oracle.make_shuf(seed, k, subk, L)

    def side(names):
        sk = [oracle.fasta_co(p, perm, named[n]).components(p)[0][0] for n in names]
        index = np.zeros(len(sk) + 1, dtype=np.uint64)
        index[1:] = np.cumsum([s.size for s in sk])
        return (np.concatenate(sk).astype(np.uint32), index), np.array([s.size for s in sk], dtype=np.uint32)

    ref_names = [str(n) for n in gold["edge/ref/names"]]
    qry_names = [str(n) for n in gold["edge/qry/names"]]
    ref, ref_ct = side(ref_names)
    qry, qry_ct = side(qry_names)
    return p, perm, ref_names, ref, ref_ct, qry_names, qry, qry_ct
