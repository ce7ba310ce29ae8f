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
