#!/usr/bin/env python
"""Golden vectors of `dist -r` (tests/golden/reference_vectors_r2b.npz): the UNMODIFIED reference binary
(oracle/_ref/metakssd, -p 1) on the seeded inputs of tests/golden_cases.py::dist_search_case() —
`dist -L <shuf> -o ref <17 genomes>`, `dist -L <shuf> -o qry <12 genomes>`, `dist -L <shuf> -A -o qryA reads.fq`, then
`dist -r ref -o out [options] qry` for every option set of DIST_SEARCH_OPTIONS (distance.out as text) and once with
--keepskf (sharedk_ct.dat, the shared k-mer count matrix).

The first `dist -r` makes the reference write its inverted index into the ref directory: mco.index.0 is
2^32 x 8 bytes = 32 GiB (co2mco.c:17-67), about two minutes and 32 GiB of disk in $TMPDIR.

    python tests/golden/make_golden_r2b.py
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle as O  # noqa: E402
import golden_cases as G  # noqa: E402


def main():
    O.build()
    assert O.have_ref(), "oracle/_ref/metakssd is missing (needs /root/reference)"
    (k, subk, L, seed), named, ref_names, qry_names, reads = G.dist_search_case()
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        sid, perm = O.make_shuf(seed, k, subk, L)
        shuf = os.path.join(tmp, "x.shuf")
        O.write_shuf_file(shuf, sid, k, subk, L, perm)
        cwd = os.getcwd()
        os.chdir(tmp)                      # short relative names end up in the sketch directories
        for n, g in named.items():
            with open(n, "wb") as f:
                f.write(g)
        with open("reads.fq", "wb") as f:
            f.write(reads)
        ref = O.ref_dist(shuf, ref_names, "ref", abundance=False, p=1)
        qry = O.ref_dist(shuf, qry_names, "qry", abundance=False, p=1)
        qryA = O.ref_dist(shuf, ["reads.fq"], "qryA", abundance=True, p=1)
        out["ref/names"] = np.array(ref.names)
        out["qry/names"] = np.array(qry.names)
        out["ref/ctx_ct"] = np.asarray(ref.ctx_ct, dtype=np.uint32)
        out["qry/ctx_ct"] = np.asarray(qry.ctx_ct, dtype=np.uint32)
        out["qryA/ctx_ct"] = np.asarray(qryA.ctx_ct, dtype=np.uint32)
        out["qryA/combco.0"] = qryA.combco[0]
        for name, opts in G.DIST_SEARCH_OPTIONS.items():
            out["out/" + name] = np.array(O.ref_dist_search("ref", "qry", "out_" + name, extra=opts))
            print(name, out["out/" + name].item().count("\n"), "lines")
        O.ref_dist_search("ref", "qry", "out_keep", extra=["--keepskf"])
        out["sharedk_ct"] = np.fromfile("out_keep/sharedk_ct.dat", dtype=np.uint32).reshape(len(qry_names), len(ref_names))
        out["outA/default"] = np.array(O.ref_dist_search("ref", "qryA", "outA"))
        O.ref_dist_search("ref", "qryA", "outA_keep", extra=["--keepskf"])
        out["sharedk_ct_A"] = np.fromfile("outA_keep/sharedk_ct.dat", dtype=np.uint32).reshape(1, len(ref_names))
        # edge: sketches without a single code on both sides (Jaccard 0/0, containment x/0 ...)
        with open("tiny_a.fasta", "wb") as f:
            f.write(b">t\nACGTACGTAC\n")
        with open("tiny_b.fasta", "wb") as f:
            f.write(b">t\nTTTTGGGGCCCCAAAA\n")
        eref = O.ref_dist(shuf, ref_names[:3] + ["tiny_a.fasta"], "eref", abundance=False, p=1)
        eqry = O.ref_dist(shuf, [ref_names[0], "tiny_b.fasta", qry_names[-1]], "eqry", abundance=False, p=1)
        out["edge/ref/names"] = np.array(eref.names)
        out["edge/qry/names"] = np.array(eqry.names)
        out["edge/ref/ctx_ct"] = np.asarray(eref.ctx_ct, dtype=np.uint32)
        out["edge/qry/ctx_ct"] = np.asarray(eqry.ctx_ct, dtype=np.uint32)
        for name in ("default", "containment", "nearest3", "corrected"):
            out["edge/out/" + name] = np.array(O.ref_dist_search("eref", "eqry", "eout_" + name, extra=G.DIST_SEARCH_OPTIONS[name]))
        O.ref_dist_search("eref", "eqry", "eout_keep", extra=["--keepskf"])
        out["edge/sharedk_ct"] = np.fromfile("eout_keep/sharedk_ct.dat", dtype=np.uint32).reshape(3, 4)
        os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "reference_vectors_r2b.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
