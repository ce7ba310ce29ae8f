#!/usr/bin/env python
"""Golden vectors of round 2 (tests/golden/reference_vectors_r2.npz): the UNMODIFIED reference binary
(oracle/_ref/metakssd) at `-p 1` on the seeded inputs of tests/golden_cases.py —
`dist -Q q -n m` on FASTQ without -A (fastq2co), `dist -u` on FASTA (uniq_fasta2co) and the `set -g / -q / -i`
MarkerDB pipeline with several genomes per species.

    python tests/golden/make_golden_r2.py
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
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, (k, subk, L, seed), text, Q, M in G.fastq_co_cases():
            sid, perm = O.make_shuf(seed, k, subk, L)
            d = os.path.join(tmp, name)
            os.makedirs(d)
            shuf = os.path.join(d, "x.shuf")
            O.write_shuf_file(shuf, sid, k, subk, L, perm)
            fq = os.path.join(d, "reads.fq")
            with open(fq, "wb") as f:
                f.write(bytes(text))
            sd = O.ref_dist(shuf, [fq], os.path.join(d, "out"), abundance=False, p=1, extra=["-Q", str(Q), "-n", str(M)])
            assert not sd.koc and sd.infile_num == 1
            out[name + "/comp_num"] = np.array([sd.comp_num])
            for c in range(sd.comp_num):
                out["%s/combco.%d" % (name, c)] = sd.combco[c]
            print(name, "codes:", sd.all_ctx_ct)
        for name, (k, subk, L, seed), text in G.uniq_cases():
            sid, perm = O.make_shuf(seed, k, subk, L)
            d = os.path.join(tmp, name)
            os.makedirs(d)
            shuf = os.path.join(d, "x.shuf")
            O.write_shuf_file(shuf, sid, k, subk, L, perm)
            fa = os.path.join(d, "g.fasta")
            with open(fa, "wb") as f:
                f.write(bytes(text))
            sd = O.ref_dist(shuf, [fa], os.path.join(d, "out"), abundance=False, p=1, extra=["-u"])
            out[name + "/comp_num"] = np.array([sd.comp_num])
            for c in range(sd.comp_num):
                out["%s/combco.%d" % (name, c)] = sd.combco[c]
            print(name, "codes:", sd.all_ctx_ct)
        if hasattr(G, "set_case"):
            G_set(out, tmp)
    np.savez_compressed(os.path.join(HERE, "reference_vectors_r2.npz"), **out)
    print("wrote reference_vectors_r2.npz with", len(out), "arrays")


def G_set(out, tmp):
    """`set -g` -> `set -q` -> `set -i` with several genomes per species: every intermediate directory"""
    (k, subk, L, seed), genomes, groups = G.set_case()
    sid, perm = O.make_shuf(seed, k, subk, L)
    d = os.path.join(tmp, "setcase")
    os.makedirs(d)
    shuf = os.path.join(d, "x.shuf")
    O.write_shuf_file(shuf, sid, k, subk, L, perm)
    paths = []
    for i, g in enumerate(genomes):
        p = os.path.join(d, "g%d.fasta" % i)
        with open(p, "wb") as f:
            f.write(bytes(g))
        paths.append(p)
    O.ref_build_markerdb(shuf, paths, groups, d, p=1)
    gsk = O.read_sketch_dir(os.path.join(d, "gsk"))
    out["set/gsk_order"] = np.array([int(os.path.basename(n)[1:].split(".")[0]) for n in gsk.names])
    for sub in ("pan", "markerdb"):
        sd = O.read_sketch_dir(os.path.join(d, sub))
        out["set/%s/names" % sub] = np.array(sd.names)
        out["set/%s/comp_num" % sub] = np.array([sd.comp_num])
        for c in range(sd.comp_num):
            out["set/%s/combco.%d" % (sub, c)] = sd.combco[c]
            out["set/%s/index.%d" % (sub, c)] = sd.index[c]
    for c in range(gsk.comp_num):
        out["set/uniq_pan.%d" % c] = np.fromfile(os.path.join(d, "union_sp", "uniq_pan.%d" % c), dtype=np.uint32)
    print("set case:", len(genomes), "genomes")


if __name__ == "__main__":
    main()
