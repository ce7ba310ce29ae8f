#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference binary
(oracle/_ref/metakssd, built from /root/reference by oracle/Makefile) at `-p 1` on seeded synthetic
inputs.  Only runs where /root/reference (or a prebuilt oracle/_ref) exists; the vectors are
committed so that the CPU test-suite and the GPU box can check the oracle / the CUDA path without
the reference.

    python tests/golden/make_golden.py

Inputs are NOT stored: tests regenerate them from the same seeds (tests/golden_cases.py).
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
        # ---- FASTQ -A sketches -------------------------------------------------------------------
        for name, (k, subk, L, shuf_seed), text in G.fastq_cases():
            sid, perm = O.make_shuf(shuf_seed, k, subk, L)
            d = os.path.join(tmp, name)
            os.makedirs(d)
            shuf = os.path.join(d, "x.shuf")
            O.write_shuf_file(shuf, sid, k, subk, L, perm)
            fq = os.path.join(d, "reads.fq")
            with open(fq, "wb") as f:
                f.write(bytes(text))
            sd = O.ref_dist(shuf, [fq], os.path.join(d, "out"), abundance=True, p=1)
            assert sd.koc and sd.infile_num == 1 and sd.shuf_id == sid & 0xFFFFFFFF
            out[name + "/comp_num"] = np.array([sd.comp_num])
            out[name + "/hdr"] = np.array([sd.kmerlen, sd.dim_rd_len, sd.all_ctx_ct, int(sd.ctx_ct[0])], dtype=np.int64)
            for c in range(sd.comp_num):
                out["%s/combco.%d" % (name, c)] = sd.combco[c]
                out["%s/abund.%d" % (name, c)] = sd.abund[c]
            print(name, "codes:", sd.all_ctx_ct)
        # ---- FASTA sketches + MarkerDB + composite -----------------------------------------------
        k, subk, L, shuf_seed = G.MDB_PARAMS
        sid, perm = O.make_shuf(shuf_seed, k, subk, L)
        d = os.path.join(tmp, "mdb")
        os.makedirs(d)
        shuf = os.path.join(d, "x.shuf")
        O.write_shuf_file(shuf, sid, k, subk, L, perm)
        S = G.mdb_synth()
        paths, groups = [], []
        gdir = os.path.join(d, "genomes")
        os.makedirs(gdir)
        for s in range(G.MDB_SPECIES):
            p = os.path.join(gdir, "sp%d.fasta" % s)
            with open(p, "wb") as f:
                f.write(bytes(G.mdb_fasta(S, s)))
            paths.append(p)
            groups.append("%d\tsp%d" % (s + 1, s))
        mdb = O.ref_build_markerdb(shuf, paths, groups, d, p=1)
        gsk = O.read_sketch_dir(os.path.join(d, "gsk"))
        for i, n in enumerate(gsk.names):       # genome sketches, keyed by species (file order is time-seeded)
            s = int(os.path.basename(n)[2:].split(".")[0])
            lo, hi = int(gsk.index[0][i]), int(gsk.index[0][i + 1])
            out["fasta/sp%d" % s] = gsk.combco[0][lo:hi]
        md = O.read_sketch_dir(mdb)
        assert md.comp_num == 1
        for i, n in enumerate(md.names):
            s = int(n.split("_sp")[1])
            lo, hi = int(md.index[0][i]), int(md.index[0][i + 1])
            out["markerdb/sp%d" % s] = md.combco[0][lo:hi]
        out["markerdb/order"] = np.array([int(n.split("_sp")[1]) for n in md.names])
        fq = os.path.join(d, "reads.fq")
        with open(fq, "wb") as f:
            f.write(bytes(G.mdb_reads(S)))
        O.ref_dist(shuf, [fq], os.path.join(d, "qry"), abundance=True, p=1)
        tsv = O.ref_composite(mdb, os.path.join(d, "qry"), p=1)
        lines = [l for l in tsv.splitlines() if "\t" in l]
        # the query name is the path given on the command line: keep only the columns after it
        out["composite/lines"] = np.array(["\t".join(l.split("\t")[1:]) for l in lines])
        print("composite lines:", len(lines))
    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_vectors.npz"), "with", len(out), "arrays")


if __name__ == "__main__":
    main()
