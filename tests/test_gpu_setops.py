"""GPU: `set -g / -q / -i` on the device (mk_set_group / mk_set_uniq_union / mk_set_operate) against the reference
binary's golden directories (pan, union_sp, markerdb) and, on randomised sketches, against the oracle; then the
whole MarkerDB pipeline through the C host program, byte for byte against the reference binary's directories."""
import os
import subprocess

import numpy as np
import pytest

import golden_cases as G
from test_oracle_golden_r2 import _set_inputs

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_vectors_r2.npz")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_set_pipeline_against_reference_vectors(lib_built, oracle, shuf):
    gold = np.load(GOLD)
    codes, index, taxids, names = _set_inputs(oracle, gold)
    taxon_of, ids = oracle.organize_taxa(taxids)
    sid, perm = shuf(1234, 11, 6, 3)
    with lib_built.Sketcher(perm, 11, 6, 3) as sk:
        pc, pi = sk.set_group(codes, index, taxon_of, len(ids))
        assert np.array_equal(pc, gold["set/pan/combco.0"]) and np.array_equal(pi, gold["set/pan/index.0"])
        u = sk.set_uniq_union(pc)
        assert np.array_equal(u, gold["set/uniq_pan.0"])
        mc, mi = sk.set_operate(u, pc, pi, True)
        assert np.array_equal(mc, gold["set/markerdb/combco.0"]) and np.array_equal(mi, gold["set/markerdb/index.0"])


def test_set_ops_random_against_oracle(lib_built, oracle, shuf):
    rng = np.random.default_rng(3)
    sid, perm = shuf(1234, 11, 6, 3)
    with lib_built.Sketcher(perm, 11, 6, 3) as sk:
        for case in range(12):
            n_gen = int(rng.integers(1, 60))
            n_tax = int(rng.integers(1, max(2, n_gen)))
            space = int(rng.choice([50, 5000, 2 ** 31]))
            sks = [rng.integers(0, space, size=int(rng.integers(0, 3000))).astype(np.uint32) for _ in range(n_gen)]
            codes = np.concatenate(sks) if sks else np.empty(0, np.uint32)
            index = np.zeros(n_gen + 1, np.uint64)
            index[1:] = np.cumsum([s.size for s in sks])
            taxon_of = rng.integers(-1, n_tax, size=n_gen).astype(np.int32)
            wc, wi = oracle.set_group(codes, index, taxon_of, n_tax)
            gc, gi = sk.set_group(codes, index, taxon_of, n_tax)
            assert np.array_equal(gi, wi) and np.array_equal(gc, wc), "group case %d" % case
            wu = oracle.set_uniq_union(wc)
            assert np.array_equal(sk.set_uniq_union(wc), wu), "uniq case %d" % case
            for inter in (True, False):
                oc, oi = oracle.set_operate(wu, wc, wi, inter)
                dc, di = sk.set_operate(wu, wc, wi, inter)
                assert np.array_equal(di, oi) and np.array_equal(dc, oc), "operate case %d" % case


def _stat_fields(path):
    raw = bytearray(open(os.path.join(path, "cofiles.stat"), "rb").read())
    raw[5:8] = b"\0\0\0"                         # padding after `bool koc`
    return raw


def test_cli_set_pipeline_byte_identical_to_reference(lib_built, oracle, tmp_path):
    """dist (reference) -> [ours | reference] set -g -> set -q -> set -i on the same genome sketch directory and
    taxfile: pan/, union_sp/, markerdb/ byte for byte (names as C strings, stat padding masked)"""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/metakssd is not built")
    subprocess.run(["make", "-C", os.path.join(ROOT, "host")], check=True, capture_output=True)
    cli = os.path.join(ROOT, "host", "metakssd-b200")
    (k, subk, L, seed), genomes, groups = G.set_case()
    sid, perm = oracle.make_shuf(seed, k, subk, L)
    d = str(tmp_path)
    shuf = os.path.join(d, "x.shuf")
    oracle.write_shuf_file(shuf, sid, k, subk, L, perm)
    paths = []
    for i, g in enumerate(genomes):
        p = os.path.join(d, "g%d.fasta" % i)
        with open(p, "wb") as f:
            f.write(bytes(g))
        paths.append(p)
    refdir = os.path.join(d, "ref")
    os.makedirs(refdir)
    oracle.ref_build_markerdb(shuf, paths, groups, refdir, p=1)           # writes gsk, group_name.txt, pan, union_sp, markerdb
    gsk, grp = os.path.join(refdir, "gsk"), os.path.join(refdir, "group_name.txt")
    ours = os.path.join(d, "ours")
    os.makedirs(ours)
    run = lambda *a: subprocess.run([cli] + list(a), check=True, capture_output=True, timeout=300)
    run("set", "-g", grp, "-o", os.path.join(ours, "pan"), gsk)
    run("set", "-q", "-o", os.path.join(ours, "union_sp"), os.path.join(ours, "pan"))
    run("set", "-i", os.path.join(ours, "union_sp"), "-o", os.path.join(ours, "markerdb"), os.path.join(ours, "pan"))
    rd = lambda *a: open(os.path.join(*a), "rb").read()
    for sub, files in (("pan", ["combco.0", "combco.index.0"]), ("union_sp", ["uniq_pan.0"]), ("markerdb", ["combco.0", "combco.index.0"])):
        for fn in files:
            assert rd(refdir, sub, fn) == rd(ours, sub, fn), "%s/%s" % (sub, fn)
    for sub in ("pan", "markerdb"):
        a, b = _stat_fields(os.path.join(refdir, sub)), _stat_fields(os.path.join(ours, sub))
        n = int.from_bytes(a[20:24], "little")
        assert len(a) == len(b) == 32 + 4 * n + 256 * n
        assert a[:32 + 4 * n] == b[:32 + 4 * n], sub
        for i in range(n):
            ra, rb = a[32 + 4 * n + 256 * i:][:256], b[32 + 4 * n + 256 * i:][:256]
            assert ra.split(b"\0", 1)[0] == rb.split(b"\0", 1)[0]
    assert _stat_fields(os.path.join(refdir, "union_sp"))[:32] == _stat_fields(os.path.join(ours, "union_sp"))[:32]
    # and the MarkerDB our pipeline wrote serves the reference's composite
    fq = os.path.join(d, "r.fq")
    S = oracle.synth(79, 10, 180_000, 150)
    S.fastq(0, 60_000).tofile(fq)
    oracle.ref_dist(shuf, [fq], os.path.join(d, "qry"), abundance=True, p=1)
    t_ref = oracle.ref_composite(os.path.join(refdir, "markerdb"), os.path.join(d, "qry"))
    t_our = oracle.ref_composite(os.path.join(ours, "markerdb"), os.path.join(d, "qry"))
    assert t_ref == t_our and t_ref.count("\n") >= 3
