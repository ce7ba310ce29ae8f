"""GPU: `dist -r <ref> <qry>` — the shared k-mer counts on the device (mk_shared_counts) against the matrix the reference
binary keeps with --keepskf and against the oracle on random and multi-component sketches; then the whole sub-command
through the C host program: distance.out byte for byte against the reference binary's for eight option sets
(tests/golden/reference_vectors_r2b.npz, made by tests/golden/make_golden_r2b.py)."""
import os
import struct
import subprocess

import numpy as np
import pytest

import golden_cases as G
from helpers import dist_search_world

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_vectors_r2b.npz")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "host", "metakssd-b200")


def test_shared_counts_against_the_reference_matrix(lib_built, oracle, shuf):
    gold = np.load(GOLD)
    p, perm, ref_names, ref, ref_ct, qry_names, qry, qry_ct = dist_search_world(oracle, gold)
    with lib_built.Sketcher(perm, 11, 6, 3) as sk:
        counts = sk.shared_counts([ref], [qry], qry_ct)
        assert np.array_equal(counts, gold["sharedk_ct"])
        qa = (gold["qryA/combco.0"], np.array([0, gold["qryA/combco.0"].size], dtype=np.uint64))
        assert np.array_equal(sk.shared_counts([ref], [qa], gold["qryA/ctx_ct"]), gold["sharedk_ct_A"])


def test_shared_counts_with_empty_sketches(lib_built, oracle):
    from helpers import dist_search_edge_world
    gold = np.load(GOLD)
    p, perm, ref_names, ref, ref_ct, qry_names, qry, qry_ct = dist_search_edge_world(oracle, gold)
    with lib_built.Sketcher(None, 11, 6, 3) as sk:
        assert np.array_equal(sk.shared_counts([ref], [qry], qry_ct), gold["edge/sharedk_ct"])


def test_shared_counts_random_against_oracle(lib_built, oracle):
    """Random sketches: many references sharing codes, empty sketches, a query whose ctx_ct is 0 (skipped like
    command_dist.c:1033), several components accumulating into one matrix; a context without a .shuf is enough."""
    rng = np.random.default_rng(15)
    with lib_built.Sketcher(None, 11, 6, 3) as sk:
        for case in range(8):
            n_ref, n_qry, n_comp = int(rng.integers(1, 400)), int(rng.integers(1, 60)), int(rng.choice([1, 1, 3, 16]))
            space = int(rng.choice([200, 20_000, 2 ** 32]))

            def side(n):
                comps = []
                for _ in range(n_comp):
                    sks = [np.unique(rng.integers(0, space, size=int(rng.integers(0, 4000)), dtype=np.uint64)).astype(np.uint32)
                           for _ in range(n)]
                    idx = np.zeros(n + 1, dtype=np.uint64)
                    idx[1:] = np.cumsum([s.size for s in sks])
                    comps.append((np.concatenate(sks) if idx[-1] else np.zeros(0, np.uint32), idx))
                return comps

            ref, qry = side(n_ref), side(n_qry)
            ct = np.array([sum(int(c[1][q + 1] - c[1][q]) for c in qry) for q in range(n_qry)], dtype=np.uint32)
            if n_qry > 2:
                ct[1] = 0
            want = oracle.shared_counts(ref, qry, ct)
            got = sk.shared_counts(ref, qry, ct)
            assert np.array_equal(got, want), case
            if n_qry > 2:
                assert not got[1].any()
        assert want.sum() > 0


def _write_dirs(lib_built, oracle, tmp, gold, mco_stat=False):
    p, perm, ref_names, ref, ref_ct, qry_names, qry, qry_ct = dist_search_world(oracle, gold)
    sid = oracle.make_shuf(1234, 11, 6, 3)[0]
    info = lib_built.MkInfo()
    info.k, info.drlevel, info.component_num = 11, 3, 1

    def sketches(side):
        codes, idx = side
        return [lib_built.Sketch([codes[int(idx[i]):int(idx[i + 1])]], None) for i in range(idx.size - 1)]

    rd, qd = os.path.join(tmp, "ref"), os.path.join(tmp, "qry")
    lib_built.write_sketch_dir(rd, sid, info, ref_names, sketches(ref), False)
    lib_built.write_sketch_dir(qd, sid, info, qry_names, sketches(qry), False)
    if mco_stat:      # what run_stageII() leaves next to the combco files (command_dist.c:524-541); cofiles.stat gone
        raw = open(os.path.join(rd, "cofiles.stat"), "rb").read()
        with open(os.path.join(rd, "mcofiles.stat"), "wb") as f:
            f.write(struct.pack("<Iiiii", sid & 0xFFFFFFFF, 22, 6, 1, len(ref_names)) + raw[32:])
        os.remove(os.path.join(rd, "cofiles.stat"))
    return rd, qd


@pytest.mark.parametrize("mco_stat", [False, True])
def test_cli_distance_out_is_byte_identical(lib_built, oracle, tmp_path, mco_stat):
    subprocess.run(["make", "-C", os.path.join(ROOT, "host")], check=True, capture_output=True)
    gold = np.load(GOLD)
    rd, qd = _write_dirs(lib_built, oracle, str(tmp_path), gold, mco_stat)
    for name, flags in G.DIST_SEARCH_OPTIONS.items():
        out = os.path.join(str(tmp_path), "out_" + name)
        subprocess.run([CLI, "dist", "-r", rd, "-o", out] + flags + [qd], check=True, capture_output=True, timeout=120)
        assert open(os.path.join(out, "distance.out")).read() == str(gold["out/" + name]), name
        assert not os.path.exists(os.path.join(out, "sharedk_ct.dat"))
    out = os.path.join(str(tmp_path), "out_keep")
    subprocess.run([CLI, "dist", "-r", rd, "-o", out, "--keepskf", qd], check=True, capture_output=True, timeout=120)
    m = np.fromfile(os.path.join(out, "sharedk_ct.dat"), dtype=np.uint32).reshape(gold["sharedk_ct"].shape)
    assert np.array_equal(m, gold["sharedk_ct"])


def test_cli_refuses_what_the_reference_refuses(lib_built, oracle, tmp_path):
    subprocess.run(["make", "-C", os.path.join(ROOT, "host")], check=True, capture_output=True)
    gold = np.load(GOLD)
    rd, qd = _write_dirs(lib_built, oracle, str(tmp_path), gold)
    r = subprocess.run([CLI, "dist", "-r", rd, "-o", os.path.join(str(tmp_path), "o1"), "-N", "18", qd], capture_output=True, text=True)
    assert r.returncode != 0 and "neighborN_max" in r.stderr            # more neighbours than references (command_dist.c:1578)
    raw = bytearray(open(os.path.join(qd, "cofiles.stat"), "rb").read())
    raw[0] ^= 1                                                          # another .shuf
    open(os.path.join(qd, "cofiles.stat"), "wb").write(bytes(raw))
    r = subprocess.run([CLI, "dist", "-r", rd, "-o", os.path.join(str(tmp_path), "o2"), qd], capture_output=True, text=True)
    assert r.returncode != 0 and "shuf_id" in r.stderr


@pytest.mark.skipif(os.environ.get("MK_TEST_MCO_DB") != "1", reason="opt-in (MK_TEST_MCO_DB=1): the reference binary writes a "
                    "32 GiB mco.index.0 and needs two minutes for it")
def test_cli_searches_a_database_built_by_the_reference(oracle, tmp_path):
    """A directory that holds only what the reference leaves in a database (mcofiles.stat, mco.0, mco.index.0): the host
    reads the inverted index back (read_mco_component) and distance.out equals the reference's own."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/metakssd is not built")
    subprocess.run(["make", "-C", os.path.join(ROOT, "host")], check=True, capture_output=True)
    (k, subk, L, seed), named, ref_names, qry_names, _ = G.dist_search_case()
    d = str(tmp_path)
    sid, perm = oracle.make_shuf(seed, k, subk, L)
    oracle.write_shuf_file(os.path.join(d, "x.shuf"), sid, k, subk, L, perm)
    for n, g in named.items():
        open(os.path.join(d, n), "wb").write(g)
    oracle.ref_dist(os.path.join(d, "x.shuf"), [os.path.join(d, n) for n in ref_names], os.path.join(d, "ref"), abundance=False, p=1)
    oracle.ref_dist(os.path.join(d, "x.shuf"), [os.path.join(d, n) for n in qry_names], os.path.join(d, "qry"), abundance=False, p=1)
    want = oracle.ref_dist_search(os.path.join(d, "ref"), os.path.join(d, "qry"), os.path.join(d, "out_ref"))
    for f in ("combco.0", "combco.index.0", "cofiles.stat"):
        os.remove(os.path.join(d, "ref", f))
    subprocess.run([CLI, "dist", "-r", os.path.join(d, "ref"), "-o", os.path.join(d, "out"), os.path.join(d, "qry")], check=True,
                   capture_output=True, timeout=600)
    assert open(os.path.join(d, "out", "distance.out")).read() == want
