"""GPU: the C host program (host/mkssd_main.c) end to end — reference flags, reference file formats."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "host", "metakssd-b200")


@pytest.fixture(scope="module")
def cli(lib_built):
    subprocess.run(["make", "-C", os.path.join(ROOT, "host")], check=True, capture_output=True)
    return CLI


def _run(*args):
    return subprocess.run([CLI] + [str(a) for a in args], check=True, capture_output=True, text=True, timeout=300)


def test_cli_dist_and_composite(cli, oracle, lib_built, tmp_path):
    from helpers import markerdb_from_sketches
    k, subk, L, seed = 11, 6, 3, 31337
    _run("shuffle", "-k", k, "-s", subk, "-l", L, "-o", tmp_path / "L3K11", "--seed", seed)
    shuf = tmp_path / "L3K11.shuf"
    sid, kk, ss, ll, perm = lib_built.read_shuf(str(shuf))
    sid2, perm2 = lib_built.make_shuf(seed, subk)
    assert (sid, kk, ss, ll) == (sid2, k, subk, L) and np.array_equal(perm, perm2)
    p = oracle.params(k, subk, L)
    S = oracle.synth(9, 16, 200_000, 150)
    # query sketch
    fq = tmp_path / "reads.fq"
    S.fastq(0, 30_000).tofile(fq)
    r = _run("dist", "-L", shuf, "-A", "-p", 4, "-o", tmp_path / "qry", fq)
    assert "hashsize=33554393" in r.stdout
    sd = oracle.read_sketch_dir(str(tmp_path / "qry"))
    want = oracle.fastq_koc(p, perm, S.fastq(0, 30_000))
    assert sd.koc and sd.shuf_id == sid and sd.kmerlen == 22 and sd.dim_rd_len == 6 and sd.names == [str(fq)]
    assert np.array_equal(sd.combco[0], want.components(p)[0][0]) and np.array_equal(sd.abund[0], want.counts)
    assert list(sd.index[0]) == [0, want.codes.size] and sd.all_ctx_ct == want.codes.size
    # genome sketches (FASTA, no -A), several files in one call
    paths = []
    for s in range(16):
        fa = tmp_path / ("sp%d.fasta" % s)
        S.fasta(s).tofile(fa)
        paths.append(fa)
    _run("dist", "-L", shuf, "-o", tmp_path / "gsk", *paths)
    gd = oracle.read_sketch_dir(str(tmp_path / "gsk"))
    assert not gd.koc and gd.infile_num == 16
    sketches = []
    for s in range(16):
        w = oracle.fasta_co(p, perm, S.fasta(s)).components(p)[0][0]
        got = gd.combco[0][int(gd.index[0][s]):int(gd.index[0][s + 1])]
        assert np.array_equal(got, w), "genome %d" % s
        sketches.append(w)
    # MarkerDB directory written in the reference format, then composite through the CLI
    codes, index = markerdb_from_sketches(sketches)
    names = ["%d_sp%d" % (s + 1, s) for s in range(16)]
    info = lib_built.MkInfo()
    info.k, info.drlevel, info.component_num = k, L, 1
    mdb = [lib_built.Sketch([codes[int(index[s]):int(index[s + 1])]], None) for s in range(16)]
    lib_built.write_sketch_dir(str(tmp_path / "markerdb"), sid, info, names, mdb, koc=False)
    out = _run("composite", "-r", tmp_path / "markerdb", "-q", tmp_path / "qry").stdout
    assert out == oracle.composite([(codes, index)], names, [want.components(p)[0]], str(fq)) and out.count("\n") >= 3


def test_cli_reports_crowded_like_the_reference(cli, oracle, lib_built, tmp_path):
    _run("shuffle", "-k", 7, "-s", 6, "-l", 3, "-o", tmp_path / "x", "--seed", 99)
    S = oracle.synth(8, 100, 200_000, 150)
    fq = tmp_path / "r.fq"
    S.fastq(0, 60_000).tofile(fq)
    r = subprocess.run([CLI, "dist", "-L", str(tmp_path / "x.shuf"), "-A", "-o", str(tmp_path / "o"), str(fq)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "too crowd" in r.stderr and "-k8" in r.stderr


def test_cli_dist_takes_a_directory_and_a_list_file(cli, oracle, tmp_path):
    """A directory argument and `-l <list>` give the sketch directory the explicit file arguments give
    (organize_infile_frm_arg / organize_infile_list, global_basic.c:169-330)."""
    S = oracle.synth(9, 6, 120_000, 150)
    _run("shuffle", "-k", 11, "-s", 6, "-l", 3, "-o", tmp_path / "L3K11", "--seed", 5)
    shuf = tmp_path / "L3K11.shuf"
    g = tmp_path / "genomes"
    g.mkdir()
    paths = []
    for s in range(6):
        fa = g / ("sp%d.fasta" % s)
        S.fasta(s).tofile(fa)
        paths.append(str(fa))
    (g / "README.txt").write_text("not a sequence file")
    (tmp_path / "list.txt").write_text("\n".join(paths) + "\n")
    _run("dist", "-L", shuf, "-o", tmp_path / "by_files", *paths)
    _run("dist", "-L", shuf, "-o", tmp_path / "by_dir", g)
    _run("dist", "-L", shuf, "-o", tmp_path / "by_list", "-l", tmp_path / "list.txt")
    want = [open(tmp_path / "by_files" / f, "rb").read() for f in ("combco.0", "combco.index.0", "cofiles.stat")]
    for d in ("by_dir", "by_list"):
        assert [open(tmp_path / d / f, "rb").read() for f in ("combco.0", "combco.index.0", "cofiles.stat")] == want, d
    assert len(want[0]) > 4 * 100
