"""GPU: BASELINE.json configs[0] end to end against the UNMODIFIED reference binary (oracle/_ref/metakssd, built
from /root/reference by oracle/Makefile; it travels to the GPU box as a built file).

  * 1 M x 150 bp reads, 100 species x 1 Mbp; the MarkerDB comes from the reference's own pipeline
    `dist` -> `set -g` -> `set -q` -> `set -i` (command_set.c:831, 427, 322);
  * `host/metakssd-b200 dist -L L3K11.shuf -A` writes a sketch directory that is compared BYTE FOR BYTE with the one
    the reference writes at `-p 1` (combco.0, combco.0.a, combco.index.0; cofiles.stat with its three
    uninitialised padding bytes 5..7 masked);
  * species_coverage of `host/metakssd-b200 composite` is compared byte for byte with the reference's;
  * cross-format: the reference `composite` reads OUR sketch directory, our `composite` reads the REFERENCE's
    sketch directory and the reference-built MarkerDB; our genome (FASTA) sketches against the reference's.
"""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "host", "metakssd-b200")
K, SUBK, L = 11, 6, 3
N_READS, N_SPECIES, GENOME = 1_000_000, 100, 1_000_000


@pytest.fixture(scope="module")
def world(lib_built, oracle, tmp_path_factory):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/metakssd is not built (needs /root/reference once)")
    subprocess.run(["make", "-C", os.path.join(ROOT, "host")], check=True, capture_output=True)
    d = str(tmp_path_factory.mktemp("cfg1"))
    sid, perm = oracle.make_shuf(0xC0F1, K, SUBK, L)
    shuf = os.path.join(d, "L3K11.shuf")
    oracle.write_shuf_file(shuf, sid, K, SUBK, L, perm)
    S = oracle.synth(0xC0F1 ^ 2, N_SPECIES, GENOME, 150)
    gdir = os.path.join(d, "genomes")
    os.makedirs(gdir)
    paths, groups = [], []
    for s in range(N_SPECIES):
        p = os.path.join(gdir, "sp%d.fasta" % s)
        S.fasta(s).tofile(p)
        paths.append(p)
        groups.append("%d\tsp%d" % (s + 1, s))
    threads = os.cpu_count() or 1
    mdb = oracle.ref_build_markerdb(shuf, paths, groups, d, p=threads)       # dist -> set -g -> set -q -> set -i
    fq = os.path.join(d, "reads.fq")
    S.fastq(0, N_READS).tofile(fq)
    ref_qry = os.path.join(d, "qry_ref")
    oracle.ref_dist(shuf, [fq], ref_qry, abundance=True, p=1)                # the parity target is -p 1
    ref_tsv = oracle.ref_composite(mdb, ref_qry, p=1)
    our_qry = os.path.join(d, "qry_gpu")
    subprocess.run([CLI, "dist", "-L", shuf, "-A", "-o", our_qry, fq], check=True, capture_output=True, timeout=600)
    return dict(d=d, shuf=shuf, paths=paths, mdb=mdb, fq=fq, ref_qry=ref_qry, our_qry=our_qry, ref_tsv=ref_tsv)


def _bytes(path):
    with open(path, "rb") as f:
        return f.read()


def test_sketch_directory_is_byte_identical(world):
    for name in ("combco.0", "combco.0.a", "combco.index.0"):
        a, b = _bytes(os.path.join(world["ref_qry"], name)), _bytes(os.path.join(world["our_qry"], name))
        assert a == b, "%s differs (%d vs %d bytes)" % (name, len(a), len(b))
    a = bytearray(_bytes(os.path.join(world["ref_qry"], "cofiles.stat")))
    b = bytearray(_bytes(os.path.join(world["our_qry"], "cofiles.stat")))
    assert len(a) == len(b) == 32 + 4 + 256
    a[5:8] = b[5:8] = b"\0\0\0"                       # struct padding after `bool koc`: uninitialised in the reference
    assert a[:36] == b[:36]
    assert a[36:].split(b"\0", 1)[0] == b[36:].split(b"\0", 1)[0]     # name record compared as a C string
    assert len(_bytes(os.path.join(world["our_qry"], "combco.0"))) > 4 * 5000


def test_species_coverage_is_byte_identical(world):
    r = subprocess.run([CLI, "composite", "-r", world["mdb"], "-q", world["our_qry"]], check=True, capture_output=True,
                       text=True, timeout=300)
    ours = [l for l in r.stdout.splitlines() if l.count("\t") >= 6]
    ref = [l for l in world["ref_tsv"].splitlines() if l.count("\t") >= 6]
    assert len(ref) >= 10
    assert ours == ref


def test_reference_composite_reads_our_sketch_directory(world, oracle):
    tsv = oracle.ref_composite(world["mdb"], world["our_qry"], p=1)
    assert [l for l in tsv.splitlines() if "\t" in l] == [l for l in world["ref_tsv"].splitlines() if "\t" in l]


def test_our_composite_reads_the_reference_sketch_directory(world):
    r = subprocess.run([CLI, "composite", "-r", world["mdb"], "-q", world["ref_qry"]], check=True, capture_output=True,
                       text=True, timeout=300)
    assert [l for l in r.stdout.splitlines() if l.count("\t") >= 6] == \
           [l for l in world["ref_tsv"].splitlines() if l.count("\t") >= 6]


def test_genome_sketches_match_the_reference(world, oracle):
    """`dist` without -A over all genomes in one call (batched FASTA path) against the reference's gsk directory."""
    out = os.path.join(world["d"], "gsk_gpu")
    subprocess.run([CLI, "dist", "-L", world["shuf"], "-o", out] + world["paths"], check=True, capture_output=True, timeout=600)
    ours = oracle.read_sketch_dir(out)
    ref = oracle.read_sketch_dir(os.path.join(world["d"], "gsk"))
    assert ours.infile_num == ref.infile_num == N_SPECIES and not ours.koc
    by_name = {n: i for i, n in enumerate(ours.names)}
    for i, n in enumerate(ref.names):                 # the reference lists the files in a time-seeded order
        j = by_name[n]
        a = ref.combco[0][int(ref.index[0][i]):int(ref.index[0][i + 1])]
        b = ours.combco[0][int(ours.index[0][j]):int(ours.index[0][j + 1])]
        assert np.array_equal(a, b), n
        assert int(ref.ctx_ct[i]) == int(ours.ctx_ct[j])
    assert ours.all_ctx_ct == ref.all_ctx_ct


def test_api_file_entry_point_matches_too(world, lib_built, oracle):
    sid, k, subk, l, perm = lib_built.read_shuf(world["shuf"])
    with lib_built.Sketcher(perm, k, subk, l) as sk:
        got = sk.fastq_koc_file(world["fq"])
    ref = oracle.read_sketch_dir(world["ref_qry"])
    assert np.array_equal(got.codes[0], ref.combco[0]) and np.array_equal(got.counts[0], ref.abund[0])
