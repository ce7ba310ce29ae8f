"""GPU: `dist` on FASTQ without -A (mk_fastq_co_*: quality threshold -Q, least occurrence -n; fastq2co(),
iseq2comem.c:323-419) and `dist -u` (mk_ctx_set_dedup + mk_fasta_co_*; uniq_fasta2co(), iseq2comem.c:729-828)
against the reference binary's golden vectors and, on randomised inputs, against the oracle."""
import os
import random
import subprocess

import numpy as np
import pytest

import golden_cases as G
from helpers import same_sketch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_vectors_r2.npz")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fastq_co_against_reference_vectors(lib_built, shuf):
    gold = np.load(GOLD)
    ctxs = {}
    for name, (k, subk, L, seed), text, Q, M in G.fastq_co_cases():
        key = (k, subk, L, seed)
        if key not in ctxs:
            sid, perm = shuf(seed, k, subk, L)
            ctxs[key] = lib_built.Sketcher(perm, k, subk, L)
        got = ctxs[key].fastq_co_host(bytes(text), Q, M)
        assert got.counts is None and len(got.codes) == int(gold[name + "/comp_num"][0])
        for c in range(len(got.codes)):
            assert np.array_equal(got.codes[c], gold["%s/combco.%d" % (name, c)]), "%s comp %d" % (name, c)
    for s in ctxs.values():
        s.close()


def test_fasta_uniq_against_reference_vectors(lib_built, shuf):
    gold = np.load(GOLD)
    for name, (k, subk, L, seed), text in G.uniq_cases():
        sid, perm = shuf(seed, k, subk, L)
        with lib_built.Sketcher(perm, k, subk, L) as sk:
            sk.set_dedup(True)
            got = sk.fasta_co_host([bytes(text)])[0]
            for c in range(len(got.codes)):
                assert np.array_equal(got.codes[c], gold["%s/combco.%d" % (name, c)]), "%s comp %d" % (name, c)
            sk.set_dedup(False)      # and the option really switches back
            plain = sk.fasta_co_host([bytes(text)])[0]
            assert sum(c.size for c in plain.codes) >= sum(c.size for c in got.codes)


def test_fastq_co_random_against_oracle(lib_built, oracle, shuf):
    """ragged reads, random qualities (including bytes >= 0x80), every Q / M, tiny tiles"""
    rnd = random.Random(2)
    k, subk, L = 11, 6, 3
    sid, perm = shuf(4321, k, subk, L)
    p = oracle.params(k, subk, L)
    S = oracle.synth(17, 10, 60_000, 150)
    genome = bytes(S.fasta(0)).split(b"\n", 1)[1].replace(b"\n", b"")
    with lib_built.Sketcher(perm, k, subk, L) as sk:
        for case in range(40):
            recs = []
            for i in range(rnd.randrange(1, 700)):
                n = rnd.choice([22, 23, 40, 100, 150, 151, 300, 1000])
                a = rnd.randrange(0, len(genome) - n - 1)
                seq = bytearray(genome[a:a + n])
                if rnd.random() < 0.2:
                    seq[rnd.randrange(n)] = ord("N")
                qual = bytearray(rnd.choice([33, 45, 60, 73, 200]) if rnd.random() < 0.1 else 73 for _ in range(n))
                recs += [b"@h%d" % i, bytes(seq), b"+", bytes(qual)]
            text = b"\n".join(recs) + rnd.choice([b"\n", b"", b"\n@t\nACGT", b"\n@t\nACGTACGTACGTACGTACGTACGTACGT\n+\nIIIIIIIIIIIIIIIIIIIIIIIIIIII"])
            Q, M = rnd.choice([0, 34, 46, 61, 74]), rnd.choice([1, 1, 2, 3, 7])
            os.environ["MK_TILE_BYTES"] = str(rnd.choice([64, 192, 1024, 12288]))
            try:
                got = sk.fastq_co_host(text, Q, M)
            finally:
                os.environ.pop("MK_TILE_BYTES", None)
            same_sketch(got, oracle.fastq_co(p, perm, np.frombuffer(text, np.uint8), Q, M), p)


def test_cli_flags(lib_built, oracle, tmp_path):
    """host program: `dist -Q 40 -n 2 reads.fq` (no -A) and `dist -u genome.fasta` write what the oracle says"""
    subprocess.run(["make", "-C", os.path.join(ROOT, "host")], check=True, capture_output=True)
    cli = os.path.join(ROOT, "host", "metakssd-b200")
    k, subk, L = 11, 6, 3
    subprocess.run([cli, "shuffle", "-k", str(k), "-s", str(subk), "-l", str(L), "-o", str(tmp_path / "s"), "--seed", "9"],
                   check=True, capture_output=True)
    sid, kk, ss, ll, perm = lib_built.read_shuf(str(tmp_path / "s.shuf"))
    p = oracle.params(k, subk, L)
    name, geo, text, Q, M = next(c for c in G.fastq_co_cases() if c[0] == "fqco_q45_n3")
    fq = tmp_path / "r.fq"
    fq.write_bytes(bytes(text))
    subprocess.run([cli, "dist", "-L", str(tmp_path / "s.shuf"), "-Q", "45", "-n", "3", "-o", str(tmp_path / "o1"), str(fq)],
                   check=True, capture_output=True, timeout=300)
    sd = oracle.read_sketch_dir(str(tmp_path / "o1"))
    want = oracle.fastq_co(p, perm, np.frombuffer(bytes(text), np.uint8), 45, 3)
    assert not sd.koc and np.array_equal(sd.combco[0], want.components(p)[0][0]) and sd.combco[0].size > 0
    uname, ugeo, utext = next(c for c in G.uniq_cases() if c[0] == "uniq_sp1")
    fa = tmp_path / "g.fasta"
    fa.write_bytes(bytes(utext))
    subprocess.run([cli, "dist", "-L", str(tmp_path / "s.shuf"), "-u", "-o", str(tmp_path / "o2"), str(fa)],
                   check=True, capture_output=True, timeout=300)
    sd = oracle.read_sketch_dir(str(tmp_path / "o2"))
    want = oracle.fasta_co_uniq(p, perm, np.frombuffer(bytes(utext), np.uint8))
    assert np.array_equal(sd.combco[0], want.components(p)[0][0]) and sd.combco[0].size > 0
