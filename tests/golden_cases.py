"""Seeded inputs of the golden vectors (tests/golden/reference_vectors.npz).  The vectors were
produced by the unmodified reference binary at `-p 1` (tests/golden/make_golden.py); the inputs are
regenerated here from the same seeds with the C generator of the oracle library, so nothing large
is stored in the repository."""
import numpy as np

import oracle as O

MDB_PARAMS = (11, 6, 3, 4242)      # k, subk, L, shuf seed
MDB_SPECIES = 24


def mdb_synth():
    return O.synth(1001, MDB_SPECIES, 250_000, 150)


def mdb_fasta(S, s):
    return S.fasta(s)


def mdb_reads(S):
    return S.fastq(0, 40_000)


def _edge_variants(fq: bytes):
    lines = fq.split(b"\n")
    yield "as_is", fq
    yield "no_final_newline", fq[:-1]
    yield "missing_quality", b"\n".join(lines[:-2]) + b"\n"
    yield "plus_without_newline", b"\n".join(lines[:-2])
    yield "seq_only_tail", b"\n".join(lines[:-3]) + b"\n"
    yield "seq_unterminated", b"\n".join(lines[:-3])
    yield "header_only_tail", b"\n".join(lines[:-4]) + b"\n"
    yield "lowercase", fq.lower().replace(b"@r", b"@R")
    yield "crlf", fq.replace(b"\n", b"\r\n")
    yield "blank_line_shift", fq[:5000] + b"\n" + fq[5000:]
    yield "long_headers", fq.replace(b"@r", b"@" + b"ACGT" * 60 + b" read/")
    yield "tiny_reads", b"".join(b"@x\nACGTACGTAC\n+\nIIIIIIIIII\n" for _ in range(300))


def edge_base() -> bytes:
    """1500 reads followed by one record that is known to contribute codes found nowhere else in
    the file, so that every trailing-record variant changes the sketch."""
    S2 = O.synth(11, 5, 50_000, 150)
    p = O.params(11, 6, 3)
    _, perm = O.make_shuf(1234, 11, 6, 3)
    head = bytes(S2.fastq(0, 1500))
    have = set(O.fastq_koc(p, perm, head).codes.tolist())
    recs = bytes(S2.fastq(1500, 3000)).split(b"\n")
    for i in range(0, len(recs) - 1, 4):
        rec = b"\n".join(recs[i:i + 4]) + b"\n"
        got = set(O.fastq_koc(p, perm, rec).codes.tolist())
        if got and not (got & have):
            return head + rec
    raise AssertionError("no suitable tail record")


def fastq_cases():
    """(name, (k, subk, L, shuf_seed), text) — every FASTQ -A golden case."""
    S = O.synth(42, 20, 200_000, 150)
    yield "l3k11", (11, 6, 3, 1234), S.fastq(0, 30_000)
    yield "l2k11", (11, 5, 2, 2234), S.fastq(1000, 9_000)          # 16 components
    yield "l3k10", (10, 6, 3, 3234), S.fastq(500, 8_500)
    yield "l1k9", (9, 4, 1, 4234), S.fastq(0, 3_000)
    yield "k12", (12, 6, 3, 5234), S.fastq(200, 6_200)             # two 16-base pre-words on the GPU
    base = edge_base()
    for name, text in _edge_variants(base):
        yield "edge_" + name, (11, 6, 3, 1234), np.frombuffer(text, dtype=np.uint8)
