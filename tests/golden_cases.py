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


# ---------------------------------------------------------------------------------------------------------
# round 2: `dist` on FASTQ without -A (fastq2co: -Q / -n) and `dist -u` (uniq_fasta2co);
# vectors in tests/golden/reference_vectors_r2.npz (tests/golden/make_golden_r2.py)
def _vary_quality(fq: bytes, seed: int) -> bytes:
    """Quality lines with a spread of values: most bytes high, runs of low ones, a few bytes >= 0x80 (negative as
    a signed char, so they fail even Q = 0)."""
    rng = np.random.default_rng(seed)
    lines = fq.split(b"\n")
    for i in range(3, len(lines), 4):
        q = np.frombuffer(lines[i], dtype=np.uint8).copy()
        if q.size == 0:
            continue
        n_low = int(rng.integers(0, 4))
        for _ in range(n_low):
            a = int(rng.integers(0, q.size))
            q[a:a + int(rng.integers(1, 12))] = int(rng.integers(33, 60))
        if rng.random() < 0.05:
            q[int(rng.integers(0, q.size))] = 0x80 + int(rng.integers(0, 100))
        lines[i] = q.tobytes()
    return b"\n".join(lines)


def fastq_co_cases():
    """(name, (k, subk, L, shuf_seed), text, Q, M)"""
    S = O.synth(77, 8, 120_000, 150)
    base = _vary_quality(bytes(S.fastq(0, 12_000)), 5)
    geo = (11, 6, 3, 1234)
    for Q, M in ((0, 1), (40, 1), (0, 2), (45, 3), (70, 1)):
        yield "fqco_q%d_n%d" % (Q, M), geo, base, Q, M
    # tails: 1500 reads with varied quality + one record (quality all 'I') that contributes codes found nowhere else,
    # so that every rule about the last record changes the sketch
    eb = edge_base().rstrip(b"\n").split(b"\n")
    special = b"\n".join(eb[-4:]) + b"\n"
    head = _vary_quality(b"\n".join(eb[:-4]) + b"\n", 6)
    tbase = head + special
    lines = tbase.split(b"\n")
    first = b"\n".join(head.split(b"\n")[:4]) + b"\n"
    tails = {
        "as_is": tbase,
        "no_final_newline": tbase[:-1],
        "missing_quality": b"\n".join(lines[:-2]) + b"\n",
        "seq_unterminated": b"\n".join(lines[:-3]),
        "extra_header": tbase + b"@tail\n",
        "extra_three_lines": tbase + b"@tail\nACGTACGTACGTACGTACGTACGTACGTACGTACGT\n+\n",
        "crlf": tbase.replace(b"\n", b"\r\n"),
        "one_record": special,
        "one_record_unterminated": special[:-1],
        "two_records_second_unterminated": first + special[:-1],
        "two_records": first + special,
    }
    for name, text in tails.items():
        yield "fqco_tail_" + name, geo, text, 40, 1
    yield "fqco_l2k11", (11, 5, 2, 2234), _vary_quality(bytes(S.fastq(100, 3_100)), 9), 45, 2


def uniq_cases():
    """(name, (k, subk, L, shuf_seed), fasta text) for `dist -u`"""
    S = O.synth(78, 6, 150_000, 150)
    for s in range(3):
        g = bytes(S.fasta(s))
        body = g.split(b"\n", 1)[1]
        # the genome followed by a copy of part of it: the codes of that part occur twice
        yield "uniq_sp%d" % s, (11, 6, 3, 1234), g + b">copy\n" + body[: len(body) * (s + 1) // 4]
    yield "uniq_plain", (11, 6, 3, 1234), bytes(S.fasta(4))
    yield "uniq_l2k11", (11, 5, 2, 2234), bytes(S.fasta(5)) + b">c\n" + bytes(S.fasta(5)).split(b"\n", 1)[1][:40_000]


def set_case():
    """((k, subk, L, shuf_seed), genome texts, "taxid\\tname" lines) for `set -g / -q / -i`: 10 species with 1-4
    strains each (1 % substitutions), taxids that do not hash in input order, members of a species not adjacent."""
    S = O.synth(79, 10, 180_000, 150)
    rng = np.random.default_rng(11)
    taxids = [562, 1280, 1351, 287, 1313, 9606, 10090, 1773, 632, 727]
    n_strains = [3, 1, 4, 2, 3, 1, 2, 4, 3, 2]
    genomes, groups = [], []
    for s in range(10):
        g = np.frombuffer(bytes(S.fasta(s)), dtype=np.uint8).copy()
        hdr_end = int(np.flatnonzero(g == 10)[0]) + 1
        for st in range(n_strains[s]):
            t = g.copy()
            if st:
                idx = rng.integers(hdr_end, t.size, size=t.size // 100)
                idx = idx[t[idx] != 10]
                t[idx] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=idx.size)]
            genomes.append(t.tobytes())
            groups.append("%d\tsp%d" % (taxids[s], s))
    order = rng.permutation(len(genomes))
    return (11, 6, 3, 1234), [genomes[i] for i in order], [groups[i] for i in order]


# ---------------------------------------------------------------------------------------------------------
# `dist -r <ref> <qry>` (shared k-mer counts + distance table); vectors in tests/golden/reference_vectors_r2b.npz
# (tests/golden/make_golden_r2b.py).  The reference orders the files of a sketch directory by a time-seeded shuffle,
# so the vectors also hold the order of the names; tests rebuild the directories in that order.
DIST_SEARCH_OPTIONS = {
    "default": [],
    "containment": ["-M", "1"],
    "distance_only": ["-O", "0"],
    "with_qvalues": ["-O", "1"],
    "nearest3": ["-N", "3"],
    "nearest2_containment": ["-N", "2", "-M", "1"],
    "max_dist": ["-D", "0.08"],
    "corrected": ["--correction", "1"],
}


def dist_search_case():
    """((k, subk, L, shuf_seed), {name: genome text}, ref names, qry genome names, reads text): the 25 strain genomes of
    set_case() — 17 of them the reference database, 12 queries (4 also in the database) and one read sample."""
    geo, genomes, _ = set_case()
    named = {"g%02d.fasta" % i: g for i, g in enumerate(genomes)}
    names = sorted(named)
    S = O.synth(79, 10, 180_000, 150)
    return geo, named, names[:17], names[13:], bytes(S.fastq(0, 20_000))
