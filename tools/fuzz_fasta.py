"""Randomised parity check of the FASTA batch path against the oracle (development tool): headers
anywhere, '>' inside lines, CR, lower case, non-ACGT bytes, empty files, file boundaries at arbitrary
offsets relative to the 512-byte steps and 32 KB tiles of the compaction kernels."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle as O
import metakssd_b200 as M
from helpers import same_sketch

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 50
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rnd = random.Random(seed)
k, subk, L = 11, 6, 3
sid, perm = O.make_shuf(4321, k, subk, L)
p = O.params(k, subk, L)
S = O.synth(23, 6, 400000, 150)
genomes = [bytes(S.fasta(i)).split(b"\n", 1)[1].replace(b"\n", b"") for i in range(6)]


def rand_file():
    kind = rnd.randrange(8)
    if kind == 0:
        return b""
    if kind == 1:
        return b">only a header" + (b"\n" if rnd.random() < 0.5 else b"")
    g = genomes[rnd.randrange(6)]
    a = rnd.randrange(0, len(g) - 1)
    n = rnd.choice([1, 20, 21, 22, 23, 100, 511, 512, 513, 3000, 32767, 32768, 32769, 70000, 200000])
    seq = bytearray(g[a:a + n])
    for _ in range(rnd.randrange(0, 6)):
        if seq:
            seq[rnd.randrange(len(seq))] = rnd.choice(b"NnRYacgt*-> 1")
    if rnd.random() < 0.3:
        seq = bytearray(bytes(seq).lower())
    width = rnd.choice([60, 80, 70, 1000, 10 ** 9])
    lines = [bytes(seq[i:i + width]) for i in range(0, len(seq), width)] or [b""]
    nl = b"\r\n" if rnd.random() < 0.2 else b"\n"
    body = nl.join(lines) + (nl if rnd.random() < 0.8 else b"")
    if rnd.random() < 0.2 and len(lines) > 2:     # a second record in the file
        cut = len(body) // 2
        body = body[:cut] + b"\n>second record ACGTACGTACGTACGTACGTACGTAC\n" + body[cut:]
    hdr = b"" if rnd.random() < 0.15 else b">" + bytes(rnd.choice(b"ACGTxyz >_|") for _ in range(rnd.randrange(0, 90))) + nl
    return hdr + body


bad = 0
with M.Sketcher(perm, k, subk, L) as sk:
    for c in range(n_cases):
        files = [rand_file() for _ in range(rnd.randrange(1, 12))]
        want = [O.fasta_co(p, perm, f) for f in files]
        try:
            got = sk.fasta_co_host(files)
            for i, (g, w) in enumerate(zip(got, want)):
                try:
                    same_sketch(g, w, p)
                except AssertionError as e:
                    raise AssertionError("file %d of %d (%d bytes): %s" % (i, len(files), len(files[i]), e))
        except Exception as e:
            bad += 1
            print("case %d (seed %d): %s" % (c, seed, str(e)[:160]))
print("cases", n_cases, "failures", bad)
