"""Randomised parity check of the FASTQ path against the oracle (development tool): ragged line
structure, tiny tiles (many tickets / groups per CTA), forced multi-chunk uploads."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle as O
import metakssd_b200 as M
from helpers import same_sketch

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rnd = random.Random(seed)
k, subk, L = 11, 6, 3
sid, perm = O.make_shuf(4321, k, subk, L)
p = O.params(k, subk, L)
S = O.synth(17, 10, 60000, 150)
genome = bytes(S.fasta(0)).split(b"\n", 1)[1].replace(b"\n", b"")
base_recs = bytes(S.fastq(0, 4000)).split(b"\n")


def rand_seq(n):
    a = rnd.randrange(0, len(genome) - n - 1)
    s = bytearray(genome[a:a + n])
    for _ in range(rnd.randrange(0, 3)):
        if n:
            s[rnd.randrange(n)] = rnd.choice(b"NnRacgt*-")
    return bytes(s)


def rand_text():
    out = []
    mode = rnd.randrange(4)
    n_rec = rnd.randrange(1, 1500)
    for i in range(n_rec):
        if mode == 0:                      # ordinary records
            j = 4 * rnd.randrange(0, 3999)
            out += base_recs[j:j + 4]
        elif mode == 1:                    # ragged read lengths, odd headers
            n = rnd.choice([0, 1, 5, 21, 22, 23, 31, 32, 33, 63, 64, 65, 150, 300, 1000, 3000])
            out += [b"@" + b"x" * rnd.randrange(0, 70), rand_seq(n), b"+" + b"y" * rnd.randrange(0, 3), b"I" * n]
        elif mode == 2:                    # structure noise: blank lines, missing lines, CRs
            r = rnd.random()
            seq = rand_seq(rnd.choice([30, 100, 150]))
            if r < 0.1:
                out.append(b"")
            elif r < 0.2:
                out += [b"@h", seq]
            elif r < 0.3:
                out += [b"@h\r", seq + b"\r", b"+\r", b"I" * len(seq) + b"\r"]
            else:
                out += [b"@h%d" % i, seq, b"+", b"I" * len(seq)]
        else:                              # sequence-looking quality / header lines
            seq = rand_seq(150)
            out += [b"@" + rand_seq(40), seq, b"+" + rand_seq(10), rand_seq(150)]
    text = b"\n".join(out)
    tail = rnd.randrange(4)
    if tail == 0:
        text += b"\n"
    elif tail == 1:
        text += b"\n@last\nACGTACGTACGTACGTACGTACGTACGTAC"
    elif tail == 2:
        text += b"\n\n"
    return text


bad = 0
with M.Sketcher(perm, k, subk, L) as sk:
    for c in range(n_cases):
        text = rand_text()
        tile = rnd.choice([64, 128, 192, 448, 1024, 4096, 12288])
        chunk = rnd.choice([0, 0, 512, 4096, 50000])
        os.environ["MK_TILE_BYTES"] = str(tile)
        if chunk:
            os.environ["MK_CHUNK_BYTES"] = str(chunk)
        else:
            os.environ.pop("MK_CHUNK_BYTES", None)
        arr = np.frombuffer(text, np.uint8).copy()
        want = O.fastq_koc(p, perm, arr)
        try:
            got = sk.fastq_koc_host(arr)
            same_sketch(got, want, p)
        except Exception as e:
            bad += 1
            print("case %d (seed %d, %d bytes, tile %d, chunk %d): %s" % (c, seed, len(text), tile, chunk, str(e)[:120]))
            open("/tmp/fuzz_fail_%d_%d.fq" % (seed, c), "wb").write(text)
print("cases", n_cases, "failures", bad)
