"""Race hunting: repeat small FASTA / FASTQ sketches and compare with the oracle (development tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle as O
import metakssd_b200 as M
from helpers import same_sketch
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
k, subk, L = 11, 6, 3
sid, perm = O.make_shuf(1234 + k * 100 + subk, k, subk, L)
p = O.params(k, subk, L)
S = O.synth(5, 8, 150000, 150)
g = [bytes(S.fasta(i)) for i in range(4)]
cases = [g[0], g[0] + g[1], g[2][:4000] + b"12 -*\n" + g[2][4000:], g[3]]
want = [O.fasta_co(p, perm, t) for t in cases]
fq = bytes(S.fastq(0, 20000))
want_fq = O.fastq_koc(p, perm, fq)
bad = 0
t0 = time.time()
with M.Sketcher(perm, k, subk, L) as sk:
    for r in range(reps):
        got = sk.fasta_co_host(cases)
        for i, (gt, w) in enumerate(zip(got, want)):
            try:
                same_sketch(gt, w, p)
            except AssertionError as e:
                bad += 1
                print("rep", r, "fasta case", i, str(e)[:80])
        try:
            same_sketch(sk.fastq_koc_host(np.frombuffer(fq, dtype=np.uint8)), want_fq, p)
        except AssertionError as e:
            bad += 1
            print("rep", r, "fastq", str(e)[:80])
print("reps", reps, "mismatches", bad, "%.1fs" % (time.time() - t0))
