"""Wall-clock breakdown of one bench step at bench size (development tool).  MK_TIMING=1 adds the
library's own phase marks."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import metakssd_b200 as M
from metakssd_b200 import workload as W
reads = int(sys.argv[1]) if len(sys.argv) > 1 else 40_000_000
shuf_id, perm = M.make_shuf(20240917 ^ 1, 6)
sk = M.Sketcher(perm, 11, 6, 3)
spec = M.synth_spec(20240917 ^ 2, 1000, 5_000_000, 150)
mdb = W.build_markerdb(sk, spec)
sk.load_markerdb(mdb.comp)
nbytes = spec.fastq_bytes(0, reads)
d = torch.empty(nbytes + 256, dtype=torch.uint8, device="cuda")
sk.synth_fastq_device(spec.P, spec.cdf32, spec.species, 0, reads, d, d.numel())
def sync(): torch.cuda.synchronize()
for it in range(4):
    sync(); t0 = time.perf_counter()
    s = sk.fastq_koc_device(d, nbytes)
    sync(); t1 = time.perf_counter()
    qry = [(s.codes[c], s.counts[c]) for c in range(len(s.codes))]
    stats = sk.composite(None, qry)
    sync(); t2 = time.perf_counter()
    tsv = M.composite_tsv("reads.fq", mdb.names, stats)
    t3 = time.perf_counter()
    print("iter %d: sketch %.3f ms  composite %.3f ms  tsv %.3f ms  total %.3f ms  (codes %d)" %
          (it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t3 - t0) * 1e3, sum(len(c) for c in s.codes)), file=sys.stderr)
