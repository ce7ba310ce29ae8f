"""Quick device-resident throughput probe (development tool, not the bench)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import metakssd_b200 as M
import oracle as O

k, subk, L = [int(x) for x in os.environ.get("KSL", "11,6,3").split(",")]
nreads = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
sid, perm = O.make_shuf(1234, k, subk, L)
S = O.synth(42, 100, 1_000_000, 150)
P = M.MksParams.from_buffer_copy(bytes(S.P))
sk = M.Sketcher(perm, k, subk, L)
nbytes = int(O.lib().ko_fastq_bytes(O.C.byref(S.P), 0, nreads))
d = torch.empty(nbytes + 256, dtype=torch.uint8, device="cuda")
t = time.time(); sk.synth_fastq_device(P, S.cdf32, S.species, 0, nreads, d, d.numel()); print("gen %.3fs, %d bytes" % (time.time() - t, nbytes))
for it in range(4):
    sk.profile(reset=True)
    t = time.time(); g = sk.fastq_koc_device(d, nbytes); dt = time.time() - t
    pr = sk.profile()
    gbp = nreads * 150 / 1e9
    print("iter %d: wall %.2f ms  stream %.3f ms (%.1f Gbp/s, %.0f GB/s text)  reduce %.3f ms  launches %d  codes %d" % (
        it, dt * 1e3, pr.stream_kernel_ms, gbp / (pr.stream_kernel_ms / 1e3), nbytes / 1e9 / (pr.stream_kernel_ms / 1e3),
        pr.reduce_ms, pr.kernel_launches, g.n_total))
