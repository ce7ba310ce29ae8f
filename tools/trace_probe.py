"""Per-warp phase timeline of CTA 0 (development tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
import metakssd_b200 as M
import oracle as O
k, subk, L = 11, 6, 3
nreads = 2_000_000
NW = int(os.environ.get("NW", "16"))
sid, perm = O.make_shuf(1234, k, subk, L)
S = O.synth(42, 100, 1_000_000, 150)
P = M.MksParams.from_buffer_copy(bytes(S.P))
sk = M.Sketcher(perm, k, subk, L)
nbytes = int(O.lib().ko_fastq_bytes(O.C.byref(S.P), 0, nreads))
d = torch.empty(nbytes + 256, dtype=torch.uint8, device="cuda")
sk.synth_fastq_device(P, S.cdf32, S.species, 0, nreads, d, d.numel())
tr = torch.zeros(64 * NW * 8, dtype=torch.int64, device="cuda")
lib = M.load()
lib.mk_debug_set_trace.argtypes = [C.c_void_p, C.c_void_p]
sk.fastq_koc_device(d, nbytes)
lib.mk_debug_set_trace(sk._h, tr.data_ptr())
sk.fastq_koc_device(d, nbytes)
lib.mk_debug_set_trace(sk._h, None)
t = tr.cpu().numpy().reshape(64, NW, 8)
base = t[:, :, 0].min()
for it in range(20, 26):
    t0 = t[it, :, 0].min()
    print("iter %d tile %d items %d  start(all warps) spread %d; iteration length %d" % (
        it, t[it, 0, 7], t[it, 0, 6], t[it, :, 0].max() - t0, t[it + 1, :, 0].min() - t0))
    for w in range(NW):
        r = t[it, w]
        print("  w%02d  resolve+%5d  P+%5d  S+%5d  wait+%5d  M+%5d   end@%6d" % (
            w, r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], r[5] - r[4], r[5] - t0))
