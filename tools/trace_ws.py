"""Role timeline of CTA 0 for the warp-specialised kernel (development tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
import metakssd_b200 as M
k, subk, L = 11, 6, 3
nreads = 2_000_000
sid, perm = M.make_shuf(1234, subk)
spec = M.synth_spec(42, 100, 1_000_000, 150)
sk = M.Sketcher(perm, k, subk, L)
nbytes = spec.fastq_bytes(0, nreads)
d = torch.empty(nbytes + 256, dtype=torch.uint8, device="cuda")
sk.synth_fastq_device(spec.P, spec.cdf32, spec.species, 0, nreads, d, d.numel())
tr = torch.zeros(48 * 32 * 4, dtype=torch.int64, device="cuda")
lib = M.load()
lib.mk_debug_set_trace.argtypes = [C.c_void_p, C.c_void_p]
sk.fastq_koc_device(d, nbytes)
lib.mk_debug_set_trace(sk._h, tr.data_ptr())
sk.fastq_koc_device(d, nbytes)
lib.mk_debug_set_trace(sk._h, None)
t = tr.cpu().numpy().reshape(48, 32, 4)
t0 = t[20, 0, 0]
for k_ in range(20, 26):
    r = t[k_]
    print("tile k=%d" % k_)
    print("  loader   issue@%6d" % (r[0, 0] - t0))
    print("  resolver start@%6d scanned@%6d resolved@%6d" % tuple(r[1, i] - t0 for i in range(3)))
    for w in (2, 4, 7):
        print("  front%d   full@%6d scan_done@%6d resolved(prev)@%6d mask_done@%6d" % ((w - 2,) + tuple(r[w, i] - t0 for i in range(4))))
    for w in (8, 12, 17):
        print("  probe%d   start@%6d ready@%6d probed@%6d" % ((w - 8,) + tuple(r[w, i] - t0 for i in range(3))))
