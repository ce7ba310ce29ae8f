"""Role timeline of CTA 0 for the warp-specialised kernel (development tool).
Needs a library built with the stamps compiled in: MK_NVCC_FLAGS=-DMK_TRACE python -m metakssd_b200.build --force"""
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
NCW, NSW, NMW, NPG, NPW = [int(x) for x in os.environ.get("WS", "4,6,6,2,7").split(",")]
R0 = 2 + NCW
for k_ in range(20, 26):
    r = t[k_]
    print("tile k=%d" % k_)
    print("  loader   freed@%6d issued@%6d" % tuple(r[0, i] - t0 for i in range(2)))
    tm = k_ % NPG
    SCT, MKT = NSW // NPG, NMW // NPG
    for w in (R0 + tm * SCT, R0 + tm * SCT + SCT - 1):
        print("  scan%d    full@%6d scan_done@%6d" % ((w - R0,) + tuple(r[w, i] - t0 for i in range(2))))
    for w in (R0 + NSW + tm * MKT, R0 + NSW + tm * MKT + MKT - 1):
        print("  mask%d    start@%6d mask_done@%6d" % ((w - R0 - NSW,) + tuple(r[w, i] - t0 for i in (2, 3))))
    g = k_ % NPG
    b = R0 + NSW + NMW + g * NPW
    for w in (b, b + NPW - 1):
        print("  probe%d.%d start@%6d ready@%6d probed@%6d" % ((g, w - b) + tuple(r[w, i] - t0 for i in range(3))))
if os.environ.get("COMPACT"):
    print("role timelines (per tile k: wait-start, go, done) relative to first stamp")
    for name, w, sl in (("loader", 0, (0, 1, 1)), ("scanT0", R0, (0, 0, 1)), ("scanT1", R0 + SCT, (0, 0, 1)),
                        ("maskT0", R0 + NSW, (2, 2, 3)), ("maskT1", R0 + NSW + MKT, (2, 2, 3)),
                        ("probe0", R0 + NSW + NMW, (0, 1, 2)), ("probe1", R0 + NSW + NMW + NPW, (0, 1, 2))):
        line = []
        for k_ in range(16, 40):
            r = t[k_, w]
            if r[sl[2]] == 0: continue
            line.append("k%d:%d/%d/%d" % (k_, r[sl[0]] - t0, r[sl[1]] - t0, r[sl[2]] - t0))
        print(name, " ".join(line))
