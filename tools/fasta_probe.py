"""FASTA (MarkerDB-build) path throughput probe (development tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import metakssd_b200 as M
n_genomes = int(sys.argv[1]) if len(sys.argv) > 1 else 200
shuf_id, perm = M.make_shuf(99, 6)
sk = M.Sketcher(perm, 11, 6, 3)
spec = M.synth_spec(7, n_genomes, 5_000_000, 150)
per = spec.fasta_bytes(0) + 16
buf = torch.empty(n_genomes * per + 256, dtype=torch.uint8, device="cuda")
off = sk.synth_fasta_device(spec.P, 0, n_genomes, buf, buf.numel())
nbytes = int(off[-1])
for it in range(4):
    torch.cuda.synchronize(); sk.profile(reset=True); t0 = time.perf_counter()
    out = sk.fasta_co_device(buf, off)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    pr = sk.profile()
    print("iter %d: %d genomes %.2f GB  wall %.2f ms (%.1f Gbp/s)  stream kernel %.3f ms  reduce %.3f ms  launches %d  codes %d" % (
        it, n_genomes, nbytes / 1e9, dt * 1e3, n_genomes * 5e6 / 1e9 / dt, pr.stream_kernel_ms, pr.reduce_ms, pr.kernel_launches,
        sum(s.n_total for s in out)), file=sys.stderr)
