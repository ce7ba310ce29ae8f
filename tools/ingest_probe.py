"""dev probe: CLI `dist -A` on a page-cached FASTQ file, phases by MK_TIMING, over ingest thread counts / chunk sizes"""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import metakssd_b200 as M
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
sid, perm = M.make_shuf(99, 6)
spec = M.synth_spec(5, 100, 1_000_000, 150)
d = "/tmp/ingest_probe"; os.makedirs(d, exist_ok=True)
M.write_shuf(d + "/L3K11.shuf", sid, 11, 6, 3, perm)
with M.Sketcher(perm, 11, 6, 3) as sk:
    nb = spec.fastq_bytes(0, n)
    t = torch.empty(nb + 256, dtype=torch.uint8, device="cuda")
    sk.synth_fastq_device(spec.P, spec.cdf32, spec.species, 0, n, t, t.numel())
    t[:nb].cpu().numpy().tofile(d + "/reads.fq")
    del t
print("file bytes", nb)
subprocess.run(["make", "-C", "host"], check=True, capture_output=True)
for env in ({}, {"MK_INGEST_THREADS": "4"}, {"MK_INGEST_THREADS": "8"}, {"MK_INGEST_THREADS": "16"}, {"MK_INGEST_CHUNK_BYTES": str(256 << 20)},
            {"MK_INGEST_CHUNK_BYTES": str(16 << 20)}, {"MK_INGEST_BUFFERS": "8"}):
    e = dict(os.environ, MK_TIMING="1", **env)
    t0 = time.time()
    r = subprocess.run(["host/metakssd-b200", "dist", "-L", d + "/L3K11.shuf", "-A", "-o", d + "/out", d + "/reads.fq"], env=e, capture_output=True, text=True)
    dt = time.time() - t0
    ph = [l for l in r.stderr.splitlines() if "[host" in l or "sketching" in l or "ctx" in l]
    print(env, "wall %.3f s" % dt, "|", " ; ".join(x.strip() for x in ph)[:400])
import shutil
shutil.rmtree(d + "/out2", ignore_errors=True); shutil.copytree(d + "/out", d + "/out2")
for i in range(2):
    t0 = time.time()
    r = subprocess.run(["host/metakssd-b200", "composite", "-r", d + "/out2", "-q", d + "/out"], env=dict(os.environ, MK_TIMING="1"), capture_output=True, text=True)
    print("composite wall %.3f s" % (time.time() - t0), "|", " ; ".join(x.strip() for x in r.stderr.splitlines() if "timing" in x)[:500])
t0 = time.time(); subprocess.run(["host/metakssd-b200", "--help"], capture_output=True); print("process start (--help) %.3f s" % (time.time() - t0))
