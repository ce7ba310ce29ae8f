
import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch, torch.distributed as dist
import metakssd_b200 as M
from metakssd_b200 import distributed as D
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
k, subk, L = 11, 6, 3
sid, perm = M.make_shuf(77, subk)
spec = M.synth_spec(5, 50, 300_000, 150)
per = 200_000
sk = M.Sketcher(perm, k, subk, L, device=local)
r0, r1 = rank * per, (rank + 1) * per
nb = spec.fastq_bytes(r0, r1)
d = torch.empty(nb + 256, dtype=torch.uint8, device=dev)
sk.synth_fastq_device(spec.P, spec.cdf32, spec.species, r0, r1, d, d.numel())
got = D.sketch_sharded(sk, d, nb, spec.fastq_bytes(0, r0), 0, rank == world - 1)
# the same shard from pinned host memory, uploaded in (forced small) chunks under the kernel
h = torch.empty(nb, dtype=torch.uint8, pin_memory=True)
h.copy_(d[:nb]); torch.cuda.synchronize(dev)
os.environ["MK_CHUNK_BYTES"] = "3000000"
got_h = D.sketch_sharded(sk, h, nb, spec.fastq_bytes(0, r0), 0, rank == world - 1, host_text=True)
del os.environ["MK_CHUNK_BYTES"]
# the same step inside the library: grouped ncclSend/ncclRecv by code range, owner merge, rank-local composite
# against the MarkerDB slice of the code range, slot order on rank 0 (csrc/mk_comm.cu)
from metakssd_b200 import workload as W
mdb = W.build_markerdb(sk, spec)
D.init_library_comm(sk)
sk.load_markerdb_sharded(mdb.comp)
cap = D.size_exchange_blocks(sk, d, nb, spec.fastq_bytes(0, r0), 0, rank == world - 1)
n_local = int(sk.fastq_partial_device(d, nb, spec.fastq_bytes(0, r0), 0, rank == world - 1).n)
assert 1000 < cap < n_local, (cap, n_local)          # balanced code ranges: a block is a fraction of a shard's runs
got_l, stats_l = sk.fastq_koc_sharded(d, nb, spec.fastq_bytes(0, r0), 0, rank == world - 1, cap)
got_lh, stats_lh = sk.fastq_koc_sharded(h, nb, spec.fastq_bytes(0, r0), 0, rank == world - 1, cap, host_text=True)
# a block capacity that is too small must be reported, not silently truncated
failed = 0
try:
    sk.fastq_koc_sharded(d, nb, spec.fastq_bytes(0, r0), 0, rank == world - 1, 64)
except M.MkError as e:
    failed = 1 if e.code == -4 else 0
f = torch.tensor([failed], dtype=torch.int64, device=dev); dist.all_reduce(f, op=dist.ReduceOp.MIN)
assert int(f.item()) == 1, "an overflowing exchange block must fail the step on every rank"
if rank == 0:
    nball = spec.fastq_bytes(0, world * per)
    full = torch.empty(nball + 256, dtype=torch.uint8, device=dev)
    sk.synth_fastq_device(spec.P, spec.cdf32, spec.species, 0, world * per, full, full.numel())
    want = sk.fastq_koc_device(full, nball)
    assert got.n_total == want.n_total > 1000, (got.n_total, want.n_total)
    for c in range(len(want.codes)):
        assert np.array_equal(got.codes[c], want.codes[c]), "codes/order differ"
        assert np.array_equal(got.counts[c], want.counts[c]), "counts differ"
        assert np.array_equal(got_h.codes[c], want.codes[c]), "host-text path: codes/order differ"
        assert np.array_equal(got_h.counts[c], want.counts[c]), "host-text path: counts differ"
        assert np.array_equal(got_l.codes[c], want.codes[c]) and np.array_equal(got_l.counts[c], want.counts[c]), "library path differs"
        assert np.array_equal(got_lh.codes[c], want.codes[c]) and np.array_equal(got_lh.counts[c], want.counts[c]), "library path (host text) differs"
    with M.Sketcher(perm, k, subk, L, device=local) as sk1:      # whole MarkerDB, whole sketch, one GPU
        want_stats = sk1.composite(mdb.comp, [(want.codes[c], want.counts[c]) for c in range(len(want.codes))])
    assert np.array_equal(stats_l, want_stats) and np.array_equal(stats_lh, want_stats), "sharded composite differs"
    assert int((want_stats["n"] >= 6).sum()) >= 5
    print("MULTI_OK", want.n_total)
dist.barrier()
dist.destroy_process_group()
