"""`dist -r` throughput probe (development tool): shared k-mer counts of n_qry query sketches against a database of
n_ref genome sketches (species clusters of 10 strains sharing 95 % of their codes), device vs the oracle's CPU loop."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import metakssd_b200 as M
import oracle as O

n_species = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
strains, per, n_qry = 10, 1200, int(sys.argv[2]) if len(sys.argv) > 2 else 2000
rng = np.random.default_rng(4)
base = [np.unique(rng.integers(0, 2 ** 32, size=per, dtype=np.uint64)).astype(np.uint32) for _ in range(n_species)]


def strain(b):
    s = b.copy()
    idx = rng.integers(0, s.size, size=s.size // 20)
    s[idx] = rng.integers(0, 2 ** 32, size=idx.size, dtype=np.uint64).astype(np.uint32)
    return np.unique(s)


def side(sks):
    idx = np.zeros(len(sks) + 1, dtype=np.uint64)
    idx[1:] = np.cumsum([s.size for s in sks])
    return np.concatenate(sks), idx


ref = side([strain(base[i // strains]) for i in range(n_species * strains)])
qry = side([strain(base[int(rng.integers(0, n_species))]) for _ in range(n_qry)])
print("reference: %d sketches, %.1f M codes; query: %d sketches, %.1f M codes; matrix %d x %d" % (
    ref[1].size - 1, ref[0].size / 1e6, n_qry, qry[0].size / 1e6, n_qry, ref[1].size - 1))
with M.Sketcher(None, 11, 6, 3) as sk:
    for it in range(3):
        t = time.perf_counter(); got = sk.shared_counts([ref], [qry]); dt = time.perf_counter() - t
        print("device: %.1f ms  (%.1f M query codes/s, %d shared pairs)" % (dt * 1e3, qry[0].size / dt / 1e6, int(got.sum())))
t = time.perf_counter(); want = O.shared_counts([ref], [qry]); dt = time.perf_counter() - t
print("oracle (1 thread): %.1f ms; equal: %s" % (dt * 1e3, bool(np.array_equal(got, want))))
