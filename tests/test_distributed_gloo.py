"""CPU, world_size 2 over gloo: the exchange step of the multi-GPU path (partition of runs by code
range, variable all-to-all, merge rule) on host tensors."""
import os
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys
    sys.path.insert(0, %r)
    import numpy as np, torch, torch.distributed as dist
    from metakssd_b200 import distributed as D
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    code_bits = 32
    rng = np.random.default_rng(100 + rank)
    # each rank: runs sorted by code, some codes shared with the other rank, some counts saturated
    common = np.random.default_rng(5).choice(1 << 32, size=3000, replace=False)
    own = rng.choice(1 << 32, size=5000, replace=False)
    codes = np.unique(np.concatenate([common[rank::1][: 2000 + 500 * rank], own])).astype(np.int64)
    pos = (rng.integers(0, 1 << 40, size=codes.size) * 2 + rank).astype(np.int64)
    cnt = rng.integers(1, 70000, size=codes.size).clip(max=65535).astype(np.int32)
    code_t, pos_t, cnt_t = map(torch.from_numpy, (codes, pos, cnt))
    sizes = D.split_sizes_by_code_range(code_t, world, code_bits)
    assert sum(sizes) == codes.size
    rc, rp, rk = D.all_to_all_runs(code_t, pos_t, cnt_t, sizes)
    edges = D.code_range_edges(world, code_bits)
    assert bool(((rc >= edges[rank]) & (rc < edges[rank + 1])).all()), "run delivered to the wrong owner"
    mc, mp, mk = D.merge_runs_reference(rc, rp, rk)
    # everything to rank 0
    gc, gp, gk = D.all_to_all_runs(mc, mp, mk, [mc.numel()] + [0] * (world - 1))
    # reference result: gather the raw runs of every rank and merge in one place
    allc = [None] * world; allp = [None] * world; allk = [None] * world
    dist.all_gather_object(allc, codes); dist.all_gather_object(allp, pos); dist.all_gather_object(allk, cnt)
    if rank == 0:
        wc, wp, wk = D.merge_runs_reference(*(torch.from_numpy(np.concatenate(x)) for x in (allc, allp, allk)))
        order = torch.argsort(gc)
        assert torch.equal(gc[order], wc) and torch.equal(gp[order], wp) and torch.equal(gk[order], wk)
        assert int(wk.max()) == 65535 and wc.numel() < sum(len(x) for x in allc)
        print("GLOO_OK", wc.numel())
    dist.barrier()
    dist.destroy_process_group()
''')


def test_exchange_world2_gloo(tmp_path, lib_built):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    port = 29500 + (os.getpid() % 400)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "GLOO_OK" in r.stdout
