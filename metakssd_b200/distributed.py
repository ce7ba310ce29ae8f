"""Multi-GPU plumbing of the sketching path: one process per GPU, reads sharded
by contiguous record ranges, one exchange step.

Each rank reduces its shard to runs (code, first global position, count) sorted by code
(mk_fastq_partial_device).  The code space [0, 2^code_bits) is cut into world_size equal ranges;
a variable-size all-to-all (torch.distributed over NCCL/NVLink on GPUs, gloo in the CPU tests)
delivers every run to the rank that owns its range, which merges them (sum of counts saturating
at 65535, minimum first position).  The merged ranges are sent to rank 0, which reproduces the
reference's hash-slot order (that needs all codes of a component in one table).

torch is plumbing here: device memory views, the process group, the collective.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


class _CudaView:
    """Zero-copy view of library-owned device memory through __cuda_array_interface__."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def device_tensor(ptr: int, n: int, dtype: torch.dtype, device) -> torch.Tensor:
    if n == 0 or not ptr:
        return torch.empty(0, dtype=dtype, device=device)
    typestr = {torch.int64: "<i8", torch.int32: "<i4"}[dtype]
    return torch.as_tensor(_CudaView(ptr, n, typestr), device=device)


def code_range_edges(world: int, code_bits: int) -> torch.Tensor:
    """First code of every rank's range (and 2^code_bits): quantiles of the density 2 (1 - x) the codes of unbiased
    sequence follow (a code leads with the high bases of the smaller of a k-mer and its reverse complement) —
    the same integer formula as range_edge() in csrc/mk_comm.cu."""
    import math
    edges = [0]
    for p in range(1, world):
        x = ((world - p) << (2 * code_bits)) // world
        r = math.isqrt(x)
        if r * r < x:
            r += 1
        edges.append((1 << code_bits) - r)
    edges.append(1 << code_bits)
    return torch.tensor(edges, dtype=torch.int64)


def split_sizes_by_code_range(codes: torch.Tensor, world: int, code_bits: int) -> list:
    """codes: sorted int64 tensor (codes < 2^63).  Number of runs falling into each rank's range."""
    edges = code_range_edges(world, code_bits).to(codes.device)
    cut = torch.searchsorted(codes, edges)
    return (cut[1:] - cut[:-1]).tolist()


def all_to_all_runs(code: torch.Tensor, pos: torch.Tensor, cnt: torch.Tensor, send_sizes: list, group=None):
    """Variable-size all-to-all of three parallel arrays; returns the received (code, pos, cnt).
    The three arrays travel as one (n, 3) int64 buffer: two collectives per exchange (sizes, payload)."""
    world = dist.get_world_size(group)
    dev = code.device
    send = torch.tensor(send_sizes, dtype=torch.int64, device=dev)
    recv = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_to_all_single(recv, send, group=group)
    recv_sizes = recv.tolist()
    n = int(sum(recv_sizes))
    payload = torch.stack((code, pos, cnt.to(torch.int64)), dim=1).contiguous()
    out = torch.empty((n, 3), dtype=torch.int64, device=dev)
    dist.all_to_all_single(out, payload, output_split_sizes=recv_sizes, input_split_sizes=send_sizes, group=group)
    return out[:, 0].contiguous(), out[:, 1].contiguous(), out[:, 2].to(cnt.dtype)


def merge_runs_reference(code: torch.Tensor, pos: torch.Tensor, cnt: torch.Tensor):
    """Host/torch statement of the merge rule (used by the CPU tests and as documentation of
    mk_runs_merge_device): per code, counts add up and saturate at 65535, first position is the
    minimum.  Returns arrays sorted by code."""
    if code.numel() == 0:
        return code, pos, cnt
    order = torch.argsort(code, stable=True)
    c, p, k = code[order], pos[order], cnt[order].to(torch.int64)
    uniq, inv = torch.unique_consecutive(c, return_inverse=True)
    ksum = torch.zeros(uniq.numel(), dtype=torch.int64, device=c.device).scatter_add_(0, inv, k)
    pmin = torch.full((uniq.numel(),), torch.iinfo(torch.int64).max, dtype=torch.int64, device=c.device)
    pmin = pmin.scatter_reduce(0, inv, p, reduce="amin")
    return uniq, pmin, torch.clamp(ksum, max=65535).to(cnt.dtype)


def bind_to_gpu_numa(device_index: int):
    """Pin this process to the CPUs of the NUMA node the GPU hangs off (sysfs), so that pinned host buffers allocated
    afterwards are local to the GPU's PCIe root.  Returns the node (or None when the topology is not exposed)."""
    import os
    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(device_index), "pci_domain_id", 0)
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        bdf = "%04x:%02x:%02x.0" % (dom, bus, dev)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def init_library_comm(sk, group=None):
    """Join the library's own NCCL communicator (csrc/mk_comm.cu): rank 0 draws the unique id, torch.distributed
    (any backend) only carries its 128 bytes to the other ranks."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", sk.info.device) if backend == "nccl" else torch.device("cpu")
    buf = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(sk.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(buf, 0, group=group)
    sk.comm_init(bytes(buf.cpu().numpy().tobytes()), rank, world)


def sketch_sharded(sk, d_text, nbytes: int, pos_base: int, line_base: int, is_last: bool, group=None,
                   host_text: bool = False):
    """The whole multi-GPU step for this rank's shard.  Returns the final Sketch on rank 0 (None
    elsewhere).  `sk` is this rank's Sketcher; d_text its shard in device memory (or, with
    host_text=True, in host memory: uploaded chunk by chunk under the kernel)."""
    import os, time
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device("cuda", sk.info.device)
    _t = [time.perf_counter()] if os.environ.get("MK_TIMING") else None

    def _mark(what):
        if _t is not None:
            torch.cuda.synchronize(dev)
            now = time.perf_counter()
            print("[mk timing r%d] %-18s %7.3f ms" % (rank, what, (now - _t[0]) * 1e3), flush=True)
            _t[0] = now
    runs = (sk.fastq_partial_host if host_text else sk.fastq_partial_device)(d_text, nbytes, pos_base, line_base, is_last)
    _mark("partial")
    n = int(runs.n)
    code = device_tensor(runs.d_code, n, torch.int64, dev)
    pos = device_tensor(runs.d_firstpos, n, torch.int64, dev)
    cnt = device_tensor(runs.d_count, n, torch.int32, dev)
    sizes = split_sizes_by_code_range(code, world, sk.info.code_bits)
    _mark("split sizes")
    rc, rp, rk = all_to_all_runs(code, pos, cnt, sizes, group)
    torch.cuda.current_stream(dev).synchronize()
    _mark("all_to_all")
    merged = sk.runs_merge_device(rc, rp, rk, int(rc.numel()))
    _mark("merge")
    m = int(merged.n)
    mc = device_tensor(merged.d_code, m, torch.int64, dev)
    mp = device_tensor(merged.d_firstpos, m, torch.int64, dev)
    mk = device_tensor(merged.d_count, m, torch.int32, dev)
    to_root = [m] + [0] * (world - 1)
    gc, gp, gk = all_to_all_runs(mc, mp, mk, to_root, group)
    torch.cuda.current_stream(dev).synchronize()
    _mark("gather to root")
    if rank != 0:
        return None
    out = sk.runs_finalize_device(gc, gp, gk, int(gc.numel()), distinct=True)   # owners' ranges are disjoint
    _mark("finalize")
    return out


def size_exchange_blocks(sk, text, nbytes: int, pos_base: int, line_base: int, is_last: bool, host_text: bool = False,
                         group=None) -> int:
    """Collective.  Block capacity (`max_runs`) for Sketcher.fastq_koc_sharded() on batches like this one: one sizing
    pass of the step; a capacity that turns out too small fails on every rank together (MK_ERR_NOMEM), the ranks
    agree on the size that was needed and repeat.  Returns the largest block any rank saw plus 1/8 headroom — a
    pipeline keeps using the value for its next batches and comes back here when a step reports MK_ERR_NOMEM."""
    from .api import MkError
    world = dist.get_world_size(group)
    dev = torch.device("cuda", sk.device)
    part = sk.fastq_partial_host(text, nbytes, pos_base, line_base, is_last) if host_text else \
        sk.fastq_partial_device(text, nbytes, pos_base, line_base, is_last)
    t = torch.tensor([int(part.n)], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    cap = int(t.item()) * 3 // (2 * world) + 4096            # 1.5 x the even share of the fullest shard
    for _ in range(4):
        try:
            sk.fastq_koc_sharded(text, nbytes, pos_base, line_base, is_last, cap, host_text=host_text, want_stats=False)
            failed = False
        except MkError as e:
            if e.code != -4:                                 # MK_ERR_NOMEM
                raise
            failed = True
        t = torch.tensor([sk.comm_last_block_need()], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        cap = int(t.item()) * 9 // 8 + 1024
        if not failed:
            return cap
    raise RuntimeError("exchange block capacity did not settle")


def balanced_shares(per_rank: int, world: int, waits_ms, stream_ms_per_read: float, moved_so_far: int = 0):
    """Reads per rank for the sharded step when rank 0 carries the tail (slot order of the gathered sketch, statistics).
    `waits_ms[r]` = mk_profile.exchange_wait_ms per step of rank r in steady state: the other ranks' wait minus
    rank 0's own is what rank 0's tail costs; rank 0 gets fewer reads by what streams in that time (times
    (world-1)/world, because the reads it gives up make the others slower), the total stays world * per_rank.
    Returns (shares, moved): `moved` = reads taken from rank 0 so far (pass it back in to refine)."""
    if world < 2:
        return [per_rank], 0
    tail_ms = sum(waits_ms[1:]) / (world - 1) - waits_ms[0]
    moved = moved_so_far + int(tail_ms / stream_ms_per_read * (world - 1) / world)
    moved = max(0, min(moved, per_rank * 2 // 5))
    shares = [per_rank - moved] + [per_rank + moved // (world - 1)] * (world - 1)
    shares[-1] += world * per_rank - sum(shares)
    return shares, moved
