#!/usr/bin/env python
"""bench.py — L3K11 `dist -A` sketching + `composite` throughput on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1..5]

One step = one pass of the hot path over the whole synthetic input of this rank.  The driver default is
BASELINE.json configs[1] (`--config 2`: 40 M interleaved 150-bp records = 6 Gbp per GPU, 1 000-species
MarkerDB): FASTQ text -> sketch codes + counts in reference slot order -> per-species coverage
statistics.  Prints ONE JSON line (rank 0).  The other configs of BASELINE.json:
  1  1 M reads, 100 species x 1 Mbp (the reference's CPU-runnable case)
  3  200 M reads (30 Gbp) in total, sharded over the ranks (strong scaling)
  4  config 2's reads with the L2K11 geometry (k = 11, subk = 5, L = 2: 16 components, 2^-8 sampling)
  5  10 000 genomes x 5 Mbp FASTA (50 Gbp, no -A) in total, genomes sharded over the ranks, no exchange

  parity    every line carries its own proof, checked BEFORE timing (exit status 3 on a mismatch): the sketch
            of a prefix against the oracle (codes, counts, order); at N > 1 the sharded sketch of reduced
            shards against the single-GPU sketch of their concatenation; on the CPU-baseline sample the
            reference binary's sketch directory and `composite` output against ours; the MarkerDB of a species
            subset against the reference's own `dist` -> `set -g` -> `set -q` -> `set -i` pipeline

  value     whole-job Gbp/s with the FASTQ text (and the MarkerDB) already resident in HBM
  e2e       same through the host-buffer entry point (pinned host text uploaded chunk by chunk under the
            kernel, MarkerDB uploaded with every step, statistics read back): H2D inside the timed region
  roofline  dominant kernel (k_stream_ws): algorithmic bytes (sequence bytes + their newlines, SURVEY
            §8(d)) / average launch duration measured with CUDA events on the library's stream,
            against the measured HBM copy peak (MEASURED_PEAKS.json)
  cpu_baseline  the reference binary (oracle/_ref/metakssd, all host cores) on a bounded sample

`--impl reference` times the unmodified reference binary on bounded samples of the same workload.
With N > 1 (torchrun) every rank sketches its own 6 Gbp shard (weak scaling); runs are exchanged by
code range with an NCCL all-to-all and rank 0 orders and reports.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SEED = 0x4D4B5353
READ_LEN = 150
CONFIGS = {      # BASELINE.json configs, 1-based
    1: dict(k=11, subk=6, L=3, reads=1_000_000, species=100, genome_len=1_000_000, scaling="weak", kind="fastq"),
    2: dict(k=11, subk=6, L=3, reads=40_000_000, species=1000, genome_len=5_000_000, scaling="weak", kind="fastq"),
    3: dict(k=11, subk=6, L=3, reads=200_000_000, species=1000, genome_len=5_000_000, scaling="strong", kind="fastq"),
    4: dict(k=11, subk=5, L=2, reads=40_000_000, species=1000, genome_len=5_000_000, scaling="weak", kind="fastq"),
    5: dict(k=11, subk=6, L=3, reads=0, species=10_000, genome_len=5_000_000, scaling="strong", kind="fasta"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json config (driver default: 2)")
    ap.add_argument("--reads", type=int, default=None, help="FASTQ records (per GPU for weak-scaling configs)")
    ap.add_argument("--species", type=int, default=None)
    ap.add_argument("--genome-len", type=int, default=None)
    ap.add_argument("--cpu-reads", type=int, default=8_000_000, help="records of the CPU-baseline / cli_e2e sample")
    ap.add_argument("--parity-reads", type=int, default=400_000, help="records of the prefix checked against the oracle")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-balance", action="store_true", help="N > 1: equal shards (rank 0 is not given fewer reads for its tail)")
    ap.add_argument("--no-parity", action="store_true")
    a = ap.parse_args()
    c = CONFIGS[a.config]
    a.k, a.subk, a.L, a.kind, a.scaling = c["k"], c["subk"], c["L"], c["kind"], c["scaling"]
    a.reads = a.reads if a.reads is not None else c["reads"]
    a.species = a.species if a.species is not None else c["species"]
    a.genome_len = a.genome_len if a.genome_len is not None else c["genome_len"]
    a.geom = "L%dK%d" % (a.L, a.k)
    a.metric = ("%s -A sketching + composite throughput" % a.geom) if a.kind == "fastq" else \
               ("%s genome (FASTA) sketching throughput" % a.geom)
    return a


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: NVML from a thread every few ms
    (the timed region of the device-resident leg lasts tens of ms), nvidia-smi as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, None
        self.thread, self.stop_flag, self.sm, self.mx, self.reasons = None, False, [], [], set()

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        except Exception:
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)

    def _loop(self, nv, h):
        R = nv
        bits = (("hw_slowdown", R.nvmlClocksThrottleReasonHwSlowdown),
                ("hw_thermal_slowdown", R.nvmlClocksThrottleReasonHwThermalSlowdown),
                ("sw_thermal_slowdown", R.nvmlClocksThrottleReasonSwThermalSlowdown),
                ("sw_power_cap", R.nvmlClocksThrottleReasonSwPowerCap))
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for name, bit in bits:
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.003)

    def start(self):
        try:
            import threading
            nv, h = self._nvml_handle()
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._loop, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.path = tempfile.mktemp(prefix="clocks_", suffix=".csv")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.thread:
            self.stop_flag = True
            self.thread.join(timeout=2)
            if self.sm:
                out = {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": max(self.mx),
                       "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml"}
            return out
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 6:
                    continue
                try:
                    sm.append(float(f[0])); mx.append(float(f[1]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm), "source": "nvidia-smi"}
        return out


def workload_name(args, world):
    if args.kind == "fasta":
        return ("config %d: %d synthetic genomes x %.1f Mbp FASTA (%.1f Gbp in total, no -A), %s sketches, genomes "
                "sharded over %d GPU(s), no exchange" % (args.config, args.species, args.genome_len / 1e6,
                                                         args.species * args.genome_len / 1e9, args.geom, world))
    per = "per GPU" if args.scaling == "weak" else "in total, sharded by reads"
    return ("config %d: synthetic %g M x %d bp interleaved paired-end FASTQ %s (%.2f Gbp), %s -A sketch + composite vs "
            "%d-species synthetic MarkerDB" % (args.config, args.reads / 1e6, READ_LEN, per, args.reads * READ_LEN / 1e9,
                                               args.geom, args.species))


class ParityError(SystemExit):
    def __init__(self, what):
        sys.stderr.write("bench.py: PARITY MISMATCH: %s\n" % what)
        super().__init__(3)


def _same_sketch(a, b) -> bool:
    """two api.Sketch objects: same codes, counts and order in every component"""
    if len(a.codes) != len(b.codes):
        return False
    for c in range(len(a.codes)):
        if not np.array_equal(a.codes[c], b.codes[c]):
            return False
        if (a.counts is None) != (b.counts is None) or (a.counts is not None and not np.array_equal(a.counts[c], b.counts[c])):
            return False
    return True


def _oracle_matches(sketch, want, p) -> bool:
    comps = want.components(p)
    if len(comps) != len(sketch.codes):
        return False
    for c, (codes, counts) in enumerate(comps):
        if not np.array_equal(sketch.codes[c], codes):
            return False
        if counts is not None and not np.array_equal(sketch.counts[c], counts):
            return False
    return True


def _ref_bin():
    p = os.path.join(ROOT, "oracle", "_ref", "metakssd")
    return p if os.path.exists(p) else None


def _cli_bin():
    p = os.path.join(ROOT, "host", "metakssd-b200")
    if not os.path.exists(p):
        try:
            subprocess.run(["make", "-C", os.path.join(ROOT, "host")], check=True, capture_output=True)
        except Exception:
            return None
    return p if os.path.exists(p) else None


def _tsv_rows(text):
    """species_coverage rows without the query-name column, as a sorted list (the reference's order of equal hit
    counts is the qsort's; with -p > 1 its dictionary is still order independent)"""
    return sorted("\t".join(l.split("\t")[1:]) for l in text.splitlines() if l.count("\t") >= 6)


# =====================================================================================================
def run_ours(args):
    import torch
    import torch.distributed as dist
    import metakssd_b200 as M
    from metakssd_b200 import workload as W
    from metakssd_b200 import distributed as D

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, SUBK, L = args.k, args.subk, args.L
    shuf_id, perm = M.make_shuf(SEED ^ 1, SUBK)
    sk = M.Sketcher(perm, K, SUBK, L, device=local)
    spec = M.synth_spec(SEED ^ 2, args.species, args.genome_len, READ_LEN)
    if args.kind == "fasta":
        return run_fasta(args, sk, spec, shuf_id, perm, rank, world, local, dev)
    t0 = time.time()
    mdb = W.build_markerdb(sk, spec)
    t_mdb = time.time() - t0

    per_rank = args.reads if args.scaling == "weak" else args.reads // world
    r0, r1 = rank * per_rank, (rank + 1) * per_rank
    nbytes = spec.fastq_bytes(r0, r1)
    pos_base = spec.fastq_bytes(0, r0)
    d_text = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
    sk.synth_fastq_device(spec.P, spec.cdf32, spec.species, r0, r1, d_text, d_text.numel())
    bases = per_rank * READ_LEN
    algo_bytes = per_rank * (READ_LEN + 1)          # sequence bytes + their '\n' 
    lib_stream = torch.cuda.ExternalStream(sk.cuda_stream(), device=dev)

    # `value` leg: everything the step reads is resident in HBM (FASTQ text and MarkerDB);
    # `e2e` leg: host buffers, the MarkerDB is uploaded with every step like the reference re-reads it
    use_lib_comm = world > 1 and os.environ.get("MK_DIST", "nccl-lib") != "torch"
    if use_lib_comm:       # exchange, owner merge and rank-local composite inside the library (csrc/mk_comm.cu)
        D.init_library_comm(sk)
        sk.load_markerdb_sharded(mdb.comp)
    elif rank == 0:
        sk.load_markerdb(mdb.comp)
    names_c = M.SpeciesNames(mdb.names)
    max_runs = 0
    if use_lib_comm:
        # block capacity of the exchange (runs one rank sends to one owner / one owner holds after the merge): one
        # untimed sizing pass of the step, the largest block any rank saw + 1/8 (a pipeline keeps the value for its
        # next batches).  A block that does not fit fails the step on every rank (MK_ERR_NOMEM), nothing is truncated.
        max_runs = D.size_exchange_blocks(sk, d_text, nbytes, pos_base, 4 * r0, rank == world - 1)

    def sharded_step(text, nb, pb, lb, last, host_text=False):
        """(sketch, tsv) on rank 0, (None, None) elsewhere"""
        if use_lib_comm:
            s_, st = sk.fastq_koc_sharded(text, nb, pb, lb, last, max_runs, host_text=host_text)
            return (s_, M.coverage_tsv("reads.fq", names_c, st)) if rank == 0 else (None, None)
        s_ = D.sketch_sharded(sk, text, nb, pb, lb, last, host_text=host_text)
        return (s_, composite(s_, not host_text)) if rank == 0 else (None, None)

    def composite(sketch, resident):
        if resident:        # MarkerDB and the sketch just produced are both on the device
            stats = sk.composite_last()
        else:
            qry = [(sketch.codes[c], sketch.counts[c]) for c in range(len(sketch.codes))]
            stats = sk.composite(mdb.comp, qry)
        return M.coverage_tsv("reads.fq", names_c, stats)

    # ---- parity, before anything is timed ----------------------------------------------------------
    parity = {}
    if not args.no_parity:
        import oracle as O
        p = O.params(K, SUBK, L)
        n_par = min(args.parity_reads, per_rank)
        nb_par = spec.fastq_bytes(r0, r0 + n_par)
        got = sk.fastq_koc_device(d_text, nb_par)
        if rank == 0:       # the oracle is single threaded: one rank checks its prefix against it
            want = O.fastq_koc(p, perm, d_text[:nb_par].cpu().numpy())
            parity["prefix_vs_oracle"] = {"ok": _oracle_matches(got, want, p), "records": n_par, "codes": int(got.n_total)}
            if not parity["prefix_vs_oracle"]["ok"]:
                raise ParityError("prefix sketch differs from the oracle")
        if world > 1:
            # reduced shards: the first n_sh records of every rank's shard; rank 0 also sketches their concatenation
            n_sh = min(200_000, per_rank)
            nb_sh = spec.fastq_bytes(r0, r0 + n_sh)
            sizes = [spec.fastq_bytes(q * per_rank, q * per_rank + n_sh) for q in range(world)]
            s_sh, tsv_sh = sharded_step(d_text, nb_sh, sum(sizes[:rank]), 4 * n_sh * rank, rank == world - 1)
            if rank == 0:
                cat = torch.empty(sum(sizes) + 256, dtype=torch.uint8, device=dev)
                o = 0
                for q in range(world):
                    sk.synth_fastq_device(spec.P, spec.cdf32, spec.species, q * per_rank, q * per_rank + n_sh, cat[o:], cat.numel() - o)
                    o += sizes[q]
                single = sk.fastq_koc_device(cat, sum(sizes))
                del cat
                if use_lib_comm:      # the single-GPU composite needs the whole MarkerDB: a second context holds it
                    with M.Sketcher(perm, K, SUBK, L, device=local) as sk1:
                        qry = [(single.codes[c], single.counts[c]) for c in range(len(single.codes))]
                        tsv_single = M.coverage_tsv("reads.fq", names_c, sk1.composite(mdb.comp, qry))
                else:
                    tsv_single = composite(single, False)
                parity["sharded_vs_single"] = {"ok": _same_sketch(s_sh, single) and tsv_sh == tsv_single,
                                               "records_per_rank": n_sh, "ranks": world, "codes": int(single.n_total),
                                               "coverage_rows": tsv_single.count("\n"),
                                               "path": "library NCCL exchange + rank-local composite" if use_lib_comm else "torch.distributed all-to-all"}
            flag = torch.tensor([1 if rank != 0 or parity["sharded_vs_single"]["ok"] else 0], device=dev)
            dist.broadcast(flag, 0)                 # every rank leaves together on a mismatch
            if int(flag.item()) == 0:
                raise ParityError("sharded sketch differs from the single-GPU sketch of the same records")

    def step_device():
        if world == 1:
            return composite(sk.fastq_koc_device(d_text, nbytes), True)
        return sharded_step(d_text, nbytes, pos_base, 4 * r0, rank == world - 1)[1]

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(lib_stream)
        out = None
        for _ in range(steps):
            out = fn()
        e1.record(lib_stream)
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out

    # timed legs: the sketch arrives in the library's pinned block and is used from there (no per-component copies on
    # the host; the C host program writes its files from the same arrays)
    sk.set_borrowed_output(True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total, tsv = timed(step_device, 0, args.warmup)      # warm-up outside the profile window
    balance = None
    my_reads = per_rank
    eq = None
    if use_lib_comm and not args.no_balance and not args.no_e2e:
        # the end-to-end leg is bound by the upload, for which equal shards are best: keep the equal shard on the host
        import psutil
        if psutil.virtual_memory().available > nbytes * world * 1.15 + (8 << 30):
            h_eq = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
            h_eq.copy_(d_text[:nbytes])
            torch.cuda.synchronize(dev)
            eq = {"h_text": h_eq, "nbytes": nbytes, "pos_base": pos_base, "r0": r0, "reads": per_rank}
    if use_lib_comm and not args.no_balance:
        # Rank 0 alone turns the gathered runs into the sketch (slot order needs one table) while the other ranks
        # are already streaming their next shard, and then everybody waits for it at the next exchange.  So rank 0
        # gets a shard that is shorter by what its tail costs: measured here as the time the other ranks spend in
        # the first exchange beyond rank 0's own (mk_profile.exchange_wait_ms), over three steady-state steps.
        # The total number of reads stays world x reads_per_gpu.
        d_total, history = 0, []
        for it in range(2):         # the second pass corrects what is left after the first
            sk.profile(reset=True)
            timed(step_device, 3, 0)
            pw = sk.profile()
            w = torch.tensor([pw.exchange_wait_ms / 3, pw.stream_kernel_ms / 3], dtype=torch.float64, device=dev)
            allw = [torch.zeros_like(w) for _ in range(world)]
            dist.all_gather(allw, w)
            waits = [float(x[0]) for x in allw]
            stream_ms_per_read = sum(float(x[1]) for x in allw) / (world * per_rank)
            tail_ms = sum(waits[1:]) / (world - 1) - waits[0]      # > 0: the others wait for rank 0
            history.append({"exchange_wait_ms_per_rank": waits, "rank0_tail_ms": tail_ms})
            shares, d_total = D.balanced_shares(per_rank, world, waits, stream_ms_per_read, d_total)
            starts = [sum(shares[:q]) for q in range(world)]
            r0, r1 = starts[rank], starts[rank] + shares[rank]
            my_reads = shares[rank]
            nbytes = spec.fastq_bytes(r0, r1)
            pos_base = spec.fastq_bytes(0, r0)
            d_text = None
            torch.cuda.empty_cache()
            d_text = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
            sk.synth_fastq_device(spec.P, spec.cdf32, spec.species, r0, r1, d_text, d_text.numel())
            max_runs = D.size_exchange_blocks(sk, d_text, nbytes, pos_base, 4 * r0, rank == world - 1)
            algo_bytes = my_reads * (READ_LEN + 1)
            timed(step_device, 0, args.warmup)
        balance = {"reads_per_rank": shares, "passes": history,
                   "what": "rank 0 (slot order of the gathered sketch, statistics) gets fewer reads so that all ranks reach "
                           "the next exchange together; total reads unchanged"}
    sk.profile(reset=True)
    ms_total, tsv = timed(step_device, args.steps, 0)
    prof = sk.profile()
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = world * bases / 1e9 / (ms_step / 1e3)

    # ---- end to end: pinned host text -> H2D -> sketch -> composite -> statistics on the host -----
    e2e = None
    if not args.no_e2e:
        import psutil
        need = nbytes * world * 1.15
        e2e_reads = my_reads
        if psutil.virtual_memory().available < need + (8 << 30):
            e2e_reads = max(1_000_000, int(my_reads * (psutil.virtual_memory().available * 0.5) / need))
            e2e_reads = min(e2e_reads, my_reads)
        e2e_total = e2e_reads
        if world > 1:
            t = torch.tensor([e2e_reads], dtype=torch.int64, device=dev)
            dist.all_reduce(t)
            e2e_total = int(t.item())
        numa = D.bind_to_gpu_numa(local)          # the pinned buffer is first touched on the GPU's own NUMA node
        e_r0, e_pos_base = r0, pos_base
        if eq is not None:                        # equal shards (see above); the exchange blocks are sized again for them
            h_text, e_nbytes, e_pos_base, e_r0, e2e_reads = eq["h_text"], eq["nbytes"], eq["pos_base"], eq["r0"], eq["reads"]
            e2e_total = world * e2e_reads
            max_runs = D.size_exchange_blocks(sk, h_text, e_nbytes, e_pos_base, 4 * e_r0, rank == world - 1, host_text=True)
        else:
            e_nbytes = spec.fastq_bytes(r0, r0 + e2e_reads)
            h_text = torch.empty(e_nbytes, dtype=torch.uint8, pin_memory=True)
            h_text.copy_(d_text[:e_nbytes])
            torch.cuda.synchronize(dev)

        def step_host():
            if world == 1:
                return composite(sk.fastq_koc_host(h_text), False)
            # multi-GPU: every rank uploads its shard from its own pinned buffer (chunks overlapped with
            # the kernel), then the sharded path
            return sharded_step(h_text, e_nbytes, e_pos_base, 4 * e_r0, rank == world - 1, host_text=True)[1]

        # what the host side can deliver: every rank copies its pinned shard to its GPU at the same time, nothing else
        # running (the end-to-end step cannot be faster than this copy; at N > 1 the GPUs share PCIe switches / host DRAM)
        probe_bytes = min(e_nbytes, 2 << 30)        # (slices of the pinned text into one 2 GB device buffer)
        d_probe = torch.empty(probe_bytes, dtype=torch.uint8, device=dev)

        def copy_only():
            for o in range(0, e_nbytes, probe_bytes):
                m = min(probe_bytes, e_nbytes - o)
                d_probe[:m].copy_(h_text[o:o + m], non_blocking=True)

        cur = torch.cuda.current_stream(dev)
        copy_only()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(cur)
        copy_only(); copy_only()
        c1.record(cur)
        torch.cuda.synchronize(dev)
        ms_copy = c0.elapsed_time(c1) / 2
        if world > 1:
            t = torch.tensor([ms_copy], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_copy = float(t.item())
        del d_probe
        torch.cuda.empty_cache()

        e_steps = max(2, min(args.steps, 3))
        timed(step_host, 0, 1)
        sk.profile(reset=True)
        ms_e, tsv_e = timed(step_host, e_steps, 0)
        pe = sk.profile()
        ms_e /= e_steps
        e2e = {"value": e2e_total * READ_LEN / 1e9 / (ms_e / 1e3), "unit": "Gbp/s",
               "h2d_bytes_per_step": int(pe.h2d_bytes // e_steps), "d2h_bytes_per_step": int(pe.d2h_bytes // e_steps),
               "bytes_source": "counted by the library per copy (mk_profile), this rank",
               "ms_per_step": ms_e, "reads_per_gpu": e2e_reads, "shards": "equal" if (eq is not None or balance is None) else "as in the value leg",
               "numa_node_of_rank0": numa,
               "h2d_copy_only": {"ms": ms_copy, "GBps_per_gpu": e_nbytes / 1e6 / ms_copy,
                                 "GBps_all_gpus": world * e_nbytes / 1e6 / ms_copy,
                                 "what": "all ranks copy their pinned shard to their GPU at once, max over ranks; "
                                         "the floor of ms_per_step on this host"}}
        del h_text

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    launches = max(1, int(prof.stream_kernel_launches))
    k_ms = prof.stream_kernel_ms / launches
    achieved = algo_bytes / 1e9 / (k_ms / 1e3)
    traffic, traffic_src = None, None
    try:    # DRAM bytes per text byte from the ncu --set full capture of this build (profiles/, not measured in this run)
        tj = json.load(open(os.path.join(ROOT, "profiles", "r2_stream_traffic.json")))
        traffic = tj["dram_bytes_per_text_byte"] * nbytes
        traffic_src = "ncu capture %s scaled to this launch's text bytes (not measured in this run)" % tj.get("source", "profiles/")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": os.environ.get("MK_STREAM_IMPL", "ws") == "v3" and "k_stream3" or "k_stream_ws",
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "kernel_ms_per_launch": k_ms, "algorithmic_bytes_per_launch": algo_bytes,
                "text_bytes_per_launch": nbytes, "kernel_share_of_step": k_ms / ms_step,
                "kernel_Gbp_s": my_reads * READ_LEN / 1e9 / (k_ms / 1e3)}

    cpu = cli = None
    if world == 1 and not args.no_cpu_baseline:
        cpu, cli, par2 = cpu_baseline(args, sk, spec, shuf_id, perm, mdb, d_text, composite)
        parity.update(par2)
    if any(isinstance(v, dict) and v.get("ok") is False for v in parity.values()):
        raise ParityError(json.dumps(parity))
    parity["all_ok"] = all(v.get("ok") for v in parity.values() if isinstance(v, dict)) if parity else None

    line = {
        "metric": args.metric, "value": value, "unit": "Gbp/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u64",
        "data": "synthetic",
        "config": {"workload": workload_name(args, world), "config": args.config, "k": K, "subk": SUBK, "L": L,
                   "reads_per_gpu": per_rank, "shard_balance": balance,
                   "read_len": READ_LEN, "species": args.species, "genome_len": args.genome_len,
                   "markerdb_codes": mdb.n_codes, "l2_policy": "input (%.1f GB per GPU) is far larger than L2" % (nbytes / 1e9),
                   "parallelism": ("reads sharded per GPU; runs exchanged by code range in one grouped ncclSend/ncclRecv step inside the "
                                   "library, MarkerDB sharded on the same boundaries (rank-local composite), slot order on rank 0"
                                   if use_lib_comm else "reads sharded per GPU, runs exchanged by code range (torch all-to-all)") if world > 1 else "1 GPU",
                   "markerdb_build_s": t_mdb, "species_reported": tsv.count("\n") if tsv else 0},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "cli_e2e": cli, "parity": parity,
        "gpu_launches": int(prof.kernel_launches), "clocks": clocks,
        "breakdown_ms_per_step": {"stream_kernel": prof.stream_kernel_ms / args.steps, "reduce_order": prof.reduce_ms / args.steps,
                                  "composite": prof.composite_ms / args.steps, "exchange": prof.exchange_ms / args.steps},
    }
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# =====================================================================================================
def run_fasta(args, sk, spec, shuf_id, perm, rank, world, local, dev):
    """config 5: genome (FASTA) sketching, genomes sharded over the ranks, no exchange (command_dist.c:365,397-398)."""
    import torch
    import torch.distributed as dist
    import metakssd_b200 as M
    K, SUBK, L = args.k, args.subk, args.L
    G = args.species
    g0, g1 = rank * G // world, (rank + 1) * G // world
    per_file = spec.fasta_bytes(0) + 16
    # batches of at most ~4 Gi bases of text per call
    per_batch = max(1, min(g1 - g0, (3 << 30) // per_file))
    bufs, offs = [], []
    for s0 in range(g0, g1, per_batch):
        s1 = min(g1, s0 + per_batch)
        b = torch.empty((s1 - s0) * per_file + 256, dtype=torch.uint8, device=dev)
        offs.append(sk.synth_fasta_device(spec.P, s0, s1, b, b.numel()))
        bufs.append(b)
    text_bytes = int(sum(int(o[-1]) for o in offs))
    bases = (g1 - g0) * args.genome_len
    algo_bytes = bases * 81 // 80                     # 80-column FASTA: one '\n' per 80 bases
    lib_stream = torch.cuda.ExternalStream(sk.cuda_stream(), device=dev)
    parity = {}
    if not args.no_parity and rank == 0:
        import oracle as O
        p = O.params(K, SUBK, L)
        n_par = min(6, g1 - g0)
        got = sk.fasta_co_device(bufs[0], offs[0][:n_par + 1])
        ok = True
        for i in range(n_par):
            t = bufs[0][int(offs[0][i]):int(offs[0][i + 1])].cpu().numpy()
            ok = ok and _oracle_matches(got[i], O.fasta_co(p, perm, t), p)
        parity["genomes_vs_oracle"] = {"ok": bool(ok), "genomes": n_par}
        if not ok:
            raise ParityError("genome sketches differ from the oracle")

    def step():
        n = 0
        for b, o in zip(bufs, offs):
            n += sum(s.n_total for s in sk.fasta_co_device(b, o))
        return n

    for _ in range(args.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    sk.profile(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(lib_stream)
    for _ in range(args.steps):
        ncodes = step()
    e1.record(lib_stream)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    prof = sk.profile()
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms / args.steps
    # e2e: host text (pinned) of one batch through mk_fasta_co_host, results read back
    e2e = None
    if not args.no_e2e:
        nb = int(offs[0][-1])
        h = torch.empty(nb, dtype=torch.uint8, pin_memory=True)
        h.copy_(bufs[0][:nb])
        torch.cuda.synchronize(dev)
        Lc = M.load()
        import ctypes as C
        from metakssd_b200 import api as A
        off = np.ascontiguousarray(offs[0], dtype=np.uint64)
        nf = off.size - 1

        def host_step():
            arr = (A.MkSketch * nf)()
            rc = Lc.mk_fasta_co_host(sk._h, int(h.data_ptr()), off.ctypes.data, nf, arr)
            assert rc == 0
            for i in range(nf):
                Lc.mk_sketch_free(C.byref(arr[i]))
        host_step()
        sk.profile(reset=True)
        t0 = time.perf_counter()
        for _ in range(2):
            host_step()
        dt = (time.perf_counter() - t0) / 2
        pe = sk.profile()
        e2e = {"value": world * nf * args.genome_len / 1e9 / dt, "unit": "Gbp/s", "h2d_bytes_per_step": int(pe.h2d_bytes // 2),
               "d2h_bytes_per_step": int(pe.d2h_bytes // 2), "ms_per_step": dt * 1e3,
               "sample": "one batch of %d genomes per rank from pinned host memory" % nf}
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peak()
    launches = max(1, int(prof.stream_kernel_launches))
    k_ms = prof.stream_kernel_ms / launches
    # the FASTA path reads the text in the tokenizer passes and the dense stream in the stream kernel: the roofline
    # is taken over the whole step (all kernels of one batch), against the algorithmic bytes of the text
    achieved = algo_bytes / 1e9 / (ms_step / 1e3)
    cpu = None
    if world == 1 and not args.no_cpu_baseline and _ref_bin():
        cpu = cpu_baseline_fasta(args, sk, spec, shuf_id, perm, bufs[0], offs[0])
        if cpu and "parity" in cpu:
            parity.update(cpu.pop("parity"))
    if any(isinstance(v, dict) and v.get("ok") is False for v in parity.values()):
        raise ParityError(json.dumps(parity))
    parity["all_ok"] = all(v.get("ok") for v in parity.values() if isinstance(v, dict)) if parity else None
    line = {"metric": args.metric, "value": world * bases / 1e9 / (ms_step / 1e3), "unit": "Gbp/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_name(args, world), "config": args.config, "k": K, "subk": SUBK, "L": L,
                       "genomes_per_gpu": g1 - g0, "genome_len": args.genome_len, "batches_per_step": len(bufs),
                       "codes_per_step": int(ncodes), "l2_policy": "input (%.1f GB per GPU) is far larger than L2" % (text_bytes / 1e9)},
            "roofline": {"bound": "hbm", "kernel": "FASTA step (k_fa_summary + k_fa_write + k_stream_ws RAW)", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_step": algo_bytes, "text_bytes_per_step": text_bytes,
                         "stream_kernel_ms_per_launch": k_ms},
            "cpu_baseline": cpu, "e2e": e2e, "parity": parity, "gpu_launches": int(prof.kernel_launches), "clocks": clocks}
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline_fasta(args, sk, spec, shuf_id, perm, buf, off):
    """reference `dist` (no -A) on a bounded set of genomes with all host cores; its sketches against ours."""
    import metakssd_b200 as M
    import oracle as O
    ref = _ref_bin()
    threads = os.cpu_count() or 1
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    workdir = tempfile.mkdtemp(prefix="mkssd_cpu_", dir=base)
    try:
        n = int(min(off.size - 1, max(threads * 3, 24)))
        paths = []
        for i in range(n):
            pth = os.path.join(workdir, "g%d.fasta" % i)
            buf[int(off[i]):int(off[i + 1])].cpu().numpy().tofile(pth)
            paths.append(pth)
        shuf_path = os.path.join(workdir, "x.shuf")
        M.write_shuf(shuf_path, shuf_id, args.k, args.subk, args.L, perm)
        best = None
        for _ in range(2):
            out = os.path.join(workdir, "gsk")
            shutil.rmtree(out, ignore_errors=True)
            t0 = time.perf_counter()
            subprocess.run([ref, "dist", "-L", shuf_path, "-p", str(threads), "-o", out] + paths, check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            dt = time.perf_counter() - t0
            best = dt if best is None or dt < best else best
        sd = O.read_sketch_dir(os.path.join(workdir, "gsk"))
        ours = sk.fasta_co_files(paths)
        ok = True
        for i, nme in enumerate(sd.names):
            j = paths.index(nme)
            for c in range(sd.comp_num):
                ok = ok and np.array_equal(sd.combco[c][int(sd.index[c][i]):int(sd.index[c][i + 1])], ours[j].codes[c])
        return {"value": n * args.genome_len / 1e9 / best, "unit": "Gbp/s", "cores": threads, "kind": "reference",
                "sample": "%d genomes of the workload, `metakssd dist -L %s.shuf -p %d`, page cache warm, best of 2" % (n, args.geom, threads),
                "dist_s": best, "parity": {"genome_sketches_vs_reference": {"ok": bool(ok), "genomes": n}}}
    finally:
        shutil.rmtree(workdir, ignore_errors=True)


# =====================================================================================================
def _write_sample(spec, d_text, n_reads, workdir):
    """First n_reads records of the device-resident FASTQ -> file (page cache / tmpfs)."""
    nb = spec.fastq_bytes(0, n_reads)
    path = os.path.join(workdir, "reads.fq")
    with open(path, "wb") as f:
        CH = 256 << 20
        for o in range(0, nb, CH):
            f.write(d_text[o:min(nb, o + CH)].cpu().numpy().tobytes())
    return path, nb


def _write_markerdb(mdb, shuf_id, workdir, info):
    import metakssd_b200 as M
    path = os.path.join(workdir, "markerdb")

    class _S:  # one "sketch" per species for write_sketch_dir
        def __init__(self, codes):
            self.codes = codes
            self.counts = None

        @property
        def n_total(self):
            return int(sum(c.size for c in self.codes))

    sk_list = []
    S = len(mdb.names)
    for s in range(S):
        sk_list.append(_S([mdb.comp[c][0][int(mdb.comp[c][1][s]):int(mdb.comp[c][1][s + 1])] for c in range(len(mdb.comp))]))
    M.write_sketch_dir(path, shuf_id, info, mdb.names, sk_list, koc=False)
    return path


def _time_cli(binary, shuf_path, fq_path, mdb_path, out, threads):
    """wall seconds of `<binary> dist -L .. -A` and of `<binary> composite`, and composite's stdout"""
    shutil.rmtree(out, ignore_errors=True)
    t0 = time.perf_counter()
    subprocess.run([binary, "dist", "-L", shuf_path, "-A", "-p", str(threads), "-o", out, fq_path], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    t1 = time.perf_counter()
    r = subprocess.run([binary, "composite", "-r", mdb_path, "-q", out, "-p", str(threads)], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1, r.stdout.decode()


def cpu_baseline(args, sk, spec, shuf_id, perm, mdb, d_text, composite):
    """(cpu_baseline, cli_e2e, parity) — the reference binary and our C host program on the SAME sample file and
    MarkerDB directory, same command lines; their outputs compared."""
    import metakssd_b200 as M
    threads = os.cpu_count() or 1
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    workdir = tempfile.mkdtemp(prefix="mkssd_cpu_", dir=base)
    parity = {}
    try:
        n = min(args.cpu_reads, args.reads)
        fq, nb = _write_sample(spec, d_text, n, workdir)
        ref = _ref_bin()
        cli_bin = _cli_bin()
        shuf_path = os.path.join(workdir, "%s.shuf" % args.geom)
        M.write_shuf(shuf_path, shuf_id, args.k, args.subk, args.L, perm)
        mdb_path = _write_markerdb(mdb, shuf_id, workdir, sk.info)
        cpu = cli = None
        ours_tsv = None
        if cli_bin:
            best = None
            for _ in range(2):
                td, tc, ours_tsv = _time_cli(cli_bin, shuf_path, fq, mdb_path, os.path.join(workdir, "qry_gpu"), threads)
                if best is None or td + tc < best[0] + best[1]:
                    best = (td, tc)
            cli = {"value": n * READ_LEN / 1e9 / (best[0] + best[1]), "unit": "Gbp/s",
                   "what": "wall clock of `host/metakssd-b200 dist -L %s.shuf -A` + `composite` (process start, CUDA context, "
                           "streaming ingest of the page-cached file, sketch directory on disk), best of 2" % args.geom,
                   "sample": "first %d records (%.2f Gbp, %.0f MB FASTQ) of the bench workload" % (n, n * READ_LEN / 1e9, nb / 1e6),
                   "dist_s": best[0], "composite_s": best[1], "text_GB_s": nb / 1e9 / best[0]}
        if ref:
            best = None
            for _ in range(2):
                td, tc, ref_tsv = _time_cli(ref, shuf_path, fq, mdb_path, os.path.join(workdir, "qry_ref"), threads)
                if best is None or td + tc < best[0] + best[1]:
                    best = (td, tc)
            cpu = {"value": n * READ_LEN / 1e9 / (best[0] + best[1]), "unit": "Gbp/s", "cores": threads,
                   "kind": "reference", "sample": "first %d records (%.2f Gbp, %.0f MB FASTQ) of the bench workload, "
                   "`metakssd dist -L %s.shuf -A -p %d` + `composite -p %d`, page cache warm, best of 2"
                   % (n, n * READ_LEN / 1e9, nb / 1e6, args.geom, threads, threads),
                   "dist_s": best[0], "composite_s": best[1]}
            # the reference's composite output on the sample against ours (API and C host program)
            api_tsv = composite(sk.fastq_koc_device(d_text, nb), True)
            rows_ref = _tsv_rows(ref_tsv)
            parity["composite_vs_reference"] = {"ok": rows_ref == _tsv_rows(api_tsv) and len(rows_ref) > 0, "rows": len(rows_ref)}
            if ours_tsv is not None:
                parity["cli_vs_reference"] = {"ok": rows_ref == _tsv_rows(ours_tsv), "rows": len(rows_ref)}
            # the reference reads OUR sketch directory (drop-in file formats)
            if cli_bin:
                r = subprocess.run([ref, "composite", "-r", mdb_path, "-q", os.path.join(workdir, "qry_gpu"), "-p", str(threads)],
                                   check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
                parity["reference_reads_our_sketch_dir"] = {"ok": _tsv_rows(r.stdout.decode()) == rows_ref}
            if not args.no_parity:
                parity.update(_markerdb_subset_parity(args, sk, spec, shuf_path, workdir, threads))
        else:
            import oracle as O
            n2 = min(n, 300_000)
            text = d_text[:spec.fastq_bytes(0, n2)].cpu().numpy()
            p = O.params(args.k, args.subk, args.L)
            t0 = time.perf_counter()
            O.fastq_koc(p, perm, text)
            dt = time.perf_counter() - t0
            cpu = {"value": n2 * READ_LEN / 1e9 / dt, "unit": "Gbp/s", "cores": 1, "kind": "port",
                   "sample": "first %d records, single-threaded C restatement (oracle/kssd_oracle.c)" % n2}
        return cpu, cli, parity
    finally:
        shutil.rmtree(workdir, ignore_errors=True)


def _markerdb_subset_parity(args, sk, spec, shuf_path, workdir, threads, n_sub=40):
    """MarkerDB of the first n_sub species built by the reference's own pipeline (dist -> set -g -> set -q -> set -i,
    command_set.c:831,427,322) against the one the device pipeline (FASTA sketches, mk_set_group / mk_set_uniq_union /
    mk_set_operate) builds from the same genomes: same species order, same code order, same offsets."""
    import torch
    import oracle as O
    from metakssd_b200 import workload as W
    n_sub = min(n_sub, args.species)
    gdir = os.path.join(workdir, "genomes")
    os.makedirs(gdir, exist_ok=True)
    per_file = spec.fasta_bytes(0) + 16
    buf = torch.empty(n_sub * per_file + 256, dtype=torch.uint8, device="cuda:%d" % sk.info.device)
    off = sk.synth_fasta_device(spec.P, 0, n_sub, buf, buf.numel())
    paths, groups = [], []
    for s in range(n_sub):
        pth = os.path.join(gdir, "sp%d.fasta" % s)
        buf[int(off[s]):int(off[s + 1])].cpu().numpy().tofile(pth)
        paths.append(pth)
        groups.append("%d\tsp%d" % (s + 1, s))
    os.makedirs(os.path.join(workdir, "refmdb"), exist_ok=True)
    mdb_ref = O.ref_build_markerdb(shuf_path, paths, groups, os.path.join(workdir, "refmdb"), p=threads)
    md = O.read_sketch_dir(mdb_ref)
    # the same genome sketches in the order the reference's `dist` listed them, through the device pipeline
    gsk = O.read_sketch_dir(os.path.join(workdir, "refmdb", "gsk"))
    order = [int(os.path.basename(nme)[2:].split(".")[0]) for nme in gsk.names]
    sks = sk.fasta_co_device(buf, off)
    ours = W.markerdb_pipeline(sk, [sks[g] for g in order], [g + 1 for g in order], ["sp%d" % g for g in order])
    ok = md.infile_num == n_sub and ours.names == md.names
    for c in range(md.comp_num):            # byte for byte: species order, code order inside a species, offsets
        ok = ok and np.array_equal(ours.comp[c][0], md.combco[c]) and np.array_equal(ours.comp[c][1], md.index[c])
    shutil.rmtree(gdir, ignore_errors=True)
    return {"markerdb_subset_vs_reference_set_pipeline": {"ok": bool(ok), "species": n_sub,
                                                          "codes": int(sum(c.size for c in md.combco))}}


# =====================================================================================================
def run_reference(args):
    """The unmodified reference binary (all host cores).  Inputs come from the CPU generator under oracle/ and the
    MarkerDB from the reference's own dist -> set -g -> set -q -> set -i pipeline: nothing of the product library
    is loaded in this process.  Each step runs the whole workload of the `ours` arm when the K + W steps fit the
    time budget (MK_REF_BUDGET_S, default 600 s), else the largest prefix that does (stated in `config`)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import oracle as O
    O.build()
    ref = _ref_bin()
    threads = os.cpu_count() or 1
    K, SUBK, L = args.k, args.subk, args.L
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    workdir = tempfile.mkdtemp(prefix="mkssd_ref_", dir=base)
    budget = float(os.environ.get("MK_REF_BUDGET_S", "600"))
    try:
        shuf_id, perm = O.make_shuf(SEED ^ 1, K, SUBK, L)
        S = O.synth(SEED ^ 2, args.species, args.genome_len, READ_LEN)
        shuf_path = os.path.join(workdir, "%s.shuf" % args.geom)
        O.write_shuf_file(shuf_path, shuf_id, K, SUBK, L, perm)
        t_setup = time.perf_counter()
        if args.kind == "fasta":
            return _run_reference_fasta(args, O, S, ref, shuf_path, workdir, threads, world, budget)
        per_rank = args.reads if args.scaling == "weak" else args.reads // world
        # MarkerDB: the reference's own pipeline on all species when the binary is here
        if ref:
            gdir = os.path.join(workdir, "genomes")
            os.makedirs(gdir)
            paths, groups = [], []
            for s in range(args.species):
                pth = os.path.join(gdir, "sp%d.fasta" % s)
                S.fasta(s).tofile(pth)
                paths.append(pth)
                groups.append("%d\tsp%d" % (s + 1, s))
            os.makedirs(os.path.join(workdir, "mdb"))
            mdb_path = O.ref_build_markerdb(shuf_path, paths, groups, os.path.join(workdir, "mdb"), p=threads)
            shutil.rmtree(gdir, ignore_errors=True)
            shutil.rmtree(os.path.join(workdir, "mdb", "gsk"), ignore_errors=True)
        # probe the reference's speed on 1 M records to size the per-step sample
        fq = os.path.join(workdir, "reads.fq")

        def write_reads(n):
            with open(fq, "wb") as f:
                for a in range(0, n, 2_000_000):
                    f.write(S.fastq(a, min(n, a + 2_000_000)).tobytes())
        if ref:
            def one():
                td, tc, _ = _time_cli(ref, shuf_path, fq, mdb_path, os.path.join(workdir, "qry"), threads)
                return td + tc
            kind, cores = "reference", threads
        else:
            p = O.params(K, SUBK, L)

            def one():
                text = np.fromfile(fq, dtype=np.uint8)
                t0 = time.perf_counter()
                O.fastq_koc(p, perm, text)
                return time.perf_counter() - t0
            kind, cores = "port", 1
        n_probe = min(1_000_000, per_rank)
        write_reads(n_probe)
        t_probe = one()
        t_gen = 1.7e-6 * 8 / max(1, min(threads, 8))          # s per record of the CPU generator (measured ~1.6 s / M on 8 cores)
        # time per record from the probe (the fixed table clearing / scan amortises over larger inputs: conservative)
        per_rec = t_probe / n_probe
        spent = time.perf_counter() - t_setup
        room = max(30.0, budget - spent)
        n = int(min(per_rank, room / ((args.steps + args.warmup) * per_rec + t_gen)))
        n = max(n_probe, n)
        if n != n_probe:
            write_reads(n)
        for _ in range(args.warmup):
            one()
        times = [one() for _ in range(args.steps)]
        sec = sum(times) / len(times)
        val = n * READ_LEN / 1e9 / sec
        sample = ("each step: %s %d records (%.2f Gbp) of the workload through `metakssd dist -L %s.shuf -A -p %d` + "
                  "`composite -p %d` (reference binary built from /root/reference, inputs from the CPU generator under oracle/, "
                  "MarkerDB from the reference's dist/set -g/-q/-i pipeline, page cache warm)"
                  % ("all" if n == per_rank else "the first", n, n * READ_LEN / 1e9, args.geom, cores, cores))
        line = {"impl": "reference", "metric": args.metric, "value": val, "unit": "Gbp/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": {"workload": workload_name(args, world), "config": args.config, "k": K, "subk": SUBK, "L": L,
                           "read_len": READ_LEN, "species": args.species, "genome_len": args.genome_len, "sample_reads": n,
                           "reads_per_gpu": per_rank, "whole_workload_per_step": n == per_rank},
                "cpu_baseline": {"value": val, "unit": "Gbp/s", "cores": cores, "kind": kind, "sample": sample},
                "e2e": {"value": val, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
    finally:
        shutil.rmtree(workdir, ignore_errors=True)


def _run_reference_fasta(args, O, S, ref, shuf_path, workdir, threads, world, budget):
    G = args.species
    g1 = G // world
    if not ref:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/metakssd was not built"}))
        return
    n = int(min(g1, max(threads * 3, 24)))
    paths = []
    for i in range(n):
        pth = os.path.join(workdir, "g%d.fasta" % i)
        S.fasta(i).tofile(pth)
        paths.append(pth)

    def one():
        out = os.path.join(workdir, "gsk")
        shutil.rmtree(out, ignore_errors=True)
        t0 = time.perf_counter()
        subprocess.run([ref, "dist", "-L", shuf_path, "-p", str(threads), "-o", out] + paths, check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        return time.perf_counter() - t0
    for _ in range(args.warmup):
        one()
    times = [one() for _ in range(args.steps)]
    sec = sum(times) / len(times)
    val = n * args.genome_len / 1e9 / sec
    sample = "each step: %d genomes of the workload through `metakssd dist -L %s.shuf -p %d`" % (n, args.geom, threads)
    print(json.dumps({"impl": "reference", "metric": args.metric, "value": val, "unit": "Gbp/s", "n_gpus": world,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
                      "scaling": args.scaling, "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                      "config": {"workload": workload_name(args, world), "config": args.config, "sample_genomes": n},
                      "cpu_baseline": {"value": val, "unit": "Gbp/s", "cores": threads, "kind": "reference", "sample": sample},
                      "e2e": {"value": val, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
