"""TEST INFRASTRUCTURE ONLY — ctypes binding of oracle/liboracle.so (the CPU restatement of the
MetaKSSD hot path, see kssd_oracle.c) plus helpers to drive the reference binary
(oracle/_ref/metakssd) and to read/write the reference's on-disk sketch directory.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product package (metakssd_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import struct
import subprocess
import tempfile
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_BIN = os.path.join(HERE, "_ref", "metakssd")


def build(force: bool = False) -> None:
    """Compile liboracle.so (and oracle/_ref/metakssd when /root/reference is present)."""
    if force and os.path.exists(LIB_PATH):
        os.remove(LIB_PATH)
    subprocess.run(["make", "-C", HERE, "liboracle.so"], check=True, capture_output=True)      # (no-op when up to date)
    if os.path.isdir("/root/reference") and (force or not os.path.exists(REF_BIN)):
        subprocess.run(["make", "-C", HERE, "ref"], check=True, capture_output=True)


class KoParams(C.Structure):
    _fields_ = [
        ("k", C.c_int), ("subk", C.c_int), ("drlevel", C.c_int), ("outctx", C.c_int),
        ("TL", C.c_int), ("crvs_shift", C.c_int),
        ("tupmask", C.c_uint64), ("domask", C.c_uint64), ("undomask", C.c_uint64),
        ("lowmask", C.c_uint64), ("dim_end", C.c_int), ("hashsize", C.c_uint32),
        ("hashlimit", C.c_uint32), ("component_num", C.c_int), ("comp_code_bits", C.c_int),
    ]


class KoSketch(C.Structure):
    _fields_ = [
        ("n", C.c_size_t), ("codes", C.POINTER(C.c_uint64)), ("counts", C.POINTER(C.c_uint16)),
        ("slots", C.POINTER(C.c_uint32)), ("status", C.c_int),
    ]


class MksParams(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64), ("n_species", C.c_uint32), ("genome_len", C.c_uint32),
        ("read_len", C.c_uint32), ("genus_size", C.c_uint32), ("shared_len", C.c_uint32),
        ("sub_thresh16", C.c_uint32), ("n_thresh16", C.c_uint32), ("n_present", C.c_uint32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.ko_params_init.argtypes = [C.POINTER(KoParams), C.c_int, C.c_int, C.c_int]
        L.ko_params_init.restype = C.c_int
        for fn in (L.ko_fastq_koc, L.ko_fasta_co):
            fn.argtypes = [C.POINTER(KoParams), C.c_void_p, C.c_void_p, C.c_size_t]
            fn.restype = C.POINTER(KoSketch)
        L.ko_fastq_co.argtypes = [C.POINTER(KoParams), C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int]
        L.ko_fastq_co.restype = C.POINTER(KoSketch)
        L.ko_fasta_co_uniq.argtypes = [C.POINTER(KoParams), C.c_void_p, C.c_void_p, C.c_size_t]
        L.ko_fasta_co_uniq.restype = C.POINTER(KoSketch)
        L.ko_organize_taxa.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ko_organize_taxa.restype = C.c_int
        L.ko_set_group.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ko_set_group.restype = None
        L.ko_set_uniq_union.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.ko_set_uniq_union.restype = C.c_size_t
        L.ko_set_operate.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.ko_set_operate.restype = None
        L.ko_sketch_free.argtypes = [C.POINTER(KoSketch)]
        L.ko_composite_component.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                             C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p]
        L.ko_composite_component.restype = C.c_int
        L.ko_composite_report.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_char_p, C.c_size_t]
        L.ko_composite_report.restype = C.c_size_t
        L.ko_synth_params.argtypes = [C.POINTER(MksParams), C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32,
                                      C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.POINTER(C.c_uint32))]
        L.ko_synth_params.restype = C.c_int
        L.ko_fastq_bytes.argtypes = [C.POINTER(MksParams), C.c_uint64, C.c_uint64]
        L.ko_fastq_bytes.restype = C.c_uint64
        L.ko_write_fastq.argtypes = [C.POINTER(MksParams), C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
        L.ko_write_fastq.restype = C.c_size_t
        L.ko_fasta_bytes.argtypes = [C.POINTER(MksParams), C.c_uint32]
        L.ko_fasta_bytes.restype = C.c_size_t
        L.ko_write_fasta.argtypes = [C.POINTER(MksParams), C.c_uint32, C.c_void_p]
        L.ko_write_fasta.restype = C.c_size_t
        L.ko_make_shuf_perm.argtypes = [C.c_uint64, C.c_int, C.c_void_p]
        L.ko_shuf_id.argtypes = [C.c_uint64]
        L.ko_shuf_id.restype = C.c_int32
        L.ko_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


# --------------------------------------------------------------------------- parameters / shuf
def params(k: int, subk: int, drlevel: int) -> KoParams:
    p = KoParams()
    if lib().ko_params_init(C.byref(p), k, subk, drlevel) != 0:
        raise ValueError("primer index out of range for k=%d L=%d" % (k, drlevel))
    return p


def make_shuf(seed: int, k: int, subk: int, drlevel: int):
    """Deterministic .shuf content: (shuf_id, int32 permutation of 16^subk)."""
    perm = np.empty(1 << (4 * subk), dtype=np.int32)
    lib().ko_make_shuf_perm(seed, subk, perm.ctypes.data)
    return int(lib().ko_shuf_id(seed)), perm


def write_shuf_file(path: str, shuf_id: int, k: int, subk: int, drlevel: int, perm: np.ndarray) -> None:
    """File format of the reference's .shuf (command_shuffle.c:164-211)."""
    with open(path, "wb") as f:
        f.write(struct.pack("<iiii", shuf_id, k, subk, drlevel))
        f.write(np.ascontiguousarray(perm, dtype=np.int32).tobytes())


# --------------------------------------------------------------------------- synthetic inputs
@dataclass
class Synth:
    P: MksParams
    cdf32: np.ndarray
    species: np.ndarray

    def fastq(self, r0: int, r1: int) -> np.ndarray:
        n = int(lib().ko_fastq_bytes(C.byref(self.P), r0, r1))
        buf = np.empty(n, dtype=np.uint8)
        w = lib().ko_write_fastq(C.byref(self.P), self.cdf32.ctypes.data, self.species.ctypes.data, r0, r1,
                                 buf.ctypes.data)
        assert w == n, (w, n)
        return buf

    def fasta(self, s: int) -> np.ndarray:
        n = int(lib().ko_fasta_bytes(C.byref(self.P), s))
        buf = np.empty(n, dtype=np.uint8)
        w = lib().ko_write_fasta(C.byref(self.P), s, buf.ctypes.data)
        assert w == n
        return buf


def synth(seed: int, n_species: int, genome_len: int, read_len: int = 150) -> Synth:
    P = MksParams()
    cdf = C.POINTER(C.c_uint32)()
    spc = C.POINTER(C.c_uint32)()
    if lib().ko_synth_params(C.byref(P), seed, n_species, genome_len, read_len, C.byref(cdf), C.byref(spc)) != 0:
        raise MemoryError
    n = P.n_present
    cdf_np = np.ctypeslib.as_array(cdf, shape=(n,)).copy()
    spc_np = np.ctypeslib.as_array(spc, shape=(n,)).copy()
    lib().ko_free(cdf)
    lib().ko_free(spc)
    return Synth(P, cdf_np, spc_np)


# --------------------------------------------------------------------------- sketching
@dataclass
class Sketch:
    codes: np.ndarray            # uint64, full codes in on-disk order
    counts: np.ndarray | None    # uint16 (None for FASTA)
    slots: np.ndarray
    status: int

    def components(self, p: KoParams):
        """[(filecodes uint32, counts uint16|None)] per component, as the reference writes them."""
        comp = (self.codes % np.uint64(p.component_num)).astype(np.int64)
        fc = (self.codes >> np.uint64(p.comp_code_bits)).astype(np.uint32)
        out = []
        for c in range(p.component_num):
            m = comp == c
            out.append((fc[m], None if self.counts is None else self.counts[m]))
        return out


def _as_bytes(text) -> np.ndarray:
    if isinstance(text, (bytes, bytearray)):
        return np.frombuffer(bytes(text), dtype=np.uint8)
    return np.ascontiguousarray(text, dtype=np.uint8)


def _take(sk_ptr, with_counts: bool) -> Sketch:
    sk = sk_ptr.contents
    n = sk.n
    codes = np.ctypeslib.as_array(sk.codes, shape=(n,)).copy() if n else np.empty(0, np.uint64)
    slots = np.ctypeslib.as_array(sk.slots, shape=(n,)).copy() if n else np.empty(0, np.uint32)
    counts = None
    if with_counts:
        counts = np.ctypeslib.as_array(sk.counts, shape=(n,)).copy() if n else np.empty(0, np.uint16)
    st = sk.status
    lib().ko_sketch_free(sk_ptr)
    return Sketch(codes, counts, slots, st)


def fastq_koc(p: KoParams, perm: np.ndarray, text) -> Sketch:
    t = _as_bytes(text)
    return _take(lib().ko_fastq_koc(C.byref(p), perm.ctypes.data, t.ctypes.data, t.size), True)


def fasta_co(p: KoParams, perm: np.ndarray, text) -> Sketch:
    t = _as_bytes(text)
    return _take(lib().ko_fasta_co(C.byref(p), perm.ctypes.data, t.ctypes.data, t.size), False)


def fastq_co(p: KoParams, perm: np.ndarray, text, quality: int = 0, min_occurrence: int = 1) -> Sketch:
    """`dist` on FASTQ without -A: fastq2co(.., Q, M) + write_fqco2file() (iseq2comem.c:323-419, 596-621)."""
    t = _as_bytes(text)
    return _take(lib().ko_fastq_co(C.byref(p), perm.ctypes.data, t.ctypes.data, t.size, quality, min_occurrence), False)


def fasta_co_uniq(p: KoParams, perm: np.ndarray, text) -> Sketch:
    """`dist -u` on FASTA: uniq_fasta2co() (iseq2comem.c:729-828), codes occurring once in the file."""
    t = _as_bytes(text)
    return _take(lib().ko_fasta_co_uniq(C.byref(p), perm.ctypes.data, t.ctypes.data, t.size), False)


# --------------------------------------------------------------------------- set -g / -q / -i
def organize_taxa(taxids):
    """organize_taxf(): (taxon_of_genome int32[], taxid_of_taxon list) in the reference's output order."""
    tid = np.ascontiguousarray(taxids, dtype=np.int32)
    taxon_of = np.empty(tid.size, dtype=np.int32)
    ids = np.empty(tid.size, dtype=np.int32)
    n = lib().ko_organize_taxa(tid.ctypes.data, tid.size, taxon_of.ctypes.data, ids.ctypes.data)
    return taxon_of, ids[:n].tolist()


def set_group(codes, index, taxon_of, n_taxa):
    codes = np.ascontiguousarray(codes, dtype=np.uint32)
    index = np.ascontiguousarray(index, dtype=np.uint64)
    taxon_of = np.ascontiguousarray(taxon_of, dtype=np.int32)
    out = np.empty(max(1, codes.size), dtype=np.uint32)
    oi = np.zeros(n_taxa + 1, dtype=np.uint64)
    lib().ko_set_group(codes.ctypes.data, index.ctypes.data, index.size - 1, taxon_of.ctypes.data, n_taxa, out.ctypes.data, oi.ctypes.data)
    return out[:int(oi[-1])].copy(), oi


def set_uniq_union(codes):
    codes = np.ascontiguousarray(codes, dtype=np.uint32)
    out = np.empty(max(1, codes.size), dtype=np.uint32)
    n = lib().ko_set_uniq_union(codes.ctypes.data, codes.size, out.ctypes.data)
    return out[:n].copy()


def set_operate(pan, codes, index, intersect=True):
    pan = np.ascontiguousarray(pan, dtype=np.uint32)
    codes = np.ascontiguousarray(codes, dtype=np.uint32)
    index = np.ascontiguousarray(index, dtype=np.uint64)
    out = np.empty(max(1, codes.size), dtype=np.uint32)
    oi = np.zeros(index.size, dtype=np.uint64)
    lib().ko_set_operate(pan.ctypes.data, pan.size, codes.ctypes.data, index.ctypes.data, index.size - 1, 1 if intersect else 0,
                         out.ctypes.data, oi.ctypes.data)
    return out[:int(oi[-1])].copy(), oi


# --------------------------------------------------------------------------- composite
def composite(ref_comp, ref_names, qry_comp, qry_name: str) -> str:
    """ref_comp: list over components of (codes uint32, index uint64[S+1]);
    qry_comp: list over components of (codes uint32, counts uint16).  Returns the TSV text."""
    S = len(ref_names)
    total = sum(int(c[0].size) for c in ref_comp)
    store = [np.zeros(max(1, total), dtype=np.int32) for _ in range(1)]
    # per-species hit buffers sized by the species' total code count over components
    sizes = np.zeros(S, dtype=np.int64)
    for codes, index in ref_comp:
        sizes += (index[1:] - index[:-1]).astype(np.int64)
    bufs = [np.zeros(max(1, int(s)), dtype=np.int32) for s in sizes]
    ptrs = (C.c_void_p * S)(*[b.ctypes.data for b in bufs])
    nhits = np.zeros(S, dtype=np.int32)
    for (rc, ri), (qc, qa) in zip(ref_comp, qry_comp):
        rc = np.ascontiguousarray(rc, dtype=np.uint32)
        ri = np.ascontiguousarray(ri, dtype=np.uint64)
        qc = np.ascontiguousarray(qc, dtype=np.uint32)
        qa = np.ascontiguousarray(qa, dtype=np.uint16)
        rcode = lib().ko_composite_component(rc.ctypes.data, ri.ctypes.data, S, qc.ctypes.data, qa.ctypes.data,
                                             0, qc.size, C.cast(ptrs, C.c_void_p), nhits.ctypes.data)
        if rcode != 0:
            raise ValueError("empty query component (the reference divides by zero here)")
    names = (C.c_char_p * S)(*[n.encode() for n in ref_names])
    cap = 512 * (S + 1)
    out = C.create_string_buffer(cap)
    w = lib().ko_composite_report(qry_name.encode(), names, S, C.cast(ptrs, C.c_void_p), nhits.ctypes.data, out, cap)
    del store
    return out.raw[:w].decode()


# --------------------------------------------------------------------------- dist -r (shared k-mer counts, distance table)
def shared_counts(ref_comp, qry_comp, qry_ctx_ct=None) -> np.ndarray:
    """ref_comp / qry_comp: lists over components of (codes uint32, index uint64[n+1]).  Returns uint32[n_qry, n_ref]."""
    n_ref, n_qry = ref_comp[0][1].size - 1, qry_comp[0][1].size - 1
    counts = np.zeros((n_qry, n_ref), dtype=np.uint32)
    qc = None if qry_ctx_ct is None else np.ascontiguousarray(qry_ctx_ct, dtype=np.uint32)
    for (rc, ri), (qcodes, qi) in zip(ref_comp, qry_comp):
        rc = np.ascontiguousarray(rc, dtype=np.uint32); ri = np.ascontiguousarray(ri, dtype=np.uint64)
        qcodes = np.ascontiguousarray(qcodes, dtype=np.uint32); qi = np.ascontiguousarray(qi, dtype=np.uint64)
        f = lib().ko_shared_counts
        f.restype = None
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        f(rc.ctypes.data, ri.ctypes.data, n_ref, qcodes.ctypes.data, qi.ctypes.data, n_qry,
          None if qc is None else qc.ctypes.data, counts.ctypes.data)
    return counts


def distance_out(counts, ref_ctx_ct, qry_ctx_ct, ref_names, qry_names, kmerlen, dim_rd_len, metric=0, outfields=2,
                 correction=0, n_max=0, max_dist=1.0) -> str:
    counts = np.ascontiguousarray(counts, dtype=np.uint32)
    n_qry, n_ref = counts.shape
    rc = np.ascontiguousarray(ref_ctx_ct, dtype=np.uint32); qc = np.ascontiguousarray(qry_ctx_ct, dtype=np.uint32)
    rn = (C.c_char_p * n_ref)(*[n.encode() for n in ref_names]); qn = (C.c_char_p * n_qry)(*[n.encode() for n in qry_names])
    f = lib().ko_distance_out
    f.restype = C.c_size_t
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                  C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    args = (counts.ctypes.data, n_ref, n_qry, rc.ctypes.data, qc.ctypes.data, C.cast(rn, C.c_void_p), C.cast(qn, C.c_void_p), metric,
            outfields, correction, n_max, float(max_dist), kmerlen, dim_rd_len)
    need = f(*args, None, 0)
    out = C.create_string_buffer(need + 1)
    f(*args, C.cast(out, C.c_void_p), need)
    return out.raw[:need].decode()


def ref_dist_search(refdir: str, qrydir: str, outdir: str, extra=(), p: int = 1) -> str:
    """`metakssd dist -r <refdir> -o <outdir> [extra] <qrydir>` with the reference binary; returns distance.out.
    The first use of a co directory as -r makes the reference write its inverted index into it (mco.index.N is
    2^32 x 8 bytes = 32 GiB per component, co2mco.c:17-67): minutes of time and 32 GiB of disk per component."""
    import shutil
    shutil.rmtree(outdir, ignore_errors=True)
    run_ref(["dist", "-r", refdir, "-o", outdir, "-p", str(p)] + list(extra) + [qrydir])
    return open(os.path.join(outdir, "distance.out")).read()


# --------------------------------------------------------------------------- sketch directories
CO_DSTAT = struct.Struct("<I?3xiiiiQ")  # co_dstat_t, global_basic.h:116-126 (32 bytes)


@dataclass
class SketchDir:
    shuf_id: int
    koc: bool
    kmerlen: int
    dim_rd_len: int
    comp_num: int
    infile_num: int
    all_ctx_ct: int
    ctx_ct: np.ndarray
    names: list
    combco: list          # per component uint32[]
    index: list           # per component uint64[infile_num+1]
    abund: list | None    # per component uint16[] (koc only)


def read_sketch_dir(path: str) -> SketchDir:
    raw = open(os.path.join(path, "cofiles.stat"), "rb").read()
    shuf_id, koc, kmerlen, dim_rd_len, comp_num, infile_num, all_ctx = CO_DSTAT.unpack_from(raw, 0)
    off = CO_DSTAT.size
    ctx_ct = np.frombuffer(raw, dtype=np.uint32, count=infile_num, offset=off).copy()
    off += 4 * infile_num
    names = []
    for i in range(infile_num):
        rec = raw[off + 256 * i: off + 256 * (i + 1)]
        names.append(rec.split(b"\0", 1)[0].decode())
    combco, index, abund = [], [], ([] if koc else None)
    for c in range(comp_num):
        combco.append(np.fromfile(os.path.join(path, "combco.%d" % c), dtype=np.uint32))
        index.append(np.fromfile(os.path.join(path, "combco.index.%d" % c), dtype=np.uint64))
        if koc:
            abund.append(np.fromfile(os.path.join(path, "combco.%d.a" % c), dtype=np.uint16))
    return SketchDir(shuf_id, bool(koc), kmerlen, dim_rd_len, comp_num, infile_num, all_ctx, ctx_ct, names,
                     combco, index, abund)


def write_sketch_dir(path: str, sd: SketchDir) -> None:
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "cofiles.stat"), "wb") as f:
        f.write(CO_DSTAT.pack(sd.shuf_id, sd.koc, sd.kmerlen, sd.dim_rd_len, sd.comp_num, sd.infile_num,
                              sd.all_ctx_ct))
        f.write(np.ascontiguousarray(sd.ctx_ct, dtype=np.uint32).tobytes())
        for n in sd.names:
            b = n.encode()[:255]
            f.write(b + b"\0" * (256 - len(b)))
    for c in range(sd.comp_num):
        np.ascontiguousarray(sd.combco[c], dtype=np.uint32).tofile(os.path.join(path, "combco.%d" % c))
        np.ascontiguousarray(sd.index[c], dtype=np.uint64).tofile(os.path.join(path, "combco.index.%d" % c))
        if sd.koc:
            np.ascontiguousarray(sd.abund[c], dtype=np.uint16).tofile(os.path.join(path, "combco.%d.a" % c))


# --------------------------------------------------------------------------- reference binary
def have_ref() -> bool:
    return os.path.exists(REF_BIN)


def run_ref(args, cwd=None, threads_env=None) -> subprocess.CompletedProcess:
    env = dict(os.environ)
    if threads_env:
        env["OMP_NUM_THREADS"] = str(threads_env)
    return subprocess.run([REF_BIN] + list(args), cwd=cwd, env=env, capture_output=True, text=True, check=True)


def ref_dist(shuf_path: str, inputs, outdir: str, abundance: bool, p: int = 1, extra=()) -> SketchDir:
    """`metakssd dist -L <shuf> [-A] [extra flags] -p <p> -o <outdir> <inputs...>` with the reference binary."""
    args = ["dist", "-L", shuf_path, "-p", str(p), "-o", outdir] + list(extra)
    if abundance:
        args.append("-A")
    run_ref(args + list(inputs))
    return read_sketch_dir(outdir)


def ref_composite(refdir: str, qrydir: str, p: int = 1) -> str:
    out = run_ref(["composite", "-r", refdir, "-q", qrydir, "-p", str(p)]).stdout
    return out


def ref_build_markerdb(shuf_path: str, genome_paths, group_lines, workdir: str, p: int = 1) -> str:
    """The reference's own MarkerDB pipeline (README.md:80-104): dist -> set -g -> set -q -> set -i.
    group_lines[i] = "<taxid>\t<name>" for genome i **in the order the sketch lists them**."""
    gsk = os.path.join(workdir, "gsk")
    sd = ref_dist(shuf_path, genome_paths, gsk, abundance=False, p=p)
    by_path = {os.path.abspath(gp): gl for gp, gl in zip(genome_paths, group_lines)}
    grp = os.path.join(workdir, "group_name.txt")
    with open(grp, "w") as f:
        for n in sd.names:
            f.write(by_path[os.path.abspath(n)] + "\n")
    pan = os.path.join(workdir, "pan")
    uni = os.path.join(workdir, "union_sp")
    mdb = os.path.join(workdir, "markerdb")
    run_ref(["set", "-g", grp, "-o", pan, gsk])
    run_ref(["set", "-q", "-o", uni, pan])
    run_ref(["set", "-i", uni, "-o", mdb, pan])
    return mdb
