"""ctypes binding of libmkssd_b200.so (include/mkssd_b200.h) and a thin host-side mirror of the
reference's interface for the hot path:

  reference (C, yhg926/MetaKSSD)                          here
  ------------------------------------------------------  -----------------------------------------
  read_dim_shuffle_file()      command_shuffle.c:215      read_shuf()
  seq2co_global_var_initial()  iseq2comem.c:54            Sketcher(perm, k, subk, drlevel)
  mt_shortreads2koc()+write_fqkoc2files()  :657 / :516    Sketcher.fastq_koc_{device,host,file}()
  fasta2co()+wrt_co2cmpn_use_inn_subctx()  :218 / :625    Sketcher.fasta_co_{device,host}()
  run_stageI() combine + cofiles.stat  command_dist.c:408 write_sketch_dir()
  get_species_abundance()      command_composite.c:446    Sketcher.composite() / composite_tsv()

The CUDA library is mandatory: importing works without a GPU (so that symbol/ABI checks can run
on a CPU box), but every compute call raises MkError when no device is present — there is no CPU
fallback and nothing here touches oracle/.
"""
from __future__ import annotations

import ctypes as C
import os
import struct
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MK_LIB_PATH") or os.path.join(HERE, "libmkssd_b200.so")      # (MK_LIB_PATH: development builds)

MK_OK = 0
ERRORS = {
    -1: "MK_ERR_ARG", -2: "MK_ERR_PARAM", -3: "MK_ERR_CUDA", -4: "MK_ERR_NOMEM", -5: "MK_ERR_CROWDED",
    -6: "MK_ERR_LONG_LINE", -7: "MK_ERR_IO", -8: "MK_ERR_EMPTY_QUERY", -9: "MK_ERR_UNSUPPORTED",
}


class MkError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("%s (%d): %s" % (ERRORS.get(code, "MK_ERR_?"), code, msg))
        self.code = code


class MkInfo(C.Structure):
    _fields_ = [
        ("k", C.c_int), ("subk", C.c_int), ("drlevel", C.c_int), ("kmer_len", C.c_int), ("outctx", C.c_int),
        ("dim_end", C.c_int), ("hashsize", C.c_uint32), ("hashlimit", C.c_uint32), ("component_num", C.c_int),
        ("comp_code_bits", C.c_int), ("code_bits", C.c_int), ("device", C.c_int), ("sm_count", C.c_int),
    ]


class MkSketch(C.Structure):
    _fields_ = [
        ("n_components", C.c_int), ("n_total", C.c_uint64), ("n", C.POINTER(C.c_uint64)),
        ("codes", C.POINTER(C.POINTER(C.c_uint32))), ("counts", C.POINTER(C.POINTER(C.c_uint16))),
        ("borrowed", C.c_int),
    ]


class MkProfile(C.Structure):
    _fields_ = [
        ("stream_kernel_ms", C.c_double), ("stream_kernel_launches", C.c_uint64),
        ("stream_kernel_bytes", C.c_uint64), ("reduce_ms", C.c_double), ("composite_ms", C.c_double),
        ("kernel_launches", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
        ("exchange_ms", C.c_double), ("exchange_wait_ms", C.c_double),
    ]


class MkSpeciesStat(C.Structure):
    _fields_ = [("n", C.c_int32), ("sum", C.c_int32), ("lastsum", C.c_int32), ("lastn", C.c_int32),
                ("median", C.c_int32), ("max", C.c_int32)]


class MkRuns(C.Structure):
    _fields_ = [("n", C.c_uint64), ("d_code", C.c_void_p), ("d_firstpos", C.c_void_p), ("d_count", C.c_void_p)]


class MksParams(C.Structure):  # include/mkssd_synth.h
    _fields_ = [
        ("seed", C.c_uint64), ("n_species", C.c_uint32), ("genome_len", C.c_uint32), ("read_len", C.c_uint32),
        ("genus_size", C.c_uint32), ("shared_len", C.c_uint32), ("sub_thresh16", C.c_uint32),
        ("n_thresh16", C.c_uint32), ("n_present", C.c_uint32),
    ]


# every symbol include/mkssd_b200.h declares
EXPORTS = [
    "mk_strerror", "mk_last_error", "mk_device_count", "mk_ctx_create", "mk_ctx_destroy", "mk_ctx_info",
    "mk_ctx_profile", "mk_ctx_synchronize", "mk_ctx_cuda_stream", "mk_fastq_koc_device", "mk_fastq_koc_host", "mk_fastq_koc_file",
    "mk_fasta_co_device", "mk_fasta_co_host", "mk_fasta_co_file", "mk_fasta_co_files", "mk_fastq_co_device",
    "mk_fastq_co_host", "mk_fastq_co_file", "mk_ctx_set_dedup", "mk_ctx_set_borrowed_output", "mk_sketch_free", "mk_composite_begin",
    "mk_composite_component", "mk_composite_stats", "mk_composite_hits", "mk_markerdb_load", "mk_markerdb_unload",
    "mk_composite_component_resident", "mk_composite_component_last", "mk_format_species_coverage",
    "mk_fastq_partial_device", "mk_fastq_partial_host",
    "mk_runs_finalize_device", "mk_runs_finalize_distinct_device", "mk_runs_merge_device", "mk_count_newlines_device", "mk_synth_fastq_device",
    "mk_synth_fasta_device", "mk_synth_build", "mk_synth_free", "mk_synth_fastq_bytes", "mk_synth_fasta_bytes",
    "mk_synth_shuf_perm", "mk_synth_shuf_id",
    "mk_comm_unique_id", "mk_comm_init", "mk_comm_destroy", "mk_comm_last_block_need", "mk_markerdb_load_sharded", "mk_fastq_koc_sharded_device",
    "mk_fastq_koc_sharded_host", "mk_set_group", "mk_set_uniq_union", "mk_set_operate", "mk_free", "mk_shared_counts",
]

_lib = None


def load():
    """Load the CUDA library.  Fails loudly when it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libmkssd_b200.so is missing: run `python -m metakssd_b200.build` "
                          "(or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, sz, u64, i32 = C.c_void_p, C.c_size_t, C.c_uint64, C.c_int
    L.mk_strerror.restype = C.c_char_p
    L.mk_strerror.argtypes = [i32]
    L.mk_last_error.restype = C.c_char_p
    L.mk_last_error.argtypes = [vp]
    L.mk_device_count.restype = i32
    L.mk_ctx_create.argtypes = [C.POINTER(vp), vp, i32, i32, i32, i32]
    L.mk_ctx_destroy.argtypes = [vp]
    L.mk_ctx_destroy.restype = None
    L.mk_ctx_info.argtypes = [vp, C.POINTER(MkInfo)]
    L.mk_ctx_profile.argtypes = [vp, C.POINTER(MkProfile), i32]
    L.mk_ctx_synchronize.argtypes = [vp]
    L.mk_ctx_cuda_stream.argtypes = [vp]
    L.mk_ctx_cuda_stream.restype = vp
    L.mk_fastq_koc_device.argtypes = [vp, vp, sz, C.POINTER(MkSketch)]
    L.mk_fastq_koc_host.argtypes = [vp, vp, sz, C.POINTER(MkSketch)]
    L.mk_fastq_koc_file.argtypes = [vp, C.c_char_p, C.c_char_p, C.POINTER(MkSketch)]
    L.mk_fasta_co_device.argtypes = [vp, vp, vp, i32, C.POINTER(MkSketch)]
    L.mk_fasta_co_host.argtypes = [vp, vp, vp, i32, C.POINTER(MkSketch)]
    L.mk_fasta_co_file.argtypes = [vp, C.c_char_p, C.c_char_p, C.POINTER(MkSketch)]
    L.mk_fasta_co_files.argtypes = [vp, vp, i32, C.c_char_p, C.POINTER(MkSketch)]
    L.mk_fastq_co_device.argtypes = [vp, vp, sz, i32, i32, C.POINTER(MkSketch)]
    L.mk_fastq_co_host.argtypes = [vp, vp, sz, i32, i32, C.POINTER(MkSketch)]
    L.mk_fastq_co_file.argtypes = [vp, C.c_char_p, C.c_char_p, i32, i32, C.POINTER(MkSketch)]
    L.mk_ctx_set_dedup.argtypes = [vp, i32]
    L.mk_ctx_set_borrowed_output.argtypes = [vp, i32]
    L.mk_sketch_free.argtypes = [C.POINTER(MkSketch)]
    L.mk_sketch_free.restype = None
    L.mk_composite_begin.argtypes = [vp, i32]
    L.mk_composite_component.argtypes = [vp, vp, vp, i32, vp, vp, u64, u64]
    L.mk_composite_stats.argtypes = [vp, vp]
    L.mk_markerdb_load.argtypes = [vp, i32, vp, vp, i32]
    L.mk_markerdb_unload.argtypes = [vp]
    L.mk_composite_component_resident.argtypes = [vp, i32, vp, vp, u64, u64]
    L.mk_composite_component_last.argtypes = [vp, i32]
    L.mk_format_species_coverage.argtypes = [C.c_char_p, vp, vp, i32, vp, sz]
    L.mk_format_species_coverage.restype = sz
    L.mk_composite_hits.argtypes = [vp, C.POINTER(C.POINTER(C.POINTER(C.c_int32)))]
    L.mk_fastq_partial_device.argtypes = [vp, vp, sz, u64, u64, i32, C.POINTER(MkRuns)]
    L.mk_fastq_partial_host.argtypes = [vp, vp, sz, u64, u64, i32, C.POINTER(MkRuns)]
    L.mk_runs_finalize_device.argtypes = [vp, vp, vp, vp, u64, C.POINTER(MkSketch)]
    L.mk_runs_finalize_distinct_device.argtypes = [vp, vp, vp, vp, u64, C.POINTER(MkSketch)]
    L.mk_runs_merge_device.argtypes = [vp, vp, vp, vp, u64, C.POINTER(MkRuns)]
    L.mk_count_newlines_device.argtypes = [vp, vp, sz, C.POINTER(u64)]
    L.mk_synth_fastq_device.argtypes = [vp, C.POINTER(MksParams), vp, vp, u64, u64, vp, sz, C.POINTER(sz)]
    L.mk_synth_fasta_device.argtypes = [vp, C.POINTER(MksParams), C.c_uint32, C.c_uint32, vp, sz, vp]
    L.mk_synth_build.argtypes = [C.POINTER(MksParams), u64, C.c_uint32, C.c_uint32, C.c_uint32,
                                 C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.POINTER(C.c_uint32))]
    L.mk_synth_free.argtypes = [vp]
    L.mk_synth_free.restype = None
    L.mk_synth_fastq_bytes.argtypes = [C.POINTER(MksParams), u64, u64]
    L.mk_synth_fastq_bytes.restype = u64
    L.mk_synth_fasta_bytes.argtypes = [C.POINTER(MksParams), C.c_uint32]
    L.mk_synth_fasta_bytes.restype = u64
    L.mk_synth_shuf_perm.argtypes = [u64, i32, vp]
    L.mk_synth_shuf_perm.restype = None
    L.mk_synth_shuf_id.argtypes = [u64]
    L.mk_synth_shuf_id.restype = C.c_int32
    L.mk_set_group.argtypes = [vp, vp, vp, i32, vp, i32, C.POINTER(C.POINTER(C.c_uint32)), vp]
    L.mk_set_uniq_union.argtypes = [vp, vp, u64, C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(u64)]
    L.mk_set_operate.argtypes = [vp, vp, u64, vp, vp, i32, i32, C.POINTER(C.POINTER(C.c_uint32)), vp]
    L.mk_free.argtypes = [vp]
    L.mk_free.restype = None
    L.mk_comm_unique_id.argtypes = [vp, sz]
    L.mk_comm_init.argtypes = [vp, vp, i32, i32]
    L.mk_comm_destroy.argtypes = [vp]
    L.mk_shared_counts.argtypes = [vp, vp, vp, C.c_int, vp, vp, C.c_int, vp, vp]
    L.mk_comm_last_block_need.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.mk_markerdb_load_sharded.argtypes = [vp, i32, vp, vp, i32]
    L.mk_fastq_koc_sharded_device.argtypes = [vp, vp, sz, u64, u64, i32, u64, C.POINTER(MkSketch), vp]
    L.mk_fastq_koc_sharded_host.argtypes = [vp, vp, sz, u64, u64, i32, u64, C.POINTER(MkSketch), vp]
    _lib = L
    return L


def device_count() -> int:
    return int(load().mk_device_count())


# ------------------------------------------------------------------------------------ generator
@dataclass
class SynthSpec:
    """Synthetic community (include/mkssd_synth.h): parameters + abundance CDF."""
    P: MksParams
    cdf32: np.ndarray
    species: np.ndarray

    def fastq_bytes(self, r0: int, r1: int) -> int:
        return int(load().mk_synth_fastq_bytes(C.byref(self.P), r0, r1))

    def fasta_bytes(self, s: int) -> int:
        return int(load().mk_synth_fasta_bytes(C.byref(self.P), s))


def synth_spec(seed: int, n_species: int, genome_len: int, read_len: int = 150) -> SynthSpec:
    L = load()
    P = MksParams()
    cdf = C.POINTER(C.c_uint32)()
    spc = C.POINTER(C.c_uint32)()
    rc = L.mk_synth_build(C.byref(P), seed, n_species, genome_len, read_len, C.byref(cdf), C.byref(spc))
    if rc != MK_OK:
        raise MkError(rc, L.mk_strerror(rc).decode())
    n = P.n_present
    out = SynthSpec(P, np.ctypeslib.as_array(cdf, shape=(n,)).copy(), np.ctypeslib.as_array(spc, shape=(n,)).copy())
    L.mk_synth_free(cdf)
    L.mk_synth_free(spc)
    return out


def make_shuf(seed: int, subk: int):
    """Deterministic .shuf content: (shuf_id, int32 permutation of 16^subk)."""
    L = load()
    perm = np.empty(1 << (4 * subk), dtype=np.int32)
    L.mk_synth_shuf_perm(seed, subk, perm.ctypes.data)
    return int(L.mk_synth_shuf_id(seed)), perm


# ------------------------------------------------------------------------------------ .shuf files
def read_shuf(path: str):
    """(shuf_id, k, subk, drlevel, int32 permutation) — format of command_shuffle.c:164-235."""
    with open(path, "rb") as f:
        hdr = f.read(16)
        shuf_id, k, subk, drlevel = struct.unpack("<iiii", hdr)
        perm = np.fromfile(f, dtype=np.int32, count=1 << (4 * subk))
    if perm.size != 1 << (4 * subk):
        raise ValueError("truncated .shuf file %s" % path)
    return shuf_id, k, subk, drlevel, perm


def write_shuf(path: str, shuf_id: int, k: int, subk: int, drlevel: int, perm: np.ndarray) -> None:
    with open(path, "wb") as f:
        f.write(struct.pack("<iiii", shuf_id, k, subk, drlevel))
        f.write(np.ascontiguousarray(perm, dtype=np.int32).tobytes())


# ------------------------------------------------------------------------------------ results
@dataclass
class Sketch:
    """Per-component arrays exactly as the reference writes `<i>.co.<c>` / `<i>.co.<c>.a`."""
    codes: list            # [component] -> uint32[]
    counts: list | None    # [component] -> uint16[]   (None for FASTA sketches)

    @property
    def n_total(self) -> int:
        return int(sum(c.size for c in self.codes))


def _take_sketch(sk: MkSketch, with_counts: bool) -> Sketch:
    """numpy arrays of a returned sketch.  A borrowed sketch (Sketcher.set_borrowed_output) is NOT copied: its arrays
    are views of the context's pinned staging block, valid until the next call on that context."""
    codes, counts = [], ([] if with_counts else None)
    borrowed = bool(sk.borrowed)
    for c in range(sk.n_components):
        n = int(sk.n[c])
        a = np.ctypeslib.as_array(sk.codes[c], shape=(n,)) if n else np.empty(0, np.uint32)
        codes.append(a if borrowed or not n else a.copy())
        if with_counts:
            b = np.ctypeslib.as_array(sk.counts[c], shape=(n,)) if n else np.empty(0, np.uint16)
            counts.append(b if borrowed or not n else b.copy())
    load().mk_sketch_free(C.byref(sk))
    return Sketch(codes, counts)


def _ptr(x) -> int:
    """Device pointer of a torch tensor / raw int, host pointer of a numpy array."""
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        return int(x.data_ptr())
    raise TypeError(type(x))


class Sketcher:
    """One sketching context bound to one GPU (mk_ctx)."""

    def __init__(self, perm: np.ndarray, k: int, subk: int, drlevel: int, device: int = 0):
        self._L = load()
        self._h = C.c_void_p()
        if perm is None:          # a context without pass-set tables: composite, set operations, dist -r
            ptr = None
        else:
            perm = np.ascontiguousarray(perm, dtype=np.int32)
            if perm.size != 1 << (4 * subk):
                raise ValueError("permutation must have 16^subk entries")
            ptr = perm.ctypes.data
        rc = self._L.mk_ctx_create(C.byref(self._h), ptr, k, subk, drlevel, device)
        if rc != MK_OK:
            self._h = C.c_void_p()
            raise MkError(rc, self._L.mk_strerror(rc).decode())
        self.info = MkInfo()
        self.device = device
        self._L.mk_ctx_info(self._h, C.byref(self.info))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.mk_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc: int):
        if rc != MK_OK:
            raise MkError(rc, self._L.mk_last_error(self._h).decode() or self._L.mk_strerror(rc).decode())

    def cuda_stream(self) -> int:
        """cudaStream_t of the context (wrap with torch.cuda.ExternalStream to record events on it)."""
        return int(self._L.mk_ctx_cuda_stream(self._h) or 0)

    def synchronize(self):
        self._ck(self._L.mk_ctx_synchronize(self._h))

    def profile(self, reset: bool = False) -> MkProfile:
        p = MkProfile()
        self._ck(self._L.mk_ctx_profile(self._h, C.byref(p), 1 if reset else 0))
        return p

    # -- FASTQ -A
    def fastq_koc_device(self, d_text, nbytes: int) -> Sketch:
        sk = MkSketch()
        self._ck(self._L.mk_fastq_koc_device(self._h, _ptr(d_text), nbytes, C.byref(sk)))
        return _take_sketch(sk, True)

    def fastq_koc_host(self, text) -> Sketch:
        a = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray)) else text
        sk = MkSketch()
        if isinstance(a, np.ndarray):
            a = np.ascontiguousarray(a, dtype=np.uint8)
            self._ck(self._L.mk_fastq_koc_host(self._h, a.ctypes.data, a.size, C.byref(sk)))
        else:  # pinned torch CPU tensor
            self._ck(self._L.mk_fastq_koc_host(self._h, int(a.data_ptr()), int(a.numel()), C.byref(sk)))
        return _take_sketch(sk, True)

    def fastq_koc_file(self, path: str, pipecmd: str = "") -> Sketch:
        sk = MkSketch()
        self._ck(self._L.mk_fastq_koc_file(self._h, path.encode(), pipecmd.encode(), C.byref(sk)))
        return _take_sketch(sk, True)

    # -- FASTQ without -A (`dist -Q q -n m`)
    def fastq_co_device(self, d_text, nbytes: int, quality: int = 0, min_occurrence: int = 1) -> Sketch:
        sk = MkSketch()
        self._ck(self._L.mk_fastq_co_device(self._h, _ptr(d_text), nbytes, quality, min_occurrence, C.byref(sk)))
        return _take_sketch(sk, False)

    def fastq_co_host(self, text, quality: int = 0, min_occurrence: int = 1) -> Sketch:
        a = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray)) else np.ascontiguousarray(text, dtype=np.uint8)
        sk = MkSketch()
        self._ck(self._L.mk_fastq_co_host(self._h, a.ctypes.data, a.size, quality, min_occurrence, C.byref(sk)))
        return _take_sketch(sk, False)

    def fastq_co_file(self, path: str, pipecmd: str = "", quality: int = 0, min_occurrence: int = 1) -> Sketch:
        sk = MkSketch()
        self._ck(self._L.mk_fastq_co_file(self._h, path.encode(), pipecmd.encode(), quality, min_occurrence, C.byref(sk)))
        return _take_sketch(sk, False)

    def set_borrowed_output(self, on: bool):
        """Sketches of the following calls come back as views into the pinned staging block (no host copies)."""
        self._ck(self._L.mk_ctx_set_borrowed_output(self._h, 1 if on else 0))

    def set_dedup(self, on: bool):
        """`dist -u`: the following fasta_co_* calls keep only codes that occur once in their genome."""
        self._ck(self._L.mk_ctx_set_dedup(self._h, 1 if on else 0))

    # -- FASTA
    def fasta_co_device(self, d_text, offsets) -> list:
        off = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = off.size - 1
        arr = (MkSketch * n)()
        self._ck(self._L.mk_fasta_co_device(self._h, _ptr(d_text), off.ctypes.data, n, arr))
        return [_take_sketch(arr[i], False) for i in range(n)]

    def fasta_co_host(self, texts) -> list:
        """texts: list of byte strings / uint8 arrays, one per FASTA file."""
        arrs = [np.frombuffer(t, dtype=np.uint8) if isinstance(t, (bytes, bytearray)) else
                np.ascontiguousarray(t, dtype=np.uint8) for t in texts]
        off = np.zeros(len(arrs) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([a.size for a in arrs])
        buf = np.concatenate(arrs) if arrs else np.empty(0, np.uint8)
        n = len(arrs)
        arr = (MkSketch * n)()
        self._ck(self._L.mk_fasta_co_host(self._h, buf.ctypes.data, off.ctypes.data, n, arr))
        return [_take_sketch(arr[i], False) for i in range(n)]

    def fasta_co_files(self, paths, pipecmd: str = "") -> list:
        """The FASTA file loop of run_stageI() as one batched call (mk_fasta_co_files)."""
        n = len(paths)
        enc = [p.encode() for p in paths]
        carr = (C.c_char_p * n)(*enc)
        arr = (MkSketch * n)()
        self._ck(self._L.mk_fasta_co_files(self._h, carr, n, pipecmd.encode(), arr))
        return [_take_sketch(arr[i], False) for i in range(n)]

    def fasta_co_file(self, path: str, pipecmd: str = "") -> Sketch:
        sk = MkSketch()
        self._ck(self._L.mk_fasta_co_file(self._h, path.encode(), pipecmd.encode(), C.byref(sk)))
        return _take_sketch(sk, False)

    # -- composite
    def load_markerdb(self, ref_comp):
        """Keep a MarkerDB (per component: codes uint32[], index uint64[S+1]) resident on the device;
        composite(None, qry) then intersects against it without re-uploading."""
        self._ck(self._L.mk_markerdb_unload(self._h))
        self._mdb_species = int(ref_comp[0][1].size - 1)
        self._mdb_components = len(ref_comp)
        for c, (rc, ri) in enumerate(ref_comp):
            rc = np.ascontiguousarray(rc, dtype=np.uint32)
            ri = np.ascontiguousarray(ri, dtype=np.uint64)
            self._ck(self._L.mk_markerdb_load(self._h, c, rc.ctypes.data, ri.ctypes.data, self._mdb_species))

    def composite_last(self):
        """Resident MarkerDB (load_markerdb) against the -A sketch this context produced last, which is
        still on the device: no upload at all.  Returns the per-species stats array."""
        S = self._mdb_species
        self._ck(self._L.mk_composite_begin(self._h, S))
        for c in range(self._mdb_components):
            self._ck(self._L.mk_composite_component_last(self._h, c))
        stats = np.zeros(S, dtype=STATS_DTYPE)
        self._ck(self._L.mk_composite_stats(self._h, stats.ctypes.data))
        return stats

    def composite(self, ref_comp, qry_comp, want_lists: bool = False):
        """ref_comp: per component (codes uint32[], index uint64[S+1]); qry_comp: per component
        (codes uint32[], counts uint16[]) of ONE query.  Returns the per-species stats array
        (structured: n,sum,lastsum,lastn,median,max) and, optionally, the raw hit lists."""
        if ref_comp is None:                      # resident MarkerDB (load_markerdb)
            S = self._mdb_species
            self._ck(self._L.mk_composite_begin(self._h, S))
            for c, (qc, qa) in enumerate(qry_comp[:self._mdb_components]):
                qc = np.ascontiguousarray(qc, dtype=np.uint32)
                qa = np.ascontiguousarray(qa, dtype=np.uint16)
                self._ck(self._L.mk_composite_component_resident(self._h, c, qc.ctypes.data, qa.ctypes.data, 0, qc.size))
            ref_comp = ()
        else:
            S = int(ref_comp[0][1].size - 1)
            self._ck(self._L.mk_composite_begin(self._h, S))
        for (rc, ri), (qc, qa) in zip(ref_comp, qry_comp):
            rc = np.ascontiguousarray(rc, dtype=np.uint32)
            ri = np.ascontiguousarray(ri, dtype=np.uint64)
            qc = np.ascontiguousarray(qc, dtype=np.uint32)
            qa = np.ascontiguousarray(qa, dtype=np.uint16)
            self._ck(self._L.mk_composite_component(self._h, rc.ctypes.data, ri.ctypes.data, S, qc.ctypes.data,
                                                    qa.ctypes.data, 0, qc.size))
        stats = np.zeros(S, dtype=STATS_DTYPE)
        self._ck(self._L.mk_composite_stats(self._h, stats.ctypes.data))
        if not want_lists:
            return stats
        lists = C.POINTER(C.POINTER(C.c_int32))()
        self._ck(self._L.mk_composite_hits(self._h, C.byref(lists)))
        out = []
        for s in range(S):
            n = lists[s][0]
            out.append(np.ctypeslib.as_array(lists[s], shape=(n + 1,)).copy())
        return stats, out

    # -- multi-GPU building blocks
    def fastq_partial_device(self, d_text, nbytes: int, pos_base: int, line_base: int, is_last: bool) -> MkRuns:
        r = MkRuns()
        self._ck(self._L.mk_fastq_partial_device(self._h, _ptr(d_text), nbytes, pos_base, line_base,
                                                 1 if is_last else 0, C.byref(r)))
        return r

    def fastq_partial_host(self, h_text, nbytes: int, pos_base: int, line_base: int, is_last: bool) -> MkRuns:
        r = MkRuns()
        self._ck(self._L.mk_fastq_partial_host(self._h, _ptr(h_text), nbytes, pos_base, line_base,
                                               1 if is_last else 0, C.byref(r)))
        return r

    def runs_finalize_device(self, d_code, d_pos, d_cnt, n: int, distinct: bool = False) -> Sketch:
        sk = MkSketch()
        fn = self._L.mk_runs_finalize_distinct_device if distinct else self._L.mk_runs_finalize_device
        self._ck(fn(self._h, _ptr(d_code), _ptr(d_pos), _ptr(d_cnt), n, C.byref(sk)))
        return _take_sketch(sk, True)

    def runs_merge_device(self, d_code, d_pos, d_cnt, n: int) -> MkRuns:
        r = MkRuns()
        self._ck(self._L.mk_runs_merge_device(self._h, _ptr(d_code), _ptr(d_pos), _ptr(d_cnt), n, C.byref(r)))
        return r

    # -- `set -g / -q / -i` (MarkerDB build), one component per call
    def _take_u32(self, ptr, n: int) -> np.ndarray:
        out = np.ctypeslib.as_array(ptr, shape=(n,)).copy() if n else np.empty(0, np.uint32)
        self._L.mk_free(ptr)
        return out

    def set_group(self, codes, index, taxon_of_genome, n_taxa: int):
        codes = np.ascontiguousarray(codes, dtype=np.uint32)
        index = np.ascontiguousarray(index, dtype=np.uint64)
        tax = np.ascontiguousarray(taxon_of_genome, dtype=np.int32)
        out = C.POINTER(C.c_uint32)()
        oi = np.zeros(n_taxa + 1, dtype=np.uint64)
        self._ck(self._L.mk_set_group(self._h, codes.ctypes.data, index.ctypes.data, index.size - 1, tax.ctypes.data, n_taxa,
                                      C.byref(out), oi.ctypes.data))
        return self._take_u32(out, int(oi[-1])), oi

    def set_uniq_union(self, codes) -> np.ndarray:
        codes = np.ascontiguousarray(codes, dtype=np.uint32)
        out = C.POINTER(C.c_uint32)()
        n = C.c_uint64()
        self._ck(self._L.mk_set_uniq_union(self._h, codes.ctypes.data, codes.size, C.byref(out), C.byref(n)))
        return self._take_u32(out, int(n.value))

    def set_operate(self, pan, codes, index, intersect: bool = True):
        pan = np.ascontiguousarray(pan, dtype=np.uint32)
        codes = np.ascontiguousarray(codes, dtype=np.uint32)
        index = np.ascontiguousarray(index, dtype=np.uint64)
        out = C.POINTER(C.c_uint32)()
        oi = np.zeros(index.size, dtype=np.uint64)
        self._ck(self._L.mk_set_operate(self._h, pan.ctypes.data, pan.size, codes.ctypes.data, index.ctypes.data, index.size - 1,
                                        1 if intersect else 0, C.byref(out), oi.ctypes.data))
        return self._take_u32(out, int(oi[-1])), oi

    # -- the multi-GPU step inside the library (NCCL; csrc/mk_comm.cu)
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        rc = load().mk_comm_unique_id(buf, 128)
        if rc != MK_OK:
            raise MkError(rc, load().mk_strerror(rc).decode())
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        self._ck(self._L.mk_comm_init(self._h, C.c_char_p(unique_id), rank, world))
        self.rank, self.world = rank, world

    def shared_counts(self, ref_comp, qry_comp, qry_ctx_ct=None) -> np.ndarray:
        """`dist -r`: ref_comp / qry_comp = lists over components of (codes uint32, index uint64[n+1]); returns the
        shared k-mer counts uint32[n_qry, n_ref] (mco_cbdco_nobin_dist, command_dist.c:1031-1046)."""
        n_ref, n_qry = int(ref_comp[0][1].size - 1), int(qry_comp[0][1].size - 1)
        counts = np.zeros((n_qry, n_ref), dtype=np.uint32)
        ct = None if qry_ctx_ct is None else np.ascontiguousarray(qry_ctx_ct, dtype=np.uint32)
        for (rc, ri), (qc, qi) in zip(ref_comp, qry_comp):
            rc = np.ascontiguousarray(rc, dtype=np.uint32); ri = np.ascontiguousarray(ri, dtype=np.uint64)
            qc = np.ascontiguousarray(qc, dtype=np.uint32); qi = np.ascontiguousarray(qi, dtype=np.uint64)
            self._ck(self._L.mk_shared_counts(self._h, rc.ctypes.data, ri.ctypes.data, n_ref, qc.ctypes.data, qi.ctypes.data,
                                              n_qry, None if ct is None else ct.ctypes.data, counts.ctypes.data))
        return counts

    def comm_last_block_need(self) -> int:
        """Largest exchange block this rank saw in the last sharded step (valid after MK_ERR_NOMEM too)."""
        v = C.c_uint64()
        self._ck(self._L.mk_comm_last_block_need(self._h, C.byref(v)))
        return int(v.value)

    def load_markerdb_sharded(self, ref_comp):
        """Every rank passes the WHOLE MarkerDB; the library keeps this rank's code-range slice resident."""
        self._ck(self._L.mk_markerdb_unload(self._h))
        self._mdb_species = int(ref_comp[0][1].size - 1)
        self._mdb_components = len(ref_comp)
        for c, (rc, ri) in enumerate(ref_comp):
            rc = np.ascontiguousarray(rc, dtype=np.uint32)
            ri = np.ascontiguousarray(ri, dtype=np.uint64)
            self._ck(self._L.mk_markerdb_load_sharded(self._h, c, rc.ctypes.data, ri.ctypes.data, self._mdb_species))

    def fastq_koc_sharded(self, text, nbytes: int, pos_base: int, line_base: int, is_last: bool, max_runs: int,
                          host_text: bool = False, want_stats: bool = True):
        """Collective.  Returns (Sketch, stats) on rank 0 and (None, None) elsewhere."""
        sk = MkSketch()
        root = getattr(self, "rank", 0) == 0
        stats = np.zeros(getattr(self, "_mdb_species", 0), dtype=STATS_DTYPE) if (root and want_stats and getattr(self, "_mdb_species", 0)) else None
        fn = self._L.mk_fastq_koc_sharded_host if host_text else self._L.mk_fastq_koc_sharded_device
        self._ck(fn(self._h, _ptr(text), nbytes, pos_base, line_base, 1 if is_last else 0, max_runs,
                    C.byref(sk) if root else None, stats.ctypes.data if stats is not None else None))
        if not root:
            return None, None
        return _take_sketch(sk, True), stats

    def count_newlines_device(self, d_text, nbytes: int) -> int:
        v = C.c_uint64()
        self._ck(self._L.mk_count_newlines_device(self._h, _ptr(d_text), nbytes, C.byref(v)))
        return int(v.value)

    # -- synthetic workload on the device
    def synth_fastq_device(self, P: MksParams, cdf32: np.ndarray, species: np.ndarray, r0: int, r1: int, d_out,
                           capacity: int) -> int:
        w = C.c_size_t()
        cdf32 = np.ascontiguousarray(cdf32, dtype=np.uint32)
        species = np.ascontiguousarray(species, dtype=np.uint32)
        self._ck(self._L.mk_synth_fastq_device(self._h, C.byref(P), cdf32.ctypes.data, species.ctypes.data, r0, r1,
                                               _ptr(d_out), capacity, C.byref(w)))
        return int(w.value)

    def synth_fasta_device(self, P: MksParams, s0: int, s1: int, d_out, capacity: int) -> np.ndarray:
        off = np.zeros(s1 - s0 + 1, dtype=np.uint64)
        self._ck(self._L.mk_synth_fasta_device(self._h, C.byref(P), s0, s1, _ptr(d_out), capacity, off.ctypes.data))
        return off


# ------------------------------------------------------------------------------------ host mirror
CO_DSTAT = struct.Struct("<I?3xiiiiQ")  # co_dstat_t, global_basic.h:116-126


def write_sketch_dir(path: str, shuf_id: int, info: MkInfo, names, sketches, koc: bool) -> None:
    """What run_stageI() leaves on disk (command_dist.c:408-500): combco.<c>, combco.<c>.a,
    combco.index.<c> and cofiles.stat, for `sketches[i]` of input `names[i]`."""
    os.makedirs(path, exist_ok=True)
    cn = info.component_num
    ctx_ct = np.array([s.n_total for s in sketches], dtype=np.uint32)
    for c in range(cn):
        idx = np.zeros(len(sketches) + 1, dtype=np.uint64)
        idx[1:] = np.cumsum([s.codes[c].size for s in sketches])
        with open(os.path.join(path, "combco.%d" % c), "wb") as f:
            for s in sketches:
                f.write(s.codes[c].astype(np.uint32).tobytes())
        idx.tofile(os.path.join(path, "combco.index.%d" % c))
        if koc:
            with open(os.path.join(path, "combco.%d.a" % c), "wb") as f:
                for s in sketches:
                    f.write(s.counts[c].astype(np.uint16).tobytes())
    with open(os.path.join(path, "cofiles.stat"), "wb") as f:
        f.write(CO_DSTAT.pack(shuf_id & 0xFFFFFFFF, bool(koc), 2 * info.k, 2 * info.drlevel, cn, len(sketches),
                              int(ctx_ct.sum())))
        f.write(ctx_ct.tobytes())
        for n in names:
            b = n.encode()[:255]
            f.write(b + b"\0" * (256 - len(b)))


def read_sketch_dir(path: str):
    """(header dict, names, per-component combco, index, abund|None)"""
    raw = open(os.path.join(path, "cofiles.stat"), "rb").read()
    shuf_id, koc, kmerlen, dim_rd_len, comp_num, infile_num, all_ctx = CO_DSTAT.unpack_from(raw, 0)
    off = CO_DSTAT.size + 4 * infile_num
    names = [raw[off + 256 * i: off + 256 * (i + 1)].split(b"\0", 1)[0].decode() for i in range(infile_num)]
    combco = [np.fromfile(os.path.join(path, "combco.%d" % c), dtype=np.uint32) for c in range(comp_num)]
    index = [np.fromfile(os.path.join(path, "combco.index.%d" % c), dtype=np.uint64) for c in range(comp_num)]
    abund = [np.fromfile(os.path.join(path, "combco.%d.a" % c), dtype=np.uint16) for c in range(comp_num)] if koc else None
    hdr = dict(shuf_id=shuf_id, koc=bool(koc), kmerlen=kmerlen, dim_rd_len=dim_rd_len, comp_num=comp_num,
               infile_num=infile_num, all_ctx_ct=all_ctx)
    return hdr, names, combco, index, abund


STATS_DTYPE = np.dtype([("n", "<i4"), ("sum", "<i4"), ("lastsum", "<i4"), ("lastn", "<i4"), ("median", "<i4"),
                        ("max", "<i4")])


class SpeciesNames:
    """Species names prepared once for mk_format_species_coverage (a C array of C strings)."""

    def __init__(self, names):
        self.names = list(names)
        self._enc = [n.encode() for n in self.names]
        self.carr = (C.c_char_p * len(self._enc))(*self._enc)


def coverage_tsv(qry_name: str, names: SpeciesNames, stats) -> str:
    """species_coverage lines through the library's C formatter (same text as composite_tsv)."""
    L = load()
    stats = np.ascontiguousarray(stats, dtype=STATS_DTYPE)
    q = qry_name.encode()
    need = L.mk_format_species_coverage(q, names.carr, stats.ctypes.data, len(names.names), None, 0)
    buf = C.create_string_buffer(need + 1)
    L.mk_format_species_coverage(q, names.carr, stats.ctypes.data, len(names.names), buf, need + 1)
    return buf.raw[:need].decode()


def composite_tsv(qry_name: str, ref_names, stats) -> str:
    """species_coverage lines exactly as command_composite.c:582-624 prints them: species ordered
    by matched k-mers (descending, ties by index as glibc's stable qsort leaves them), stopping at
    the first with fewer than MIN_KM_S = 6; ratios are computed in float32 and printed with %f."""
    n_all = stats["n"].astype(np.int64)
    order = np.argsort(-n_all, kind="stable")
    sel = order[:int(np.count_nonzero(n_all >= 6))]          # a prefix of the descending order
    n = n_all[sel]
    with np.errstate(divide="ignore", invalid="ignore"):
        mean = stats["sum"][sel].astype(np.int32).astype(np.float32) / n.astype(np.float32)
        last = stats["lastsum"][sel].astype(np.int32).astype(np.float32) / stats["lastn"][sel].astype(np.int64).astype(np.float32)
    lines = ["%s\t%s\t%d\t%f\t%f\t%d\t%d\n" % (qry_name, ref_names[s], a, b, c, d, e)
             for s, a, b, c, d, e in zip(sel.tolist(), n.tolist(), mean.tolist(), last.tolist(),
                                         stats["median"][sel].tolist(), stats["max"][sel].tolist())]
    return "".join(lines)
