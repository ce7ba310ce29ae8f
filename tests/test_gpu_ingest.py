"""GPU: the streaming file entry points (mk_fastq_koc_file / mk_fasta_co_files, csrc/mk_ingest.cu) against the
oracle — plain files read directly, gzip input and an explicit pipe command through popen(), chunk sizes
small enough that a file is cut into hundreds of line-aligned chunks, and every way a FASTQ file can end."""
import gzip
import os
import subprocess

import numpy as np
import pytest

from helpers import same_sketch

pytestmark = pytest.mark.gpu
K, SUBK, L = 11, 6, 3


@pytest.fixture(scope="module")
def ctx(lib_built, oracle, shuf):
    sid, perm = shuf(777, K, SUBK, L)
    sk = lib_built.Sketcher(perm, K, SUBK, L)
    yield sk, perm, oracle.params(K, SUBK, L)
    sk.close()


def _fastq(oracle, n_reads, tail):
    S = oracle.synth(21, 12, 150_000, 150)
    text = bytes(S.fastq(0, n_reads))
    if tail == "no_newline":
        text = text[:-1]
    elif tail == "partial_record":
        text += b"@last\nACGTACGTACGTACGTACGTACGTACGTAC"
    elif tail == "blank_lines":
        text += b"\n\n"
    elif tail == "half_record":
        text += b"@last\nACGTACGTACGTACGTACGTACGTACGTAC\n+\n"
    return text


@pytest.mark.parametrize("tail", ["newline", "no_newline", "partial_record", "blank_lines", "half_record"])
@pytest.mark.parametrize("chunk", [8192, 20000, 1 << 20, 1 << 26])
def test_plain_file_chunks(ctx, oracle, tmp_path, chunk, tail):
    sk, perm, p = ctx
    text = _fastq(oracle, 9000, tail)
    path = tmp_path / "reads.fq"
    path.write_bytes(text)
    os.environ["MK_INGEST_CHUNK_BYTES"] = str(chunk)
    try:
        got = sk.fastq_koc_file(str(path))
    finally:
        os.environ.pop("MK_INGEST_CHUNK_BYTES", None)
    same_sketch(got, oracle.fastq_koc(p, perm, np.frombuffer(text, np.uint8)), p)


@pytest.mark.parametrize("mode", ["gz", "pipecmd"])
@pytest.mark.parametrize("chunk", [8192, 70000, 1 << 26])
def test_pipe_sources(ctx, oracle, tmp_path, mode, chunk):
    sk, perm, p = ctx
    text = _fastq(oracle, 7000, "partial_record")
    os.environ["MK_INGEST_CHUNK_BYTES"] = str(chunk)
    try:
        if mode == "gz":
            path = tmp_path / "reads.fq.gz"
            with gzip.open(path, "wb", compresslevel=1) as f:
                f.write(text)
            got = sk.fastq_koc_file(str(path))
        else:
            path = tmp_path / "reads.fq"
            path.write_bytes(text)
            got = sk.fastq_koc_file(str(path), "cat")
    finally:
        os.environ.pop("MK_INGEST_CHUNK_BYTES", None)
    same_sketch(got, oracle.fastq_koc(p, perm, np.frombuffer(text, np.uint8)), p)


def test_chunk_boundary_sweep(ctx, oracle, tmp_path):
    """The text length sweeps across a chunk boundary byte by byte (the last chunk must never be a tiny rest)."""
    sk, perm, p = ctx
    base = _fastq(oracle, 60, "newline")
    os.environ["MK_INGEST_CHUNK_BYTES"] = "8192"
    try:
        for cutoff in list(range(8192 - 40, 8192 + 40, 3)) + list(range(16384 - 330, 16384 + 10, 7)):
            text = base[:cutoff]
            path = tmp_path / "sweep.fq"
            path.write_bytes(text)
            want = oracle.fastq_koc(p, perm, np.frombuffer(text, np.uint8))
            same_sketch(sk.fastq_koc_file(str(path)), want, p)
            same_sketch(sk.fastq_koc_file(str(path), "cat"), want, p)
    finally:
        os.environ.pop("MK_INGEST_CHUNK_BYTES", None)


def test_long_line_is_refused_across_chunks(ctx, oracle, lib_built, tmp_path):
    sk, perm, p = ctx
    text = _fastq(oracle, 200, "newline") + b"@long\n" + b"ACGT" * 1100 + b"\n+\n" + b"I" * 4400 + b"\n"
    path = tmp_path / "long.fq"
    path.write_bytes(text)
    os.environ["MK_INGEST_CHUNK_BYTES"] = "8192"
    try:
        with pytest.raises(lib_built.MkError) as e:
            sk.fastq_koc_file(str(path))
        assert e.value.code == -6
    finally:
        os.environ.pop("MK_INGEST_CHUNK_BYTES", None)


def test_missing_file_is_an_io_error(ctx, lib_built, tmp_path):
    sk, _, _ = ctx
    with pytest.raises(lib_built.MkError) as e:
        sk.fastq_koc_file(str(tmp_path / "nope.fq"))
    assert e.value.code == -7


def test_fasta_files_batched(ctx, oracle, tmp_path):
    """mk_fasta_co_files: plain and gzip genomes in one call, batch buffer smaller than the input."""
    sk, perm, p = ctx
    S = oracle.synth(5, 10, 120_000, 150)
    paths, want = [], []
    for s in range(10):
        fa = bytes(S.fasta(s))
        if s % 3 == 2:
            path = tmp_path / ("g%d.fasta.gz" % s)
            with gzip.open(path, "wb", compresslevel=1) as f:
                f.write(fa)
        else:
            path = tmp_path / ("g%d.fasta" % s)
            path.write_bytes(fa)
        paths.append(str(path))
        want.append(oracle.fasta_co(p, perm, np.frombuffer(fa, np.uint8)))
    for batch in (1 << 20, 1 << 30):
        os.environ["MK_FASTA_BATCH_BYTES"] = str(batch)
        try:
            got = sk.fasta_co_files(paths)
        finally:
            os.environ.pop("MK_FASTA_BATCH_BYTES", None)
        for g, w in zip(got, want):
            same_sketch(g, w, p)


def test_cli_rss_is_bounded(lib_built, oracle, tmp_path):
    """`metakssd-b200 dist -A` on a file of ~190 MB: peak RSS stays far below the input size + pinned ring."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-C", os.path.join(root, "host")], check=True, capture_output=True)
    cli = os.path.join(root, "host", "metakssd-b200")
    subprocess.run([cli, "shuffle", "-k", "11", "-s", "6", "-l", "3", "-o", str(tmp_path / "L3K11"), "--seed", "5"],
                   check=True, capture_output=True)
    S = oracle.synth(3, 8, 100_000, 150)
    one = bytes(S.fastq(0, 100_000))
    fq = tmp_path / "big.fq"
    with open(fq, "wb") as f:
        for _ in range(6):
            f.write(one)
    env = dict(os.environ, MK_INGEST_CHUNK_BYTES=str(16 << 20))
    # a fresh parent process, so that ru_maxrss of its children is this one command's high-water mark
    code = ("import resource, subprocess, sys; r = subprocess.run(sys.argv[1:], capture_output=True); "
            "print(r.returncode, resource.getrusage(resource.RUSAGE_CHILDREN).ru_maxrss)")
    import sys
    r = subprocess.run([sys.executable, "-c", code, cli, "dist", "-L", str(tmp_path / "L3K11.shuf"), "-A", "-o",
                        str(tmp_path / "sk"), str(fq)], capture_output=True, text=True, env=env, timeout=600)
    rc, rss_kb = [int(x) for x in r.stdout.split()]
    assert rc == 0, r.stderr[-2000:]
    # CUDA context + library image + the pinned ring (4 x 16 MB here); the text itself (190 MB) is never resident
    assert rss_kb < 1_000_000, rss_kb
