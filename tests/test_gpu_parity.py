"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.
Bar: bit-exact codes, counts and on-disk order (all integer work)."""
import numpy as np
import pytest

from helpers import markerdb_from_sketches, same_sketch

pytestmark = pytest.mark.gpu

L3K11 = (11, 6, 3)


@pytest.fixture(scope="module")
def sk311(oracle, lib_built, shuf):
    sid, perm = shuf(1234, *L3K11)
    s = lib_built.Sketcher(perm, *L3K11)
    yield s, perm, oracle.params(*L3K11)
    s.close()


@pytest.mark.parametrize("k,subk,L,nreads", [(11, 6, 3, 20000), (11, 5, 2, 20000), (10, 6, 3, 8000),
                                             (9, 4, 1, 3000), (12, 6, 3, 6000), (13, 6, 4, 6000)])
def test_fastq_koc_matches_oracle(oracle, lib_built, shuf, k, subk, L, nreads):
    sid, perm = shuf(1234 + k * 100 + subk, k, subk, L)
    p = oracle.params(k, subk, L)
    S = oracle.synth(42, 20, 200000, 150)
    fq = S.fastq(0, nreads)
    want = oracle.fastq_koc(p, perm, fq)
    assert want.status == 0 and want.codes.size > 0
    with lib_built.Sketcher(perm, k, subk, L) as sk:
        same_sketch(sk.fastq_koc_host(fq), want, p)


@pytest.mark.parametrize("tile", [64, 192, 1024, 4096, 32768])
def test_tile_boundaries(oracle, sk311, monkeypatch, tile):
    """Small tiles force reads, k-mers and records to straddle tile boundaries and exercise the
    decoupled look-back across thousands of tiles."""
    s, perm, p = sk311
    S = oracle.synth(7, 5, 50000, 150)
    fq = S.fastq(100, 2100)
    want = oracle.fastq_koc(p, perm, fq)
    monkeypatch.setenv("MK_TILE_BYTES", str(tile))
    same_sketch(s.fastq_koc_host(fq), want, p)


def _mutations(fq: bytes):
    recs = fq.split(b"\n")
    lines = [l for l in recs]
    yield "as_is", fq
    yield "no_final_newline", fq[:-1]
    yield "missing_quality", b"\n".join(lines[:-2]) + b"\n"               # last record: 3 lines
    yield "plus_without_newline", b"\n".join(lines[:-2])                   # ... "+" then EOF
    yield "seq_only_tail", b"\n".join(lines[:-3]) + b"\n"                  # last record: 2 lines
    yield "seq_unterminated", b"\n".join(lines[:-3])
    yield "header_only_tail", b"\n".join(lines[:-4]) + b"\n"
    yield "lowercase", fq.lower().replace(b"@r", b"@R")
    yield "crlf", fq.replace(b"\n", b"\r\n")
    yield "blank_line_shift", fq[:5000] + b"\n" + fq[5000:]                # shifts the record phase
    yield "long_headers", fq.replace(b"@r", b"@" + b"ACGT" * 60 + b" read/")
    yield "empty", b""
    yield "one_newline", b"\n"
    yield "tiny_reads", b"".join(b"@x\nACGTACGTAC\n+\nIIIIIIIIII\n" for _ in range(300))


def test_fastq_edge_cases(oracle, sk311):
    s, perm, p = sk311
    S = oracle.synth(11, 5, 50000, 150)
    fq = bytes(S.fastq(0, 1500))
    for name, text in _mutations(fq):
        want = oracle.fastq_koc(p, perm, text)
        got = s.fastq_koc_host(np.frombuffer(text, dtype=np.uint8).copy() if text else np.empty(0, np.uint8))
        try:
            same_sketch(got, want, p)
        except AssertionError as e:
            raise AssertionError("%s: %s" % (name, e))


def test_variable_read_lengths_and_ns(oracle, sk311):
    s, perm, p = sk311
    rng = np.random.default_rng(5)
    S = oracle.synth(3, 4, 30000, 150)
    base = bytes(S.fastq(0, 800)).split(b"\n")
    out = []
    for i in range(0, len(base) - 1, 4):
        seq = bytearray(base[i + 1])
        n = int(rng.integers(1, len(seq) + 1))
        seq = seq[:n]
        for _ in range(int(rng.integers(0, 3))):
            seq[int(rng.integers(0, n))] = ord("N")
        out += [base[i], bytes(seq), b"+", b"I" * n]
    text = b"\n".join(out) + b"\n"
    same_sketch(s.fastq_koc_host(np.frombuffer(text, np.uint8).copy()), oracle.fastq_koc(p, perm, text), p)


def test_count_saturation(oracle, sk311):
    """More than 65535 occurrences of a k-mer saturate at 0xFFFF (iseq2comem.c:712)."""
    s, perm, p = sk311
    S = oracle.synth(21, 3, 20000, 150)
    recs = bytes(S.fastq(0, 400)).split(b"\n")
    one = None
    for i in range(0, len(recs) - 1, 4):       # a read that contributes at least one code
        r = b"\n".join(recs[i:i + 4]) + b"\n"
        if oracle.fastq_koc(p, perm, r).codes.size:
            one = r
            break
    assert one is not None
    text = one * 66000 + bytes(S.fastq(400, 600))
    want = oracle.fastq_koc(p, perm, text)
    assert want.counts.max() == 65535
    same_sketch(s.fastq_koc_host(np.frombuffer(text, np.uint8).copy()), want, p)


def test_crowded_is_an_error(oracle, lib_built, shuf):
    """distinct codes > hashlimit is fatal in the reference (iseq2comem.c:708-709)."""
    k, subk, L = 7, 6, 3
    sid, perm = shuf(99, k, subk, L)
    p = oracle.params(k, subk, L)
    assert p.hashlimit == 305
    S = oracle.synth(8, 100, 200000, 150)
    fq = S.fastq(0, 60000)
    assert oracle.fastq_koc(p, perm, fq).status == 1
    with lib_built.Sketcher(perm, k, subk, L) as sk:
        with pytest.raises(lib_built.MkError) as ei:
            sk.fastq_koc_host(fq)
        assert ei.value.code == -5


def test_device_resident_and_file_inputs(oracle, sk311, tmp_path):
    import torch
    s, perm, p = sk311
    S = oracle.synth(13, 6, 60000, 150)
    fq = S.fastq(0, 5000)
    want = oracle.fastq_koc(p, perm, fq)
    d = torch.from_numpy(fq).cuda()
    same_sketch(s.fastq_koc_device(d, d.numel()), want, p)
    path = tmp_path / "reads.fq"
    fq.tofile(path)
    same_sketch(s.fastq_koc_file(str(path)), want, p)
    import gzip
    gz = tmp_path / "reads2.fq.gz"
    with gzip.open(gz, "wb") as f:
        f.write(bytes(fq))
    same_sketch(s.fastq_koc_file(str(gz)), want, p)


def test_device_generator_matches_host_generator(oracle, sk311, lib_built):
    import torch
    s, perm, p = sk311
    S = oracle.synth(77, 12, 100000, 150)
    r0, r1 = 95, 10250   # crosses digit-count boundaries of the record header
    want = S.fastq(r0, r1)
    P = lib_built.MksParams.from_buffer_copy(bytes(S.P))
    d = torch.empty(want.size + 64, dtype=torch.uint8, device="cuda")
    n = s.synth_fastq_device(P, S.cdf32, S.species, r0, r1, d, d.numel())
    assert n == want.size
    assert np.array_equal(d[:n].cpu().numpy(), want)
    fa = np.concatenate([S.fasta(i) for i in range(3, 6)])
    d2 = torch.empty(fa.size + 64, dtype=torch.uint8, device="cuda")
    off = s.synth_fasta_device(P, 3, 6, d2, d2.numel())
    assert int(off[-1]) == fa.size
    assert np.array_equal(d2[:fa.size].cpu().numpy(), fa)


# ------------------------------------------------------------------------------------ FASTA
def _fasta_cases(S):
    g = [bytes(S.fasta(i)) for i in range(4)]
    yield "plain", g[0]
    yield "multi_record", g[0] + g[1]
    yield "lowercase_mix", g[2][:2000] + g[2][2000:9000].lower() + g[2][9000:]
    yield "n_runs", g[3][:5000] + b"NNNNNNNNNN\n" + g[3][5000:12000] + b"RYK" + g[3][12000:]
    yield "crlf", g[0].replace(b"\n", b"\r\n")
    yield "gt_mid_line", g[1][:3000] + b">oops a header in the middle ACGTACGTACGTACGTACGTACGTACGT\n" + g[1][3000:]
    yield "header_at_eof", g[0] + b">trailing header without newline ACGT"
    yield "no_header", g[0].split(b"\n", 1)[1]
    yield "single_line", g[0].split(b"\n", 1)[0] + b"\n" + g[0].split(b"\n", 1)[1].replace(b"\n", b"") + b"\n"
    yield "digits_and_gaps", g[2][:4000] + b"12 -*\n" + g[2][4000:]
    yield "empty", b""
    yield "only_header", b">nothing here\n"


@pytest.mark.parametrize("k,subk,L", [(11, 6, 3), (11, 5, 2)])
def test_fasta_matches_oracle(oracle, lib_built, shuf, k, subk, L):
    sid, perm = shuf(1234 + k * 100 + subk, k, subk, L)
    p = oracle.params(k, subk, L)
    S = oracle.synth(5, 8, 150000, 150)
    cases = list(_fasta_cases(S))
    with lib_built.Sketcher(perm, k, subk, L) as sk:
        # one by one ...
        for name, text in cases:
            want = oracle.fasta_co(p, perm, text)
            got = sk.fasta_co_host([text])[0]
            try:
                same_sketch(got, want, p)
            except AssertionError as e:
                raise AssertionError("%s: %s" % (name, e))
        # ... and as one batch (files must not leak k-mers or header state into each other)
        got = sk.fasta_co_host([t for _, t in cases])
        for (name, text), g in zip(cases, got):
            try:
                same_sketch(g, oracle.fasta_co(p, perm, text), p)
            except AssertionError as e:
                raise AssertionError("batch/%s: %s" % (name, e))


# ------------------------------------------------------------------------------------ composite
def test_composite_matches_oracle(oracle, sk311):
    s, perm, p = sk311
    S = oracle.synth(1001, 40, 300000, 150)
    sp = [oracle.fasta_co(p, perm, S.fasta(i)).components(p)[0][0] for i in range(40)]
    ref_codes, ref_index = markerdb_from_sketches(sp)
    names = ["%d_sp%d" % (i + 1, i) for i in range(40)]
    q = oracle.fastq_koc(p, perm, S.fastq(0, 60000))
    qc, qa = q.components(p)[0]
    want = oracle.composite([(ref_codes, ref_index)], names, [(qc, qa)], "reads.fq")
    assert want.count("\n") >= 3
    from metakssd_b200 import composite_tsv
    stats, lists = s.composite([(ref_codes, ref_index)], [(qc, qa)], want_lists=True)
    assert composite_tsv("reads.fq", names, stats) == want
    for i, l in enumerate(lists):
        assert l[0] == stats["n"][i] == l.size - 1
    # end to end: GPU sketch -> GPU composite
    g = s.fastq_koc_host(S.fastq(0, 60000))
    stats2 = s.composite([(ref_codes, ref_index)], [(g.codes[0], g.counts[0])])
    assert composite_tsv("reads.fq", names, stats2) == want
    # resident MarkerDB: same answer, repeatedly, without re-uploading the database
    s.load_markerdb([(ref_codes, ref_index)])
    for _ in range(2):
        assert composite_tsv("reads.fq", names, s.composite(None, [(g.codes[0], g.counts[0])])) == want
    # ... and with the query taken from the device: the sketch this context produced last
    from metakssd_b200 import SpeciesNames, coverage_tsv
    g2 = s.fastq_koc_host(S.fastq(0, 60000))
    assert np.array_equal(g2.codes[0], g.codes[0])
    assert coverage_tsv("reads.fq", SpeciesNames(names), s.composite_last()) == want


def test_composite_multi_component(oracle, lib_built, shuf):
    k, subk, L = 11, 5, 2
    sid, perm = shuf(1234 + k * 100 + subk, k, subk, L)
    p = oracle.params(k, subk, L)
    assert p.component_num == 16
    S = oracle.synth(2002, 12, 120000, 150)
    from metakssd_b200 import composite_tsv
    sp = [oracle.fasta_co(p, perm, S.fasta(i)).components(p) for i in range(12)]
    ref = [markerdb_from_sketches([sp[i][c][0] for i in range(12)]) for c in range(16)]
    names = ["%d_sp%d" % (i + 1, i) for i in range(12)]
    q = oracle.fastq_koc(p, perm, S.fastq(0, 20000)).components(p)
    want = oracle.composite(ref, names, q, "q.fq")
    with lib_built.Sketcher(perm, k, subk, L) as sk:
        stats = sk.composite(ref, q)
        sk.load_markerdb(ref)
        stats_res = sk.composite(None, q)
    assert composite_tsv("q.fq", names, stats) == want and want
    assert composite_tsv("q.fq", names, stats_res) == want


def test_empty_and_one_code_query_components(sk311, lib_built, oracle):
    """command_composite.c:535: a query component with NO code gets a dictionary of 0 slots — both loops run zero
    times, no hit, no error (a 16-component sketch of a shallow sample has such components); with exactly ONE code
    the table has 1 slot and HASH() divides by zero: the reference dies, the library reports MK_ERR_EMPTY_QUERY."""
    s, perm, p = sk311
    ref = (np.arange(100, dtype=np.uint32), np.array([0, 50, 100], dtype=np.uint64))
    stats = s.composite([ref], [(np.empty(0, np.uint32), np.empty(0, np.uint16))])
    assert int(stats["n"].sum()) == 0
    assert oracle.composite([ref], ["a", "b"], [(np.empty(0, np.uint32), np.empty(0, np.uint16))], "q") == ""
    # an empty component next to a populated one: the populated one still counts
    qc = np.arange(10, 60, dtype=np.uint32)
    stats = s.composite([ref, ref], [(np.empty(0, np.uint32), np.empty(0, np.uint16)), (qc, np.full(50, 3, np.uint16))])
    assert stats["n"].tolist() == [40, 10]
    with pytest.raises(lib_built.MkError) as ei:
        s.composite([ref], [(np.array([5], np.uint32), np.array([1], np.uint16))])
    assert ei.value.code == -8
    with pytest.raises(ValueError):
        oracle.composite([ref], ["a", "b"], [(np.array([5], np.uint32), np.array([1], np.uint16))], "q")


def test_host_upload_pipeline_chunks(oracle, sk311, monkeypatch):
    """mk_fastq_koc_host uploads and sketches chunk by chunk (one launch per chunk, line count and
    candidates carried over): many small chunks must give the sketch of the whole text."""
    s, perm, p = sk311
    S = oracle.synth(77, 12, 200000, 150)
    text = S.fastq(0, 25000)
    want = oracle.fastq_koc(p, perm, text)
    for chunk in ("60000", "1000000", "49152"):
        monkeypatch.setenv("MK_CHUNK_BYTES", chunk)
        same_sketch(s.fastq_koc_host(text), want, p)
    monkeypatch.delenv("MK_CHUNK_BYTES")
    same_sketch(s.fastq_koc_host(text), want, p)


def test_long_lines_exact_threshold(oracle, sk311):
    """fgets(buf, 4096) splits a line once 4095 bytes came without a newline: such input is refused
    (MK_ERR_LONG_LINE) exactly from that length on, wherever the line sits; shorter lines are sketched."""
    import metakssd_b200 as M
    s, perm, p = sk311
    S = oracle.synth(91, 6, 100000, 150)
    genome = bytes(S.fasta(0)).split(b"\n", 1)[1].replace(b"\n", b"")
    recs = bytes(S.fastq(0, 3000))

    def with_line(seq_len, where):
        seq = genome[:seq_len]
        rec = b"@long\n" + seq + b"\n+\n" + b"I" * seq_len + b"\n"
        cut = recs.find(b"\n@r", where) + 1
        return recs[:cut] + rec + recs[cut:]

    for seq_len, where in ((3000, 100), (4094, 50000), (4094, 400000)):
        text = with_line(seq_len, where)
        want = oracle.fastq_koc(p, perm, np.frombuffer(text, np.uint8))
        same_sketch(s.fastq_koc_host(np.frombuffer(text, np.uint8).copy()), want, p)
    for seq_len, where in ((4095, 100), (4095, 300000), (5000, 200000), (13000, 7), (40000, 123456)):
        text = with_line(seq_len, where)
        with pytest.raises(M.MkError) as e:
            s.fastq_koc_host(np.frombuffer(text, np.uint8).copy())
        assert e.value.code == -6, (seq_len, where, e.value)   # MK_ERR_LONG_LINE


def test_classic_stream_kernel_agrees(oracle, sk311, monkeypatch):
    """MK_STREAM_IMPL=classic selects the earlier unit-pulling kernel (kept as a cross-check of the
    warp-specialised one): same sketch, from device and from host text."""
    s, perm, p = sk311
    S = oracle.synth(31, 15, 150000, 150)
    text = S.fastq(0, 30000)
    want = oracle.fastq_koc(p, perm, text)
    monkeypatch.setenv("MK_STREAM_IMPL", "classic")
    same_sketch(s.fastq_koc_host(text), want, p)
    monkeypatch.delenv("MK_STREAM_IMPL")
    same_sketch(s.fastq_koc_host(text), want, p)



def test_fallback_when_the_shared_memory_base_check_fails(lib_built, oracle, shuf, capfd):
    """k_stream_ws reads the filter through a fixed CTA-shared address and checks the assumption in every launch; when
    the check fails (forced here) the context switches to the unit-pulling kernel for good and says so once."""
    import os
    sid, perm = shuf(1234, 11, 6, 3)
    p = oracle.params(11, 6, 3)
    text = oracle.synth(5, 6, 100_000, 150).fastq(0, 20_000)
    want = oracle.fastq_koc(p, perm, text)
    os.environ["MK_DEBUG_FORCE_SMEM_BASE_FLAG"] = "1"
    try:
        with lib_built.Sketcher(perm, 11, 6, 3) as sk:
            got = sk.fastq_koc_host(np.asarray(text))
            again = sk.fastq_koc_host(np.asarray(text))
    finally:
        del os.environ["MK_DEBUG_FORCE_SMEM_BASE_FLAG"]
    err = capfd.readouterr().err
    assert err.count("using the unit-pulling stream kernel") == 1
    for g in (got, again):
        comps = want.components(p)
        assert np.array_equal(g.codes[0], comps[0][0]) and np.array_equal(g.counts[0], comps[0][1])
