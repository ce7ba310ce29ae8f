"""CPU: host-side mirror of the reference's formats and reporting (no GPU compute)."""
import numpy as np
import pytest


def test_generators_agree(oracle, lib_built):
    """Product generator helpers == oracle generator (same header, two libraries)."""
    a = lib_built.synth_spec(42, 100, 1_000_000)
    b = oracle.synth(42, 100, 1_000_000)
    assert bytes(a.P) == bytes(b.P) and np.array_equal(a.cdf32, b.cdf32) and np.array_equal(a.species, b.species)
    assert a.fastq_bytes(95, 10250) == b.fastq(95, 10250).size
    assert a.fasta_bytes(7) == b.fasta(7).size
    sid, perm = lib_built.make_shuf(77, 5)
    sid2, perm2 = oracle.make_shuf(77, 11, 5, 2)
    assert sid == sid2 and np.array_equal(perm, perm2)
    assert np.array_equal(np.sort(perm), np.arange(1 << 20))


def test_fastq_text_shape(oracle):
    S = oracle.synth(3, 10, 100_000)
    t = bytes(S.fastq(8, 12))
    lines = t.split(b"\n")
    assert lines[0] == b"@r8" and lines[8] == b"@r10" and lines[2] == b"+" and lines[3] == b"I" * 150
    assert all(len(lines[i]) == 150 for i in (1, 5, 9, 13)) and t.endswith(b"\n")


def test_shuf_file_round_trip(lib_built, tmp_path):
    sid, perm = lib_built.make_shuf(5, 4)
    p = tmp_path / "x.shuf"
    lib_built.write_shuf(str(p), sid, 9, 4, 1, perm)
    assert p.stat().st_size == 16 + 4 * (1 << 16)
    sid2, k, subk, L, perm2 = lib_built.read_shuf(str(p))
    assert (sid2, k, subk, L) == (sid, 9, 4, 1) and np.array_equal(perm, perm2)


def test_sketch_dir_round_trip(lib_built, oracle, tmp_path):
    """write_sketch_dir() produces what run_stageI() leaves on disk; the oracle-side reader (which
    parses the reference's own directories in make_golden.py) reads it back."""
    info = lib_built.MkInfo()
    info.k, info.drlevel, info.component_num = 11, 2, 16
    rng = np.random.default_rng(1)
    sk = []
    for _ in range(3):
        codes = [rng.integers(0, 1 << 32, size=int(rng.integers(0, 50)), dtype=np.uint32) for _ in range(16)]
        counts = [rng.integers(1, 65535, size=c.size, dtype=np.uint16) for c in codes]
        sk.append(lib_built.Sketch(codes, counts))
    names = ["a.fq", "dir/b.fastq", "c" * 300]
    d = str(tmp_path / "sk")
    lib_built.write_sketch_dir(d, 0x12345678, info, names, sk, koc=True)
    sd = oracle.read_sketch_dir(d)
    assert (sd.shuf_id, sd.koc, sd.kmerlen, sd.dim_rd_len, sd.comp_num, sd.infile_num) == (0x12345678, True, 22, 4, 16, 3)
    assert sd.all_ctx_ct == sum(s.n_total for s in sk) and list(sd.ctx_ct) == [s.n_total for s in sk]
    assert sd.names[:2] == names[:2] and sd.names[2] == "c" * 255
    for c in range(16):
        assert np.array_equal(sd.combco[c], np.concatenate([s.codes[c] for s in sk]))
        assert np.array_equal(sd.abund[c], np.concatenate([s.counts[c] for s in sk]))
        assert list(sd.index[c]) == [0] + list(np.cumsum([s.codes[c].size for s in sk]))
    hdr, names2, combco, index, abund = lib_built.read_sketch_dir(d)
    assert hdr["comp_num"] == 16 and names2 == sd.names and np.array_equal(combco[3], sd.combco[3])


def _stats_numpy(lists):
    """Restatement of command_composite.c:598-613 on plain lists (what mk_composite_stats computes)."""
    out = np.zeros(len(lists), dtype=[("n", "<i4"), ("sum", "<i4"), ("lastsum", "<i4"), ("lastn", "<i4"),
                                      ("median", "<i4"), ("max", "<i4")])
    for s, v in enumerate(lists):
        a = [len(v)] + sorted(int(x) for x in v)
        n = len(v)
        out[s]["n"] = n
        out[s]["sum"] = sum(a[1:])
        j, ls, ln = int(n * 0.98), 0, 0
        while j <= n * 0.99:
            ls += a[j]; ln += 1; j += 1
        out[s]["lastsum"], out[s]["lastn"] = ls, ln
        out[s]["median"] = a[n // 2] if n else 0
        out[s]["max"] = a[n] if n else 0
    return out


def test_composite_tsv_matches_reference_printf(oracle, lib_built):
    rng = np.random.default_rng(7)
    S = 60
    sizes = rng.integers(0, 400, size=S)
    sizes[:5] = [0, 5, 6, 6, 1]                       # MIN_KM_S boundary and ties
    ref_index = np.zeros(S + 1, dtype=np.uint64)
    ref_index[1:] = np.cumsum(sizes)
    ref_codes = rng.permutation(int(ref_index[-1]) * 3)[: int(ref_index[-1])].astype(np.uint32)
    hit = rng.random(ref_codes.size) < 0.8
    hit[: int(ref_index[5])] = True
    qry_codes = np.concatenate([ref_codes[hit], np.arange(10**6, 10**6 + 500, dtype=np.uint32)])
    qry_counts = rng.integers(1, 3000, size=qry_codes.size).astype(np.uint16)
    perm = rng.permutation(qry_codes.size)
    qry_codes, qry_counts = qry_codes[perm], qry_counts[perm]
    names = ["%d_species%d" % (i + 1, i) for i in range(S)]
    want = oracle.composite([(ref_codes, ref_index)], names, [(qry_codes, qry_counts)], "sample.fq")
    lut = dict(zip(qry_codes.tolist(), qry_counts.tolist()))
    lists = [[lut[c] for c in ref_codes[int(ref_index[s]):int(ref_index[s + 1])].tolist() if c in lut] for s in range(S)]
    got = lib_built.composite_tsv("sample.fq", names, _stats_numpy(lists))
    assert got == want and want.count("\n") > 40


def test_organize_taxa_matches_oracle(lib_built, oracle):
    """workload.organize_taxa (host logic of the MarkerDB pipeline) lists the taxa like organize_taxf()
    (command_set.c:635-704), here against the oracle's restatement (itself pinned to the reference binary)"""
    from metakssd_b200.workload import organize_taxa
    rng = np.random.default_rng(3)
    for n in (2, 7, 100, 1000):          # (a one-line taxfile divides by zero in the reference: hash size 1)
        taxids = rng.integers(1, 3_000_000, size=n).tolist()
        taxids += taxids[: n // 3]                       # several genomes per taxon
        a, ids_a = organize_taxa(taxids)
        b, ids_b = oracle.organize_taxa(taxids)
        assert ids_a == ids_b and np.array_equal(a, b)
    a, ids = organize_taxa([s + 1 for s in range(1000)])
    assert sorted(ids) == list(range(1, 1001))


def test_c_coverage_formatter_matches_python(lib_built):
    """mk_format_species_coverage (host C, command_composite.c:582-624) against the Python statement of
    the same rule: order by matches descending with ties in index order, cut below 6, float32 ratios."""
    from metakssd_b200 import SpeciesNames, composite_tsv, coverage_tsv
    from metakssd_b200.api import STATS_DTYPE
    rng = np.random.default_rng(5)
    S = 400
    st = np.zeros(S, dtype=STATS_DTYPE)
    st["n"] = rng.choice([0, 3, 5, 6, 6, 7, 50, 50, 1000, 123456], size=S)
    st["sum"] = rng.integers(1, 2 ** 31 - 1, size=S)
    st["sum"][::7] = -5                      # int wrap-around of the reference's `int sum`
    st["lastsum"] = rng.integers(0, 70000, size=S)
    st["lastn"] = np.maximum(1, st["n"] // 100 + 1)
    st["median"] = rng.integers(0, 65536, size=S)
    st["max"] = rng.integers(0, 65536, size=S)
    names = ["%d_species_%d" % (i + 1, i) for i in range(S)]
    want = composite_tsv("some/query.fq", names, st)
    got = coverage_tsv("some/query.fq", SpeciesNames(names), st)
    assert got == want and want.count("\n") > 100
    assert coverage_tsv("q", SpeciesNames(names), np.zeros(S, dtype=STATS_DTYPE)) == ""



def test_code_range_edges_balance_the_owners(oracle):
    """distributed.code_range_edges / range_edge() in csrc/mk_comm.cu: quantiles of the min-of-two-uniforms law the
    codes follow, so every owner gets about the same share of a sketch (equal-width ranges gave the first of two
    owners two thirds)."""
    from metakssd_b200 import distributed as D
    p = oracle.params(11, 6, 3)
    _, perm = oracle.make_shuf(5, 11, 6, 3)
    S = oracle.synth(7, 20, 200_000, 150)
    codes = np.asarray(oracle.fastq_koc(p, perm, S.fastq(0, 400_000)).codes, dtype=np.int64)
    assert codes.size > 1000
    for world in (2, 4, 8):
        e = D.code_range_edges(world, 32).numpy()
        assert e[0] == 0 and e[-1] == 1 << 32 and np.all(np.diff(e) > 0)
        share = np.histogram(codes, bins=e)[0] / codes.size
        assert share.max() < 1.35 / world and share.min() > 0.65 / world, (world, share)
    even = np.histogram(codes, bins=[0, 1 << 31, 1 << 32])[0] / codes.size
    assert even[0] > 0.6                                   # what equal-width ranges would have done


def test_balanced_shares_keep_the_total_and_equalise_the_ranks():
    """distributed.balanced_shares: rank 0 (which carries the tail of the sharded step) gets fewer reads; with the
    shares it returns, stream time + tail of rank 0 equals the stream time of the others."""
    from metakssd_b200 import distributed as D
    per, ms_per_read = 40_000_000, 8.0 / 40_000_000
    for world, tail in ((2, 0.9), (4, 1.2), (8, 1.5)):
        waits = [0.1] + [0.1 + tail] * (world - 1)
        shares, moved = D.balanced_shares(per, world, waits, ms_per_read)
        assert sum(shares) == world * per and shares[0] == per - moved and min(shares) > 0
        assert abs((shares[0] * ms_per_read + tail) - shares[1] * ms_per_read) < 1e-3
        again, moved2 = D.balanced_shares(per, world, [0.1] * world, ms_per_read, moved)      # nothing left to correct
        assert again == shares and moved2 == moved
        back, moved3 = D.balanced_shares(per, world, [0.1 + tail] + [0.1] * (world - 1), ms_per_read, moved)   # over-corrected
        assert moved3 < moved and sum(back) == world * per
    assert D.balanced_shares(per, 1, [0.0], ms_per_read) == ([per], 0)
    shares, moved = D.balanced_shares(per, 4, [0.0, 100.0, 100.0, 100.0], ms_per_read)        # capped at 40 % of a shard
    assert moved == per * 2 // 5 and sum(shares) == 4 * per


def test_host_reads_the_reference_mco_format(tmp_path, lib_built):
    """host/mkssd_main.c::read_mco_component: a database directory in the reference's format (mcofiles.stat, mco.<c> =
    the references holding each code, mco.index.<c> = END offset of every code's row, co2mco.c:56-79) comes back as
    the sketches it was built from (`dist -r <db> --dump-ref <dir>`, no device needed).  The real index has 2^32 rows
    (32 GiB); the reader takes as many rows as the file holds, so the test uses codes below 2^16."""
    import os
    import struct
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-C", os.path.join(root, "host")], check=True, capture_output=True)
    rng = np.random.default_rng(8)
    n_ref, rows = 9, 1 << 16
    sks = [np.unique(rng.integers(0, rows, size=int(rng.integers(0, 900)))).astype(np.uint32) for _ in range(n_ref)]
    sks[3] = np.zeros(0, np.uint32)                                      # an empty sketch
    db, qd = tmp_path / "db", tmp_path / "qry"
    db.mkdir(); qd.mkdir()
    names = [b"genome_%d.fa" % i for i in range(n_ref)]
    stat_tail = np.array([s.size for s in sks], dtype=np.uint32).tobytes() + b"".join(n.ljust(256, b"\0") for n in names)
    (db / "mcofiles.stat").write_bytes(struct.pack("<Iiiii", 77, 22, 6, 1, n_ref) + stat_tail)
    holders = [[] for _ in range(rows)]                                  # the inverted index, rows in code order
    for g, s in enumerate(sks):
        for c in s:
            holders[int(c)].append(g)
    np.array([g for h in holders for g in h], dtype=np.uint32).tofile(db / "mco.0")
    np.cumsum([len(h) for h in holders]).astype(np.uint64).tofile(db / "mco.index.0")
    (qd / "cofiles.stat").write_bytes(struct.pack("<I?3xiiiiQ", 77, False, 22, 6, 1, 1, 0) + struct.pack("<I", 0) + b"q".ljust(256, b"\0"))
    np.zeros(0, np.uint32).tofile(qd / "combco.0")
    np.zeros(2, np.uint64).tofile(qd / "combco.index.0")
    dump = tmp_path / "dump"
    r = subprocess.run([os.path.join(root, "host", "metakssd-b200"), "dist", "-r", str(db), "-o", str(tmp_path / "o"),
                        "--dump-ref", str(dump), str(qd)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    hdr, got_names, combco, index, _ = lib_built.read_sketch_dir(str(dump))
    assert got_names == [n.decode() for n in names] and hdr["shuf_id"] == 77 and hdr["infile_num"] == n_ref
    for g, s in enumerate(sks):
        assert np.array_equal(combco[0][int(index[0][g]):int(index[0][g + 1])], s), g
    # a truncated gid file is refused, not read past
    (db / "mco.0").write_bytes((db / "mco.0").read_bytes()[:-8])
    r = subprocess.run([os.path.join(root, "host", "metakssd-b200"), "dist", "-r", str(db), "-o", str(tmp_path / "o"),
                        "--dump-ref", str(dump), str(qd)], capture_output=True, text=True)
    assert r.returncode != 0 and "mco" in r.stderr


def test_host_dist_expands_directories_and_list_files(tmp_path):
    """`dist` takes its inputs like the reference (organize_infile_frm_arg / organize_infile_list, global_basic.c:169-330):
    a directory stands for the sequence files in it (accepted extensions, optionally compressed), -l names one file per
    line; `--list-inputs` prints the expansion without touching a device.  The query of `dist -r` is left alone."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-C", os.path.join(root, "host")], check=True, capture_output=True)
    cli = os.path.join(root, "host", "metakssd-b200")
    g = tmp_path / "genomes"
    g.mkdir()
    for n in ("b.fasta", "a.fq.gz", "notes.txt", "d.FA", "e.fna.bz2", "sub"):
        (g / n).mkdir() if n == "sub" else (g / n).write_bytes(b"")
    lst = tmp_path / "list.txt"
    lst.write_text("  %s\n\n%s\r\n" % (g / "b.fasta", g / "a.fq.gz"))
    r = subprocess.run([cli, "dist", "--list-inputs", str(g), str(g / "d.FA"), "-l", str(lst)], capture_output=True, text=True, check=True)
    want = [str(g / n) for n in ("a.fq.gz", "b.fasta", "d.FA", "e.fna.bz2")] + [str(g / "d.FA"), str(g / "b.fasta"), str(g / "a.fq.gz")]
    assert r.stdout.split("\n")[:-1] == want
    r = subprocess.run([cli, "dist", "--list-inputs", str(g), "-r", "refdir", "-o", "out"], capture_output=True, text=True, check=True)
    assert r.stdout == str(g) + "\n"
    bad = tmp_path / "bad.txt"
    bad.write_text(str(g / "notes.txt") + "\n")
    r = subprocess.run([cli, "dist", "--list-inputs", "-l", str(bad)], capture_output=True, text=True)
    assert r.returncode != 0 and "wrong format" in r.stderr


def test_host_prints_the_reference_distance_table(tmp_path, lib_built, oracle):
    """host/mkssd_main.c::dist_search printing (dist_print_nobin / output_ctrl, command_dist.c:1531-1680) on CPU: with
    `-f <sharedk_ct.dat>` the table is printed from a kept count matrix (command_dist.c:984-987), so the golden matrix of
    the reference binary must give the reference's distance.out for all eight option sets, byte for byte."""
    import os
    import subprocess
    import golden_cases as G
    from helpers import dist_search_world
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-C", os.path.join(root, "host")], check=True, capture_output=True)
    cli = os.path.join(root, "host", "metakssd-b200")
    gold = np.load(os.path.join(root, "tests", "golden", "reference_vectors_r2b.npz"))
    p, perm, ref_names, ref, ref_ct, qry_names, qry, qry_ct = dist_search_world(oracle, gold)
    sid = oracle.make_shuf(1234, 11, 6, 3)[0]
    info = lib_built.MkInfo()
    info.k, info.drlevel, info.component_num = 11, 3, 1

    def sketches(side):
        codes, idx = side
        return [lib_built.Sketch([codes[int(idx[i]):int(idx[i + 1])]], None) for i in range(idx.size - 1)]

    rd, qd = str(tmp_path / "ref"), str(tmp_path / "qry")
    lib_built.write_sketch_dir(rd, sid, info, ref_names, sketches(ref), False)
    lib_built.write_sketch_dir(qd, sid, info, qry_names, sketches(qry), False)
    skf = str(tmp_path / "sharedk_ct.dat")
    gold["sharedk_ct"].astype(np.uint32).tofile(skf)
    for name, flags in G.DIST_SEARCH_OPTIONS.items():
        out = str(tmp_path / ("out_" + name))
        subprocess.run([cli, "dist", "-r", rd, "-o", out, "-f", skf] + flags + [qd], check=True, capture_output=True, timeout=60)
        assert open(os.path.join(out, "distance.out")).read() == str(gold["out/" + name]), name
    open(skf, "ab").write(b"\0\0\0\0")
    r = subprocess.run([cli, "dist", "-r", rd, "-o", str(tmp_path / "o"), "-f", skf, qd], capture_output=True, text=True)
    assert r.returncode != 0 and "does not fit" in r.stderr
    # sketches without a code: the nan / inf lines of the reference
    from helpers import dist_search_edge_world
    p, perm, ref_names, ref, ref_ct, qry_names, qry, qry_ct = dist_search_edge_world(oracle, gold)
    rd, qd = str(tmp_path / "eref"), str(tmp_path / "eqry")
    lib_built.write_sketch_dir(rd, sid, info, ref_names, sketches(ref), False)
    lib_built.write_sketch_dir(qd, sid, info, qry_names, sketches(qry), False)
    gold["edge/sharedk_ct"].astype(np.uint32).tofile(skf)
    for name in ("default", "containment", "nearest3", "corrected"):
        out = str(tmp_path / ("eout_" + name))
        subprocess.run([cli, "dist", "-r", rd, "-o", out, "-f", skf] + G.DIST_SEARCH_OPTIONS[name] + [qd], check=True,
                       capture_output=True, timeout=60)
        assert open(os.path.join(out, "distance.out")).read() == str(gold["edge/out/" + name]), name
