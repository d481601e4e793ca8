"""BAM ingest: dropest_b200/host/BamIngest against BAM files written by tests/bam_utils.py.  The parsing tests are CPU
only; the end-to-end test (BAM -> container on the GPU -> count matrix vs the oracle on the same reads) needs a GPU."""
import os
import subprocess

import numpy as np
import pytest

import parity_utils as pu
from bam_utils import alignment, write_bam

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DUMP = os.path.join(ROOT, "dropest_b200", "lib", "test_bam_ingest")
REFS = [("chr1", 1000000), ("chr2", 900000), ("chrM", 16000)]


def _dump(files, filled=True, min_q=0, gene_in_chr=False, type_tag="-", intronic="-", intergenic="-", threads=3, expect_ok=True, n_columns=6, env=None):
    assert os.path.exists(DUMP), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    r = subprocess.run([DUMP, "1" if filled else "0", str(min_q), "1" if gene_in_chr else "0", type_tag, intronic, intergenic, str(threads)] + list(files),
                       capture_output=True, text=True, env=dict(os.environ, **(env or {})))
    assert (r.returncode == 0) == expect_ok, r.stdout[-500:] + r.stderr[-500:]
    lines = r.stdout.strip().split("\n")
    return [tuple(l.split("\t"))[:n_columns] for l in lines if not l.startswith("#")], [l.split("\t") for l in lines if l.startswith("#")]


def _random_reads(n, seed):
    rng = np.random.default_rng(seed)
    acgt = np.array(list("ACGT"))
    reads = []
    for i in range(n):
        cb = "".join(rng.choice(acgt, 16)) if rng.random() > 0.3 else "ACGTACGTACGTACGT"
        umi = "".join(rng.choice(acgt, 10))
        gene = f"G{int(rng.integers(0, 50))}" if rng.random() > 0.1 else None
        reads.append((cb, umi, gene, int(rng.integers(0, 3)), int(rng.integers(0, 900000))))
    return reads


def test_filled_bam_tags_filters_and_counters(tmp_path):
    """-f mode: CB / UB / GX tags; unmapped and secondary alignments, unknown reference ids, missing barcode tags and low base qualities
    are skipped and counted as BamController::process_alignment does; records straddle many BGZF blocks; several inflate threads."""
    reads = _random_reads(40000, 1)
    als, exp = [], []
    n_unmapped = n_noref = n_notag = n_lowq = 0
    for i, (cb, umi, gene, ref, pos) in enumerate(reads):
        tags = [("NH", ("i", 1)), ("CB", ("Z", cb)), ("UB", ("Z", umi)), ("CQ", ("Z", "I" * 16)), ("UQ", ("Z", "I" * 10)), ("xs", ("B", [1, 2, 3]))]
        if gene:
            tags.insert(1, ("GX", ("Z", gene)))
        flag = 0
        kind = i % 23
        if kind == 3:
            flag = 4; n_unmapped += 1
        elif kind == 5:
            flag = 0x100; n_unmapped += 1
        elif kind == 7:
            ref = 9; n_noref += 1
        elif kind == 11:
            tags = [t for t in tags if t[0] != "UB"]; n_notag += 1
        elif kind == 13:
            tags = [t if t[0] != "UQ" else ("UQ", ("Z", "IIII#IIIII")) for t in tags]; n_lowq += 1
        als.append(alignment(f"r{i}", ref, pos, flag, tags))
        if kind not in (3, 5, 7, 11, 13):
            exp.append((cb, umi, gene or "-", REFS[ref][0], "2" if gene else "1", "I" * 16))
    path = str(tmp_path / "a.bam")
    write_bam(path, REFS, als, block_bytes=30000)
    for threads in (1, 4):
        got, meta = _dump([path], min_q=20, threads=threads)
        assert got == exp
        total = len(reads) - n_unmapped - n_noref
        assert meta[-1] == ["#stats", str(total), str(n_noref + n_notag), str(n_lowq), str(n_unmapped)]
    got, meta = _dump([path], min_q=0)   # quality filter off: the low-quality reads come through
    assert len(got) == len(exp) + n_lowq and meta[-1][3] == "0"


def test_fast_inflate_equals_zlib():
    """host/FastInflate.h (the reader's own DEFLATE decoder) on ~3000 streams written by zlib at every level / strategy -- stored, fixed and
    dynamic blocks, codes long enough for subtables, sizes around every boundary --, refusal of wrong output sizes, and ~3000 damaged
    streams that must fail or decode without touching a byte outside the output, and 40 000 fuzz cases (valid streams with bit flips,
    overwritten bytes, truncation; random bytes; wrong sizes) in which whatever the decoder accepts zlib accepts too, with the same bytes
    (tests/cpp/test_fast_inflate.cpp; 300 000 such cases ran clean under ASan / UBSan)."""
    exe = os.path.join(ROOT, "dropest_b200", "lib", "test_fast_inflate")
    assert os.path.exists(exe), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok\t"), r.stdout[-500:] + r.stderr[-500:]


def test_fast_inflate_and_zlib_paths_read_the_same(tmp_path):
    """The reader with its own decoder and with zlib only (DGE_BAM_ZLIB_INFLATE) returns the same reads; BAMs deflated at level 0 (stored
    blocks), 1 and 9."""
    reads = _random_reads(12000, 9)
    als = [alignment(f"q{i}", ref, pos, 0, [("CB", ("Z", cb)), ("UB", ("Z", umi))] + ([("GX", ("Z", gene))] if gene else []))
           for i, (cb, umi, gene, ref, pos) in enumerate(reads)]
    for level in (0, 1, 9):
        path = str(tmp_path / f"l{level}.bam")
        write_bam(path, REFS, als, block_bytes=60000, level=level)
        ours, m1 = _dump([path], threads=2)
        zl, m2 = _dump([path], threads=2, env={"DGE_BAM_ZLIB_INFLATE": "1"})
        assert ours == zl and m1 == m2 and len(ours) == len(reads)


def test_records_across_chunk_boundaries_and_loader_errors(tmp_path):
    """The reader inflates the next chunk in the background while the current one is framed and parsed: a record cut by a chunk boundary is
    completed in the new chunk's headroom (or, when it is longer than the headroom, in a buffer of its own).  Tiny chunks and no headroom put
    a cut into nearly every chunk; the output must not depend on either.  What the loader throws (truncated file, corrupt block, CRC)
    reaches the caller."""
    reads = _random_reads(20000, 5)
    als = [alignment(f"read{i}", ref, pos, 0, [("CB", ("Z", cb)), ("UB", ("Z", umi))] + ([("GX", ("Z", gene))] if gene else []) +
                     ([("xl", ("Z", "x" * 70000))] if i % 4001 == 7 else []))   # a few records longer than a BGZF block
           for i, (cb, umi, gene, ref, pos) in enumerate(reads)]
    path = str(tmp_path / "a.bam")
    write_bam(path, REFS, als, block_bytes=9000)
    want, meta = _dump([path], threads=2)
    assert len(want) == len(reads)
    for chunk, headroom in ((20000, 1 << 20), (20000, 0), (1, 64), (300000, 100)):
        env = {"DGE_BAM_CHUNK_BYTES": str(chunk), "DGE_BAM_HEADROOM_BYTES": str(headroom)}
        for packed in (False, True):
            got, m = _dump([path], threads=3, env=dict(env, **({"DGE_BAM_PACKED": "1"} if packed else {})))
            assert got == want and m[-1] == meta[-1]
    raw = open(path, "rb").read()
    cut = str(tmp_path / "cut.bam")
    open(cut, "wb").write(raw[:len(raw) * 2 // 3])
    _, m = _dump([cut], env={"DGE_BAM_CHUNK_BYTES": "20000"}, expect_ok=False)
    assert m[-1][0] == "#error" and "truncated" in m[-1][1]
    bad = bytearray(raw)
    bad[len(raw) // 2] ^= 0x55
    flipped = str(tmp_path / "flipped.bam")
    open(flipped, "wb").write(bytes(bad))
    _, m = _dump([flipped], env={"DGE_BAM_CHUNK_BYTES": "20000"}, expect_ok=False)
    assert m[-1][0] == "#error"


def test_read_name_mode_read_types_and_several_files(tmp_path):
    """Without -f the barcode and UMI come from the read name ("...!CB#UMI", ReadParameters::parse_encoded_id); the read-type tag maps to
    intron / not-annotated / exon marks (ReadParamsParser::parse_read_type); files are read one after the other."""
    recs = [("x!AAAACCCCGGGGTTTT#ACGTACGTAC", [("GX", ("Z", "Gene1")), ("XF", ("Z", "INTRONIC"))], "4"),
            ("y!AAAACCCCGGGGTTTT#TTTTTTTTTT", [("GX", ("Z", "Gene1")), ("XF", ("Z", "CODING"))], "2"),
            ("z!CCCCCCCCGGGGTTTT#ACGTACGTAC", [("GX", ("Z", "Gene2")), ("XF", ("A", "I"))], "2"),
            ("w!CCCCCCCCGGGGTTTT#ACGTACGTAA", [("GX", ("Z", "Gene2")), ("XF", ("Z", "INTERGENIC"))], "1"),
            ("v!CCCCCCCCGGGGTTTT#ACGTACGTAG", [("XF", ("Z", "INTRONIC"))], "1"),          # no gene tag: not annotated, whatever the type
            ("no_codec_here", [("GX", ("Z", "Gene2"))], None)]
    a = [alignment(n, 0, 10, 0, t) for n, t, _ in recs[:3]]
    b = [alignment(n, 1, 10, 16, t) for n, t, _ in recs[3:]]
    pa, pb = str(tmp_path / "a.bam"), str(tmp_path / "b.bam")
    write_bam(pa, REFS, a)
    write_bam(pb, REFS, b)
    got, meta = _dump([pa, pb], filled=False, type_tag="XF", intronic="INTRONIC", intergenic="INTERGENIC")
    assert [(g[0], g[1], g[2], g[3], g[4]) for g in got] == [
        ("AAAACCCCGGGGTTTT", "ACGTACGTAC", "Gene1", "chr1", "4"), ("AAAACCCCGGGGTTTT", "TTTTTTTTTT", "Gene1", "chr1", "2"),
        ("CCCCCCCCGGGGTTTT", "ACGTACGTAC", "Gene2", "chr1", "2"), ("CCCCCCCCGGGGTTTT", "ACGTACGTAA", "Gene2", "chr2", "1"),
        ("CCCCCCCCGGGGTTTT", "ACGTACGTAG", "-", "chr2", "1")]
    assert meta[-1] == ["#stats", "6", "1", "0", "0"]
    got, _ = _dump([pa], filled=False, gene_in_chr=True)
    assert [g[2] for g in got] == ["chr1"] * 3 and [g[4] for g in got] == ["2"] * 3


def test_bulk_packed_path_equals_the_one_read_path(tmp_path):
    """parse_batch_packed (what parse_bam_files uses for -f input: views + 2-bit packed barcode / UMI, no string per read) takes the same
    decisions as read_info_from_alignment on every kind of record: missing / non-string / character-typed tags, empty values, N and
    lower-case bases (not packable), odd lengths, low base qualities, read-type values, unmapped / secondary / unknown-reference records."""
    rng = np.random.default_rng(21)
    acgt = np.array(list("ACGT"))
    als = []
    for i in range(30000):
        cb, umi = "".join(rng.choice(acgt, 16)), "".join(rng.choice(acgt, 10))
        kind = i % 47
        if kind == 2:
            umi = umi[:4] + "N" + umi[5:]
        elif kind == 4:
            cb = cb[:3] + "n" + cb[4:]
        elif kind == 6:
            cb = cb[:11]
        elif kind == 8:
            umi = umi + "ACGTACGT"          # 18 bases: not packable into 32 bits
        tags = [("NH", ("i", int(rng.integers(1, 5)))), ("CB", ("Z", cb)), ("UB", ("Z", umi))]
        if kind == 10:
            tags[1] = ("CB", ("A", "G"))
        elif kind == 12:
            tags[2] = ("UB", ("i", 7))
        elif kind == 14:
            tags[1] = ("CB", ("Z", ""))
        if i % 3:
            q = ["I"] * 10
            if kind == 16:
                q[int(rng.integers(0, 10))] = "$"
            tags += [("CQ", ("Z", "I" * 16)), ("UQ", ("Z", "".join(q)))]
        if i % 7:
            tags.append(("GX", ("Z", f"ENSG{int(rng.integers(0, 300)):09d}" if kind != 18 else "")))
            if i % 2:
                tags.append(("XF", ("Z", str(rng.choice(["CODING", "INTRONIC", "INTERGENIC", "UTR"])))) if kind != 20 else ("XF", ("A", "I")))
        flag = 4 if kind == 22 else 0x100 if kind == 24 else 0
        ref = 7 if kind == 26 else int(rng.integers(0, 3))
        als.append(alignment(f"r{i}", ref, i, flag, tags))
    path = str(tmp_path / "a.bam")
    write_bam(path, REFS, als, block_bytes=40000)
    for kw in (dict(min_q=20, type_tag="XF", intronic="INTRONIC", intergenic="INTERGENIC"), dict(min_q=0), dict(gene_in_chr=True)):
        one, meta_one = _dump([path], n_columns=7, threads=1, **kw)
        for threads in (1, 4):
            bulk, meta_bulk = _dump([path], n_columns=10, threads=threads, env={"DGE_BAM_PACKED": "1"}, **kw)
            assert [b[:7] for b in bulk] == one and meta_bulk[-1] == meta_one[-1] and len(one) > 20000
        code = {"A": 0, "C": 1, "G": 2, "T": 3}
        for b in bulk:
            cb_ok, umi_ok = set(b[0]) <= set("ACGT"), set(b[1]) <= set("ACGT") and len(b[1]) <= 16
            assert int(b[7]) == (1 if cb_ok else 0) | (2 if umi_ok else 0)
            if cb_ok:
                v = 0
                for ch in b[0]:
                    v = v * 4 + code[ch]
                assert int(b[8]) == v
            if umi_ok:
                v = 0
                for ch in b[1]:
                    v = v * 4 + code[ch]
                assert int(b[9]) == v


def test_bulk_packed_path_read_name_mode(tmp_path):
    """Without -f the bulk path takes barcode and UMI from the read name like ReadParameters::parse_encoded_id: the last '#', the last '!'
    before it, empty parts and names without the codec cannot be parsed; N bases and odd lengths are kept (not packable)."""
    rng = np.random.default_rng(33)
    acgt = np.array(list("ACGT"))
    names = []
    for i in range(20000):
        cb, umi = "".join(rng.choice(acgt, 12)), "".join(rng.choice(acgt, 8))
        kind = i % 31
        name = f"read{i}:x!{cb}#{umi}"
        if kind == 1: name = f"read{i}"                        # no codec
        elif kind == 3: name = f"read{i}#{umi}"                # no '!'
        elif kind == 5: name = f"read{i}!#{umi}"               # empty barcode
        elif kind == 7: name = f"read{i}!{cb}#"                # empty UMI
        elif kind == 9: name = f"a!b#c!{cb}#{umi}"             # several separators: the last ones count
        elif kind == 11: name = f"r!{cb}#{umi[:3]}N{umi[4:]}"  # N in the UMI
        elif kind == 13: name = f"r#1!{cb[:7]}#{umi}"          # '#' before the '!': still the last '#'
        elif kind == 15: name = f"r!{cb}!{cb}#{umi}"           # two '!': the last one before the '#'
        tags = [("GX", ("Z", f"G{int(rng.integers(0, 99))}"))] if i % 5 else []
        if i % 3 == 0:
            tags.append(("XF", ("Z", str(rng.choice(["INTRONIC", "CODING", "INTERGENIC"])))))
        names.append(alignment(name, int(rng.integers(0, 3)), i, 0, tags))
    path = str(tmp_path / "n.bam")
    write_bam(path, REFS, names, block_bytes=50000)
    for kw in (dict(type_tag="XF", intronic="INTRONIC", intergenic="INTERGENIC"), dict(), dict(gene_in_chr=True)):
        one, meta_one = _dump([path], filled=False, n_columns=7, threads=1, **kw)
        bulk, meta_bulk = _dump([path], filled=False, n_columns=7, threads=4, env={"DGE_BAM_PACKED": "1"}, **kw)
        assert bulk == one and meta_bulk[-1] == meta_one[-1] and 15000 < len(one) < 20000
        assert int(meta_one[-1][2]) > 2000   # the names that cannot be parsed are counted


HOST_PIPE = os.path.join(ROOT, "dropest_b200", "lib", "test_ingest_pipeline_host")


def _host_pipeline(files, threads=4, name_mode=False, min_q=0, env=None, expect_ok=True):
    assert os.path.exists(HOST_PIPE), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    r = subprocess.run([HOST_PIPE, str(threads), "1" if name_mode else "0", str(min_q)] + list(files), capture_output=True, text=True,
                       env=dict(os.environ, **(env or {})))
    assert (r.returncode == 0) == expect_ok, r.stdout[-500:] + r.stderr[-500:]
    return r.stdout.strip().split("\n")[-1].split("\t")


def test_host_side_of_parse_bam_files_threads_counters_and_errors(tmp_path):
    """BamProcessing::parse_bam_files itself (background loader + producer thread + hand-over thread) with the device entry points replaced
    by a no-op stand-in (tests/cpp/stub_device.c: test infrastructure, computes nothing): the counters, the number of reads that reach
    dge_add_batch*, bulk path == one-read path, two files, tiny chunks (many hand-overs between the three threads), and an error raised on
    the loader thread (truncated file) arriving at the caller after the earlier batches were handed over."""
    rng = np.random.default_rng(77)
    acgt = np.array(list("ACGT"))
    als, n_ok = [], 0
    for i in range(60000):
        cb, umi = "".join(rng.choice(acgt, 16)), "".join(rng.choice(acgt, 10))
        tags = [("CB", ("Z", cb)), ("UB", ("Z", umi)), ("CQ", ("Z", "I" * 16)), ("UQ", ("Z", "I" * 9 + ("#" if i % 13 == 0 else "I")))]
        if i % 5:
            tags.append(("GX", ("Z", f"G{int(rng.integers(0, 700))}")))
        if i % 17 == 0:
            tags = tags[1:]                   # no barcode tag: cannot be parsed
        flag = 4 if i % 19 == 0 else 0
        als.append(alignment(f"r{i}", int(rng.integers(0, 3)), i, flag, tags))
    a, b = str(tmp_path / "a.bam"), str(tmp_path / "b.bam")
    write_bam(a, REFS, als[:35000], block_bytes=20000)
    write_bam(b, REFS, als[35000:], block_bytes=20000)
    skipped = sum(1 for i in range(60000) if i % 19 == 0)
    cant = sum(1 for i in range(60000) if i % 19 and i % 17 == 0)
    lowq = sum(1 for i in range(60000) if i % 19 and i % 17 and i % 13 == 0)
    total = 60000 - skipped
    want = ["stats", str(total), str(cant), str(lowq), str(skipped), str(total - cant - lowq), "0", "0"]
    digests = set()
    for env in ({}, {"DGE_BAM_CHUNK_BYTES": "30000"}, {"DGE_BAM_ONE_BY_ONE": "1"}):
        for threads in (1, 5):
            got = _host_pipeline([a, b], threads=threads, min_q=10, env=env)
            assert got[:8] == want
            digests.add(got[8])
    assert len(digests) == 1   # packed keys, gene ids + marks, stream positions, chromosome ids: the same read by read on every path
    # read-name mode with N bases, barcodes of another length and names without the codec: the reads the packed record cannot carry take
    # the one-read path inside add_records, and the device still sees the same stream
    names = []
    for i in range(30000):
        cb, umi = "".join(rng.choice(acgt, 12)), "".join(rng.choice(acgt, 8))
        if i % 41 == 2: umi = umi[:3] + "N" + umi[4:]
        if i % 43 == 3: cb = cb[:5] + "N" + cb[6:]
        if i % 47 == 4: cb = cb[:9]
        if i % 53 == 5: umi = umi + "AC"
        name = f"x{i}!{cb}#{umi}" if i % 59 else f"x{i}"
        tags = [("GX", ("Z", f"G{int(rng.integers(0, 300))}"))] if i % 4 else []
        if i % 3 == 0:
            tags.append(("XF", ("Z", str(rng.choice(["INTRONIC", "CODING", "INTERGENIC"])))))
        names.append(alignment(name, int(rng.integers(0, 3)), i, 0, tags))
    c = str(tmp_path / "c.bam")
    write_bam(c, REFS, names, block_bytes=30000)
    bulk = _host_pipeline([c], name_mode=True)
    one = _host_pipeline([c], name_mode=True, env={"DGE_BAM_ONE_BY_ONE": "1"})
    assert bulk == one and int(bulk[5]) > 27000 and int(bulk[7]) > 0   # reads arrived; UMIs of another length were skipped and counted
    raw = open(b, "rb").read()
    cut = str(tmp_path / "cut.bam")
    open(cut, "wb").write(raw[:len(raw) // 2])
    for env in ({"DGE_BAM_CHUNK_BYTES": "30000"}, {"DGE_BAM_ONE_BY_ONE": "1"}):
        got = _host_pipeline([a, cut], env=env, expect_ok=False)
        assert got[0] == "error" and "truncated" in got[1]


REF_FLOW = os.path.join(ROOT, "oracle", "_ref", "ref_bam_flow")


@pytest.mark.skipif(not os.path.exists(REF_FLOW), reason="compiled reference (oracle/_ref/ref_bam_flow) not built")
def test_read_parameter_files_mode_matches_the_compiled_reference(tmp_path):
    """-r: barcode / UMI / UMI quality by read name from droptag's gzipped read-parameter files (ReadMapParamsParser.cpp:22-109): rows that
    cannot be parsed, empty barcodes, repeated names (first row wins), names with '@', reads missing from the files, a second alignment of
    a name (the row is handed out once), the base-quality threshold evaluated when the row is loaded -- the accepted reads, in order, are
    those of the reference's own BamController + ReadMapParamsParser."""
    import gzip

    rng = np.random.default_rng(9)
    acgt = np.array(list("ACGT"))
    rows_a, rows_b, als, text = [], [], [], [f"@SQ\t{n}\t{l}" for n, l in REFS]
    for i in range(6000):
        name = f"read{i}"
        cb, umi = "".join(rng.choice(acgt, 12)), "".join(rng.choice(acgt, 8))
        cbq = "".join(chr(33 + int(q)) for q in rng.integers(12, 41, 12))
        umiq = "".join(chr(33 + int(q)) for q in rng.integers(12, 41, 8))
        kind = i % 37
        row = f"{'@' if i % 2 else ''}{name} {cb} {umi} {cbq} {umiq}"
        if kind == 3:
            row = f"{name} {cb} {umi}"                       # too few fields
        elif kind == 5:
            row = f"{name}  {umi} {cbq} {umiq}"              # empty barcode
        elif kind == 7:
            row = None                                        # not in the files at all
        elif kind == 11:
            umiq = umiq[:3] + "#" + umiq[4:]                  # below the threshold
            row = f"{name} {cb} {umi} {cbq} {umiq}"
        if row is not None:
            (rows_a if i % 3 else rows_b).append(row)
            if kind == 13:
                rows_b.append(f"{name} {cb[::-1]} {umi} {cbq} {umiq}")   # a second row of the same name: ignored
        tags = [("NH", ("i", 1))] + ([("GX", ("Z", f"g{i % 40}"))] if i % 5 else [])
        flag = 4 if kind == 17 else 0
        ref = int(rng.integers(0, 3))
        n_copies = 2 if kind == 19 else 1                     # the same read name aligned twice
        for c in range(n_copies):
            als.append(alignment(("@" if kind == 23 else "") + name, ref, 10 + i + c, flag, tags))
            text.append("\t".join([("@" if kind == 23 else "") + name, str(ref), str(10 + i + c), "8M", str(flag)] + [f"{t}:{k}:{v}" for t, (k, v) in tags]))
    fa, fb = str(tmp_path / "params_a.gz"), str(tmp_path / "params_b.gz")
    gzip.open(fa, "wt").write("\n".join(rows_a) + "\n\n")
    gzip.open(fb, "wt").write("\n".join(rows_b))            # no newline at the end
    bam = str(tmp_path / "in.bam")
    write_bam(bam, REFS, als, block_bytes=45000)
    ref_dir = tmp_path / "ref"
    ref_dir.mkdir()
    open(str(ref_dir / "in.bam"), "w").write("\n".join(text) + "\n")
    r = subprocess.run([REF_FLOW, "--filled", "0", "--read-params", f"{fa} {fb}", "--min-quality", "10", "--bam-output", "1", "--filtered", "0",
                        "--min-genes-before", "1", "--min-genes-after", "1", "in.bam"], cwd=str(ref_dir), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-500:] + r.stderr[-500:]
    exp = []
    for line in open(str(ref_dir / "in.tagged.bam")).read().strip().split("\n"):
        f = line.split("\t")
        tag = {x[:2]: x[5:] for x in f[1:]}
        exp.append((tag["CR"], tag["UR"], tag.get("GX", "-"), tag["UQ"]))
    for threads in (1, 4):
        got, meta = _dump([bam], filled=False, min_q=10, threads=threads, n_columns=7, env={"DGE_BAM_READ_PARAMS": f"{fa}\t{fb}"})
        assert [(g[0], g[1], g[2], g[6]) for g in got] == exp and len(exp) > 4500
        assert all(g[5] == "-" for g in got)                  # the barcode quality is not kept (ReadParametersEfficient)
        n_mapped = sum(1 for i in range(6000) if i % 37 != 17) + sum(1 for i in range(6000) if i % 37 == 19)
        assert int(meta[-1][1]) == n_mapped and int(meta[-1][1]) - int(meta[-1][2]) - int(meta[-1][3]) == len(exp) and int(meta[-1][3]) > 100


def test_corrupt_files_fail_loudly(tmp_path):
    als = [alignment(f"r{i}", 0, i, 0, [("CB", ("Z", "ACGTACGTACGTACGT")), ("UB", ("Z", "ACGTACGTAC"))]) for i in range(5000)]
    path = str(tmp_path / "a.bam")
    write_bam(path, REFS, als, block_bytes=20000)
    data = bytearray(open(path, "rb").read())
    bad = str(tmp_path / "bad.bam")
    data[len(data) // 2] ^= 0x55
    open(bad, "wb").write(data)
    _, meta = _dump([bad], expect_ok=False)
    assert meta and meta[-1][0] == "#error"
    trunc = str(tmp_path / "trunc.bam")
    open(trunc, "wb").write(bytes(data[: len(data) // 3]))
    _, meta = _dump([trunc], expect_ok=False)
    assert meta and meta[-1][0] == "#error"
    _, meta = _dump([str(tmp_path / "missing.bam")], expect_ok=False)
    assert "Can't open BAM file" in meta[-1][1]


REF_RP = os.path.join(ROOT, "oracle", "_ref", "ref_readparams")


@pytest.mark.skipif(not os.path.exists(REF_RP), reason="compiled reference (oracle/_ref/ref_readparams) not built")
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_gene_assignment_from_annotation_matches_the_compiled_reference(tmp_path, seed):
    """-g: gene and mark from the annotation at the first and the last aligned base (ReadParamsParser::get_gene_from_reference), CIGARs with
    insertions / deletions / skips / clips, reads on chromosomes the annotation does not know; every alignment goes through the BAM file and
    BamIngest on one side and, as text, through the reference's own ReadParamsParser (BamAlignment shimmed) on the other."""
    import test_gene_annotation as tga

    rng = np.random.default_rng(100 + seed)
    if seed == 0:
        genes_file = os.path.join(ROOT, "tests", "golden", "ref_gtf", "gtf_test.gtf.gz")
        refs = [("chr1", 250000), ("chr2", 250000), ("chrX", 250000), ("chrNotInGtf", 1000)]
        spans = [(0, 11800, 72100), (1, 11800, 72100), (2, 34000, 36100)]
    else:
        text, genes = tga._random_gtf(rng, with_introns=seed == 2, all_have_transcripts=True)
        genes_file = str(tmp_path / "a.gtf")
        open(genes_file, "w").write(text)
        refs = [("chr1", 100000), ("chr2", 100000), ("chr3", 100000), ("chrNotInGtf", 1000)]
        spans = [(int(c[3:]) - 1, max(0, a - 40), b + 40) for c, a, b in genes]
    als, tsv = [], []
    for i in range(6000):
        ref, a, b = spans[int(rng.integers(0, len(spans)))] if rng.random() > 0.02 else (3, 0, 900)
        pos = int(rng.integers(a, b))
        ops = [("M", int(rng.integers(1, 60)))]
        for _ in range(int(rng.integers(0, 3))):
            ops.append((str(rng.choice(list("IDNS"))), int(rng.integers(1, 400 if seed == 0 else 40))))
            ops.append(("M", int(rng.integers(1, 60))))
        if rng.random() < 0.2:
            ops.insert(0, ("S", 5))
        name = f"r{i}!ACGTACGTACGTACGT#ACGTACGTAC"
        tags = [("GX", ("Z", "IGNORED"))]   # the gene tag must not be looked at in this mode
        als.append(alignment(name, ref, pos, 0, tags, seq_len=sum(l for o, l in ops if o in "MIS=X"), cigar_ops=ops))
        tsv.append(f"{name}\t{refs[ref][0]}\t{pos}\t{''.join(f'{l}{o}' for o, l in ops)}\tGX:Z:IGNORED")
    bam = str(tmp_path / "g.bam")
    write_bam(bam, refs, als, block_bytes=40000)
    tp = str(tmp_path / "g.tsv")
    open(tp, "w").write("\n".join(tsv) + "\n")
    ref_out = subprocess.run([REF_RP, genes_file, "0", "0", "0", "-", "-", "-", tp], capture_output=True, text=True)
    assert ref_out.returncode == 0, ref_out.stdout[-300:]
    exp, n_chr = [], 0
    for line, (ref_name) in zip(ref_out.stdout.strip().split("\n"), [x.split("\t")[1] for x in tsv]):
        if line == "!chr":
            n_chr += 1
            continue
        cb, umi, gene, mark, _ = line.split(" ")
        exp.append((cb, umi, gene, ref_name, mark))
    r = subprocess.run([DUMP, "0", "0", "0", "-", "-", "-", "3", bam], capture_output=True, text=True, env=dict(os.environ, DGE_BAM_GENES=genes_file))
    assert r.returncode == 0, r.stdout[-300:]
    lines = r.stdout.strip().split("\n")
    got = [tuple(l.split("\t")[:5]) for l in lines if not l.startswith("#")]
    assert got == exp
    assert lines[-1].split("\t")[:3] == ["#stats", str(len(als)), str(n_chr)]
    marks = {g[4] for g in got}
    assert {"0", "2"} <= marks and (marks & {"3", "4", "6"}) and n_chr > 50   # no gene, exonic, and mixed / intronic reads all occur
    # the bulk path (parse_batch_packed: read-name codec + annotation lookups on the parsing threads) takes the same decisions
    rp = subprocess.run([DUMP, "0", "0", "0", "-", "-", "-", "3", bam], capture_output=True, text=True,
                        env=dict(os.environ, DGE_BAM_GENES=genes_file, DGE_BAM_PACKED="1"))
    assert rp.returncode == 0, rp.stdout[-300:]
    plines = rp.stdout.strip().split("\n")
    assert [tuple(l.split("\t")[:5]) for l in plines if not l.startswith("#")] == exp and plines[-1] == lines[-1]


@pytest.mark.gpu
def test_bam_to_count_matrix_matches_the_reference(tmp_path):
    """End to end: a BAM with CB / UB / GX / XF tags (two files, records straddling BGZF blocks) -> BamIngest -> the container on the GPU
    -> filtered count matrix and per-chromosome exon table, against the oracle run on the same reads given as text."""
    import oracle_io
    from dropest_b200.capi import unpack_seq
    from dropest_b200.synth import SynthTables

    case = pu.small_case(n_reads=40000, n_cells=30, n_genes=80, merge="real", seed=23)
    recs = SynthTables(case.spec).generate_host(0, case.spec.n_reads)
    chr_ids = pu.synth_chr_ids(recs, 3)
    type_of = {1: "INTERGENIC", 2: "CODING", 4: "INTRONIC", 6: "CODING"}   # a read-type tag carries one type
    als, tsv = [], []
    for r, c in zip(recs, chr_ids):
        cb, umi = unpack_seq(int(r["key"]) >> 24, 16), unpack_seq(int(r["key"]) & 0xFFFFFF, case.umi_len)
        gene_id, mark = int(r["gene"]) & 0xFFFFFF, (int(r["gene"]) >> 24) & 7
        tags = [("CB", ("Z", cb)), ("UB", ("Z", umi))]
        gene = None if gene_id == 0xFFFFFF else f"g{gene_id}"
        if gene:
            tags += [("GX", ("Z", gene)), ("XF", ("Z", type_of[mark]))]
        eff_mark = 1 if gene is None else {"INTERGENIC": 1, "CODING": 2, "INTRONIC": 4}[type_of[mark]]
        als.append(alignment(f"q{len(als)}", int(c), 100 + len(als), 0, tags))
        tsv.append(f"{cb}\t{umi}\t{gene or '-'}\t{REFS[int(c)][0]}\t{eff_mark}")
    half = len(als) // 2
    pa, pb = str(tmp_path / "a.bam"), str(tmp_path / "b.bam")
    write_bam(pa, REFS, als[:half], block_bytes=50000)
    write_bam(pb, REFS, als[half:], block_bytes=33333)
    exe = os.path.join(ROOT, "dropest_b200", "lib", "test_bam_pipeline")
    out = subprocess.run([exe, case.barcodes, str(case.min_genes_before), str(case.min_genes_after), "XF", "INTRONIC", "INTERGENIC", pa, pb],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-800:] + out.stderr[-800:]
    rows = [l.split("\t") for l in out.stdout.strip().split("\n")]
    tsv_path = str(tmp_path / "reads.tsv")
    open(tsv_path, "w").write("\n".join(tsv) + "\n")
    ora = oracle_io.run_oracle(tsv_path, merge="real", barcodes=case.barcodes, barcodes_type="const", min_genes_before=case.min_genes_before,
                               min_genes_after=case.min_genes_after)
    o_bc = oracle_io.strings(ora["cell_barcodes"])
    o_genes = oracle_io.strings(ora["gene_names"])
    stats = [r for r in rows if r[0] == "stats"][0]
    assert int(stats[1]) == len(als) and int(stats[2]) == 0 and int(stats[4]) == int(ora["n_cells"][0]) and int(stats[5]) == int(ora["real_cells_number"][0])
    assert int(stats[6]) == int(ora["intergenic_reads"][0])
    assert [r[1] for r in rows if r[0] == "cell"] == [o_bc[i] for i in ora["filtered_cells"]]
    got = sorted((int(r[1]), r[2], int(r[3])) for r in rows if r[0] == "cm")
    exp = sorted((int(c), o_genes[int(g)], int(v)) for c, g, v in zip(ora["cm_col"], ora["cm_gene"], ora["cm_val"]))
    assert got == exp and len(got) > 500
    o_cells, o_chrs = oracle_io.strings(ora["chr_exon_cells"]), oracle_io.strings(ora["chr_exon_chrs"])
    o_counts = ora["chr_exon_counts"].reshape(len(o_cells), len(o_chrs))
    exp_exon = sorted((o_cells[r], o_chrs[c], int(o_counts[r, c])) for r in range(len(o_cells)) for c in range(len(o_chrs)) if o_counts[r, c])
    assert sorted((r[1], r[2], int(r[3])) for r in rows if r[0] == "exon") == exp_exon and exp_exon
