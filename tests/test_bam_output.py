"""The BAM writers: dropest_b200/host/BamOutput -- `-b` tagged BAMs and the `-F` filtered BAM with corrected barcodes and
UMIs.  CPU tests: BGZF / BAM framing, tag editing, round trip through our own reader.  GPU tests: the whole flow (BAM -> container on the
device -> merge_and_filter -> second pass) against the reference's own BamController / BamProcessor / FilteringBamProcessor compiled
unmodified (oracle/_ref/ref_bam_flow: BamTools shimmed onto text files) on the same alignments."""
import os
import subprocess
import zlib

import numpy as np
import pytest

from bam_utils import alignment, read_bam, write_bam

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "dropest_b200", "lib", "test_bam_output")
DUMP = os.path.join(ROOT, "dropest_b200", "lib", "test_bam_ingest")
REF_FLOW = os.path.join(ROOT, "oracle", "_ref", "ref_bam_flow")
REFS = [("chr1", 1000000), ("chr2", 900000), ("chrM", 16000)]


def _tag_text(tags):
    out = []
    for name, (kind, v) in tags:
        if kind == "B":
            out.append(f"{name}:B:S," + ",".join(str(x) for x in v))
        else:
            out.append(f"{name}:{kind}:{v}")
    return out


def _expected_edit(tag_texts, edits):
    """BamAlignment::EditTag for a list of (tag, value): an existing tag is removed, the new one is appended; names that are not two characters
    are ignored"""
    cur = list(tag_texts)
    for tag, value in edits:
        if len(tag) != 2:
            continue
        cur = [t for t in cur if t[:2] != tag]
        cur.append(f"{tag}:Z:{value}")
    return cur


def test_tagged_bam_framing_tags_and_round_trip(tmp_path):
    """Every accepted read comes out with the edits of BamProcessorAbstract::save_alignment in their order (gene, raw barcode, raw UMI,
    qualities, read type, corrected barcode / UMI); other tags -- integer, array, character -- and the fixed part of the record are carried
    over untouched; skipped reads are not written; the output is valid BGZF (many blocks, several deflate threads) and our own reader
    reads it back."""
    rng = np.random.default_rng(5)
    acgt = np.array(list("ACGT"))
    type_vals = {1: "INTERGENIC", 2: "CODING", 4: "INTRONIC"}
    als, exp = [], []
    for i in range(30000):
        cb, umi = "".join(rng.choice(acgt, 16)), "".join(rng.choice(acgt, 10))
        gene = f"G{int(rng.integers(0, 50))}" if rng.random() > 0.2 else None
        mark = int(rng.choice([1, 2, 4]))
        tags = [("NH", ("i", int(rng.integers(1, 9)))), ("CR", ("Z", "stale")), ("CB", ("Z", cb)), ("xs", ("B", [1, 2, i % 65536])), ("UB", ("Z", umi))]
        if i % 3 == 0:
            tags += [("CQ", ("Z", "I" * 16)), ("UQ", ("Z", "F" * 10))]
        if gene:
            tags += [("GX", ("Z", gene)), ("XF", ("Z", type_vals[mark])), ("ch", ("A", "x"))]
        flag = 4 if i % 17 == 3 else 0x100 if i % 17 == 5 else 16 if i % 2 else 0
        if i % 29 == 11:
            tags = [t for t in tags if t[0] != "CB"]   # cannot be parsed: skipped
        name = f"read{i}" + "x" * (i % 7)
        als.append(alignment(name, int(rng.integers(0, 3)), int(rng.integers(0, 900000)), flag, tags, seq_len=8 + i % 5))
        if flag & 0x104 or i % 29 == 11:
            continue
        edits = []
        eff_mark = 1 if gene is None else mark
        if gene:
            edits.append(("GX", gene))
        edits += [("CR", cb), ("UR", umi)]
        if i % 3 == 0:
            edits += [("CQ", "I" * 16), ("UQ", "F" * 10)]
        edits.append(("XF", {1: "INTERGENIC", 2: "EXONIC", 4: "INTRONIC"}[eff_mark]))
        edits.append(("CB", cb[::-1]))
        if umi[0] == "A":
            edits.append(("UB", umi))
        exp.append((name, _expected_edit(_tag_text(tags), edits)))
    half = len(als) // 2
    pa, pb = str(tmp_path / "in.a.bam"), str(tmp_path / "b.bam")
    write_bam(pa, REFS, als[:half], block_bytes=50000, header_text="@HD\tVN:1.6\n@PG\tID:test\n")
    write_bam(pb, REFS, als[half:], block_bytes=33333)
    out_dir = tmp_path / "out"
    out_dir.mkdir()
    r = subprocess.run([EXE, "cpu", str(out_dir), "3", "XF", "INTRONIC", "INTERGENIC", pa, pb], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-500:] + r.stderr[-500:]
    stats = r.stdout.strip().split("\t")
    assert int(stats[4]) == len(exp)
    # names: "<input without its extension>.tagged.bam", directory stripped (BamProcessor::get_result_bam_name + update_bam)
    oa, ob = str(out_dir / "in.a.tagged.bam"), str(out_dir / "b.tagged.bam")
    ta, ra, reca = read_bam(oa)
    tb, rb, recb = read_bam(ob)
    assert ta == "@HD\tVN:1.6\n@PG\tID:test\n" and tb == "@HD\tVN:1.6\n" and ra == REFS and rb == REFS
    got = [(n, t) for n, _, _, _, t in reca + recb]
    assert got == exp
    # fixed fields survive: compare with the input records
    _, _, in_a = read_bam(pa)
    by_name = {n: (ref, pos, flag) for n, ref, pos, flag, _ in in_a}
    assert all(by_name[n] == (ref, pos, flag) for n, ref, pos, flag, _ in reca)
    assert os.path.getsize(oa) > 3 * 65536   # several BGZF blocks
    # our own reader reads the output back: -f mode now sees the corrected barcode in CB and the raw one in CR
    d = subprocess.run([DUMP, "1", "0", "0", "XF", "INTRONIC", "INTERGENIC", "2", oa, ob], capture_output=True, text=True)
    assert d.returncode == 0, d.stdout[-300:] + d.stderr[-300:]
    lines = [l.split("\t") for l in d.stdout.strip().split("\n") if not l.startswith("#")]
    n_with_ub = sum(1 for _, t in exp if any(x.startswith("UB:") for x in t))
    assert len(lines) == n_with_ub   # reads whose UB tag was dropped by a stale-free edit cannot be parsed any more: none here lose it


def test_edit_of_the_same_tag_twice_and_empty_tag_names(tmp_path):
    """A read-type tag that is not configured has an empty name: BamTools' AddTag refuses it, nothing is written for it; configuring the raw
    barcode tag to the name of the corrected one leaves the later edit."""
    als = [alignment("r0", 0, 5, 0, [("CB", ("Z", "ACGTACGTACGTACGT")), ("UB", ("Z", "AAAAAAAAAA")), ("GX", ("Z", "g"))])]
    p = str(tmp_path / "x.bam")
    write_bam(p, REFS, als)
    r = subprocess.run([EXE, "cpu", str(tmp_path), "1", "-", "-", "-", p], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    _, _, rec = read_bam(str(tmp_path / "x.tagged.bam"))
    assert rec[0][4] == ["GX:Z:g", "CR:Z:ACGTACGTACGTACGT", "UR:Z:AAAAAAAAAA", "CB:Z:TGCATGCATGCATGCA", "UB:Z:AAAAAAAAAA"]


def _flow_case(seed, n_reads, n_cells, umi_len, with_n, whitelist):
    """alignments of a small experiment: a few large cells + error barcodes one mismatch away, few genes and short UMIs (so that UMIs
    collide and the directional merge has work), optional N in UMIs; returns (list of (name, ref, pos, flag, tags)), whitelist lines"""
    rng = np.random.default_rng(seed)
    acgt = np.array(list("ACGT"))
    p1 = ["".join(rng.choice(acgt, 8)) for _ in range(12)]
    p2 = ["".join(rng.choice(acgt, 8)) for _ in range(12)]
    cells = [p1[int(rng.integers(0, 12))] + p2[int(rng.integers(0, 12))] for _ in range(n_cells)]
    # per (cell, gene) a small pool of true UMIs with skewed read counts; errors = one substitution of a true UMI
    reads = []
    for i in range(n_reads):
        c = true_c = cells[int(rng.integers(0, n_cells))]
        r = rng.random()
        if r < 0.12:   # one of two recurring error barcodes of the cell (enough reads to become real cells that the whitelist merge folds back)
            er = np.random.default_rng(zlib.crc32(c.encode()) + int(r < 0.06))
            k = int(er.integers(0, 16))
            c = c[:k] + "ACGT"[("ACGT".index(c[k]) + 1 + int(er.integers(0, 3))) % 4] + c[k + 1:]
        elif r < 0.15:   # scattered errors
            k = int(rng.integers(0, 16))
            c = c[:k] + str(rng.choice(acgt)) + c[k + 1:]
        g = int(rng.zipf(1.6)) % 25
        lr = np.random.default_rng(zlib.crc32(f"{true_c}:{g}".encode()))
        pool = ["".join(lr.choice(acgt, umi_len)) for _ in range(4)]
        umi = pool[min(3, int(rng.exponential(0.8)))]
        if rng.random() < 0.2:
            k = int(rng.integers(0, umi_len))
            umi = umi[:k] + str(rng.choice(acgt)) + umi[k + 1:]
        if with_n and rng.random() < 0.06:
            k = int(rng.integers(0, umi_len))
            umi = umi[:k] + "N" + umi[k + 1:]
        gene = None if rng.random() < 0.08 else f"g{g}"
        xf = str(rng.choice(["CODING", "INTRONIC", "INTERGENIC"], p=[0.7, 0.2, 0.1]))
        tags = [("NH", ("i", 1)), ("CB", ("Z", c)), ("UB", ("Z", umi))]
        if gene:
            tags += [("GX", ("Z", gene)), ("XF", ("Z", xf))]
        flag = 4 if i % 41 == 7 else 0
        reads.append((f"q{i}", int(rng.integers(0, 3)), 100 + i, flag, tags))
    wl = [" ".join(p1), " ".join(p2)] if whitelist else None
    return reads, wl


def _as_text(reads):
    lines = [f"@SQ\t{n}\t{l}" for n, l in REFS]
    for name, ref, pos, flag, tags in reads:
        lines.append("\t".join([name, str(ref), str(pos), "8M", str(flag)] + _tag_text(tags)))
    return "\n".join(lines) + "\n"


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(REF_FLOW), reason="compiled reference (oracle/_ref/ref_bam_flow) not built")
@pytest.mark.parametrize("merge,umi_merge,with_n,umi_len", [("real", "simple", True, 8), ("none", "directional", False, 6), ("real", "directional", True, 6),
                                                            ("none", "simple", False, 8)])
def test_filtered_and_tagged_bams_match_the_compiled_reference(tmp_path, merge, umi_merge, with_n, umi_len):
    """-b and -F end to end: two input BAMs -> tagged BAMs per input + ONE filtered BAM (named after the first input) whose reads carry the
    merged cell barcode (CB) and the merged / repaired UMI (UB) next to the raw values (CR / UR); same reads, same order, same tag blocks
    as the unmodified reference flow."""
    reads, wl = _flow_case(seed=3 + umi_len + int(with_n), n_reads=30000, n_cells=12, umi_len=umi_len, with_n=with_n, whitelist=merge == "real")
    half = len(reads) // 2
    ours, theirs = tmp_path / "ours", tmp_path / "ref"
    ours.mkdir(); theirs.mkdir()
    wl_path = "-"
    if wl:
        wl_path = str(tmp_path / "wl.txt")
        open(wl_path, "w").write("\n".join(wl) + "\n")
    # the reference reverse-complements whitelist tokens on load (BarcodesParser.cpp): write the file so that the loaded tokens are ours
    if wl:
        comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
        rc = lambda s: "".join(comp[c] for c in reversed(s))
        open(wl_path, "w").write("\n".join(" ".join(rc(t) for t in line.split()) for line in wl) + "\n")
    for k, part in enumerate((reads[:half], reads[half:])):
        write_bam(str(tmp_path / f"in{k}.bam"), REFS, [alignment(n, r, p, f, t) for n, r, p, f, t in part], block_bytes=40000 + 1000 * k)
        open(str(theirs / f"in{k}.bam"), "w").write(_as_text(part))
    cmd = [REF_FLOW, "--merge", merge, "--umi-merge", umi_merge, "--min-genes-before", "5", "--min-genes-after", "8", "--type-tag", "XF", "--intronic", "INTRONIC",
           "--intergenic", "INTERGENIC", "--bam-output", "1", "--filtered", "1"]
    if wl:
        cmd += ["--barcodes", wl_path, "--barcodes-type", "const"]
    r = subprocess.run(cmd + ["in0.bam", "in1.bam"], cwd=str(theirs), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-500:] + r.stderr[-500:]
    ref_counts = r.stdout.strip().split("\t")
    o = subprocess.run([EXE, "gpu", str(ours), wl_path, "5", "8", umi_merge, "XF", "INTRONIC", "INTERGENIC", "1", str(tmp_path / "in0.bam"), str(tmp_path / "in1.bam")],
                       capture_output=True, text=True)
    assert o.returncode == 0, o.stdout[-800:] + o.stderr[-800:]
    stats = o.stdout.strip().split("\n")[-1].split("\t")
    assert stats[0] == "stats" and [stats[2], stats[3], stats[4]] == [ref_counts[1], ref_counts[3], ref_counts[5]]
    assert int(stats[6]) == 0 and int(stats[7]) == 0 and stats[8].endswith("in0.filtered.bam")

    def ref_lines(path):
        return [(l.split("\t")[0], l.split("\t")[1:]) for l in open(path).read().strip().split("\n") if l]

    for name in ("in0.tagged.bam", "in1.tagged.bam", "in0.filtered.bam"):
        _, refs, rec = read_bam(str(ours / name))
        assert refs == REFS
        exp = ref_lines(str(theirs / name))
        assert [(n, t) for n, _, _, _, t in rec] == exp, name
    assert not os.path.exists(str(ours / "in1.filtered.bam"))
    _, _, filt = read_bam(str(ours / "in0.filtered.bam"))
    assert int(stats[5]) == len(filt) and len(filt) > 5000
    tag = lambda t, k: next(x[5:] for x in t if x.startswith(k + ":"))
    if merge == "real":
        assert sum(1 for _, _, _, _, t in filt if tag(t, "CB") != tag(t, "CR")) > 200      # reads of merged error barcodes
    if umi_merge == "directional" or with_n:
        assert sum(1 for _, _, _, _, t in filt if tag(t, "UB") != tag(t, "UR")) > 100      # merged / repaired UMIs
