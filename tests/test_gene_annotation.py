"""Gene annotation: dropest_b200/host/GeneAnnotation against the reference's own RefGenesContainer, compiled unmodified
into oracle/_ref/ref_gtf, on the reference's test annotation and on randomly generated GTF / BED files with every awkward shape the
format allows (touching exons in both insertion orders, intron records, overlapping transcripts, records without transcript or gene id,
"." columns, comments, gzip).  CPU only.  The outputs for the fixture annotation are also committed (tests/golden/ref_gtf), so the pin
holds on machines without the compiled reference."""
import gzip
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MINE = os.path.join(ROOT, "dropest_b200", "lib", "test_gene_annotation")
REF = os.path.join(ROOT, "oracle", "_ref", "ref_gtf")
FIX = os.path.join(ROOT, "tests", "golden", "ref_gtf")
needs_ref = pytest.mark.skipif(not os.path.exists(REF), reason="compiled reference (oracle/_ref/ref_gtf) not built")


def _run(exe, genes, queries_path):
    r = subprocess.run([exe, genes, queries_path], capture_output=True, text=True)
    return r.returncode, r.stdout


def _write_queries(path, queries):
    with open(path, "w") as f:
        for c, s, e in queries:
            f.write(f"{c}\t{s}\t{e}\n")


def _fixture_queries():
    q = []
    for chrom in ("chr1", "chr2", "chrX", "chr7"):
        for p in range(11800, 72100, 37):
            q.append((chrom, p, p + 1))
        for p in range(11000, 73000, 501):
            q.append((chrom, p, p + 10))
    q += [("chr1", 20000, 19000), ("chr1", 0, 10 ** 9)]
    return q


def test_fixture_annotation_matches_committed_reference_output(tmp_path):
    assert os.path.exists(MINE), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    qp = str(tmp_path / "q.tsv")
    _write_queries(qp, _fixture_queries())
    rc, out = _run(MINE, os.path.join(FIX, "gtf_test.gtf.gz"), qp)
    assert rc == 0
    assert out == open(os.path.join(FIX, "gtf_test.expected.txt")).read()
    lines = out.split("\n")
    assert "WASH7P:1" in lines and "AR4F5:2,BR4F5:2,OR4F5:2" in lines and "!chr" in lines   # Tests/TestTools.cpp:127-183,262-284


@needs_ref
def test_committed_expectation_is_what_the_reference_prints(tmp_path):
    qp = str(tmp_path / "q.tsv")
    _write_queries(qp, _fixture_queries())
    rc, out = _run(REF, os.path.join(FIX, "gtf_test.gtf.gz"), qp)
    assert rc == 0 and out == open(os.path.join(FIX, "gtf_test.expected.txt")).read()


def _random_gtf(rng, with_introns, all_have_transcripts, clean=True):
    lines = ["# comment line", "#!genome-build test"]
    genes = []
    for g in range(int(rng.integers(3, 12))):
        chrom = f"chr{int(rng.integers(1, 4))}"
        gstart = int(rng.integers(100, 5000))
        for t in range(int(rng.integers(1, 4))):
            tid = f"T{g}_{t}"
            pos = gstart + int(rng.integers(0, 60))
            exons = []
            for _ in range(int(rng.integers(1, 6))):
                length = int(rng.integers(5, 120))
                exons.append((pos, pos + length))
                gap = int(rng.choice([0, 0, 1, 1, 7, 30, 200]))          # 0 / 1: touching or abutting exons
                pos += length + gap
            order = rng.permutation(len(exons))                          # insertion order matters for touching spans
            for k in order:
                s, e = exons[k]
                attrs = f'gene_id "G{g}"; gene_name "N{g}";'
                if all_have_transcripts or rng.random() > 0.3:
                    attrs += f' transcript_id "{tid}";'
                if rng.random() < 0.1:
                    attrs = f'gene_name "N{g}"; transcript_id "{tid}";'   # no gene_id: the name stands in
                lines.append(f"{chrom}\tsrc\texon\t{s + 1}\t{e}\t.\t+\t.\t{attrs} extra \"x\";")
            if with_introns:
                for (s0, e0), (s1, e1) in zip(exons[:-1], exons[1:]):
                    if s1 - e0 >= 2:
                        lines.append(f'{chrom}\tsrc\tintron\t{e0 + 1}\t{s1}\t.\t+\t.\tgene_id "G{g}"; gene_name "N{g}"; transcript_id "{tid}";')
            genes.append((chrom, gstart, pos))
    lines.append("chr1\tsrc\tCDS\t10\t20\t.\t+\t.\tgene_id \"G0\"; transcript_id \"T0_0\";")      # other feature types are ignored
    lines.append(".\tsrc\texon\t10\t20\t.\t+\t.\tgene_id \"G0\"; transcript_id \"T0_0\";")          # "." columns are skipped
    lines.append("chr1\tsrc\texon\t10\t20\t.\t+\t.\tgene_id")                                        # 9 tokens: skipped
    if not clean:
        lines.append("chr1 src exon 5")                                                              # too short: reported, skipped
        lines.append('chr2\tsrc\texon\t50\t60\t.\t+\t.\ttss_id "Z"; other "y";')                     # neither id nor name: reported, skipped
    return "\n".join(lines) + ("\n" if rng.random() > 0.3 else ""), genes


def _random_queries(rng, genes, n=1500):
    q = []
    for _ in range(n):
        chrom, a, b = genes[int(rng.integers(0, len(genes)))]
        p = int(rng.integers(max(0, a - 50), b + 50))
        width = int(rng.choice([1, 1, 1, 2, 10, 100]))
        q.append((chrom, p, p + width))
    q += [("chr1", 0, 1), ("chrUn", 5, 6), ("chr2", 100, 50)]
    return q


@needs_ref
@pytest.mark.parametrize("seed", range(12))
def test_random_gtf_and_bed_match_the_compiled_reference(tmp_path, seed):
    assert os.path.exists(MINE)
    rng = np.random.default_rng(seed)
    text, genes = _random_gtf(rng, with_introns=seed % 3 == 0, all_have_transcripts=seed % 2 == 0, clean=seed % 4 != 1)
    gz = seed % 2 == 1
    path = str(tmp_path / ("a.gtf.gz" if gz else "a.gtf"))
    if gz:
        with gzip.open(path, "wt") as f:
            f.write(text)
    else:
        open(path, "w").write(text)
    qp = str(tmp_path / "q.tsv")
    _write_queries(qp, _random_queries(rng, genes))
    a, b = _run(REF, path, qp), _run(MINE, path, qp)
    assert a == b
    if a[0] == 0:
        assert sum(1 for l in a[1].split("\n") if l not in ("-", "!chr", "")) > 100
    # the same intervals as a BED file (chrom, start, end, name): one "transcript" per gene name
    bed = []
    for line in text.split("\n"):
        c = line.split("\t")
        if len(c) >= 9 and c[2] == "exon" and c[0] != "." and "gene_id \"G" in c[8]:
            name = c[8].split('"')[1]
            bed.append(f"{c[0]}\t{int(c[3]) - 1}\t{c[4]}\t{name}\t0\t+")
    bp = str(tmp_path / "a.bed")
    open(bp, "w").write("# bed\n\n" + "\n".join(bed) + "\n")
    assert _run(REF, bp, qp) == _run(MINE, bp, qp)


@needs_ref
def test_error_cases_match_the_compiled_reference(tmp_path):
    qp = str(tmp_path / "q.tsv")
    _write_queries(qp, [("chr1", 5, 6)])
    cases = {
        "overlap.gtf": 'chr1\ts\texon\t10\t50\t.\t+\t.\tgene_id "A"; transcript_id "T";\nchr1\ts\tintron\t40\t80\t.\t+\t.\tgene_id "A"; transcript_id "T";\n',
        "two_genes.gtf": 'chr1\ts\texon\t10\t50\t.\t+\t.\tgene_id "A"; transcript_id "T";\nchr1\ts\texon\t60\t80\t.\t+\t.\tgene_id "B"; transcript_id "T";\n',
        "empty_line.gtf": 'chr1\ts\texon\t10\t50\t.\t+\t.\tgene_id "A"; transcript_id "T";\n\nchr1\ts\texon\t60\t80\t.\t+\t.\tgene_id "A"; transcript_id "T";\n',
        "wrong.txt": "x\n",
    }
    for name, text in cases.items():
        p = str(tmp_path / name)
        open(p, "w").write(text)
        a, b = _run(REF, p, qp), _run(MINE, p, qp)
        assert a[0] == b[0] == 1 and a[1] == b[1], name
    assert _run(REF, str(tmp_path / "missing.gtf"), qp) == _run(MINE, str(tmp_path / "missing.gtf"), qp)
