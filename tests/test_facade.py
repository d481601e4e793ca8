"""The host-side C++ mirror of CellsDataContainer / ResultsPrinter (dropest_b200/host) running the reference's own test
sequence on the GPU, and a structural check of the files it writes (.rds parsed back by a minimal reader, .mtx, tsv)."""
import gzip
import os
import struct
import subprocess

import pytest

import parity_utils as pu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "dropest_b200", "lib", "test_facade")


class Rds:
    """Just enough of R's XDR serialization (version 2) to read back what ResultsPrinter::save_rds writes."""

    def __init__(self, data):
        self.d, self.o, self.syms = data, 0, []

    def i32(self):
        v = struct.unpack(">i", self.d[self.o:self.o + 4])[0]
        self.o += 4
        return v

    def item(self):
        fl = self.i32()
        t = fl & 255
        has_attr, has_tag = bool(fl & (1 << 9)), bool(fl & (1 << 10))
        if t == 254:
            return None
        if t == 255:
            return self.syms[(fl >> 8) - 1]
        if t == 1:
            s = self.item()
            self.syms.append(s)
            return s
        if t == 9:
            n = self.i32()
            s = self.d[self.o:self.o + n].decode()
            self.o += n
            return s
        if t == 2:  # pairlist -> dict
            out = {}
            while True:
                tag = self.item() if has_tag else None
                out[tag] = self.item()
                fl = self.i32()
                if (fl & 255) == 254:
                    return out
                assert (fl & 255) == 2
                has_tag = bool(fl & (1 << 10))
        if t in (13, 14, 16, 19):
            n = self.i32()
            if t == 13:
                v = list(struct.unpack(f">{n}i", self.d[self.o:self.o + 4 * n])); self.o += 4 * n
            elif t == 14:
                v = list(struct.unpack(f">{n}d", self.d[self.o:self.o + 8 * n])); self.o += 8 * n
            else:
                v = [self.item() for _ in range(n)]
            attr = self.item() if has_attr else None
            return {"v": v, "attr": attr} if attr else v
        if t == 25:
            return {"S4": self.item()}
        raise ValueError(f"unexpected SEXP type {t}")


@pytest.mark.gpu
def test_reference_container_tests_through_the_cpp_facade(tmp_path):
    assert os.path.exists(BIN), "build the facade test first: python -c 'import __graft_entry__ as g; g.build()'"
    r = subprocess.run([BIN, pu.WL_TEST_EST, str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OK (0 failures)" in r.stdout
    # ---- MatrixMarket + names (ResultsPrinter.cpp:81-91)
    lines = open(tmp_path / "cell.counts.mtx").read().split("\n")
    assert lines[0].startswith("%%MatrixMarket matrix coordinate")
    nrow, ncol, nnz = map(int, lines[1].split())
    assert (ncol, nnz) == (2, 7) and nrow == 6
    cells = open(tmp_path / "cell.counts.cells.tsv").read().split()
    genes = open(tmp_path / "cell.counts.genes.tsv").read().split()
    assert cells == ["AAATTAGGTCCC", "AAATTAGGTCCA"] and sorted(genes) == sorted(["Gene1", "Gene10", "Gene20", "Gene2", "Gene3", "Gene4"])
    trip = [tuple(map(int, l.split())) for l in lines[2:2 + nnz]]
    assert all(1 <= i <= nrow and 1 <= j <= ncol for i, j, _ in trip)
    assert [t[1] for t in trip] == sorted(t[1] for t in trip)  # column-major
    by = {(genes[i - 1], cells[j - 1]): v for i, j, v in trip}
    assert by[("Gene1", "AAATTAGGTCCA")] == 2 and by[("Gene3", "AAATTAGGTCCA")] == 2 and by[("Gene10", "AAATTAGGTCCC")] == 1
    # ---- .rds: list(cm = dgCMatrix, cm_raw, merge_targets, ...), the contract of docs/dropest.rst:178-194
    raw = gzip.open(tmp_path / "cell.counts.rds").read()
    assert raw[:2] == b"X\n"
    rd = Rds(raw)
    rd.o = 2
    assert rd.i32() == 2
    rd.i32(); rd.i32()
    d = rd.item()
    names = d["attr"]["names"]
    assert names == ["cm", "cm_raw", "reads_per_chr_per_cells", "mean_reads_per_umi", "saturation_info", "merge_targets", "aligned_reads_per_cell",
                     "aligned_umis_per_cell", "requested_umis_per_cb", "requested_reads_per_cb"]   # ResultsPrinter.cpp:47-57
    cm = d["v"][0]["S4"]
    assert cm["class"]["v"] == ["dgCMatrix"] and cm["class"]["attr"]["package"] == ["Matrix"]
    assert cm["Dim"] == [6, 2] and cm["p"] == [0, 3, 7] and len(cm["i"]) == 7 and sum(cm["x"]) == 9.0
    assert cm["Dimnames"][1] == cells and cm["Dimnames"][0] == genes
    for c in range(2):
        col = cm["i"][cm["p"][c]:cm["p"][c + 1]]
        assert col == sorted(col)  # row indices ascending inside a column (dgCMatrix invariant)
    # reads_per_chr_per_cells = list(Exon, Intron, Intergenic) of cells x chromosomes integer matrices (ResultsPrinter.cpp:140-166)
    rpc = d["v"][2]
    assert rpc["attr"]["names"] == ["Exon", "Intron", "Intergenic"]
    exon = rpc["v"][0]
    assert exon["attr"]["dim"] == [2, 3] and exon["attr"]["dimnames"][0] == ["AAATTAGGTCCA", "AAATTAGGTCCC"]
    chrs = exon["attr"]["dimnames"][1]
    assert sorted(chrs) == ["chr1", "chr2", "chr3"]
    got = {(r, chrs[c]): exon["v"][c * 2 + r] for r in range(2) for c in range(3)}   # column-major
    assert got == {(0, "chr1"): 4, (0, "chr2"): 4, (0, "chr3"): 4, (1, "chr1"): 2, (1, "chr2"): 0, (1, "chr3"): 2}
    assert rpc["v"][1]["attr"]["dim"] == [0, 0] and rpc["v"][2]["attr"]["dim"] == [0, 0]
    mrpu = d["v"][3]
    assert mrpu["attr"]["names"] == ["AAATTAGGTCCA", "AAATTAGGTCCC"] and mrpu["v"] == [12 / 6, 4 / 3]   # reads / distinct UMIs held after the merge
    sat = d["v"][4]
    assert sat["attr"]["names"] == ["reads", "cbs", "umis"]
    sat_rows = sorted(zip(sat["v"][1], sat["v"][2], sat["v"][0]))
    assert len(sat_rows) == 9 and sum(r for _, _, r in sat_rows) == 16 and ("AAATTAGGTCCA", "CCCCCT", 4) in sat_rows
    mt = d["v"][5]
    assert dict(zip(mt["attr"]["names"], mt["v"])) == {"AAATTAGGTCCG": "AAATTAGGTCCC", "AAATTAGGTCGG": "AAATTAGGTCCA",
                                                      "CCCTTAGGTCCA": "AAATTAGGTCCA", "CAATTAGGTCCG": "AAATTAGGTCCA"}
    umis = d["v"][7]
    assert dict(zip(umis["attr"]["names"], umis["v"])) == {"AAATTAGGTCCA": 12, "AAATTAGGTCCC": 4}
    rreads = d["v"][9]
    assert dict(zip(rreads["attr"]["names"], rreads["v"])) == {"AAATTAGGTCCA": 12, "AAATTAGGTCCC": 4}
    # ---- -V: list(exon, intron, spanning) of dgCMatrix over the filtered cells (ResultsPrinter.cpp:455-474); the fixture's reads are all exonic
    raw = gzip.open(tmp_path / "cell.counts.matrices.rds").read()
    rd = Rds(raw)
    rd.o = 2
    assert rd.i32() == 2
    rd.i32(); rd.i32()
    m = rd.item()
    assert m["attr"]["names"] == ["exon", "intron", "spanning"]
    ex, intr, span = (x["S4"] for x in m["v"])
    assert ex["Dim"] == [6, 2] and ex["p"] == cm["p"] and ex["i"] == cm["i"] and ex["x"] == cm["x"] and ex["Dimnames"] == cm["Dimnames"]
    for e in (intr, span):
        assert e["Dim"] == [0, 2] and e["p"] == [0, 0, 0] and e["i"] == [] and e["Dimnames"][1] == cells


def test_merge_strategy_factory_reads_the_xml_configuration():
    """MergeStrategyFactory::from_xml (the <Estimation> block of configs/*.xml): keys, defaults, mandatory max_cb_merge_edit_distance,
    barcodes_file resolved against the configuration file, -G override, strategy selection -- CPU only."""
    exe = os.path.join(os.path.dirname(BIN), "test_factory_xml")
    assert os.path.exists(exe), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    r = subprocess.run([exe, os.path.join(pu.GOLDEN, "configs")], capture_output=True, text=True)
    assert r.returncode == 0 and "OK (0 failures)" in r.stdout, r.stdout + r.stderr
