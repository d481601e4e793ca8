"""`reads_per_umi_per_cell` of the .rds (ResultsPrinter::get_reads_per_umi_per_cell, ResultsPrinter.cpp:261-314) with
UMI::mean_quality (UMI.cpp:46-55): per filtered cell and requested gene, every requested UMI with its read count and the per-base "mean"
quality.  The sums belong to UMI OBJECTS and merges move or drop objects (Gene.cpp:26-58), so the values depend on the order in which cells
were merged and on which UMIs the UMI merge created -- compared here with the unmodified reference on streams with barcode merges (with and
without chains), directional UMI merges and N repair."""
import gzip
import os
import subprocess

import numpy as np
import pytest

import oracle_io
import parity_utils as pu
from test_bam_output import _flow_case
from test_facade import Rds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "dropest_b200", "lib", "test_rpupc")

pytestmark = pytest.mark.gpu


def _tsv(reads, umi_len, seed, with_quality=True):
    rng = np.random.default_rng(seed)
    lines = []
    for _name, ref, _pos, flag, tags in reads:
        if flag & 4:
            continue
        d = {t: v[1] for t, v in tags}
        mark = {"CODING": 2, "INTRONIC": 4, "INTERGENIC": 1}[d["XF"]] if "GX" in d else 1
        qual = "".join(chr(33 + int(q)) for q in rng.integers(2, 41, umi_len)) if with_quality else ""
        lines.append("\t".join([d["CB"], d["UB"], d.get("GX", "-"), f"chr{ref}", str(mark), qual]))
    return "\n".join(lines) + "\n"


@pytest.mark.skipif(not os.path.exists(oracle_io.REF_BIN), reason="compiled reference (oracle/_ref/dropest_ref) not built")
@pytest.mark.parametrize("merge,umi_merge,with_n,umi_len,with_quality", [("real", "simple", True, 8, True), ("none", "directional", False, 6, True),
                                                                         ("real", "directional", True, 6, True), ("simple", "simple", False, 8, True),
                                                                         ("real", "simple", False, 8, False)])
def test_reads_per_umi_per_cell_matches_the_compiled_reference(tmp_path, merge, umi_merge, with_n, umi_len, with_quality):
    reads, wl = _flow_case(seed=11 + umi_len + int(with_n), n_reads=24000, n_cells=10, umi_len=umi_len, with_n=with_n, whitelist=merge == "real")
    tsv = str(tmp_path / "reads.tsv")
    open(tsv, "w").write(_tsv(reads, umi_len, 5, with_quality))
    wl_path = "-"
    if wl:
        comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
        wl_path = str(tmp_path / "wl.txt")   # the reference reverse-complements whitelist tokens on load (BarcodesParser.cpp)
        open(wl_path, "w").write("\n".join(" ".join("".join(comp[c] for c in reversed(t)) for t in line.split()) for line in wl) + "\n")
    ora = oracle_io.run_oracle(tsv, kind="reference", merge=merge, barcodes=None if wl_path == "-" else wl_path, barcodes_type="const", min_genes_before=5,
                               min_genes_after=8, umi_merge=umi_merge, dump_rpupc=True)
    rds = str(tmp_path / "out.rds")
    o = subprocess.run([EXE, wl_path, merge, umi_merge, "5", "8", tsv, rds], capture_output=True, text=True)
    assert o.returncode == 0, o.stdout[-800:] + o.stderr[-800:]
    rows = [l.split("\t") for l in o.stdout.strip().split("\n")]
    cells = [r[1] for r in rows if r[0] == "cell"]
    genes = [r[1] for r in rows if r[0] == "gene"]
    entries = [(int(r[1]), int(r[2])) for r in rows if r[0] == "entry"]
    got = [dict() for _ in entries]
    for r in rows:
        if r[0] == "umi":
            got[int(r[1])][r[2]] = (int(r[3]), [float(x) for x in r[4].split(",")] if len(r) > 4 and r[4] else [])
    # ---- the reference
    assert cells == oracle_io.strings(ora["rp_cells"]) and genes == oracle_io.strings(ora["rp_genes"])
    assert entries == list(zip(ora["rp_cell_indexes"].tolist(), ora["rp_gene_indexes"].tolist()))
    exp = [dict() for _ in entries]
    qoff = np.concatenate([[0], np.cumsum(ora["rp_umi_qlen"])]).astype(np.int64)
    for k, (e, seq, n) in enumerate(zip(ora["rp_umi_entry"], oracle_io.strings(ora["rp_umi_seq"]), ora["rp_umi_reads"])):
        exp[int(e)][seq] = (int(n), ora["rp_umi_quality"][qoff[k]:qoff[k + 1]].tolist())
    assert got == exp
    n_umis = sum(len(e) for e in exp)
    assert n_umis > 700 and len(cells) >= 8
    if with_quality:
        assert all(len(q) == umi_len for e in exp for _, q in e.values())
    else:
        assert all(q == [] for e in exp for _, q in e.values())
    # ---- the .rds carries it as the last field: list(cells, genes, cell_indexes, gene_indexes, reads_per_umi)
    raw = gzip.open(rds).read()
    rd = Rds(raw)
    rd.o = 2
    assert rd.i32() == 2
    rd.i32(); rd.i32()
    d = rd.item()
    assert d["attr"]["names"][-1] == "reads_per_umi_per_cell"
    rp = d["v"][-1]
    assert rp["attr"]["names"] == ["cells", "genes", "cell_indexes", "gene_indexes", "reads_per_umi"]
    assert rp["v"][0] == cells and rp["v"][1] == genes and [int(x) for x in rp["v"][2]] == [e[0] for e in entries]
    first = rp["v"][4][0]
    assert set(first["attr"]["names"]) == set(exp[0]) and len(first["v"][0]) == 2
    name0 = first["attr"]["names"][0]
    assert first["v"][0][0] == [float(exp[0][name0][0])] and list(first["v"][0][1]) == exp[0][name0][1]
