"""Shared machinery of the parity tests: run one workload through the CUDA path (C ABI) and through the CPU oracle, compare.

The oracle (oracle/_ref = compiled reference, or oracle/_build = our restatement) is only ever the CHECKER here.
"""
from __future__ import annotations

import os
import tempfile
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

import dropest_b200 as dg
from dropest_b200.synth import SynthSpec, SynthTables, read_whitelist, write_packed

import oracle_io

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
WL_SYNTH_7_9 = os.path.join(GOLDEN, "wl_synth_7x9.txt")      # 96 x 128 tokens of 7 and 9 bp (10x-like split)
WL_SYNTH_8_8 = os.path.join(GOLDEN, "wl_synth_8x8.txt")      # 64 x 64 tokens of 8 bp (inDrop v3-like)
WL_CLOSE_4_4 = os.path.join(GOLDEN, "wl_close_4x4.txt")      # tiny, tokens at distance 1 of each other: exercises ties
WL_TEST_EST = os.path.join(GOLDEN, "wl_test_est.txt")        # the reference's own 3 x 3 inDrop fixture

MERGE_NAMES = {"none": dg.MERGE_NONE, "real": dg.MERGE_REAL, "simple": dg.MERGE_SIMPLE,
               "poisson_real": dg.MERGE_POISSON_REAL, "poisson_simple": dg.MERGE_POISSON_SIMPLE, "all": dg.MERGE_ALL}


@dataclass
class Case:
    name: str
    spec: Optional[SynthSpec] = None
    recs: Optional[np.ndarray] = None            # explicit records instead of a synthetic spec
    cb_len: int = 16
    umi_len: int = 12
    n_genes: int = 1
    gene_names: Optional[list] = None
    merge: str = "none"
    barcodes: Optional[str] = None
    barcodes_type: str = "const"
    min_genes_before: int = 10
    min_genes_after: int = 10
    max_cb_ed: int = 2
    min_frac: float = 0.2
    marks: str = "eEBA"
    max_cells: int = -1
    reads_output: bool = False
    umi_merge: str = "simple"
    max_umi_ed: int = 1
    umi_mult: float = 2.0
    max_merge_prob: float = 1e-4
    max_real_merge_prob: float = 1e-7
    dump_umis: bool = True
    chr_ids: Optional[np.ndarray] = None         # uint8 chromosome id per read -> per-chromosome Stats tables are fed and compared
    n_lists: Optional[object] = None             # dropest_b200.synth.NLists: the records carry barcodes / UMIs with N as indices into it
    n_batches: int = 3
    shuffle: bool = True
    extra: dict = field(default_factory=dict)


def small_case(n_reads=20000, n_cells=40, n_genes=60, merge="real", **kw) -> Case:
    wl = read_whitelist(WL_SYNTH_7_9)
    spec = SynthSpec(n_reads=n_reads, n_cells=n_cells, n_genes=n_genes, cb_len=16, umi_len=10, whitelist_parts=wl,
                     cb_error_ppm=kw.pop("cb_error_ppm", 60000), reads_per_umi=kw.pop("reads_per_umi", 3), seed=kw.pop("seed", 7))
    return Case(name=f"small_{n_reads}", spec=spec, cb_len=16, umi_len=10, n_genes=n_genes, merge=merge,
                barcodes=WL_SYNTH_7_9 if merge in ("real", "poisson_real") else None, barcodes_type="const",
                min_genes_before=kw.pop("min_genes_before", 5), min_genes_after=kw.pop("min_genes_after", 10), **kw)


def gpu_run(case: Case, recs: Optional[np.ndarray], device_generate: bool = False, tables: Optional[SynthTables] = None):
    cfg = dg.Config(cb_len=case.cb_len, umi_len=case.umi_len, n_genes=case.n_genes, merge_type=MERGE_NAMES[case.merge],
                    barcodes_type=dg.BARCODES_INDROP if case.barcodes_type == "indrop" else dg.BARCODES_CONST,
                    barcodes_file=case.barcodes, min_genes_before_merge=case.min_genes_before,
                    min_genes_after_merge=case.min_genes_after, max_cb_merge_edit_distance=case.max_cb_ed,
                    min_merge_fraction=case.min_frac, marks=case.marks, max_cells=case.max_cells,
                    umi_merge_type=dg.UMI_MERGE_DIRECTIONAL if case.umi_merge == "directional" else dg.UMI_MERGE_SIMPLE,
                    max_umi_merge_edit_distance=case.max_umi_ed, umi_merge_mult=case.umi_mult,
                    max_merge_prob=case.max_merge_prob, max_real_merge_prob=case.max_real_merge_prob,
                    reads_output=case.reads_output, max_barcodes_hint=case.extra.get("max_barcodes_hint", 1 << 16),
                    allow_n=case.n_lists is not None)
    c = dg.Container(cfg)
    if case.n_lists is not None:
        c.set_n_strings(0, case.n_lists.umis)
        c.set_n_strings(1, case.n_lists.cbs)
    if device_generate:
        import torch

        n = case.spec.n_reads
        buf = torch.empty(n * 16, dtype=torch.uint8, device="cuda:0")
        tables.generate_device(0, 0, n, buf.data_ptr())
        torch.cuda.synchronize()
        c.add_batch_device(buf.data_ptr(), n, keepalive=buf)
    elif case.extra.get("segments"):
        # ONE fill launch over several device arrays of very different lengths (dge_add_batch_segments_device), empty ones included
        import torch

        cuts = [int(recs.shape[0] * f) for f in case.extra["segments"]]
        bounds = [0] + sorted(cuts) + [recs.shape[0]]
        perm = np.random.default_rng(7).permutation(recs.shape[0]) if case.shuffle else np.arange(recs.shape[0])
        bufs = [torch.from_numpy(np.ascontiguousarray(recs[perm[a:b]]).view(np.uint8).reshape(-1).copy()).cuda() for a, b in zip(bounds[:-1], bounds[1:])]
        bufs = [b if b.numel() else torch.empty(16, dtype=torch.uint8, device="cuda:0") for b in bufs]
        c.add_batch_segments_device([b.data_ptr() for b in bufs], [b1 - a1 for a1, b1 in zip(bounds[:-1], bounds[1:])], keepalive=bufs)
    elif case.extra.get("soa"):
        # structure-of-arrays batches in stream order: read_idx is implicit (dge_add_batch_soa)
        assert np.array_equal(recs["read_idx"], recs["read_idx"][0] + np.arange(recs.shape[0], dtype=np.uint32))
        bounds = np.linspace(0, recs.shape[0], max(1, case.n_batches) + 1).astype(np.int64)
        for a, b in zip(bounds[:-1], bounds[1:]):
            if b > a:
                c.add_batch_soa(recs["key"][a:b], recs["gene"][a:b], first_read_idx=int(recs["read_idx"][a]))
    else:
        order = np.arange(recs.shape[0])
        if case.shuffle:
            np.random.default_rng(123).shuffle(order)
        for k, part in enumerate(np.array_split(order, max(1, case.n_batches))):
            if case.chr_ids is None:
                c.add_batch(recs[part])
            elif k % 2 == 0 or not np.array_equal(recs["read_idx"][part], recs["read_idx"][part[0]] + np.arange(part.shape[0], dtype=np.uint32)):
                c.add_batch_chr(recs[part], case.chr_ids[part])
            else:   # consecutive reads: the structure-of-arrays form
                c.add_batch_soa_chr(recs["key"][part], recs["gene"][part], case.chr_ids[part], first_read_idx=int(recs["read_idx"][part[0]]))
    c.set_initialized()
    pre = c.cells(dg.CELLS_FILTERED)
    c.merge_and_filter()
    out = {
        "summary": c.summary(),
        "timings": c.timings(),
        "filtered_pre": pre,
        "all": c.cells(dg.CELLS_ALL),
        "real": c.cells(dg.CELLS_REAL),
        "filtered": c.cells(dg.CELLS_FILTERED),
        "cm": c.matrix(dg.MATRIX_CM),
        "cm_raw": c.matrix(dg.MATRIX_CM_RAW),
        "gene_order": c.gene_order(),
        "merge_pairs": c.merge_pairs(),
    }
    if case.chr_ids is not None:
        out["chr_stats"] = c.chr_stats()
    if case.dump_umis:
        out["umigs"] = c.umigs(dg.CELLS_ALL)
    if case.extra.get("matrix_marks"):   # the filtered matrix for other query marks (the -V matrices)
        out["cm_marks"] = {code: c.matrix_marks(code) for code in case.extra["matrix_marks"]}
    c.close()
    return out


def run_case(case: Case, device_generate: bool = False, kind: str = "any"):
    tables = None
    if case.spec is not None:
        tables = SynthTables(case.spec)
        recs = tables.generate_host(0, case.spec.n_reads)
    else:
        recs = case.recs
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "reads.bin")
        if case.extra.get("n_chr") and case.chr_ids is None:
            case.chr_ids = synth_chr_ids(recs, int(case.extra["n_chr"]))
        write_packed(path, recs, case.cb_len, case.umi_len, case.n_genes, case.gene_names, n_lists=case.n_lists, chr_ids=case.chr_ids)
        ora = oracle_io.run_oracle(path, kind=kind, merge=case.merge, barcodes=case.barcodes, barcodes_type=case.barcodes_type,
                                   min_genes_before=case.min_genes_before, min_genes_after=case.min_genes_after,
                                   max_cb_ed=case.max_cb_ed, min_frac=case.min_frac, marks=case.marks, max_cells=case.max_cells,
                                   reads_output=case.reads_output, dump_umis=case.dump_umis, umi_merge=case.umi_merge,
                                   max_umi_ed=case.max_umi_ed, umi_mult=case.umi_mult, max_merge_prob=case.max_merge_prob,
                                   max_real_merge_prob=case.max_real_merge_prob)
    gpu = gpu_run(case, recs, device_generate=device_generate, tables=tables)
    return {"case": case, "oracle": ora, "gpu": gpu, "recs": recs}


def synth_chr_ids(recs: np.ndarray, n_chr: int) -> np.ndarray:
    """A chromosome per read: genes live on one chromosome each, reads without a gene fall anywhere; the last chromosome only ever
    sees intergenic reads (the per-statistic column sets differ, Stats::presented_chromosomes)."""
    gene = (recs["gene"] & np.uint32(0xFFFFFF)).astype(np.uint64)
    by_gene = ((gene * np.uint64(2654435761)) >> np.uint64(13)) % np.uint64(max(1, n_chr - 1))
    anywhere = ((recs["read_idx"].astype(np.uint64) * np.uint64(40503)) >> np.uint64(3)) % np.uint64(n_chr)
    return np.where(gene == 0xFFFFFF, anywhere, by_gene).astype(np.uint8)


def _gene_id_of_name(case: Case, names):
    if case.gene_names:
        lut = {n: i for i, n in enumerate(case.gene_names)}
        return np.array([lut[n] for n in names], dtype=np.int64)
    return np.array([int(n[1:]) for n in names], dtype=np.int64)


def assert_parity(res, check_umigs: bool = True):
    case, ora, gpu = res["case"], res["oracle"], res["gpu"]
    # ---- cells, in first-seen order
    pack_cb = case.n_lists.pack_cb if case.n_lists is not None else dg.pack_seq
    pack_umi = case.n_lists.pack_umi if case.n_lists is not None else dg.pack_seq
    o_bc = np.array([pack_cb(s) for s in oracle_io.strings(ora["cell_barcodes"])], dtype=np.uint64)
    g_all = gpu["all"]
    assert g_all.shape[0] == o_bc.shape[0] == int(ora["n_cells"][0]) == gpu["summary"]["total_cells_number"], "total cells"
    np.testing.assert_array_equal(g_all["barcode"], o_bc, err_msg="cell id order (first seen)")
    np.testing.assert_array_equal(g_all["flags"], ora["cell_flags"].astype(np.uint32), err_msg="real/merged/excluded flags")
    np.testing.assert_array_equal(g_all["umis_stat"], ora["cell_umis_stat"], err_msg="TOTAL_UMIS_PER_CB")
    np.testing.assert_array_equal(g_all["reads_stat"], ora["cell_reads_stat"], err_msg="TOTAL_READS_PER_CB")
    np.testing.assert_array_equal(g_all["n_genes"], ora["cell_n_genes"], err_msg="genes per cell")
    np.testing.assert_array_equal(g_all["requested_genes_num"], ora["cell_req_genes"].astype(np.int32), err_msg="requested genes")
    np.testing.assert_array_equal(g_all["requested_umis_num"], ora["cell_req_umis"].astype(np.int32), err_msg="requested umis")
    np.testing.assert_array_equal(g_all["merge_target"].astype(np.int64), ora["merge_targets"], err_msg="merge_targets")
    # ---- filtered cells (order matters), before and after the merge
    np.testing.assert_array_equal(gpu["filtered_pre"]["barcode"], o_bc[ora["filtered_pre_merge"]], err_msg="filtered cells at set_initialized")
    np.testing.assert_array_equal(gpu["filtered"]["barcode"], o_bc[ora["filtered_cells"]], err_msg="filtered cells")
    assert gpu["summary"]["real_cells_number"] == int(ora["real_cells_number"][0])
    for k in ("intergenic_reads", "has_exon_reads", "has_intron_reads", "has_not_annotated_reads"):
        assert gpu["summary"][k] == int(ora[k][0]), k
    # ---- gene first-seen order
    o_gene_ids = _gene_id_of_name(case, oracle_io.strings(ora["gene_names"]))
    np.testing.assert_array_equal(gpu["gene_order"].astype(np.int64), o_gene_ids, err_msg="gene indexer order")
    # ---- matrices: (column, gene, value) with genes ascending inside a column
    for name, pref in (("cm", "cm"), ("cm_raw", "cm_raw")):
        indptr, genes, vals = gpu[name]
        col = np.repeat(np.arange(indptr.shape[0] - 1), np.diff(indptr))
        o_col, o_gene, o_val = ora[pref + "_col"], o_gene_ids[ora[pref + "_gene"]] if ora[pref + "_gene"].size else ora[pref + "_gene"], ora[pref + "_val"]
        # oracle triplets are ordered by gene_indexer id inside a column; ours by caller gene id
        o_order = np.lexsort((o_gene, o_col))
        assert col.shape[0] == o_col.shape[0], f"{name} nnz {col.shape[0]} vs {o_col.shape[0]}"
        np.testing.assert_array_equal(col, o_col[o_order], err_msg=f"{name} columns")
        np.testing.assert_array_equal(genes.astype(np.int64), o_gene[o_order], err_msg=f"{name} genes")
        np.testing.assert_array_equal(vals.astype(np.int64), o_val[o_order], err_msg=f"{name} values")
    # ---- per-chromosome Stats tables: rows = real cells that counted anything for the statistic (cell-id order), columns = the
    # chromosomes the statistic has seen in ANY cell; compared as {(barcode, chromosome id): count} + the row list + the column set
    if "chr_stats" in gpu:
        counts, presented = gpu["chr_stats"]
        real_bc = gpu["real"]["barcode"]
        assert counts.shape[0] == real_bc.shape[0]
        for t, name in enumerate(("chr_exon", "chr_intron", "chr_intergenic")):
            o_cells = np.array([pack_cb(x) for x in oracle_io.strings(ora[name + "_cells"])], dtype=np.uint64)
            o_chrs = np.array([int(x[3:]) for x in oracle_io.strings(ora[name + "_chrs"])], dtype=np.int64)
            o_counts = ora[name + "_counts"].reshape(o_cells.shape[0], o_chrs.shape[0]) if o_cells.size else np.zeros((0, o_chrs.shape[0]), dtype=np.int32)
            mine = counts[:, :, t]
            rows = np.flatnonzero(mine.sum(axis=1) > 0)
            np.testing.assert_array_equal(real_bc[rows], o_cells, err_msg=f"{name}: cells listed")
            assert sorted(np.flatnonzero(presented[t]).tolist()) == sorted(o_chrs.tolist()), f"{name}: presented chromosomes"
            np.testing.assert_array_equal(mine[rows][:, o_chrs], o_counts, err_msg=f"{name}: counts")
    n_cols_raw = gpu["cm_raw"][0].shape[0] - 1
    assert n_cols_raw == ora["cm_raw_cells"].shape[0]
    np.testing.assert_array_equal(gpu["real"]["barcode"], o_bc[ora["cm_raw_cells"]], err_msg="cm_raw column order")
    # ---- every (cell, gene, UMI, reads, mark)
    if check_umigs and case.dump_umis and "umi_cell" in ora:
        o_umi = np.array([pack_umi(s) for s in oracle_io.strings(ora["umi_seq"])], dtype=np.uint64)
        o = np.stack([ora["umi_cell"].astype(np.uint64), o_gene_ids[ora["umi_gene"]].astype(np.uint64), o_umi,
                      ora["umi_count"].astype(np.uint64), ora["umi_mark"].astype(np.uint64)], axis=1)
        u = gpu["umigs"]
        g = np.stack([u["cell"].astype(np.uint64), u["gene"].astype(np.uint64), u["umi"].astype(np.uint64),
                      u["count"].astype(np.uint64), u["mark"].astype(np.uint64)], axis=1)
        o = o[np.lexsort((o[:, 2], o[:, 1], o[:, 0]))]
        g = g[np.lexsort((g[:, 2], g[:, 1], g[:, 0]))]
        assert o.shape == g.shape, f"distinct (cell,gene,umi): {g.shape[0]} vs oracle {o.shape[0]}"
        np.testing.assert_array_equal(g, o, err_msg="(cell, gene, umi, reads, mark)")
