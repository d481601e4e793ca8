"""Cross-rank whitelist merge (dge_dist_step) with every 'rank' a handle on ONE GPU: the collectives are plain copies
(dropest_b200.dist.merge_across_handles), everything else is the code that runs under NCCL.  The union of the shards' results
must equal the single-handle result bit for bit -- including children whose nearest candidates are in farther distance classes
and order-dependent ties of the best fraction (RealBarcodesMergeStrategy.cpp:63-109), with the candidate on another rank."""
import numpy as np
import pytest

import dropest_b200 as dg
from dropest_b200 import dist as dgdist
from dropest_b200.synth import SynthSpec, SynthTables, rank_of, read_whitelist, records_from_strings

import parity_utils as pu

pytestmark = pytest.mark.gpu


def _cm_triplets(c, which_cells, which_matrix):
    cells = c.cells(which_cells)
    indptr, genes, vals = c.matrix(which_matrix)
    col = np.repeat(np.arange(indptr.shape[0] - 1), np.diff(indptr))
    return np.stack([cells["barcode"][col].astype(np.uint64), genes.astype(np.uint64), vals.astype(np.uint64)], axis=1)


def _order(t):
    return t[np.lexsort(tuple(t[:, k] for k in reversed(range(t.shape[1]))))] if t.shape[0] else t


def _collect(c):
    filt = c.cells(dg.CELLS_FILTERED)
    a, b = c.merge_pairs()
    allc = c.cells(dg.CELLS_ALL)
    return {
        "cm": _cm_triplets(c, dg.CELLS_FILTERED, dg.MATRIX_CM),
        "raw": _cm_triplets(c, dg.CELLS_REAL, dg.MATRIX_CM_RAW),
        "pairs": np.stack([a, b], axis=1) if a.shape[0] else np.zeros((0, 2), dtype=np.uint64),
        "filt": np.stack([filt["barcode"], filt["umis_stat"].astype(np.uint64), filt["reads_stat"].astype(np.uint64),
                          filt["requested_genes_num"].astype(np.uint64), filt["requested_umis_num"].astype(np.uint64)], axis=1),
        "cells": np.stack([allc["barcode"], allc["flags"].astype(np.uint64), allc["umis_stat"].astype(np.uint64), allc["reads_stat"].astype(np.uint64)], axis=1),
        "summary": c.summary(),
    }


def _sync_umi_first(conts):
    """what dropest_b200.dist.sync_umi_first_seen does with an all-reduce(min), for handles of one process"""
    import torch

    n = conts[0].umi_first_size()
    if n == 0:
        return
    tabs = []
    for c in conts:
        t = torch.empty(n, dtype=torch.int32, device="cuda:0")
        c.umi_first_export(t.data_ptr())
        tabs.append(t.to(torch.int64) & 0xFFFFFFFF)
    m = torch.stack(tabs).min(dim=0).values
    m = torch.where(m >= 2 ** 31, m - 2 ** 32, m).to(torch.int32)
    torch.cuda.synchronize()
    for c in conts:
        c.umi_first_import(m.data_ptr())


def _run(cfg_kwargs, recs, world):
    single = dg.Container(dg.Config(**cfg_kwargs))
    single.add_batch(recs)
    single.set_initialized()
    single.merge_and_filter()
    ref = _collect(single)
    single.close()
    owner = rank_of((recs["key"] >> np.uint64(24)).astype(np.uint64), world)
    conts = []
    for r in range(world):
        c = dg.Container(dg.Config(sharded=True, **cfg_kwargs))
        c.add_batch(recs[owner == r])
        conts.append(c)
    _sync_umi_first(conts)
    for c in conts:
        c.set_initialized()
    dgdist.merge_across_handles(conts)
    outs = []
    for c in conts:
        c.merge_and_filter()
        outs.append(_collect(c))
        c.close()
    for name in ("cm", "raw", "pairs", "filt", "cells"):
        got = np.concatenate([o[name] for o in outs])
        np.testing.assert_array_equal(_order(got), _order(ref[name]), err_msg=name)
    for k in ("n_merged", "n_excluded", "real_cells_number", "filtered_cells_number", "cm_nnz", "cm_raw_nnz", "total_cells_number"):
        assert sum(o["summary"][k] for o in outs) == ref["summary"][k], k
    assert all(o["summary"]["n_unresolved"] == 0 for o in outs)
    return ref, outs


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_whitelist_merge_equals_single_handle(world):
    wl = read_whitelist(pu.WL_SYNTH_7_9)
    spec = SynthSpec(n_reads=300_000, n_cells=80, n_genes=150, cb_len=16, umi_len=10, whitelist_parts=wl, cb_error_ppm=60000, seed=9)
    recs = SynthTables(spec).generate_host(0, spec.n_reads)
    cfg = dict(cb_len=16, umi_len=10, n_genes=150, merge_type=dg.MERGE_REAL, barcodes_type=dg.BARCODES_CONST, barcodes_file=pu.WL_SYNTH_7_9,
               min_genes_before_merge=5, min_genes_after_merge=10, max_barcodes_hint=1 << 16)
    ref, outs = _run(cfg, recs, world)
    assert ref["summary"]["n_merged"] > 0
    # the merges really crossed ranks
    assert sum(o["pairs"].shape[0] for o in outs) == ref["pairs"].shape[0]


@pytest.mark.parametrize("min_frac", [0.0, 0.2])
@pytest.mark.parametrize("world", [2, 5])
def test_sharded_merge_far_classes_and_ties(world, min_frac):
    """Whitelist tokens one substitution apart: several candidates per class (exact ties of the best fraction, decided by the
    reference's neighbour order) and children with two errors (distance classes >= 2), candidates spread over the ranks."""
    rng = np.random.default_rng(5)
    wl = read_whitelist(pu.WL_CLOSE_4_4)
    true_cbs = [a + b for a in wl[0] for b in wl[1]][:20]
    reads = []
    genes = [f"G{i}" for i in range(12)]
    for cb in true_cbs:
        for _ in range(int(rng.integers(5, 40))):
            c = list(cb)
            r = rng.random()
            if r < 0.25:
                p = int(rng.integers(0, 8)); c[p] = "ACGT"[(("ACGT".index(c[p])) + int(rng.integers(1, 4))) % 4]
            if r < 0.08:
                p = int(rng.integers(0, 8)); c[p] = "ACGT"[(("ACGT".index(c[p])) + int(rng.integers(1, 4))) % 4]
            umi = "".join("ACGT"[int(x)] for x in rng.integers(0, 2, size=4))
            reads.append(("".join(c), umi, genes[int(rng.integers(0, 12))], int(rng.choice([1, 2, 4, 6]))))
    gene_ids = {}
    recs = records_from_strings(reads, gene_ids)
    cfg = dict(cb_len=8, umi_len=4, n_genes=len(gene_ids), merge_type=dg.MERGE_REAL, barcodes_type=dg.BARCODES_CONST, barcodes_file=pu.WL_CLOSE_4_4,
               min_genes_before_merge=1, min_genes_after_merge=2, min_merge_fraction=min_frac, max_barcodes_hint=1 << 12)
    ref, outs = _run(cfg, recs, world)
    assert ref["summary"]["n_merged"] > 0


def test_sharded_merge_then_directional_umi_merge():
    """-u on sharded handles: the cross-rank CB merge runs on the device, the UMI merge continues from the merged state; the UMI
    first-seen table (StringIndexer ids) is min-reduced across the shards first."""
    wl = read_whitelist(pu.WL_SYNTH_7_9)
    spec = SynthSpec(n_reads=120_000, n_cells=25, n_genes=20, cb_len=16, umi_len=5, whitelist_parts=wl, cb_error_ppm=50000, reads_per_umi=3, seed=16)
    recs = SynthTables(spec).generate_host(0, spec.n_reads)
    cfg = dict(cb_len=16, umi_len=5, n_genes=20, merge_type=dg.MERGE_REAL, barcodes_type=dg.BARCODES_CONST, barcodes_file=pu.WL_SYNTH_7_9,
               min_genes_before_merge=3, min_genes_after_merge=5, max_barcodes_hint=1 << 16, umi_merge_type=dg.UMI_MERGE_DIRECTIONAL)
    ref, outs = _run(cfg, recs, 3)
    assert ref["summary"]["n_umis_merged"] > 0 and ref["summary"]["n_merged"] > 0
    assert sum(o["summary"]["n_umis_merged"] for o in outs) == ref["summary"]["n_umis_merged"]


def test_sharded_merge_with_empty_shards():
    """More ranks than cells: some shards hold no read at all and still take part in every collective."""
    wl = read_whitelist(pu.WL_SYNTH_7_9)
    spec = SynthSpec(n_reads=20_000, n_cells=3, n_genes=40, cb_len=16, umi_len=10, whitelist_parts=wl, cb_error_ppm=100000, seed=4)
    recs = SynthTables(spec).generate_host(0, spec.n_reads)
    cfg = dict(cb_len=16, umi_len=10, n_genes=40, merge_type=dg.MERGE_REAL, barcodes_type=dg.BARCODES_CONST, barcodes_file=pu.WL_SYNTH_7_9,
               min_genes_before_merge=3, min_genes_after_merge=5, max_barcodes_hint=1 << 12)
    _run(cfg, recs, 8)
