"""Parity AT THE SHAPES THE BENCH NUMBERS ARE QUOTED ON (BASELINE.json configs[0..2]): the CUDA path (through the C ABI) against
the compiled, unmodified reference (oracle/_ref, kind="reference") on scaled replicas that keep the key layout, the whitelist,
the thresholds and the reads-per-cell of the full configuration.

  C2  10x v3: 16 + 12 bp, 30 k genes, the bench's own 2048 x 3328 product whitelist, default max_barcodes_hint
      (key layout tb=22 / gb=15 / ub=24), before=20 / after=100, 40 k reads per cell -> error barcodes become real cells and merge
  C3  inDrop v3: 8+8 / 6 bp, the reference's own data/barcodes/indrop_v3 (configs/indrop_v3.xml:6-8), barcodes_type=indrop
  C1  10x.xml plumbing case at FULL size: 1 M reads / 1 k cells / 500 genes on data/barcodes/10x_aug_2016_split (16 / 10 bp)
The whitelist files under tests/golden/ref_barcodes/ are copies of the reference's DATA files (not sources): /root/reference does
not exist on the GPU box.
"""
import os

import numpy as np
import pytest

import dropest_b200 as dg
from dropest_b200.synth import SynthSpec, product_whitelist, read_whitelist

import oracle_io
import parity_utils as pu

pytestmark = pytest.mark.gpu

REF_BARCODES = os.path.join(pu.GOLDEN, "ref_barcodes")


def _need_reference():
    if not oracle_io.available("reference"):
        pytest.skip("oracle/_ref (compiled reference) is not built")


def test_c2_shape_replica_20m_reads_500_cells(tmp_path):
    """The bench's C2 shape, 1/20 of its size with the same reads per cell."""
    _need_reference()
    wl_path = str(tmp_path / "wl_7x9_2048x3328.txt")
    product_whitelist(wl_path)
    wl = read_whitelist(wl_path)
    assert (len(wl[0]), len(wl[1])) == (2048, 3328)
    spec = SynthSpec(n_reads=20_000_000, n_cells=500, n_genes=30_000, cb_len=16, umi_len=12, whitelist_parts=wl, seed=43)
    case = pu.Case(name="c2_replica", spec=spec, cb_len=16, umi_len=12, n_genes=30_000, merge="real", barcodes=wl_path, barcodes_type="const",
                   min_genes_before=20, min_genes_after=100, max_cb_ed=2, min_frac=0.2, dump_umis=False, n_batches=4,
                   extra={"max_barcodes_hint": 0})
    res = pu.run_case(case, kind="reference")
    assert res["oracle"]["_kind"] == "reference"
    pu.assert_parity(res)
    s = res["gpu"]["summary"]
    # the replica must exercise what the bench exercises: error barcodes that became real cells and were merged back
    assert s["n_merged"] > 1000 and s["filtered_cells_number"] == 500
    assert s["real_cells_number"] == 500


def test_c3_shape_indrop_v3_reference_whitelist():
    """inDrop v3 on the reference's own whitelist file, two-part split (InDropBarcodesParser.cpp:31-38)."""
    _need_reference()
    wl_path = os.path.join(REF_BARCODES, "indrop_v3")
    wl = read_whitelist(wl_path, indrop=True)
    assert (len(wl[0]), len(wl[1])) == (384, 384)
    spec = SynthSpec(n_reads=5_000_000, n_cells=125, n_genes=30_000, cb_len=16, umi_len=6, whitelist_parts=wl, seed=44)
    case = pu.Case(name="c3_replica", spec=spec, cb_len=16, umi_len=6, n_genes=30_000, merge="real", barcodes=wl_path, barcodes_type="indrop",
                   min_genes_before=20, min_genes_after=100, max_cb_ed=2, min_frac=0.2, dump_umis=False, n_batches=3,
                   extra={"max_barcodes_hint": 0})
    res = pu.run_case(case, kind="reference")
    pu.assert_parity(res)
    assert res["gpu"]["summary"]["n_merged"] > 100


def test_c1_full_size_10x_aug_2016_split():
    """BASELINE configs[0] at full size through the CUDA path (configs/10x.xml: 16 bp CB, 10 bp UMI, before=20 / after=100)."""
    _need_reference()
    wl_path = os.path.join(REF_BARCODES, "10x_aug_2016_split")
    wl = read_whitelist(wl_path)
    assert (len(wl[0]), len(wl[1])) == (480, 1536)
    spec = SynthSpec(n_reads=1_000_000, n_cells=1000, n_genes=500, cb_len=16, umi_len=10, whitelist_parts=wl, seed=42)
    case = pu.Case(name="c1_full", spec=spec, cb_len=16, umi_len=10, n_genes=500, merge="real", barcodes=wl_path, barcodes_type="const",
                   min_genes_before=20, min_genes_after=100, max_cb_ed=2, min_frac=0.2, dump_umis=True, n_batches=3,
                   extra={"max_barcodes_hint": 0})
    res = pu.run_case(case, kind="reference")
    pu.assert_parity(res)
