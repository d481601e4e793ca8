"""Reads whose barcode / UMI contains N: the CUDA path keeps them like the reference does -- the barcode is a cell of its own,
the UMI a UMI of its own until MergeUMIsStrategySimple repairs the N-UMIs of the real cells (nearest N-free UMI of the same (cell, gene)
within max_umi_merge_edit_distance, ties: more reads, then UMI id; else the N's are replaced with the process-wide rand() seeded 42,
consumed in the iteration order of the reference's unordered_set) -- checked field by field against the compiled reference."""
import numpy as np
import pytest

import dropest_b200 as dg
from dropest_b200.synth import SynthSpec, SynthTables, inject_n, read_whitelist

import oracle_io
import parity_utils as pu

pytestmark = pytest.mark.gpu


def _case(merge, umi_len, n_genes, n_reads, n_cells, seed, umi_ppm, cb_ppm, **kw):
    wl = read_whitelist(pu.WL_SYNTH_7_9)
    spec = SynthSpec(n_reads=n_reads, n_cells=n_cells, n_genes=n_genes, cb_len=16, umi_len=umi_len, whitelist_parts=wl, cb_error_ppm=50000,
                     reads_per_umi=kw.pop("reads_per_umi", 3), seed=seed)
    recs = SynthTables(spec).generate_host(0, n_reads)
    recs, lists = inject_n(recs, 16, umi_len, umi_ppm, cb_ppm, seed=seed)
    return pu.Case(name=f"n_reads_{merge}_{umi_len}", recs=recs, cb_len=16, umi_len=umi_len, n_genes=n_genes, merge=merge,
                   barcodes=pu.WL_SYNTH_7_9 if merge == "real" else None, barcodes_type="const", min_genes_before=kw.pop("min_genes_before", 5),
                   min_genes_after=kw.pop("min_genes_after", 10), n_lists=lists, **kw), lists


@pytest.mark.parametrize("merge", ["none", "real"])
@pytest.mark.parametrize("umi_len,n_genes,max_umi_ed", [(10, 120, 1), (5, 15, 1), (6, 30, 2), (12, 60, 0)])
def test_n_umis_are_repaired_like_the_reference(merge, umi_len, n_genes, max_umi_ed):
    """3 % of the reads get one or two N's in the UMI.  Short UMIs / few genes make dense segments where a nearest N-free UMI exists
    (ties on distance and read count included), long ones mostly take the random fill; both orders come from the reference's containers."""
    if not oracle_io.available("reference"):
        pytest.skip("oracle/_ref (compiled reference) is not built")
    case, lists = _case(merge, umi_len, n_genes, 60000, 30, seed=40 + umi_len, umi_ppm=30000, cb_ppm=0, max_umi_ed=max_umi_ed)
    res = pu.run_case(case, kind="reference")
    pu.assert_parity(res)
    assert res["gpu"]["summary"]["n_umis_merged"] > 0
    u = res["gpu"]["umigs"]
    real_cells = np.flatnonzero(res["gpu"]["all"]["flags"] & 1)
    assert not np.any((u["umi"][np.isin(u["cell"], real_cells)] & dg.UMI_N_BIT) != 0), "an N-UMI survived in a real cell"


@pytest.mark.parametrize("merge", ["none", "real"])
@pytest.mark.parametrize("umi_len,n_genes,max_umi_ed,reads_per_umi", [(10, 120, 1, 3), (5, 15, 1, 6), (6, 30, 2, 5), (8, 25, 1, 8)])
def test_n_umis_under_the_directional_umi_merge(merge, umi_len, n_genes, max_umi_ed, reads_per_umi):
    """-u with N-UMIs (MergeUMIsStrategyDirectional.cpp:57-116): a source with N only stops at distance 0, and without any target is renamed by
    fix_n_umi_with_random from the never-seeded rand(); targets travel through the unordered_map's one-hop compression and are applied in the
    map's order (Cell::merge_umis).  Segments holding an N-UMI are replayed literally on the host, in the reference's traversal order."""
    if not oracle_io.available("reference"):
        pytest.skip("oracle/_ref (compiled reference) is not built")
    case, lists = _case(merge, umi_len, n_genes, 60000, 30, seed=70 + umi_len, umi_ppm=30000, cb_ppm=0, max_umi_ed=max_umi_ed,
                        umi_merge="directional", reads_per_umi=reads_per_umi)
    res = pu.run_case(case, kind="reference")
    pu.assert_parity(res)
    assert res["gpu"]["summary"]["n_umis_merged"] > 100


def test_n_barcodes_are_cells_of_their_own_and_merge_through_the_whitelist_walk():
    """1 % of the reads get an N in the barcode: those barcodes are cells of their own; the ones that become real take the exact host
    enumeration (N matches any base, BarcodesParser.cpp:21-74) and merge into their whitelist neighbour."""
    if not oracle_io.available("reference"):
        pytest.skip("oracle/_ref (compiled reference) is not built")
    case, lists = _case("real", 10, 80, 80000, 12, seed=51, umi_ppm=5000, cb_ppm=10000, min_genes_before=2, min_genes_after=5)
    res = pu.run_case(case, kind="reference")
    pu.assert_parity(res)
    allc = res["gpu"]["all"]
    is_n = (allc["barcode"] & np.uint64(dg.CB_N_BIT)) != 0
    assert is_n.sum() > 50
    assert np.any(is_n & ((allc["flags"] & 2) != 0)), "no barcode with N was merged: the test does not exercise the host walk"


@pytest.mark.parametrize("wl,btype", [(pu.WL_SYNTH_8_8, "indrop"), (pu.WL_SYNTH_7_9, "const")])
@pytest.mark.parametrize("merge", ["none", "real"])
def test_variable_length_barcodes_travel_as_escaped_cells(merge, wl, btype):
    """inDrop v1 / v2: barcodes of several lengths in one run (a9).  A quarter of the barcodes lose or gain a base; they are cells of their
    own (dge_set_cb_strings), and with a whitelist their merge runs through the exact host walk on the strings (InDropBarcodesParser's
    split takes the last len(part 2) characters, ConstLengthBarcodesParser fixed offsets) -- against the compiled reference."""
    from dropest_b200.synth import inject_odd_length_barcodes

    if not oracle_io.available("reference"):
        pytest.skip("oracle/_ref (compiled reference) is not built")
    parts = read_whitelist(wl, indrop=btype == "indrop")
    spec = SynthSpec(n_reads=60000, n_cells=25, n_genes=80, cb_len=16, umi_len=10, whitelist_parts=parts, cb_error_ppm=40000, seed=61)
    recs = SynthTables(spec).generate_host(0, spec.n_reads)
    recs, lists = inject_n(recs, 16, 10, 3000, 3000, seed=5)
    recs, lists = inject_odd_length_barcodes(recs, 16, 0.25, lists, seed=6)
    case = pu.Case(name="varlen", recs=recs, cb_len=16, umi_len=10, n_genes=80, merge=merge, barcodes=wl if merge == "real" else None, barcodes_type=btype,
                   min_genes_before=3, min_genes_after=6, n_lists=lists)
    if merge == "real" and btype == "const":
        # ConstLengthBarcodesParser::split_barcode (.cpp:34-48) refuses a barcode of another length: the reference run fails, and so does ours,
        # with the same message
        with pytest.raises(RuntimeError, match="oracle failed"):
            pu.run_case(case, kind="reference")
        with pytest.raises(dg.DgeError, match="has wrong length \\(16 expected\\)"):
            pu.gpu_run(case, recs)
        return
    res = pu.run_case(case, kind="reference")
    pu.assert_parity(res)
    lens = {len(s) for s in lists.cbs}
    assert {15, 16, 17} <= lens
    allc = res["gpu"]["all"]
    esc = (allc["barcode"] & np.uint64(dg.CB_N_BIT)) != 0
    assert esc.sum() > 100
    if merge == "real":
        assert np.any(esc & ((allc["flags"] & 2) != 0)), "no variable-length barcode was merged"
    else:
        assert np.any(esc & ((allc["flags"] & 1) != 0)), "no variable-length barcode is a real cell"


def test_flagged_records_need_allow_n():
    recs = np.zeros(1, dtype=dg.RECORD_DTYPE)
    recs["key"] = [(0x1234 << 24) | 3]
    recs["gene"] = [1 | (2 << 24) | dg.FLAG_UMI_N]
    c = dg.Container(dg.Config(cb_len=8, umi_len=4, n_genes=2))
    c.add_batch(recs)
    with pytest.raises(dg.DgeError) as e:
        c.set_initialized()
    assert e.value.code == 1
    c.close()
