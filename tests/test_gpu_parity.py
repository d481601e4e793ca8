"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.  Bit-exact bar."""
import numpy as np
import pytest

import dropest_b200 as dg
from dropest_b200.synth import SynthSpec, SynthTables, read_whitelist, records_from_strings

import oracle_io
import parity_utils as pu

pytestmark = pytest.mark.gpu

import golden_cases
from golden_cases import fixture_case


@pytest.mark.parametrize("name", ["fixture", "real_7x9", "none_7x9", "simple_7x9", "real_8x8_reads", "poisson_simple_7x9", "poisson_real_7x9", "all_7x9",
                                  "directional_7x9", "real_7x9_chr"])
def test_gpu_reproduces_golden_reference_outputs(name):
    """CUDA path vs the committed outputs of the compiled, unmodified reference (no oracle binary needed on the GPU box)."""
    case = golden_cases.cases()[name]
    recs = golden_cases.case_records(case)
    golden_cases.case_chr_ids(case, recs)
    gpu = pu.gpu_run(case, recs)
    pu.assert_parity({"case": case, "oracle": golden_cases.load_golden(name), "gpu": gpu})


def test_synth_device_matches_host():
    import torch

    spec = SynthSpec(n_reads=200_000, n_cells=300, n_genes=500, cb_len=16, umi_len=12, whitelist_parts=read_whitelist(pu.WL_SYNTH_7_9))
    t = SynthTables(spec)
    host = t.generate_host(1000, 100_000)
    buf = torch.empty(100_000 * 16, dtype=torch.uint8, device="cuda:0")
    t.generate_device(0, 1000, 100_000, buf.data_ptr())
    dev = np.frombuffer(buf.cpu().numpy().tobytes(), dtype=dg.RECORD_DTYPE)
    np.testing.assert_array_equal(dev, host)


def test_reference_fixture_merge_by_real_barcodes():
    """Tests/TestEstimation.cpp:237-280 (testMergeByRealBarcodes) through the CUDA path, checked against the oracle AND the
    literal expectations written in the reference test."""
    res = pu.run_case(fixture_case())
    pu.assert_parity(res)
    g = res["gpu"]
    assert g["summary"]["total_cells_number"] == 7
    assert g["filtered"].shape[0] == 2
    assert list(g["filtered"]["n_genes"]) == [3, 4]
    assert list(g["all"]["merge_target"]) == [0, 1, 1, 0, 0, 0, 6]
    assert [bool(f & 2) for f in g["all"]["flags"]] == [False, False, True, True, True, True, False]
    assert sum(bool(f & 4) for f in g["all"]["flags"]) == 1
    u = g["umigs"]
    names = res["case"].gene_names

    def reads(cell, gene, umi):
        m = (u["cell"] == cell) & (u["gene"] == names.index(gene)) & (u["umi"] == dg.pack_seq(umi))
        assert m.sum() == 1
        return int(u["count"][m][0])

    c0, c1 = 1, 0  # filtered order is [cell 1, cell 0]
    assert reads(c0, "Gene1", "CAACCT") == 2
    assert reads(c1, "Gene1", "AAACCT") == 3
    assert reads(c1, "Gene2", "CCCCCT") == 4
    assert reads(c1, "Gene3", "ACCCCT") == 2
    assert reads(c1, "Gene3", "CCATTC") == 1


def test_fill_only_small():
    res = pu.run_case(pu.small_case(n_reads=30000, n_cells=50, n_genes=80, merge="none"))
    pu.assert_parity(res)


def test_fill_device_resident_input():
    res = pu.run_case(pu.small_case(n_reads=50000, n_cells=60, n_genes=100, merge="none"), device_generate=True)
    pu.assert_parity(res)


@pytest.mark.parametrize("seed", [7, 8])
def test_merge_real_small(seed):
    res = pu.run_case(pu.small_case(n_reads=60000, n_cells=30, n_genes=120, merge="real", seed=seed))
    pu.assert_parity(res)
    assert res["gpu"]["summary"]["n_merged"] > 0


def test_merge_real_medium():
    res = pu.run_case(pu.small_case(n_reads=600_000, n_cells=400, n_genes=2000, merge="real", min_genes_before=20, min_genes_after=50,
                                    cb_error_ppm=20000, reads_per_umi=4))
    pu.assert_parity(res)


def test_merge_real_indrop_like_8x8():
    wl = read_whitelist(pu.WL_SYNTH_8_8)
    spec = SynthSpec(n_reads=80000, n_cells=40, n_genes=150, cb_len=16, umi_len=6, whitelist_parts=wl, cb_error_ppm=80000, seed=3)
    case = pu.Case(name="indrop_like", spec=spec, cb_len=16, umi_len=6, n_genes=150, merge="real", barcodes=pu.WL_SYNTH_8_8,
                   barcodes_type="indrop", min_genes_before=5, min_genes_after=10)
    res = pu.run_case(case)
    pu.assert_parity(res)


@pytest.mark.parametrize("min_frac", [0.0, 0.2])
def test_close_whitelist_ties_and_far_classes(min_frac):
    """Whitelist tokens one substitution apart: several neighbours per class, exact ties, distance classes >= 2."""
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
    names = [n for n, _ in sorted(gene_ids.items(), key=lambda kv: kv[1])]
    case = pu.Case(name="close_wl", recs=recs, cb_len=8, umi_len=4, n_genes=len(names), gene_names=names, merge="real",
                   barcodes=pu.WL_CLOSE_4_4, barcodes_type="const", min_genes_before=1, min_genes_after=2, min_frac=min_frac)
    res = pu.run_case(case)
    pu.assert_parity(res)


def test_reads_output_max_cells_and_marks():
    res = pu.run_case(pu.small_case(n_reads=40000, n_cells=30, n_genes=90, merge="real", reads_output=True, max_cells=12, marks="eB"))
    pu.assert_parity(res)
    assert res["gpu"]["filtered"].shape[0] == 12


@pytest.mark.parametrize("reads_output,merge", [(False, "real"), (True, "none")])
def test_velocyto_matrices_for_other_query_marks(reads_output, merge):
    """-V (ResultsPrinter::save_intron_exon_matrices, ResultsPrinter.cpp:455-474): the filtered matrix over the SAME filtered cells for the
    query marks "e", "i" and "BA".  Expected values come from the reference's own (cell, gene, UMI, reads, mark) dump after merge_and_filter:
    Gene::number_of_requested_umis (Gene.cpp:60-79) counts the UMIs -- or sums their reads -- whose accumulated mark equals one of the query's."""
    case = pu.small_case(n_reads=60000, n_cells=30, n_genes=90, merge=merge, reads_output=reads_output, seed=41)
    case.extra["matrix_marks"] = ["e", "i", "BA", "eEBA"]
    res = pu.run_case(case)
    pu.assert_parity(res)
    ora, gpu = res["oracle"], res["gpu"]
    o_gene_ids = pu._gene_id_of_name(case, oracle_io.strings(ora["gene_names"]))
    col_of_cell = {int(c): k for k, c in enumerate(ora["filtered_cells"])}
    marks_of = {"e": {2}, "i": {4}, "BA": {6, 7}, "eEBA": {2, 3, 6, 7}}
    for code, (indptr, genes, vals) in gpu["cm_marks"].items():
        exp = {}
        for cell, gene, count, mark in zip(ora["umi_cell"], ora["umi_gene"], ora["umi_count"], ora["umi_mark"]):
            if int(cell) in col_of_cell and int(mark) in marks_of[code]:
                key = (col_of_cell[int(cell)], int(o_gene_ids[int(gene)]))
                exp[key] = exp.get(key, 0) + (int(count) if reads_output else 1)
        col = np.repeat(np.arange(indptr.shape[0] - 1), np.diff(indptr))
        got = {(int(c), int(g)): int(v) for c, g, v in zip(col, genes, vals)}
        assert indptr.shape[0] - 1 == len(col_of_cell) and got == exp and (len(exp) > 50 or code == "i"), code
        assert all(np.all(np.diff(genes[indptr[k]:indptr[k + 1]]) > 0) for k in range(indptr.shape[0] - 1))   # genes ascending inside a column
    # the container's own marks give the container's own matrix
    np.testing.assert_array_equal(gpu["cm_marks"]["eEBA"][2], gpu["cm"][2])
    assert sum(len(v[1]) for k, v in gpu["cm_marks"].items() if k != "eEBA") > 500


def test_intergenic_only_and_empty_inputs():
    # barcodes that only ever appear without a gene still become cells (CellsDataContainer.cpp:64-78)
    gene_ids = {}
    recs = records_from_strings([("ACGTACGT", "ACGT", None, 2), ("ACGTACGT", "ACGA", None, 2), ("TTTTACGT", "ACGT", "G1", 2),
                                 ("TTTTACGT", "ACGT", "G1", 4), ("TTTTACGT", "ACGT", "G2", 1)], gene_ids)
    case = pu.Case(name="intergenic", recs=recs, cb_len=8, umi_len=4, n_genes=2, gene_names=["G1", "G2"], merge="none",
                   min_genes_before=0, min_genes_after=0, shuffle=False, n_batches=2)
    res = pu.run_case(case)
    pu.assert_parity(res)
    assert res["gpu"]["summary"]["intergenic_reads"] == 2
    empty = pu.Case(name="empty", recs=np.zeros(0, dtype=dg.RECORD_DTYPE), cb_len=8, umi_len=4, n_genes=2, merge="none")
    out = pu.gpu_run(empty, empty.recs)
    assert out["summary"]["total_cells_number"] == 0 and out["cm"][0].shape[0] == 1


def test_call_order_errors_mirror_reference_throws():
    c = dg.Container(dg.Config(cb_len=8, umi_len=4, n_genes=2))
    with pytest.raises(dg.DgeError) as e:
        c.merge_and_filter()  # "You must initialize container" (CellsDataContainer.cpp:41-42)
    assert e.value.code == 2 and "initialize" in str(e.value)
    c.set_initialized()
    with pytest.raises(dg.DgeError) as e:
        c.add_batch(np.zeros(1, dtype=dg.RECORD_DTYPE))  # "Container is already initialized" (CellsDataContainer.cpp:61-62)
    assert e.value.code == 2
    with pytest.raises(dg.DgeError):
        c.set_initialized()
    c.close()


@pytest.mark.parametrize("key,gene,idx", [((0x1234 << 24) | (1 << 8), 1, 0),            # UMI bits beyond umi_len
                                          ((1 << 16) << 24, 1, 0),                      # barcode bits beyond cb_len
                                          (0x1234 << 24, 1 | (1 << 27), 0),             # reserved bits of the gene word
                                          (0x1234 << 24, 1, 0xFFFFFFFF),                # read_idx collides with the "none" sentinel
                                          (0x1234 << 24, 5, 0)])                        # gene id >= n_genes
def test_malformed_records_are_rejected(key, gene, idx):
    """A record whose fields do not fit the configured lengths would silently corrupt the grouping key: DGE_ERR_INVALID instead."""
    c = dg.Container(dg.Config(cb_len=8, umi_len=4, n_genes=2))
    recs = np.zeros(3, dtype=dg.RECORD_DTYPE)
    recs["key"] = [0x1111 << 24, key, 0x2222 << 24]
    recs["gene"] = [0 | (2 << 24), gene | (2 << 24) if gene < (1 << 24) else gene, 1 | (2 << 24)]
    recs["read_idx"] = [1, idx, 2]
    c.add_batch(recs)
    with pytest.raises(dg.DgeError) as e:
        c.set_initialized()
    assert e.value.code == 1
    c.close()


def test_full_size_invariants_20m():
    """Size-independent properties on a stream the oracle cannot chew in seconds: totals are conserved and the result does not
    depend on batching / order."""
    import torch

    wl = read_whitelist(pu.WL_SYNTH_7_9)
    spec = SynthSpec(n_reads=20_000_000, n_cells=2000, n_genes=5000, cb_len=16, umi_len=12, whitelist_parts=wl)
    t = SynthTables(spec)
    n = spec.n_reads
    buf = torch.empty(n * 16, dtype=torch.uint8, device="cuda:0")
    t.generate_device(0, 0, n, buf.data_ptr())
    torch.cuda.synchronize()

    def run(split):
        cfg = dg.Config(cb_len=16, umi_len=12, n_genes=5000, merge_type=dg.MERGE_REAL, barcodes_type=dg.BARCODES_CONST,
                        barcodes_file=pu.WL_SYNTH_7_9, min_genes_before_merge=20, min_genes_after_merge=50)
        c = dg.Container(cfg)
        if split == 1:
            c.add_batch_device(buf.data_ptr(), n)
        else:
            step = n // split
            for k in range(split):
                lo, hi = k * step, (n if k == split - 1 else (k + 1) * step)
                c.add_batch_device(buf.data_ptr() + lo * 16, hi - lo)
        c.set_initialized()
        c.merge_and_filter()
        s = c.summary()
        allc = c.cells(dg.CELLS_ALL)
        cm = c.matrix(dg.MATRIX_CM)
        raw = c.matrix(dg.MATRIX_CM_RAW)
        filt = c.cells(dg.CELLS_FILTERED)
        c.close()
        return s, allc, cm, raw, filt

    s1, all1, cm1, raw1, f1 = run(1)
    s2, all2, cm2, raw2, f2 = run(7)
    assert s1 == s2
    np.testing.assert_array_equal(all1, all2)
    for a, b in zip(cm1 + raw1, cm2 + raw2):
        np.testing.assert_array_equal(a, b)
    # conservation: every read is either intergenic or counted once in TOTAL_READS of an unmerged cell
    unmerged = (all1["flags"] & 2) == 0
    assert int(all1["reads_stat"][unmerged].sum()) + s1["intergenic_reads"] == n
    assert s1["has_exon_reads"] + s1["has_intron_reads"] + s1["has_not_annotated_reads"] == n - s1["intergenic_reads"]
    # filtered cells ascend in (requested genes, requested umis)
    k = f1["requested_genes_num"].astype(np.int64) * (1 << 32) + f1["requested_umis_num"]
    assert np.all(np.diff(k) >= 0)
    # cm_raw column sums == distinct UMIs of real cells that were never merge targets
    assert raw1[0][-1] == raw1[1].shape[0]


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_route_kernel_matches_host_mirror(world):
    """dge_route_by_barcode_device (the partition step before the all-to-all) against its numpy mirror."""
    import torch
    from dropest_b200 import dist as dgdist

    spec = SynthSpec(n_reads=50_000, n_cells=300, n_genes=100, cb_len=16, umi_len=10, whitelist_parts=read_whitelist(pu.WL_SYNTH_7_9))
    recs = SynthTables(spec).generate_host(0, spec.n_reads)
    raw = torch.from_numpy(np.frombuffer(recs.tobytes(), dtype=np.uint8).copy()).cuda()
    out = torch.empty_like(raw)
    counts = dgdist.route_device(0, raw.data_ptr(), recs.shape[0], world, out.data_ptr())
    exp_sorted, exp_counts = dgdist.route_host(recs, world)
    np.testing.assert_array_equal(counts, exp_counts)
    got = dgdist.records_from_tensor(out)
    off = 0
    for r in range(world):
        seg, exp = got[off:off + int(counts[r])], exp_sorted[off:off + int(counts[r])]
        np.testing.assert_array_equal(np.sort(seg, order=["read_idx"]), np.sort(exp, order=["read_idx"]))
        off += int(counts[r])


def _directional_case(umi_len, n_genes, n_reads, seed, **kw):
    wl = read_whitelist(pu.WL_SYNTH_7_9)
    spec = SynthSpec(n_reads=n_reads, n_cells=25, n_genes=n_genes, cb_len=16, umi_len=umi_len, whitelist_parts=wl, cb_error_ppm=50000,
                     reads_per_umi=kw.pop("reads_per_umi", 3), seed=seed)
    return pu.Case(name=f"directional_{umi_len}_{n_genes}", spec=spec, cb_len=16, umi_len=umi_len, n_genes=n_genes, merge=kw.pop("merge", "real"),
                   barcodes=pu.WL_SYNTH_7_9, min_genes_before=3, min_genes_after=5, umi_merge="directional", **kw)


@pytest.mark.parametrize("umi_len,n_genes,max_ed,mult", [(5, 40, 1, 2.0), (4, 12, 1, 2.0), (6, 8, 2, 2.0), (5, 20, 1, 1.0), (5, 6, 3, 0.5),
                                                         (8, 3, 1, 2.0)])
def test_directional_umi_merge(umi_len, n_genes, max_ed, mult):
    """`-u`: dense UMI neighbourhoods (short UMIs) so that segments of every size class merge, with ties between equal read
    counts (<= 16 UMIs: stable order on the device; larger: exact host replay of std::sort)."""
    case = _directional_case(umi_len, n_genes, 60000, seed=11 + umi_len, max_umi_ed=max_ed, umi_mult=mult)
    res = pu.run_case(case)
    pu.assert_parity(res)
    s = res["gpu"]["summary"]
    assert s["n_umis_merged"] > 0


def test_directional_umi_merge_large_segments():
    """Two genes and 5-base UMIs: segments of several hundred UMIs (block-per-segment kernel and the host replay)."""
    case = _directional_case(5, 2, 150000, seed=5, reads_per_umi=6, merge="none")
    res = pu.run_case(case)
    pu.assert_parity(res)


# ---- SimpleMergeStrategy (no whitelist: Drop-seq style protocols), reference Merge/SimpleMergeStrategy.cpp -----------------------
@pytest.mark.parametrize("seed", [7, 9])
def test_merge_simple_small(seed):
    res = pu.run_case(pu.small_case(n_reads=60000, n_cells=30, n_genes=120, merge="simple", seed=seed))
    pu.assert_parity(res)
    assert res["gpu"]["summary"]["n_merged"] > 0


def test_merge_simple_medium_strict_edit_distance():
    """max_cb_merge_edit_distance is a STRICT bound in this strategy (SimpleMergeStrategy.cpp:70): 1 forbids every merge."""
    for max_ed in (1, 2, 3):
        res = pu.run_case(pu.small_case(n_reads=300_000, n_cells=200, n_genes=800, merge="simple", min_genes_before=10, min_genes_after=20,
                                        cb_error_ppm=30000, reads_per_umi=3, max_cb_ed=max_ed))
        pu.assert_parity(res)
        if max_ed == 1:
            assert res["gpu"]["summary"]["n_merged"] == 0


@pytest.mark.parametrize("min_frac", [0.0, 0.1, 0.2])
def test_merge_simple_ties_replayed_in_reference_order(min_frac):
    """Bases with two or three admissible candidates (edit distance 1) that share the SAME number of UMI-genes and have the SAME
    size: exactly equal fractions, so the target depends on the iteration order of the reference's unordered containers --
    the host replay has to reproduce it."""
    rng = np.random.default_rng(11)
    genes = [f"G{i}" for i in range(6)]
    umis = ["".join("ACGT"[(v >> (2 * k)) & 3] for k in range(4)) for v in range(256)]
    reads = []

    def mutate(cb, pos, step):
        return cb[:pos] + "ACGT"[("ACGT".index(cb[pos]) + step) % 4] + cb[pos + 1:]

    for grp in range(16):
        base = "".join("ACGT"[int(x)] for x in rng.integers(0, 4, 10))
        n_parents = 2 + grp % 2
        parents = [mutate(base, 1 + 3 * k, 1 + (grp + k) % 3) for k in range(n_parents)]
        pool = [(g, u) for g in genes[:4] for u in rng.choice(umis, 3, replace=False)]      # the base's UMI-genes: 4 genes
        for g, u in pool:
            for _ in range(int(rng.integers(1, 3))):
                reads.append((base, u, g, 2))
        share = len(pool) // n_parents
        for k, p in enumerate(parents):
            mine = pool[k * share:(k + 1) * share]                                          # same number shared with every parent
            extra = [(g, u) for g in genes for u in rng.choice(umis, 2, replace=False)]
            extra = [x for x in extra if x not in pool][:10]                                # same total size for every parent
            for g, u in mine + extra:
                reads.append((p, u, g, 2))
    order = rng.permutation(len(reads))
    reads = [reads[i] for i in order]
    gene_ids = {}
    recs = records_from_strings(reads, gene_ids)
    names = [n for n, _ in sorted(gene_ids.items(), key=lambda kv: kv[1])]
    case = pu.Case(name="simple_ties", recs=recs, cb_len=10, umi_len=4, n_genes=len(names), gene_names=names, merge="simple",
                   min_genes_before=2, min_genes_after=2, min_frac=min_frac, max_cb_ed=2, shuffle=False, n_batches=2)
    res = pu.run_case(case)
    pu.assert_parity(res)
    assert res["gpu"]["summary"]["n_cb_merge_replayed"] > 0


# ---- MergeAllMergeStrategy (merge_type=all), reference Merge/MergeAllMergeStrategy.h:16-50 ----------------------------------------
@pytest.mark.parametrize("max_ed", [1, 2, 3])
def test_merge_all_small(max_ed):
    res = pu.run_case(pu.small_case(n_reads=60000, n_cells=30, n_genes=120, merge="all", seed=21, max_cb_ed=max_ed))
    pu.assert_parity(res)
    assert res["gpu"]["summary"]["n_merged"] > 0


def test_merge_all_short_barcodes_many_neighbours():
    """8-base barcodes drawn at random: plenty of cells within the edit distance of each other, chains of merges and equal
    (distance, size) candidates where the earliest filtered cell has to win."""
    rng = np.random.default_rng(3)
    cbs = ["".join("ACGT"[int(x)] for x in rng.integers(0, 3, 8)) for _ in range(120)]
    genes = [f"G{i}" for i in range(8)]
    reads = []
    for cb in cbs:
        for _ in range(int(rng.integers(3, 12))):
            reads.append((cb, "".join("ACGT"[int(x)] for x in rng.integers(0, 4, 5)), genes[int(rng.integers(0, 8))], 2))
    reads = [reads[i] for i in rng.permutation(len(reads))]
    gene_ids = {}
    recs = records_from_strings(reads, gene_ids)
    names = [n for n, _ in sorted(gene_ids.items(), key=lambda kv: kv[1])]
    case = pu.Case(name="all_short", recs=recs, cb_len=8, umi_len=5, n_genes=len(names), gene_names=names, merge="all",
                   min_genes_before=2, min_genes_after=2, max_cb_ed=2, shuffle=False, n_batches=2)
    res = pu.run_case(case)
    pu.assert_parity(res)
    assert res["gpu"]["summary"]["n_merged"] > 10


def test_soa_batches_with_implicit_read_index():
    """dge_add_batch_soa: key / gene arrays, read_idx = stream position; same results as the 16-byte records."""
    case = pu.small_case(n_reads=80000, n_cells=40, n_genes=150, merge="real", seed=31)
    case.extra["soa"] = True
    case.n_batches = 4
    res = pu.run_case(case)
    pu.assert_parity(res)


@pytest.mark.parametrize("cuts", [(0.5,), (0.001, 0.001, 0.4, 0.95), (0.3, 0.31, 0.9)])
def test_segment_batches_in_one_launch(cuts):
    """dge_add_batch_segments_device: several device arrays (here: shuffled reads cut into very uneven pieces, an empty one included)
    consumed by ONE fill launch with interleaved tiles -- the shape of the peer-memory exchange -- give the reference's result."""
    case = pu.small_case(n_reads=120000, n_cells=40, n_genes=150, merge="real", seed=33)
    case.extra["segments"] = cuts
    case.shuffle = True
    res = pu.run_case(case)
    pu.assert_parity(res)


@pytest.mark.parametrize("merge", ["none", "real"])
def test_per_chromosome_stats(merge):
    """Stats' EXON / INTRON / INTERGENIC_READS_PER_CHR_PER_CELL (row a4): the chromosome side array feeds the (cell, chromosome) counter
    table; with a CB merge the counters of merged cells are added to their targets (Stats::merge); equal to the reference's
    get_stat_by_real_cells output, including which cells and which chromosomes each table lists."""
    case = pu.small_case(n_reads=150000, n_cells=50, n_genes=200, merge=merge, seed=41)
    case.extra["n_chr"] = 7
    case.n_batches = 4
    case.shuffle = False
    res = pu.run_case(case)
    pu.assert_parity(res)
    counts, presented = res["gpu"]["chr_stats"]
    assert counts[:, :, 0].sum() > 0 and counts[:, :, 1].sum() > 0 and counts[:, :, 2].sum() > 0
    assert presented[2, 6] and not presented[0, 6] and not presented[1, 6]   # the last chromosome only ever holds intergenic reads


def test_oversized_sub_buckets_have_no_capacity_limit():
    """One barcode with 9 M reads is ONE L1 bucket of ~9 M keys: with at most 2048 sub-buckets per bucket almost every sub-bucket is larger
    than the biggest sort class (4096 keys) and holds thousands of distinct keys, plus one (gene, UMI) repeated 60 k times.  Such
    sub-buckets are sorted as segments in global memory and run-length encoded (k_tail_offsets / cub segmented sort / k_tail_dedup);
    they used to overflow a shared-memory hash table.  Checked against numpy: every distinct (cell, gene, UMI) with its read count and
    accumulated mark."""
    rng = np.random.default_rng(3)
    n = 9_000_000
    recs = np.zeros(n, dtype=dg.RECORD_DTYPE)
    cb = np.where(rng.random(n) < 0.97, np.uint64(0x1234567), rng.integers(0, 1 << 20, size=n).astype(np.uint64))
    umi = rng.integers(0, 1 << 20, size=n).astype(np.uint64)          # 10-base UMIs: ~3 M distinct values, so (gene, UMI) pairs repeat
    gene = rng.integers(0, 40, size=n).astype(np.uint32)
    mark = rng.choice(np.array([1, 2, 4, 6], dtype=np.uint32), size=n)
    umi[:60_000] = 77; gene[:60_000] = 5; cb[:60_000] = np.uint64(0x1234567)
    recs["key"] = (cb << np.uint64(24)) | umi
    recs["gene"] = gene | (mark << np.uint32(24))
    recs["read_idx"] = np.arange(n, dtype=np.uint32)
    c = dg.Container(dg.Config(cb_len=16, umi_len=10, n_genes=40, merge_type=dg.MERGE_NONE, min_genes_before_merge=1, min_genes_after_merge=1,
                               max_barcodes_hint=1 << 21))
    c.add_batch(recs)
    c.set_initialized()
    c.merge_and_filter()
    u = c.umigs(dg.CELLS_ALL)
    cells = c.cells(dg.CELLS_ALL)
    got_key = (cells["barcode"][u["cell"]].astype(np.uint64) << np.uint64(32)) | (u["gene"].astype(np.uint64) << np.uint64(24)) | u["umi"].astype(np.uint64)
    exp_key = (cb << np.uint64(32)) | (gene.astype(np.uint64) << np.uint64(24)) | umi
    uniq, inv, cnt = np.unique(exp_key, return_inverse=True, return_counts=True)
    exp_mark = np.zeros(uniq.shape[0], dtype=np.uint32)
    np.bitwise_or.at(exp_mark, inv, mark)
    order = np.argsort(got_key)
    np.testing.assert_array_equal(got_key[order], uniq)
    np.testing.assert_array_equal(u["count"][order].astype(np.int64), cnt)
    np.testing.assert_array_equal(u["mark"][order].astype(np.uint32), exp_mark)
    assert cnt.max() >= 60_000
    c.close()


@pytest.mark.parametrize("giant", [1000, 50000])
def test_long_l1_buckets_shared_by_clusters(monkeypatch, giant):
    """k_l2_bucket_cluster (8-CTA clusters, histograms combined through distributed shared memory) takes the L1 buckets longer than
    DGE_L2_GIANT keys; with the threshold lowered every bucket (1000) or only the longest ones (50000) go through it."""
    monkeypatch.setenv("DGE_L2_GIANT", str(giant))
    wl = read_whitelist(pu.WL_SYNTH_7_9)
    spec = SynthSpec(n_reads=3_000_000, n_cells=300, n_genes=2000, cb_len=16, umi_len=10, whitelist_parts=wl, cb_error_ppm=20000, seed=29)
    case = pu.Case(name="giant_buckets", spec=spec, cb_len=16, umi_len=10, n_genes=2000, merge="real", barcodes=pu.WL_SYNTH_7_9,
                   min_genes_before=10, min_genes_after=30, dump_umis=False, n_batches=2)
    res = pu.run_case(case)
    pu.assert_parity(res)


def test_merge_simple_dropseq_like_2m():
    """Drop-seq shaped stream (12 bp barcodes, 8 bp UMIs, no whitelist -> SimpleMergeStrategy) at a size where the inverted index and
    the pair lists are non-trivial; checked against the oracle."""
    wl = read_whitelist(pu.WL_SYNTH_8_8)
    parts = [[t[:6] for t in wl[0]], [t[:6] for t in wl[1]]]
    parts = [sorted(set(p)) for p in parts]
    spec = SynthSpec(n_reads=2_000_000, n_cells=1500, n_genes=3000, cb_len=12, umi_len=8, whitelist_parts=parts, cb_error_ppm=40000,
                     reads_per_umi=3, seed=17)
    case = pu.Case(name="dropseq_simple", spec=spec, cb_len=12, umi_len=8, n_genes=3000, merge="simple", min_genes_before=10,
                   min_genes_after=30, max_cb_ed=2, dump_umis=False, n_batches=5)
    res = pu.run_case(case)
    pu.assert_parity(res)
    assert res["gpu"]["summary"]["n_merged"] > 300


def test_full_size_invariants_simple_merge_20m():
    """SimpleMergeStrategy at 20 M reads: the result does not depend on batching, totals are conserved."""
    import torch

    wl = read_whitelist(pu.WL_SYNTH_7_9)
    spec = SynthSpec(n_reads=20_000_000, n_cells=2000, n_genes=5000, cb_len=16, umi_len=10, whitelist_parts=wl, seed=23)
    t = SynthTables(spec)
    n = spec.n_reads
    buf = torch.empty(n * 16, dtype=torch.uint8, device="cuda:0")
    t.generate_device(0, 0, n, buf.data_ptr())
    torch.cuda.synchronize()

    def run(split):
        cfg = dg.Config(cb_len=16, umi_len=10, n_genes=5000, merge_type=dg.MERGE_SIMPLE, min_genes_before_merge=20, min_genes_after_merge=50,
                        max_cb_merge_edit_distance=2)
        c = dg.Container(cfg)
        step = n // split
        for k in range(split):
            lo, hi = k * step, (n if k == split - 1 else (k + 1) * step)
            c.add_batch_device(buf.data_ptr() + lo * 16, hi - lo)
        c.set_initialized()
        c.merge_and_filter()
        out = (c.summary(), c.cells(dg.CELLS_ALL), c.matrix(dg.MATRIX_CM))
        c.close()
        return out

    s1, all1, cm1 = run(1)
    s2, all2, cm2 = run(5)
    assert s1 == s2 and s1["n_merged"] > 1000
    np.testing.assert_array_equal(all1, all2)
    for a, b in zip(cm1, cm2):
        np.testing.assert_array_equal(a, b)
    unmerged = (all1["flags"] & 2) == 0
    assert int(all1["reads_stat"][unmerged].sum()) + s1["intergenic_reads"] == n


# ---- precise merge (-M): PoissonSimpleMergeStrategy / PoissonRealBarcodesMergeStrategy on top of PoissonTargetEstimator ----------------------
@pytest.mark.parametrize("seed,probs", [(21, (1e-4, 1e-7)), (22, (1e-2, 1e-3)), (23, (0.5, 0.2))])
def test_merge_poisson_simple_small(seed, probs):
    """-M without a whitelist; the thresholds range from the defaults to values that accept most candidates (so that argmin over
    several neighbours and chains of merges occur)."""
    res = pu.run_case(pu.small_case(n_reads=60000, n_cells=30, n_genes=120, merge="poisson_simple", seed=seed, max_merge_prob=probs[0],
                                    max_real_merge_prob=probs[1]), kind="reference")
    pu.assert_parity(res)
    if probs[1] >= 1e-3:
        assert res["gpu"]["summary"]["n_merged"] > 0


def test_merge_poisson_simple_dropseq_like():
    """BASELINE configs[4] shape (12 bp barcodes, 8 bp UMIs -> 65 536-UMI space where collisions are material, no whitelist, -M with the
    drop_seq.xml probabilities 1e-5 / 1e-7), scaled down."""
    wl = read_whitelist(pu.WL_SYNTH_8_8)
    parts = [sorted(set(t[:6] for t in wl[0])), sorted(set(t[:6] for t in wl[1]))]
    spec = SynthSpec(n_reads=400_000, n_cells=300, n_genes=2000, cb_len=12, umi_len=8, whitelist_parts=parts, cb_error_ppm=40000, reads_per_umi=3, seed=27)
    case = pu.Case(name="dropseq_poisson", spec=spec, cb_len=12, umi_len=8, n_genes=2000, merge="poisson_simple", min_genes_before=10, min_genes_after=30,
                   max_cb_ed=2, max_merge_prob=1e-5, max_real_merge_prob=1e-7, dump_umis=False, n_batches=3)
    res = pu.run_case(case, kind="reference")
    pu.assert_parity(res)
    assert res["gpu"]["summary"]["n_merged"] > 50


@pytest.mark.parametrize("seed,probs", [(31, (1e-4, 1e-7)), (32, (0.3, 0.05))])
def test_merge_poisson_real_small(seed, probs):
    """-M with a whitelist: neighbours beyond the nearest distance class (get_max_merge_dist = min + 1), whitelist barcodes that merge
    into other whitelist barcodes, chains in phase 2."""
    res = pu.run_case(pu.small_case(n_reads=40000, n_cells=25, n_genes=100, merge="poisson_real", seed=seed, max_merge_prob=probs[0],
                                    max_real_merge_prob=probs[1]), kind="reference")
    pu.assert_parity(res)
