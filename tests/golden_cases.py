"""Seeded workloads whose reference outputs are committed under tests/golden/case_*.npz (made by tests/golden/make_golden.py)."""
import os
import tempfile

import numpy as np

import parity_utils as pu
from dropest_b200.synth import SynthSpec, read_whitelist, records_from_strings, write_packed
import oracle_io

FIXTURE_READS = [
    ("AAATTAGGTCCA", "AAACCT", "Gene1"), ("AAATTAGGTCCA", "CCCCCT", "Gene2"), ("AAATTAGGTCCA", "ACCCCT", "Gene3"),
    ("AAATTAGGTCCA", "ACCCCT", "Gene4"), ("AAATTAGGTCCC", "CAACCT", "Gene1"), ("AAATTAGGTCCC", "CAACCT", "Gene10"),
    ("AAATTAGGTCCC", "CAACCT", "Gene20"), ("AAATTAGGTCCG", "CAACCT", "Gene1"), ("AAATTAGGTCGG", "AAACCT", "Gene1"),
    ("AAATTAGGTCGG", "CCCCCT", "Gene2"), ("CCCTTAGGTCCA", "CCATTC", "Gene3"), ("CCCTTAGGTCCA", "CCCCCT", "Gene2"),
    ("CCCTTAGGTCCA", "ACCCCT", "Gene3"), ("CAATTAGGTCCG", "CAACCT", "Gene1"), ("CAATTAGGTCCG", "AAACCT", "Gene1"),
    ("CAATTAGGTCCG", "CCCCCT", "Gene2"), ("AAAAAAAAAAAA", "CCCCCT", "Gene2"),
]


def fixture_case(**kw) -> pu.Case:
    """Fixture of the reference's Tests/TestEstimation.cpp:33-80 (17 reads, 7 barcodes, 3x3 inDrop whitelist)."""
    gene_ids = {}
    recs = records_from_strings([(cb, umi, g, 2) for cb, umi, g in FIXTURE_READS], gene_ids)
    names = [n for n, _ in sorted(gene_ids.items(), key=lambda kv: kv[1])]
    return pu.Case(name="test_est_fixture", recs=recs, cb_len=12, umi_len=6, n_genes=len(names), gene_names=names, merge="real",
                   barcodes=pu.WL_TEST_EST, barcodes_type="indrop", min_genes_before=0, min_genes_after=0, max_cb_ed=7,
                   min_frac=0.0, shuffle=False, n_batches=1, **kw)


def cases():
    wl8 = read_whitelist(pu.WL_SYNTH_8_8)
    return {
        "fixture": fixture_case(),
        "real_7x9": pu.small_case(n_reads=30000, n_cells=30, n_genes=80, merge="real", seed=11),
        "none_7x9": pu.small_case(n_reads=20000, n_cells=25, n_genes=60, merge="none", seed=12),
        "simple_7x9": pu.small_case(n_reads=30000, n_cells=30, n_genes=80, merge="simple", seed=14),
        # -M strategies: oracle port and CUDA path are both pinned against these
        "poisson_simple_7x9": pu.small_case(n_reads=30000, n_cells=30, n_genes=80, merge="poisson_simple", seed=15),
        "poisson_real_7x9": pu.small_case(n_reads=30000, n_cells=30, n_genes=80, merge="poisson_real", seed=16),
        "all_7x9": pu.small_case(n_reads=30000, n_cells=30, n_genes=80, merge="all", seed=17),
        "directional_7x9": pu.small_case(n_reads=30000, n_cells=30, n_genes=40, merge="real", seed=18, umi_merge="directional", reads_per_umi=6),
        # per-chromosome Stats tables (row a4): a chromosome id per read, whitelist merge so that Stats::merge matters
        "real_7x9_chr": pu.small_case(n_reads=30000, n_cells=30, n_genes=80, merge="real", seed=19, extra={"n_chr": 7}),
        "real_8x8_reads": pu.Case(name="real_8x8_reads",
                                  spec=SynthSpec(n_reads=25000, n_cells=20, n_genes=70, cb_len=16, umi_len=6, whitelist_parts=wl8,
                                                 cb_error_ppm=80000, seed=13),
                                  cb_len=16, umi_len=6, n_genes=70, merge="real", barcodes=pu.WL_SYNTH_8_8, barcodes_type="indrop",
                                  min_genes_before=5, min_genes_after=8, reads_output=True, marks="eiB"),
    }


def case_records(case: pu.Case) -> np.ndarray:
    if case.spec is not None:
        from dropest_b200.synth import SynthTables

        return SynthTables(case.spec).generate_host(0, case.spec.n_reads)
    return case.recs


def case_chr_ids(case: pu.Case, recs: np.ndarray):
    if case.chr_ids is None and case.extra.get("n_chr"):
        case.chr_ids = pu.synth_chr_ids(recs, int(case.extra["n_chr"]))
    return case.chr_ids


def run_oracle_on(case: pu.Case, kind: str = "any"):
    recs = case_records(case)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "reads.bin")
        write_packed(path, recs, case.cb_len, case.umi_len, case.n_genes, case.gene_names, chr_ids=case_chr_ids(case, recs))
        return oracle_io.run_oracle(path, kind=kind, merge=case.merge, barcodes=case.barcodes, barcodes_type=case.barcodes_type,
                                    min_genes_before=case.min_genes_before, min_genes_after=case.min_genes_after,
                                    max_cb_ed=case.max_cb_ed, min_frac=case.min_frac, marks=case.marks, max_cells=case.max_cells,
                                    reads_output=case.reads_output, dump_umis=True, umi_merge=case.umi_merge, max_umi_ed=case.max_umi_ed,
                                    umi_mult=case.umi_mult)


def load_golden(name: str):
    path = os.path.join(pu.GOLDEN, f"case_{name}.npz")
    with np.load(path) as z:
        return {k: z[k] for k in z.files}
