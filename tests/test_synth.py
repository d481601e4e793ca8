"""Host-side logic: synthetic stream generator, record packing, routing hash (CPU only)."""
import numpy as np

import dropest_b200 as dg
from dropest_b200 import synth
import parity_utils as pu


def test_pack_unpack_roundtrip_and_order():
    rng = np.random.default_rng(0)
    for _ in range(200):
        n = int(rng.integers(1, 21))
        s = "".join("ACGT"[int(x)] for x in rng.integers(0, 4, n))
        assert dg.unpack_seq(dg.pack_seq(s), n) == s
    a, b = "ACGTACGTACGTACGT", "ACGTACGTACGTACTT"
    assert (a < b) == (dg.pack_seq(a) < dg.pack_seq(b))  # numeric order == string order (compare_cells tie-break)


def test_host_generator_is_counter_based():
    wl = synth.read_whitelist(pu.WL_SYNTH_7_9)
    t = synth.SynthTables(synth.SynthSpec(n_reads=100000, n_cells=50, n_genes=200, cb_len=16, umi_len=12, whitelist_parts=wl))
    whole = t.generate_host(0, 5000)
    np.testing.assert_array_equal(whole[1234:2345], t.generate_host(1234, 1111))
    assert whole["read_idx"].tolist() == list(range(5000))
    marks = (whole["gene"] >> 24) & 7
    assert set(np.unique(marks)) <= {1, 2, 4}
    frac_inter = np.mean((whole["gene"] & 0xFFFFFF) == dg.NO_GENE)
    assert 0.03 < frac_inter < 0.07
    # true barcodes are whitelist products; ~2 % of reads carry one substitution
    true = set(int(x) for x in t.cell_barcode)
    cbs = (whole["key"] >> np.uint64(24)).astype(np.uint64)
    frac_err = np.mean([int(c) not in true for c in cbs])
    assert 0.01 < frac_err < 0.035


def test_whitelist_reader_reverse_complements_like_reference():
    parts = synth.read_whitelist(pu.WL_TEST_EST)
    assert parts == [["AAT", "GAA", "AAA"], ["TTAGGTCCA", "TTAGGGGCC", "TTAGGTCCC"]]  # Tests/TestEstimation.cpp:98-121


def test_rank_of_partitions_barcodes():
    rng = np.random.default_rng(1)
    cb = rng.integers(0, 1 << 32, size=100000, dtype=np.uint64)
    for n in (1, 2, 4, 8):
        r = synth.rank_of(cb, n)
        assert r.min() >= 0 and r.max() < n
        counts = np.bincount(r, minlength=n)
        assert counts.min() > 0.9 * len(cb) / n
        np.testing.assert_array_equal(r, synth.rank_of(cb.copy(), n))  # pure function of the barcode
