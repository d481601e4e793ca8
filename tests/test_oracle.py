"""Pins of the CPU oracle (test infrastructure): the restatement in oracle/port must reproduce
  (a) the literal expectations written in the reference's own unit tests,
  (b) tests/golden/* = outputs of the compiled, unmodified reference (made by tests/golden/make_golden.py),
  (c) the compiled reference itself on fresh seeded streams, whenever oracle/_ref is present (build container)."""
import json
import os
import subprocess

import numpy as np
import pytest

import golden_cases
import oracle_io
import parity_utils as pu
from dropest_b200.synth import SynthSpec, read_whitelist

needs_port = pytest.mark.skipif(not oracle_io.available("port"), reason="oracle/_build/dropest_port not built")
needs_ref = pytest.mark.skipif(not oracle_io.available("reference"), reason="oracle/_ref needs /root/reference (build container only)")


def pins():
    with open(os.path.join(pu.GOLDEN, "ref_pins.json")) as f:
        return json.load(f)


def same(a, b):
    keys = [k for k in a if not k.startswith("_") and not k.startswith("t_")]
    assert keys
    for k in keys:
        assert k in b, k
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)


def test_golden_pins_equal_literal_expectations_of_reference_tests():
    p = pins()
    # Tests/TestEstimation.cpp:98-121 testBarcodesFile
    assert p["testBarcodesFile"] == {"part0": ["AAT", "GAA", "AAA"], "part1": ["TTAGGTCCA", "TTAGGGGCC", "TTAGGTCCC"]}
    # :160-178 testUmigsIntersection
    assert p["testUmigsIntersection"] == [2, 1, 0]
    # :180-206 testFillDistances
    assert p["testFillDistances"]["values0"] == [1, 1, 2] and p["testFillDistances"]["index0"][2] == 1
    assert p["testFillDistances"]["values1"] == [1, 1, 2] and p["testFillDistances"]["index1"][2] == 1
    # :208-225 testRealNeighboursCbs
    assert p["testRealNeighboursCbs"]["CAATTAGGTCCG"] == ["AAATTAGGTCCA", "AAATTAGGTCCC"]
    assert p["testRealNeighboursCbs"]["AAATTAGGTCCC"] == ["AAATTAGGTCCC"]
    # :227-235 testRealNeighbours
    assert p["testRealNeighbours"][:6] == [0, 1, 1, 0, 0, 0]
    # :237-280 testMergeByRealBarcodes
    m = p["testMergeByRealBarcodes"]
    assert m["total_cells"] == 7 and len(m["filtered"]) == 2
    assert m["gene_sizes"] == [3, 4, 1, 2, 1, 2] and m["read_counts"] == [2, 3, 4, 2, 1]
    assert m["merged"] == [0, 0, 1, 1, 1, 1, 0] and sum(m["excluded"]) == 1
    # :282-320 testSplitBarcode / testConstLengthBarcodeParser
    c = p["testConstLengthBarcodeParser"]
    assert c["indrop_lengths"] == [8, 8] and c["indrop_sizes"] == [384, 384]
    assert c["tenx_lengths"] == [7, 9] and c["tenx_sizes"] == [480, 1536] and c["tenx_min_dists"] == [0, 0]
    assert c["split_indrop"] == ["TAATGAGC", "ACTAATGA"]
    # :490-540 testUMIMergeStrategySimple
    u = p["testUMIMergeStrategySimple"]
    assert u["Gene1"] == {"AAACCG": 1, "AAACCT": 3, "ACCCCT": 1, "CCCCCT": 1} and len(u["Gene2"]) == 3
    assert all("N" not in k for k in u["Gene2"])
    # :588-608 testUMIMergeStrategyDirectional
    assert p["testUMIMergeStrategyDirectional"] == {"AAA": "AGT", "AAT": "AGT", "CCC": "TCC"}
    # Tests/TestTools.cpp:47-54 testEditDistance
    assert p["testEditDistance"][:5] == [1, 1, 2, 2, 0]
    # Tests/TestEstimationMergeProbs.cpp:93-140 (the asserts that hold on the compiled reference, SURVEY section 4 caveat)
    assert p["poisson"]["umi_distribution_size"] == 8 and p["poisson"]["probs"][0] == 1
    assert p["poisson"]["merge_targets_phase1"][7] == -1


@needs_port
@pytest.mark.parametrize("name", ["fixture", "real_7x9", "none_7x9", "simple_7x9", "poisson_simple_7x9", "poisson_real_7x9", "all_7x9", "directional_7x9", "real_8x8_reads", "real_7x9_chr"])
def test_port_reproduces_golden_reference_outputs(name):
    res = golden_cases.run_oracle_on(golden_cases.cases()[name], kind="port")
    assert res["_kind"] == "port"
    same(golden_cases.load_golden(name), res)


@needs_port
def test_port_fixture_matches_literal_expectations():
    res = golden_cases.run_oracle_on(golden_cases.fixture_case(), kind="port")
    assert list(res["merge_targets"]) == [0, 1, 1, 0, 0, 0, 6]
    assert list(res["filtered_cells"]) == [1, 0]
    assert list(res["cell_n_genes"][[1, 0]]) == [3, 4]
    assert [int(f) & 2 for f in res["cell_flags"]] == [0, 0, 2, 2, 2, 2, 0]


@needs_port
def test_port_edit_distance_pins():
    exp = pins()["testEditDistance"]
    args = [("ATTTTC", "ATTTGC", 1, 10000), ("ATTTTCC", "ATTTGNC", 1, 10000), ("ATTTTCC", "ATTTGNC", 0, 10000),
            ("ATTTTCC", "ATTTGTC", 1, 10000), ("ATTTTCC", "ATTTTCC", 1, 10000), ("ACGTACG", "ACGTACGT", 1, 1),
            ("AAAA", "TTTT", 1, 1), ("ACGTAC", "ACGAAC", 1, 1), ("ACGTACGT", "ACGACGTT", 1, 2)]
    for (a, b, sn, me), e in zip(args, exp):
        out = subprocess.run([oracle_io.PORT_BIN, "--edit-distance", a, b, str(sn), str(me)], capture_output=True, text=True, check=True)
        assert int(out.stdout) == e, (a, b, sn, me)


@needs_port
@needs_ref
@pytest.mark.parametrize("merge", ["none", "real", "simple", "all", "poisson_real", "poisson_simple"])
@pytest.mark.parametrize("seed", [21, 22])
def test_port_matches_compiled_reference_on_fresh_streams(merge, seed):
    case = pu.small_case(n_reads=25000, n_cells=25, n_genes=70, merge=merge, seed=seed)
    if merge in ("real", "poisson_real"):
        case.barcodes = pu.WL_SYNTH_7_9
    if seed == 22:
        case.extra["n_chr"] = 5   # per-chromosome Stats tables as well
    a = golden_cases.run_oracle_on(case, kind="reference")
    b = golden_cases.run_oracle_on(case, kind="port")
    assert a["_kind"] == "reference" and b["_kind"] == "port"
    same(a, b)


@needs_ref
def test_golden_files_are_current():
    """The committed fixtures are what the compiled reference produces today."""
    out = subprocess.run([oracle_io.REF_PINS_BIN, "/root/reference/data"], capture_output=True, text=True, check=True).stdout
    assert json.loads(out) == pins()
    same(golden_cases.load_golden("real_7x9"), golden_cases.run_oracle_on(golden_cases.cases()["real_7x9"], kind="reference"))
