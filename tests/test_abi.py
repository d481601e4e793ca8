"""The C-ABI library loads without a GPU and exports every symbol include/dropest_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

import dropest_b200 as dg
from dropest_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "dropest_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dge_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = dg.load_library()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/dropest_b200.h but not exported"
    assert set(capi.EXPORTS) <= set(syms)


def test_every_export_has_a_ctypes_prototype():
    """A missing argtypes list makes ctypes pass 64-bit sizes / pointers as 32-bit ints."""
    lib = dg.load_library()
    missing = [s for s in capi.EXPORTS if getattr(lib, s).argtypes is None]
    assert not missing, missing


def test_struct_layouts_match_header():
    assert ctypes.sizeof(capi._Config) == 120
    assert capi.CELL_INFO_DTYPE.itemsize == 40
    assert capi.RECORD_DTYPE.itemsize == 16
    assert ctypes.sizeof(capi._Summary) == 19 * 8
    assert ctypes.sizeof(capi._SynthParams) == 96


def test_config_defaults_follow_merge_strategy_factory():
    """MergeStrategyFactory.cpp:23-59 defaults; query marks 'eEBA' (CellsDataContainer.cpp:17)."""
    lib = dg.load_library()
    c = capi._Config()
    lib.dge_config_default(ctypes.byref(c))
    assert (c.min_genes_before_merge, c.min_genes_after_merge) == (10, 10)
    assert c.min_merge_fraction == 0.2 and c.max_merge_prob == 1e-4 and c.max_real_merge_prob == 1e-7
    assert c.max_umi_merge_edit_distance == 1 and c.umi_merge_mult == 2.0
    assert c.barcodes_type == dg.BARCODES_INDROP
    assert c.query_mark_mask == dg.marks_to_mask("eEBA") == 0xCC
    assert dg.marks_to_mask("e") == 1 << 2 and dg.marks_to_mask("iIBA") == (1 << 4) | (1 << 5) | (1 << 6) | (1 << 7)


def test_no_cpu_fallback_without_device():
    """The product path must fail loudly when no CUDA device is usable -- never fall back to a CPU implementation."""
    import torch

    if torch.cuda.is_available():
        return
    try:
        dg.Container(dg.Config(cb_len=8, umi_len=4, n_genes=2))
    except dg.DgeError as e:
        assert e.code == 3 and "CUDA" in str(e)
    else:
        raise AssertionError("dge_create succeeded without a CUDA device")


def test_edit_and_hamming_distance_reference_pins():
    """Tests/TestTools.cpp:47-54 literal expectations + banded behaviour observed on the compiled reference."""
    ed = capi.edit_distance
    assert ed("ATTTTC", "ATTTGC") == 1
    assert ed("ATTTTCC", "ATTTGNC") == 1
    assert ed("ATTTTCC", "ATTTGNC", skip_n=False) == 2
    assert ed("ATTTTCC", "ATTTGTC") == 2
    assert ed("ATTTTCC", "ATTTTCC") == 0
    assert ed("ACGTACG", "ACGTACGT", True, 1) == 8
    assert ed("AAAA", "TTTT", True, 1) == 2
    assert capi.hamming_distance("AAANTTT", "AAACTTT") == 0
    assert capi.hamming_distance("AAANTTT", "AAACTTT", skip_n=False) == 1
    assert capi.hamming_distance("AAA", "AAAA") == 0xFFFFFFFF
