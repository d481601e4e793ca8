"""dropest_b200 -- B200-native (sm_100a CUDA) implementation of dropEst's count-matrix hot path.

The product is the C-ABI shared library ``dropest_b200/lib/libdropest_b200.so`` (declared in ``include/dropest_b200.h``);
this package is the thin Python binding used by the tests, ``bench.py`` and ``__graft_entry__.py``.
There is no CPU fallback: importing works anywhere, but every compute call needs a CUDA device.
"""
from .capi import (  # noqa: F401
    Config,
    Container,
    DgeError,
    lib_path,
    load_library,
    MERGE_NONE,
    MERGE_REAL,
    MERGE_SIMPLE,
    MERGE_POISSON_REAL,
    MERGE_POISSON_SIMPLE,
    MERGE_ALL,
    UMI_MERGE_SIMPLE,
    UMI_MERGE_DIRECTIONAL,
    BARCODES_CONST,
    BARCODES_INDROP,
    CELLS_ALL,
    CELLS_REAL,
    CELLS_FILTERED,
    MATRIX_CM,
    MATRIX_CM_RAW,
    RECORD_DTYPE,
    NO_GENE,
    FLAG_UMI_N,
    FLAG_CB_N,
    CB_N_BIT,
    UMI_N_BIT,
    pack_seq,
    unpack_seq,
    marks_to_mask,
    collisions_adjusted_sizes,
)
