"""ctypes binding of include/dropest_b200.h (the drop-in C ABI).  No torch types cross this boundary."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

MERGE_NONE, MERGE_REAL, MERGE_SIMPLE, MERGE_POISSON_REAL, MERGE_POISSON_SIMPLE, MERGE_ALL = range(6)
BARCODES_CONST, BARCODES_INDROP = 0, 1
UMI_MERGE_SIMPLE, UMI_MERGE_DIRECTIONAL = 0, 1
CELLS_ALL, CELLS_REAL, CELLS_FILTERED = 0, 1, 2
MATRIX_CM, MATRIX_CM_RAW = 0, 1
NO_GENE = 0xFFFFFF
FLAG_UMI_N, FLAG_CB_N = 1 << 27, 1 << 28     # dge_record16.gene flags: key field = index into the N-UMI / N-barcode list
CB_N_BIT, UMI_N_BIT = 1 << 40, 1 << 31       # how the query surface reports such entries
ABI_VERSION = 2

RECORD_DTYPE = np.dtype([("key", "<u8"), ("gene", "<u4"), ("read_idx", "<u4")])

CELL_INFO_DTYPE = np.dtype(
    [
        ("barcode", "<u8"),
        ("first_read_idx", "<u4"),
        ("flags", "<u4"),
        ("n_genes", "<i4"),
        ("umis_stat", "<i4"),
        ("reads_stat", "<i4"),
        ("requested_genes_num", "<i4"),
        ("requested_umis_num", "<i4"),
        ("merge_target", "<i4"),
    ]
)
assert CELL_INFO_DTYPE.itemsize == 40

DIST_MAX_WORLD = 64
DIST_DONE, DIST_ALLGATHER, DIST_ALLTOALL = 0, 1, 2

# exported symbols of include/dropest_b200.h (checked by tests/test_abi.py)
EXPORTS = [
    "dge_config_default", "dge_create", "dge_destroy", "dge_last_error", "dge_add_batch", "dge_add_batch_device", "dge_add_batch_segments_device", "dge_add_batch_soa",
    "dge_add_batch_chr", "dge_add_batch_soa_chr", "dge_add_batch_chr_device", "dge_get_chr_stats",
    "dge_set_initialized", "dge_merge_and_filter", "dge_reset", "dge_set_stream", "dge_set_n_strings", "dge_set_cb_strings", "dge_get_summary", "dge_get_timings", "dge_get_cells",
    "dge_get_matrix", "dge_get_gene_order", "dge_get_merge_pairs", "dge_get_umigs", "dge_get_umi_merge_targets", "dge_get_matrix_marks", "dge_get_merge_events", "dge_edit_distance",
    "dge_hamming_distance", "dge_whitelist_shape", "dge_whitelist_token", "dge_synth_generate_device",
    "dge_route_by_barcode_device", "dge_route_count_slices_device", "dge_route_scatter_slice_device", "dge_dist_step",
    "dge_route_scatter_bounded_device", "dge_peer_alloc", "dge_peer_free", "dge_peer_open", "dge_peer_close",
    "dge_umi_first_size", "dge_umi_first_export", "dge_umi_first_import", "dge_collisions_adjusted_sizes",
]


class DgeError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"dge error {code}: {msg}")
        self.code = code


class _Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32), ("device", C.c_int32), ("cb_len", C.c_uint32), ("umi_len", C.c_uint32),
        ("n_genes", C.c_uint32), ("merge_type", C.c_uint32), ("barcodes_type", C.c_uint32), ("umi_merge_type", C.c_uint32),
        ("min_genes_before_merge", C.c_uint32), ("min_genes_after_merge", C.c_uint32),
        ("max_cb_merge_edit_distance", C.c_uint32), ("max_umi_merge_edit_distance", C.c_uint32),
        ("min_merge_fraction", C.c_double), ("max_merge_prob", C.c_double), ("max_real_merge_prob", C.c_double),
        ("umi_merge_mult", C.c_double), ("query_mark_mask", C.c_uint32), ("max_cells", C.c_int32),
        ("reads_output", C.c_uint32), ("sharded", C.c_uint32), ("barcodes_file", C.c_char_p),
        ("max_barcodes_hint", C.c_uint64), ("allow_n", C.c_uint32), ("save_umi_merge_targets", C.c_uint32),
    ]


class _Summary(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "n_reads", "total_cells_number", "real_cells_number", "filtered_cells_number", "n_genes_seen", "n_umigs",
        "intergenic_reads", "has_exon_reads", "has_intron_reads", "has_not_annotated_reads", "cm_nnz", "cm_raw_nnz",
        "n_merged", "n_excluded", "n_unresolved", "n_umis_merged", "n_umi_segments_replayed", "n_cb_merge_replayed", "n_host_flow")]


class _DistIO(C.Structure):
    """dge_dist_io (include/dropest_b200.h): one step of the cross-rank merge state machine."""
    _fields_ = [("world", C.c_uint32), ("rank", C.c_uint32), ("collective", C.c_uint32), ("stage", C.c_uint32), ("send", C.c_void_p),
                ("send_bytes", C.c_uint64 * DIST_MAX_WORLD), ("recv", C.c_void_p), ("recv_bytes", C.c_uint64 * DIST_MAX_WORLD)]


class _Timings(C.Structure):
    _fields_ = [("ms_fill", C.c_float), ("ms_init", C.c_float), ("ms_merge", C.c_float), ("ms_finish", C.c_float),
                ("ms_total", C.c_float), ("ms_dedup_kernel", C.c_float), ("n_kernel_launches", C.c_uint32),
                ("n_dedup_launches", C.c_uint32), ("ms_fill_kernel", C.c_float), ("n_fill_launches", C.c_uint32)]


class _SynthParams(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64), ("n_reads_total", C.c_uint64), ("n_cells", C.c_uint32), ("n_genes", C.c_uint32),
        ("cb_len", C.c_uint32), ("umi_len", C.c_uint32), ("cell_cdf", C.c_void_p), ("cell_barcode", C.c_void_p),
        ("cell_reads", C.c_void_p), ("gene_cdf", C.c_void_p), ("gene_weight", C.c_void_p), ("cb_error_ppm", C.c_uint32),
        ("intergenic_ppm", C.c_uint32), ("intron_ppm", C.c_uint32), ("not_annotated_ppm", C.c_uint32),
        ("reads_per_umi", C.c_uint32), ("reserved", C.c_uint32),
    ]


def lib_path() -> str:
    # DGE_LIB: development override to A/B two builds of the same library (compile-time kernel variants)
    return os.environ.get("DGE_LIB") or os.path.join(_HERE, "lib", "libdropest_b200.so")


_lib = None


def load_library():
    """Load the CUDA shared library.  Fails loudly when it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(path)
    lib.dge_config_default.argtypes = [C.POINTER(_Config)]
    lib.dge_config_default.restype = None
    lib.dge_create.argtypes = [C.POINTER(_Config), C.POINTER(C.c_void_p)]
    lib.dge_destroy.argtypes = [C.c_void_p]
    lib.dge_destroy.restype = None
    lib.dge_last_error.argtypes = [C.c_void_p]
    lib.dge_last_error.restype = C.c_char_p
    lib.dge_add_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.dge_add_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.dge_add_batch_soa.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64]
    lib.dge_set_initialized.argtypes = [C.c_void_p]
    lib.dge_merge_and_filter.argtypes = [C.c_void_p]
    lib.dge_reset.argtypes = [C.c_void_p]
    lib.dge_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    lib.dge_get_summary.argtypes = [C.c_void_p, C.POINTER(_Summary)]
    lib.dge_get_timings.argtypes = [C.c_void_p, C.POINTER(_Timings)]
    lib.dge_get_cells.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.dge_get_matrix.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    lib.dge_get_matrix_marks.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    lib.dge_get_matrix_marks.restype = C.c_int
    lib.dge_get_gene_order.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.dge_get_merge_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.dge_get_umi_merge_targets.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.dge_get_merge_events.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.dge_get_merge_events.restype = C.c_int
    lib.dge_get_umi_merge_targets.restype = C.c_int
    lib.dge_get_umigs.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.dge_edit_distance.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_uint]
    lib.dge_edit_distance.restype = C.c_uint
    lib.dge_hamming_distance.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    lib.dge_collisions_adjusted_sizes.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.POINTER(C.c_uint32)]
    lib.dge_hamming_distance.restype = C.c_uint
    lib.dge_whitelist_shape.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.c_void_p, C.c_void_p, C.c_size_t]
    lib.dge_whitelist_token.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_char_p, C.c_size_t]
    lib.dge_synth_generate_device.argtypes = [C.c_int, C.POINTER(_SynthParams), C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
    lib.dge_route_by_barcode_device.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.dge_route_count_slices_device.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_uint32, C.c_size_t, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.dge_route_scatter_slice_device.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.dge_dist_step.argtypes = [C.c_void_p, C.POINTER(_DistIO)]
    lib.dge_add_batch_segments_device.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.c_uint32]
    lib.dge_add_batch_chr.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    lib.dge_add_batch_soa_chr.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64]
    lib.dge_add_batch_chr_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    lib.dge_get_chr_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_uint32), C.c_void_p]
    lib.dge_route_scatter_bounded_device.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_size_t, C.c_void_p, C.c_void_p, C.c_uint32,
                                                     C.c_size_t, C.c_void_p, C.c_void_p]
    lib.dge_peer_alloc.argtypes = [C.c_int, C.c_size_t, C.POINTER(C.c_void_p), C.c_char_p]
    lib.dge_peer_free.argtypes = [C.c_int, C.c_void_p]
    lib.dge_peer_open.argtypes = [C.c_int, C.c_char_p, C.POINTER(C.c_void_p)]
    lib.dge_peer_close.argtypes = [C.c_int, C.c_void_p]
    lib.dge_set_n_strings.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_size_t]
    lib.dge_set_cb_strings.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]
    lib.dge_umi_first_size.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
    lib.dge_umi_first_export.argtypes = [C.c_void_p, C.c_void_p]
    lib.dge_umi_first_import.argtypes = [C.c_void_p, C.c_void_p]
    _lib = lib
    return lib


_CODE = {"A": 0, "C": 1, "G": 2, "T": 3}


def pack_seq(s: str) -> int:
    v = 0
    for ch in s:
        v = (v << 2) | _CODE[ch]
    return v


def unpack_seq(v: int, length: int) -> str:
    return "".join("ACGT"[(int(v) >> (2 * (length - 1 - i))) & 3] for i in range(length))


def marks_to_mask(code: str = "eEBA") -> int:
    """UMI::Mark::get_by_code (reference Estimation/UMI.cpp:123-154) -> bit mask over accumulated mark values."""
    table = {"e": 2, "i": 4, "E": 3, "I": 5, "B": 6, "A": 7}
    mask = 0
    for ch in code:
        if ch not in table:
            raise ValueError(f"Unexpected gene match levels: {ch}")
        mask |= 1 << table[ch]
    return mask


@dataclass
class Config:
    cb_len: int = 16
    umi_len: int = 10
    n_genes: int = 1
    device: int = 0
    merge_type: int = MERGE_NONE
    barcodes_type: int = BARCODES_INDROP
    barcodes_file: Optional[str] = None
    umi_merge_type: int = UMI_MERGE_SIMPLE
    min_genes_before_merge: int = 10
    min_genes_after_merge: int = 10
    max_cb_merge_edit_distance: int = 2
    max_umi_merge_edit_distance: int = 1
    min_merge_fraction: float = 0.2
    max_merge_prob: float = 1e-4
    max_real_merge_prob: float = 1e-7
    umi_merge_mult: float = 2.0
    marks: str = "eEBA"
    max_cells: int = -1
    reads_output: bool = False
    max_barcodes_hint: int = 0
    sharded: bool = False
    allow_n: bool = False
    save_umi_merge_targets: bool = False
    _keep: list = field(default_factory=list, repr=False)

    def to_c(self) -> _Config:
        lib = load_library()
        c = _Config()
        lib.dge_config_default(C.byref(c))
        for name in ("cb_len", "umi_len", "n_genes", "device", "merge_type", "barcodes_type", "umi_merge_type",
                     "min_genes_before_merge", "min_genes_after_merge", "max_cb_merge_edit_distance",
                     "max_umi_merge_edit_distance", "min_merge_fraction", "max_merge_prob", "max_real_merge_prob",
                     "umi_merge_mult", "max_cells", "max_barcodes_hint"):
            setattr(c, name, getattr(self, name))
        c.query_mark_mask = marks_to_mask(self.marks)
        c.reads_output = 1 if self.reads_output else 0
        c.sharded = 1 if self.sharded else 0
        c.allow_n = 1 if self.allow_n else 0
        c.save_umi_merge_targets = 1 if self.save_umi_merge_targets else 0
        if self.barcodes_file:
            b = self.barcodes_file.encode()
            self._keep.append(b)
            c.barcodes_file = b
        return c


class Container:
    """Python mirror of the reference's CellsDataContainer fill/query surface over the C ABI."""

    def __init__(self, cfg: Config):
        self._lib = load_library()
        self.cfg = cfg
        self._h = C.c_void_p()
        cc = cfg.to_c()
        rc = self._lib.dge_create(C.byref(cc), C.byref(self._h))
        if rc != 0:
            raise DgeError(rc, self._lib.dge_last_error(None).decode())
        self._keepalive = []

    def close(self):
        if self._h:
            self._lib.dge_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int):
        if rc != 0:
            raise DgeError(rc, self._lib.dge_last_error(self._h).decode())

    # ---- fill
    def add_batch(self, recs: np.ndarray):
        """Host records (numpy array of RECORD_DTYPE).  = n x CellsDataContainer::add_record."""
        recs = np.ascontiguousarray(recs, dtype=RECORD_DTYPE)
        self._check(self._lib.dge_add_batch(self._h, recs.ctypes.data, recs.shape[0]))

    def add_batch_ptr(self, host_ptr: int, n: int):
        self._check(self._lib.dge_add_batch(self._h, C.c_void_p(host_ptr), n))

    def add_batch_soa_ptr(self, keys_ptr: int, genes_ptr: int, n: int, first_read_idx: int = 0):
        self._check(self._lib.dge_add_batch_soa(self._h, C.c_void_p(keys_ptr), C.c_void_p(genes_ptr), n, first_read_idx))

    def add_batch_soa(self, keys: np.ndarray, genes: np.ndarray, first_read_idx: int = 0):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        genes = np.ascontiguousarray(genes, dtype=np.uint32)
        assert keys.shape == genes.shape
        self._check(self._lib.dge_add_batch_soa(self._h, keys.ctypes.data, genes.ctypes.data, keys.shape[0], first_read_idx))

    def add_batch_chr(self, recs: np.ndarray, chr_ids: np.ndarray):
        """Host records + the chromosome id of every read (uint8): also feeds the per-chromosome Stats counters."""
        recs = np.ascontiguousarray(recs, dtype=RECORD_DTYPE)
        chr_ids = np.ascontiguousarray(chr_ids, dtype=np.uint8)
        assert chr_ids.shape[0] == recs.shape[0]
        self._check(self._lib.dge_add_batch_chr(self._h, recs.ctypes.data, chr_ids.ctypes.data, recs.shape[0]))

    def add_batch_soa_chr(self, keys: np.ndarray, genes: np.ndarray, chr_ids: np.ndarray, first_read_idx: int = 0):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        genes = np.ascontiguousarray(genes, dtype=np.uint32)
        chr_ids = np.ascontiguousarray(chr_ids, dtype=np.uint8)
        assert keys.shape == genes.shape == chr_ids.shape
        self._check(self._lib.dge_add_batch_soa_chr(self._h, keys.ctypes.data, genes.ctypes.data, chr_ids.ctypes.data, keys.shape[0], first_read_idx))

    def add_batch_chr_device(self, dev_ptr: int, chr_dev_ptr: int, n: int, keepalive=None):
        if keepalive is not None:
            self._keepalive.append(keepalive)
        self._check(self._lib.dge_add_batch_chr_device(self._h, C.c_void_p(dev_ptr), C.c_void_p(chr_dev_ptr), n))

    def chr_stats(self):
        """(counts[n_real_cells, n_chr, 3] int32 -- exon / intron / intergenic reads, cells in CELLS_REAL order --, presented[3, n_chr] bool)"""
        n_cells, n_chr = C.c_size_t(0), C.c_uint32(0)
        self._check(self._lib.dge_get_chr_stats(self._h, None, 0, C.byref(n_cells), C.byref(n_chr), None))
        counts = np.zeros((n_cells.value, n_chr.value, 3), dtype=np.int32)
        presented = np.zeros((3, n_chr.value), dtype=np.uint8)
        if n_chr.value:
            self._check(self._lib.dge_get_chr_stats(self._h, counts.ctypes.data, n_cells.value, C.byref(n_cells), C.byref(n_chr), presented.ctypes.data))
        return counts, presented.astype(bool)

    def add_batch_device(self, dev_ptr: int, n: int, keepalive=None):
        if keepalive is not None:
            self._keepalive.append(keepalive)
        self._check(self._lib.dge_add_batch_device(self._h, C.c_void_p(dev_ptr), n))

    def add_batch_segments_device(self, dev_ptrs, counts, keepalive=None):
        """ONE fill launch over several device record arrays (local or peer-mapped); see dge_add_batch_segments_device."""
        if keepalive is not None:
            self._keepalive.append(keepalive)
        k = len(dev_ptrs)
        ptrs = (C.c_void_p * max(k, 1))(*[C.c_void_p(int(p)) for p in dev_ptrs])
        cnts = (C.c_uint64 * max(k, 1))(*[int(x) for x in counts])
        self._check(self._lib.dge_add_batch_segments_device(self._h, ptrs, cnts, k))

    def set_stream(self, cuda_stream: int):
        self._check(self._lib.dge_set_stream(self._h, C.c_void_p(cuda_stream)))

    def set_initialized(self):
        self._check(self._lib.dge_set_initialized(self._h))
        self._keepalive.clear()

    def reset(self):
        self._check(self._lib.dge_reset(self._h))
        self._keepalive.clear()

    def merge_and_filter(self):
        self._check(self._lib.dge_merge_and_filter(self._h))

    # ---- cross-rank merge steps (sharded runs); the collectives live in dropest_b200/dist.py
    def set_n_strings(self, which: int, strings):
        """The strings behind DGE_FLAG_UMI_N (which = 0) / DGE_FLAG_CB_N (which = 1) record indices (dge_set_n_strings); barcode lists
        with strings of several lengths go through dge_set_cb_strings."""
        strings = list(strings)
        if which == 1 and any(len(s) != int(self.cfg.cb_len) for s in strings):
            lengths = np.array([len(s) for s in strings], dtype=np.uint32)
            self._check(self._lib.dge_set_cb_strings(self._h, "".join(strings).encode(), lengths.ctypes.data, len(strings)))
            return
        blob = "".join(strings).encode()
        self._check(self._lib.dge_set_n_strings(self._h, which, blob, len(strings)))

    def dist_io(self, world: int, rank: int) -> "_DistIO":
        io = _DistIO()
        io.world, io.rank = world, rank
        return io

    def dist_step(self, io: "_DistIO") -> int:
        """One step of the cross-rank whitelist merge (dge_dist_step); returns io.collective."""
        self._check(self._lib.dge_dist_step(self._h, C.byref(io)))
        return int(io.collective)

    def umi_first_size(self) -> int:
        n = C.c_size_t(0)
        self._check(self._lib.dge_umi_first_size(self._h, C.byref(n)))
        return int(n.value)

    def umi_first_export(self, dst_device_ptr: int):
        self._check(self._lib.dge_umi_first_export(self._h, dst_device_ptr))

    def umi_first_import(self, src_device_ptr: int):
        self._check(self._lib.dge_umi_first_import(self._h, src_device_ptr))

    # ---- query
    def summary(self) -> dict:
        s = _Summary()
        self._check(self._lib.dge_get_summary(self._h, C.byref(s)))
        return {n: int(getattr(s, n)) for n, _ in _Summary._fields_}

    def timings(self) -> dict:
        t = _Timings()
        self._check(self._lib.dge_get_timings(self._h, C.byref(t)))
        return {n: (float(getattr(t, n)) if n.startswith("ms") else int(getattr(t, n))) for n, _ in _Timings._fields_}

    def cells(self, which: int) -> np.ndarray:
        n = C.c_size_t(0)
        self._check(self._lib.dge_get_cells(self._h, which, None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=CELL_INFO_DTYPE)
        if n.value:
            self._check(self._lib.dge_get_cells(self._h, which, out.ctypes.data, n.value, C.byref(n)))
        return out

    def matrix(self, which: int):
        """(indptr int64[n_cols+1], gene_ids int32[nnz], values int32[nnz])"""
        nc, nnz = C.c_size_t(0), C.c_size_t(0)
        self._check(self._lib.dge_get_matrix(self._h, which, None, None, None, C.byref(nc), C.byref(nnz)))
        indptr = np.zeros(nc.value + 1, dtype=np.int64)
        genes = np.zeros(nnz.value, dtype=np.int32)
        vals = np.zeros(nnz.value, dtype=np.int32)
        self._check(self._lib.dge_get_matrix(self._h, which, indptr.ctypes.data, genes.ctypes.data, vals.ctypes.data,
                                             C.byref(nc), C.byref(nnz)))
        return indptr, genes, vals

    def matrix_marks(self, marks: str):
        """The filtered matrix for another query-mark code (e.g. "e", "i", "BA": the -V matrices); same layout as matrix()"""
        mask = marks_to_mask(marks)
        nc, nnz = C.c_size_t(0), C.c_size_t(0)
        self._check(self._lib.dge_get_matrix_marks(self._h, mask, None, None, None, C.byref(nc), C.byref(nnz)))
        indptr = np.zeros(nc.value + 1, dtype=np.int64)
        genes = np.zeros(nnz.value, dtype=np.int32)
        vals = np.zeros(nnz.value, dtype=np.int32)
        self._check(self._lib.dge_get_matrix_marks(self._h, mask, indptr.ctypes.data, genes.ctypes.data, vals.ctypes.data, C.byref(nc), C.byref(nnz)))
        return indptr, genes, vals

    def matrix_shape(self, which: int):
        nc, nnz = C.c_size_t(0), C.c_size_t(0)
        self._check(self._lib.dge_get_matrix(self._h, which, None, None, None, C.byref(nc), C.byref(nnz)))
        return int(nc.value), int(nnz.value)

    def matrix_into(self, which: int, indptr_ptr: int, genes_ptr: int, vals_ptr: int):
        """Copy the matrix into caller-owned host buffers (int64[n_cols+1], int32[nnz], int32[nnz]); page-locked buffers get
        full PCIe speed."""
        nc, nnz = C.c_size_t(0), C.c_size_t(0)
        self._check(self._lib.dge_get_matrix(self._h, which, C.c_void_p(indptr_ptr), C.c_void_p(genes_ptr), C.c_void_p(vals_ptr),
                                             C.byref(nc), C.byref(nnz)))
        return int(nc.value), int(nnz.value)

    def gene_order(self) -> np.ndarray:
        n = C.c_size_t(0)
        self._check(self._lib.dge_get_gene_order(self._h, None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=np.int32)
        if n.value:
            self._check(self._lib.dge_get_gene_order(self._h, out.ctypes.data, n.value, C.byref(n)))
        return out

    def merge_pairs(self):
        n = C.c_size_t(0)
        self._check(self._lib.dge_get_merge_pairs(self._h, None, None, 0, C.byref(n)))
        a = np.zeros(n.value, dtype=np.uint64)
        b = np.zeros(n.value, dtype=np.uint64)
        if n.value:
            self._check(self._lib.dge_get_merge_pairs(self._h, a.ctypes.data, b.ctypes.data, n.value, C.byref(n)))
        return a, b

    def umi_merge_targets(self) -> dict:
        """Gene::merge_targets() of every (cell, gene): rows (cell barcode code, gene, source UMI, target UMI)."""
        n = C.c_size_t(0)
        self._check(self._lib.dge_get_umi_merge_targets(self._h, None, None, None, None, None, 0, C.byref(n)))
        m = n.value
        out = {"cb": np.zeros(m, np.uint64), "gene": np.zeros(m, np.int32), "src": np.zeros(m, np.uint32), "dst": np.zeros(m, np.uint32),
               "created": np.zeros(m, np.uint8)}
        if m:
            self._check(self._lib.dge_get_umi_merge_targets(self._h, out["cb"].ctypes.data, out["gene"].ctypes.data, out["src"].ctypes.data,
                                                            out["dst"].ctypes.data, out["created"].ctypes.data, m, C.byref(n)))
        return out

    def merge_events(self):
        """(from, to) barcode codes of the merge_cells calls in application order"""
        n = C.c_size_t(0)
        self._check(self._lib.dge_get_merge_events(self._h, None, None, 0, C.byref(n)))
        a = np.zeros(n.value, dtype=np.uint64)
        b = np.zeros(n.value, dtype=np.uint64)
        if n.value:
            self._check(self._lib.dge_get_merge_events(self._h, a.ctypes.data, b.ctypes.data, n.value, C.byref(n)))
        return a, b

    def umigs(self, which: int) -> dict:
        n = C.c_size_t(0)
        self._check(self._lib.dge_get_umigs(self._h, which, None, None, None, None, None, 0, C.byref(n)))
        m = n.value
        out = {"cell": np.zeros(m, np.uint32), "gene": np.zeros(m, np.int32), "umi": np.zeros(m, np.uint32),
               "count": np.zeros(m, np.uint32), "mark": np.zeros(m, np.uint8)}
        if m:
            self._check(self._lib.dge_get_umigs(self._h, which, out["cell"].ctypes.data, out["gene"].ctypes.data,
                                                out["umi"].ctypes.data, out["count"].ctypes.data, out["mark"].ctypes.data,
                                                m, C.byref(n)))
        return out

    def whitelist(self):
        n_parts = C.c_uint32(0)
        sizes = np.zeros(8, np.uint32)
        lens = np.zeros(8, np.uint32)
        self._check(self._lib.dge_whitelist_shape(self._h, C.byref(n_parts), sizes.ctypes.data, lens.ctypes.data, 8))
        parts = []
        buf = C.create_string_buffer(64)
        for p in range(n_parts.value):
            toks = []
            for i in range(int(sizes[p])):
                self._check(self._lib.dge_whitelist_token(self._h, p, i, buf, 64))
                toks.append(buf.value.decode())
            parts.append(toks)
        return parts


def edit_distance(a: str, b: str, skip_n: bool = True, max_ed: int = 10000) -> int:
    return int(load_library().dge_edit_distance(a.encode(), b.encode(), 1 if skip_n else 0, max_ed))


def collisions_adjusted_sizes(umi_probabilities, max_gene_expression: int, device: int = 0):
    """Tools::CollisionsAdjuster (CollisionsAdjuster.cpp:12-49) on the device: adjusted sizes for s = 1..max_gene_expression.
    Returns (uint64 array, first step that forced the exact-order rerun or 0)."""
    p = np.ascontiguousarray(umi_probabilities, dtype=np.float64)
    out = np.zeros(int(max_gene_expression), dtype=np.uint64)
    rerun = C.c_uint32(0)
    lib = load_library()
    rc = lib.dge_collisions_adjusted_sizes(device, p.ctypes.data, p.shape[0], int(max_gene_expression), out.ctypes.data, C.byref(rerun))
    if rc != 0:
        raise DgeError(rc, lib.dge_last_error(None).decode())
    return out, int(rerun.value)


def hamming_distance(a: str, b: str, skip_n: bool = True) -> int:
    return int(load_library().dge_hamming_distance(a.encode(), b.encode(), 1 if skip_n else 0))
