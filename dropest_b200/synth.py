"""Synthetic read streams: table construction, the numpy mirror of csrc/synth.cu, and file output for the oracle.

Record i depends only on (seed, i) and the tables built here, so the host stream (fed to the CPU oracle as strings) and the
device stream (fed to the CUDA pipeline) are identical bit for bit -- tests/test_synth.py checks that.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from .capi import CB_N_BIT, FLAG_CB_N, FLAG_UMI_N, NO_GENE, RECORD_DTYPE, UMI_N_BIT, _SynthParams, load_library, pack_seq, unpack_seq

U64 = np.uint64
_M64 = (1 << 64) - 1


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + U64(0x9E3779B97F4A7C15)).astype(U64)
        z = x
        z = ((z ^ (z >> U64(30))) * U64(0xBF58476D1CE4E5B9)).astype(U64)
        z = ((z ^ (z >> U64(27))) * U64(0x94D049BB133111EB)).astype(U64)
        return (z ^ (z >> U64(31))).astype(U64)


def mix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = x.astype(U64)
        x = x ^ (x >> U64(33)); x = (x * U64(0xFF51AFD7ED558CCD)).astype(U64)
        x = x ^ (x >> U64(33)); x = (x * U64(0xC4CEB9FE1A85EC53)).astype(U64)
        return (x ^ (x >> U64(33))).astype(U64)


def barcode_hash(cb: np.ndarray) -> np.ndarray:
    """Host mirror of dge::barcode_hash (csrc/common.cuh)."""
    with np.errstate(over="ignore"):
        return mix64((cb.astype(U64) + U64(0x9E3779B97F4A7C15)).astype(U64))


def rank_of(cb: np.ndarray, n_ranks: int) -> np.ndarray:
    """Owner rank of a barcode: host mirror of rank_of in csrc/synth.cu (multi-GPU routing)."""
    return (((barcode_hash(cb) & U64(0xFFFFFFFF)) * U64(n_ranks)) >> U64(32)).astype(np.uint32)


def read_whitelist(path: str, indrop: bool = False) -> List[List[str]]:
    """Whitelist parts as the reference stores them: every token reverse-complemented (BarcodesParser.cpp:117-144)."""
    comp = {"A": "T", "T": "A", "G": "C", "C": "G", "N": "N"}
    parts = []
    with open(path) as f:
        for line in f:
            toks = line.split()
            if not toks:
                continue
            parts.append(["".join(comp[c] for c in reversed(t)) for t in toks])
            if indrop and len(parts) == 2:
                break
    return parts


@dataclass
class SynthSpec:
    n_reads: int
    n_cells: int
    n_genes: int
    cb_len: int = 16
    umi_len: int = 12
    seed: int = 42
    cb_error_ppm: int = 20000      # 2 % of reads get one substituted barcode base
    intergenic_ppm: int = 50000    # 5 % of reads have no gene
    intron_ppm: int = 150000       # 15 % intronic
    not_annotated_ppm: int = 50000  # 5 % not annotated, remaining 80 % exonic
    reads_per_umi: int = 4
    sigma: float = 1.0             # log-normal cell sizes
    zipf: float = 1.0              # gene weights ~ rank^-zipf
    whitelist_parts: Optional[Sequence[Sequence[str]]] = None  # product-form whitelist to draw true barcodes from


class SynthTables:
    """Sampling tables shared by host and device generators."""

    def __init__(self, spec: SynthSpec):
        self.spec = spec
        rng = np.random.default_rng(spec.seed)
        w = rng.lognormal(mean=0.0, sigma=spec.sigma, size=spec.n_cells)
        w = w / w.sum()
        cdf = np.cumsum(w)
        cell_cdf = np.minimum(np.floor(cdf * float(2 ** 64)), float(2 ** 64 - 2 ** 11)).astype(U64)
        cell_cdf[-1] = U64(_M64)
        self.cell_cdf = np.ascontiguousarray(cell_cdf)
        self.cell_reads = np.ascontiguousarray(np.maximum(1, np.round(w * spec.n_reads * (1 - spec.intergenic_ppm * 1e-6))).astype(U64))
        gw = 1.0 / np.power(np.arange(1, spec.n_genes + 1, dtype=np.float64), spec.zipf)
        gw = gw / gw.sum()
        gcdf = np.cumsum(gw)
        gene_cdf = np.minimum(np.floor(gcdf * float(2 ** 64)), float(2 ** 64 - 2 ** 11)).astype(U64)
        gene_cdf[-1] = U64(_M64)
        self.gene_cdf = np.ascontiguousarray(gene_cdf)
        self.gene_weight = np.ascontiguousarray(np.minimum(np.floor(gw * float(2 ** 32)), float(2 ** 32 - 1)).astype(np.uint32))
        # true barcodes: distinct whitelist combinations, or distinct random barcodes
        if spec.whitelist_parts is not None:
            parts = [np.array([pack_seq(t) for t in p], dtype=U64) for p in spec.whitelist_parts]
            lens = [len(p[0]) for p in spec.whitelist_parts]
            assert sum(lens) == spec.cb_len, "whitelist part lengths must add up to cb_len"
            total = 1
            for p in parts:
                total *= len(p)
            assert total >= spec.n_cells, "whitelist too small"
            if total <= 50_000_000:
                combo = rng.choice(total, size=spec.n_cells, replace=False)
            else:
                combo = np.unique(rng.integers(0, total, size=int(spec.n_cells * 1.2) + 16))
                rng.shuffle(combo)
                combo = combo[: spec.n_cells]
                assert combo.shape[0] == spec.n_cells
            cb = np.zeros(spec.n_cells, dtype=U64)
            rem = combo.astype(np.int64)
            idxs = []
            for p in reversed(parts):
                idxs.append(rem % len(p))
                rem = rem // len(p)
            idxs = list(reversed(idxs))
            for p, ln, idx in zip(parts, lens, idxs):
                cb = (cb << U64(2 * ln)) | p[idx]
            self.cell_barcode = np.ascontiguousarray(cb)
        else:
            space = 1 << (2 * spec.cb_len)
            cb = np.unique(rng.integers(0, space, size=int(spec.n_cells * 1.2) + 16, dtype=np.uint64))
            rng.shuffle(cb)
            assert cb.shape[0] >= spec.n_cells
            self.cell_barcode = np.ascontiguousarray(cb[: spec.n_cells].astype(U64))

    # ---- host generator (numpy mirror of k_synth) --------------------------------------------------------------
    def generate_host(self, first: int, count: int) -> np.ndarray:
        s = self.spec
        with np.errstate(over="ignore"):
            i = (np.arange(count, dtype=U64) + U64(first)).astype(U64)
            base = splitmix64(U64(s.seed) ^ (i * U64(0xD6E8FEB86659FD93)).astype(U64))
            r0, r1, r2, r3, r4 = (splitmix64((base + U64(k)).astype(U64)) for k in range(5))
            c = np.searchsorted(self.cell_cdf, r0, side="left").astype(np.int64)
            g = np.searchsorted(self.gene_cdf, r1, side="left").astype(np.int64)
            pool = ((self.cell_reads[c] * self.gene_weight[g].astype(U64)) >> U64(32)) // U64(s.reads_per_umi)
            pool = np.maximum(pool, U64(1))
            u_index = r2 % pool
            cg = (c.astype(U64) << U64(32)) | g.astype(U64)
            umi = splitmix64((splitmix64(cg) + u_index).astype(U64)) & U64(0xFFFFFFFF)
            umi = umi & U64((1 << (2 * s.umi_len)) - 1)
            cb = self.cell_barcode[c].copy()
            err = (r3 % U64(1000000)) < U64(s.cb_error_ppm)
            pos = (r4 % U64(s.cb_len)).astype(np.int64)
            delta = U64(1) + ((r4 >> U64(8)) % U64(3))
            shift = (2 * (s.cb_len - 1 - pos)).astype(U64)
            old = (cb >> shift) & U64(3)
            new_cb = (cb & ~(U64(3) << shift)) | (((old + delta) & U64(3)) << shift)
            cb = np.where(err, new_cb, cb)
            intergenic = ((r3 >> U64(20)) % U64(1000000)) < U64(s.intergenic_ppm)
            z = (r3 >> U64(40)) % U64(1000000)
            mark = np.where(z < U64(s.intron_ppm), 4, np.where(z < U64(s.intron_ppm + s.not_annotated_ppm), 1, 2)).astype(np.uint32)
            out = np.zeros(count, dtype=RECORD_DTYPE)
            out["key"] = (cb << U64(24)) | umi
            out["gene"] = np.where(intergenic, np.uint32(NO_GENE), g.astype(np.uint32)) | (mark << np.uint32(24))
            out["read_idx"] = i.astype(np.uint32)
        return out

    # ---- device generator -------------------------------------------------------------------------------------
    def params(self) -> _SynthParams:
        s = self.spec
        p = _SynthParams()
        p.seed = s.seed
        p.n_reads_total = s.n_reads
        p.n_cells, p.n_genes, p.cb_len, p.umi_len = s.n_cells, s.n_genes, s.cb_len, s.umi_len
        p.cell_cdf = self.cell_cdf.ctypes.data
        p.cell_barcode = self.cell_barcode.ctypes.data
        p.cell_reads = self.cell_reads.ctypes.data
        p.gene_cdf = self.gene_cdf.ctypes.data
        p.gene_weight = self.gene_weight.ctypes.data
        p.cb_error_ppm, p.intergenic_ppm = s.cb_error_ppm, s.intergenic_ppm
        p.intron_ppm, p.not_annotated_ppm = s.intron_ppm, s.not_annotated_ppm
        p.reads_per_umi = s.reads_per_umi
        return p

    def generate_device(self, device: int, first: int, count: int, out_ptr: int, stream: int = 0):
        lib = load_library()
        p = self.params()
        rc = lib.dge_synth_generate_device(device, C.byref(p), first, count, C.c_void_p(out_ptr), C.c_void_p(stream))
        if rc != 0:
            raise RuntimeError(f"dge_synth_generate_device failed with {rc}")


def product_whitelist(path: str, n1: int = 2048, n2: int = 3328, seed: int = 11, len1: int = 7, len2: int = 9):
    """Synthetic product-form whitelist file (default 7+9 bp, 2048 x 3328 tokens: the 10x v3-like list of BASELINE configs[1];
    the real 10x v3 list is not a product of parts).  Tokens of one part all have base-sum == 0 mod 4, so any
    two differ in >= 2 positions, like a real error-tolerant whitelist."""
    rng = np.random.default_rng(seed)

    def part(length, n):
        free = rng.choice(4 ** (length - 1), size=n, replace=False)
        toks = []
        for v in free:
            b = [(int(v) >> (2 * i)) & 3 for i in range(length - 1)]
            b.append((-sum(b)) % 4)
            toks.append("".join("ACGT"[x] for x in b))
        return toks

    with open(path, "w") as f:
        f.write(" ".join(part(len1, n1)) + "\n")
        f.write(" ".join(part(len2, n2)) + "\n")


def write_packed(path: str, recs: np.ndarray, cb_len: int, umi_len: int, n_genes: int, gene_names: Optional[Sequence[str]] = None,
                 n_lists: Optional["NLists"] = None, chr_ids: Optional[np.ndarray] = None):
    """DGER0001 stream for the oracle drivers (oracle/common/dge_io.h); DGER0002 (+ the two N-string lists) when `n_lists` is given;
    `chr_ids` (uint8 per record) is appended after the records and announced by the header's n_chr field."""
    blob = ("\n".join(gene_names)).encode() if gene_names else b""
    with open(path, "wb") as f:
        f.write(b"DGER0002" if n_lists is not None else b"DGER0001")
        f.write(np.array([recs.shape[0]], dtype="<u8").tobytes())
        f.write(np.array([cb_len, umi_len, n_genes, 0 if chr_ids is None else int(chr_ids.max()) + 1 if chr_ids.size else 1], dtype="<u4").tobytes())
        f.write(np.array([len(blob)], dtype="<u8").tobytes())
        f.write(blob)
        if n_lists is not None:
            for lst in (n_lists.umis, n_lists.cbs):
                b = ("\n".join(lst)).encode()
                f.write(np.array([len(b)], dtype="<u8").tobytes())
                f.write(b)
        f.write(np.ascontiguousarray(recs, dtype=RECORD_DTYPE).tobytes())
        if chr_ids is not None:
            assert chr_ids.shape[0] == recs.shape[0]
            f.write(np.ascontiguousarray(chr_ids, dtype=np.uint8).tobytes())


class NLists:
    """The caller-side lists behind DGE_FLAG_UMI_N / DGE_FLAG_CB_N record indices: equal strings get equal indices."""

    def __init__(self):
        self.umis, self.cbs = [], []
        self._u, self._c = {}, {}

    def umi_index(self, s: str) -> int:
        if s not in self._u:
            self._u[s] = len(self.umis)
            self.umis.append(s)
        return self._u[s]

    def cb_index(self, s: str) -> int:
        if s not in self._c:
            self._c[s] = len(self.cbs)
            self.cbs.append(s)
        return self._c[s]

    def pack_cb(self, s: str) -> int:
        """barcode string -> what dge_cell_info.barcode reports (escaped barcodes -- with N, or of another length -- report their list index)"""
        return (CB_N_BIT | self._c[s]) if s in self._c else pack_seq(s)

    def pack_umi(self, s: str) -> int:
        """UMI string -> what dge_get_umigs reports"""
        return (UMI_N_BIT | self._u[s]) if "N" in s else pack_seq(s)


def inject_n(recs: np.ndarray, cb_len: int, umi_len: int, umi_ppm: int, cb_ppm: int, seed: int = 1) -> "tuple[np.ndarray, NLists]":
    """Replaces, in a random subset of reads, one or two UMI bases (umi_ppm) / one barcode base (cb_ppm) by N and re-encodes those reads as
    indices into the returned N-string lists (DGE_FLAG_UMI_N / DGE_FLAG_CB_N)."""
    rng = np.random.default_rng(seed)
    out = recs.copy()
    lists = NLists()
    n = recs.shape[0]
    pick_u = np.flatnonzero(rng.integers(0, 1_000_000, n) < umi_ppm)
    pick_c = np.flatnonzero(rng.integers(0, 1_000_000, n) < cb_ppm)
    for i in pick_u:
        u = list(unpack_seq(int(recs["key"][i]) & 0xFFFFFF, umi_len))
        for p in rng.choice(umi_len, size=int(rng.integers(1, 3)), replace=False):
            u[int(p)] = "N"
        idx = lists.umi_index("".join(u))
        out["key"][i] = (int(out["key"][i]) & ~0xFFFFFF) | idx
        out["gene"][i] = int(out["gene"][i]) | FLAG_UMI_N
    for i in pick_c:
        c = list(unpack_seq(int(recs["key"][i]) >> 24, cb_len))
        c[int(rng.integers(0, cb_len))] = "N"
        idx = lists.cb_index("".join(c))
        out["key"][i] = (idx << 24) | (int(out["key"][i]) & 0xFFFFFF)
        out["gene"][i] = int(out["gene"][i]) | FLAG_CB_N
    return out, lists


def inject_odd_length_barcodes(recs: np.ndarray, cb_len: int, fraction: float, lists: "Optional[NLists]" = None, seed: int = 1):
    """Variable-length barcodes (inDrop v1 / v2): a random subset of the BARCODES (all reads of the barcode alike) loses its first base or gains
    one in front; those reads are re-encoded as indices into the escaped-barcode list (DGE_FLAG_CB_N), like barcodes with N."""
    rng = np.random.default_rng(seed)
    out = recs.copy()
    lists = lists if lists is not None else NLists()
    cbs = (recs["key"] >> np.uint64(24)).astype(np.uint64)
    plain = (recs["gene"] & np.uint32(FLAG_CB_N)) == 0
    uniq = np.unique(cbs[plain])
    chosen = uniq[rng.random(uniq.shape[0]) < fraction]
    new_of = {}
    for v in chosen:
        s = unpack_seq(int(v), cb_len)
        new_of[int(v)] = s[1:] if rng.random() < 0.5 else "ACGT"[int(rng.integers(0, 4))] + s
    for i in np.flatnonzero(plain & np.isin(cbs, chosen)):
        idx = lists.cb_index(new_of[int(cbs[i])])
        out["key"][i] = (idx << 24) | (int(out["key"][i]) & 0xFFFFFF)
        out["gene"][i] = int(out["gene"][i]) | FLAG_CB_N
    return out, lists


def records_from_strings(reads, gene_ids: dict, first_idx: int = 0) -> np.ndarray:
    """reads: iterable of (cb, umi, gene_name_or_None, mark_bits).  Gene ids are assigned by `gene_ids` (name -> id)."""
    reads = list(reads)
    out = np.zeros(len(reads), dtype=RECORD_DTYPE)
    for i, (cb, umi, gene, mark) in enumerate(reads):
        gid = NO_GENE if not gene else gene_ids.setdefault(gene, len(gene_ids))
        out["key"][i] = (pack_seq(cb) << 24) | pack_seq(umi)
        out["gene"][i] = gid | (mark << 24)
        out["read_idx"][i] = first_idx + i
    return out
