"""oracle/oracle_io.py -- TEST INFRASTRUCTURE ONLY: run the CPU oracles and read their DGEO0001 output.

Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
Two oracles exist:
  * oracle/_ref/dropest_ref   -- the UNMODIFIED reference hot path compiled with shims (oracle/Makefile `ref`); kind "reference"
  * oracle/_build/dropest_port -- our CPU restatement (oracle/port); kind "port"
"""
from __future__ import annotations

import os
import subprocess
import tempfile
from typing import Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_BIN = os.path.join(_HERE, "_ref", "dropest_ref")
REF_PINS_BIN = os.path.join(_HERE, "_ref", "ref_pins")
PORT_BIN = os.path.join(_HERE, "_build", "dropest_port")

_DTYPES = {0: np.uint8, 1: np.int32, 2: np.uint32, 3: np.int64, 4: np.uint64, 5: np.float64}


def read_dgeo(path: str) -> Dict[str, np.ndarray]:
    with open(path, "rb") as f:
        buf = f.read()
    assert buf[:8] == b"DGEO0001", "bad oracle output"
    n = int(np.frombuffer(buf, "<u8", 1, 8)[0])
    out = {}
    for i in range(n):
        o = 16 + i * 56
        name = buf[o:o + 32].split(b"\0", 1)[0].decode()
        dtype = int(np.frombuffer(buf, "<u4", 1, o + 32)[0])
        count = int(np.frombuffer(buf, "<u8", 1, o + 40)[0])
        off = int(np.frombuffer(buf, "<u8", 1, o + 48)[0])
        out[name] = np.frombuffer(buf, _DTYPES[dtype], count, off).copy()
    return out


def strings(arr: np.ndarray):
    s = arr.tobytes().decode()
    return s.split("\n")[:-1] if s else []


def available(kind: str = "any") -> bool:
    if kind == "reference":
        return os.path.exists(REF_BIN)
    if kind == "port":
        return os.path.exists(PORT_BIN)
    return os.path.exists(REF_BIN) or os.path.exists(PORT_BIN)


def oracle_binary(kind: str = "any") -> str:
    if kind in ("reference", "any") and os.path.exists(REF_BIN):
        return REF_BIN
    if kind in ("port", "any") and os.path.exists(PORT_BIN):
        return PORT_BIN
    raise FileNotFoundError("no oracle binary built: run `make -C oracle ref port`")


def run_oracle(in_path: str, kind: str = "any", merge: str = "none", barcodes: Optional[str] = None,
               barcodes_type: str = "const", min_genes_before: int = 10, min_genes_after: int = 10, max_cb_ed: int = 2,
               min_frac: float = 0.2, marks: str = "eEBA", max_cells: int = -1, reads_output: bool = False,
               dump_umis: bool = False, umi_merge: str = "simple", max_umi_ed: int = 1, umi_mult: float = 2.0, limit: int = 0,
               max_merge_prob: float = 1e-4, max_real_merge_prob: float = 1e-7, init_only: bool = False,
               timeout: Optional[float] = None, dump_rpupc: bool = False) -> Dict[str, np.ndarray]:
    exe = oracle_binary(kind)
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "out.dgeo")
        cmd = [exe, "--in", in_path, "--out", out, "--merge", merge, "--min-genes-before", str(min_genes_before),
               "--min-genes-after", str(min_genes_after), "--max-cb-ed", str(max_cb_ed), "--min-frac", repr(min_frac),
               "--marks", marks, "--max-cells", str(max_cells), "--umi-merge", umi_merge, "--max-umi-ed", str(max_umi_ed), "--umi-mult", repr(float(umi_mult)),
               "--max-merge-prob", repr(max_merge_prob), "--max-real-merge-prob", repr(max_real_merge_prob)]
        if barcodes:
            cmd += ["--barcodes", barcodes, "--barcodes-type", barcodes_type]
        if reads_output:
            cmd.append("--reads-output")
        if dump_umis:
            cmd.append("--dump-umis")
        if dump_rpupc:   # the compiled reference only (reads_per_umi_per_cell with UMI::mean_quality)
            cmd.append("--dump-rpupc")
        if init_only:
            cmd.append("--init-only")
        if limit:
            cmd += ["--limit", str(limit)]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
        if r.returncode != 0:
            raise RuntimeError(f"oracle failed: {' '.join(cmd)}\n{r.stderr}")
        res = read_dgeo(out)
        res["_stderr"] = r.stderr
        res["_kind"] = "reference" if exe == REF_BIN else "port"
        return res
