"""Golden vectors for Tools::CollisionsAdjuster from the COMPILED UNMODIFIED REFERENCE (oracle/_ref/ref_collisions, built from
/root/reference by oracle/Makefile).  Run in the build container only: python tests/golden/make_collisions_golden.py
Writes tests/golden/collisions.npz: for every case the probability vector (exact float64 bits) and the adjusted sizes 1..max."""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "ref_collisions")


def cases():
    rng = np.random.default_rng(5)
    out = {}
    out["uniform4096"] = (np.full(4096, 1.0 / 4096), 3000)                     # 6 bp UMIs, SURVEY.md A9 probe values
    z = 1.0 / (1.0 + np.arange(256)); out["zipf256"] = (z / z.sum(), 150)
    w = rng.random(65536) ** 3 + 1e-3; out["random65536"] = (w / w.sum(), 400)  # Drop-seq 8 bp UMI space, skewed
    w = rng.random(1024); w[:4] += 200.0; out["spiky1024"] = (w / w.sum(), 600) # a few dominant UMIs: strong collisions early
    out["single"] = (np.array([1.0 - 1e-9]), 5)                                  # degenerate: one UMI
    return out


def run_ref(p: np.ndarray, max_size: int, binary: str = REF) -> np.ndarray:
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "p.f64")
        np.ascontiguousarray(p, dtype="<f8").tofile(path)
        txt = subprocess.run([binary, path, str(max_size)], check=True, capture_output=True, text=True).stdout
    return np.array([int(x) for x in txt.split()], dtype=np.uint64)


if __name__ == "__main__":
    assert os.path.exists(REF), "build oracle/_ref first (make -C oracle ref)"
    keep = {}
    for name, (p, max_size) in cases().items():
        keep["p_" + name] = np.ascontiguousarray(p, dtype=np.float64)
        keep["adjusted_" + name] = run_ref(p, max_size)
        print(name, p.shape[0], max_size, keep["adjusted_" + name][[0, max_size // 2, max_size - 1]])
    np.savez_compressed(os.path.join(HERE, "collisions.npz"), **keep)
