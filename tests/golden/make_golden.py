"""Regenerates the golden fixtures from the COMPILED UNMODIFIED REFERENCE (oracle/_ref, built from /root/reference by
oracle/Makefile).  Run in the build container only: python tests/golden/make_golden.py
  tests/golden/ref_pins.json      known-answer values of the reference's own unit tests, as observed on the compiled reference
  tests/golden/case_*.npz         full canonical outputs of oracle/_ref/dropest_ref for small seeded workloads
Inputs are not stored: they are re-derived from the seeds in golden_cases.py (host generator, bit-exact by construction).
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import oracle_io  # noqa: E402
import golden_cases  # noqa: E402

if __name__ == "__main__":
    assert oracle_io.available("reference"), "build oracle/_ref first (make -C oracle ref)"
    pins = subprocess.run([oracle_io.REF_PINS_BIN, "/root/reference/data"], check=True, capture_output=True, text=True).stdout
    with open(os.path.join(HERE, "ref_pins.json"), "w") as f:
        f.write(pins)
    for name, case in golden_cases.cases().items():
        res = golden_cases.run_oracle_on(case, kind="reference")
        keep = {k: v for k, v in res.items() if not k.startswith("_") and not k.startswith("t_")}
        np.savez_compressed(os.path.join(HERE, f"case_{name}.npz"), **keep)
        print(name, {k: v.shape for k, v in keep.items() if v.size > 1 and k in ("cell_flags", "cm_val", "umi_count")})
