#!/usr/bin/env python
"""BAM -> count matrix rate through dropest_b200/host/BamIngest + the container (GPU box): python tools/bam_rate.py [reads] [copies]
Writes a synthetic 10x-like BAM (CB / UB / GX / CQ / UQ tags, 91-base reads) with tests/bam_utils.py and times lib/test_bam_pipeline on it;
`copies` passes the same file that many times (the Python writer is slow: 2 M reads x 20 copies = the 40 M-read run of
profiles/r2_bam_ingest_final_x20.log; the one-time device setup of ~0.6 s then no longer dominates)."""
import os, subprocess, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from bam_utils import alignment, write_bam

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
copies = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(0)
cbs = ["".join(rng.choice(list("ACGT"), 16)) for _ in range(3000)]
umis = ["".join(rng.choice(list("ACGT"), 10)) for _ in range(4096)]
als = []
for i in range(n):
    als.append(alignment(f"A00123:45:HXXXX:1:{1100 + i % 100}:{i}:{i % 9999}", i % 3, i, 0,
                         [("NH", ("i", 1)), ("CB", ("Z", cbs[int(rng.integers(0, 3000))])), ("UB", ("Z", umis[int(rng.integers(0, 4096))])),
                          ("GX", ("Z", f"ENSG{int(rng.integers(0, 20000)):011d}")), ("CQ", ("Z", "I" * 16)), ("UQ", ("Z", "I" * 10))], seq_len=91))
path = "/tmp/bam_rate.bam"
write_bam(path, [("chr1", 1 << 28), ("chr2", 1 << 28), ("chrM", 16000)], als, block_bytes=65000)
print("BAM:", n, "reads,", round(os.path.getsize(path) / 1e6, 1), "MB")
exe = os.path.join(ROOT, "dropest_b200", "lib", "test_bam_pipeline")
for rep in range(2):
    r = subprocess.run([exe, "-", "5", "5", "", "", ""] + [path] * copies, capture_output=True, text=True)
    print([l for l in r.stdout.split("\n") if l.startswith(("timing", "stats", "error"))], r.stderr[-200:])
