"""Generates the synthetic whitelist fixtures used by the parity tests (the reference's real whitelists live under
/root/reference/data/barcodes and do not travel to the GPU box).  File format = the reference's own: one line per barcode
part, whitespace separated tokens, stored reverse-complemented on load (BarcodesParser.cpp:117-144).
Run: python tests/golden/make_whitelists.py
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def tokens(rng, n, length, min_dist):
    out = []
    while len(out) < n:
        t = rng.integers(0, 4, size=length)
        if all((t != o).sum() >= min_dist for o in out):
            out.append(t)
    return ["".join("ACGT"[b] for b in t) for t in out]


def write(name, parts):
    with open(os.path.join(HERE, name), "w") as f:
        for p in parts:
            f.write(" ".join(p) + "\n")


if __name__ == "__main__":
    rng = np.random.default_rng(20261017)
    write("wl_synth_7x9.txt", [tokens(rng, 96, 7, 2), tokens(rng, 128, 9, 3)])
    write("wl_synth_8x8.txt", [tokens(rng, 64, 8, 3), tokens(rng, 64, 8, 3)])
    # tokens at Hamming distance 1 of each other: neighbours tie, distance classes >= 2 get used
    write("wl_close_4x4.txt", [["AAAA", "AAAC", "AACC", "GGGG", "GGGT", "TTTT"], ["CCCC", "CCCA", "CCAA", "TGTG", "TGTA", "ACAC"]])
    # the reference's own test fixture (data/barcodes/test_est), re-typed: 3 x 3 inDrop whitelist
    write("wl_test_est.txt", [["ATT", "TTC", "TTT"], ["TGGACCTAA", "GGCCCCTAA", "GGGACCTAA"]])
