"""Tools::CollisionsAdjuster (reference Tools/CollisionsAdjuster.cpp:12-49): oracle port and CUDA path against golden vectors of the
compiled, unmodified reference (tests/golden/collisions.npz, made by tests/golden/make_collisions_golden.py).  Integer outputs:
bit-exact bar."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_collisions_golden as mcg  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "collisions.npz")
PORT = os.path.join(ROOT, "oracle", "_build", "collisions_port")
CASES = ["uniform4096", "zipf256", "random65536", "spiky1024", "single"]


def golden():
    with np.load(GOLDEN) as z:
        return {k: z[k] for k in z.files}


def test_golden_holds_the_reference_probe_values():
    """SURVEY.md A9 / tests/golden/ref_pins.json: uniform 4096-UMI space -> adjusted(100) = 101, (1000) = 1146, (3000) = 5400."""
    g = golden()
    a = g["adjusted_uniform4096"]
    assert (int(a[99]), int(a[999]), int(a[2999])) == (101, 1146, 5400)
    pins = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_pins.json")))["collisionsAdjuster"]
    assert [int(a[s - 1]) for s in (1, 10, 100, 500, 1000, 2000, 3000)] == pins["uniform4096"]
    assert [int(g["adjusted_zipf256"][s - 1]) for s in (1, 5, 20, 50, 100, 150)] == pins["zipf256"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_port_matches_reference(name):
    assert os.path.exists(PORT), "build the oracle port first: make -C oracle port"
    g = golden()
    got = mcg.run_ref(g["p_" + name], g["adjusted_" + name].shape[0], binary=PORT)
    np.testing.assert_array_equal(got, g["adjusted_" + name])


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_matches_reference(name):
    import dropest_b200 as dg

    g = golden()
    got, rerun = dg.collisions_adjusted_sizes(g["p_" + name], g["adjusted_" + name].shape[0])
    np.testing.assert_array_equal(got, g["adjusted_" + name])


@pytest.mark.gpu
def test_cuda_exact_order_path(monkeypatch):
    """The sequential-summation rerun (taken when the parallel sum comes too close to a rounding boundary) forced on."""
    import dropest_b200 as dg

    monkeypatch.setenv("DGE_CA_FORCE_EXACT", "1")
    g = golden()
    for name in ("zipf256", "spiky1024"):
        got, _ = dg.collisions_adjusted_sizes(g["p_" + name], g["adjusted_" + name].shape[0])
        np.testing.assert_array_equal(got, g["adjusted_" + name])
