"""The oracle's restated arithmetic ("port") against golden vectors produced by the reference's own
headers (tests/golden/make_golden.py).  Bit-exact: both are host fp32 builds with
-ffp-contract=off, so any difference is a restatement error, not rounding."""
from pathlib import Path

import numpy as np
import pytest

import kat

GOLD = Path(__file__).parent / "golden"


def _same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.dtype.kind == "f":
        return np.array_equal(a.view(np.uint32), b.view(np.uint32)) or np.array_equal(a, b, equal_nan=True)
    return np.array_equal(a, b)


@pytest.fixture(scope="module")
def port_kat(port_lib):
    return kat.run(port_lib)


def test_kat_keys_cover_all_functions(port_kat):
    gold = np.load(GOLD / "kat_reference.npz")
    assert set(gold.files) == set(port_kat.keys())


@pytest.mark.parametrize("key", sorted(np.load(GOLD / "kat_reference.npz").files))
def test_port_matches_reference_golden(port_kat, key):
    gold = np.load(GOLD / "kat_reference.npz")
    assert _same(port_kat[key], gold[key]), f"{key}: port differs from the reference-header golden"


def test_rng_known_values(port_kat):
    # integer-exact stream: state after TEA(4 rounds) on (pixel 0, seed 0), first LCG outputs in [0,1)
    s = port_kat["rng_stream"]
    assert s.min() >= 0.0 and s.max() < 1.0
    # (s * 2^24) must be integers: 24-bit mantissa draws (cuda/random.h:35-37)
    assert np.all(np.mod(s.astype(np.float64) * 2 ** 24, 1.0) == 0)


def test_render_goldens(port_lib):
    gold = np.load(GOLD / "render_reference.npz")
    got = kat.run_renders(port_lib)
    assert set(gold.files) == set(got.keys())
    for k in gold.files:
        assert _same(got[k], gold[k]), f"{k}: port render differs from the reference-header render"
