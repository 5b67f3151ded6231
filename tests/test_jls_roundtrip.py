"""CPU: encode -> decode round trips of the oracle's streams through an independent T.87 decoder
(tests/tools/jls_decoder.c).  NEAR = 0 must give the pixels back, NEAR > 0 every sample within NEAR -- for the golden
fixtures, the edge cases and random shapes.  (The reference ships no decoder; this is the conformance check that does not
go through the reference's own code.)"""
import os

import numpy as np
import pytest

from imcvt_b200 import synth
from tests._util import GOLDEN, jls_decode, load_pnm, oracle_jls
from tests.test_oracle_jls import edge_images


def planes_of(img):
    return img[None] if img.ndim == 2 else np.ascontiguousarray(np.moveaxis(img, 2, 0))


def check(img, near):
    got, near_out = jls_decode(oracle_jls(img, near))
    want = planes_of(img)
    assert got.shape == want.shape and near_out == near
    diff = np.abs(got.astype(np.int32) - want.astype(np.int32)).max() if want.size else 0
    assert diff <= near, f"max error {diff} > NEAR {near}"
    if near == 0:
        assert np.array_equal(got, want)


@pytest.mark.parametrize("near", [0, 1, 2, 3, 4])
def test_golden_fixtures_round_trip(near):
    for name in ("P5", "P6", "P2", "P3", "P4"):
        img, _ = load_pnm(os.path.join(GOLDEN, name + ".pnm"))
        check(img, near)


@pytest.mark.parametrize("near", [0, 2])
def test_edge_cases_and_random_shapes_round_trip(near):
    for _, img in edge_images():
        check(img, near)
    rng = np.random.default_rng(3)
    for _ in range(12):
        h, w = int(rng.integers(1, 70)), int(rng.integers(1, 90))
        kind = ["natural", "texture", "noise", "constant"][int(rng.integers(0, 4))]
        check(synth.frame(kind, int(rng.integers(0, 100)), h, w), near)
