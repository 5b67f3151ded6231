"""CPU: the SOURCE of the candidate-table JPEG-LS scan kernel (imcvt_b200/csrc/jls_nearscan_kernel.cuh -- the path for NEAR 1..4:
stager warps make one decision entry per sample and per candidate of the reconstructed left neighbour, one lane walks the
reconstruction chain through them, a warp codes a segment later, a warp packs the segment before that), compiled for the
host and run under the fiber SIMT emulator of tests/emu, must match the oracle and the golden fixtures bit for bit."""
import os

import numpy as np
import pytest

from imcvt_b200 import synth
from tests._util import GOLDEN, emu_jls_nearscan, golden_table, load_pnm, oracle_jls, sha256
from tests.test_oracle_jls import edge_images


@pytest.mark.parametrize("near", [1, 2, 3, 4])
def test_known_answer_p5(near):
    img, _ = load_pnm(os.path.join(GOLDEN, "P5.pnm"))
    entry = [e for e in golden_table() if e["codec"] == "jls" and e["fixture"] == "P5" and e["setting"] == near][0]
    s = emu_jls_nearscan(img, near)
    assert len(s) == entry["length"] and sha256(s) == entry["sha256"]


@pytest.mark.parametrize("near", [1, 2, 3, 4])
def test_edge_cases_match_oracle(near):
    for name, img in edge_images():
        if img.ndim == 2:
            assert emu_jls_nearscan(img, near) == oracle_jls(img, near), (name, near)


def test_rows_of_several_segments_and_run_heavy_content():
    rng = np.random.default_rng(21)
    cases = [("two-segments", synth.frame("natural", 2, 9, 1025)), ("five-segments", synth.frame("texture", 1, 5, 4100)),
             ("exact-segment", synth.frame("natural", 4, 6, 2048)), ("noise-wide", rng.integers(0, 256, (3, 5000), dtype=np.uint8)),
             ("const-wide", np.full((4, 4500), 77, np.uint8)),                                       # runs across segments and rows
             ("near-flat", (100 + rng.integers(0, 3, (20, 700))).astype(np.uint8)),                   # within NEAR: long runs
             ("dark", rng.integers(0, 4, (16, 600), dtype=np.uint8)),                                 # reconstruction clamped at 0
             ("bright", (252 + rng.integers(0, 4, (16, 600))).astype(np.uint8)),                      # ... and at 255
             ("binary", (rng.integers(0, 2, (40, 333)) * 255).astype(np.uint8)),                      # 0xFF stuffing
             ("steps", np.repeat(rng.integers(0, 256, (12, 150), dtype=np.uint8), 16, axis=1))]
    row = np.repeat(rng.integers(0, 256, 200), rng.integers(1, 300, 200))[:6000].astype(np.uint8)
    cases.append(("random-runs", np.stack([np.roll(row, 37 * i) for i in range(4)])))
    for name, img in cases:
        for near in (1, 2, 3, 4):
            assert emu_jls_nearscan(img, near) == oracle_jls(img, near), (name, near)


def test_random_shapes():
    rng = np.random.default_rng(5)
    for _ in range(30):
        h, w = int(rng.integers(1, 40)), int(rng.integers(1, 1300))
        kind = int(rng.integers(0, 3))
        if kind == 0:
            img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        elif kind == 1:
            img = (rng.integers(0, 6, (h, w)) * 51).astype(np.uint8)
        else:
            img = np.clip(128 + np.cumsum(rng.integers(-3, 4, (h, w)), axis=1), 0, 255).astype(np.uint8)
        near = int(rng.integers(1, 5))
        assert emu_jls_nearscan(img, near) == oracle_jls(img, near), (h, w, kind, near)
