"""CPU: the SOURCE of the CUDA JPEG-LS tile kernel (imcvt_b200/csrc/jls_tile_kernel.cuh: producers / chains / packer
software pipeline over shared-memory tiles), compiled for the host and run under the fiber SIMT emulator of
tests/emu, must match the oracle and the golden fixtures bit for bit."""
import os

import numpy as np

from imcvt_b200 import synth
from tests._util import GOLDEN, emu_jls, golden_table, load_pnm, oracle_jls, sha256
from tests.test_oracle_jls import edge_images


import pytest


@pytest.mark.parametrize("pair", [False, True], ids=["one-cta", "cluster-pair"])
def test_known_answer_p5(pair):
    img, _ = load_pnm(os.path.join(GOLDEN, "P5.pnm"))
    entry = [e for e in golden_table() if e["codec"] == "jls" and e["fixture"] == "P5" and e["setting"] == 0][0]
    s = emu_jls(img, pair)
    assert len(s) == entry["length"] and sha256(s) == entry["sha256"]


@pytest.mark.parametrize("pair", [False, True], ids=["one-cta", "cluster-pair"])
def test_edge_cases_match_oracle(pair):
    for name, img in edge_images():
        if img.ndim == 2:
            assert emu_jls(img, pair) == oracle_jls(img, 0), name


@pytest.mark.parametrize("pair", [False, True], ids=["one-cta", "cluster-pair"])
def test_multi_tile_planes_match_oracle(pair):
    rng = np.random.default_rng(8)
    cases = [(kind, synth.frame(kind, 3, 200, 480)) for kind in ["natural", "texture", "noise", "constant"]]
    cases.append(("binary", (rng.integers(0, 2, (90, 700)) * 255).astype(np.uint8)))          # 0xFF stuffing across chunks
    cases.append(("few-levels", (rng.integers(0, 4, (120, 333)) * 85).astype(np.uint8)))      # runs + interruptions
    cases.append(("widest", rng.integers(0, 256, (5, 6144), dtype=np.uint8)))                  # one row per tile
    cases.append(("steps", np.repeat(rng.integers(0, 256, (30, 375), dtype=np.uint8), 16, axis=1)))
    # rows split into segments across producer warps: runs that cross segment boundaries
    cases.append(("const-wide", np.full((4, 5000), 77, np.uint8)))
    row = np.repeat(rng.integers(0, 256, 200), rng.integers(1, 400, 200))[:6000].astype(np.uint8)
    cases.append(("random-runs", np.stack([np.roll(row, 37 * i) for i in range(6)])))
    img = np.repeat(rng.integers(0, 3, (7, 40), dtype=np.uint8) * 100, 50, axis=1)[:, :1999].copy()
    img[rng.integers(0, 7, 30), rng.integers(0, 1999, 30)] += 1
    cases.append(("runs+interrupts", img))
    for name, img in cases:
        assert emu_jls(img, pair) == oracle_jls(img, 0), name


@pytest.mark.parametrize("pair", [False, True], ids=["one-cta", "cluster-pair"])
def test_line_ring_output_matches(pair):
    """The output path for page-locked host slots (imcvtJLSHostMode 2): stuffed bytes leave the packer as whole 128-byte
    lines through a shared-memory ring; the slot starts inside a line (offset 37 + 25 header bytes)."""
    rng = np.random.default_rng(11)
    cases = [("natural", synth.frame("natural", 5, 100, 480)), ("noise", synth.frame("noise", 1, 40, 333)),
             ("constant", synth.frame("constant", 0, 64, 64)), ("tiny", rng.integers(0, 256, (1, 1), dtype=np.uint8)),
             ("binary", (rng.integers(0, 2, (60, 700)) * 255).astype(np.uint8)), ("short", rng.integers(0, 256, (3, 17), dtype=np.uint8))]
    for name, img in cases:
        assert emu_jls(img, pair, line_out=True) == oracle_jls(img, 0), name
