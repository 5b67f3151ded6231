"""CPU: pins oracle/hevc_oracle.c (the restatement) to the reference: known answers of SURVEY.md 8(c) and
byte-for-byte equality (stream and reconstruction) with the compiled reference on edge-case images."""
import os
import numpy as np
import pytest
from tests._util import GOLDEN, golden_table, load_pnm, oracle_hevc, ref_hevc, ref_lib, sha256

SURVEY_HEVC = {
    "P5": [(33045, "ca32937104b8aeed"), (22435, "8c6c7d5241e0b3ee"), (15077, "c5405305d9386ef0"), (9771, "f142528c4fee7d33"), (5670, "0eb0c7c2e82dc34a")],
    "P6": [(32578, "e6b6bb1eaf2dbf49"), (22249, "1f14265660abe0a0"), (15097, "69b7fd272ebb3cc1"), (9869, "7d3e7cad78c2c0da"), (5718, "63434ae2e4c8410e")],
    "P2": [(191, "1312f3c9de91eae0"), (163, "be9ae87dcbee4a0b"), (139, "188ebdeae836d4bf"), (117, "4aad2dd21f43b533"), (89, "51abf694da17c29c")],
}
P5_Q0_FULL = "ca32937104b8aeed9c9b9a3281e19407d88d7b18083ab6729523541d659a52e6"


def fixture_gray(name):
    img, is_rgb = load_pnm(os.path.join(GOLDEN, name + ".pnm"))
    return img[:, :, 1] if is_rgb else img          # imageio_hevc.c:24-26: green channel of RGB


def test_golden_table_matches_survey():
    tab = {(e["fixture"], e["setting"]): e for e in golden_table() if e["codec"] == "hevc"}
    for name, rows in SURVEY_HEVC.items():
        for q, (length, sha) in enumerate(rows):
            e = tab[(name, q)]
            assert (e["length"], e["sha256"][:16]) == (length, sha)
    assert tab[("P5", 0)]["sha256"] == P5_Q0_FULL


@pytest.mark.parametrize("entry", [e for e in golden_table() if e["codec"] == "hevc" and (e["fixture"] == "P2" or e["setting"] in (0, 4))],
                         ids=lambda e: f"{e['fixture']}-qpd6_{e['setting']}")
def test_oracle_reproduces_known_answers(entry):
    s, _ = oracle_hevc(fixture_gray(entry["fixture"]), entry["setting"])
    assert len(s) == entry["length"]
    assert sha256(s) == entry["sha256"]


def hevc_edge_images():
    rng = np.random.default_rng(5)
    yield "noise40x72", rng.integers(0, 256, (40, 72), dtype=np.uint8)
    yield "1x1", np.array([[77]], dtype=np.uint8)
    yield "const33x65", np.full((33, 65), 200, dtype=np.uint8)
    yy, xx = np.mgrid[0:64, 0:96]
    yield "ramp", ((xx * 2 + yy * 3) & 255).astype(np.uint8)
    yield "edges", ((xx // 7 + yy // 5) % 2 * 230 + 10).astype(np.uint8)
    yield "binary", (rng.integers(0, 2, (32, 64)) * 255).astype(np.uint8)
    yield "smooth+noise", np.clip(128 + 60 * np.sin(xx / 9.0) + rng.normal(0, 3, xx.shape), 0, 255).astype(np.uint8)


@pytest.mark.parametrize("qpd6", [0, 1, 2, 3, 4])
def test_oracle_equals_compiled_reference(qpd6):
    if ref_lib() is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    for name, img in hevc_edge_images():
        s, r = oracle_hevc(img, qpd6)
        s2, r2 = ref_hevc(img, qpd6)
        assert s == s2, name
        assert np.array_equal(r, r2), name
