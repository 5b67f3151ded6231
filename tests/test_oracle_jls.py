"""CPU: pins oracle/jls_oracle.c (the restatement) to the reference.

 1. known answers: tests/golden/known_answers.json (made by tests/golden/make_golden.py from the compiled
    reference) must agree with the table printed in SURVEY.md 8(c) (reference CLI + shipped binary), and the
    oracle must reproduce every entry byte for byte (sha256 of the whole stream);
 2. oracle == compiled reference (oracle/_ref) on random and edge-case images, NEAR 0..4, gray and RGB.
"""
import os
import numpy as np
import pytest
from tests._util import GOLDEN, golden_table, load_pnm, oracle_jls, ref_jls, ref_lib, sha256

# (length, sha256[:16]) from SURVEY.md 8(c) -- produced by the reference CLI, independent of make_golden.py
SURVEY_JLS = {
    "P5": [(33388, "57f1959687d927c3"), (21703, "448583aae7f08ecf"), (17035, "f7e7c0953e0d5548"), (14419, "4f57dd7bb57a4566"), (12640, "4a34f2fa91e96371")],
    "P6": [(101008, "442874ef73fc6a92"), (66269, "c04a6ba16d0185da"), (52124, "601716f4fdaf3fc4"), (44193, "d4df02fda365cb4c"), (38800, "6368173a31929c41")],
    "P2": [(84, "3501d606c52d352d"), (69, "22cd26cc112a1316"), (62, "968704d615129472"), (54, "b1539f2b3563c980"), (54, "29314fdbf86d7ec8")],
    "P3": [(92, "65151b6bcc173bf1"), (88, "0ccfb7c9425f36ae"), (87, "088d97e7a7aef4e0"), (80, "2e155f8e5fad8ad0"), (77, "6db9c208de7ea968")],
    "P4": [(101, "b9fe0efbbd6e5349"), (95, "4c2b2fb62492c025"), (95, "a985ec24a40308f6"), (119, "895fc895460ac900"), (119, "fa634183a958b421")],
}
P5_NEAR0_FULL = "57f1959687d927c328432713ea152d5c5e39eac8ae74792051d5df95dfea2843"


def test_golden_table_matches_survey():
    tab = {(e["fixture"], e["setting"]): e for e in golden_table() if e["codec"] == "jls"}
    for name, rows in SURVEY_JLS.items():
        for near, (length, sha) in enumerate(rows):
            e = tab[(name, near)]
            assert (e["length"], e["sha256"][:16]) == (length, sha)
    assert tab[("P5", 0)]["sha256"] == P5_NEAR0_FULL


@pytest.mark.parametrize("entry", [e for e in golden_table() if e["codec"] == "jls"],
                         ids=lambda e: f"{e['fixture']}-near{e['setting']}")
def test_oracle_reproduces_known_answers(entry):
    img, _ = load_pnm(os.path.join(GOLDEN, entry["fixture"] + ".pnm"))
    s = oracle_jls(img, entry["setting"])
    assert len(s) == entry["length"]
    assert sha256(s) == entry["sha256"]


def edge_images():
    rng = np.random.default_rng(7)
    yield "noise131x97", rng.integers(0, 256, (97, 131), dtype=np.uint8)
    yield "1x1", np.array([[200]], dtype=np.uint8)
    yield "1col", rng.integers(0, 256, (37, 1), dtype=np.uint8)
    yield "1row", rng.integers(0, 256, (1, 53), dtype=np.uint8)
    yield "const", np.full((40, 70), 93, dtype=np.uint8)
    yield "zeros", np.zeros((33, 65), dtype=np.uint8)
    yield "ff-stuffing", (rng.integers(0, 2, (64, 96)) * 255).astype(np.uint8)
    yy, xx = np.mgrid[0:80, 0:120]
    yield "posterised", ((np.sin(xx / 9.0) + np.cos(yy / 7.0)) * 3).astype(np.int64).astype(np.uint8) * 20
    yield "long-runs", np.repeat(rng.integers(0, 256, (50, 8), dtype=np.uint8), 40, axis=1)
    yield "rgb-noise", rng.integers(0, 256, (31, 45, 3), dtype=np.uint8)
    yield "rgb-smooth", np.stack([(xx + yy) // 2, xx, yy], axis=-1).astype(np.uint8)
    yield "wide", rng.integers(100, 110, (3, 4099), dtype=np.uint8)


@pytest.mark.parametrize("near", [0, 1, 2, 3, 4])
def test_oracle_equals_compiled_reference(near):
    if ref_lib() is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    for name, img in edge_images():
        assert oracle_jls(img, near) == ref_jls(img, near), name
