"""Regenerates tests/golden/known_answers.json from the UNMODIFIED reference compiled into
oracle/_ref/libimcvt_ref.so (run in the build container, where /root/reference exists):

    python tests/golden/make_golden.py

Each entry: fixture, codec, setting (NEAR or qpd6), stream length, sha256 of the stream.
The (length, sha256[:16]) pairs must equal the table in SURVEY.md 8(c), which was produced
independently (reference CLI + shipped static binary) -- tests/test_oracle_jls.py checks that too.
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests._util import GOLDEN, load_pnm, ref_jls, ref_hevc, sha256

out = []
for name in ["P2", "P3", "P4", "P5", "P6"]:
    img, is_rgb = load_pnm(os.path.join(GOLDEN, name + ".pnm"))
    for near in range(5):
        s = ref_jls(img, near)
        out.append(dict(fixture=name, codec="jls", setting=near, length=len(s), sha256=sha256(s)))
for name in ["P2", "P5", "P6"]:
    img, is_rgb = load_pnm(os.path.join(GOLDEN, name + ".pnm"))
    if is_rgb:
        img = img[:, :, 1]                      # imageio_hevc.c:24-26 encodes the green channel
    for q in range(5):
        s, _ = ref_hevc(img, q)
        out.append(dict(fixture=name, codec="hevc", setting=q, length=len(s), sha256=sha256(s)))
json.dump(out, open(os.path.join(GOLDEN, "known_answers.json"), "w"), indent=1)
print(len(out), "entries")
