"""A few CTUs of H.265 on both launch shapes, for `compute-sanitizer --tool racecheck --racecheck-report hazard`.
Arguments: height width (default 32 32)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import imcvt_b200
from imcvt_b200 import synth
from tests._util import oracle_hevc
h = int(sys.argv[1]) if len(sys.argv) > 1 else 32
w = int(sys.argv[2]) if len(sys.argv) > 2 else 32
imcvt_b200.init(0)
img = synth.batch("natural", 1, h, w)
want = oracle_hevc(img[0], 3)[0]
for path in (1, 2):                     # one CTA per frame, cluster pair
    imcvt_b200.debug_hevc_path(path)
    s, _ = imcvt_b200.encode_hevc_batch(img, 3)
    assert s[0] == want, path
print("ok")
