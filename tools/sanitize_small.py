"""Small encodes of every device path, meant to run under compute-sanitizer (memcheck / racecheck) on a GPU box:
    compute-sanitizer --tool memcheck python tools/sanitize_small.py
Outputs are still compared with the oracle."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import imcvt_b200
from imcvt_b200 import synth
from tests._util import oracle_hevc, oracle_jls

imcvt_b200.init(0)
L = imcvt_b200.lib()
gray = synth.batch("natural", 3, 40, 200)
# JPEG-LS NEAR = 0: cluster pair (default), one CTA per plane, multi-kernel pipeline; NEAR = 2: three-warp row scan; RGB
for path in (0, 2, 1):
    imcvt_b200.debug_jls_path(path)
    got = imcvt_b200.encode_jls_batch(gray, 0)
    assert all(got[i] == oracle_jls(gray[i], 0) for i in range(3)), path
imcvt_b200.debug_jls_path(0)
# NEAR > 0: the candidate-table kernel (rows of one and of several segments, every NEAR), and the three-warp row scan (path 5)
wide = synth.batch("texture", 2, 6, 1100)
for near in (1, 2, 3, 4):
    got = imcvt_b200.encode_jls_batch(gray, near)
    assert all(got[i] == oracle_jls(gray[i], near) for i in range(3)), near
    got = imcvt_b200.encode_jls_batch(wide, near)
    assert all(got[i] == oracle_jls(wide[i], near) for i in range(2)), near
imcvt_b200.debug_jls_path(5)
got = imcvt_b200.encode_jls_batch(gray, 2)
assert all(got[i] == oracle_jls(gray[i], 2) for i in range(3))
imcvt_b200.debug_jls_path(0)
rgb = synth.batch("texture", 2, 24, 70, rgb=True)
got = imcvt_b200.encode_jls_batch(rgb, 1)
assert all(got[i] == oracle_jls(rgb[i], 1) for i in range(2))
# page-locked buffers coded in place (rows requested ahead, tile rows pulled over DSMEM, 128-byte line ring)
n, h, w = gray.shape
stride = imcvt_b200.jls_bound(0, h, w)
h_in = torch.from_numpy(gray).pin_memory()
h_out = torch.zeros(n * stride, dtype=torch.uint8).pin_memory()
lens = (C.c_size_t * n)()
assert L.imcvtEncodeBatchJLS(h_in.numpy().ctypes.data, h * w, n, 0, h, w, 0, h_out.numpy().ctypes.data, stride, C.cast(lens, C.c_void_p)) == 0
assert all(h_out.numpy()[i * stride:i * stride + lens[i]].tobytes() == oracle_jls(gray[i], 0) for i in range(n))
# H.265: both launch shapes, a sweep, the green plane
small = synth.batch("natural", 2, 40, 72)
for path in (1, 2):
    imcvt_b200.debug_hevc_path(path)
    s, r = imcvt_b200.encode_hevc_batch(small, 3)
    for i in range(2):
        want, want_r = oracle_hevc(small[i], 3)
        assert s[i] == want and np.array_equal(r[i], want_r), path
imcvt_b200.debug_hevc_path(0)
sw = imcvt_b200.encode_hevc_sweep(small[:1], [4, 1])
assert sw[0][0] == oracle_hevc(small[0], 4)[0] and sw[0][1] == oracle_hevc(small[0], 1)[0]
g = imcvt_b200.encode_hevc_green_batch(rgb, 2)
assert g[0] == oracle_hevc(rgb[0, :, :, 1], 2)[0]
imcvt_b200.shutdown()
print("sanitize_small: all paths ran and match the oracle")
