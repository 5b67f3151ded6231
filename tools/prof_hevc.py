"""Phase profile of hevc_frame_kernel (development aid; needs `make -C imcvt_b200/csrc PROF=1`).
   python tools/prof_hevc.py [kind] [n] [qpd6] [h] [w] [path]"""
import ctypes as C, os, sys
os.environ.setdefault("IMCVT_B200_LIB", "libimcvt_b200_prof.so")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import imcvt_b200
from imcvt_b200 import synth

kind = sys.argv[1] if len(sys.argv) > 1 else "natural"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
q = int(sys.argv[3]) if len(sys.argv) > 3 else 4
h = int(sys.argv[4]) if len(sys.argv) > 4 else 128
w = int(sys.argv[5]) if len(sys.argv) > 5 else 256
path = int(sys.argv[6]) if len(sys.argv) > 6 else 0          # imcvtDebugHEVCPath: launch shape | search form << 2
frames = synth.batch(kind, n, h, w)
imcvt_b200.init(0)
imcvt_b200.debug_hevc_path(path)
d = torch.from_numpy(frames).cuda()
out = torch.empty((n, imcvt_b200.hevc_bound(h, w)), dtype=torch.uint8, device="cuda")
lens = torch.zeros(n, dtype=torch.int64, device="cuda")
L = imcvt_b200.lib()
buf = (C.c_ulonglong * 128)()
import time
for it in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    imcvt_b200.encode_hevc_device(d, q, out, lens)
    torch.cuda.synchronize(); wall = time.time() - t0
    L.imcvtDebugHevcProf(buf)
print(f"wall {wall*1e3:.1f} ms for the batch")
v = np.array(list(buf), dtype=np.float64)
nctu = n * ((h + 31) // 32) * ((w + 31) // 32)
names = ["entry", "split+children sse", "A one-TU", "A four-TU", "ctx copy", "B trials / CU8: wait for four-TU trials after NxN", "decide", "commit",
         "4x4 A", "4x4 B: thread 0's own trial", "4x4 B: wait for the other trials", "4x4 decide+keep", "CU8: NxN trial (thread 0)", "final copy"]
tot = v.sum()
print(f"{kind} n={n} qpd6={q} {h}x{w}: {tot/nctu:.0f} cycles per CTU (thread 0 timeline), bytes={int(lens.sum())}")
for lv, nm in enumerate(["CU32", "CU16", "CU8"]):
    for p, pn in enumerate(names):
        c = v[lv * 16 + p]
        if c > 0:
            print(f"  {nm} {pn:20s} {c/nctu:10.0f} cyc/CTU  {100*c/tot:5.1f}%")
print(f"  pair spine: enter+push {v[50]/nctu:.0f}  wait#1 {v[51]/nctu:.0f}  split-cost {v[52]/nctu:.0f}  wait#2 {v[53]/nctu:.0f}  decide..finish {v[54]/nctu:.0f}  wait#3 {v[55]/nctu:.0f}")
print(f"  CTU load {v[48]/nctu:10.0f}  {100*v[48]/tot:5.1f}%   CTU store+terminate {v[49]/nctu:10.0f}  {100*v[49]/tot:5.1f}%")
stages = ["refs+smooth", "predict", "fwd transform", "rdoq", "last/cg scan", "dequant+inverse", "recon+sse", "binarize"]
for sc, nm in enumerate(["TU4", "TU8", "TU16", "TU32"]):
    row = v[64 + sc * 8: 64 + sc * 8 + 8]
    if row.sum() > 0:
        print(f"  {nm} (thread 0, cyc/CTU): " + "  ".join(f"{a} {b/nctu:.0f}" for a, b in zip(stages, row)) + f"   total {row.sum()/nctu:.0f}")
