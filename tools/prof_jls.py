"""Role profile of jls_tile_kernel (development aid; needs `make -C imcvt_b200/csrc PROF=1`).
   python tools/prof_jls.py [kind] [n]"""
import ctypes as C, os, sys
os.environ.setdefault("IMCVT_B200_LIB", "libimcvt_b200_prof.so")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import imcvt_b200
from imcvt_b200 import synth

kind = sys.argv[1] if len(sys.argv) > 1 else "natural"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
path = int(sys.argv[3]) if len(sys.argv) > 3 else 0
h, w = 1080, 1920
base = synth.batch(kind, min(n, 8), h, w)
frames = np.concatenate([base] * ((n + len(base) - 1) // len(base)))[:n]
imcvt_b200.init(0)
imcvt_b200.debug_jls_path(path)
d = torch.from_numpy(frames).cuda()
out = torch.empty((n, 2 * h * w + 4096), dtype=torch.uint8, device="cuda")
lens = torch.zeros(n, dtype=torch.int64, device="cuda")
L = imcvt_b200.lib()
buf = (C.c_ulonglong * 16)()
for it in range(2):
    imcvt_b200.encode_jls_device(d, 0, out, lens)
    torch.cuda.synchronize()
    L.imcvtDebugJlsProf(buf)
v = [float(x) for x in buf]
rows = max(1, 6144 // w)
ticks = n * ((h + rows - 1) // rows + (3 if path == 3 else 2))
names = {0: "chain loop (warp 0)", 6: "chain: run-event thread", 7: "chain loop + barrier", 4: "chain + code pass", 1: "producer warp 0", 5: "producer last warp", 2: "packer", 3: "tick (incl. barrier)"}
print(f"{kind} n={n} path={path}: cycles per tick" + ("  (pair: 0 = chain stage 1, 6/7 = chain stage 2 first thread / run thread, 1 = producers, 2 = packer)" if path == 3 else ""))
names.update({8: "  prod: classify done", 9: "  prod: barrier 1", 10: "  prod: scan+order done", 11: "  prod: scatter done"})
names.update({12: "  pack: zero+lens+scan (sum)", 13: "  pack: placed (sum)", 14: "  pack: stuffed (sum)"})
for k in (3, 0, 6, 7, 4, 1, 8, 9, 10, 11, 5, 2, 12, 13, 14):
    print(f"  {names[k]:24s} {v[k]/ticks:9.0f}")
