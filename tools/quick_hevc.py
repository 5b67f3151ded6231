"""Quick device-side timing of the HEVC frame kernel (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import imcvt_b200
from imcvt_b200 import synth

kind = sys.argv[1] if len(sys.argv) > 1 else "natural"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
q = int(sys.argv[3]) if len(sys.argv) > 3 else 4
h = int(sys.argv[4]) if len(sys.argv) > 4 else 1080
w = int(sys.argv[5]) if len(sys.argv) > 5 else 1920
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 2
path = int(sys.argv[7]) if len(sys.argv) > 7 else 0
base = synth.batch(kind, min(n, 8), h, w)
frames = np.concatenate([base] * ((n + len(base) - 1) // len(base)))[:n]
imcvt_b200.init(0)
imcvt_b200.debug_hevc_path(path)
d = torch.from_numpy(frames).cuda()
stride = imcvt_b200.hevc_bound(h, w)
out = torch.empty((n, stride), dtype=torch.uint8, device="cuda")
lens = torch.zeros(n, dtype=torch.int64, device="cuda")
for it in range(iters):
    torch.cuda.synchronize(); t = time.time()
    imcvt_b200.encode_hevc_device(d, q, out, lens)
    torch.cuda.synchronize(); dt = time.time() - t
    print(f"iter {it}: {dt*1e3:.1f} ms  {n*h*w/dt/1e6:.2f} Mpix/s  bytes={int(lens.sum())}  per-CTU {dt*1e6/(((h+31)//32)*((w+31)//32)*max(1,(n+147)//148)):.1f} us", flush=True)
