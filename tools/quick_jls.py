"""Quick device-side timing of the JPEG-LS pipeline (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import imcvt_b200
from imcvt_b200 import synth

kind = sys.argv[1] if len(sys.argv) > 1 else "natural"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
near = int(sys.argv[3]) if len(sys.argv) > 3 else 0
path = int(sys.argv[4]) if len(sys.argv) > 4 else 0
h, w = 1080, 1920
base = synth.batch(kind, min(n, 8), h, w)
frames = np.concatenate([base] * ((n + len(base) - 1) // len(base)))[:n]
imcvt_b200.init(0)
imcvt_b200.debug_jls_path(path)
d = torch.from_numpy(frames).cuda()
stride = 2 * h * w + 4096
out = torch.empty((n, stride), dtype=torch.uint8, device="cuda")
lens = torch.zeros(n, dtype=torch.int64, device="cuda")
imcvt_b200.profile_enable(True)
for it in range(3):
    imcvt_b200.profile_reset()
    torch.cuda.synchronize(); t = time.time()
    imcvt_b200.encode_jls_device(d, near, out, lens)
    torch.cuda.synchronize(); dt = time.time() - t
    st = imcvt_b200.profile_stats()
    print(f"iter {it}: {dt*1e3:.2f} ms  {n*h*w/dt/1e6:.0f} Mpix/s  bytes={int(lens.sum())}")
for k, (c, ms) in sorted(st.items(), key=lambda kv: -kv[1][1]):
    print(f"   {k:28s} {c:3d} launches {ms:9.3f} ms")
