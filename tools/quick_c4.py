"""configs[3]'s unit through the host API (development aid): n 3840x2160 RGB frames at NEAR, host buffers, with the per-kernel
event times of the library.   python tools/quick_c4.py [n] [near] [iters]"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import imcvt_b200
from imcvt_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
near = int(sys.argv[2]) if len(sys.argv) > 2 else 2
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3          # host-buffer calls (0: device-resident path only)
h, w = 2160, 3840
base = np.stack([synth.frame("natural", i, h, w, True) for i in range(min(n, 4))])
frames = np.concatenate([base] * ((n + len(base) - 1) // len(base)))[:n]
imcvt_b200.init(0)
L = imcvt_b200.lib()
stride = 3 * h * w
hf = torch.from_numpy(frames).pin_memory()
ho = torch.empty(n * stride, dtype=torch.uint8).pin_memory()
lens = (C.c_size_t * n)()
imcvt_b200.profile_enable(True)
for it in range(iters):
    imcvt_b200.profile_reset()
    t = time.time()
    rc = L.imcvtEncodeBatchJLS(hf.numpy().ctypes.data, 3 * h * w, n, 1, h, w, near, ho.numpy().ctypes.data, stride, C.cast(lens, C.c_void_p))
    dt = time.time() - t
    assert rc == 0, imcvt_b200.last_error()
    print(f"host call {it}: {dt*1e3:.1f} ms  {n*h*w/dt/1e6:.0f} Mpix/s  bytes={sum(lens)}", flush=True)
    for k, (c, ms) in sorted(imcvt_b200.profile_stats().items(), key=lambda kv: -kv[1][1]):
        print(f"   {k:24s} {c:3d} launches {ms:9.3f} ms")
d = torch.from_numpy(frames).cuda()
out = torch.empty((n, stride), dtype=torch.uint8, device="cuda")
dl = torch.zeros(n, dtype=torch.int64, device="cuda")
for it in range(2):
    torch.cuda.synchronize(); t = time.time()
    imcvt_b200.encode_jls_device(d, near, out, dl)
    torch.cuda.synchronize(); dt = time.time() - t
    print(f"device call {it}: {dt*1e3:.1f} ms  {n*h*w/dt/1e6:.0f} Mpix/s")
