"""Host-buffer (page-locked) JPEG-LS batch through the C-ABI in the three transfer modes (development aid)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import imcvt_b200
from imcvt_b200 import synth

kind = sys.argv[1] if len(sys.argv) > 1 else "natural"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
h, w = 1080, 1920
base = synth.batch(kind, min(n, 8), h, w)
frames = np.concatenate([base] * ((n + len(base) - 1) // len(base)))[:n]
imcvt_b200.init(0)
L = imcvt_b200.lib()
stride = imcvt_b200.jls_bound(0, h, w)
h_in = torch.from_numpy(frames).pin_memory()
h_out = torch.empty(n * stride, dtype=torch.uint8).pin_memory()
lens = (C.c_size_t * n)()
ref = None
for mode in (0, 1, 2, 0, 1, 2):
    L.imcvtJLSHostMode(mode)
    ts = []
    for it in range(6):
        torch.cuda.synchronize(); t = time.perf_counter()
        rc = L.imcvtEncodeBatchJLS(h_in.numpy().ctypes.data, h * w, n, 0, h, w, 0, h_out.numpy().ctypes.data, stride, C.cast(lens, C.c_void_p))
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
        assert rc == 0, imcvt_b200.last_error()
    got = [h_out.numpy()[i * stride:i * stride + lens[i]].tobytes() for i in range(n)]
    if ref is None: ref = got
    assert got == ref, f"mode {mode} differs"
    best = min(ts[1:])
    print(f"mode {mode}: best {best*1e3:.2f} ms  median {sorted(ts[1:])[2]*1e3:.2f} ms  {n*h*w/best/1e6:.0f} Mpix/s  bytes={sum(lens)}")
