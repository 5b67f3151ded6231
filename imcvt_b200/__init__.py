"""imcvt_b200 -- host-side Python mirror of the C-ABI in include/imcvt_b200.h (libimcvt_b200.so).

The library is the product; this module is a thin ctypes binding that keeps the reference's names and
argument meaning (src/imageio.h:20-21, src/HEVCe/HEVCe.h:5-12):

    write_jls_image_file(filename, buf, is_rgb, height, width, near)   -> 0 ok / 1 failed
    write_hevc_image_file(filename, buf, is_rgb, height, width, qpd6)  -> 0 ok / 1 failed
    hevc_image_encoder(img, qpd6)                                      -> (stream bytes, reconstructed picture)

plus the batch entry points the single-image API cannot express:

    encode_jls_batch(frames, near)            frames: uint8 [n,H,W] or [n,H,W,3]  -> list of bytes
    encode_hevc_batch(frames, qpd6)           frames: uint8 [n,H,W]               -> (list of bytes, rcon [n,H',W'])
    encode_jls_device / encode_hevc_device    torch CUDA tensors in and out, no host copies

There is no CPU fallback: importing works anywhere (so the symbol table can be checked), but every encode
call raises ImcvtError unless libimcvt_b200.so is built and a Blackwell GPU is present.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# IMCVT_B200_LIB selects a development build of the same library (e.g. libimcvt_b200_prof.so, `make PROF=1`)
LIB_PATH = os.path.join(_HERE, os.environ.get("IMCVT_B200_LIB", "libimcvt_b200.so"))

# every symbol include/imcvt_b200.h and include/imcvt_imageio.h declare
EXPORTED_SYMBOLS = [
    "imcvtInit", "imcvtShutdown", "imcvtLastError", "imcvtVersion",
    "imcvtJLSBound", "imcvtHEVCBound",
    "imcvtEncodeBatchJLS", "imcvtEncodeBatchJLSDevice",
    "imcvtEncodeBatchHEVC", "imcvtEncodeBatchHEVCDevice", "imcvtEncodeBatchHEVCGreen", "imcvtEncodeBatchHEVCSweep",
    "imcvtDeviceCount", "imcvtEncodeBatchJLSMulti", "imcvtEncodeBatchHEVCMulti", "imcvtEncodeBatchHEVCGreenMulti",
    "imcvtEncodeBatchHEVCSweepMulti",
    "imcvtHostAlloc", "imcvtHostFree", "imcvtJLSHostMode",
    "imcvtProfileEnable", "imcvtProfileReset", "imcvtProfileGet", "imcvtProfileList", "imcvtLaunchCount",
    "imcvtDebugJLSForceSerial", "imcvtDebugJLSPath", "imcvtDebugHEVCPath",
    "writeJLSImageFile", "writeHEVCImageFile", "HEVCImageEncoder",
]


class ImcvtError(RuntimeError):
    pass


_lib = None


def lib():
    """Loads libimcvt_b200.so (in-tree build) and declares the prototypes.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImcvtError(f"{LIB_PATH} is missing: build it with `make -C imcvt_b200/csrc` "
                         "(or python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, sz, u32, i32 = C.c_void_p, C.c_size_t, C.c_uint32, C.c_int
    L.imcvtInit.argtypes = [i32]; L.imcvtInit.restype = i32
    L.imcvtShutdown.argtypes = []; L.imcvtShutdown.restype = None
    L.imcvtLastError.argtypes = []; L.imcvtLastError.restype = C.c_char_p
    L.imcvtVersion.argtypes = []; L.imcvtVersion.restype = i32
    L.imcvtJLSBound.argtypes = [i32, u32, u32]; L.imcvtJLSBound.restype = sz
    L.imcvtHEVCBound.argtypes = [u32, u32]; L.imcvtHEVCBound.restype = sz
    L.imcvtEncodeBatchJLS.argtypes = [vp, sz, i32, i32, u32, u32, i32, vp, sz, vp]; L.imcvtEncodeBatchJLS.restype = i32
    L.imcvtEncodeBatchJLSDevice.argtypes = [vp, sz, i32, i32, u32, u32, i32, vp, sz, vp, vp]; L.imcvtEncodeBatchJLSDevice.restype = i32
    L.imcvtEncodeBatchHEVC.argtypes = [vp, sz, i32, u32, u32, i32, vp, sz, vp, vp, sz]; L.imcvtEncodeBatchHEVC.restype = i32
    L.imcvtEncodeBatchHEVCDevice.argtypes = [vp, sz, i32, u32, u32, i32, vp, sz, vp, vp, sz, vp]; L.imcvtEncodeBatchHEVCDevice.restype = i32
    L.imcvtEncodeBatchHEVCGreen.argtypes = [vp, sz, i32, u32, u32, i32, vp, sz, vp]; L.imcvtEncodeBatchHEVCGreen.restype = i32
    L.imcvtEncodeBatchHEVCSweep.argtypes = [vp, sz, i32, u32, u32, C.POINTER(i32), i32, vp, sz, vp]; L.imcvtEncodeBatchHEVCSweep.restype = i32
    ip = C.POINTER(i32)
    L.imcvtDeviceCount.argtypes = []; L.imcvtDeviceCount.restype = i32
    L.imcvtEncodeBatchJLSMulti.argtypes = [vp, sz, i32, i32, u32, u32, i32, vp, sz, vp, ip, i32]; L.imcvtEncodeBatchJLSMulti.restype = i32
    L.imcvtEncodeBatchHEVCMulti.argtypes = [vp, sz, i32, u32, u32, i32, vp, sz, vp, vp, sz, ip, i32]; L.imcvtEncodeBatchHEVCMulti.restype = i32
    L.imcvtEncodeBatchHEVCGreenMulti.argtypes = [vp, sz, i32, u32, u32, i32, vp, sz, vp, ip, i32]; L.imcvtEncodeBatchHEVCGreenMulti.restype = i32
    L.imcvtEncodeBatchHEVCSweepMulti.argtypes = [vp, sz, i32, u32, u32, ip, i32, vp, sz, vp, ip, i32]; L.imcvtEncodeBatchHEVCSweepMulti.restype = i32
    L.imcvtHostAlloc.argtypes = [sz]; L.imcvtHostAlloc.restype = vp
    L.imcvtHostFree.argtypes = [vp]; L.imcvtHostFree.restype = None
    L.imcvtJLSHostMode.argtypes = [i32]; L.imcvtJLSHostMode.restype = None
    L.imcvtProfileEnable.argtypes = [i32]; L.imcvtProfileEnable.restype = None
    L.imcvtProfileReset.argtypes = []; L.imcvtProfileReset.restype = None
    L.imcvtProfileGet.argtypes = [C.c_char_p, C.POINTER(i32), C.POINTER(C.c_double)]; L.imcvtProfileGet.restype = i32
    L.imcvtProfileList.argtypes = [C.c_char_p, sz]; L.imcvtProfileList.restype = i32
    L.imcvtLaunchCount.argtypes = []; L.imcvtLaunchCount.restype = C.c_long
    L.imcvtDebugJLSForceSerial.argtypes = [i32]; L.imcvtDebugJLSForceSerial.restype = None
    L.imcvtDebugJLSPath.argtypes = [i32]; L.imcvtDebugJLSPath.restype = None
    L.imcvtDebugHEVCPath.argtypes = [i32]; L.imcvtDebugHEVCPath.restype = None
    L.writeJLSImageFile.argtypes = [C.c_char_p, vp, i32, u32, u32, i32]; L.writeJLSImageFile.restype = i32
    L.writeHEVCImageFile.argtypes = [C.c_char_p, vp, i32, u32, u32, i32]; L.writeHEVCImageFile.restype = i32
    L.HEVCImageEncoder.argtypes = [vp, vp, vp, C.POINTER(i32), C.POINTER(i32), i32]; L.HEVCImageEncoder.restype = i32
    _lib = L
    return L


def last_error():
    return lib().imcvtLastError().decode("utf-8", "replace")


def _check(rc, what):
    if rc != 0:
        raise ImcvtError(f"{what} failed: {last_error()}")


def init(device=0):
    _check(lib().imcvtInit(int(device)), "imcvtInit")


def shutdown():
    lib().imcvtShutdown()


def jls_bound(is_rgb, height, width):
    return int(lib().imcvtJLSBound(int(is_rgb), height, width))


def hevc_bound(height, width):
    return int(lib().imcvtHEVCBound(height, width))


def pad32(v):
    return (min(int(v), 8192) + 31) // 32 * 32


# ---------------------------------------------------------------------------------------------
# host-buffer batch API
# ---------------------------------------------------------------------------------------------
def device_count():
    return int(lib().imcvtDeviceCount())


def _devlist(devices):
    """None -> (NULL, 0) = all visible devices; an int k -> devices 0..k-1; a list -> those ordinals."""
    if devices is None:
        return None, 0
    if isinstance(devices, int):
        devices = list(range(devices))
    arr = (C.c_int * len(devices))(*[int(d) for d in devices])
    return arr, len(devices)


def encode_jls_batch(frames, near=0, out=None, out_stride=None, devices=0):
    """frames: uint8 array [n,H,W] (gray) or [n,H,W,3] (RGB), C-contiguous (may be pinned).
    Returns a list of n `bytes` (.jls streams, bit-identical to the reference's writeJLSImageFile).
    devices: 0 = the engine's device (imcvtEncodeBatchJLS); None / k / [ordinals] = sharded over all / the first k / those
    devices (imcvtEncodeBatchJLSMulti) -- the same bytes."""
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    if frames.ndim not in (3, 4) or (frames.ndim == 4 and frames.shape[3] != 3):
        raise ValueError("frames must be [n,H,W] or [n,H,W,3] uint8")
    n, h, w = frames.shape[:3]
    is_rgb = int(frames.ndim == 4)
    if out_stride is None:
        out_stride = jls_bound(is_rgb, h, w)
    if out is None:
        out = np.empty(n * out_stride, dtype=np.uint8)
    lens = (C.c_size_t * n)()
    frame_stride = h * w * (3 if is_rgb else 1)
    if devices == 0:
        _check(lib().imcvtEncodeBatchJLS(frames.ctypes.data, frame_stride, n, is_rgb, h, w, int(near),
                                         out.ctypes.data, out_stride, C.cast(lens, C.c_void_p)), "imcvtEncodeBatchJLS")
    else:
        dl, nd = _devlist(devices)
        _check(lib().imcvtEncodeBatchJLSMulti(frames.ctypes.data, frame_stride, n, is_rgb, h, w, int(near),
                                              out.ctypes.data, out_stride, C.cast(lens, C.c_void_p), dl, nd), "imcvtEncodeBatchJLSMulti")
    return [out[i * out_stride:i * out_stride + lens[i]].tobytes() for i in range(n)]


def encode_hevc_batch(frames, qpd6=4, want_rcon=True, devices=0):
    """frames: uint8 [n,H,W].  Returns (list of n .h265 streams, rcon uint8 [n,H',W'] or None).  devices: see encode_jls_batch."""
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    if frames.ndim != 3:
        raise ValueError("frames must be [n,H,W] uint8")
    n, h, w = frames.shape
    out_stride = hevc_bound(h, w)
    out = np.empty(n * out_stride, dtype=np.uint8)
    lens = (C.c_size_t * n)()
    hp, wp = pad32(h), pad32(w)
    rcon = np.empty((n, hp, wp), dtype=np.uint8) if want_rcon else None
    if devices == 0:
        _check(lib().imcvtEncodeBatchHEVC(frames.ctypes.data, h * w, n, h, w, int(qpd6),
                                          out.ctypes.data, out_stride, C.cast(lens, C.c_void_p),
                                          rcon.ctypes.data if want_rcon else None, hp * wp), "imcvtEncodeBatchHEVC")
    else:
        dl, nd = _devlist(devices)
        _check(lib().imcvtEncodeBatchHEVCMulti(frames.ctypes.data, h * w, n, h, w, int(qpd6),
                                               out.ctypes.data, out_stride, C.cast(lens, C.c_void_p),
                                               rcon.ctypes.data if want_rcon else None, hp * wp, dl, nd), "imcvtEncodeBatchHEVCMulti")
    return [out[i * out_stride:i * out_stride + lens[i]].tobytes() for i in range(n)], rcon


def encode_hevc_sweep(frames, qpd6_list, devices=0):
    """frames: uint8 [n,H,W]; qpd6_list: 1..8 settings.  One launch codes every frame at every setting.
    Returns a list of n lists: streams[i][k] = frame i at qpd6_list[k] (same bytes as encode_hevc_batch per setting)."""
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    if frames.ndim != 3:
        raise ValueError("frames must be [n,H,W] uint8")
    n, h, w = frames.shape
    nq = len(qpd6_list)
    q = (C.c_int * nq)(*[int(x) for x in qpd6_list])
    out_stride = hevc_bound(h, w)
    out = np.empty(n * nq * out_stride, dtype=np.uint8)
    lens = (C.c_size_t * (n * nq))()
    if devices == 0:
        _check(lib().imcvtEncodeBatchHEVCSweep(frames.ctypes.data, h * w, n, h, w, q, nq, out.ctypes.data, out_stride,
                                               C.cast(lens, C.c_void_p)), "imcvtEncodeBatchHEVCSweep")
    else:
        dl, nd = _devlist(devices)
        _check(lib().imcvtEncodeBatchHEVCSweepMulti(frames.ctypes.data, h * w, n, h, w, q, nq, out.ctypes.data, out_stride,
                                                    C.cast(lens, C.c_void_p), dl, nd), "imcvtEncodeBatchHEVCSweepMulti")
    return [[out[(i * nq + k) * out_stride:(i * nq + k) * out_stride + lens[i * nq + k]].tobytes() for k in range(nq)] for i in range(n)]


def encode_hevc_green_batch(frames, qpd6=4, devices=0):
    """frames: uint8 [n,H,W,3] interleaved RGB.  Encodes the green channel (what the reference's writeHEVCImageFile does
    with RGB input, src/imageio_hevc.c:21-26); the channel is extracted on the device.  Returns the n .h265 streams."""
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    if frames.ndim != 4 or frames.shape[3] != 3:
        raise ValueError("frames must be [n,H,W,3] uint8")
    n, h, w = frames.shape[:3]
    out_stride = hevc_bound(h, w)
    out = np.empty(n * out_stride, dtype=np.uint8)
    lens = (C.c_size_t * n)()
    if devices == 0:
        _check(lib().imcvtEncodeBatchHEVCGreen(frames.ctypes.data, h * w * 3, n, h, w, int(qpd6),
                                               out.ctypes.data, out_stride, C.cast(lens, C.c_void_p)), "imcvtEncodeBatchHEVCGreen")
    else:
        dl, nd = _devlist(devices)
        _check(lib().imcvtEncodeBatchHEVCGreenMulti(frames.ctypes.data, h * w * 3, n, h, w, int(qpd6),
                                                    out.ctypes.data, out_stride, C.cast(lens, C.c_void_p), dl, nd), "imcvtEncodeBatchHEVCGreenMulti")
    return [out[i * out_stride:i * out_stride + lens[i]].tobytes() for i in range(n)]


# ---------------------------------------------------------------------------------------------
# device-resident API (torch CUDA tensors; used by bench.py for the HBM-resident number)
# ---------------------------------------------------------------------------------------------
def encode_jls_device(d_frames, near, d_out, d_len, stream_ptr=0):
    """d_frames: torch.uint8 CUDA [n,H,W] or [n,H,W,3]; d_out: torch.uint8 CUDA [n, out_stride];
    d_len: torch.int64 CUDA [n].  Queues the encode on the given CUDA stream (raw pointer, 0 = default)."""
    n, h, w = d_frames.shape[:3]
    is_rgb = int(d_frames.dim() == 4)
    assert d_frames.is_contiguous() and d_out.is_contiguous() and d_len.is_contiguous()
    _check(lib().imcvtEncodeBatchJLSDevice(d_frames.data_ptr(), d_frames.stride(0), n, is_rgb, h, w, int(near),
                                           d_out.data_ptr(), d_out.stride(0), d_len.data_ptr(), stream_ptr),
           "imcvtEncodeBatchJLSDevice")


def encode_hevc_device(d_frames, qpd6, d_out, d_len, d_rcon=None, stream_ptr=0):
    n, h, w = d_frames.shape
    assert d_frames.is_contiguous() and d_out.is_contiguous() and d_len.is_contiguous()
    _check(lib().imcvtEncodeBatchHEVCDevice(d_frames.data_ptr(), d_frames.stride(0), n, h, w, int(qpd6),
                                            d_out.data_ptr(), d_out.stride(0), d_len.data_ptr(),
                                            d_rcon.data_ptr() if d_rcon is not None else None,
                                            d_rcon.stride(0) if d_rcon is not None else 0, stream_ptr),
           "imcvtEncodeBatchHEVCDevice")


# ---------------------------------------------------------------------------------------------
# the reference's own interface, same names / argument meaning / return codes
# ---------------------------------------------------------------------------------------------
def write_jls_image_file(filename, buf, is_rgb, height, width, near):
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    return int(lib().writeJLSImageFile(os.fsencode(filename), buf.ctypes.data, int(is_rgb), height, width, int(near)))


def write_hevc_image_file(filename, buf, is_rgb, height, width, qpd6):
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    return int(lib().writeHEVCImageFile(os.fsencode(filename), buf.ctypes.data, int(is_rgb), height, width, int(qpd6)))


def hevc_image_encoder(img, qpd6):
    """Mirror of HEVCImageEncoder(pbuffer, img, img_rcon, &ysz, &xsz, qpd6): returns (stream, rcon[H',W'])."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    hp, wp = pad32(h), pad32(w)
    out = np.empty(hevc_bound(h, w), dtype=np.uint8)
    rcon = np.empty((hp, wp), dtype=np.uint8)
    ch, cw = C.c_int(h), C.c_int(w)
    n = lib().HEVCImageEncoder(out.ctypes.data, img.ctypes.data, rcon.ctypes.data, C.byref(ch), C.byref(cw), int(qpd6))
    if n <= 0:
        raise ImcvtError(f"HEVCImageEncoder failed: {last_error()}")
    assert (ch.value, cw.value) == (hp, wp)
    return out[:n].tobytes(), rcon


# ---------------------------------------------------------------------------------------------
# instrumentation
# ---------------------------------------------------------------------------------------------
def profile_enable(on=True):
    lib().imcvtProfileEnable(int(bool(on)))


def profile_reset():
    lib().imcvtProfileReset()


def profile_stats():
    """{kernel name: (launches, total device ms)} accumulated since profile_reset()."""
    buf = C.create_string_buffer(1 << 16)
    if lib().imcvtProfileList(buf, len(buf)) != 0:
        return {}
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.rsplit(":", 2)
        out[name] = (int(cnt), float(ms))
    return out


def launch_count():
    return int(lib().imcvtLaunchCount())


def debug_jls_force_serial(on):
    lib().imcvtDebugJLSForceSerial(int(bool(on)))


def debug_jls_path(path):
    """0 = default NEAR=0 path (tile kernel), 1 = force the multi-kernel pipeline (used for planes wider than 6144)."""
    lib().imcvtDebugJLSPath(int(path))


def debug_hevc_path(path):
    """0 = default, 1 = one CTA per frame, 2 = a cluster of two CTAs per frame (see imcvtDebugHEVCPath)."""
    lib().imcvtDebugHEVCPath(int(path))
