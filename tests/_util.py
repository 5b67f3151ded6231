"""Shared test helpers: fixture loading, the two CPU checkers (oracle restatement + compiled reference).

Everything under oracle/ is TEST INFRASTRUCTURE: it is loaded here (tests), in __graft_entry__.smoke()
and in bench.py's cpu_baseline / --impl reference legs only.
"""
import ctypes as C
import hashlib
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
ORACLE_DIR = os.path.join(ROOT, "oracle")


def load_pnm(path):
    """Minimal PNM reader with the reference loader's conventions (imageio_pnm.c:73-148):
    maxval is NOT rescaled, PBM 1 -> 0 and 0 -> 255.  Returns (array HxW or HxWx3 u8, is_rgb)."""
    data = open(path, "rb").read()
    assert data[0:1] == b"P"
    t = data[1] - ord("0")
    pos = 2
    nums = []
    want = 2 if t in (1, 4) else 3

    def next_number(pos):
        while True:
            ch = data[pos:pos + 1]
            if ch == b"#":
                while data[pos:pos + 1] not in (b"\r", b"\n", b""):
                    pos += 1
            elif ch.isdigit():
                v = 0
                while data[pos:pos + 1].isdigit():
                    v = v * 10 + int(data[pos:pos + 1])
                    pos += 1
                return v, pos + 1          # the reference consumes one char after the number
            elif ch == b"":
                return -1, pos
            pos += 1

    last_ch_pos = pos
    while len(nums) < want:
        v, pos = next_number(pos)
        nums.append(v)
    w, h = nums[0], nums[1]
    # skip to end of the header line (imageio_pnm.c:100-102)
    if data[pos - 1:pos] != b"\n":
        while data[pos:pos + 1] not in (b"\n", b""):
            pos += 1
        pos += 1
    is_rgb = t in (3, 6)
    n = w * h * (3 if is_rgb else 1)
    if t in (5, 6):
        arr = np.frombuffer(data, dtype=np.uint8, count=n, offset=pos).copy()
    elif t == 4:
        rowbytes = (w + 7) // 8
        bits = np.unpackbits(np.frombuffer(data, dtype=np.uint8, count=rowbytes * h, offset=pos).reshape(h, rowbytes), axis=1)[:, :w]
        arr = np.where(bits != 0, 0, 255).astype(np.uint8).reshape(-1)
    else:
        vals = []
        while len(vals) < n:
            v, pos = next_number(pos)
            vals.append(v)
        arr = np.array(vals, dtype=np.int64)
        if t == 1:
            arr = np.where(arr != 0, 0, 255)
        arr = arr.astype(np.uint8)
    return (arr.reshape(h, w, 3) if is_rgb else arr.reshape(h, w)), is_rgb


def sha256(b):
    return hashlib.sha256(bytes(b)).hexdigest()


def golden_table():
    with open(os.path.join(GOLDEN, "known_answers.json")) as f:
        return json.load(f)


def _build_oracle():
    subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True, capture_output=True)


_libs = {}


def oracle_lib():
    """ctypes handle of oracle/liboracle.so (this repo's plain-C restatement)."""
    if "oracle" not in _libs:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            _build_oracle()
        lib = C.CDLL(path)
        lib.oracle_jls_encode.restype = C.c_size_t
        lib.oracle_jls_encode.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_size_t]
        if hasattr(lib, "oracle_hevc_encode"):
            lib.oracle_hevc_encode.restype = C.c_long
            lib.oracle_hevc_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]
        _libs["oracle"] = lib
    return _libs["oracle"]


def ref_lib():
    """ctypes handle of oracle/_ref/libimcvt_ref.so (the UNMODIFIED reference, compiled by oracle/Makefile),
    or None when it has not been built (it can only be built where /root/reference exists)."""
    if "ref" not in _libs:
        path = os.path.join(ORACLE_DIR, "_ref", "libimcvt_ref.so")
        if not os.path.exists(path) and os.path.isdir("/root/reference/src"):
            _build_oracle()
        if not os.path.exists(path):
            _libs["ref"] = None
        else:
            lib = C.CDLL(path)
            lib.ref_jls_encode.restype = C.c_size_t
            lib.ref_jls_encode.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_size_t]
            lib.HEVCImageEncoder.restype = C.c_int
            lib.HEVCImageEncoder.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]
            _libs["ref"] = lib
    return _libs["ref"]


def _jls_call(fn, img, near):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    is_rgb = int(img.ndim == 3)
    h, w = img.shape[:2]
    cap = 8 * h * w * (3 if is_rgb else 1) + 65536
    out = np.empty(cap, dtype=np.uint8)
    n = fn(img.ctypes.data, is_rgb, h, w, near, out.ctypes.data, cap)
    assert n > 0
    return out[:n].tobytes()


def oracle_jls(img, near):
    return _jls_call(oracle_lib().oracle_jls_encode, img, near)


def ref_jls(img, near):
    return _jls_call(ref_lib().ref_jls_encode, img, near)


def _hevc_call(fn, img, qpd6):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    assert img.ndim == 2
    h, w = img.shape
    hp, wp = (min(h, 8192) + 31) // 32 * 32, (min(w, 8192) + 31) // 32 * 32
    out = np.zeros(2 * (wp + 32) * (hp + 32) + 65536, dtype=np.uint8)
    rcon = np.zeros(hp * wp + 64, dtype=np.uint8)
    ch, cw = C.c_int(h), C.c_int(w)
    n = fn(out.ctypes.data, img.ctypes.data, rcon.ctypes.data, C.byref(ch), C.byref(cw), qpd6)
    assert n > 0 and ch.value == hp and cw.value == wp
    return out[:n].tobytes(), rcon[:hp * wp].reshape(hp, wp).copy()


def oracle_hevc(img, qpd6):
    return _hevc_call(oracle_lib().oracle_hevc_encode, img, qpd6)


def ref_hevc(img, qpd6):
    return _hevc_call(ref_lib().HEVCImageEncoder, img, qpd6)


def check_jls(img, near):
    """The checker of the GPU tier: the UNMODIFIED reference (oracle/_ref, which travels to the GPU box prebuilt) whenever
    it is there, this repo's restatement only as the fallback."""
    return ref_jls(img, near) if ref_lib() is not None else oracle_jls(img, near)


def check_hevc(img, qpd6):
    return ref_hevc(img, qpd6) if ref_lib() is not None else oracle_hevc(img, qpd6)


def fullsize_table():
    """(length, sha256) of the reference's streams for synthetic frames at BASELINE sizes (tests/golden/make_golden_fullsize.py)."""
    with open(os.path.join(GOLDEN, "fullsize_answers.json")) as f:
        return json.load(f)


def emu_hevc(img, qpd6, pair=False, lockstep=False):
    """The SOURCE of the CUDA HEVC kernel run under the CPU SIMT emulator (tests/emu, test infrastructure).
    lockstep: the 8x8 units searched with all warps in lock step (Params::cu8_lockstep) instead of by three warp groups."""
    if "emu" not in _libs:
        d = os.path.join(ROOT, "tests", "emu")
        subprocess.run(["make", "-C", d, "-s", "libhevc_emu.so"], check=True, capture_output=True)
        lib = C.CDLL(os.path.join(d, "libhevc_emu.so"))
        lib.emu_hevc_encode.restype = C.c_long
        lib.emu_hevc_encode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]
        lib.emu_hevc_encode_mode.restype = C.c_long
        lib.emu_hevc_encode_mode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int]
        _libs["emu"] = lib
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    hp, wp = (min(h, 8192) + 31) // 32 * 32, (min(w, 8192) + 31) // 32 * 32
    out = np.zeros(2 * (wp + 32) * (hp + 32) + 65536, dtype=np.uint8)
    rcon = np.zeros(hp * wp, dtype=np.uint8)
    ch, cw = C.c_int(h), C.c_int(w)
    n = _libs["emu"].emu_hevc_encode_mode(out.ctypes.data, out.size, img.ctypes.data, rcon.ctypes.data, C.byref(ch), C.byref(cw), qpd6, (1 if pair else 0) | (2 if lockstep else 0))
    assert n > 0 and ch.value == hp and cw.value == wp
    return out[:n].tobytes(), rcon.reshape(hp, wp).copy()


def emu_hevc_sweep(img, qlist, pair=False):
    """A quantiser sweep (one frame at every qpd6 of qlist, the job queue of imcvtEncodeBatchHEVCSweep) on one emulated CTA /
    CTA pair: the per-job switch of quantiser and context tables.  Returns [(stream, rcon), ...] in qlist order."""
    emu_hevc(np.zeros((8, 8), np.uint8), 4)                       # loads the library
    lib = _libs["emu"]
    lib.emu_hevc_sweep.restype = C.c_int
    lib.emu_hevc_sweep.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int,
                                   C.POINTER(C.c_ulonglong), C.c_int]
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    hp, wp = (min(h, 8192) + 31) // 32 * 32, (min(w, 8192) + 31) // 32 * 32
    nq = len(qlist)
    cap = 2 * (wp + 32) * (hp + 32) + 65536
    out = np.zeros(nq * cap, dtype=np.uint8)
    rcon = np.zeros((nq, hp, wp), dtype=np.uint8)
    lens = (C.c_ulonglong * nq)()
    q = (C.c_int * nq)(*[int(x) for x in qlist])
    assert lib.emu_hevc_sweep(out.ctypes.data, cap, img.ctypes.data, rcon.ctypes.data, h, w, q, nq, lens, 1 if pair else 0) == 0
    return [(out[k * cap:k * cap + lens[k]].tobytes(), rcon[k].copy()) for k in range(nq)]


def emu_jls(img, pair=False, line_out=False):
    """The SOURCE of the CUDA JPEG-LS tile kernel (NEAR = 0, gray) run under the CPU SIMT emulator (tests/emu);
    pair=True: the two-CTA cluster variant (rank 0 producers, rank 1 chains + packers)."""
    if "emu_jls" not in _libs:
        d = os.path.join(ROOT, "tests", "emu")
        subprocess.run(["make", "-C", d, "-s", "libjls_emu.so"], check=True, capture_output=True)
        lib = C.CDLL(os.path.join(d, "libjls_emu.so"))
        lib.emu_jls_encode.restype = C.c_size_t
        lib.emu_jls_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        lib.emu_jls_encode_mode.restype = C.c_size_t
        lib.emu_jls_encode_mode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int]
        _libs["emu_jls"] = lib
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    # line_out: the packer's 128-byte line ring (the output path for page-locked host slots); the slot then starts at an
    # arbitrary offset inside a line so that the partial first line is exercised
    out = np.zeros(8 * h * w + 65536 + 256, dtype=np.uint8)
    off = (int(line_out) * (37 + 128 - out.ctypes.data % 128)) % 128 if line_out else 0
    n = _libs["emu_jls"].emu_jls_encode_mode(img.ctypes.data, h, w, out.ctypes.data + off, out.size - 256,
                                             (1 if pair else 0) | (2 if line_out else 0))
    assert n > 0
    return out[off:off + n].tobytes()


def jls_decode(stream):
    """Independent JPEG-LS decoder (tests/tools/jls_decoder.c, written from ITU-T T.87's decoding procedures; the
    reference ships no decoder).  Returns (planes uint8 [ncomp, H, W], NEAR of the last scan)."""
    if "jlsdec" not in _libs:
        d = os.path.join(ROOT, "tests", "tools")
        so = os.path.join(d, "libjlsdec.so")
        src = os.path.join(d, "jls_decoder.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-o", so, src], check=True, capture_output=True)
        lib = C.CDLL(so)
        lib.jls_decode.restype = C.c_long
        lib.jls_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _libs["jlsdec"] = lib
    buf = np.frombuffer(stream, dtype=np.uint8)
    # dimensions from the SOF55 segment right after SOI (FFD8 FFF7 len P Y X Nf)
    h, w, nc = (int(buf[7]) << 8) | int(buf[8]), (int(buf[9]) << 8) | int(buf[10]), int(buf[11])
    out = np.zeros(nc * h * w + 16, dtype=np.uint8)
    cw, ch, cn, near = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    rc = _libs["jlsdec"].jls_decode(buf.ctypes.data, len(stream), out.ctypes.data, C.byref(cw), C.byref(ch), C.byref(cn), C.byref(near))
    assert rc == 0, f"jls_decode failed with {rc}"
    assert (cw.value, ch.value, cn.value) == (w, h, nc)
    return out[: nc * h * w].reshape(nc, h, w), near.value


def emu_jls_rowscan(img, near):
    """The SOURCE of the row-pipelined JPEG-LS scan kernel (imcvt_b200/csrc/jls_rowscan_kernel.cuh, the path for NEAR > 0)
    run under the CPU SIMT emulator (tests/emu); gray images, any NEAR."""
    if "emu_rowscan" not in _libs:
        d = os.path.join(ROOT, "tests", "emu")
        subprocess.run(["make", "-C", d, "-s", "libjls_rowscan_emu.so"], check=True, capture_output=True)
        lib = C.CDLL(os.path.join(d, "libjls_rowscan_emu.so"))
        lib.emu_jls_rowscan_encode.restype = C.c_size_t
        lib.emu_jls_rowscan_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        _libs["emu_rowscan"] = lib
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    out = np.zeros(8 * h * w + 65536, dtype=np.uint8)
    n = _libs["emu_rowscan"].emu_jls_rowscan_encode(img.ctypes.data, h, w, int(near), out.ctypes.data, out.size)
    assert n > 0
    return out[:n].tobytes()


def emu_jls_nearscan(img, near):
    """The SOURCE of the candidate-table JPEG-LS scan kernel (imcvt_b200/csrc/jls_nearscan_kernel.cuh, the path for NEAR 1..4)
    run under the CPU SIMT emulator (tests/emu); gray images."""
    if "emu_nearscan" not in _libs:
        d = os.path.join(ROOT, "tests", "emu")
        subprocess.run(["make", "-C", d, "-s", "libjls_nearscan_emu.so"], check=True, capture_output=True)
        lib = C.CDLL(os.path.join(d, "libjls_nearscan_emu.so"))
        lib.emu_jls_nearscan_encode.restype = C.c_size_t
        lib.emu_jls_nearscan_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        _libs["emu_nearscan"] = lib
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    out = np.zeros(8 * h * w + 65536, dtype=np.uint8)
    n = _libs["emu_nearscan"].emu_jls_nearscan_encode(img.ctypes.data, h, w, int(near), out.ctypes.data, out.size)
    assert n > 0
    return out[:n].tobytes()
