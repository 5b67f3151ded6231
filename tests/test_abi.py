"""CPU: the C-ABI library builds, loads, and exports every symbol the headers declare (no compute calls)."""
import ctypes as C
import os
import re

import pytest

import imcvt_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for hdr in ("imcvt_b200.h", "imcvt_imageio.h"):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", text):
            name = m.group(1)
            if name.startswith("imcvt") or name in ("writeJLSImageFile", "writeHEVCImageFile", "HEVCImageEncoder"):
                names.add(name)
    return names


def test_binding_lists_every_declared_symbol():
    assert declared_symbols() == set(imcvt_b200.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(imcvt_b200.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = C.CDLL(imcvt_b200.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_bounds_and_argument_checks_without_gpu():
    L = imcvt_b200.lib()
    assert L.imcvtVersion() >= 100
    # the reference's stream allocation for HEVC (imageio_hevc.c:14)
    assert imcvt_b200.hevc_bound(263, 300) == 2 * (300 + 32) * (263 + 32) + 65536
    assert imcvt_b200.jls_bound(0, 1, 1) >= 27
    import numpy as np
    out = np.zeros(1024, dtype=np.uint8)
    ln = (C.c_size_t * 1)()
    img = np.zeros(4, dtype=np.uint8)
    # bad dimensions are rejected before any device work, like imageio_jls.c:436
    assert L.imcvtEncodeBatchJLS(img.ctypes.data, 4, 1, 0, 0, 4, 0, out.ctypes.data, 1024, C.cast(ln, C.c_void_p)) == 1
    assert L.imcvtEncodeBatchJLS(img.ctypes.data, 4, 1, 0, 2, 40000, 0, out.ctypes.data, 1024, C.cast(ln, C.c_void_p)) == 1
    assert b"32767" in L.imcvtLastError()
    assert L.writeJLSImageFile(b"/tmp/never_written.jls", img.ctypes.data, 0, 0, 4, 0) == 1
    assert not os.path.exists("/tmp/never_written.jls")


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    with pytest.raises(imcvt_b200.ImcvtError):
        imcvt_b200.encode_jls_batch(np.zeros((1, 4, 4), dtype=np.uint8), 0)
    assert "no CPU fallback" in imcvt_b200.last_error() or "CUDA" in imcvt_b200.last_error()
