"""GPU: the batch command line (imcvt_b200/ImCvt_b200 = host/imcvt_cli.c + host/pnm_io.c on libimcvt_b200.so) -- SURVEY
section 8(f) rows f1 / f2.  One call with many `<in> -o <out>` pairs of several shapes, codecs grouped into batches:
every .jls / .h265 must be the oracle's stream, and stdout / exit code the reference CLI's (src/main.c) byte for byte
(oracle/_ref/ImCvt_ref, the reference compiled unmodified, when it travelled with the snapshot).  Also the two staging
entry points the front-end uses: imcvtEncodeBatchHEVCGreen (green plane picked on the device, replaces
src/imageio_hevc.c:24-26) and imcvtHostAlloc (page-locked staging)."""
import ctypes as C
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

import imcvt_b200
from imcvt_b200 import synth
from tests._util import GOLDEN, ROOT, load_pnm
from tests._util import check_hevc as oracle_hevc, check_jls as oracle_jls   # the unmodified reference when present

pytestmark = pytest.mark.gpu

CLI = os.path.join(ROOT, "imcvt_b200", "ImCvt_b200")
REF = os.path.join(ROOT, "oracle", "_ref", "ImCvt_ref")


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(CLI):
        subprocess.run(["make", "-C", os.path.join(ROOT, "imcvt_b200", "csrc")], check=True, capture_output=True)


def write_pnm(path, img):
    h, w = img.shape[:2]
    with open(path, "wb") as f:
        f.write(b"P%c\n%d %d\n255\n" % (b"6" if img.ndim == 3 else b"5", w, h))
        f.write(np.ascontiguousarray(img).tobytes())


def run(exe, args, cwd, env=None):
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([exe] + args, cwd=cwd, capture_output=True, timeout=900, env=e)
    return p.returncode, p.stdout, p.stderr


def make_inputs(d):
    """name -> image; several shapes so that the front-end has to open new batches, gray and RGB"""
    imgs = {}
    for i, f in enumerate(synth.batch("natural", 5, 96, 160)):
        imgs[f"n{i}.pgm"] = f
    for i, f in enumerate(synth.batch("texture", 3, 70, 131)):
        imgs[f"t{i}.pgm"] = f
    rgb = np.stack([synth.batch(k, 2, 64, 96) for k in ("natural", "texture", "noise")], axis=-1)
    for i in range(2):
        imgs[f"c{i}.ppm"] = rgb[i]
    for name, img in imgs.items():
        write_pnm(os.path.join(d, name), img)
    for f in ("P5", "P6", "P2"):
        shutil.copy(os.path.join(GOLDEN, f + ".pnm"), os.path.join(d, f + ".pnm"))
        imgs[f + ".pnm"] = load_pnm(os.path.join(d, f + ".pnm"))[0]
    return imgs


def pairs(names, suffix):
    args = []
    for n in names:
        args += [n, "-o", os.path.splitext(n)[0] + suffix]
    return args


@pytest.mark.parametrize("setting,batch", [(0, "256"), (2, "3"), (4, "1")])
def test_many_files_one_call_jls(setting, batch):
    with tempfile.TemporaryDirectory() as d, tempfile.TemporaryDirectory() as d2:
        imgs = make_inputs(d)
        make_inputs(d2)
        order = ["n0.pgm", "n1.pgm", "n2.pgm", "t0.pgm", "n3.pgm", "c0.ppm", "c1.ppm", "P5.pnm", "missing.pgm", "t1.pgm", "t2.pgm",
                 "n4.pgm", "P6.pnm", "P2.pnm"]
        args = [f"-{setting}"] + pairs(order, ".jls") + ["n0.pgm", "-o", "n1.jls", "n1.pgm", "-o", "copy.pgm", "n2.pgm", "-o", "x.JLS"]
        rc, out, err = run(CLI, args, d, {"IMCVT_BATCH": batch})
        assert rc == 2, (out, err)          # missing.pgm, and n1.jls "already exist" (written earlier in the same call, no -f)
        for n in order:
            if n != "missing.pgm":
                got = open(os.path.join(d, os.path.splitext(n)[0] + ".jls"), "rb").read()
                assert got == oracle_jls(imgs[n], setting), n
        assert open(os.path.join(d, "x.JLS"), "rb").read() == oracle_jls(imgs["n2.pgm"], setting)
        if os.path.exists(REF):
            rc2, out2, _ = run(REF, args, d2)
            assert (rc2, out2) == (rc, out)


@pytest.mark.parametrize("setting,batch", [(4, "256"), (1, "2")])
def test_many_files_one_call_hevc_and_mixed_outputs(setting, batch):
    with tempfile.TemporaryDirectory() as d, tempfile.TemporaryDirectory() as d2:
        imgs = make_inputs(d)
        make_inputs(d2)
        args = [f"-{setting}", "-f"] + pairs(["n0.pgm", "n1.pgm", "t0.pgm"], ".h265") + ["c0.ppm", "-o", "c0.hevc", "c1.ppm", "-o", "c1.265"]
        args += ["n2.pgm", "-o", "n2.jls", "n3.pgm", "-o", "n3.pgm.out.pgm", "n4.pgm", "-o", "n4.h265", "n0.h265", "-o", "back.pgm",
                 "P2.pnm", "-o", "P2.h265", "t1.pgm", "-o", "t1.txt"]
        rc, out, err = run(CLI, args, d, {"IMCVT_BATCH": batch})
        assert rc == 2, (out, err)          # a .h265 file is not a readable input; unsupported suffix .txt
        assert out.count(b"Only compress the green channel") == 2
        for n, o in [("n0.pgm", "n0.h265"), ("n1.pgm", "n1.h265"), ("t0.pgm", "t0.h265"), ("n4.pgm", "n4.h265"), ("P2.pnm", "P2.h265")]:
            assert open(os.path.join(d, o), "rb").read() == oracle_hevc(imgs[n], setting)[0], o
        for n, o in [("c0.ppm", "c0.hevc"), ("c1.ppm", "c1.265")]:
            assert open(os.path.join(d, o), "rb").read() == oracle_hevc(imgs[n][:, :, 1], setting)[0], o
        assert open(os.path.join(d, "n2.jls"), "rb").read() == oracle_jls(imgs["n2.pgm"], setting)
        if os.path.exists(REF):
            rc2, out2, _ = run(REF, args, d2)
            assert (rc2, out2) == (rc, out)


def test_green_plane_on_the_device_and_through_the_file_api():
    rgb = np.stack([synth.batch(k, 3, 75, 133) for k in ("noise", "natural", "texture")], axis=-1)       # odd pixel count: tail path
    for qpd6 in (0, 4):
        got = imcvt_b200.encode_hevc_green_batch(rgb, qpd6)
        for i in range(3):
            assert got[i] == oracle_hevc(rgb[i, :, :, 1], qpd6)[0], (qpd6, i)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "g.h265")
        assert imcvt_b200.write_hevc_image_file(path, rgb[0], 1, 75, 133, 3) == 0
        assert open(path, "rb").read() == oracle_hevc(rgb[0, :, :, 1], 3)[0]


def test_pinned_staging_buffer_feeds_the_batch_call():
    L = imcvt_b200.lib()
    frames = synth.batch("natural", 9, 64, 200)
    p = L.imcvtHostAlloc(frames.nbytes)
    assert p, imcvt_b200.last_error()
    try:
        C.memmove(p, frames.ctypes.data, frames.nbytes)
        staged = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(frames.nbytes,)).reshape(frames.shape)
        got = imcvt_b200.encode_jls_batch(staged, 0)
        for i in range(9):
            assert got[i] == oracle_jls(frames[i], 0)
    finally:
        L.imcvtHostFree(p)


def test_bmp_qoi_png_inputs_are_staged_into_codec_batches():
    """Row f4 on the device: pictures that arrive as BMP / QOI / PNG (this repo's own readers, host/fmt_*.c) go through the same
    staging and batch calls as PNM files; streams against the oracle, stdout / exit code / every file against the reference CLI."""
    with tempfile.TemporaryDirectory() as d, tempfile.TemporaryDirectory() as d2:
        for dd in (d, d2):
            imgs = make_inputs(dd)
            # the other formats of the same pictures, written by the command line itself (CPU paths; tests/test_cli_formats.py
            # holds them against the reference byte for byte)
            conv = ["-f"]
            for n in ("n0.pgm", "n1.pgm", "c0.ppm", "c1.ppm", "t0.pgm"):
                stem = os.path.splitext(n)[0]
                conv += [n, "-o", stem + ".bmp", n, "-o", stem + ".png", n, "-o", stem + ".qoi"]
            assert run(CLI, conv, dd)[0] == 0
        args = ["-f", "-1", "n0.bmp", "-o", "a0.jls", "n1.png", "-o", "a1.jls", "n1.bmp", "-o", "a2.jls", "c0.png", "-o", "a3.jls", "c1.bmp", "-o", "a4.jls",
                "c0.qoi", "-o", "a5.jls", "n0.qoi", "-o", "a6.jls", "n0.png", "-o", "b0.h265", "t0.bmp", "-o", "b1.h265", "c1.png", "-o", "b2.h265",
                "t0.qoi", "-o", "b3.h265", "a0.jls", "-o", "nope.png"]
        rc, out, err = run(CLI, args, d, {"IMCVT_BATCH": "4"})
        assert rc == 1, (out, err)                                   # a .jls file is not a readable input
        g = lambda n: np.repeat(imgs[n][:, :, None], 3, axis=2)     # a gray picture comes back from QOI as RGB (imageio_qoi.c:149-186)
        want = {"a0.jls": oracle_jls(imgs["n0.pgm"], 1), "a1.jls": oracle_jls(imgs["n1.pgm"], 1), "a2.jls": oracle_jls(imgs["n1.pgm"], 1),
                "a3.jls": oracle_jls(imgs["c0.ppm"], 1), "a4.jls": oracle_jls(imgs["c1.ppm"], 1), "a5.jls": oracle_jls(imgs["c0.ppm"], 1),
                "a6.jls": oracle_jls(g("n0.pgm"), 1),
                "b0.h265": oracle_hevc(imgs["n0.pgm"], 1)[0], "b1.h265": oracle_hevc(imgs["t0.pgm"], 1)[0],
                "b2.h265": oracle_hevc(imgs["c1.ppm"][:, :, 1], 1)[0], "b3.h265": oracle_hevc(imgs["t0.pgm"], 1)[0]}
        for name, stream in want.items():
            assert open(os.path.join(d, name), "rb").read() == stream, name
        if os.path.exists(REF):
            rc2, out2, _ = run(REF, args, d2)
            assert (rc2, out2) == (rc, out)
            for name in want:
                assert open(os.path.join(d, name), "rb").read() == open(os.path.join(d2, name), "rb").read(), name


def test_resident_engine_serves_codec_batches_warm():
    """IMCVT_SERVE / IMCVT_SERVER (host/imcvt_serve.c): the same command lines executed by a resident process -- identical streams,
    text and exit codes, and without the CUDA context creation a cold process pays (the wall-clock times are printed)."""
    import signal
    import time
    with tempfile.TemporaryDirectory() as s, tempfile.TemporaryDirectory() as d, tempfile.TemporaryDirectory() as d2:
        imgs = make_inputs(d)
        make_inputs(d2)
        sock = os.path.join(s, "sock")
        srv = subprocess.Popen([CLI], env=dict(os.environ, IMCVT_SERVE=sock), cwd=s, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
        try:
            for _ in range(1200):
                if os.path.exists(sock):
                    break
                time.sleep(0.05)
            assert os.path.exists(sock)
            names = ["n0.pgm", "n1.pgm", "n2.pgm", "n3.pgm", "n4.pgm", "t0.pgm", "c0.ppm", "c1.ppm"]
            args_j = ["-f", "-0"] + pairs(names, ".jls")
            args_h = ["-f", "-4"] + pairs(["n0.pgm", "n1.pgm", "t0.pgm", "c0.ppm"], ".h265")
            t = time.perf_counter(); cold = run(CLI, args_j, d2); t_cold = time.perf_counter() - t
            walls = []
            for rep in range(3):
                t = time.perf_counter(); warm = run(CLI, args_j, d, {"IMCVT_SERVER": sock}); walls.append(time.perf_counter() - t)
                assert warm[:2] == cold[:2]
                for n in names:
                    assert open(os.path.join(d, os.path.splitext(n)[0] + ".jls"), "rb").read() == oracle_jls(imgs[n], 0), n
            h_cold = run(CLI, args_h, d2)
            h_warm = run(CLI, args_h, d, {"IMCVT_SERVER": sock})
            assert h_warm[:2] == h_cold[:2] and h_warm[0] == 0
            for n in ("n0.pgm", "n1.pgm", "t0.pgm"):
                assert open(os.path.join(d, os.path.splitext(n)[0] + ".h265"), "rb").read() == oracle_hevc(imgs[n], 4)[0], n
            assert open(os.path.join(d, "c0.h265"), "rb").read() == oracle_hevc(imgs["c0.ppm"][:, :, 1], 4)[0]
            print("resident engine: cold process %.3f s, through the server %s s" % (t_cold, ", ".join("%.3f" % w for w in walls)))
            assert min(walls[1:]) < t_cold, (t_cold, walls)
            assert srv.poll() is None
        finally:
            srv.send_signal(signal.SIGTERM)
            assert srv.wait(timeout=120) == 0
