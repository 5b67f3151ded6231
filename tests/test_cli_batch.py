"""CPU: the batch command line (imcvt_b200/host/imcvt_cli.c + pnm_io.c) -- everything that does not need the device:
usage text and exit codes, per-file error lines, PNM in / PNM out, and that codec outputs FAIL (no CPU fallback) on a box
without a GPU.  Where oracle/_ref/ImCvt_ref (the reference's CLI, compiled unmodified) exists, stdout, exit codes and
files are compared with it byte for byte (reference behaviour: src/main.c:87-223, src/imageio_pnm.c)."""
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

from tests._util import GOLDEN, ROOT, load_pnm

CLI = os.path.join(ROOT, "imcvt_b200", "ImCvt_b200")
REF = os.path.join(ROOT, "oracle", "_ref", "ImCvt_ref")


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(CLI):
        import __graft_entry__
        __graft_entry__.build()


def run(exe, args, cwd):
    p = subprocess.run([exe] + args, cwd=cwd, capture_output=True, timeout=600)
    return p.returncode, p.stdout, p.stderr


def stage(d):
    for f in ("P2", "P3", "P4", "P5", "P6"):
        shutil.copy(os.path.join(GOLDEN, f + ".pnm"), os.path.join(d, f + ".pnm"))
    # plain PBM with a comment, samples glued together ("11" is ONE sample for the reference's reader, imageio_pnm.c:134-138),
    # a raw PGM whose header line carries trailing text, and a plain PGM with samples above 255
    open(os.path.join(d, "p1.pbm"), "w").write("P1\n# c\n3 2\n1 0 1\n0 11 0\n")
    open(os.path.join(d, "p5c.pgm"), "wb").write(b"P5 # x\n4 2 # y\n200 trailing\n" + bytes(range(8)))
    open(os.path.join(d, "p2big.pgm"), "w").write("P2\n2 2\n255\n0 256 511 7\n")
    open(os.path.join(d, "short.pgm"), "wb").write(b"P5\n4 4\n255\n" + bytes(5))
    open(os.path.join(d, "maxval.pgm"), "wb").write(b"P5\n2 2\n65535\n" + bytes(8))


def test_usage_and_exit_code():
    with tempfile.TemporaryDirectory() as d:
        rc, out, _ = run(CLI, [], d)
        assert rc == 255 and out.count(b"\n") == 20 and b"ImCvt [-switches]  <in1> -o <out1>" in out
        rc2, out2, _ = run(CLI, ["-f", "-3"], d)
        assert (rc2, out2) == (rc, out)
        if os.path.exists(REF):
            assert run(REF, [], d)[:2] == (rc, out)
        # "-o name" in front of any input: ignored (the reference indexes dst_fnames[-1] there)
        assert run(CLI, ["-o", "x.pgm"], d)[:2] == (rc, out)


CASES = [
    ["-f", "P2.pnm", "-o", "a.pgm", "P3.pnm", "-o", "b.ppm", "P4.pnm", "-o", "c.pnm", "p1.pbm", "-o", "d.PGM",
     "missing.pnm", "-o", "e.pgm", "P5.pnm", "-o", "f.xyz", "P6.pnm", "-o", "g.ppm", "short.pgm", "-o", "h.pgm",
     "maxval.pgm", "-o", "i.pgm", "p5c.pgm", "-o", "j.pgm", "p2big.pgm", "-o", "k.pgm"],
    ["P5.pnm", "-o", "a.pgm"],                                  # exists now, no -f: refused
    ["P5.pnm", "-o", "l.pgm", "l.pgm", "-o", "m.pgm", "P2.pnm", "-o", "l.pgm"],     # output reused as input; same name twice
    ["P5.pnm", "-fo", "n.ppm"],                                 # switches may be glued (main.c:103-110)
]


def test_pnm_conversions_and_error_lines():
    with tempfile.TemporaryDirectory() as d, tempfile.TemporaryDirectory() as d2:
        stage(d)
        stage(d2)
        for args in CASES:
            rc, out, _ = run(CLI, args, d)
            if os.path.exists(REF):
                rc2, out2, _ = run(REF, args, d2)
                assert (rc, out) == (rc2, out2), args
        want = {"a.pgm": "P2.pnm", "b.ppm": "P3.pnm", "c.pnm": "P4.pnm", "d.PGM": "p1.pbm", "g.ppm": "P6.pnm", "j.pgm": "p5c.pgm",
                "k.pgm": "p2big.pgm", "l.pgm": "P5.pnm", "m.pgm": "P5.pnm", "n.ppm": "P5.pnm"}
        for name, src in want.items():
            img, is_rgb = load_pnm(os.path.join(d, src))
            h, w = img.shape[:2]
            head = b"P%c\n%d %d\n255\n" % (b"6" if is_rgb else b"5", w, h)
            assert open(os.path.join(d, name), "rb").read() == head + img.tobytes(), name
            if os.path.exists(REF):
                assert open(os.path.join(d2, name), "rb").read() == head + img.tobytes(), name
        for name in ("e.pgm", "f.xyz", "h.pgm", "i.pgm"):
            assert not os.path.exists(os.path.join(d, name))
        assert np.array_equal(load_pnm(os.path.join(d, "d.PGM"))[0], np.array([[0, 255, 0], [255, 0, 255]], dtype=np.uint8))
        assert load_pnm(os.path.join(d, "k.pgm"))[0].tolist() == [[0, 0], [255, 7]]


def test_default_output_name_needs_the_png_writer_of_the_reference():
    # no -o: "<src>.png" (main.c:171-175).  PNG is one of the reference's format files (out of scope):
    # the stand-alone build has no writer for it and says so with the reference's own error line
    with tempfile.TemporaryDirectory() as d:
        stage(d)
        rc, out, _ = run(CLI, ["sub.dir/../P5.pnm", "P2.pnm", "-o", "q.pgm"], d)
        text = out.decode()
        assert "(1/2)  sub.dir/../P5.pnm -> sub.dir/../P5.png\n" in text
        assert rc == 1 and "summary:  1 file converted  1 failed" in text


def test_codec_outputs_fail_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with tempfile.TemporaryDirectory() as d:
        stage(d)
        rc, out, err = run(CLI, ["P5.pnm", "-o", "a.jls", "P2.pnm", "-o", "b.pgm", "P6.pnm", "-o", "c.h265", "P5.pnm", "-o", "d.jls"], d)
        text = out.decode()
        assert rc == 3
        assert text == ("(1/4)  P5.pnm -> a.jls\n   ***ERROR: write a.jls failed\n(2/4)  P2.pnm -> b.pgm\n(3/4)  P6.pnm -> c.h265\n"
                        "   warning: this HEVCencoder currently only support gray 8-bit image instead of RGB image. "
                        "Only compress the green channel of this image.\n   ***ERROR: write c.h265 failed\n"
                        "(4/4)  P5.pnm -> d.jls\n   ***ERROR: write d.jls failed\n\nsummary:  1 file converted  3 failed\n")
        assert b"no CPU fallback" in err
        assert not os.path.exists(os.path.join(d, "a.jls")) and not os.path.exists(os.path.join(d, "c.h265"))
