"""GPU: the reference's OWN command line (src/main.c + its PNM/BMP/QOI/PNG loaders, compiled unmodified by oracle/Makefile)
linked against libimcvt_b200.so instead of its imageio_jls.c / imageio_hevc.c / HEVCe.c -- the drop-in of
INTEGRATION.md section 1.  Outputs must be the reference's bytes, stdout and exit codes the reference's own."""
import os
import subprocess
import tempfile

import pytest

from tests._util import GOLDEN, ROOT, golden_table, sha256

pytestmark = pytest.mark.gpu

CLI = os.path.join(ROOT, "oracle", "_ref", "ImCvt_b200")
REF = os.path.join(ROOT, "oracle", "_ref", "ImCvt_ref")


def run(exe, args, cwd):
    p = subprocess.run([exe] + args, cwd=cwd, capture_output=True, text=True, timeout=600)
    return p.returncode, p.stdout


@pytest.mark.skipif(not os.path.exists(CLI), reason="oracle/_ref/ImCvt_b200 not built (needs /root/reference at build time)")
def test_reference_cli_on_the_b200_engine_writes_the_reference_streams():
    table = {(e["codec"], e["fixture"], e["setting"]): e for e in golden_table()}
    with tempfile.TemporaryDirectory() as d:
        for fixture in ("P5", "P6"):
            src = os.path.join(GOLDEN, fixture + ".pnm")
            for setting in (0, 2, 4):
                for codec, suffix in (("jls", ".jls"), ("hevc", ".h265")):
                    out = os.path.join(d, f"{fixture}_{setting}{suffix}")
                    rc, text = run(CLI, [f"-{setting}", "-f", src, "-o", out], d)
                    assert rc == 0, text
                    e = table[(codec, fixture, setting)]
                    data = open(out, "rb").read()
                    assert len(data) == e["length"] and sha256(data) == e["sha256"], (fixture, codec, setting)
        # several files in one call, one of them unreadable: the reference's per-file error path and exit code (main.c:223)
        a, b = os.path.join(d, "a.jls"), os.path.join(d, "b.h265")
        args = ["-1", "-f", os.path.join(GOLDEN, "P5.pnm"), "-o", a, os.path.join(d, "missing.pnm"), "-o", os.path.join(d, "m.jls"),
                os.path.join(GOLDEN, "P2.pnm"), "-o", b]
        rc, text = run(CLI, args, d)
        assert rc == 1
        if os.path.exists(REF):
            a2, b2 = os.path.join(d, "a2.jls"), os.path.join(d, "b2.h265")
            args2 = [x.replace(a, a2).replace(b, b2) for x in args]
            rc2, text2 = run(REF, args2, d)
            assert rc2 == rc
            assert text2.replace("a2.jls", "a.jls").replace("b2.h265", "b.h265") == text
            assert open(a, "rb").read() == open(a2, "rb").read() and open(b, "rb").read() == open(b2, "rb").read()
