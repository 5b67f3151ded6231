"""CPU: the staging pipeline of the batch command line (imcvt_b200/host/imcvt_cli.c: batches by shape, postponed parallel pixel
reads, hazard flushes, the writer threads, the resident process's arenas) with a TEST DOUBLE in place of libimcvt_b200.so
(tests/tools/fake_engine.c: a "stream" is a tag + the frame's own bytes; every batch call is logged).  What the engine would be
handed and which file its answer lands in is visible without a GPU; the real streams are the GPU tier's business
(tests/test_gpu_cli_batch.py).  The double is built into a temporary directory and never shipped."""
import os
import signal
import struct
import subprocess
import tempfile
import time

import numpy as np
import pytest

from tests._util import GOLDEN, ROOT, load_pnm

HOST = os.path.join(ROOT, "imcvt_b200", "host")
SRCS = ["imcvt_cli.c", "imcvt_serve.c", "pnm_io.c", "fmt_bmp.c", "fmt_qoi.c", "fmt_png.c"]


@pytest.fixture(scope="module")
def cli():
    with tempfile.TemporaryDirectory() as d:
        so = os.path.join(d, "libimcvt_b200.so")
        subprocess.run(["gcc", "-O1", "-g", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests", "tools", "fake_engine.c")], check=True)
        exe = os.path.join(d, "ImCvt_fake")
        subprocess.run(["gcc", "-O1", "-g", "-Wall", "-pthread", "-I" + os.path.join(ROOT, "include"), "-o", exe] + [os.path.join(HOST, f) for f in SRCS]
                       + ["-L" + d, "-limcvt_b200", "-Wl,-rpath," + d], check=True)
        yield exe


def run(exe, args, cwd, **env):
    e = dict(os.environ)
    for k in ("IMCVT_SERVER", "IMCVT_SERVE", "IMCVT_BATCH", "IMCVT_DEVICES", "IMCVT_DEVICE"):
        e.pop(k, None)
    e.update({k: str(v) for k, v in env.items()})
    p = subprocess.run([exe] + args, cwd=cwd, capture_output=True, timeout=600, env=e)
    return p.returncode, p.stdout, p.stderr


def write_pnm(path, img):
    h, w = img.shape[:2]
    with open(path, "wb") as f:
        f.write(b"P%c\n%d %d\n255\n" % (b"6" if img.ndim == 3 else b"5", w, h) + np.ascontiguousarray(img).tobytes())


def parse(stream):
    codec, setting, flag = chr(stream[0]), stream[1], stream[2]
    h, w = struct.unpack("<II", stream[4:12])
    return codec, setting, flag, h, w, stream[16:]


def expect(path, codec, setting, img, green=False):
    c, s, f, h, w, px = parse(open(path, "rb").read())
    want = img[:, :, 1] if green else img
    assert (c, s) == (codec, setting) and (h, w) == img.shape[:2], path
    assert f == (1 if (green or (img.ndim == 3 and codec == "J")) else 0), path
    assert px == np.ascontiguousarray(want).tobytes(), path


def make(d, rng, n=40):
    imgs = {}
    for i in range(n):
        shape = [(33, 47), (33, 47), (33, 47), (20, 64), (64, 20)][i % 5] if i % 7 else (33, 47, 3)
        imgs["f%02d.p%cm" % (i, "p" if len(shape) == 3 else "g")] = rng.integers(0, 256, shape, dtype=np.uint8)
    for name, img in imgs.items():
        write_pnm(os.path.join(d, name), img)
    return imgs


@pytest.mark.parametrize("batch", [256, 5, 1])
def test_every_file_gets_its_own_pixels_whatever_the_batching(cli, batch):
    rng = np.random.default_rng(3)
    with tempfile.TemporaryDirectory() as d:
        imgs = make(d, rng)
        log = os.path.join(d, "calls.log")
        args = ["-2"]
        for k, n in enumerate(imgs):
            args += [n, "-o", n + (".jls" if k % 3 else ".h265")]
        rc, out, err = run(cli, args, d, IMCVT_BATCH=batch, FAKE_ENGINE_LOG=log)
        assert rc == 0, (out, err)
        for k, (n, img) in enumerate(imgs.items()):
            if k % 3:
                expect(os.path.join(d, n + ".jls"), "J", 2, img)
            else:
                expect(os.path.join(d, n + ".h265"), "H", 2, img, green=img.ndim == 3)
        calls = open(log).read().splitlines()
        assert sum(int(c.split("n=")[1].split()[0]) for c in calls) == len(imgs)
        assert all(int(c.split("n=")[1].split()[0]) <= batch for c in calls)
        lines = out.decode().splitlines()
        assert [l for l in lines if l.startswith("(")] == ["(%d/%d)  %s -> %s" % (k + 1, len(imgs), n, n + (".jls" if k % 3 else ".h265")) for k, n in enumerate(imgs)]


def test_hazards_between_postponed_reads_and_later_writes(cli):
    """A later job that overwrites the SOURCE of an earlier, still pending codec job must not change what that job encodes;
    an output used as a later input; short and missing files between good ones."""
    rng = np.random.default_rng(5)
    with tempfile.TemporaryDirectory() as d:
        a, b, c = (rng.integers(0, 256, (30, 50), dtype=np.uint8) for _ in range(3))
        for n, img in (("a.pgm", a), ("b.pgm", b), ("c.pgm", c)):
            write_pnm(os.path.join(d, n), img)
        open(os.path.join(d, "short.pgm"), "wb").write(b"P5\n50 30\n255\n" + bytes(100))
        args = ["-f", "a.pgm", "-o", "a.jls", "b.pgm", "-o", "b.jls",
                "c.pgm", "-o", "a.pgm",                      # PNM output over a source whose pixels are still to be read
                "a.pgm", "-o", "a2.jls",                     # ... and the new a.pgm as an input again
                "short.pgm", "-o", "s.jls", "missing.pgm", "-o", "m.jls",
                "b.pgm", "-o", "b.bmp", "b.bmp", "-o", "b2.jls",
                "c.pgm", "-o", "c.jls", "c.jls", "-o", "never.pgm"]
        rc, out, err = run(cli, args, d, FAKE_ENGINE_LOG=os.path.join(d, "log"))
        assert rc == 3, (out, err)                           # short, missing, and a .jls file is not an input
        expect(os.path.join(d, "a.jls"), "J", 0, a)
        expect(os.path.join(d, "b.jls"), "J", 0, b)
        expect(os.path.join(d, "a2.jls"), "J", 0, c)
        expect(os.path.join(d, "b2.jls"), "J", 0, b)
        expect(os.path.join(d, "c.jls"), "J", 0, c)
        assert load_pnm(os.path.join(d, "a.pgm"))[0].tobytes() == c.tobytes()
        text = out.decode()
        assert "***ERROR: open short.pgm failed" in text and "***ERROR: missing.pgm not exist" in text and "***ERROR: open c.jls failed" in text
        assert not os.path.exists(os.path.join(d, "s.jls")) and not os.path.exists(os.path.join(d, "m.jls"))


def test_batches_are_offered_to_several_devices_only_when_worth_it(cli):
    rng = np.random.default_rng(9)
    with tempfile.TemporaryDirectory() as d:
        imgs = {"g%02d.pgm" % i: rng.integers(0, 256, (16, 16), dtype=np.uint8) for i in range(40)}
        for n, img in imgs.items():
            write_pnm(os.path.join(d, n), img)
        log = os.path.join(d, "log")
        args = []
        for n in imgs:
            args += [n, "-o", n + ".jls"]
        assert run(cli, ["-f"] + args, d, FAKE_ENGINE_LOG=log, FAKE_ENGINE_DEVICES=8)[0] == 0
        assert open(log).read().splitlines() == ["jls n=40 rgb=0 h=16 w=16 setting=0 ndev=5"]            # at least 8 frames per device
        os.remove(log)
        assert run(cli, ["-f"] + args[:9], d, FAKE_ENGINE_LOG=log, FAKE_ENGINE_DEVICES=8)[0] == 0
        assert open(log).read().splitlines() == ["jls n=3 rgb=0 h=16 w=16 setting=0 ndev=0"]             # too small to shard
        os.remove(log)
        assert run(cli, ["-f"] + args, d, FAKE_ENGINE_LOG=log, FAKE_ENGINE_DEVICES=8, IMCVT_DEVICES="2")[0] == 0
        assert open(log).read().splitlines() == ["jls n=40 rgb=0 h=16 w=16 setting=0 ndev=2"]
        for n, img in imgs.items():
            expect(os.path.join(d, n + ".jls"), "J", 0, img)


def test_resident_process_reuses_its_arenas_across_requests_of_different_shapes(cli):
    rng = np.random.default_rng(11)
    with tempfile.TemporaryDirectory() as s, tempfile.TemporaryDirectory() as d:
        sock = os.path.join(s, "sock")
        env = dict(os.environ, IMCVT_SERVE=sock)
        srv = subprocess.Popen([cli], env=env, cwd=s, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        try:
            for _ in range(400):
                if os.path.exists(sock):
                    break
                time.sleep(0.02)
            for round_, shape in enumerate([(20, 30), (200, 300), (8, 8, 3), (200, 300), (64, 64)]):
                imgs = {"r%d_%02d.%s" % (round_, i, "ppm" if len(shape) == 3 else "pgm"): rng.integers(0, 256, shape, dtype=np.uint8) for i in range(12)}
                args = ["-f", "-%d" % (round_ % 5)]
                for n, img in imgs.items():
                    write_pnm(os.path.join(d, n), img)
                    args += [n, "-o", n + ".jls"]
                rc, out, err = run(cli, args, d, IMCVT_SERVER=sock, IMCVT_BATCH=5)
                assert rc == 0, (out, err)
                for n, img in imgs.items():
                    expect(os.path.join(d, n + ".jls"), "J", round_ % 5, img)
            assert srv.poll() is None
        finally:
            srv.send_signal(signal.SIGTERM)
            assert srv.wait(timeout=60) == 0


def test_many_small_files_in_one_batch_and_a_tight_descriptor_limit(cli):
    """300 files in one batch (postponed reads are carried out in lots), once under RLIMIT_NOFILE = 40: every file still gets its
    own pixels and nothing fails for want of descriptors."""
    import resource
    rng = np.random.default_rng(13)
    with tempfile.TemporaryDirectory() as d:
        imgs = {"s%03d.pgm" % i: rng.integers(0, 256, (8, 8), dtype=np.uint8) for i in range(300)}
        args = ["-f"]
        for n, img in imgs.items():
            write_pnm(os.path.join(d, n), img)
            args += [n, "-o", n + ".jls"]
        for limit in (None, 40):
            for n in imgs:
                if os.path.exists(os.path.join(d, n + ".jls")):
                    os.remove(os.path.join(d, n + ".jls"))
            e = {k: v for k, v in os.environ.items() if not k.startswith("IMCVT_")}
            e["IMCVT_BATCH"] = "999"
            pre = (lambda: resource.setrlimit(resource.RLIMIT_NOFILE, (limit, limit))) if limit else None
            p = subprocess.run([cli] + args, cwd=d, capture_output=True, timeout=600, env=e, preexec_fn=pre)
            assert p.returncode == 0, (limit, p.stdout[-300:], p.stderr[-300:])
            for n, img in imgs.items():
                expect(os.path.join(d, n + ".jls"), "J", 0, img)
