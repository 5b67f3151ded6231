"""CPU: the resident engine of the batch command line (imcvt_b200/host/imcvt_serve.c; SURVEY 8(f) row f2 -- a one-shot
process pays CUDA context creation per invocation).  `IMCVT_SERVE=<socket> ImCvt_b200` serves,
`IMCVT_SERVER=<socket> ImCvt_b200 <args>` is the ordinary command line executed by the server's process with the client's
stdout / stderr / working directory.  Everything a command line shows -- text, exit code, files -- must be what the same
command shows when it runs in-process (and, where oracle/_ref/ImCvt_ref exists, what the reference's CLI shows).  No GPU
here: codec outputs fail through the server exactly as they fail in-process (no CPU fallback)."""
import os
import signal
import socket
import subprocess
import tempfile
import time

import pytest

from tests._util import ROOT
from tests.test_cli_batch import CASES, CLI, REF, stage


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(CLI):
        import __graft_entry__
        __graft_entry__.build()


class Server:
    def __init__(self, d):
        self.path = os.path.join(d, "s")
        self.err = open(os.path.join(d, "server.err"), "wb")
        env = dict(os.environ, IMCVT_SERVE=self.path)
        env.pop("IMCVT_SERVER", None)
        self.p = subprocess.Popen([CLI], env=env, cwd=d, stdout=subprocess.DEVNULL, stderr=self.err)
        up = False
        for _ in range(600):                       # engine start: a CUDA context on a GPU box, an error message here
            c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            try:
                c.connect(self.path)               # (a connection that says nothing: the server drops it)
                up = True
            except OSError:
                time.sleep(0.05)
            finally:
                c.close()
            if up:
                break
        assert up, "server did not come up"

    def stop(self):
        self.p.send_signal(signal.SIGTERM)
        rc = self.p.wait(timeout=60)
        self.err.close()
        return rc


def run(args, cwd, server=None, extra_env=None):
    env = dict(os.environ)
    env.pop("IMCVT_SERVER", None)
    env.pop("IMCVT_SERVE", None)
    if server:
        env["IMCVT_SERVER"] = server
    env.update(extra_env or {})
    p = subprocess.run([CLI] + args, cwd=cwd, capture_output=True, timeout=600, env=env)
    return p.returncode, p.stdout, p.stderr


def test_commands_through_the_server_equal_in_process_runs():
    with tempfile.TemporaryDirectory() as s, tempfile.TemporaryDirectory() as d, tempfile.TemporaryDirectory() as d2, tempfile.TemporaryDirectory() as d3:
        srv = Server(s)
        try:
            for dd in (d, d2, d3):
                stage(dd)
            for args in CASES + [[], ["-f", "P5.pnm", "-o", "q.bmp", "q.bmp", "-o", "q.png", "q.png", "-o", "q.qoi", "q.qoi", "-o", "q.ppm"],
                                 ["P5.pnm", "-o", "codec.jls", "P6.pnm", "-o", "codec.h265"]]:
                got = run(args, d, srv.path)
                want = run(args, d2)
                assert got[:2] == want[:2], args
                assert (b"no CUDA device" in got[2]) == (b"no CUDA device" in want[2])     # stderr reaches the client too
                if os.path.exists(REF) and "codec.jls" not in args:
                    p = subprocess.run([REF] + args, cwd=d3, capture_output=True, timeout=600)
                    assert (p.returncode, p.stdout) == got[:2], args
            names = sorted(os.listdir(d))
            assert names == sorted(os.listdir(d2)) and len(names) > 15
            for n in names:
                assert open(os.path.join(d, n), "rb").read() == open(os.path.join(d2, n), "rb").read(), n
        finally:
            assert srv.stop() == 0
        assert not os.path.exists(srv.path)            # the socket is removed on SIGTERM


def test_relative_paths_follow_the_clients_directory_and_requests_queue():
    with tempfile.TemporaryDirectory() as s, tempfile.TemporaryDirectory() as d:
        srv = Server(s)
        try:
            subs = []
            for k in range(4):
                sub = os.path.join(d, "w%d" % k)
                os.mkdir(sub)
                stage(sub)
                subs.append(sub)
            env = dict(os.environ, IMCVT_SERVER=srv.path)
            ps = [subprocess.Popen([CLI, "-f", "P%d.pnm" % (2 + k), "-o", "out.bmp", "out.bmp", "-o", "back.pnm"], cwd=sub, env=env,
                                   stdout=subprocess.PIPE, stderr=subprocess.PIPE) for k, sub in enumerate(subs)]
            for k, (p, sub) in enumerate(zip(ps, subs)):
                out, _ = p.communicate(timeout=120)
                assert p.returncode == 0 and out.count(b"\n") == 4, out
                assert sorted(os.listdir(sub)).count("out.bmp") == 1 and os.path.exists(os.path.join(sub, "back.pnm"))
            assert not os.path.exists(os.path.join(s, "out.bmp"))
        finally:
            srv.stop()


def test_no_server_on_the_path_means_an_ordinary_run_and_the_server_survives_bad_clients():
    with tempfile.TemporaryDirectory() as s, tempfile.TemporaryDirectory() as d:
        stage(d)
        rc, out, _ = run(["-f", "P5.pnm", "-o", "a.pgm"], d, os.path.join(s, "nobody-listens"))
        assert rc == 0 and b"(1/1)  P5.pnm -> a.pgm" in out
        srv = Server(s)
        try:
            for junk in (b"", b"garbage that is not a request", b"\x49\x4d\x56\x31" + b"\xff" * 64):
                c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
                c.connect(srv.path)
                if junk:
                    c.sendall(junk)
                c.close()
            rc, out, _ = run(["-f", "P6.pnm", "-o", "b.ppm"], d, srv.path)
            assert rc == 0 and b"(1/1)  P6.pnm -> b.ppm" in out and srv.p.poll() is None
            # a client whose stdout is a closed pipe must not take the server down (SIGPIPE ignored)
            p = subprocess.Popen([CLI, "-f", "P5.pnm", "-o", "c.pgm", "P6.pnm", "-o", "c.ppm"], cwd=d, env=dict(os.environ, IMCVT_SERVER=srv.path),
                                 stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
            p.stdout.close()
            p.wait(timeout=60)
            rc, out, _ = run(["-f", "P5.pnm", "-o", "d.pgm"], d, srv.path)
            assert rc == 0 and srv.p.poll() is None
        finally:
            srv.stop()


def test_a_second_server_does_not_take_a_live_servers_path_but_takes_a_stale_one():
    with tempfile.TemporaryDirectory() as s, tempfile.TemporaryDirectory() as d:
        stage(d)
        srv = Server(s)
        try:
            env = dict(os.environ, IMCVT_SERVE=srv.path)
            env.pop("IMCVT_SERVER", None)
            second = subprocess.run([CLI], env=env, cwd=s, capture_output=True, timeout=60)
            assert second.returncode == 1 and b"already listening" in second.stderr
            rc, out, _ = run(["-f", "P5.pnm", "-o", "a.pgm"], d, srv.path)          # the first server is untouched
            assert rc == 0 and srv.p.poll() is None
        finally:
            srv.p.kill()                                                            # dies without removing its socket file
            srv.p.wait(timeout=60)
            srv.err.close()
        assert os.path.exists(srv.path)
        third = Server(s)                                                           # the stale file is replaced
        try:
            rc, out, _ = run(["-f", "P6.pnm", "-o", "b.ppm"], d, third.path)
            assert rc == 0 and b"(1/1)  P6.pnm -> b.ppm" in out
        finally:
            assert third.stop() == 0
