"""End-to-end timing of the batch command line (files in -> files out), next to the reference's CLI on the same files.
Wall clock around the whole process (CUDA context creation, file reads, H2D, kernels, D2H, file writes): this is what a
user of the binary sees, NOT a kernel number (those come from bench.py).  Usage: python tools/cli_bench.py [n] [codec]"""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imcvt_b200 import synth  # noqa: E402

CLI = os.path.join(ROOT, "imcvt_b200", "ImCvt_b200")
REF = os.path.join(ROOT, "oracle", "_ref", "ImCvt_ref")


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    codec = sys.argv[2] if len(sys.argv) > 2 else "jls"
    suffix, switch = (".jls", "-0") if codec == "jls" else (".h265", "-4")
    h, w = 1080, 1920
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.TemporaryDirectory(dir=base) as d:
        frames = synth.batch("natural", n, h, w)
        for i in range(n):
            with open(os.path.join(d, f"f{i:03d}.pgm"), "wb") as f:
                f.write(b"P5\n%d %d\n255\n" % (w, h))
                f.write(frames[i].tobytes())
        args = [switch, "-f"]
        for i in range(n):
            args += [f"f{i:03d}.pgm", "-o", f"f{i:03d}{suffix}"]
        res = {"codec": codec, "frames": n, "shape": [h, w]}
        for rep in range(2):                    # second run: files in the page cache, driver warm
            t0 = time.perf_counter()
            p = subprocess.run([CLI] + args, cwd=d, capture_output=True, env=dict(os.environ, IMCVT_TIMING="1"))
            dt = time.perf_counter() - t0
            assert p.returncode == 0, (p.stdout[-400:], p.stderr[-400:])
            res[f"batch_cli_s_run{rep}"] = round(dt, 3)
            res[f"phases_run{rep}"] = p.stderr.decode().strip().splitlines()[-1]
        res["batch_cli_mpixel_s"] = round(n * h * w / dt / 1e6, 1)
        # the same command line through a resident engine (IMCVT_SERVE / IMCVT_SERVER, host/imcvt_serve.c): no context creation
        sock = os.path.join(d, "sock")
        srv = subprocess.Popen([CLI], cwd=d, env=dict(os.environ, IMCVT_SERVE=sock), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        try:
            for _ in range(1200):
                if os.path.exists(sock):
                    break
                time.sleep(0.05)
            cold = {f: open(os.path.join(d, f), "rb").read() for f in os.listdir(d) if f.endswith(suffix)}
            for rep in range(3):                # the first request still pins the staging arena and sizes the workspaces
                t0 = time.perf_counter()
                p = subprocess.run([CLI] + args, cwd=d, capture_output=True, env=dict(os.environ, IMCVT_TIMING="1", IMCVT_SERVER=sock))
                dt = time.perf_counter() - t0
                assert p.returncode == 0, (p.stdout[-400:], p.stderr[-400:])
                res[f"resident_cli_s_run{rep}"] = round(dt, 3)
                res[f"resident_phases_run{rep}"] = p.stderr.decode().strip().splitlines()[-1]
            res["resident_cli_mpixel_s"] = round(n * h * w / dt / 1e6, 1)
            res["resident_same_bytes_as_cold"] = all(open(os.path.join(d, f), "rb").read() == b for f, b in cold.items())
        finally:
            srv.terminate()
            srv.wait(timeout=120)
        if os.path.exists(REF):
            m = 4 if codec == "jls" else 1
            rargs = [switch, "-f"]
            for i in range(m):
                rargs += [f"f{i:03d}.pgm", "-o", f"r{i:03d}{suffix}"]
            t0 = time.perf_counter()
            p = subprocess.run([REF] + rargs, cwd=d, capture_output=True)
            dt = time.perf_counter() - t0
            assert p.returncode == 0
            res["ref_cli_files"] = m
            res["ref_cli_s"] = round(dt, 3)
            res["ref_cli_mpixel_s_one_thread"] = round(m * h * w / dt / 1e6, 2)
            same = all(open(os.path.join(d, f"f{i:03d}{suffix}"), "rb").read() == open(os.path.join(d, f"r{i:03d}{suffix}"), "rb").read()
                       for i in range(m))
            res["same_bytes_as_reference_cli"] = same
        print(json.dumps(res))


if __name__ == "__main__":
    main()
