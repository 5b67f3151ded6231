"""GPU, 2 or more devices on the box (skipped below that): the multi-device entry points (imcvtEncodeBatch*Multi) and the
batch command line shard a batch over the devices in contiguous frame blocks -- one process, one engine and one host thread
per device, no collective (the reference's file loop src/main.c:162-211 spread over GPUs).  The bytes and the
size tables must not depend on the number of devices."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import imcvt_b200
from imcvt_b200 import synth
from tests._util import ROOT, check_hevc, check_jls

pytestmark = pytest.mark.gpu

CLI = os.path.join(ROOT, "imcvt_b200", "ImCvt_b200")


def ndev():
    return imcvt_b200.device_count()


@pytest.fixture(autouse=True)
def need_two_devices():
    if ndev() < 2:
        pytest.skip("needs at least 2 visible CUDA devices")


def device_sets():
    n = ndev()
    sets = [[0], [1], [0, 1], [1, 0], None]
    if n >= 4:
        sets += [4, [3, 1]]
    if n >= 8:
        sets += [8]
    return sets


def test_jls_bytes_do_not_depend_on_the_number_of_devices():
    for near, rgb, n, h, w in ((0, False, 13, 96, 160), (2, True, 7, 64, 88), (0, False, 3, 33, 65)):
        frames = synth.batch("natural", n, h, w, rgb=rgb)
        one = imcvt_b200.encode_jls_batch(frames, near)
        assert one[0] == check_jls(frames[0], near) and one[-1] == check_jls(frames[-1], near)
        for devs in device_sets():
            got = imcvt_b200.encode_jls_batch(frames, near, devices=devs)
            assert got == one, (near, rgb, devs)


def test_full_hd_jls_batch_on_all_devices_from_page_locked_buffers():
    """The bench's configs[1] batch, page-locked (portable) buffers, sharded over every device: blocks run the in-place kernel."""
    import ctypes as C
    n, h, w = 64, 1080, 1920
    frames = synth.batch("natural", n, h, w)
    L = imcvt_b200.lib()
    stride = 2 * h * w + 4096
    p_in, p_out = L.imcvtHostAlloc(n * h * w), L.imcvtHostAlloc(n * stride)
    assert p_in and p_out
    try:
        C.memmove(p_in, frames.ctypes.data, n * h * w)
        lens = (C.c_size_t * n)()
        rc = L.imcvtEncodeBatchJLSMulti(p_in, h * w, n, 0, h, w, 0, p_out, stride, C.cast(lens, C.c_void_p), None, 0)
        assert rc == 0, imcvt_b200.last_error()
        out = np.ctypeslib.as_array(C.cast(p_out, C.POINTER(C.c_uint8)), shape=(n * stride,))
        one = imcvt_b200.encode_jls_batch(frames, 0, out_stride=stride)
        for i in range(n):
            assert out[i * stride:i * stride + lens[i]].tobytes() == one[i], i
    finally:
        L.imcvtHostFree(p_in)
        L.imcvtHostFree(p_out)


def test_hevc_bytes_and_reconstructions_do_not_depend_on_the_number_of_devices():
    frames = synth.batch("texture", 9, 72, 136)
    one, rc1 = imcvt_b200.encode_hevc_batch(frames, 3)
    assert one[4] == check_hevc(frames[4], 3)[0]
    for devs in device_sets():
        got, rc = imcvt_b200.encode_hevc_batch(frames, 3, devices=devs)
        assert got == one and np.array_equal(rc, rc1), devs
    sw1 = imcvt_b200.encode_hevc_sweep(frames[:5], [1, 4, 2])
    for devs in device_sets():
        assert imcvt_b200.encode_hevc_sweep(frames[:5], [1, 4, 2], devices=devs) == sw1, devs
    rgb = synth.batch("natural", 5, 40, 72, rgb=True)
    g1 = imcvt_b200.encode_hevc_green_batch(rgb, 2)
    assert imcvt_b200.encode_hevc_green_batch(rgb, 2, devices=None) == g1


def test_bad_device_lists_fail_loudly():
    frames = synth.batch("natural", 2, 16, 16)
    for devs in ([0, 0], [ndev()], [-1]):
        with pytest.raises(imcvt_b200.ImcvtError):
            imcvt_b200.encode_jls_batch(frames, 0, devices=devs)


def test_command_line_writes_the_same_files_on_one_and_on_all_devices():
    frames = synth.batch("natural", 40, 48, 80)
    with tempfile.TemporaryDirectory() as d:
        args_by_env = {}
        for tag, env in (("one", {"IMCVT_DEVICE": "0"}), ("all", {"IMCVT_DEVICES": "all"}), ("two", {"IMCVT_DEVICES": "1,0"})):
            args = []
            for i in range(40):
                src = os.path.join(d, f"f{i}.pgm")
                if not os.path.exists(src):
                    with open(src, "wb") as f:
                        f.write(b"P5\n80 48\n255\n" + frames[i].tobytes())
                args += [src, "-o", os.path.join(d, f"{tag}{i}.{'jls' if i < 24 else 'h265'}")]
            r = subprocess.run([CLI, "-2"] + args, capture_output=True, env={**os.environ, **env})
            assert r.returncode == 0, r.stdout.decode() + r.stderr.decode()
            args_by_env[tag] = r.stdout.decode().replace(f"{tag}", "X")
        assert args_by_env["one"] == args_by_env["all"] == args_by_env["two"]
        for i in range(40):
            ext = "jls" if i < 24 else "h265"
            a = open(os.path.join(d, f"one{i}.{ext}"), "rb").read()
            assert a == open(os.path.join(d, f"all{i}.{ext}"), "rb").read() == open(os.path.join(d, f"two{i}.{ext}"), "rb").read(), i
            assert a == (check_jls(frames[i], 2) if i < 24 else check_hevc(frames[i], 2)[0]), i
