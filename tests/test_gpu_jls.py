"""GPU: parity of the CUDA JPEG-LS path (through the C-ABI) with the oracle and the golden fixtures."""
import os
import tempfile

import numpy as np
import pytest

import imcvt_b200
from imcvt_b200 import synth
from tests._util import GOLDEN, golden_table, load_pnm, sha256
from tests._util import check_jls as oracle_jls      # the unmodified reference (oracle/_ref) when present, else the restatement
from tests.test_oracle_jls import edge_images

pytestmark = pytest.mark.gpu


def first_diff(a, b):
    n = min(len(a), len(b))
    for i in range(n):
        if a[i] != b[i]:
            return f"first difference at byte {i} of {len(a)}/{len(b)}"
    return f"lengths differ: {len(a)} vs {len(b)}"


def check(img, near):
    got = imcvt_b200.encode_jls_batch(img[None], near)[0]
    want = oracle_jls(img, near)
    assert got == want, first_diff(got, want)


@pytest.fixture(params=[2, 3, 1], ids=["tile-kernel", "cluster-pair", "multi-kernel"], autouse=True)
def lossless_path(request):
    """Every test runs on all NEAR=0 device paths: the shared-memory tile kernel with one CTA per plane, the same on a
    cluster of two CTAs per plane (the default for small batches), and the multi-kernel pipeline that planes wider
    than 6144 use."""
    imcvt_b200.debug_jls_path(request.param)
    yield request.param
    imcvt_b200.debug_jls_path(0)


@pytest.mark.parametrize("serial", [False, True], ids=["pipeline", "serial"])
@pytest.mark.parametrize("entry", [e for e in golden_table() if e["codec"] == "jls"],
                         ids=lambda e: f"{e['fixture']}-near{e['setting']}")
def test_known_answers(entry, serial, lossless_path):
    if serial and lossless_path != 2:
        pytest.skip("the serial kernel does not depend on the lossless path")
    img, _ = load_pnm(os.path.join(GOLDEN, entry["fixture"] + ".pnm"))
    imcvt_b200.debug_jls_force_serial(serial)
    try:
        s = imcvt_b200.encode_jls_batch(img[None], entry["setting"])[0]
    finally:
        imcvt_b200.debug_jls_force_serial(False)
    assert len(s) == entry["length"]
    assert sha256(s) == entry["sha256"]


@pytest.mark.parametrize("near", [0, 1, 2, 3, 4])
def test_edge_cases_match_oracle(near):
    for name, img in edge_images():
        got = imcvt_b200.encode_jls_batch(img[None], near)[0]
        want = oracle_jls(img, near)
        assert got == want, f"{name}: {first_diff(got, want)}"


def test_random_shapes_lossless():
    rng = np.random.default_rng(11)
    for _ in range(40):
        h, w = int(rng.integers(1, 70)), int(rng.integers(1, 300))
        kind = rng.integers(0, 4)
        if kind == 0:
            img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        elif kind == 1:
            img = rng.integers(0, 4, (h, w), dtype=np.uint8) * 85          # few levels: many runs + interruptions
        elif kind == 2:
            img = np.repeat(rng.integers(0, 256, (h, (w + 15) // 16), dtype=np.uint8), 16, axis=1)[:, :w].copy()
        else:
            img = (rng.integers(0, 2, (h, w)) * 255).astype(np.uint8)       # 0xFF stuffing stress
        check(img, 0)


def test_batch_of_synthetic_frames_all_kinds():
    for kind in ["natural", "texture", "noise", "constant"]:
        frames = synth.batch(kind, 3, 135, 240)
        got = imcvt_b200.encode_jls_batch(frames, 0)
        for i in range(3):
            want = oracle_jls(frames[i], 0)
            assert got[i] == want, f"{kind}[{i}]: {first_diff(got[i], want)}"


def test_rgb_batch_lossless_and_lossy():
    frames = synth.batch("natural", 2, 90, 160, rgb=True)
    for near in (0, 2):
        got = imcvt_b200.encode_jls_batch(frames, near)
        for i in range(2):
            want = oracle_jls(frames[i], near)
            assert got[i] == want, f"near={near}[{i}]: {first_diff(got[i], want)}"


@pytest.mark.parametrize("path", [0, 5], ids=["candidate-table", "three-warp-rowscan"])
def test_near_lossy_kernels_match_oracle(path, lossless_path):
    """NEAR 1..4 on both kernels that code it: the candidate-table kernel (jls_nearscan_kernel.cuh, the default) and the
    three-warp row scan (imcvtDebugJLSPath 5): rows of one and of several segments, run-heavy and clamped content."""
    if lossless_path != 2:
        pytest.skip("NEAR > 0 does not depend on the lossless path")
    rng = np.random.default_rng(33)
    cases = [synth.frame("natural", 1, 40, 1500), synth.frame("texture", 2, 9, 4100), synth.frame("natural", 3, 7, 1024),
             synth.frame("natural", 4, 5, 2049), rng.integers(0, 256, (6, 2500), dtype=np.uint8),
             np.full((5, 3000), 200, np.uint8), (100 + rng.integers(0, 3, (20, 700))).astype(np.uint8),
             rng.integers(0, 4, (16, 600), dtype=np.uint8), (252 + rng.integers(0, 4, (16, 600))).astype(np.uint8),
             (rng.integers(0, 2, (40, 333)) * 255).astype(np.uint8), synth.frame("texture", 5, 1, 1), synth.frame("noise", 5, 1, 700)]
    imcvt_b200.debug_jls_path(path)
    try:
        for near in (1, 2, 3, 4):
            for k, img in enumerate(cases):
                got = imcvt_b200.encode_jls_batch(img[None], near)[0]
                want = oracle_jls(img, near)
                assert got == want, f"case {k} near {near}: {first_diff(got, want)}"
        # a batch with more planes than SMs (plane = CTA), RGB
        frames = synth.batch("texture", 70, 20, 300, rgb=True)
        got = imcvt_b200.encode_jls_batch(frames, 2)
        for i in (0, 33, 69):
            assert got[i] == oracle_jls(frames[i], 2), i
    finally:
        imcvt_b200.debug_jls_path(0)


def test_full_hd_frame_matches_oracle():
    for kind in ["natural", "texture"]:
        f = synth.frame(kind, 5, 1080, 1920)
        check(f, 0)


def test_row_tiles_by_bulk_copy_and_through_registers_give_the_same_streams(lossless_path):
    """The tile kernels stage their rows with cp.async.bulk + mbarrier when the plane's width is a multiple of 16 (and fall back
    to loads through registers otherwise); imcvtDebugJLSPath(6) forces the register path.  Same bytes either way, on both
    launch shapes (64 frames: cluster pairs; 160: one CTA per plane), also with page-locked buffers coded in place."""
    if lossless_path != 2:
        pytest.skip("sets its own paths")
    import ctypes as C
    import torch
    for n, h, w in ((3, 1080, 1920), (160, 33, 160), (5, 40, 1912)):
        frames = synth.batch("natural", n, h, w)
        imcvt_b200.debug_jls_path(0)
        a = imcvt_b200.encode_jls_batch(frames, 0)
        imcvt_b200.debug_jls_path(6)
        b = imcvt_b200.encode_jls_batch(frames, 0)
        imcvt_b200.debug_jls_path(0)
        assert a == b, (n, h, w)
        assert a[0] == oracle_jls(frames[0], 0) and a[-1] == oracle_jls(frames[-1], 0)
    # page-locked frames and slots: the bulk copies read host memory over PCIe
    frames = synth.batch("texture", 4, 270, 480)
    stride = imcvt_b200.jls_bound(0, 270, 480)
    h_in = torch.from_numpy(frames).pin_memory()
    h_out = torch.zeros(4 * stride, dtype=torch.uint8).pin_memory()
    lens = (C.c_size_t * 4)()
    rc = imcvt_b200.lib().imcvtEncodeBatchJLS(h_in.numpy().ctypes.data, 270 * 480, 4, 0, 270, 480, 0, h_out.numpy().ctypes.data, stride, C.cast(lens, C.c_void_p))
    assert rc == 0, imcvt_b200.last_error()
    for i in range(4):
        assert h_out.numpy()[i * stride:i * stride + lens[i]].tobytes() == oracle_jls(frames[i], 0), i


def test_serial_kernel_equals_pipeline_at_full_size():
    frames = synth.batch("natural", 4, 1080, 1920)
    a = imcvt_b200.encode_jls_batch(frames, 0)
    imcvt_b200.debug_jls_force_serial(True)
    try:
        b = imcvt_b200.encode_jls_batch(frames, 0)
    finally:
        imcvt_b200.debug_jls_force_serial(False)
    assert a == b


def test_write_jls_image_file_is_a_drop_in():
    img, _ = load_pnm(os.path.join(GOLDEN, "P5.pnm"))
    entry = [e for e in golden_table() if e["codec"] == "jls" and e["fixture"] == "P5" and e["setting"] == 0][0]
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "a.jls")
        assert imcvt_b200.write_jls_image_file(path, img, 0, img.shape[0], img.shape[1], 0) == 0
        assert sha256(open(path, "rb").read()) == entry["sha256"]
        # reference error behaviour: bad dimensions -> 1, nothing written (imageio_jls.c:436)
        bad = os.path.join(d, "b.jls")
        assert imcvt_b200.write_jls_image_file(bad, img, 0, 0, 10, 0) == 1
        assert not os.path.exists(bad)
        # unwritable path -> 1 (imageio_jls.c:467-472)
        assert imcvt_b200.write_jls_image_file(os.path.join(d, "no/such/dir.jls"), img, 0, img.shape[0], img.shape[1], 0) == 1


def test_planes_wider_than_the_tile_kernel_take_the_multi_kernel_path():
    rng = np.random.default_rng(4)
    img = np.repeat(rng.integers(0, 256, (6, 2100), dtype=np.uint8), 4, axis=1)[:, :8200].copy()
    check(img, 0)


def test_more_planes_than_sms_use_the_plane_queue():
    frames = synth.batch("natural", 70, 40, 96, rgb=True)                  # 210 planes
    got = imcvt_b200.encode_jls_batch(frames, 0)
    for i in (0, 33, 69):
        assert got[i] == oracle_jls(frames[i], 0), i


def test_small_output_slot_is_an_error_not_a_crash():
    img = np.random.default_rng(1).integers(0, 256, (64, 64), dtype=np.uint8)
    with pytest.raises(imcvt_b200.ImcvtError):
        imcvt_b200.encode_jls_batch(img[None], 0, out_stride=512)


@pytest.mark.parametrize("near", [0, 1, 3])
def test_plain_one_thread_scan_kernel_still_matches(near, lossless_path):
    """jls_path 4 keeps the plain one-thread-per-plane scan (the first kernel of the lossy path) as a cross-check of the
    row-pipelined one."""
    if lossless_path != 2:
        pytest.skip("independent of the lossless path")
    frames = synth.batch("texture", 3, 45, 77)
    imcvt_b200.debug_jls_path(4)
    imcvt_b200.debug_jls_force_serial(True)
    try:
        got = imcvt_b200.encode_jls_batch(frames, near)
    finally:
        imcvt_b200.debug_jls_force_serial(False)
        imcvt_b200.debug_jls_path(0)
    for i in range(3):
        assert got[i] == oracle_jls(frames[i], near), i


def test_lossy_wide_and_tall_planes_cross_segment_boundaries():
    """The row-pipelined scan cuts rows into segments of 2048 samples: runs, contexts and the bit packer carry across them."""
    rng = np.random.default_rng(5)
    wide = np.repeat(rng.integers(0, 256, (3, 300), dtype=np.uint8), 20, axis=1)[:, :5000].copy()
    wide[1, 2047:2050] += 9
    for near in (1, 2, 4):
        for img in (wide, rng.integers(0, 256, (2, 4097), dtype=np.uint8), np.full((3, 2048), 200, np.uint8)):
            got = imcvt_b200.encode_jls_batch(img[None], near)[0]
            assert got == oracle_jls(img, near), (near, img.shape)


def test_baseline_size_streams_round_trip_through_an_independent_decoder(lossless_path):
    """BASELINE.json's shapes, checked with a property that does not need the oracle at that size: the emitted stream must
    decode (tests/tools/jls_decoder.c, written from T.87) to the pixels (NEAR = 0) or to within NEAR of them (NEAR = 2):
    1920x1080 gray, 1920x1080 RGB, and one 3840x2160 RGB frame (the lossy configuration's shape)."""
    if lossless_path == 1:
        pytest.skip("the multi-kernel path is covered at small sizes; this test is about the bench paths")
    from tests._util import jls_decode
    gray = synth.batch("natural", 3, 1080, 1920)
    for near in (0, 2):
        for i, s in enumerate(imcvt_b200.encode_jls_batch(gray, near)):
            got, near_out = jls_decode(s)
            assert near_out == near
            err = np.abs(got[0].astype(np.int16) - gray[i].astype(np.int16)).max()
            assert err <= near, (near, i, int(err))
    rgb = np.stack([synth.frame(k, 7, 1080, 1920) for k in ("natural", "texture", "noise")], axis=-1)      # H x W x 3
    got, _ = jls_decode(imcvt_b200.encode_jls_batch(rgb[None], 0)[0])
    assert np.array_equal(np.moveaxis(got, 0, 2), rgb)
    if lossless_path == 2:
        big = np.stack([synth.frame("natural", 11 + c, 2160, 3840) for c in range(3)], axis=-1)
        got, near_out = jls_decode(imcvt_b200.encode_jls_batch(big[None], 2)[0])
        assert near_out == 2
        assert np.abs(np.moveaxis(got, 0, 2).astype(np.int16) - big.astype(np.int16)).max() <= 2


def test_shutdown_and_reinit_keep_working():
    """imcvtShutdown frees every device / pinned allocation and per-device kernel attribute flag; the next call starts over."""
    frames = synth.batch("natural", 9, 40, 64)          # >= 8 frames: the chunked host path with its pinned length buffer
    a = imcvt_b200.encode_jls_batch(frames, 0)
    imcvt_b200.shutdown()
    imcvt_b200.init(0)
    b = imcvt_b200.encode_jls_batch(frames, 0)
    s, _ = imcvt_b200.encode_hevc_batch(frames[:2, :32, :32], 3)
    imcvt_b200.shutdown()
    imcvt_b200.init(0)
    s2, _ = imcvt_b200.encode_hevc_batch(frames[:2, :32, :32], 3)
    assert a == b and s == s2
    assert a[0] == oracle_jls(frames[0], 0)


def test_device_calls_on_two_streams_share_the_workspace_in_order():
    """The device-buffer entry points only queue work.  Two calls on DIFFERENT streams carve the same engine workspace one
    after the other: the second stream has to wait for the first call's kernels (Workspace::wait_previous)."""
    import torch
    a = synth.batch("natural", 6, 200, 333)
    b = synth.batch("texture", 5, 120, 517)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    jobs = []
    for frames, near, k in ((a, 0, 0), (b, 2, 1), (a, 1, 1), (b, 0, 0)):
        n, h, w = frames.shape
        jobs.append(dict(frames=frames, near=near, stream=streams[k], d=torch.from_numpy(frames).cuda(),
                         out=torch.zeros((n, imcvt_b200.jls_bound(0, h, w)), dtype=torch.uint8, device="cuda"),
                         lens=torch.zeros(n, dtype=torch.int64, device="cuda")))
    torch.cuda.synchronize()
    for j in jobs:                                      # no synchronisation between the calls
        imcvt_b200.encode_jls_device(j["d"], j["near"], j["out"], j["lens"], j["stream"].cuda_stream)
    torch.cuda.synchronize()
    for j in jobs:
        for i in range(j["frames"].shape[0]):
            got = j["out"][i, : int(j["lens"][i])].cpu().numpy().tobytes()
            assert got == oracle_jls(j["frames"][i], j["near"]), (j["near"], i)


def test_host_calls_from_two_threads_are_serialised():
    import threading
    jobs = [(synth.batch("natural", 9, 96, 160), 0), (synth.batch("noise", 3, 64, 64), 3), (synth.batch("texture", 8, 70, 131), 0)]
    got = [None] * len(jobs)

    def work(k):
        got[k] = imcvt_b200.encode_jls_batch(jobs[k][0], jobs[k][1])

    threads = [threading.Thread(target=work, args=(k,)) for k in range(len(jobs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for k, (frames, near) in enumerate(jobs):
        assert got[k] is not None
        for i in range(frames.shape[0]):
            assert got[k][i] == oracle_jls(frames[i], near), (k, i)


@pytest.mark.parametrize("mode", [-1, 1, 2])
def test_page_locked_buffers_read_and_written_in_place(mode):
    """imcvtJLSHostMode: with page-locked buffers the kernels read the frames (mode 1) and write the streams (mode 2) in
    place over PCIe instead of staging them through device memory (-1, the default: when the batch runs on the cluster-pair
    kernel).  Same bytes in every mode; pageable buffers fall back to staging whatever the mode says."""
    import ctypes as C
    import torch
    L = imcvt_b200.lib()
    try:
        for frames, near in ((synth.batch("natural", 9, 120, 333), 0), (synth.batch("texture", 3, 64, 96, rgb=True), 0),
                             (synth.batch("natural", 5, 70, 131), 2), (synth.batch("noise", 2, 33, 6200), 0)):
            L.imcvtJLSHostMode(0)
            want = imcvt_b200.encode_jls_batch(frames, near)
            L.imcvtJLSHostMode(mode)
            assert imcvt_b200.encode_jls_batch(frames, near) == want            # pageable numpy buffers: staged
            n, h, w = frames.shape[:3]
            is_rgb = int(frames.ndim == 4)
            stride = imcvt_b200.jls_bound(is_rgb, h, w)
            h_in = torch.from_numpy(frames).pin_memory()
            h_out = torch.zeros(n * stride, dtype=torch.uint8).pin_memory()
            lens = (C.c_size_t * n)()
            rc = L.imcvtEncodeBatchJLS(h_in.numpy().ctypes.data, frames[0].nbytes, n, is_rgb, h, w, near,
                                       h_out.numpy().ctypes.data, stride, C.cast(lens, C.c_void_p))
            assert rc == 0, imcvt_b200.last_error()
            got = [h_out.numpy()[i * stride:i * stride + lens[i]].tobytes() for i in range(n)]
            assert got == want, (mode, frames.shape, near)
            for i in range(n):
                assert got[i] == oracle_jls(frames[i], near)
    finally:
        L.imcvtJLSHostMode(-1)
