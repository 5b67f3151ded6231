"""GPU: parity of the CUDA H.265 intra path (through the C-ABI) with the oracle and the golden fixtures."""
import os
import tempfile

import numpy as np
import pytest

import imcvt_b200
from imcvt_b200 import synth
from tests._util import GOLDEN, golden_table, load_pnm, sha256
from tests._util import check_hevc as oracle_hevc    # the unmodified reference (oracle/_ref) when present, else the restatement
from tests.test_oracle_hevc import fixture_gray, hevc_edge_images

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[1, 2, 5, 10], ids=["one-cta", "cluster-pair", "one-cta-lockstep", "cluster-pair-groups"], autouse=True)
def hevc_path(request):
    """Every test runs on both launch shapes -- one CTA per frame, and a cluster of two CTAs per frame (the default for
    batches of at most SMs/2 frames) -- and on both forms of the 8x8 search (imcvtDebugHEVCPath: + 4 lock step, + 8 warp
    groups; each shape's default form first)."""
    imcvt_b200.debug_hevc_path(request.param)
    yield request.param
    imcvt_b200.debug_hevc_path(0)


@pytest.mark.parametrize("entry", [e for e in golden_table() if e["codec"] == "hevc"],
                         ids=lambda e: f"{e['fixture']}-qpd6_{e['setting']}")
def test_known_answers(entry):
    img = fixture_gray(entry["fixture"])
    s, _ = imcvt_b200.encode_hevc_batch(img[None], entry["setting"])
    assert len(s[0]) == entry["length"]
    assert sha256(s[0]) == entry["sha256"]


@pytest.mark.parametrize("qpd6", [0, 1, 2, 3, 4])
def test_edge_cases_match_oracle(qpd6):
    for name, img in hevc_edge_images():
        s, r = imcvt_b200.encode_hevc_batch(img[None], qpd6)
        s2, r2 = oracle_hevc(img, qpd6)
        assert s[0] == s2, name
        assert np.array_equal(r[0], r2), name


def test_batch_of_synthetic_frames_all_kinds():
    for kind in ["natural", "texture", "noise", "constant"]:
        frames = synth.batch(kind, 3, 72, 136)
        s, r = imcvt_b200.encode_hevc_batch(frames, 4)
        for i in range(3):
            s2, r2 = oracle_hevc(frames[i], 4)
            assert s[i] == s2, f"{kind}[{i}]"
            assert np.array_equal(r[i], r2), f"{kind}[{i}]"


def test_batch_larger_than_the_grid_uses_the_frame_queue():
    frames = synth.batch("natural", 200, 32, 64)
    s, _ = imcvt_b200.encode_hevc_batch(frames, 3, want_rcon=False)
    for i in (0, 57, 147, 148, 199):
        assert s[i] == oracle_hevc(frames[i], 3)[0], i
    assert len(set(s)) > 150


def test_hevc_image_encoder_and_file_writer_are_drop_ins():
    img, _ = load_pnm(os.path.join(GOLDEN, "P5.pnm"))
    entry = [e for e in golden_table() if e["codec"] == "hevc" and e["fixture"] == "P5" and e["setting"] == 4][0]
    s, rcon = imcvt_b200.hevc_image_encoder(img, 4)
    assert sha256(s) == entry["sha256"]
    assert rcon.shape == (288, 320)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "a.h265")
        assert imcvt_b200.write_hevc_image_file(path, img, 0, img.shape[0], img.shape[1], 4) == 0
        assert sha256(open(path, "rb").read()) == entry["sha256"]
        # RGB input: the green channel is coded (imageio_hevc.c:24-26)
        rgb, _ = load_pnm(os.path.join(GOLDEN, "P6.pnm"))
        e6 = [e for e in golden_table() if e["codec"] == "hevc" and e["fixture"] == "P6" and e["setting"] == 2][0]
        assert imcvt_b200.write_hevc_image_file(path, rgb, 1, rgb.shape[0], rgb.shape[1], 2) == 0
        assert sha256(open(path, "rb").read()) == e6["sha256"]
        assert imcvt_b200.write_hevc_image_file(os.path.join(d, "no/such/dir.h265"), img, 0, img.shape[0], img.shape[1], 4) == 1


def test_small_output_slot_is_an_error_not_a_crash():
    import ctypes as C
    img = np.random.default_rng(1).integers(0, 256, (64, 64), dtype=np.uint8)
    out = np.zeros(600, dtype=np.uint8)
    ln = (C.c_size_t * 1)()
    rc = imcvt_b200.lib().imcvtEncodeBatchHEVC(img.ctypes.data, 64 * 64, 1, 64, 64, 0, out.ctypes.data, 600, C.cast(ln, C.c_void_p), None, 0)
    assert rc == 1


def test_full_hd_streams_decode_to_the_encoders_reconstruction(hevc_path):
    """BASELINE-size frames (1920x1080) cannot be checked against the CPU oracle inside a test (45 s per frame), so an
    independent decoder stands in: FFmpeg (through OpenCV) must decode every emitted stream to exactly the reconstruction
    the encoder kept -- at qpd6 <= 1 the stream's in-loop deblocking leaves the picture unchanged.
    The first CTU row of frame 0 is also compared with the oracle bit for bit."""
    if hevc_path > 3:
        pytest.skip("the default search form of each launch shape is enough at this size")
    cv2 = pytest.importorskip("cv2")
    frames = synth.batch("natural", 2, 1080, 1920)
    streams, rcon = imcvt_b200.encode_hevc_batch(frames, 1)
    assert rcon.shape == (2, 1088, 1920)
    with tempfile.TemporaryDirectory() as d:
        for i, s in enumerate(streams):
            path = os.path.join(d, f"f{i}.h265")
            open(path, "wb").write(s)
            cap = cv2.VideoCapture(path, cv2.CAP_FFMPEG)
            cap.set(cv2.CAP_PROP_CONVERT_RGB, 0)
            ok, img = cap.read()
            assert ok, f"frame {i}: FFmpeg could not decode the stream"
            y = img.reshape(-1)[: 1088 * 1920].reshape(1088, 1920)
            assert np.array_equal(y, rcon[i]), f"frame {i}: decoded picture differs from the encoder's reconstruction"
            psnr = 10 * np.log10(255.0 ** 2 / np.mean((y[:1080].astype(np.float64) - frames[i]) ** 2))
            assert psnr > 45, psnr
    # one CTU row against the oracle: the first 32 rows of the picture are coded before anything below them is read
    s_row, r_row = imcvt_b200.encode_hevc_batch(frames[:1, :32], 1)
    want, want_r = oracle_hevc(frames[0, :32], 1)
    assert s_row[0] == want and np.array_equal(r_row[0], want_r)
    assert np.array_equal(rcon[0][:32], want_r)      # intra prediction never reads below a CTU row


def test_4k_qp_sweep_decodes_and_orders(hevc_path):
    """BASELINE configs[4] shape: one 3840x2160 gray frame through the sweep the reference can express ("QP
    22/27/32/37" runs as qpd6 1..4 = QP 10/16/22/28).  180 s per encode on the CPU oracle, so the checks are the ones that
    do not depend on size: FFmpeg decodes every stream (exactly the encoder's reconstruction at qpd6 1, within the in-loop
    deblocking's reach above that), sizes and PSNR fall with the quantiser, and the first CTU row of every stream's
    reconstruction is the oracle's bit for bit."""
    if hevc_path != 2:
        pytest.skip("one launch shape is enough for the size check (a single frame runs on the cluster pair by default)")
    cv2 = pytest.importorskip("cv2")
    frame = synth.batch("natural", 1, 2160, 3840)
    sizes, psnrs = [], []
    with tempfile.TemporaryDirectory() as d:
        for qpd6 in (1, 2, 3, 4):
            streams, rcon = imcvt_b200.encode_hevc_batch(frame, qpd6)
            assert rcon.shape == (1, 2176, 3840)
            path = os.path.join(d, f"q{qpd6}.h265")
            open(path, "wb").write(streams[0])
            cap = cv2.VideoCapture(path, cv2.CAP_FFMPEG)
            cap.set(cv2.CAP_PROP_CONVERT_RGB, 0)
            ok, img = cap.read()
            assert ok, f"qpd6 {qpd6}: FFmpeg could not decode the stream"
            y = img.reshape(-1)[: 2176 * 3840].reshape(2176, 3840)
            diff = np.abs(y.astype(np.int32) - rcon[0].astype(np.int32))
            if qpd6 == 1:
                assert diff.max() == 0
            else:
                assert diff.max() <= 6 and np.mean(diff != 0) < 0.2, (qpd6, int(diff.max()), float(np.mean(diff != 0)))
            sizes.append(len(streams[0]))
            psnrs.append(10 * np.log10(255.0 ** 2 / np.mean((rcon[0][:2160].astype(np.float64) - frame[0]) ** 2)))
            want, want_r = oracle_hevc(frame[0, :32], qpd6)
            assert np.array_equal(rcon[0][:32], want_r), qpd6
    assert sizes == sorted(sizes, reverse=True) and psnrs == sorted(psnrs, reverse=True), (sizes, psnrs)


def test_sweep_in_one_launch_equals_one_batch_per_setting():
    """imcvtEncodeBatchHEVCSweep (BASELINE configs[4]: every frame at several QPs): job i*nq + k = frame i at the k-th setting,
    all jobs in one launch; the bytes are those of one batch call per setting, i.e. the oracle's."""
    frames = synth.batch("natural", 3, 72, 136)
    qs = [1, 4, 2, 3, 0]
    got = imcvt_b200.encode_hevc_sweep(frames, qs)
    for k, q in enumerate(qs):
        per, _ = imcvt_b200.encode_hevc_batch(frames, q, want_rcon=False)
        for i in range(3):
            assert got[i][k] == per[i], (i, q)
    for i in range(3):
        assert got[i][1] == oracle_hevc(frames[i], 4)[0] and got[i][4] == oracle_hevc(frames[i], 0)[0]
    # more jobs than SMs: the job queue
    many = synth.batch("texture", 60, 32, 64)
    sw = imcvt_b200.encode_hevc_sweep(many, [4, 1, 3])
    ref = {q: imcvt_b200.encode_hevc_batch(many, q, want_rcon=False)[0] for q in (4, 1, 3)}
    for i in (0, 17, 49, 59):
        assert [sw[i][0], sw[i][1], sw[i][2]] == [ref[4][i], ref[1][i], ref[3][i]], i
    L = imcvt_b200.lib()
    import ctypes as C
    bad = (C.c_int * 2)(1, 7)
    out = np.zeros(4096, np.uint8)
    ln = (C.c_size_t * 2)()
    assert L.imcvtEncodeBatchHEVCSweep(many.ctypes.data, 32 * 64, 1, 32, 64, bad, 2, out.ctypes.data, 2048, C.cast(ln, C.c_void_p)) == 1
