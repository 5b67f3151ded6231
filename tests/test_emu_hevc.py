"""CPU: the SOURCE of the CUDA HEVC frame kernel (imcvt_b200/csrc/hevc_kernel.cuh), compiled for the host and run
under the fiber SIMT emulator of tests/emu, must match the oracle bit for bit (stream and reconstruction).
This checks the kernel's logic (and the absence of divergent barriers / partial-warp shuffles, which the emulator
reports as deadlocks) without a GPU; the `-m gpu` tests check the real thing through the C-ABI."""
import numpy as np
import pytest

from tests._util import emu_hevc, oracle_hevc


def images():
    rng = np.random.default_rng(21)
    yy, xx = np.mgrid[0:64, 0:64]
    yield "smooth+noise64x64", np.clip(128 + 60 * np.sin(xx / 9.0) + 40 * np.cos(yy / 7.0) + rng.normal(0, 3, xx.shape), 0, 255).astype(np.uint8)
    yield "noise33x40", rng.integers(0, 256, (33, 40), dtype=np.uint8)
    yield "edges32x64", ((xx[:32] // 7 + yy[:32] // 5) % 2 * 230 + 10).astype(np.uint8)
    yield "1x1", np.array([[77]], dtype=np.uint8)


@pytest.mark.parametrize("lockstep", [False, True], ids=["warp-groups", "lock-step"])
@pytest.mark.parametrize("pair", [False, True], ids=["one-cta", "cluster-pair"])
@pytest.mark.parametrize("qpd6", [0, 2, 4])
def test_kernel_source_matches_oracle(qpd6, pair, lockstep):
    if lockstep and qpd6 == 2:
        pytest.skip("two settings are enough for the second search form")
    for name, img in images():
        s, r = emu_hevc(img, qpd6, pair, lockstep)
        s2, r2 = oracle_hevc(img, qpd6)
        assert s == s2, f"{name}: stream differs"
        assert np.array_equal(r, r2), f"{name}: reconstruction differs"


@pytest.mark.parametrize("pair", [False, True], ids=["one-cta", "cluster-pair"])
def test_sweep_jobs_switch_quantiser_and_context_tables(pair):
    """imcvtEncodeBatchHEVCSweep's job queue on one emulated CTA / CTA pair: consecutive jobs of different qpd6 (the CTA rebuilds
    its context-initialisation table in between), and a repeated setting."""
    from imcvt_b200 import synth
    from tests._util import emu_hevc_sweep
    img = synth.frame("natural", 3, 40, 64)
    qs = [4, 1, 1, 3]
    got = emu_hevc_sweep(img, qs, pair)
    for k, q in enumerate(qs):
        want, want_r = oracle_hevc(img, q)
        assert got[k][0] == want, (k, q)
        assert np.array_equal(got[k][1], want_r), (k, q)
