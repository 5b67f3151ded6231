"""GPU: seeded random sweep -- random sizes (1x1, odd widths, narrower than a warp, wider than the shared-memory tile kernel
takes, taller than wide), random content (noise, smooth ramps, piecewise-constant runs, two-level, sparse spikes) and every
setting, byte-compared with the oracle through the batch C-ABI.  SURVEY section 4's "property" tier; the seeds are fixed so
that a failure names a reproducible case."""
import numpy as np
import pytest

import imcvt_b200
from tests._util import check_hevc as oracle_hevc, check_jls as oracle_jls   # the unmodified reference when present

pytestmark = pytest.mark.gpu


def random_image(rng, h, w, rgb=False):
    kind = rng.integers(0, 6)
    shape = (h, w, 3) if rgb else (h, w)
    if kind == 0:
        img = rng.integers(0, 256, shape)
    elif kind == 1:                                             # smooth ramp + small noise
        yy, xx = np.mgrid[0:h, 0:w]
        base = (xx * rng.integers(1, 5) + yy * rng.integers(1, 5)) // rng.integers(1, 4) + rng.integers(0, 200)
        img = (base[..., None] if rgb else base) + rng.integers(-2, 3, shape)
    elif kind == 2:                                             # piecewise constant: long runs, run interruptions
        img = np.repeat(np.repeat(rng.integers(0, 256, ((h + 6) // 7, (w + 12) // 13) + ((3,) if rgb else ())), 7, axis=0), 13, axis=1)[:h, :w]
    elif kind == 3:                                             # two levels
        img = np.where(rng.random(shape) < 0.5, rng.integers(0, 40), rng.integers(200, 256))
    elif kind == 4:                                             # flat with sparse spikes
        img = np.full(shape, rng.integers(0, 256))
        mask = rng.random(shape) < 0.02
        img = np.where(mask, rng.integers(0, 256, shape), img)
    else:                                                       # near-flat: differences within NEAR most of the time
        img = rng.integers(100, 104, shape) + (rng.random(shape) < 0.01) * rng.integers(0, 100, shape)
    return np.clip(img, 0, 255).astype(np.uint8)


JLS_SHAPES = [(1, 1), (1, 2), (2, 1), (1, 37), (41, 1), (3, 3), (5, 31), (7, 33), (17, 64), (33, 65), (9, 300), (64, 257), (23, 1023),
              (6, 2049), (4, 6143), (3, 6145), (2, 9000), (300, 7), (129, 129), (50, 1921)]


@pytest.mark.parametrize("seed", range(4))
def test_jls_random_sizes_and_content(seed):
    rng = np.random.default_rng(1000 + seed)
    for k, (h, w) in enumerate(JLS_SHAPES):
        rgb = bool(rng.integers(0, 2)) and h * w < 200000
        near = int(rng.integers(0, 5)) if (k + seed) % 2 else 0
        n = int(rng.integers(1, 4))
        frames = np.stack([random_image(rng, h, w, rgb) for _ in range(n)])
        got = imcvt_b200.encode_jls_batch(frames, near)
        for i in range(n):
            assert got[i] == oracle_jls(frames[i], near), f"seed {seed} shape {h}x{w} rgb={rgb} near={near} frame {i}"


HEVC_SHAPES = [(1, 1), (8, 8), (31, 33), (32, 32), (33, 31), (64, 96), (40, 100), (97, 35), (7, 130), (65, 65)]


@pytest.mark.parametrize("seed", range(3))
def test_hevc_random_sizes_and_content(seed):
    rng = np.random.default_rng(2000 + seed)
    for k, (h, w) in enumerate(HEVC_SHAPES):
        qpd6 = int(rng.integers(0, 5))
        imcvt_b200.debug_hevc_path(1 + (k + seed) % 2)          # alternate one CTA per frame / cluster pair
        try:
            n = int(rng.integers(1, 3))
            frames = np.stack([random_image(rng, h, w) for _ in range(n)])
            streams, rcon = imcvt_b200.encode_hevc_batch(frames, qpd6)
            for i in range(n):
                want, want_r = oracle_hevc(frames[i], qpd6)
                assert streams[i] == want, f"seed {seed} shape {h}x{w} qpd6={qpd6} frame {i}"
                assert np.array_equal(rcon[i], want_r), f"seed {seed} shape {h}x{w} qpd6={qpd6} frame {i} (reconstruction)"
        finally:
            imcvt_b200.debug_hevc_path(0)
