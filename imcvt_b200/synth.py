"""Deterministic synthetic frames for the benchmark and the parity tests.

Integer-only numpy (uint64 counter hash + a parabolic sine), so every machine produces the same bytes:
seed = 1000 + frame index.  Distributions:
  S1 "natural" : smooth sinusoid field + ~N(0,3) noise          (headline workload)
  S2 "texture" : tests/golden/P5.pnm tiled with a per-frame cyclic shift
  S3 "noise"   : uniform 0..255                                  (worst case)
  S4 "constant": one grey level                                  (best case: all run mode)
"""
import os

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(z):
    """splitmix64 finaliser on a uint64 array (wrapping arithmetic)."""
    with np.errstate(over="ignore"):
        z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def _hash_field(seed, h, w, stream=0):
    with np.errstate(over="ignore"):
        idx = np.arange(h * w, dtype=np.uint64).reshape(h, w)
        key = _mix64(np.uint64(seed) * np.uint64(0x100000001B3) + np.uint64(stream))
        return _mix64(idx * np.uint64(0xD6E8FEB86659FD93) + key)


def _isin(phase):
    """phase: int64 array in 1/65536 turns -> parabolic sine scaled to [-32768, 32768]."""
    t = phase & 0xFFFF
    u = t & 0x7FFF
    y = (u * (32768 - u)) >> 13
    return np.where(t >= 32768, -y, y)


def _natural_plane(seed, h, w, px, py, ax, ay, stream):
    yy, xx = np.mgrid[0:h, 0:w].astype(np.int64)
    sx = _isin(xx * 65536 // px + seed * 977)
    cy = _isin(yy * 65536 // py + 16384 + seed * 1409)
    bits = _hash_field(seed, h, w, stream)
    noise = np.zeros((h, w), dtype=np.int64)
    for k in range(4):                                    # 4 x uniform{0..4}: mean 8, sigma ~2.8
        noise += ((bits >> np.uint64(16 * k)) & np.uint64(0xFFFF)).astype(np.int64) * 5 >> 16
    v = 128 + ((ax * sx) >> 15) + ((ay * cy) >> 15) + noise - 8
    return np.clip(v, 0, 255).astype(np.uint8)


_P5 = None


def _p5_tile():
    global _P5
    if _P5 is None:
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "P5.pnm")
        data = open(path, "rb").read()
        # raw PGM "P5\n300 263\n255\n"
        parts = data.split(b"\n", 3)
        w, h = (int(t) for t in parts[1].split())
        _P5 = np.frombuffer(parts[3], dtype=np.uint8, count=w * h).reshape(h, w)
    return _P5


def frame(kind, index, h, w, rgb=False):
    """One synthetic frame: uint8 [h,w] or [h,w,3]."""
    seed = 1000 + index
    if kind == "natural":
        trip = [(609, 383, 60, 50), (487, 521, 55, 45), (701, 311, 40, 65)]
        planes = [_natural_plane(seed, h, w, *trip[c], stream=c) for c in range(3 if rgb else 1)]
    elif kind == "texture":
        t = _p5_tile()
        th, tw = t.shape
        ys = (np.arange(h) + 13 * index) % th
        xs = (np.arange(w) + 7 * index) % tw
        base = t[np.ix_(ys, xs)]
        planes = [base if c == 0 else np.roll(base, (5 * c, 3 * c), axis=(0, 1)) for c in range(3 if rgb else 1)]
    elif kind == "noise":
        planes = [(_hash_field(seed, h, w, c) & np.uint64(0xFF)).astype(np.uint8) for c in range(3 if rgb else 1)]
    elif kind == "constant":
        planes = [np.full((h, w), (128 + 37 * c + index) & 0xFF, dtype=np.uint8) for c in range(3 if rgb else 1)]
    else:
        raise ValueError(kind)
    return np.stack(planes, axis=-1) if rgb else planes[0]


def batch(kind, n, h, w, rgb=False, first=0):
    """n synthetic frames: uint8 [n,h,w] or [n,h,w,3]."""
    return np.stack([frame(kind, first + i, h, w, rgb) for i in range(n)], axis=0)
