"""Regenerates tests/golden/fullsize_answers.json: (length, sha256) of the UNMODIFIED reference's streams
(oracle/_ref/libimcvt_ref.so) for synthetic frames at BASELINE.json's full sizes, where the reference needs
30 s .. 5 min of CPU per frame (run in the build container, where /root/reference exists):

    python tests/golden/make_golden_fullsize.py [threads] [--all]

The frames come from imcvt_b200/synth.py (integer-only, the same bytes on every machine), so the GPU tier can
check byte-level parity at 1920x1080 and 3840x2160 against these hashes without re-running the reference; the
tests that DO re-run it on the GPU box's host cores (tests/test_gpu_zz_fullsize.py) cover the same frames.
Each entry also records the wall-clock seconds one host thread of this container needed.
"""
import json, os, sys, time
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from imcvt_b200 import synth
from tests._util import GOLDEN, ref_jls, ref_hevc, sha256

# (codec, kind, frame index, H, W, rgb, setting)
JOBS = [
    ("hevc", "natural", 0, 2160, 3840, False, 1), ("hevc", "natural", 0, 2160, 3840, False, 2),
    ("hevc", "natural", 0, 2160, 3840, False, 3), ("hevc", "natural", 0, 2160, 3840, False, 4),
    ("hevc", "texture", 0, 1080, 1920, False, 1),
    ("hevc", "natural", 0, 1080, 1920, False, 4), ("hevc", "natural", 1, 1080, 1920, False, 4),
    ("hevc", "natural", 2, 1080, 1920, False, 4), ("hevc", "natural", 3, 1080, 1920, False, 4),
    ("hevc", "texture", 0, 1080, 1920, False, 4), ("hevc", "noise", 0, 1080, 1920, False, 4),
    ("jls", "natural", 0, 2160, 3840, True, 2), ("jls", "texture", 1, 2160, 3840, True, 2),
    ("jls", "natural", 0, 1080, 1920, False, 0), ("jls", "natural", 63, 1080, 1920, False, 0),
    ("jls", "texture", 0, 1080, 1920, False, 0), ("jls", "noise", 0, 1080, 1920, False, 0),
    ("jls", "natural", 0, 1080, 1920, False, 2),
]
# the sample frames of the 256-frame 3840x2160 batches (configs[3] / configs[4]) the reference is run on in full
JOBS += [("hevc", "natural", i, 2160, 3840, False, q) for i in (85, 170, 255) for q in (1, 2, 3, 4)]
JOBS += [("jls", "natural", i, 2160, 3840, True, 2) for i in (85, 170, 255)]


def run(job):
    codec, kind, idx, h, w, rgb, setting = job
    f = synth.frame(kind, idx, h, w, rgb)
    t = time.perf_counter()
    s = ref_jls(f, setting) if codec == "jls" else ref_hevc(f, setting)[0]
    dt = time.perf_counter() - t
    print(job, len(s), f"{dt:.1f} s", flush=True)
    return dict(codec=codec, kind=kind, index=idx, height=h, width=w, rgb=rgb, setting=setting,
                length=len(s), sha256=sha256(s), ref_seconds_one_thread=round(dt, 2))


if __name__ == "__main__":
    threads = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else (os.cpu_count() or 1)
    # entries already in the table are kept (the reference and synth.py are fixed); only missing jobs run
    path = os.path.join(GOLDEN, "fullsize_answers.json")
    have = json.load(open(path)) if os.path.exists(path) and "--all" not in sys.argv else []
    key = lambda e: (e["codec"], e["kind"], e["index"], e["height"], e["width"], e["rgb"], e["setting"])
    done = {key(e): e for e in have}
    todo = [j for j in JOBS if j not in done]
    with ThreadPoolExecutor(max_workers=threads) as ex:
        for e in ex.map(run, todo):
            done[key(e)] = e
    out = [done[j] for j in JOBS]
    json.dump(out, open(os.path.join(GOLDEN, "fullsize_answers.json"), "w"), indent=1)
    print(len(out), "entries")
