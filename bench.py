#!/usr/bin/env python
"""bench.py -- the driver contract.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--codec jls|hevc]
    torchrun ... bench.py --gpus N ...          (one rank per GPU, N > 1)

Main line (BASELINE.json configs[1], the configuration the metric is quoted on): JPEG-LS NEAR=0 encode of a batch
of 64 synthetic 1920x1080 gray8 frames per GPU ("natural" content: smooth field + ~N(0,3) noise, seed 1000+frame; see
imcvt_b200/synth.py).  A step = one pass of the hot path over the batch.  Frames are independent, so N GPUs encode N
disjoint batches (weak scaling, no data-path collective); the only collective is the all-gather of per-frame stream
sizes that gives every rank the global output offsets.

Sections that ride along in the same JSON line (each skippable, see --skip):
  content   the same batch with S2 "texture" (second headline), S3 "noise" and S4 "constant" (bounds) content
  hevc      configs[2]: H.265 intra, "QP 27" -> qpd6=4 = QP 28, 64 x 1080p per GPU, + its own content bounds
  c4        configs[3]: JPEG-LS NEAR=2, 256 x 3840x2160 RGB24, the fixed batch split over 8 GPUs (shard.frame_block):
            every rank codes the 32-frame block the 8-way split gives it (ranks r < N run blocks 0..N-1)
  c5        configs[4]: H.265 quantiser sweep qpd6 1..4 ("QP 22/27/32/37"), 256 x 3840x2160 gray8, same split,
            one imcvtEncodeBatchHEVCSweep launch per rank (128 jobs)
Every section byte-compares the streams the GPU produced in its timed host-buffer step with the reference's own CPU
implementation on the same frames ("parity"); a mismatch makes the run exit non-zero.

Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 1080, 1920
FRAMES_PER_GPU = 64
HEVC_FRAMES_PER_GPU = 64
KIND = "natural"
HEVC_QPD6 = 4
# the two 8-GPU configurations: a FIXED batch of 256 frames, split into 8 contiguous blocks
BIG_H, BIG_W, BIG_TOTAL, BIG_SPLIT = 2160, 3840, 256, 8
C4_NEAR = 2
C5_QLIST = [1, 2, 3, 4]
C5_SAMPLE_FRAMES = (0, 85, 170, 255)              # the frames the reference is run on in full
MACS_PER_CTU = 12615680                           # dense-contraction count of the reference per 32x32 CTU
INT32_MAC_PEAK = 148 * 64 * 1.965e9               # nominal: SMs x INT32 lanes x max SM clock (not in MEASURED_PEAKS.json)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--codec", default="jls", choices=["jls", "hevc"])
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU (default: the named config's 64)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hevc", action="store_true", help="same as --skip hevc")
    ap.add_argument("--skip", default="", help="comma list of sections to leave out: content,hevc,c4,c5,parity,parity4k")
    ap.add_argument("--big-frames", type=int, default=0, help="c4/c5: frames per GPU (default 32 = the 8-way split of 256)")
    a = ap.parse_args()
    a.skip = set(s for s in a.skip.split(",") if s)
    if a.no_hevc:
        a.skip.add("hevc")
    return a


def workload_name(codec, n, kind=KIND):
    """The same string in both arms (the driver compares them)."""
    if codec == "jls":
        return f"JPEG-LS near=0, batch {n} synthetic ({kind}) 1920x1080 gray8 frames per GPU, bit-exact .jls"
    return (f"H.265 intra qpd6={HEVC_QPD6} (QP 28; 'QP 27' is not representable), batch {n} synthetic ({kind}) "
            f"1920x1080 gray8 frames per GPU")


def metric_name(codec):
    return f"Mpixels/s encode ({'JLS near=0' if codec == 'jls' else 'H.265-I qpd6=4 (QP 28)'})"


# ------------------------------------------------------------------------------------------------
# the reference's own CPU implementation (oracle/_ref, else the oracle port), all host threads
# ------------------------------------------------------------------------------------------------
def cpu_kind():
    from tests import _util
    return "reference" if _util.ref_lib() is not None else "port"


def cpu_jls(frame, near):
    """Only bench.py's cpu_baseline / parity / --impl reference legs execute anything under oracle/."""
    from tests import _util
    return _util.ref_jls(frame, near) if _util.ref_lib() is not None else _util.oracle_jls(frame, near)


def cpu_hevc(frame, qpd6):
    from tests import _util
    return (_util.ref_hevc(frame, qpd6) if _util.ref_lib() is not None else _util.oracle_hevc(frame, qpd6))[0]


def cpu_threads():
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return max(1, n)


def cpu_map(fn, jobs, threads):
    """Runs fn over jobs on `threads` host threads (ctypes releases the GIL).  Returns (wall seconds, results, per-job seconds)."""
    def timed(j):
        t = time.perf_counter()
        r = fn(j)
        return r, time.perf_counter() - t
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
        res = list(ex.map(timed, jobs))
    return time.perf_counter() - t0, [r for r, _ in res], [s for _, s in res]


def hevc_cpu_sample(frames):
    """--impl reference --codec hevc: HEVC on the CPU costs 25-45 core-seconds per 1080p frame, so that arm encodes a
    crop: the top-left 256x128 window (32 CTUs) of each sampled frame.  Per-CTU work is frame-size independent."""
    return [np.ascontiguousarray(f[:128, :256]) for f in frames]


def run_reference(args, rank, world):
    if rank != 0:
        return
    from imcvt_b200 import synth
    kind = cpu_kind()
    threads = cpu_threads()
    if args.codec == "jls":
        n = args.frames or FRAMES_PER_GPU
        frames = list(synth.batch(KIND, n, H, W))
        enc = lambda f: cpu_jls(f, 0)
        sample = f"all {n} frames of the step, {threads} threads, one frame per task"
        pix = n * H * W
    else:
        n = args.frames or HEVC_FRAMES_PER_GPU
        k = min(threads, 16)
        frames = hevc_cpu_sample(list(synth.batch(KIND, k, H, W)))
        enc = lambda f: cpu_hevc(f, HEVC_QPD6)
        sample = f"{k} crops of 256x128 (32 CTUs each) from frames 0..{k-1}, {threads} threads; extrapolated per pixel"
        pix = k * 128 * 256
    for _ in range(args.warmup):
        cpu_map(enc, frames[:max(1, min(len(frames), threads))], threads)
    t = 0.0
    for _ in range(args.steps):
        dt, _, _ = cpu_map(enc, frames, threads)
        t += dt
    v = pix * args.steps / t / 1e6
    line = {
        "impl": "reference", "metric": metric_name(args.codec),
        "value": v, "unit": "Mpixels/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(args.codec, n), "frames_per_gpu": n, "height": H, "width": W, "sample": sample},
        "cpu_baseline": {"value": v, "unit": "Mpixels/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clocks and throttle reasons DURING the timed region: an NVML polling thread (5 ms period; the timed
    region of the JPEG-LS config is only tens of milliseconds), falling back to `nvidia-smi -lms` when NVML is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self.thread = self.stop_flag = None
        self.f = self.p = None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            # CUDA_VISIBLE_DEVICES may renumber devices: match by PCI bus id of the CUDA device when torch is there
            try:
                import torch
                bus = torch.cuda.get_device_properties(index).pci_bus_id
                dom = torch.cuda.get_device_properties(index).pci_domain_id
                dev = torch.cuda.get_device_properties(index).pci_device_id
                self.h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{dom:08X}:{bus:02X}:{dev:02X}.0")
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _poll(self):
        nv = self.nvml
        names = [("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap)]
        while not self.stop_flag.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for nme, bit in names:
                    if r & bit:
                        self.reasons.add(nme)
            except Exception:
                pass
            self.stop_flag.wait(0.005)

    def start(self):
        if self.nvml is not None:
            import threading
            self.stop_flag = threading.Event()
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            if not self.samples:
                return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"]}
            return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                    "samples": len(self.samples), "source": "nvml, 5 ms period"}
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for k, nme in enumerate(names):
                if r[5 + k].strip().lower() == "active":
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def traffic_of(kernel):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture of this workload (profiles/traffic_*.json,
    written by tools/ncu_traffic.py); labelled with the capture it comes from.  None when there is no capture."""
    tr = os.path.join(ROOT, "profiles", f"traffic_{kernel}.json")
    if os.path.exists(tr):
        try:
            d = json.load(open(tr))
            return d.get("dram_bytes_per_launch"), d.get("source", os.path.relpath(tr, ROOT))
        except Exception:
            pass
    return None, None


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local_rank):
    """One process per GPU: run on the CPUs NVML names as local to this GPU, so that the page-locked frame / stream buffers
    (first touched by this process) sit on the NUMA node whose PCIe root the GPU hangs on.  The host-buffer leg moves
    ~200 MB per step and GPU; with eight ranks floating over two sockets half of that crossed the socket link."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local_rank)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{pr.pci_domain_id:08X}:{pr.pci_bus_id:02X}:{pr.pci_device_id:02X}.0")
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [i for i in range(ncpu) if (words[i // 64] >> (i % 64)) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception as e:                      # no NVML / no affinity information: stay where the launcher put us
        print(f"bench: NUMA binding skipped ({e})", file=sys.stderr)
    return 0


class Ctx:
    """What every section needs: rank layout, the host threads this rank may use for the CPU legs, collectives."""

    def __init__(self, args, rank, world, local_rank):
        import torch
        self.args, self.rank, self.world, self.local_rank = args, rank, world, local_rank
        self.torch = torch
        self.threads = max(1, cpu_threads() // max(1, world)) if world > 1 else cpu_threads()
        self.failed = []                          # parity failures (section names)

    def barrier(self):
        import torch.distributed as dist
        self.torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        import torch.distributed as dist
        t = self.torch.tensor(values, dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def sum_over_ranks(self, values):
        import torch.distributed as dist
        t = self.torch.tensor(values, dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(x) for x in t.tolist()]


def parity_report(ctx, name, got, want, what):
    """got / want: lists of bytes.  Returns the "parity" object (counts summed over ranks); remembers a failure."""
    bad = [i for i, (a, b) in enumerate(zip(got, want)) if a != b]
    n, nbad = ctx.sum_over_ranks([len(got), len(bad)])
    if bad and not isinstance(got[bad[0]], (bytes, bytearray)):
        print(f"bench: PARITY MISMATCH in {name}: item {bad[0]}: (length, sha256) {got[bad[0]]} (GPU) vs {want[bad[0]]} (reference)", file=sys.stderr)
    elif bad:
        i = bad[0]
        k = next((j for j in range(min(len(got[i]), len(want[i]))) if got[i][j] != want[i][j]), min(len(got[i]), len(want[i])))
        print(f"bench: PARITY MISMATCH in {name}: item {i}, first difference at byte {k} of {len(got[i])} (GPU) / {len(want[i])} (CPU)", file=sys.stderr)
    if nbad:
        ctx.failed.append(name)
    return {"frames": int(n), "equal": nbad == 0, "against": cpu_kind(), "what": what}


def host_link_probe(ctx, nin, nout, iters=8):
    """What the host side of the box gives when every rank moves its step's bytes at once: `nin` bytes page-locked host -> HBM
    and `nout` bytes HBM -> page-locked host per rank and iteration, both directions in flight on two streams, all ranks
    together (barrier on both sides, max over ranks).  Returns aggregate GB/s.  The e2e leg cannot move its bytes faster."""
    torch = ctx.torch
    h_in = torch.empty(nin, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(nout, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(nin, dtype=torch.uint8, device="cuda")
    d_out = torch.zeros(nout, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def once():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)

    once()
    ctx.barrier()
    t = time.perf_counter()
    for _ in range(iters):
        once()
    ctx.barrier()
    dt = ctx.max_over_ranks([time.perf_counter() - t])[0]
    return ctx.world * (nin + nout) * iters / dt / 1e9


def measure(ctx, codec, steps, warmup, kind=KIND, with_host=True, with_cpu=True):
    """One codec on the 64 x 1080p batch; returns the JSON line's fields on rank 0 (None elsewhere)."""
    import torch
    import torch.distributed as dist
    import imcvt_b200
    from imcvt_b200 import shard, synth
    args, rank, world = ctx.args, ctx.rank, ctx.world

    n = args.frames or (FRAMES_PER_GPU if codec == "jls" else HEVC_FRAMES_PER_GPU)
    first = rank * n                                   # disjoint frames per rank (weak scaling)
    frames = synth.batch(kind, n, H, W, first=first)   # uint8 [n,H,W]
    npix = n * H * W

    # --- device-resident buffers (value) ---
    d_frames = torch.from_numpy(frames).cuda()
    if codec == "jls":
        out_stride = 2 * H * W + 4096                  # >2x any 8-bit content's JPEG-LS size; the API fails loudly if too small
    else:
        out_stride = imcvt_b200.hevc_bound(H, W)
    d_out = torch.empty((n, out_stride), dtype=torch.uint8, device="cuda")
    d_len = torch.zeros(n, dtype=torch.int64, device="cuda")
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream

    def step_device():
        if codec == "jls":
            imcvt_b200.encode_jls_device(d_frames, 0, d_out, d_len, sptr)
        else:
            imcvt_b200.encode_hevc_device(d_frames, HEVC_QPD6, d_out, d_len, None, sptr)
        if world > 1:
            # the path's only exchange: per-frame stream sizes -> global offsets on every rank (imcvt_b200/shard.py)
            sizes = shard.gather_sizes(d_len, world * n, world)
            return shard.global_offsets(sizes)
        return d_len

    # --- warm-up ---
    for _ in range(warmup):
        step_device()
    ctx.barrier()

    # --- timed region: K steps, CUDA events on the launching stream, kernel events inside the library ---
    sampler = ClockSampler(ctx.local_rank)
    sampler.start()
    imcvt_b200.profile_enable(True)
    imcvt_b200.profile_reset()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    ctx.barrier()
    t_wall = time.perf_counter()
    for a, b in ev:
        a.record(stream)
        step_device()
        b.record(stream)
    ctx.barrier()
    t_wall = time.perf_counter() - t_wall
    launches = imcvt_b200.launch_count()
    stats = imcvt_b200.profile_stats()
    imcvt_b200.profile_enable(False)
    clocks = sampler.stop()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    total_bytes = int(d_len.sum().item())
    assert int((d_len <= 0).sum().item()) == 0, "an encode failed"
    dev_lens = d_len.cpu().numpy().copy()

    # --- e2e: host pinned buffers in, host buffers out, copies inside the timed region ---
    t_e2e, d2h, host_streams = 0.0, 0, None
    if with_host:
        h_frames = torch.from_numpy(frames).pin_memory()
        h_out = torch.empty(n * out_stride, dtype=torch.uint8).pin_memory()
        h_frames_np, h_out_np = h_frames.numpy(), h_out.numpy()
        h_lens = (C.c_size_t * n)()

        def step_host():
            L = imcvt_b200.lib()
            if codec == "jls":
                rc = L.imcvtEncodeBatchJLS(h_frames_np.ctypes.data, H * W, n, 0, H, W, 0, h_out_np.ctypes.data, out_stride,
                                           C.cast(h_lens, C.c_void_p))
            else:
                rc = L.imcvtEncodeBatchHEVC(h_frames_np.ctypes.data, H * W, n, H, W, HEVC_QPD6, h_out_np.ctypes.data, out_stride,
                                            C.cast(h_lens, C.c_void_p), None, 0)
            if rc:
                raise RuntimeError(imcvt_b200.last_error())

        step_host()
        ctx.barrier()
        t_e2e = time.perf_counter()
        for _ in range(steps):
            step_host()
        ctx.barrier()
        t_e2e = time.perf_counter() - t_e2e
        d2h = sum(int(x) for x in h_lens) + 8 * n + 4
        host_streams = [h_out_np[i * out_stride:i * out_stride + h_lens[i]].tobytes() for i in range(n)]
        assert [len(s) for s in host_streams] == [int(x) for x in dev_lens], "device-buffer and host-buffer paths disagree on stream sizes"
        # the device-resident path's bytes must be the host-buffer path's bytes
        d_host = d_out.cpu().numpy()
        for i in (0, n // 2, n - 1):
            assert d_host[i, :dev_lens[i]].tobytes() == host_streams[i], f"frame {i}: device-buffer and host-buffer paths differ"
        del h_frames, h_out
    link_gbs = None
    if with_host and codec == "jls" and kind == KIND:
        link_gbs = host_link_probe(ctx, npix, max(1, total_bytes))

    # --- max over ranks ---
    dev_ms, e2e_ms, wall_ms = ctx.max_over_ranks([dev_ms, t_e2e * 1e3, t_wall * 1e3])

    # --- CPU legs: the reference's own CPU path on this box's host cores.  Every rank byte-compares what its GPU produced
    #     in the host-buffer step above; rank 0 at N=1 also times it (cpu_baseline) ---
    cpu = parity = None
    if with_cpu and with_host and "parity" not in args.skip:
        threads = ctx.threads
        if codec == "jls":
            enc = lambda f: cpu_jls(f, 0)
            idx = list(range(n))
            sample = f"all {n} frames of the step on {threads} threads (second of two passes)"
            cpu_map(enc, [frames[i] for i in idx[:threads]], threads)           # warm the code, tables and page cache
        else:
            enc = lambda f: cpu_hevc(f, HEVC_QPD6)
            k = max(4, min(threads, 16))
            idx = list(range(min(k, n)))
            sample = f"frames 0..{len(idx)-1} of the step in full (1920x1080, 2040 CTUs each) on {min(threads, len(idx))} threads"
        dt, want, _ = cpu_map(enc, [frames[i] for i in idx], threads)
        parity = parity_report(ctx, f"{codec}/{kind}", [host_streams[i] for i in idx], want,
                               f"every byte of the streams of {'all' if len(idx) == n else 'the first ' + str(len(idx))} frames of each rank's timed batch")
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            cpu = {"value": len(idx) * H * W / dt / 1e6, "unit": "Mpixels/s", "cores": min(threads, len(idx)), "kind": cpu_kind(), "sample": sample}

    if rank != 0:
        return None
    K = steps
    value = world * npix * K / (dev_ms / 1e3) / 1e6
    peak, peak_src = measured_peak()
    # dominant kernel = the one with the largest share of the timed region (per-kernel CUDA events in the library)
    dom = max(stats, key=lambda k: stats[k][1]) if stats else None
    if dom is not None and stats[dom][0] > 0:
        cnt, ms = stats[dom]
        alg_bytes = npix + total_bytes                     # H*W*C in + stream out, per launch (whole batch)
        achieved = alg_bytes / (ms / cnt / 1e3) / 1e9
        traffic, traffic_src = traffic_of(dom) if kind == KIND else (None, None)
        roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": ms / cnt,
                "kernel_ms_per_step": {k: v[1] / K for k, v in sorted(stats.items())}}
        if codec == "hevc":
            ctus = ((H + 31) // 32) * ((W + 31) // 32)
            macs = MACS_PER_CTU * ctus * n
            rate = macs / (ms / cnt / 1e3)
            roof["int_mac"] = {"macs_per_launch": macs, "achieved_tmacs": rate / 1e12, "peak_tmacs_nominal": INT32_MAC_PEAK / 1e12,
                               "frac": rate / INT32_MAC_PEAK,
                               "note": "dense-contraction count of the reference (12 615 680 MAC per CTU, SURVEY 8d); the kernel uses partial butterflies"}
    else:
        roof = {"bound": "hbm", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None}
    line = {
        "metric": metric_name(codec),
        "value": value, "unit": "Mpixels/s", "n_gpus": world, "steps": K, "warmup": warmup,
        "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {
            "workload": workload_name(codec, n, kind),
            "frames_per_gpu": n, "height": H, "width": W, "parallelism": f"frame-sharded x{world}",
            "l2": "the inputs of a step (132.7 MB per GPU) exceed the 126 MB L2 and are read from HBM again every step; no explicit flush",
            "bytes_out_per_step_rank0": total_bytes,
        },
        "gpu_launches": launches,
        "wall_ms_per_step": wall_ms / K,
        "clocks": clocks,
        "roofline": roof,
    }
    if with_host:
        line["e2e"] = {"value": world * npix * K / (e2e_ms / 1e3) / 1e6, "unit": "Mpixels/s",
                       "h2d_bytes_per_step": npix, "d2h_bytes_per_step": d2h,
                       "transfer": ("page-locked host frames and slots through imcvtEncodeBatchJLS: the kernel reads the frames and writes "
                                    "the streams in place over PCIe every step (no staging copy in HBM)") if codec == "jls" else
                                   "page-locked host frames uploaded, streams downloaded through imcvtEncodeBatchHEVC every step"}
    if with_host and link_gbs is not None:
        line["e2e"]["host_link"] = {
            "copy_ceiling_gbs": link_gbs, "achieved_gbs": world * (npix + d2h) * K / (e2e_ms / 1e3) / 1e9,
            "what": ("copy_ceiling = all ranks copying this step's bytes (frames in, streams out) between page-locked host memory and "
                     "HBM at once, both directions in flight, measured in this run; achieved = the same bytes over the e2e step time")}
    if parity is not None:
        line["parity"] = parity
    if cpu is not None:
        line["cpu_baseline"] = cpu
    return line


def content_bounds(ctx, codec, steps):
    """S2 texture (second headline), S3 noise / S4 constant (bounds), device-resident; rank 0 gets {kind: {...}}."""
    out = {}
    for kind in ("texture", "noise", "constant"):
        sec = measure(ctx, codec, steps, 1 if codec == "hevc" else 3, kind=kind, with_host=False, with_cpu=False)
        if sec is not None:
            out[kind] = {"value": sec["value"], "unit": sec["unit"], "ms_per_step": sec["ms_per_step"], "steps": sec["steps"],
                         "bytes_out_per_step_rank0": sec["config"]["bytes_out_per_step_rank0"],
                         "roofline_frac": sec["roofline"].get("frac"), "kernel": sec["roofline"].get("kernel")}
            if codec == "hevc" and "int_mac" in sec["roofline"]:
                out[kind]["int_mac_tmacs"] = sec["roofline"]["int_mac"]["achieved_tmacs"]
    return out


def big_frames(kind, first, count, rgb, threads):
    """The rank's block of the 256-frame 3840x2160 batch (synthetic frame index = global frame index)."""
    from imcvt_b200 import synth
    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
        fr = list(ex.map(lambda i: synth.frame(kind, first + i, BIG_H, BIG_W, rgb), range(count)))
    return np.stack(fr, axis=0)


def measure_c4(ctx):
    """configs[3]: JPEG-LS NEAR=2, 256 x 3840x2160 RGB24 (three scans per frame), frame-sharded over 8 GPUs."""
    import torch
    import imcvt_b200
    from imcvt_b200 import shard
    args, rank, world = ctx.args, ctx.rank, ctx.world
    per = args.big_frames or shard.frame_block(0, BIG_SPLIT, BIG_TOTAL)[1]
    first = rank * per
    frames = big_frames(KIND, first, per, True, ctx.threads)                   # [per, H, W, 3]
    fbytes = BIG_H * BIG_W * 3
    out_stride = fbytes                                                         # NEAR=2 streams are ~1.5 bits per sample; too small fails loudly
    d_frames = torch.from_numpy(frames).cuda()
    d_out = torch.empty((per, out_stride), dtype=torch.uint8, device="cuda")
    d_len = torch.zeros(per, dtype=torch.int64, device="cuda")
    stream = torch.cuda.current_stream()

    def step_device():
        imcvt_b200.encode_jls_device(d_frames, C4_NEAR, d_out, d_len, stream.cuda_stream)
        if world > 1:
            return shard.global_offsets(shard.gather_sizes(d_len, world * per, world))    # the size all-gather, inside the step
        return d_len

    step_device()
    ctx.barrier()
    imcvt_b200.profile_enable(True)
    imcvt_b200.profile_reset()
    sampler = ClockSampler(ctx.local_rank)
    sampler.start()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    a.record(stream)
    step_device()
    b.record(stream)
    ctx.barrier()
    clocks = sampler.stop()
    dev_ms = a.elapsed_time(b)
    stats = imcvt_b200.profile_stats()
    launches = imcvt_b200.launch_count()
    imcvt_b200.profile_enable(False)
    assert int((d_len <= 0).sum().item()) == 0, "an encode failed"
    total_bytes = int(d_len.sum().item())
    del d_out

    h_frames = torch.from_numpy(frames).pin_memory()
    h_out = torch.empty(per * out_stride, dtype=torch.uint8).pin_memory()
    h_lens = (C.c_size_t * per)()
    hf, ho = h_frames.numpy(), h_out.numpy()

    def step_host():
        rc = imcvt_b200.lib().imcvtEncodeBatchJLS(hf.ctypes.data, fbytes, per, 1, BIG_H, BIG_W, C4_NEAR, ho.ctypes.data, out_stride, C.cast(h_lens, C.c_void_p))
        if rc:
            raise RuntimeError(imcvt_b200.last_error())

    step_host()                                    # warm-up: the staging arena and the chunk workspaces (several GB) are allocated on first use
    ctx.barrier()
    t = time.perf_counter()
    step_host()
    ctx.barrier()
    e2e_ms = (time.perf_counter() - t) * 1e3
    streams = [ho[i * out_stride:i * out_stride + h_lens[i]].tobytes() for i in range(per)]
    dev_ms, e2e_ms = ctx.max_over_ranks([dev_ms, e2e_ms])

    cpu = parity = None
    if "parity" not in args.skip:
        dt, want, _ = cpu_map(lambda f: cpu_jls(f, C4_NEAR), list(frames), ctx.threads)
        parity = parity_report(ctx, "c4", streams, want, "every byte of every frame's stream of each rank's block")
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            cpu = {"value": per * BIG_H * BIG_W / dt / 1e6, "unit": "Mpixels/s", "cores": min(ctx.threads, per), "kind": cpu_kind(),
                   "sample": f"all {per} frames of rank 0's block on {min(ctx.threads, per)} threads"}
    if rank != 0:
        return None
    pix = world * per * BIG_H * BIG_W
    peak, peak_src = measured_peak()
    dom = max(stats, key=lambda k: stats[k][1]) if stats else None
    roof = None
    if dom:
        cnt, ms = stats[dom]
        alg = per * fbytes + total_bytes
        # the committed capture is of exactly this launch (configs[3]'s per-GPU share, default frame count and content)
        traffic, traffic_src = traffic_of(dom) if not args.big_frames else (None, None)
        roof = {"bound": "hbm", "kernel": dom, "achieved": alg / (ms / cnt / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": alg / (ms / cnt / 1e3) / 1e9 / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg, "avg_launch_ms": ms / cnt,
                "kernel_ms_per_step": {k: v[1] for k, v in sorted(stats.items())}}
    sec = {"metric": f"Mpixels/s encode (JLS near={C4_NEAR}, RGB24: 3 samples per pixel)",
           "value": pix / (dev_ms / 1e3) / 1e6, "unit": "Mpixels/s", "msamples_per_s": 3 * pix / (dev_ms / 1e3) / 1e6,
           "steps": 1, "warmup": 1, "ms_per_step": dev_ms, "scaling": "strong (fixed 256-frame batch, 8-way split)",
           "config": {"workload": f"JPEG-LS near={C4_NEAR} lossy, batch {BIG_TOTAL} synthetic ({KIND}) 3840x2160 RGB24 (per-plane), frame-sharded over {BIG_SPLIT} GPUs",
                      "frames_per_gpu": per, "blocks_run": world, "of_blocks": BIG_SPLIT if not args.big_frames else None,
                      "share": (f"{world} of the {BIG_SPLIT} blocks of {per} frames ran (ranks 0..{world-1}); value counts the frames that ran"
                                if world < BIG_SPLIT else "the whole 256-frame batch"),
                      "planes_per_gpu": 3 * per, "bytes_out_per_step_rank0": total_bytes},
           "e2e": {"value": pix / (e2e_ms / 1e3) / 1e6, "unit": "Mpixels/s", "h2d_bytes_per_step": per * fbytes,
                   "d2h_bytes_per_step": sum(int(x) for x in h_lens) + 8 * per + 4,
                   "transfer": "page-locked host frames uploaded in chunks overlapped with the kernels, streams downloaded, through imcvtEncodeBatchJLS"},
           "gpu_launches": launches, "clocks": clocks, "roofline": roof}
    if parity is not None:
        sec["parity"] = parity
    if cpu is not None:
        sec["cpu_baseline"] = cpu
    return sec


def measure_c5(ctx):
    """configs[4]: H.265 quantiser sweep, 256 x 3840x2160 gray8, frame-sharded over 8 GPUs; one sweep launch per rank."""
    import torch
    import imcvt_b200
    from imcvt_b200 import shard
    args, rank, world = ctx.args, ctx.rank, ctx.world
    per = args.big_frames or shard.frame_block(0, BIG_SPLIT, BIG_TOTAL)[1]
    first = rank * per
    nq = len(C5_QLIST)
    jobs = per * nq
    frames = big_frames(KIND, first, per, False, ctx.threads)                  # [per, H, W]
    fbytes = BIG_H * BIG_W
    out_stride = fbytes + 65536                                                # a stream longer than its picture fails loudly
    h_frames = torch.from_numpy(frames).pin_memory()
    h_out = torch.empty(jobs * out_stride, dtype=torch.uint8).pin_memory()
    h_lens = (C.c_size_t * jobs)()
    hf, ho = h_frames.numpy(), h_out.numpy()
    q = (C.c_int * nq)(*C5_QLIST)
    live = None
    mine = [f - first for f in C5_SAMPLE_FRAMES if first <= f < first + per]
    if "parity" not in args.skip:
        fut = job = None
        if mine and "parity4k" not in args.skip:
            job = (mine[0], nq - 1)

            def run_live(fr=frames[mine[0]].copy(), qd=C5_QLIST[-1]):
                t0 = time.perf_counter()
                sb = cpu_hevc(fr, qd)
                return sb, time.perf_counter() - t0
            fut = ThreadPoolExecutor(max_workers=1).submit(run_live)
        live = (job, fut, time.perf_counter())
    imcvt_b200.profile_enable(True)
    imcvt_b200.profile_reset()
    sampler = ClockSampler(ctx.local_rank)
    sampler.start()
    ctx.barrier()
    t = time.perf_counter()
    rc = imcvt_b200.lib().imcvtEncodeBatchHEVCSweep(hf.ctypes.data, fbytes, per, BIG_H, BIG_W, q, nq, ho.ctypes.data, out_stride, C.cast(h_lens, C.c_void_p))
    if rc:
        raise RuntimeError(imcvt_b200.last_error())
    if world > 1:
        lens_t = torch.tensor([int(x) for x in h_lens], dtype=torch.int64, device="cuda")
        shard.global_offsets(shard.gather_sizes(lens_t, world * jobs, world))  # the size all-gather, inside the step
    ctx.barrier()
    e2e_ms = (time.perf_counter() - t) * 1e3
    clocks = sampler.stop()
    stats = imcvt_b200.profile_stats()
    launches = imcvt_b200.launch_count()
    imcvt_b200.profile_enable(False)
    kern = [k for k in stats if k.startswith("hevc_")]
    dev_ms = sum(stats[k][1] for k in kern)                                    # device time of the sweep kernel (frames resident in HBM)
    total_bytes = sum(int(x) for x in h_lens)
    dev_ms, e2e_ms = ctx.max_over_ranks([dev_ms, e2e_ms])

    # parity: BASELINE.md 3.5 -- the sample frames 0/85/170/255, every byte, at each setting.  The reference needs 95 s .. 5 min
    # of one host thread per (frame, setting) at this size, so (1) every owned sample frame x setting is checked against the
    # (length, sha256) of the UNMODIFIED reference's stream committed in tests/golden/fullsize_answers.json (generated by
    # tests/golden/make_golden_fullsize.py; costs no host time here), and (2) the reference itself is run live on this box for
    # the rank's first sample frame at the last setting -- started on a host thread BEFORE the sweep so that it overlaps the
    # GPU's work -- and compared byte for byte; its measured thread-seconds also give the cpu_baseline.
    cpu = parity = None
    if live is not None:
        import hashlib
        from tests import _util
        table = {(e["index"], e["setting"]): e for e in _util.fullsize_table()
                 if e["codec"] == "hevc" and e["kind"] == KIND and (e["height"], e["width"]) == (BIG_H, BIG_W)}
        got, want = [], []
        for i in mine:
            for k in range(nq):
                e = table.get((first + i, C5_QLIST[k]))
                if e is None:
                    continue
                b = ho[(i * nq + k) * out_stride:(i * nq + k) * out_stride + h_lens[i * nq + k]].tobytes()
                got.append((len(b), hashlib.sha256(b).hexdigest()))
                want.append((e["length"], e["sha256"]))
        live_job, fut, t_live0 = live
        if fut is not None:
            ref_stream, secs = fut.result()
            i, k = live_job
            got.append(ho[(i * nq + k) * out_stride:(i * nq + k) * out_stride + h_lens[i * nq + k]].tobytes())
            want.append(ref_stream)
        parity = parity_report(ctx, "c5", got, want,
                               f"frames {[f for f in C5_SAMPLE_FRAMES if f < world * per]} (of the ranks that own them) at each of the {nq} settings: length + sha256 of the "
                               f"unmodified reference's stream (tests/golden/fullsize_answers.json); the first owned sample frame at qpd6 {C5_QLIST[-1]} "
                               "also byte for byte against the reference run live on this box")
        if rank == 0 and world == 1 and not args.no_cpu_baseline and fut is not None:
            # thread-seconds of the other settings relative to the live one, from the golden table's own timing column (build container)
            rel = sum(table[(first + i, q)]["ref_seconds_one_thread"] for q in C5_QLIST if (first + i, q) in table) / table[(first + i, C5_QLIST[k])]["ref_seconds_one_thread"] \
                if all((first + i, q) in table for q in C5_QLIST) else float(nq)
            per_thread = nq * BIG_H * BIG_W / (secs * rel) / 1e6
            cpu = {"value": per_thread * ctx.threads, "unit": "Mpixels/s", "cores": ctx.threads, "kind": cpu_kind(),
                   "sample": (f"frame {first + i} in full at qpd6 {C5_QLIST[k]} on one host thread of this box ({secs:.0f} s, run while the GPU sweep ran); the other settings scaled "
                              f"by their measured cost relative to it ({rel:.2f}x for all {nq}, tests/golden/fullsize_answers.json); per-thread rate x {ctx.threads} host threads "
                              "(frames are independent; the survey measured linear scaling)")}
    if rank != 0:
        return None
    pix = world * jobs * BIG_H * BIG_W                                         # every (frame, setting) encode counts its pixels
    ctus = ((BIG_H + 31) // 32) * ((BIG_W + 31) // 32)
    macs = MACS_PER_CTU * ctus * jobs
    peak, peak_src = measured_peak()
    alg = per * fbytes + total_bytes
    sec = {"metric": "Mpixels/s encode (H.265-I sweep qpd6 1..4 = QP 10/16/22/28; 'QP 22/27/32/37' is not representable)",
           "value": pix / (dev_ms / 1e3) / 1e6, "unit": "Mpixels/s", "steps": 1, "warmup": 0, "ms_per_step": dev_ms,
           "scaling": "strong (fixed 256-frame batch, 8-way split)",
           "config": {"workload": f"H.265 intra QP sweep qpd6 {C5_QLIST}, batch {BIG_TOTAL} synthetic ({KIND}) 3840x2160 gray8, frame-sharded over {BIG_SPLIT} GPUs",
                      "frames_per_gpu": per, "jobs_per_gpu": jobs, "blocks_run": world,
                      "share": (f"{world} of the {BIG_SPLIT} blocks of {per} frames ran (ranks 0..{world-1}); value counts the encodes that ran"
                                if world < BIG_SPLIT else "the whole 256-frame batch at 4 settings = 1024 encodes"),
                      "bytes_out_per_step_rank0": total_bytes,
                      "timing": "value = the sweep kernel's CUDA-event time (frames resident in HBM); e2e = the imcvtEncodeBatchHEVCSweep call, upload and download included"},
           "e2e": {"value": pix / (e2e_ms / 1e3) / 1e6, "unit": "Mpixels/s", "h2d_bytes_per_step": per * fbytes, "d2h_bytes_per_step": total_bytes + 8 * jobs + 4},
           "gpu_launches": launches, "clocks": clocks,
           "roofline": {"bound": "hbm", "kernel": kern[0] if kern else None, "achieved": alg / (dev_ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": alg / (dev_ms / 1e3) / 1e9 / peak, "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg,
                        "int_mac": {"macs_per_launch": macs, "achieved_tmacs": macs / (dev_ms / 1e3) / 1e12,
                                    "peak_tmacs_nominal": INT32_MAC_PEAK / 1e12, "frac": macs / (dev_ms / 1e3) / INT32_MAC_PEAK}}}
    if parity is not None:
        sec["parity"] = parity
    if cpu is not None:
        sec["cpu_baseline"] = cpu
    return sec


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import imcvt_b200

    torch.cuda.set_device(local_rank)
    if world > 1:
        bind_to_gpu_numa_node(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    imcvt_b200.init(local_rank)
    ctx = Ctx(args, rank, world, local_rank)

    line = measure(ctx, args.codec, args.steps, max(args.warmup, 3))
    if "content" not in args.skip:
        c = content_bounds(ctx, args.codec, 3 if args.codec == "jls" else 1)
        if rank == 0:
            line["content"] = c
    if args.codec == "jls":
        if "hevc" not in args.skip:
            # the metric names both encoders: the H.265 configuration (configs[2]) rides along, with fewer steps (one step is seconds)
            sec = measure(ctx, "hevc", max(3, min(3, args.steps)), 1)
            if "content" not in args.skip:
                c = content_bounds(ctx, "hevc", 1)
                if rank == 0:
                    sec["content"] = c
            if rank == 0:
                line["hevc"] = {k: sec[k] for k in ("metric", "value", "unit", "steps", "warmup", "ms_per_step", "config", "e2e", "gpu_launches",
                                                      "clocks", "roofline", "parity", "cpu_baseline", "content") if k in sec}
        if "c4" not in args.skip:
            sec = measure_c4(ctx)
            if rank == 0:
                line["c4"] = sec
        if "c5" not in args.skip:
            sec = measure_c5(ctx)
            if rank == 0:
                line["c5"] = sec
    failed = ctx.failed
    if rank == 0:
        if failed:
            line["parity_failed"] = failed
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    if failed:
        sys.exit(3)


_JSON_FD = None


def emit(line):
    """The ONE JSON line of the contract, written to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    # Native libraries print to stdout too (NCCL's "NCCL version ..." banner on the multi-GPU boxes): everything but the
    # JSON line goes to stderr, so that stdout carries exactly one line.
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.abspath(__file__)] + sys.argv[1:]
        os.dup2(_JSON_FD, 1)            # the re-executed ranks redirect for themselves
        os.execv(sys.executable, cmd)
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
