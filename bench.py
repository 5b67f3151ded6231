#!/usr/bin/env python
"""bench.py -- the driver contract.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--codec jls|hevc]
    torchrun ... bench.py --gpus N ...          (one rank per GPU, N > 1)

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): JPEG-LS NEAR=0 encode of a
batch of 64 synthetic 1920x1080 gray8 frames per GPU ("natural" content: smooth field + ~N(0,3) noise, seed
1000+frame; see imcvt_b200/synth.py).  A step = one pass of the hot path over the batch.  Frames are
independent, so N GPUs encode N disjoint batches (weak scaling, no data-path collective); the only
collective is the all-gather of per-frame stream sizes that gives every rank the global output offsets.
`--codec hevc` switches to configs[2] (H.265 intra, "QP 27" -> qpd6=4 = QP 28).

Prints ONE JSON line on rank 0.
"""
import argparse
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--codec", default="jls", choices=["jls", "hevc"])
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU (default: the named config's 64)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hevc", action="store_true", help="default run: skip the H.265 measurement that rides along with the JPEG-LS line")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# the reference's own CPU implementation (oracle/_ref, else the oracle port), all host threads
# ------------------------------------------------------------------------------------------------
def cpu_encoder(codec):
    """Returns (callable(frame) -> stream bytes, kind).  Only bench.py's cpu_baseline / --impl reference legs
    execute anything under oracle/."""
    from tests import _util
    if _util.ref_lib() is not None:
        if codec == "jls":
            return (lambda f: _util.ref_jls(f, 0)), "reference"
        return (lambda f: _util.ref_hevc(f, HEVC_QPD6)[0]), "reference"
    if codec == "jls":
        return (lambda f: _util.oracle_jls(f, 0)), "port"
    return (lambda f: _util.oracle_hevc(f, HEVC_QPD6)[0]), "port"


def cpu_threads():
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return max(1, n)


def cpu_run(enc, frames, threads):
    """Encodes every frame once on `threads` host threads (ctypes releases the GIL); returns seconds."""
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        sizes = list(ex.map(lambda f: len(enc(f)), frames))
    return time.perf_counter() - t0, sizes


def hevc_cpu_sample(frames):
    """HEVC on the CPU costs ~45 core-seconds per 1080p frame, so the CPU legs encode a crop:
    the top-left 256x128 window (32 CTUs) of each sampled frame.  Per-CTU work is frame-size independent."""
    return [np.ascontiguousarray(f[:128, :256]) for f in frames]


def run_reference(args, rank, world):
    if rank != 0:
        return
    from imcvt_b200 import synth
    enc, kind = cpu_encoder(args.codec)
    threads = cpu_threads()
    if args.codec == "jls":
        n = args.frames or FRAMES_PER_GPU
        frames = list(synth.batch(KIND, n, H, W))
        sample = f"all {n} frames of the step, {threads} threads, one frame per task"
        pix = n * H * W
        workload = f"JPEG-LS near=0, batch {n} synthetic ({KIND}) 1920x1080 gray8 frames"
    else:
        n = min(threads, 16)
        frames = hevc_cpu_sample(list(synth.batch(KIND, n, H, W)))
        sample = f"{n} crops of 256x128 (32 CTUs each) from frames 0..{n-1}, {threads} threads; extrapolated per pixel"
        pix = n * 128 * 256
        workload = f"H.265 intra qpd6={HEVC_QPD6} (QP 28), synthetic ({KIND}) 1920x1080 gray8 frames"
    for _ in range(args.warmup):
        cpu_run(enc, frames[:max(1, min(len(frames), threads))], threads)
    t = 0.0
    for _ in range(args.steps):
        dt, _ = cpu_run(enc, frames, threads)
        t += dt
    v = pix * args.steps / t / 1e6
    line = {
        "impl": "reference", "metric": f"Mpixels/s encode ({'JLS near=0' if args.codec == 'jls' else 'H.265-I qpd6=4'})",
        "value": v, "unit": "Mpixels/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload, "sample": sample},
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


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import imcvt_b200
    from imcvt_b200 import synth

    torch.cuda.set_device(local_rank)
    if world > 1:
        bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    imcvt_b200.init(local_rank)

    line = measure(args, args.codec, args.steps, max(args.warmup, 3), rank, world, local_rank)
    if args.codec == "jls" and not args.no_hevc:
        # the metric names both encoders: the H.265 configuration (configs[2]) rides along, with fewer steps (one step is seconds)
        sec = measure(args, "hevc", max(1, min(2, args.steps)), 1, rank, world, local_rank)
        if rank == 0:
            line["hevc"] = {k: sec[k] for k in ("metric", "value", "unit", "steps", "warmup", "ms_per_step", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline") if k in sec}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def measure(args, codec, steps, warmup, rank, world, local_rank):
    """One codec's measurement; returns the JSON line's fields on rank 0 (None elsewhere)."""
    import torch
    import torch.distributed as dist
    import imcvt_b200
    from imcvt_b200 import shard, synth

    n = args.frames or (FRAMES_PER_GPU if codec == "jls" else HEVC_FRAMES_PER_GPU)
    first = rank * n                                   # disjoint frames per rank (weak scaling)
    frames = synth.batch(KIND, n, H, W, first=first)   # uint8 [n,H,W]
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

    # --- host-buffer path (e2e) ---
    h_frames = torch.from_numpy(frames).pin_memory()
    h_out = torch.empty(n * out_stride, dtype=torch.uint8).pin_memory()
    h_frames_np, h_out_np = h_frames.numpy(), h_out.numpy()
    import ctypes as C
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

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # --- warm-up ---
    for _ in range(warmup):
        step_device()
    barrier()

    # --- timed region: K steps, CUDA events on the launching stream, kernel events inside the library ---
    sampler = ClockSampler(local_rank)
    sampler.start()
    imcvt_b200.profile_enable(True)
    imcvt_b200.profile_reset()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    t_wall = time.perf_counter()
    for a, b in ev:
        a.record(stream)
        sizes = step_device()
        b.record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall
    launches = imcvt_b200.launch_count()
    stats = imcvt_b200.profile_stats()
    imcvt_b200.profile_enable(False)
    clocks = sampler.stop()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    total_bytes = int(d_len.sum().item())
    assert int((d_len <= 0).sum().item()) == 0, "an encode failed"

    # --- e2e: host pinned buffers in, host buffers out, copies inside the timed region ---
    step_host()
    barrier()
    t_e2e = time.perf_counter()
    for _ in range(steps):
        step_host()
    barrier()
    t_e2e = time.perf_counter() - t_e2e
    d2h = sum(int(x) for x in h_lens) + 8 * n + 4

    # --- max over ranks ---
    t = torch.tensor([dev_ms, t_e2e * 1e3, t_wall * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, wall_ms = (float(x) for x in t.tolist())

    # --- CPU baseline (rank 0, N=1 only): the reference's own CPU path on this box's host cores ---
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        enc, kind = cpu_encoder(codec)
        threads = cpu_threads()
        if codec == "jls":
            sample_frames = list(frames)
            sample = f"all {n} frames of the step on {threads} threads"
            spix = npix
        else:
            k = min(threads, 16)
            sample_frames = hevc_cpu_sample(list(frames[:k]))
            sample = f"{k} crops of 256x128 (32 CTUs each) from frames 0..{k-1} on {threads} threads; per-pixel rate"
            spix = k * 128 * 256
        dt, _ = cpu_run(enc, sample_frames, threads)
        cpu = {"value": spix / dt / 1e6, "unit": "Mpixels/s", "cores": threads, "kind": kind, "sample": sample}

    if rank == 0:
        K = steps
        value = world * npix * K / (dev_ms / 1e3) / 1e6
        peak, peak_src = measured_peak()
        # dominant kernel = the one with the largest share of the timed region (per-kernel CUDA events in the library)
        dom = max(stats, key=lambda k: stats[k][1]) if stats else None
        if dom is not None and stats[dom][0] > 0:
            cnt, ms = stats[dom]
            alg_bytes = npix + total_bytes                     # H*W*C in + stream out, per launch (whole batch)
            achieved = alg_bytes / (ms / cnt / 1e3) / 1e9
            roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": ms / cnt,
                    "kernel_ms_per_step": {k: v[1] / K for k, v in sorted(stats.items())}}
            tr = os.path.join(ROOT, "profiles", f"traffic_{dom}.json")
            if os.path.exists(tr):
                try:
                    roof["traffic"] = json.load(open(tr)).get("dram_bytes_per_launch")
                except Exception:
                    pass
        else:
            roof = {"bound": "hbm", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None}
        line = {
            "metric": f"Mpixels/s encode ({'JLS near=0' if codec == 'jls' else 'H.265-I qpd6=4 (QP 28)'})",
            "value": value, "unit": "Mpixels/s", "n_gpus": world, "steps": K, "warmup": warmup,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {
                "workload": (f"JPEG-LS near=0, batch {n} synthetic ({KIND}) 1920x1080 gray8 frames per GPU, bit-exact .jls"
                             if codec == "jls" else
                             f"H.265 intra qpd6={HEVC_QPD6} (QP 28; 'QP 27' is not representable), batch {n} synthetic ({KIND}) 1920x1080 gray8 frames per GPU"),
                "frames_per_gpu": n, "height": H, "width": W, "parallelism": f"frame-sharded x{world}",
                "l2": "the inputs of a step (132.7 MB per GPU) exceed the 126 MB L2 and are read from HBM again every step; no explicit flush",
                "bytes_out_per_step_rank0": total_bytes,
            },
            "e2e": {"value": world * npix * K / (e2e_ms / 1e3) / 1e6, "unit": "Mpixels/s",
                    "h2d_bytes_per_step": npix, "d2h_bytes_per_step": d2h,
                    "transfer": ("page-locked host frames and slots through imcvtEncodeBatchJLS: the kernel reads the frames and writes "
                                 "the streams in place over PCIe every step (no staging copy in HBM)") if codec == "jls" else
                                "page-locked host frames uploaded, streams downloaded through imcvtEncodeBatchHEVC every step"},
            "gpu_launches": launches,
            "wall_ms_per_step": wall_ms / K,
            "clocks": clocks,
            "roofline": roof,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        return line
    return None


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
