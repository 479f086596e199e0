#!/usr/bin/env python
"""Benchmark of the RGNir per-pixel analysis path (BASELINE.json metric: RGNir Mpix/s for
WB + NDVI/GNDVI/NDWI + stats [+ colormap], % of HBM peak).

    python bench.py --gpus N --steps K --warmup W                    # our CUDA path, BASELINE config 2
    python bench.py --workload c3|c4|c5 --gpus N ...                 # the batched / mosaic / survey configs
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's NumPy path on host cores

A "step" is one pass of the hot path over the workload's batch of synthetic frames:

  c2 (default)  16 distinct 4000x3000 uint8 frames per GPU (weak scaling: frames are independent units).
  c3            1,024 frames 5472x3648 uint16, round-robin over the GPUs (strong scaling), all resident in HBM,
                processed in launch groups of 8 (two-level histogram, stretch build, fused pass per group).
  c4            one 32,768^2 uint8 mosaic as 64 tiles of 4,096^2, a row band of tiles per GPU (strong scaling):
                ONE white-balance histogram over the local tiles, the 3 x 256 counters summed over the GPUs INSIDE
                the step (--exchange peer: by the LUT kernel itself over NVLink peer memory; nccl: all-reduce),
                one LUT, fused pass over the tiles, image-wide statistics by one all-gather.
  c5            100,000 frames 1280x960 uint8 over the GPUs (strong scaling) in launch groups of 256 cycling through
                a device ring of 2,048 distinct resident frames (368.6 GB do not fit in HBM; refilling the ring is
                the ingest side and outside `value`), one dataset merge at the end.

One JSON line on stdout (rank 0).  `value` is device-resident throughput (inputs in HBM when the timed region
starts, CUDA events, max over ranks); `e2e` goes through the host-array API with pinned host buffers and H2D / D2H
inside the timed region; `roofline` is the fused Pass-2 kernel against the measured HBM copy bandwidth;
`sustained` is the same step looped for >= 2 s; `cpu_baseline` is the NumPy oracle port timed on this box's host
cores on a bounded sample.  Frames come from a counter-based integer generator that torch (device) and NumPy
(host) evaluate identically, so both arms and the CPU baseline see the same pixels.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "RGNir Mpix/s (WB+NDVI/GNDVI/NDWI+stats+colormap)"
UNIT = "Mpix/s"
FALLBACK_HBM_GBS = 6650.0            # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent

# name -> frame geometry and how the work is laid out
WORKLOADS = {
    "c2": dict(tag="C2", dtype="u8", width=4000, height=3000, frames_per_gpu=16, scaling="weak", seed0=2_000_000),
    "c3": dict(tag="C3", dtype="u16", width=5472, height=3648, total_frames=1024, group=8, scaling="strong", seed0=3000),
    "c4": dict(tag="C4", dtype="u8", width=4096, height=4096, total_frames=64, scaling="strong", seed0=4_000_000),
    "c5": dict(tag="C5", dtype="u8", width=1280, height=960, total_frames=100_000, group=256, ring_groups=8,
               scaling="strong", seed0=5_000_000),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=0, help="c2: frames per rank per step (default 16)")
    ap.add_argument("--total-frames", type=int, default=0, help="c3 / c4 / c5: frames (tiles) of the whole job")
    ap.add_argument("--group", type=int, default=0, help="c3 / c5: frames per launch group")
    ap.add_argument("--ring-groups", type=int, default=0, help="c5: groups of distinct frames in the device ring")
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--dtype", default="", choices=["", "u8", "u16"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = same as --steps (capped)")
    ap.add_argument("--cpu-frames", type=int, default=3, help="frames of the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-single", action="store_true", help="skip the one-frame-per-pass leg (profiling runs)")
    ap.add_argument("--sustain-s", type=float, default=-1.0,
                    help="seconds of the sustained leg (default: 2 s for c2, off for the other workloads)")
    ap.add_argument("--chunk", type=int, default=0, help="frames per pipeline chunk in the e2e path")
    ap.add_argument("--no-parity", action="store_true", help="c4: skip the reduced-mosaic check against the oracle")
    ap.add_argument("--exchange", default="peer", choices=["nccl", "peer"],
                    help="c4: white-balance histogram exchange -- NCCL all-reduce between Pass 1 and the LUT build, or the "
                         "LUT kernel exchanging the counters itself over NVLink peer memory (falls back to nccl if "
                         "symmetric memory cannot be set up)")
    a = ap.parse_args()
    w = dict(WORKLOADS[a.workload])
    for k in ("height", "width", "dtype", "group"):
        if getattr(a, k):
            w[k] = getattr(a, k)
    if a.frames:
        w["frames_per_gpu"] = a.frames
    if a.total_frames:
        w["total_frames"] = a.total_frames
    if a.ring_groups:
        w["ring_groups"] = a.ring_groups
    a.w = w
    if a.sustain_s < 0:
        a.sustain_s = 2.0 if a.workload == "c2" else 0.0
    return a


def workload_name(a):
    w = a.w
    kind = "uint8" if w["dtype"] == "u8" else "uint16"
    std = all(w[k] == WORKLOADS[a.workload][k] for k in ("width", "height", "dtype"))
    tag = w["tag"] if std else "custom"
    what = ("white balance + NDVI/GNDVI/NDWI fp32 maps + statistics/histograms + colormap RGB")
    if a.workload == "c2":
        return f"{tag}: {w['width']}x{w['height']} {kind} RGNir frames, {what}; {w['frames_per_gpu']} distinct frames per GPU per step"
    if a.workload == "c3":
        return (f"{tag}: batch of {w['total_frames']} {w['width']}x{w['height']} {kind} RGNir frames round-robin over the GPUs, "
                f"{what}; launch groups of {w['group']}")
    if a.workload == "c4":
        return (f"{tag}: one {w['total_frames']}-tile {kind} RGNir mosaic of {w['width']}x{w['height']} tiles, row band per GPU, "
                f"image-wide white balance (histogram SUM all-reduce) + {what} + image-wide statistics")
    return (f"{tag}: survey of {w['total_frames']} {w['width']}x{w['height']} {kind} RGNir frames over the GPUs, {what}; "
            f"launch groups of {w['group']} through a device ring of {w['ring_groups'] * w['group']} distinct frames")


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def profiled_traffic(a):
    """DRAM bytes per K2 launch from a committed `ncu --set full` capture of this exact workload, with the file it
    was read from (it is NOT measured in this run); (None, None) for shapes without a capture."""
    w = a.w
    if not (a.workload == "c2" and w["dtype"] == "u8" and w["width"] == 4000 and w["height"] == 3000
            and w["frames_per_gpu"] == 16):
        return None, None
    for name in ("r02_k2_fused_index_ncu_summary.txt", "r01_k2_fused_index_ncu_summary.txt"):
        try:
            total = 0.0
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            with open(os.path.join(ROOT, "profiles", name)) as fh:
                for line in fh:
                    parts = line.split()
                    if len(parts) >= 3 and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                        total += float(parts[1]) * mult[parts[2]]
            if total:
                return total, f"profiles/{name}"
        except Exception:
            continue
    return None, None


# ------------------------------------------------------------------------------------------
# synthetic frames: a counter-based integer generator, bit-identical under NumPy and torch
# ------------------------------------------------------------------------------------------
_GEN_MEAN = (90, 110, 150)        # "vegetation-like" channels: R, G, NIR
_GEN_STD = (35, 35, 45)
_M32 = 0xFFFFFFFF


def _seed32(seed: int) -> int:
    return (int(seed) * 0x85EBCA6B + 0xC2B2AE35) & _M32


def _gen_samples(idx, seed32, chan_std, chan_base, top):
    """idx: int64 sample indices (pixel * 3 + channel), seed32 / chan_std / chan_base: int64, broadcastable.
    One 32-bit hash per sample; the sum of its four bytes (Irwin-Hall, mean 510, sd 147.8) is the deviate.
    Only integer operations that NumPy and torch define identically (wrapping multiply, shifts of
    non-negative values, floor division of non-negative values)."""
    x = (idx * 0x9E3779B1 + seed32) & _M32
    x = x ^ (x >> 16)
    x = (x * 0x7FEB352D) & _M32
    x = x ^ (x >> 15)
    x = (x * 0x846CA68B) & _M32
    x = x ^ (x >> 16)
    s = (x & 255) + ((x >> 8) & 255) + ((x >> 16) & 255) + (x >> 24)
    v = chan_base + (s * chan_std) // 148
    return v.clip(0, top)


def _gen_consts(sample_bytes):
    scale = 1 if sample_bytes == 1 else 257
    top = 255 if sample_bytes == 1 else 65535
    std = [s * scale for s in _GEN_STD]
    base = [m * scale - (510 * sd) // 148 for m, sd in zip(_GEN_MEAN, std)]
    return std, base, top


def counter_frame_np(seed: int, h: int, w: int, sample_bytes: int = 1):
    """Host form: HxWx3 uint8 / uint16 frame of generator seed ``seed``."""
    import numpy as np
    std, base, top = _gen_consts(sample_bytes)
    idx = np.arange(h * w * 3, dtype=np.int64).reshape(h * w, 3)
    v = _gen_samples(idx, np.int64(_seed32(seed)), np.array(std, np.int64), np.array(base, np.int64), top)
    return v.astype(np.uint8 if sample_bytes == 1 else np.uint16).reshape(h, w, 3)


def counter_fill_device(frames, seeds, batch_samples: int = 1 << 26):
    """Device form: fill the DeviceFrames batch ``frames`` (3 channels), frame f from generator seed ``seeds[f]``."""
    import torch
    dev = frames.data.device
    npx, sb = frames.n_pixels, frames.sample_bytes
    std, base, top = _gen_consts(sb)
    std_t = torch.tensor(std, dtype=torch.int64, device=dev)
    base_t = torch.tensor(base, dtype=torch.int64, device=dev)
    idx = torch.arange(npx * 3, dtype=torch.int64, device=dev).view(1, npx, 3)
    per = max(1, batch_samples // (npx * 3))
    for a in range(0, frames.n_frames, per):
        b = min(frames.n_frames, a + per)
        s32 = torch.tensor([_seed32(s) for s in seeds[a:b]], dtype=torch.int64, device=dev).view(-1, 1, 1)
        v = _gen_samples(idx, s32, std_t, base_t, top)
        if sb == 1:
            frames.data[a:b, :npx * 3] = v.to(torch.uint8).view(b - a, -1)
        else:
            frames.data[a:b, :npx * 6] = v.to(torch.int32).to(torch.int16).view(b - a, -1).view(torch.uint8)
        del v, s32
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)
    return frames


# ------------------------------------------------------------------------------------------
# clocks: sampled DURING the timed region through NVML
# ------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.recording = False          # the thread runs from before the barrier; samples count only inside the timed region
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            if not self.recording:
                time.sleep(0.0005)
                continue
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def snapshot(self, reset=True):
        """Clocks of the recording window so far (the thread keeps running)."""
        s = sorted(self.samples)
        out = {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
               "samples": len(s)}
        if reset:
            self.samples, self.reasons = [], set()
        return out

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()


# ------------------------------------------------------------------------------------------
# CPU side: the reference's NumPy path (its own functions from oracle/_ref when built, else the oracle port)
# ------------------------------------------------------------------------------------------
_CPU_FRAMES = None
_CPU_REF = None


def _reference_functions():
    """The reference's own fix_white_balance / calculate_index / analyze_index (oracle/_ref, made by
    oracle/build_ref.py where /root/reference exists) or None -> the oracle port."""
    try:
        from oracle._ref import process_images as R
        return R
    except Exception:
        return None


def _cpu_worker_init(seeds, h, w, sb):
    """Each worker process builds its synthetic frames ONCE, outside any timed region: worker k takes
    seeds[k % len(seeds)] -- the generator seeds of the GPU arm's rank-0 frames."""
    global _CPU_FRAMES, _CPU_REF
    import multiprocessing as mp
    ident = mp.current_process()._identity
    k = (ident[0] - 1) if ident else 0
    _CPU_FRAMES = [counter_frame_np(seeds[k % len(seeds)], h, w, sb)]
    _CPU_REF = _reference_functions()


def _cpu_frames(n):
    """The reference's NumPy path on this worker's cached frame, n times."""
    import warnings
    from oracle import oracle_np as o
    t0 = time.perf_counter()
    for _ in range(n):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            o.reference_cpu_path(_CPU_FRAMES[0], reference=_CPU_REF)
    return time.perf_counter() - t0


def cpu_baseline_single(frames_host):
    """Sequential, one core: the NumPy calls of the path are single-threaded."""
    import warnings
    from oracle import oracle_np as o
    ref = _reference_functions()
    t0 = time.perf_counter()
    for img in frames_host:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            o.reference_cpu_path(img, reference=ref)
    dt = time.perf_counter() - t0
    npx = sum(f.shape[0] * f.shape[1] for f in frames_host)
    return npx / dt / 1e6, dt, ("reference" if ref is not None else "port")


def frame_seeds(a, rank, world):
    """Generator seeds of the frames rank ``rank`` holds (the reference arm replays rank 0's)."""
    w = a.w
    if a.workload == "c2":
        F = w["frames_per_gpu"]
        return [w["seed0"] + rank * F + f for f in range(F)]
    total = w["total_frames"]
    if a.workload == "c3":
        return [w["seed0"] + g for g in range(rank, total, world)]                      # round-robin
    if a.workload == "c4":
        base, extra = divmod(total, world)                                              # contiguous row band per rank
        b0 = rank * base + min(rank, extra)
        return [w["seed0"] + t for t in range(b0, b0 + base + (1 if rank < extra else 0))]
    ring = w["ring_groups"] * w["group"]
    return [w["seed0"] + rank * ring + i for i in range(ring)]


def common_config(a, world):
    w = a.w
    cfg = {"workload": workload_name(a), "height": w["height"], "width": w["width"]}
    if a.workload == "c2":
        cfg["frames_per_gpu"] = w["frames_per_gpu"]
    else:
        cfg["total_frames"] = w["total_frames"]
        cfg["frames_per_gpu"] = (w["total_frames"] + world - 1) // world
    return cfg


def run_reference(a):
    """--impl reference: the reference's CPU implementation of the path on all host cores: independent frames, one
    worker process per core, like the reference's own per-file loop (backend-process.py:92-97).  The functions are
    the reference's own (oracle/_ref, extracted unmodified by oracle/build_ref.py) when that directory travelled
    here, else the oracle port.  Frames: the generator seeds of the GPU arm's rank-0 batch."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import multiprocessing as mp
    w = a.w
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 64))
    ctx = mp.get_context("spawn")
    sb = 1 if w["dtype"] == "u8" else 2
    # bounded sample: one frame per worker per step costs ~3 s at C2 size; for long runs the frames are cropped to
    # their first rows (same width, same per-pixel work) so that the whole run stays within minutes
    rows = w["height"] if a.steps <= 20 else max(256, w["height"] * 20 // a.steps)
    per_step = 1
    if rows * w["width"] < 2_000_000 and a.steps <= 20:                  # small frames (C5): several per worker per step
        per_step = max(1, 6_000_000 // (rows * w["width"]))
    npx = rows * w["width"]
    seeds = frame_seeds(a, 0, max(1, a.gpus))
    kind = "reference" if _reference_functions() is not None else "port"
    with ctx.Pool(workers, initializer=_cpu_worker_init, initargs=(seeds, rows, w["width"], sb)) as pool:
        for _ in range(max(1, min(a.warmup, 1))):
            pool.map(_cpu_frames, [1] * workers, chunksize=1)
        t0 = time.perf_counter()
        for _ in range(a.steps):
            pool.map(_cpu_frames, [per_step] * workers, chunksize=1)
        dt = time.perf_counter() - t0
    value = workers * per_step * a.steps * npx / dt / 1e6
    what = ("the reference's own fix_white_balance / calculate_index / analyze_index (process-images.py:424-513, "
            "extracted unmodified into oracle/_ref)" if kind == "reference"
            else "NumPy oracle port of process-images.py:424-513")
    sample = (f"{workers * per_step * a.steps} frames of {w['width']}x{rows} over {workers} processes ({what} "
              "+ np.std + np.histogram(50) + colormap gather)")
    cfg = common_config(a, max(1, a.gpus if a.gpus else world))
    cfg["sample"] = f"{workers * per_step} frames of {w['width']}x{rows} per step, {per_step} per worker process"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True,
        "scaling": w["scaling"], "vs_baseline": None, "dtype": w["dtype"], "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------
def reduced_mosaic_parity(a, eng, ld, rank, world, s):
    """"C4 on a reduced mosaic with the same tiling logic" -- a 1,024^2 mosaic as 64 tiles of
    128^2, row band per rank, through the same plan / hook / exchange code as the timed run, against the oracle on
    the WHOLE image (checker only: nothing here is timed)."""
    import warnings

    import numpy as np
    import torch
    import torch.distributed as dist
    from lars_image_processing_b200.engine import ALL_OUTPUTS, FramePlan
    from oracle import oracle_np as o
    side, tile, grid = 1024, 128, 8
    img = counter_frame_np(4_000_000 - 1, side, side, 1)
    img[: side // 8] //= 3                                   # the first tile row differs: per-rank percentiles would be wrong
    tiles = [np.ascontiguousarray(img[r * tile:(r + 1) * tile, c * tile:(c + 1) * tile]) for r in range(grid) for c in range(grid)]
    b0, b1 = ld.shard_range(len(tiles), rank, world)
    dev = eng.upload(tiles[b0:b1], stream=s)

    def hook(hist):
        if world > 1:
            dist.all_reduce(hist, op=dist.ReduceOp.SUM)
    peer, _ = make_peer_exchange(a, eng, ld, world)
    plan = FramePlan(eng, dev, ALL_OUTPUTS, stream=s, tiles_of_one_image=True, hist_hook=None if peer else hook, peer_exchange=peer)
    res = plan.run()
    whole = ld.dataset_statistics(eng, res.stats, None, s)
    out = eng.download(res, stream=s)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = o.analyze_frame(img)
    ok = True
    for k, t in enumerate(range(b0, b1)):
        r, c = divmod(t, grid)
        sl = (slice(r * tile, (r + 1) * tile), slice(c * tile, (c + 1) * tile))
        ok &= np.array_equal(out[k]["wb"], want["wb"][sl])
        for name in o.INDEX_TYPES:
            ok &= np.array_equal(out[k]["maps"][name].view(np.uint32), want["maps"][name][sl].view(np.uint32))
            ok &= np.array_equal(out[k]["rgb"][name], want["rgb"][name][sl])
    rec = ld.records_to_numpy(whole)
    for i, name in enumerate(o.INDEX_TYPES):
        ws = want["stats"][name]
        ok &= int(rec["count"][i]) == ws["count"] and int(rec["count_above"][i]) == ws["count_above"]
        ok &= np.array_equal(rec["hist"][i][:50], ws["hist"])
        ok &= float(rec["min"][i]) == ws["min"] and float(rec["max"][i]) == ws["max"]
        ref_mean, ref_std = float(np.mean(want["maps"][name])), float(np.std(want["maps"][name]))
        ok &= abs(float(rec["mean"][i]) - ref_mean) <= 1e-6 * max(abs(ref_mean), ref_std, 1e-3)
    flag = torch.tensor([1 if ok else 0], device=eng.device, dtype=torch.int32)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    assert int(flag[0]) == 1, "reduced C4 mosaic differs from the oracle"
    assert peer is None or not peer.timed_out(), "peer exchange timed out"
    return {"checked": f"{side}x{side} mosaic, {grid * grid} tiles of {tile}x{tile}, row band per rank, against the NumPy "
                       "oracle on the whole image: WB bytes, fp32 maps (bits), RGB, histograms, counts, min / max exact; "
                       "mean within 1e-6", "ok": True}


class TimedPeerExchange:
    """A PeerHistogramExchange whose fused exchange + LUT launches are bracketed by CUDA events (bench evidence)."""

    def __init__(self, inner, events):
        self.inner, self.events = inner, events

    def lut_build(self, hist, lut, pct, quantiles=(0.02, 0.98), stream=None, chain=0):
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        self.inner.lut_build(hist, lut, pct, quantiles, stream=stream, chain=chain)
        e1.record(stream)
        self.events.append((e0, e1))

    def timed_out(self):
        return self.inner.timed_out()


def make_peer_exchange(a, eng, ld, world):
    """--exchange peer on more than one rank -> a PeerHistogramExchange, or None (with the reason) -> NCCL hook."""
    if a.exchange != "peer" or world == 1:
        return None, None
    try:
        return ld.PeerHistogramExchange(eng), None
    except Exception as exc:
        return None, str(exc)[:200]


class Workload:
    """Device-resident state of one workload on one rank and its step."""

    def __init__(self, a, eng, ld, rank, world, s):
        import torch
        import torch.distributed as dist
        from lars_image_processing_b200._lib import INDEX_STATS_DTYPE, check
        from lars_image_processing_b200.engine import ALL_OUTPUTS, DeviceFrames, FramePlan
        self.a, self.eng, self.ld, self.rank, self.world, self.s = a, eng, ld, rank, world, s
        self.torch, self.dist, self.check = torch, dist, check
        w = a.w
        self.h, self.wd = w["height"], w["width"]
        self.npx = self.h * self.wd
        self.sb = 1 if w["dtype"] == "u8" else 2
        self.k2_events, self.allreduce_events = [], []
        self.exchange = ld.AsyncDatasetStatistics(eng, timing=True)
        seeds = frame_seeds(a, rank, world)
        self.resident = len(seeds)                       # distinct frames resident in HBM on this rank
        self.frames = eng.alloc_frames(self.resident, self.h, self.wd, 3, s, sample_bytes=self.sb)
        counter_fill_device(self.frames, seeds)
        rec = INDEX_STATS_DTYPE.itemsize
        if a.workload == "c2":
            self.local_frames = self.resident
            self.plan = FramePlan(eng, self.frames, ALL_OUTPUTS, stream=s)
            self.groups = None
            self.launches_per_step = 5 if self.sb == 1 else 7     # hist, lut, fused, finalize, merge (u16: hi, select, lo, build, ...)
        elif a.workload == "c4":
            self.local_frames = self.resident

            def hook(hist):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(s)
                if world > 1:
                    dist.all_reduce(hist, op=dist.ReduceOp.SUM)
                e1.record(s)
                self.allreduce_events.append((e0, e1))
            self.peer, self.peer_note = make_peer_exchange(a, eng, ld, world)
            if self.peer is not None:
                self.plan = FramePlan(eng, self.frames, ALL_OUTPUTS, stream=s, tiles_of_one_image=True,
                                      peer_exchange=TimedPeerExchange(self.peer, self.allreduce_events))
            else:
                self.plan = FramePlan(eng, self.frames, ALL_OUTPUTS, stream=s, tiles_of_one_image=True, hist_hook=hook)
            self.groups = None
            self.launches_per_step = 5
        else:
            G = w["group"]
            total = w["total_frames"]
            self.local_frames = len(range(rank, total, world)) if a.workload == "c3" else \
                ld.shard_range(total, rank, world)[1] - ld.shard_range(total, rank, world)[0]
            n_full, rem = divmod(self.local_frames, G)
            if a.workload == "c3" and self.resident != self.local_frames:
                raise RuntimeError("c3 keeps every local frame resident")
            ring_groups = max(1, self.resident // G)
            self.groups = []                             # (plan, first resident frame) per launch group of the step
            view = lambda f0, n: DeviceFrames(self.frames.data[f0:f0 + n], self.npx, 3, (self.h, self.wd), self.sb)
            self.plan = FramePlan(eng, view(0, min(G, self.resident)), ALL_OUTPUTS, stream=s) if n_full else None
            self.plan_rem = FramePlan(eng, view(0, rem), ALL_OUTPUTS, stream=s) if rem else None
            for g in range(n_full):
                self.groups.append((self.plan, (g % ring_groups) * G if a.workload == "c5" else g * G, G))
            if rem:
                self.groups.append((self.plan_rem, 0 if a.workload == "c5" else n_full * G, rem))
            self.view = view
            self.group_records = torch.zeros((len(self.groups), 3, rec), dtype=torch.uint8, device=eng.device)
            self.launches_per_step = (5 if self.sb == 1 else 7) * len(self.groups)
        self.pixels_per_step = self.local_frames * self.npx
        torch.cuda.synchronize(eng.device)

    def step(self, timed):
        torch, eng, s = self.torch, self.eng, self.s
        ev = lambda: (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        if self.groups is None:
            e = ev() if timed else None
            res = self.plan.run(fused_events=e)
            if timed:
                self.k2_events.append((e, self.plan.frames.n_frames))
            self.exchange.submit(res.stats, stream=s)
            return
        for k, (plan, f0, n) in enumerate(self.groups):
            plan.rebind(self.view(f0, n))
            e = ev() if timed else None
            res = plan.run(fused_events=e)
            if timed:
                self.k2_events.append((e, n))
            self.check(eng.lib.lars_stats_merge(res.stats.data_ptr(), n, self.group_records[k].data_ptr(), s.cuda_stream),
                       "lars_stats_merge")
        self.exchange.submit(self.group_records, stream=s)

    def last_stats_records(self):
        """Per-frame records of the most recent launch (sanity checks)."""
        plan = self.plan if self.plan is not None else self.plan_rem
        return plan.out.stats


def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    from lars_image_processing_b200 import distributed as ld
    from lars_image_processing_b200.engine import ALL_OUTPUTS, DeviceFrames, Engine, FramePlan

    rank, world, local_rank = ld.init_from_env()
    if world != a.gpus and rank == 0:
        print(f"[bench] note: --gpus {a.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)
    torch.cuda.set_device(local_rank)
    affinity = "unchanged"
    if world > 1 and os.environ.get("LARS_BENCH_GPU_LOCAL_CPUS", "1") == "1":
        # pinned staging buffers should live on the NUMA node the GPU hangs off: bind this rank to the
        # GPU-local CPUs before anything is allocated (ignored when the container's cpuset forbids it)
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
            affinity = f"gpu-local ({len(os.sched_getaffinity(0))} cpus)"
        except Exception:
            pass
    eng = Engine(local_rank)
    s = eng.stream()
    w = a.w
    h, wd = w["height"], w["width"]
    npx = h * wd
    sb = 1 if w["dtype"] == "u8" else 2

    parity = None
    if a.workload == "c4" and not a.no_parity:
        parity = reduced_mosaic_parity(a, eng, ld, rank, world, s)

    wl = Workload(a, eng, ld, rank, world, s)
    exchange = wl.exchange
    host_s = [0.0]

    def step(timed):
        t_host = time.perf_counter()
        wl.step(timed)
        host_s[0] += time.perf_counter() - t_host

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    align = torch.zeros(1, device=eng.device)

    def aligned_start(ev):
        """Every rank's GPU timeline starts together: after the host barrier the ranks' threads wake at different
        times, so a tiny all-reduce is enqueued right in front of the start event -- it completes on every GPU at
        the same moment, and from there each GPU runs its own queue back to back."""
        with torch.cuda.stream(s):
            if world > 1:
                dist.all_reduce(align)
            ev.record(s)

    def reduce_max(vals):
        t = torch.tensor(vals, device=eng.device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def gather_ranks(vals):
        t = torch.tensor(vals, device=eng.device, dtype=torch.float64)
        if world == 1:
            return [[float(x)] for x in t]
        out = torch.empty((world, len(vals)), device=eng.device, dtype=torch.float64)
        dist.all_gather_into_tensor(out, t)
        return [[float(out[r, k]) for r in range(world)] for k in range(len(vals))]

    def k2_stats():
        """(mean ms per K2 launch, frames per launch weighted) over the recorded event pairs; resets the list."""
        ms = [e0.elapsed_time(e1) for (e0, e1), _ in wl.k2_events]
        fr = [n for _, n in wl.k2_events]
        wl.k2_events = []
        return sum(ms) / len(ms), sum(fr) / len(fr), sum(ms)

    # NVML is initialised and the sampling thread started BEFORE the barrier: nvmlInit takes a different
    # time on every rank, and anything between the barrier and t_start becomes start skew
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(a.warmup):
        step(False)
    t_start, t_work, t_end = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    host_s[0] = 0.0
    barrier()
    exchange.gather_us()                                    # drop the warm-up's timings
    wl.allreduce_events.clear()
    sampler.recording = True
    aligned_start(t_start)
    for _ in range(a.steps):
        step(True)
    t_work.record(s)                                        # this rank's own work is done ...
    dataset = exchange.result(s)                            # ... and every exchange has landed inside the timed region
    t_end.record(s)
    barrier()
    sampler.recording = False
    clocks = sampler.snapshot()
    host_enqueue_ms = host_s[0] / max(1, a.steps) * 1e3
    ms_own = t_start.elapsed_time(t_end)
    tail_us = t_work.elapsed_time(t_end) * 1e3
    k2_ms_own, k2_frames, k2_total_ms = k2_stats()
    ms, k2_ms = reduce_max([ms_own, k2_ms_own])
    per_rank = gather_ranks([ms_own, k2_ms_own, tail_us])
    gather_us = exchange.gather_us()
    allreduce_us = [e0.elapsed_time(e1) * 1e3 for e0, e1 in wl.allreduce_events]
    px_all = torch.tensor([wl.pixels_per_step], device=eng.device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(px_all)
    total_px = float(px_all[0])
    value = total_px * a.steps / (ms * 1e-3) / 1e6

    # sanity: the timed work really happened (histogram totals == pixels; the exchange merged every rank)
    rec = ld.records_to_numpy(wl.last_stats_records())
    assert int(rec["hist"][0, 0].sum()) == npx and int(rec["count"][-1, 2]) == npx
    assert int(ld.records_to_numpy(dataset)["count"][0]) == int(total_px)

    # ---- the same step looped for >= sustain_s seconds (clocks under a long run, sw_power_cap)
    sustained = None
    if a.sustain_s > 0:
        n_sus = max(a.steps, int(a.sustain_s / (ms / a.steps * 1e-3)) + 1)
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        sampler.recording = True
        aligned_start(t0e)
        for i in range(n_sus):
            wl.step(i % 8 == 0)                             # K2 events on every 8th step
        exchange.result(s)
        t1e.record(s)
        barrier()
        sampler.recording = False
        sus_clocks = sampler.snapshot()
        sus_k2_ms, sus_k2_frames, _ = k2_stats()
        sus_ms, sus_k2_ms = reduce_max([t0e.elapsed_time(t1e), sus_k2_ms])
        exchange.gather_us()
        wl.allreduce_events.clear()
        sustained = {"seconds": sus_ms * 1e-3, "steps": n_sus, "value": total_px * n_sus / (sus_ms * 1e-3) / 1e6,
                     "unit": UNIT, "k2_ms_per_launch": sus_k2_ms, "clocks": sus_clocks}
    sampler.stop()

    # ---- BASELINE config 2 read literally: ONE frame per pass (latency-bound: ~60 us of traffic per frame),
    # the whole pass replayed as a CUDA graph
    single = None
    if a.workload == "c2" and sb == 1 and rank == 0 and not a.no_single:
        one = eng.alloc_frames(1, h, wd, 3, s)
        counter_fill_device(one, [99])
        plan1 = FramePlan(eng, one, ALL_OUTPUTS, stream=s).capture()
        for _ in range(5):
            plan1.replay()
        s.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 50
        g0.record(s)
        for _ in range(reps):
            plan1.replay()
        g1.record(s)
        s.synchronize()
        us = g0.elapsed_time(g1) / reps * 1e3
        single = {"frames_per_pass": 1, "us_per_pass": us, "value": npx / us, "unit": UNIT,
                  "note": "one frame per pass, CUDA-graph replay of Pass 1 + LUT + fused Pass 2 + finalize; "
                          "the frame (36 MB) stays L2-resident between the passes"}
        del plan1, one

    # ---- end to end through the host-array API (pinned buffers, H2D + D2H in the timed region)
    e2e = None
    if not a.no_e2e:
        e2e = run_e2e(a, eng, ld, wl, rank, world, barrier, reduce_max)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_hbm_peak()
    pass1_b = 3 if sb == 1 else 6                             # algorithmic; uint16 reads the frame a second time
    pass2_b = 27 if sb == 1 else 30                           # read raw, write WB u8 + 3 fp32 maps + 3 RGB images
    launch_bytes = pass2_b * k2_frames * npx
    achieved = launch_bytes / (k2_ms * 1e-3) / 1e9
    traffic, traffic_src = profiled_traffic(a)
    cfg = common_config(a, world)
    resident_mb = wl.resident * npx * 3 * sb / 1e6
    cfg.update({
        "l2": f"inputs larger than L2 ({resident_mb:.0f} MB of distinct raw frames resident per GPU, "
              f"{launch_bytes / 1e6:.0f} MB touched per fused launch)",
        "cpu_affinity": affinity,
        "parallelism": ("single GPU" if world == 1 else {
            "c2": f"frames sharded over {world} GPUs, one dataset-statistics all-gather per step on a side stream (overlaps the next step)",
            "c3": f"frames round-robin over {world} GPUs, no data-path collective, one dataset-statistics all-gather per step",
            "c4": (f"row bands of tiles over {world} GPUs, white-balance histogram exchanged inside every step by the LUT kernel "
                   "itself over NVLink peer memory (symmetric buffers, per-rank flags), one image-statistics all-gather per step"
                   if getattr(wl, "peer", None) is not None else
                   f"row bands of tiles over {world} GPUs, SUM all-reduce of the white-balance histogram between Pass 1 and the LUT "
                   "build inside every step, one image-statistics all-gather per step"),
            "c5": f"contiguous shards over {world} GPUs, no data-path collective, one dataset-statistics all-gather per step"}[a.workload]),
    })
    if a.workload in ("c3", "c5"):
        cfg["group"] = w["group"]
        cfg["launch_groups_per_step_per_gpu"] = len(wl.groups)
    if a.workload == "c5":
        cfg["ring"] = (f"{wl.resident} distinct frames resident per GPU, reused cyclically by the {len(wl.groups)} launch groups of a step "
                       "(the whole survey, 368.6 GB, does not fit in HBM; refilling the ring is ingest, outside `value`)")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None,
        "dtype": w["dtype"], "data": "synthetic (counter-based integer generator, identical on device and host)",
        "config": cfg,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram__bytes_read+write)",
                     "traffic_source": traffic_src,
                     "algorithmic_bytes_per_launch": launch_bytes,
                     "kernel": f"fused_index_kernel<3,{sb}> (+ workspace memset and statistics finalize inside the event pair)",
                     "algorithmic_bytes_per_px": pass2_b, "ms_per_launch": k2_ms, "frames_per_launch": k2_frames,
                     "peak_source": peak_src,
                     "whole_step_GBps": (pass1_b + pass2_b) * total_px / world * a.steps / (ms * 1e-3) / 1e9,
                     "whole_step_frac": (pass1_b + pass2_b) * total_px / world * a.steps / (ms * 1e-3) / 1e9 / peak,
                     "whole_step_bytes_per_px": pass1_b + pass2_b},
        "clocks": clocks, "host_enqueue_ms_per_step": host_enqueue_ms,
        "ranks": {"ms": per_rank[0], "k2_ms_per_launch": per_rank[1], "exchange_tail_us": per_rank[2],
                  "note": "per rank: own timed region, mean fused-launch time, wait between the end of the rank's own "
                          "work and the arrival of the last exchange (`value` uses the maximum over ranks)"},
        "collectives": {"stats_all_gather_us": summarize(gather_us), "wb_hist_all_reduce_us": summarize(allreduce_us),
                        "note": "device time on rank 0 (CUDA events on the stream the collective runs on); a collective "
                                "also waits for the slowest rank to arrive"},
        "gpu_launches": (wl.launches_per_step + (1 if world > 1 else 0)) * a.steps,
    }
    if parity is not None:
        line["parity_check"] = parity
    if a.workload == "c4":
        line["collectives"]["wb_hist_exchange"] = "peer" if getattr(wl, "peer", None) is not None else "nccl"
        if getattr(wl, "peer", None) is not None:
            line["collectives"]["note"] += ("; with the peer exchange wb_hist_all_reduce_us is the fused exchange + LUT kernel "
                                            "(3.9 us of it is the LUT build)")
            assert not wl.peer.timed_out(), "peer exchange timed out"
        elif getattr(wl, "peer_note", None):
            line["collectives"]["wb_hist_exchange_note"] = "peer exchange unavailable, NCCL used: " + wl.peer_note
    if sustained is not None:
        sustained["roofline_frac"] = pass2_b * sus_k2_frames * npx / (sustained["k2_ms_per_launch"] * 1e-3) / 1e9 / peak
        line["sustained"] = sustained
    if e2e is not None:
        line["e2e"] = e2e
    if single is not None:
        line["single_frame"] = single
    if world == 1 and not a.no_cpu_baseline:
        n_cpu = min(a.cpu_frames, wl.resident)
        if npx >= 15_000_000:
            n_cpu = min(n_cpu, 2)
        elif npx < 2_000_000:
            n_cpu = min(wl.resident, max(n_cpu, 24))
        raw = [wl.frames.data[i, :npx * 3 * sb].cpu().numpy() for i in range(n_cpu)]
        sample = [(r if sb == 1 else r.view(np.uint16)).reshape(h, wd, 3) for r in raw]
        assert np.array_equal(sample[0], counter_frame_np(frame_seeds(a, 0, 1)[0], h, wd, sb)), "device / host generators differ"
        v, dt, kind = cpu_baseline_single(sample)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": kind,
                                "sample": f"{len(sample)} of the step's {wd}x{h} frames, sequential NumPy path "
                                          f"(WB + 3 indices + analyze_index + std + hist(50) + colormap), {dt:.1f} s; "
                                          f"host has {os.cpu_count()} cores, the reference path is single-threaded"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def summarize(us):
    if not us:
        return None
    s = sorted(us)
    return {"n": len(s), "median": s[len(s) // 2], "min": s[0], "max": s[-1]}


def run_e2e(a, eng, ld, wl, rank, world, barrier, reduce_max):
    """The metric through the host-array API: pinned host frames in, host results out, copies inside the timed
    region.  c2: the step's 16 frames, every product back.  c3: a bounded sample of the rank's frames (16), every
    product back.  c4: the rank's tiles through Engine.run_host_mosaic (image-wide histogram all-reduce inside),
    every product back.  c5: the ring's frames in survey mode (statistics records only come back)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from lars_image_processing_b200.engine import ALL_OUTPUTS
    w = a.w
    h, wd = w["height"], w["width"]
    npx, sb = h * wd, wl.sb
    shape = (h, wd, 3)
    rec_dtype = ld.INDEX_STATS_DTYPE
    k_e2e = a.e2e_steps or min(a.steps, 5)
    if a.workload == "c5":
        F = wl.resident
        outputs, chunk = ("stats",), a.chunk or 64
        k_e2e = a.e2e_steps or min(a.steps, 2)
    elif a.workload == "c3":
        F = min(wl.resident, 16)
        outputs, chunk = ALL_OUTPUTS, a.chunk or 2
        k_e2e = a.e2e_steps or min(a.steps, 3)
    elif a.workload == "c4":
        F = wl.resident
        outputs, chunk = ALL_OUTPUTS, a.chunk or 2
        k_e2e = a.e2e_steps or min(a.steps, 3)
    else:
        F = wl.resident
        outputs, chunk = ALL_OUTPUTS, a.chunk or 2
    host_in = torch.empty((F, npx * 3 * sb), dtype=torch.uint8, pin_memory=True)
    host_in.copy_(wl.frames.data[:F, :npx * 3 * sb])
    torch.cuda.synchronize()
    host_out = eng.alloc_host_outputs(F, h, wd, 3, outputs)

    def hook(hist):
        if world > 1:
            dist.all_reduce(hist, op=dist.ReduceOp.SUM)

    def once(out):
        if a.workload == "c4":
            stats = eng.run_host_mosaic(host_in, shape, out, chunk=chunk, hist_hook=hook)
            ld.dataset_statistics(eng, stats)
        else:
            eng.run_host_batch(host_in, shape, out, chunk=chunk, sample_bytes=sb)

    def timed(out):
        once(out)                                           # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            once(out)
        torch.cuda.synchronize()
        return reduce_max([time.perf_counter() - t0])[0]

    dt = timed(host_out)
    px = torch.tensor([F * npx], device=eng.device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(px)
    total = float(px[0])
    h2d = F * npx * 3 * sb
    per_px_back = (3 if "wb" in outputs else 0) + (12 if "maps" in outputs else 0) + (9 if "rgb" in outputs else 0)
    d2h = F * (npx * per_px_back + 3 * 576)
    api = {"c4": "Engine.run_host_mosaic + distributed.dataset_statistics (pinned, pipelined)",
           "c5": "Engine.run_host_batch, statistics records only (pinned, pipelined)"}.get(
        a.workload, "Engine.run_host_batch (pinned, pipelined)")
    e2e = {"value": total * k_e2e / dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "steps": k_e2e, "api": api,
           "sample": f"{F} frames per GPU per e2e step" + ("" if F == wl.local_frames else
                                                            f" (bounded sample of the rank's {wl.local_frames})")}
    rec_h = host_out["stats"].numpy().view(rec_dtype).reshape(F, 3)
    assert int(rec_h["count"][0, 0]) == npx and int(rec_h["hist"][-1, 2].sum()) == npx    # the host results are the real thing
    # what the box allows: every rank copies device -> pinned host at the same time (the e2e path returns 24 B/px and
    # takes 3 B/px in, so D2H is its wire); the aggregate saturates on the host side of multi-GPU boxes
    # (profiles/r02_pcie_probe_n*.log), and `e2e` is to be read against this ceiling, not against N x one GPU's link
    if d2h > (64 << 20):
        nbytes = 256 << 20
        dsrc = torch.empty(nbytes, dtype=torch.uint8, device=eng.device)
        hdst = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        hdst.copy_(dsrc, non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(4):
            hdst.copy_(dsrc, non_blocking=True)
        torch.cuda.synchronize()
        dt_c = reduce_max([time.perf_counter() - t0])[0]
        ceiling = world * 4 * nbytes / dt_c / 1e9
        used = world * d2h * k_e2e / dt / 1e9
        e2e["host_ceiling_GBps"] = ceiling
        e2e["d2h_GBps"] = used
        e2e["d2h_frac_of_ceiling"] = used / ceiling
        e2e["ceiling_note"] = (f"aggregate pinned D2H of {world} rank(s) copying at once, 4 x 256 MiB each, measured right after the "
                               "e2e leg; the e2e path moves 8x more bytes back than in")
        del dsrc, hdst
    if a.workload == "c2":
        # survey mode (BASELINE config 5's result: statistics only): same call, only the records come back
        host_stats = eng.alloc_host_outputs(F, h, wd, 3, ("stats",))
        dt2 = timed(host_stats)
        e2e["stats_only"] = {"value": total * k_e2e / dt2 / 1e6, "unit": UNIT,
                             "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": F * 3 * 576}
        assert np.array_equal(host_stats["stats"].numpy().view(rec_dtype).reshape(F, 3)["hist"], rec_h["hist"])
    return e2e


_RESULT_OUT = None


def emit(line: dict) -> None:
    """The ONE JSON line of the contract, on the process's original stdout."""
    (_RESULT_OUT or sys.stdout).write(json.dumps(line) + "\n")
    (_RESULT_OUT or sys.stdout).flush()


def main():
    global _RESULT_OUT
    # stdout carries exactly one JSON line: anything a library writes to fd 1 (NCCL prints its version
    # banner there) is sent to stderr instead
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
