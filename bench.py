#!/usr/bin/env python
"""Benchmark of the RGNir per-pixel analysis path (BASELINE.json metric: RGNir Mpix/s for
WB + NDVI/GNDVI/NDWI + stats [+ colormap], % of HBM peak).

    python bench.py --gpus N --steps K --warmup W          # our CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's NumPy path on host cores

A "step" is one pass of the hot path (Pass 1 histogram -> LUT -> fused Pass 2 -> statistics
finalize [-> dataset-statistics exchange when N > 1]) over one batch of synthetic frames of the
workload configuration (BASELINE config 2: 4000x3000 uint8 RGNir frames).  Each rank holds its
own batch (weak scaling: frames are independent units).

One JSON line on stdout (rank 0).  `value` is device-resident throughput (inputs in HBM when
the timed region starts, CUDA events, max over ranks); `e2e` goes through the host-array API
with pinned host buffers and H2D / D2H inside the timed region; `roofline` is the fused Pass-2
kernel against the measured HBM copy bandwidth; `cpu_baseline` is the NumPy oracle port timed
on this box's host cores on a bounded sample.
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
U8_PASS1_BYTES_PER_PX = 3            # K1 reads the raw frame once
U8_PASS2_BYTES_PER_PX = 3 + 3 + 12 + 9   # K2: read raw, write WB u8 + 3 fp32 maps + 3 RGB images
FALLBACK_HBM_GBS = 6650.0            # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=16, help="frames per rank per step")
    ap.add_argument("--height", type=int, default=3000)
    ap.add_argument("--width", type=int, default=4000)
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = same as --steps (capped at 5)")
    ap.add_argument("--cpu-frames", type=int, default=3, help="frames of the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-single", action="store_true", help="skip the one-frame-per-pass leg (profiling runs)")
    ap.add_argument("--chunk", type=int, default=2, help="frames per pipeline chunk in the e2e path")
    ap.add_argument("--dtype", default="u8", choices=["u8", "u16"],
                    help="sample type of the synthetic frames (u16: 16-bit frames as in BASELINE config 3)")
    return ap.parse_args()


def workload_name(a):
    shape = (a.dtype, a.width, a.height)
    tag = {("u8", 4000, 3000): "C2", ("u16", 5472, 3648): "C3 (per-GPU slice of the 20 MP 16-bit batch)",
           ("u8", 1280, 960): "C5 (one launch group of the small-frame survey)",
           ("u8", 4096, 4096): "C4 (4096^2 tiles, per-tile white balance)"}.get(shape, "custom")
    return (f"{tag}: {a.width}x{a.height} {'uint8' if a.dtype == 'u8' else 'uint16'} RGNir frames, white balance + "
            f"NDVI/GNDVI/NDWI fp32 maps + statistics/histograms + colormap RGB; {a.frames} distinct frames per GPU per step")


def profiled_traffic(a):
    """DRAM bytes per K2 launch from the committed `ncu --set full` capture of this exact workload
    (profiles/r01_k2_fused_index_ncu_summary.txt: 16 C2 frames per launch); None for other shapes."""
    if not (a.dtype == "u8" and a.width == 4000 and a.height == 3000 and a.frames == 16):
        return None
    try:
        total = 0.0
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        with open(os.path.join(ROOT, "profiles", "r01_k2_fused_index_ncu_summary.txt")) as fh:
            for line in fh:
                parts = line.split()
                if len(parts) >= 3 and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    total += float(parts[1]) * mult[parts[2]]
        return total or None
    except Exception:
        return None


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


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

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------
# CPU side: the oracle port of the reference's NumPy path
# ------------------------------------------------------------------------------------------
_CPU_FRAME = None


def _cpu_worker_init(h, w):
    """Each worker process builds its synthetic frame ONCE, outside any timed region."""
    global _CPU_FRAME
    from oracle import synth
    _CPU_FRAME = synth.vegetation_frame(2 + os.getpid() % 1000, h, w)


def _cpu_one_frame(_):
    """The reference's NumPy path (oracle port) on this worker's cached frame."""
    import warnings
    from oracle import oracle_np as o
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        o.reference_cpu_path(_CPU_FRAME)
    return time.perf_counter() - t0


def cpu_baseline_single(frames_host):
    """Sequential, one core: the NumPy calls of the path are single-threaded."""
    import warnings
    from oracle import oracle_np as o
    t0 = time.perf_counter()
    for img in frames_host:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            o.reference_cpu_path(img)
    dt = time.perf_counter() - t0
    npx = sum(f.shape[0] * f.shape[1] for f in frames_host)
    return npx / dt / 1e6, dt


def run_reference(a):
    """--impl reference: the reference's CPU implementation (NumPy oracle port; the reference is
    Python and cannot travel to this box) on all host cores: one independent frame per worker
    process per step, like the reference's own per-file loop (backend-process.py:92-97)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 64))
    ctx = mp.get_context("spawn")
    # bounded sample: one frame per worker per step costs ~3 s at C2 size; for long runs the frames are
    # cropped to their first rows (same width, same per-pixel work) so that the whole run stays within minutes
    rows = a.height if a.steps <= 20 else max(256, a.height * 20 // a.steps)
    npx = rows * a.width
    with ctx.Pool(workers, initializer=_cpu_worker_init, initargs=(rows, a.width)) as pool:
        for _ in range(max(1, min(a.warmup, 1))):
            pool.map(_cpu_one_frame, range(workers), chunksize=1)
        t0 = time.perf_counter()
        for _ in range(a.steps):
            pool.map(_cpu_one_frame, range(workers), chunksize=1)
        dt = time.perf_counter() - t0
    value = workers * a.steps * npx / dt / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(a),
                   "sample": f"{workers} frames of {a.width}x{rows} per step, one per worker process"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port",
                         "sample": f"{workers * a.steps} frames of {a.width}x{rows} over {workers} processes "
                                   "(NumPy oracle port of process-images.py:424-513 + std + hist(50) + colormap)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------
def synth_frames_device(eng, n_frames, h, w, seed, sample_bytes=1):
    """Vegetation-like frames generated on the device: R~N(90,35) G~N(110,35) NIR~N(150,45)
    (x257 for uint16)."""
    import torch
    frames = eng.alloc_frames(n_frames, h, w, 3, sample_bytes=sample_bytes)
    npx = h * w
    g = torch.Generator(device=eng.device)
    g.manual_seed(seed)
    scale = 1.0 if sample_bytes == 1 else 257.0
    top = 255 if sample_bytes == 1 else 65535
    mean = torch.tensor([90.0, 110.0, 150.0], device=eng.device) * scale
    std = torch.tensor([35.0, 35.0, 45.0], device=eng.device) * scale
    for f in range(n_frames):
        x = (torch.randn((npx, 3), generator=g, device=eng.device) * std + mean).round_().clamp_(0, top)
        if sample_bytes == 1:
            frames.data[f, :npx * 3] = x.to(torch.uint8).reshape(-1)
        else:
            frames.data[f, :npx * 6] = x.to(torch.int32).to(torch.int16).reshape(-1).view(torch.uint8)
        del x
    torch.cuda.synchronize(eng.device)
    return frames


def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    from lars_image_processing_b200 import distributed as ld
    from lars_image_processing_b200.engine import ALL_OUTPUTS, Engine, FramePlan

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
    F, h, w = a.frames, a.height, a.width
    npx = h * w
    sb = 1 if a.dtype == "u8" else 2
    frames = synth_frames_device(eng, F, h, w, seed=2 + rank, sample_bytes=sb)
    # uint8: pre-bound plan (buffers, workspace and the C argument block are created once; a step is
    # three C-ABI calls with no allocation), so eight ranks sharing the host cores stay ahead of their GPUs
    plan = FramePlan(eng, frames, ALL_OUTPUTS, stream=s)
    res = plan.out

    fused_ms = []
    host_s = [0.0]
    exchange = ld.AsyncDatasetStatistics(eng)   # all-gather + merge on a side stream, overlapped with the next step

    def step(timed):
        t_host = time.perf_counter()
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) if timed else None
        plan.run(fused_events=ev)           # uint8: K1, K1b, K2 (+ finalize); uint16: two-level histogram, stretch build, K2
        if timed:
            fused_ms.append(ev)
        exchange.submit(res.stats, stream=s)                # local merge (+ one NCCL all-gather when N > 1)
        host_s[0] += time.perf_counter() - t_host

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # NVML is initialised and the sampling thread started BEFORE the barrier: nvmlInit takes a different
    # time on every rank, and anything between the barrier and t_start becomes start skew that the last
    # step's exchange turns into time for every rank
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(a.warmup):
        step(False)
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host_s[0] = 0.0
    barrier()
    sampler.recording = True
    t_start.record(s)
    for _ in range(a.steps):
        step(True)
    dataset = exchange.result(s)                            # every exchange has landed inside the timed region
    t_end.record(s)
    barrier()
    sampler.recording = False
    clocks = sampler.stop()
    host_enqueue_ms = host_s[0] / max(1, a.steps) * 1e3
    ms = t_start.elapsed_time(t_end)
    k2_ms = sum(e0.elapsed_time(e1) for e0, e1 in fused_ms) / len(fused_ms)
    t = torch.tensor([ms, k2_ms], device=eng.device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, k2_ms = float(t[0]), float(t[1])
    value = world * F * npx * a.steps / (ms * 1e-3) / 1e6

    # sanity: the timed work really happened (histogram totals == pixels)
    rec = ld.records_to_numpy(res.stats)
    assert int(rec["hist"][0, 0].sum()) == npx and int(rec["count"][-1, 2]) == npx
    assert int(ld.records_to_numpy(dataset)["count"][0]) == world * F * npx   # the exchange really merged every rank

    # ---- BASELINE config 2 read literally: ONE frame per pass (latency-bound: ~60 us of traffic per frame),
    # the whole pass replayed as a CUDA graph
    single = None
    if sb == 1 and rank == 0 and not a.no_single:
        one = synth_frames_device(eng, 1, h, w, seed=99, sample_bytes=1)
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
        host_in = torch.empty((F, npx * 3 * sb), dtype=torch.uint8, pin_memory=True)
        host_in.copy_(frames.data[:, :npx * 3 * sb])
        torch.cuda.synchronize()
        host_out = eng.alloc_host_outputs(F, h, w, 3, ALL_OUTPUTS)
        k_e2e = a.e2e_steps or min(a.steps, 5)
        eng.run_host_batch(host_in, (h, w, 3), host_out, chunk=a.chunk, sample_bytes=sb)      # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            eng.run_host_batch(host_in, (h, w, 3), host_out, chunk=a.chunk, sample_bytes=sb)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], device=eng.device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt[0])
        h2d = F * npx * 3 * sb
        d2h = F * (npx * 3 + 3 * npx * 4 + 3 * npx * 3 + 3 * 576)
        e2e = {"value": world * F * npx * k_e2e / dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": k_e2e, "api": "Engine.run_host_batch (pinned, pipelined)"}
        # the host results are the real thing
        rec_h = host_out["stats"].numpy().view(rec.dtype).reshape(F, 3)
        assert np.array_equal(rec_h["hist"], rec["hist"])
        # survey mode (BASELINE config 5's result: statistics only): same call, only the records come back
        host_stats = eng.alloc_host_outputs(F, h, w, 3, ("stats",))
        eng.run_host_batch(host_in, (h, w, 3), host_stats, chunk=a.chunk, sample_bytes=sb)
        barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            eng.run_host_batch(host_in, (h, w, 3), host_stats, chunk=a.chunk, sample_bytes=sb)
        torch.cuda.synchronize()
        tt = torch.tensor([time.perf_counter() - t0], device=eng.device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e["stats_only"] = {"value": world * F * npx * k_e2e / float(tt[0]) / 1e6, "unit": UNIT,
                             "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": F * 3 * 576}
        assert np.array_equal(host_stats["stats"].numpy().view(rec.dtype).reshape(F, 3)["hist"], rec["hist"])

    if rank != 0:
        return
    peak, peak_src = measured_hbm_peak()
    pass1_b = U8_PASS1_BYTES_PER_PX if sb == 1 else 12        # u16: two-level histogram reads the frame twice
    pass2_b = U8_PASS2_BYTES_PER_PX if sb == 1 else 30        # u16: read 6, write 3 + 12 + 9
    launch_bytes = pass2_b * F * npx
    achieved = launch_bytes / (k2_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": a.dtype, "data": "synthetic",
        "config": {"workload": workload_name(a), "frames_per_gpu": F, "height": h, "width": w,
                   "l2": f"inputs larger than L2 ({F * npx * 3 * sb / 1e6:.0f} MB raw per GPU per step)",
                   "cpu_affinity": affinity,
                   "parallelism": f"frames sharded over {world} GPU(s), one dataset-statistics all-gather per step on a side stream (overlaps the next step)"
                   if world > 1 else "single GPU"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": profiled_traffic(a), "traffic_unit": "bytes per launch (ncu dram__bytes_read+write)",
                     "algorithmic_bytes_per_launch": launch_bytes,
                     "kernel": "fused_index_kernel<3,1> (+ workspace memset and statistics finalize inside the event pair)",
                     "algorithmic_bytes_per_px": pass2_b, "ms_per_launch": k2_ms,
                     "peak_source": peak_src,
                     "whole_step_GBps": (pass1_b + pass2_b) * F * npx * a.steps
                     / (ms * 1e-3) / 1e9},
        "clocks": clocks, "host_enqueue_ms_per_step": host_enqueue_ms,
        # wb_hist, wb_lut_build, fused_index, fused_finalize, stats_merge (+1 merge after the all-gather)
        "gpu_launches": (5 if world == 1 else 6) * a.steps,
    }
    if e2e is not None:
        line["e2e"] = e2e
    if single is not None:
        line["single_frame"] = single
    if world == 1 and not a.no_cpu_baseline:
        raw = [frames.data[i, :npx * 3 * sb].cpu().numpy() for i in range(min(a.cpu_frames, F))]
        sample = [(r if sb == 1 else r.view(np.uint16)).reshape(h, w, 3) for r in raw]
        v, dt = cpu_baseline_single(sample)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": f"{len(sample)} of the step's {w}x{h} frames, sequential NumPy oracle port "
                                          f"(WB + 3 indices + analyze_index + std + hist(50) + colormap), {dt:.1f} s; "
                                          f"host has {os.cpu_count()} cores, the reference path is single-threaded"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


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
