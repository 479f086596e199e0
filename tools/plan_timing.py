#!/usr/bin/env python
"""Host-overhead check: frames/s of small batches through Engine.process_device (allocating,
Python-heavy), FramePlan.run (pre-bound) and FramePlan.replay (CUDA graph)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lars_image_processing_b200.engine import Engine, FramePlan, ALL_OUTPUTS
sys.path.insert(0, ".")
from bench import counter_fill_device

eng = Engine(0)
for (h, w, F) in ((960, 1280, 1), (960, 1280, 16), (3000, 4000, 1), (3000, 4000, 2)):
    frames = counter_fill_device(eng.alloc_frames(F, h, w, 3), [3 + i for i in range(F)])
    s = eng.stream()
    res = eng.alloc_outputs(frames, ALL_OUTPUTS, s)
    plan = FramePlan(eng, frames).capture()
    def timed(fn, n=200):
        for _ in range(10): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(n): fn()
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
    t_dev = timed(lambda: eng.process_device(frames, out=res, stream=s))
    t_run = timed(plan.run)
    t_rep = timed(plan.replay)
    mp = F * h * w / 1e6
    print(f"{w}x{h} x{F}: process_device {t_dev:7.1f} us  plan.run {t_run:7.1f} us  graph replay {t_rep:7.1f} us  "
          f"-> {mp / t_rep * 1e6 / 1e3:8.1f} Gpix/s with the graph")
