"""K10 timing: Pillow-exact Lanczos resize of C2 frames to the app's 1024-pixel working size,
CUDA events on the launching stream, next to Pillow on one host core."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from PIL import Image
from lars_image_processing_b200.engine import get_engine
from oracle import synth

eng = get_engine()
s = eng.stream()
for (h, w, F) in ((3000, 4000, 16), (3648, 5472, 8), (960, 1280, 64)):
    t = eng.preprocess_target(h, w, 1024) or (h // 2, w // 2)
    frames = [synth.vegetation_frame(100 + i, h, w) for i in range(min(F, 2))]
    dev = eng.upload([frames[i % len(frames)] for i in range(F)], stream=s)
    for _ in range(3):                                   # warm-up: plan, allocator pools
        out = eng.resize_device(dev, t[0], t[1], s)
    s.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    with torch.cuda.stream(s):
        e0.record(s)
        for _ in range(n):
            out = eng.resize_device(dev, t[0], t[1], s)
        e1.record(s)
    s.synchronize()
    ms = e0.elapsed_time(e1) / n
    in_b, out_b = F * h * w * 3, F * t[0] * t[1] * 3
    t0 = time.perf_counter()
    ref = np.array(Image.fromarray(frames[0]).resize((t[1], t[0]), Image.Resampling.LANCZOS))
    cpu_ms = (time.perf_counter() - t0) * 1e3
    got = out.data[0, :t[0] * t[1] * 3].cpu().numpy().reshape(t[0], t[1], 3)
    print(f"{w}x{h} x{F} -> {t[1]}x{t[0]}: {ms * 1e3:8.1f} us/batch  {ms * 1e3 / F:7.1f} us/frame  "
          f"{(in_b + out_b) / ms / 1e6:7.0f} GB/s algorithmic (in+out)  {F * h * w / ms / 1e6:7.1f} Gpix/s  |  "
          f"Pillow 1 core {cpu_ms:6.1f} ms/frame  bit-identical={np.array_equal(got, ref)}")
