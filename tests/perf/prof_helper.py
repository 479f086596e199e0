import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from oracle import synth
from lars_image_processing_b200.engine import get_engine
eng = get_engine(); s = eng.stream()
img = synth.vegetation_frame(2, 3000, 4000)
for it in range(6):
    t0 = time.perf_counter(); dev = eng.upload([img], stream=s); s.synchronize()
    t1 = time.perf_counter(); res = eng.process_device(dev, stream=s); s.synchronize()
    t2 = time.perf_counter(); out = eng.download(res, stream=s)
    t3 = time.perf_counter()
    print(f"iter {it}: upload {1e3*(t1-t0):.2f} ms  process {1e3*(t2-t1):.2f}  download {1e3*(t3-t2):.2f}")
    del out
# pinned alloc cost alone
for it in range(4):
    t0 = time.perf_counter(); x = torch.empty((3, 1, 12000000), dtype=torch.float32, pin_memory=True); t1 = time.perf_counter()
    print(f"pinned alloc 144 MB: {1e3*(t1-t0):.2f} ms"); del x
