"""Device-side LZW decode against the host decoder on all cores and against Pillow:
16 BASELINE config 2 frames (4000x3000 uint8) stored as LZW TIFF by Pillow / libtiff, noise and 4x4-block
content.  Timed: host parse + staging + H2D + decode kernels (what a caller pays), and the kernels alone."""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from concurrent.futures import ThreadPoolExecutor
from PIL import Image
from oracle import synth
from lars_image_processing_b200 import ingest
from lars_image_processing_b200.engine import get_engine

eng = get_engine()
tmp = tempfile.mkdtemp(prefix="lars_lzw_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
N = 16
for name in ("noise", "4x4 blocks"):
    paths, imgs = [], []
    for i in range(N):
        img = synth.vegetation_frame(100 + i, 3000, 4000)
        if name != "noise":
            img = np.ascontiguousarray(np.repeat(np.repeat(img[::4, ::4], 4, 0), 4, 1))
        p = os.path.join(tmp, f"f{i}.tif")
        Image.fromarray(img).save(p, compression="tiff_lzw")
        paths.append(p)
        imgs.append(img)
    mb = sum(os.path.getsize(p) for p in paths) / 1e6
    dev = ingest.decode_tiff_batch_on_device(paths, eng)            # warm-up + correctness
    torch.cuda.synchronize()
    n = 3000 * 4000 * 3
    assert all(np.array_equal(dev.data[i, :n].cpu().numpy().reshape(3000, 4000, 3), imgs[i]) for i in (0, N - 1))
    tm = {}
    t = time.perf_counter()
    for _ in range(3):
        dev = ingest.decode_tiff_batch_on_device(paths, eng, threads=16, timings=tm)
    torch.cuda.synchronize()
    t_dev = (time.perf_counter() - t) / 3
    dst = [np.empty_like(imgs[0]) for _ in range(N)]
    with ThreadPoolExecutor(16) as pool:
        t = time.perf_counter()
        list(pool.map(lambda k: ingest.read_frame(paths[k], out=dst[k], threads=1), range(N)))
        t_host = time.perf_counter() - t
    t = time.perf_counter()
    for p_ in paths[:2]:
        np.array(Image.open(p_))
    t_pil = (time.perf_counter() - t) / 2 * N
    px = N * 12e6
    print(f"{name:10s} {N} x 12 MP, {mb:5.0f} MB of LZW: device decode {t_dev * 1e3:7.1f} ms ({px / t_dev / 1e9:5.2f} Gpix/s, "
          f"kernels alone {tm['kernel_ms']:6.1f} ms = {px * 3 / tm['kernel_ms'] / 1e6:5.1f} GB/s decoded over {tm['strips']} strips) | host decoder, 16 threads {t_host * 1e3:7.1f} ms ({px / t_host / 1e9:5.2f} Gpix/s) | "
          f"Pillow, 1 thread {t_pil * 1e3:7.0f} ms", flush=True)
    for p in paths:
        os.remove(p)
os.rmdir(tmp)
