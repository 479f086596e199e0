"""Measurement of the device-side TIFF decoders (run on the GPU box): LZW and Deflate strips, kernel time through
CUDA events, next to the host readers on all cores.  Round-2 log with the variants that were removed afterwards
(LZW ring variant, device PNG): profiles/r02_device_decode.log."""
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
tmp = tempfile.mkdtemp(prefix="lars_dec_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
cores = os.cpu_count() or 1


def host_ms(paths, n):
    with ThreadPoolExecutor(min(cores, 16)) as pool:
        t = time.perf_counter()
        list(pool.map(lambda k: ingest.read_frame(paths[k % len(paths)], threads=1), range(n)))
        return (time.perf_counter() - t) * 1e3


big = synth.vegetation_frame(100, 3000, 4000)
blocky = np.ascontiguousarray(np.repeat(np.repeat(big[::4, ::4], 4, 0), 4, 1))
for label, codec in (("LZW", "tiff_lzw"), ("Deflate", "tiff_adobe_deflate")):
    for name, img in (("noise", big), ("4x4 blocks", blocky)):
        p = os.path.join(tmp, "f.tif")
        Image.fromarray(img).save(p, compression=codec)
        tm = {}
        for _ in range(3):
            dev = ingest.decode_tiff_batch_on_device([p] * 16, eng, threads=16, timings=tm)
        torch.cuda.synchronize()
        ok = np.array_equal(dev.data[15, :img.size].cpu().numpy().reshape(img.shape), img)
        print(f"{label:7s} {name:10s} 16 x 12 MP, "
              f"{tm['strips']} strips: kernels {tm['kernel_ms']:7.1f} ms = {16 * img.size / tm['kernel_ms'] / 1e6:5.1f} GB/s decoded, "
              f"correct {ok} | host readers, {min(cores, 16)} threads: {host_ms([p], 16):7.1f} ms", flush=True)
