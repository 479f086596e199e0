"""Round-2 measurement of the device-side decoders (run on the GPU box):
    python tests/perf/bench_device_decode.py                       # LZW variant 1
    LARS_LZW_VARIANT=2 python tests/perf/bench_device_decode.py    # LZW variant 2 (input / output rings)
    LARS_EXPERIMENTAL_DEVICE_INFLATE=1 python tests/perf/bench_device_decode.py   # + Deflate TIFF and PNG
Kernel time through CUDA events (TIFF) / wall clock around a synchronised call (PNG), next to the host readers on all cores."""
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
codecs = [("LZW", "tiff_lzw")] + ([("Deflate", "tiff_adobe_deflate")] if ingest.EXPERIMENTAL_DEVICE_INFLATE else [])
for label, codec in codecs:
    for name, img in (("noise", big), ("4x4 blocks", blocky)):
        p = os.path.join(tmp, "f.tif")
        Image.fromarray(img).save(p, compression=codec)
        tm = {}
        for _ in range(3):
            dev = ingest.decode_tiff_batch_on_device([p] * 16, eng, threads=16, timings=tm)
        torch.cuda.synchronize()
        ok = np.array_equal(dev.data[15, :img.size].cpu().numpy().reshape(img.shape), img)
        print(f"{label:7s} variant {os.environ.get('LARS_LZW_VARIANT', '1') if label == 'LZW' else '-'} {name:10s} 16 x 12 MP, "
              f"{tm['strips']} strips: kernels {tm['kernel_ms']:7.1f} ms = {16 * img.size / tm['kernel_ms'] / 1e6:5.1f} GB/s decoded, "
              f"correct {ok} | host readers, {min(cores, 16)} threads: {host_ms([p], 16):7.1f} ms", flush=True)
if ingest.EXPERIMENTAL_DEVICE_INFLATE:
    small = [synth.vegetation_frame(200 + i, 960, 1280) for i in range(8)]
    paths = []
    for i, img in enumerate(small):
        p = os.path.join(tmp, f"s{i}.png")
        Image.fromarray(img).save(p, compress_level=1)
        paths.append(p)
    for n in (64, 256, 1024):
        batch = [paths[i % 8] for i in range(n)]
        ingest.decode_png_batch_on_device(batch[:8], eng)
        torch.cuda.synchronize()
        t = time.perf_counter()
        dev = ingest.decode_png_batch_on_device(batch, eng)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t) * 1e3
        ok = np.array_equal(dev.data[n - 1, :small[0].size].cpu().numpy().reshape(small[0].shape), small[(n - 1) % 8])
        print(f"PNG     {n:5d} x 1280x960: device path {dt:8.1f} ms ({n * 1.2288 / dt:6.2f} Gpix/s), correct {ok} | "
              f"host readers, {min(cores, 16)} threads: {host_ms(paths, n):8.1f} ms", flush=True)
