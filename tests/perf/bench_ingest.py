"""Ingest measurement: files -> decode threads -> pinned ring -> GPU path,
statistics back.  BASELINE config 5 shape (1280x960 uint8) as uncompressed TIFF and as PNG, and
config 3 shape (5472x3648 uint16 TIFF, which Pillow cannot deliver), next to the reference's serial
loop (np.array(Image.open(f)) + the NumPy port of its per-frame path) on a sample of the same files."""
import os, shutil, sys, tempfile, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from PIL import Image
from oracle import oracle_np as o, synth
from lars_image_processing_b200 import ingest
from lars_image_processing_b200.engine import get_engine

eng = get_engine()
tmp = tempfile.mkdtemp(prefix="lars_ingest_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)


def make_files(n, h, w, dtype, ext, distinct=8, **tiff_kw):
    base = [synth.vegetation_frame(7000 + i, h, w, dtype) for i in range(min(n, distinct))]
    paths = []
    for i in range(n):
        p = os.path.join(tmp, f"{ext}_{np.dtype(dtype).name}_{i}.{ext}")
        if i >= len(base):
            shutil.copyfile(paths[i % len(base)], p)                                    # encode each distinct frame once
        elif ext == "png":
            Image.fromarray(base[i]).save(p, compress_level=1)
        elif ext == "lzw.tif":
            Image.fromarray(base[i]).save(p, compression="tiff_lzw")                    # libtiff-written
        else:
            ingest.write_tiff(p, base[i], **(tiff_kw or dict(rows_per_strip=64)))
        paths.append(p)
    return paths, base


def run(label, paths, h, w, dtype, chunk, threads, cpu_sample):
    pipe = ingest.SurveyPipeline(h, w, dtype=dtype, chunk=chunk, depth=3, decode_threads=threads, engine=eng)
    pipe.run(paths[:chunk * 2])                       # warm-up: pools, plans, page cache
    t0 = time.perf_counter()
    out = pipe.run(paths)
    dt = time.perf_counter() - t0
    npx = len(paths) * h * w
    file_mb = sum(os.path.getsize(p) for p in paths) / 1e6
    line = (f"{label:34s} {len(paths):4d} files {file_mb:8.0f} MB  {dt * 1e3:8.1f} ms  {len(paths) / dt:8.1f} frames/s  "
            f"{npx / dt / 1e6:9.1f} Mpix/s  {file_mb / dt / 1e3:6.2f} GB/s of file bytes")
    if cpu_sample:
        t0 = time.perf_counter()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for p in paths[:cpu_sample]:
                o.reference_cpu_path(np.array(Image.open(p)), colormap=False)
        cdt = (time.perf_counter() - t0) / cpu_sample
        line += f"  | reference loop (Pillow + NumPy, 1 core): {h * w / cdt / 1e6:6.2f} Mpix/s"
    assert out["frames"] == len(paths) and out["dataset"]["NDVI"]["count"] == npx
    print(line, flush=True)


p8, _ = make_files(512, 960, 1280, np.uint8, "tif")
run("C5 shape, uncompressed TIFF", p8, 960, 1280, np.uint8, 64, 8, 4)
run("C5 shape, uncompressed TIFF, 16 thr", p8, 960, 1280, np.uint8, 64, 16, 0)
pp, _ = make_files(128, 960, 1280, np.uint8, "png")
run("C5 shape, PNG (native decode)", pp, 960, 1280, np.uint8, 32, 16, 4)
pl, _ = make_files(128, 960, 1280, np.uint8, "lzw.tif")
run("C5 shape, LZW TIFF (native decode)", pl, 960, 1280, np.uint8, 32, 16, 4)
p16, _ = make_files(16, 3648, 5472, np.uint16, "tif")
run("C3 shape, 16-bit TIFF", p16, 3648, 5472, np.uint16, 4, 8, 0)
p16z, _ = make_files(16, 3648, 5472, np.uint16, "z.tif", distinct=2, tile=(256, 256), compression="deflate", predictor=True)
run("C3 shape, 16-bit Deflate tiles", p16z, 3648, 5472, np.uint16, 4, 16, 0)
for p in p8 + pp + pl + p16 + p16z:
    os.remove(p)
os.rmdir(tmp)
