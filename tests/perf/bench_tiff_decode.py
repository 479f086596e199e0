"""Host-side decode measurement of the native TIFF reader next to Pillow / libtiff:
one BASELINE config 2 frame (4000x3000 uint8) per codec and thread count, one config 3 frame (5472x3648
uint16, which Pillow cannot deliver) as Deflate + predictor tiles, and one rank's share of a tiled mosaic read
as a region.  No GPU work; run on the GPU box because this container has about one host core."""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from PIL import Image
from oracle import synth
from lars_image_processing_b200 import ingest

tmp = tempfile.mkdtemp(prefix="lars_tiff_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
p = os.path.join(tmp, "a.tif")
THREADS = (1, 4, 8, 16)


def best(fn, reps=3):
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t)
    return min(ts)


print(f"host cores: {os.cpu_count()}")
noise = synth.vegetation_frame(1, 3000, 4000)
blocky = np.ascontiguousarray(np.repeat(np.repeat(noise[::4, ::4], 4, 0), 4, 1))
for name, img in (("noise", noise), ("4x4 blocks", blocky)):
    for codec in ("tiff_lzw", "tiff_adobe_deflate", "packbits", None):
        Image.fromarray(img).save(p, **({} if codec is None else {"compression": codec}))
        mb = os.path.getsize(p) / 1e6
        t_pil = best(lambda: np.array(Image.open(p)))
        dst = np.empty_like(img)
        ts = [best(lambda th=th: ingest.read_frame(p, out=dst, threads=th)) for th in THREADS]
        assert np.array_equal(dst, img)
        print(f"12 MP u8 {name:10s} {str(codec):18s} {mb:6.1f} MB  Pillow {t_pil * 1e3:6.1f} ms | native "
              + " / ".join(f"{t * 1e3:6.1f}" for t in ts) + f" ms at {THREADS} threads "
              f"({img.size / min(ts) / 1e9:5.2f} GB/s decoded)", flush=True)

img16 = synth.vegetation_frame(2, 3648, 5472, np.uint16)
for kw, label in ((dict(rows_per_strip=64), "uncompressed strips"),
                  (dict(tile=(256, 256), compression="deflate", predictor=True), "Deflate+predictor 256x256 tiles")):
    ingest.write_tiff(p, img16, **kw)
    mb = os.path.getsize(p) / 1e6
    dst = np.empty_like(img16)
    ts = [best(lambda th=th: ingest.read_frame(p, out=dst, threads=th)) for th in THREADS]
    assert np.array_equal(dst, img16)
    print(f"20 MP u16 RGB {label:32s} {mb:6.1f} MB | native " + " / ".join(f"{t * 1e3:6.1f}" for t in ts)
          + f" ms at {THREADS} threads ({img16.nbytes / min(ts) / 1e9:5.2f} GB/s decoded)", flush=True)

# one of 8 ranks' share (8 tiles of 1024x1024) of an 8192x8192 mosaic stored as 512x512 Deflate tiles
mosaic = np.ascontiguousarray(np.tile(blocky[:2048, :2048], (4, 4, 1)))
ingest.write_tiff(p, mosaic, tile=(512, 512), compression="deflate", predictor=True)
t_all = best(lambda: ingest.read_frame(p, threads=16), reps=2)
t_mine = best(lambda: ingest.read_mosaic_tiles(p, 1024, 1024, 3, 8, threads=16), reps=2)
tiles, origins = ingest.read_mosaic_tiles(p, 1024, 1024, 3, 8, threads=16)
assert all(np.array_equal(t, mosaic[r:r + 1024, c:c + 1024]) for t, (r, c) in zip(tiles, origins))
print(f"8192x8192 u8 mosaic, {os.path.getsize(p) / 1e6:.0f} MB file: whole image {t_all * 1e3:.0f} ms, "
      f"rank 3 of 8 (8 tiles of 1024x1024) {t_mine * 1e3:.0f} ms")
os.remove(p)
os.rmdir(tmp)
