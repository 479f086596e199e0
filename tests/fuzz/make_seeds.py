"""Seed files for the sanitizer fuzz of the native readers: every TIFF layout the reader accepts (codec x strips /
tiles x classic / BigTIFF x byte order x chunky / planar x uint8 / uint16 / float32) and hand-built PNGs."""
import itertools
import os
import struct
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lars_image_processing_b200 import ingest  # noqa: E402

out = sys.argv[1]
os.makedirs(out, exist_ok=True)
rng = np.random.default_rng(5)
k = 0
for dtype in (np.uint8, np.uint16, np.float32):
    img = rng.integers(0, 256, (19, 23, 3))
    img = (img * (1 if dtype == np.uint8 else 200)).astype(dtype) if dtype != np.float32 else rng.uniform(-1, 1, (19, 23, 3)).astype(np.float32)
    for comp, tile, big, be, planar in itertools.product((None, "deflate", "lzw", "packbits"), (None, (16, 16)), (False, True),
                                                         (False, True), (False, True)):
        pred = dtype != np.float32 and comp in ("lzw", "deflate") and k % 2 == 0
        ingest.write_tiff(os.path.join(out, f"t{k}.tif"), img, big_endian=be, compression=comp, predictor=pred, tile=tile,
                          bigtiff=big, rows_per_strip=None if tile else 5, planar=planar)
        k += 1


def png_bytes(img, filters, idat):
    h, w = img.shape[:2]
    ch = 1 if img.ndim == 2 else img.shape[2]
    bits = img.dtype.itemsize * 8
    rb = w * ch * bits // 8
    rows = np.frombuffer(img.astype(img.dtype.newbyteorder(">")).tobytes(), np.uint8).reshape(h, rb)
    raw = b"".join(bytes([0]) + rows[r].tobytes() for r in range(h))       # filter 0 rows; the fuzzer flips the bytes anyway
    z = zlib.compress(raw, 6)
    chunk = lambda t, d: struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))
    body = b"".join(chunk(b"IDAT", z[a:a + idat]) for a in range(0, len(z), idat))
    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, bits, {1: 0, 3: 2, 4: 6}[ch], 0, 0, 0)) + body
            + chunk(b"IEND", b""))


n = 0
for shape, dtype in (((19, 23, 3), np.uint8), ((19, 23), np.uint16), ((9, 14, 4), np.uint16), ((30, 1, 3), np.uint8)):
    img = rng.integers(0, np.iinfo(dtype).max + 1, shape).astype(dtype)
    for idat in (40, 1 << 20):
        with open(os.path.join(out, f"p{n}.png"), "wb") as fh:
            fh.write(png_bytes(img, [0], idat))
        n += 1
print(k, "TIFF seeds,", n, "PNG seeds in", out)
