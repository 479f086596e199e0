"""Ingest in front of the path and the streaming many-frame runner (SURVEY.md section 8(f) rank 4,
BASELINE configs 3 and 5).

The reference loads every frame with ``np.array(PIL.Image.open(...))`` (process-images.py:183-193;
backend-process.py:52; process-ndvi.py:18; process-rgn.py:18) inside a serial per-file loop
(backend-process.py:92-97; process-images.py:633-663).  Here:

* :func:`read_frame` keeps that behaviour for every format Pillow decodes, and adds a native
  TIFF reader (``lars_tiff_probe`` / ``lars_tiff_read_region``, host side of the C ABI: strips or
  tiles, uncompressed / LZW / Deflate / PackBits, predictor, BigTIFF) that moves the chunks of a
  memory-mapped file straight into a pinned buffer, compressed chunks decoded by several host
  threads -- including 16-bit RGB TIFFs, which Pillow opens as 8-bit;
* PNG frames (BASELINE config 1; gray / RGB / RGBA, 8- or 16-bit, non-interlaced) decode natively as well
  (``lars_png_probe`` / ``lars_png_read``: one C call per frame, so decode threads do not queue on the
  interpreter lock as they do inside Pillow);
* :func:`read_region` / :func:`read_mosaic_tiles` read only the strips / tiles that touch a
  rectangle: every rank of a tile-sharded orthomosaic (BASELINE config 4) reads its own tiles;
* :class:`SurveyPipeline` streams any number of equally-shaped frames through the GPU path:
  host threads decode into a ring of pinned chunk buffers, one stream copies H2D, one runs
  Pass 1 + LUT + Pass 2, one copies results back, and the dataset-wide statistics are folded on
  the device chunk by chunk (``lars_stats_merge``), with the usual single all-gather across ranks
  at the end.  Nothing in here computes on the CPU: without the CUDA library it raises.
"""
from __future__ import annotations

import ctypes as C
import io
import mmap
import os
import queue
import struct
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib
from ._lib import INDEX_STATS_DTYPE, LarsError, check
from .engine import ALL_OUTPUTS, DeviceFrames, DeviceOutputs, Engine, get_engine, stats_records_to_dicts

Source = Union[str, os.PathLike, bytes, bytearray, memoryview, np.ndarray]

LARS_ERR_UNSUPPORTED = -3


# ------------------------------------------------------------------------------------------
# files -> arrays
# ------------------------------------------------------------------------------------------
def _tiff_probe(buf) -> Optional[_lib.TiffInfo]:
    """TiffInfo if the native reader handles this buffer, None if Pillow has to; raises on a
    corrupt TIFF."""
    lib = _lib.load()
    info = _lib.TiffInfo()
    view = np.frombuffer(buf, dtype=np.uint8)
    try:        # no view of a memory-mapped file may outlive this call (closing the map would fail)
        if view.size < 4 or bytes(view[:2]) not in (b"II", b"MM"):
            return None
        rc = lib.lars_tiff_probe(view.ctypes.data, view.size, C.byref(info))
    finally:
        del view
    if rc == LARS_ERR_UNSUPPORTED:
        return None
    check(rc, "lars_tiff_probe")
    return info


PNG_SIGNATURE = b"\x89PNG\r\n\x1a\n"


def _png_probe(buf) -> Optional[_lib.PngInfo]:
    """PngInfo if the native reader handles this buffer, None if Pillow has to; raises on a corrupt PNG."""
    view = np.frombuffer(buf, dtype=np.uint8)
    try:
        if view.size < 8 or bytes(view[:8]) != PNG_SIGNATURE:
            return None
        info = _lib.PngInfo()
        rc = _lib.load().lars_png_probe(view.ctypes.data, view.size, C.byref(info))
    finally:
        del view
    if rc == LARS_ERR_UNSUPPORTED:
        return None
    check(rc, "lars_png_probe")
    return info


def _png_shape(info) -> tuple:
    return (info.height, info.width) if info.channels == 1 else (info.height, info.width, info.channels)


def _tiff_dtype(info) -> np.dtype:
    return np.dtype({8: np.uint8, 16: np.uint16, 32: np.float32}[info.bits_per_sample])


def _tiff_shape(info) -> tuple:
    spp = info.samples_per_pixel
    return (info.height, info.width) if spp == 1 else (info.height, info.width, spp)


def default_decode_threads() -> int:
    return max(1, min(8, os.cpu_count() or 1))


def read_frame(source: Source, out: Optional[np.ndarray] = None, threads: Optional[int] = None) -> np.ndarray:
    """One frame as the array ``np.array(Image.open(source))`` would give -- except that 16-bit
    TIFFs keep their 16 bits.  ``source``: path, encoded bytes, or an array (returned as is).
    ``out``: optional destination (e.g. a row of a pinned buffer) of the right size and dtype.
    ``threads``: host threads decoding the strips / tiles of a compressed TIFF (default: up to 8)."""
    if isinstance(source, np.ndarray):
        if out is not None:
            np.copyto(out.reshape(source.shape), source)
            return out.reshape(source.shape)
        return source
    if isinstance(source, (bytes, bytearray, memoryview)):
        return _decode_buffer(source, out, threads)
    with open(os.fspath(source), "rb") as fh:
        size = os.fstat(fh.fileno()).st_size
        if size == 0:
            raise ValueError(f"{source}: empty file")
        with mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ) as mm:
            return _decode_buffer(mm, out, threads)


def _tiff_read_into(buf, info, region, dst: np.ndarray, threads: Optional[int]) -> None:
    r0, r1, c0, c1 = region
    view = np.frombuffer(buf, dtype=np.uint8)
    try:        # no view of a memory-mapped file may outlive this call (closing the map would fail)
        rc = _lib.load().lars_tiff_read_region(view.ctypes.data, view.size, C.byref(info), r0, r1, c0, c1,
                                               dst.ctypes.data, dst.nbytes,
                                               default_decode_threads() if threads is None else int(threads))
    finally:
        del view
    check(rc, "lars_tiff_read_region")


def _decode_buffer(buf, out: Optional[np.ndarray], threads: Optional[int] = None) -> np.ndarray:
    info = _tiff_probe(buf)
    if info is not None:
        dtype = _tiff_dtype(info)
        shape = _tiff_shape(info)
        dst = out if out is not None else np.empty(shape, dtype)
        if dst.dtype != dtype or dst.size != int(np.prod(shape)) or not dst.flags.c_contiguous:
            raise ValueError(f"destination must be a contiguous {np.dtype(dtype).name} array of {shape}")
        _tiff_read_into(buf, info, (0, info.height, 0, info.width), dst, threads)
        return dst.reshape(shape)
    png = _png_probe(buf)
    if png is not None:
        dtype = np.uint8 if png.bit_depth == 8 else np.uint16
        shape = _png_shape(png)
        dst = out if out is not None else np.empty(shape, dtype)
        if dst.dtype != dtype or dst.size != int(np.prod(shape)) or not dst.flags.c_contiguous:
            raise ValueError(f"destination must be a contiguous {np.dtype(dtype).name} array of {shape}")
        view = np.frombuffer(buf, dtype=np.uint8)
        try:
            rc = _lib.load().lars_png_read(view.ctypes.data, view.size, C.byref(png), dst.ctypes.data, dst.nbytes)
        finally:
            del view
        check(rc, "lars_png_read")
        return dst.reshape(shape)
    from PIL import Image
    data = buf if isinstance(buf, (bytes, bytearray)) else bytes(buf)
    arr = np.array(Image.open(io.BytesIO(data)))              # process-images.py:183-193
    if out is not None:
        if out.dtype != arr.dtype or out.size != arr.size:
            raise ValueError(f"destination must hold {arr.shape} {arr.dtype}, got {out.shape} {out.dtype}")
        np.copyto(out.reshape(arr.shape), arr)
        return out.reshape(arr.shape)
    return arr


def read_region(source: Source, rows: Sequence[int], cols: Optional[Sequence[int]] = None,
                out: Optional[np.ndarray] = None, threads: Optional[int] = None) -> np.ndarray:
    """Rows ``[rows[0], rows[1])`` x columns ``[cols[0], cols[1])`` (default: every column) of a frame.
    For TIFF files only the strips / tiles that touch the rectangle are read and decoded, so a rank can
    fetch its share of an orthomosaic far larger than its host memory; every other format is decoded
    as a whole (Pillow) and cropped."""
    if isinstance(source, np.ndarray):
        shape = source.shape
    elif isinstance(source, (bytes, bytearray, memoryview)):
        return _region_of_buffer(source, rows, cols, out, threads)
    else:
        with open(os.fspath(source), "rb") as fh, mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ) as mm:
            return _region_of_buffer(mm, rows, cols, out, threads)
    r0, r1, c0, c1 = _check_region(shape, rows, cols)
    return _crop_into(source, (r0, r1, c0, c1), out)


def _check_region(shape, rows, cols) -> tuple:
    h, w = shape[:2]
    r0, r1 = int(rows[0]), int(rows[1])
    c0, c1 = (0, w) if cols is None else (int(cols[0]), int(cols[1]))
    if not (0 <= r0 < r1 <= h and 0 <= c0 < c1 <= w):
        raise ValueError(f"region rows [{r0}, {r1}) x columns [{c0}, {c1}) is empty or outside the {h} x {w} frame")
    return r0, r1, c0, c1


def _crop_into(arr: np.ndarray, region, out: Optional[np.ndarray]) -> np.ndarray:
    r0, r1, c0, c1 = region
    crop = arr[r0:r1, c0:c1]
    if out is None:
        return np.ascontiguousarray(crop)
    if out.dtype != crop.dtype or out.size != crop.size:
        raise ValueError(f"destination must hold {crop.shape} {crop.dtype}, got {out.shape} {out.dtype}")
    np.copyto(out.reshape(crop.shape), crop)
    return out.reshape(crop.shape)


def _region_of_buffer(buf, rows, cols, out, threads) -> np.ndarray:
    info = _tiff_probe(buf)
    if info is None:
        whole = _decode_buffer(buf, None)
        return _crop_into(whole, _check_region(whole.shape, rows, cols), out)
    dtype = _tiff_dtype(info)
    full = _tiff_shape(info)
    r0, r1, c0, c1 = _check_region(full, rows, cols)
    shape = (r1 - r0, c1 - c0) + tuple(full[2:])
    dst = out if out is not None else np.empty(shape, dtype)
    if dst.dtype != dtype or dst.size != int(np.prod(shape)) or not dst.flags.c_contiguous:
        raise ValueError(f"destination must be a contiguous {np.dtype(dtype).name} array of {shape}")
    _tiff_read_into(buf, info, (r0, r1, c0, c1), dst, threads)
    return dst.reshape(shape)


def mosaic_tile_grid(height: int, width: int, tile_h: int, tile_w: int) -> List[tuple]:
    """Row-major (row0, col0) origins of the equally-sized tiles of a mosaic.  The device batch holds
    equally-sized frames and the white-balance percentiles count every pixel exactly once, so the tile
    size has to divide the image size (pick a divisor, e.g. with :func:`largest_divisor`)."""
    if tile_h < 1 or tile_w < 1 or height % tile_h or width % tile_w:
        raise ValueError(f"{tile_h} x {tile_w} tiles do not divide a {height} x {width} mosaic")
    return [(r, c) for r in range(0, height, tile_h) for c in range(0, width, tile_w)]


def largest_divisor(n: int, at_most: int) -> int:
    """Largest divisor of ``n`` that is <= ``at_most`` (a tile edge for :func:`mosaic_tile_grid`)."""
    for d in range(max(1, min(n, int(at_most))), 0, -1):
        if n % d == 0:
            return d
    return 1


def read_mosaic_band(source: Source, rank: int = 0, world: int = 1, out: Optional[np.ndarray] = None,
                     threads: Optional[int] = None):
    """The full-width row band rank ``rank`` of ``world`` owns of one huge image: ``(band, row0)``.
    Nothing in the path is two-dimensional, so a band of any height is simply one frame of
    ``rows * width`` pixels -- this is the sharding for mosaics whose sizes have no convenient divisor
    (``Engine.upload([band])`` then ``distributed.process_mosaic_tiles``: the bands of different ranks
    may differ in height, the white-balance counters are all-reduced as for tiles)."""
    from .distributed import shard_range
    shape, _ = frame_info(source)
    if world > shape[0]:
        raise ValueError(f"{world} ranks for a mosaic of {shape[0]} rows")
    r0, r1 = shard_range(shape[0], rank, world)
    return read_region(source, (r0, r1), None, out=out, threads=threads), r0


def read_mosaic_tiles(source: Source, tile_h: int, tile_w: int, rank: int = 0, world: int = 1,
                      out: Optional[np.ndarray] = None, threads: Optional[int] = None):
    """The tiles rank ``rank`` of ``world`` owns of one huge image (contiguous block of the row-major
    tile grid, :func:`..distributed.shard_range`), read straight from the file: ``(tiles, origins)`` with
    ``tiles`` of shape [n, tile_h, tile_w(, C)] -- what ``Engine.upload`` and
    ``distributed.process_mosaic_tiles`` take -- and ``origins`` the (row0, col0) of each.  ``out``:
    optional (pinned) destination of that shape."""
    from .distributed import shard_range
    shape, dtype = frame_info(source)
    grid = mosaic_tile_grid(shape[0], shape[1], tile_h, tile_w)
    a, b = shard_range(len(grid), rank, world)
    mine = grid[a:b]
    tshape = (len(mine), tile_h, tile_w) + tuple(shape[2:])
    tiles = out if out is not None else np.empty(tshape, dtype)
    if tiles.dtype != dtype or tiles.size != int(np.prod(tshape)) or not tiles.flags.c_contiguous:
        raise ValueError(f"destination must be a contiguous {np.dtype(dtype).name} array of {tshape}")
    tiles = tiles.reshape(tshape)
    if isinstance(source, (np.ndarray, bytes, bytearray, memoryview)):
        for k, (r, c) in enumerate(mine):
            read_region(source, (r, r + tile_h), (c, c + tile_w), out=tiles[k], threads=threads)
    elif mine:
        with open(os.fspath(source), "rb") as fh, mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ) as mm:
            if _is_tiff(mm):
                for k, (r, c) in enumerate(mine):
                    _region_of_buffer(mm, (r, r + tile_h), (c, c + tile_w), tiles[k], threads)
            else:                                   # decoded once, cropped per tile
                whole = _decode_buffer(mm, None)
                for k, (r, c) in enumerate(mine):
                    _crop_into(whole, (r, r + tile_h, c, c + tile_w), tiles[k])
    return tiles, mine


def frame_info(source: Source) -> tuple:
    """(shape, dtype) of a frame without decoding the pixels where the format allows it."""
    if isinstance(source, np.ndarray):
        return tuple(source.shape), source.dtype
    if not isinstance(source, (bytes, bytearray, memoryview)):
        with open(os.fspath(source), "rb") as fh, mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ) as mm:
            return frame_info(memoryview(mm)) if (_is_tiff(mm) or bytes(mm[:8]) == PNG_SIGNATURE) else _info_via_pillow(bytes(mm))
    info = _tiff_probe(source)
    if info is not None:
        return _tiff_shape(info), _tiff_dtype(info)
    png = _png_probe(source)
    if png is not None:
        return _png_shape(png), np.dtype(np.uint8 if png.bit_depth == 8 else np.uint16)
    return _info_via_pillow(source)


def _is_tiff(buf) -> bool:
    return len(buf) >= 4 and bytes(buf[:2]) in (b"II", b"MM")


def _info_via_pillow(buf) -> tuple:
    arr = _decode_buffer(buf, None)
    return tuple(arr.shape), arr.dtype


TIFF_COMPRESSION = {None: 1, "none": 1, "lzw": 5, "deflate": 8, "packbits": 32773}


def _packbits_encode(data: bytes) -> bytes:
    """TIFF 6.0 section 9, literal runs only (valid PackBits; a writer needs no more)."""
    out = bytearray()
    for a in range(0, len(data), 128):
        piece = data[a:a + 128]
        out.append(len(piece) - 1)
        out += piece
    return bytes(out)


def _lzw_encode(data: bytes) -> bytes:
    """TIFF 6.0 section 13 (MSB-first 9..12-bit codes, early width change).  Plain Python: meant for the
    modest files a test or an export writes, not for bulk storage -- use "deflate" there."""
    out = bytearray()
    acc = nb = 0

    def put(code, width):
        nonlocal acc, nb
        acc = (acc << width) | code
        nb += width
        while nb >= 8:
            nb -= 8
            out.append((acc >> nb) & 0xFF)
        acc &= (1 << nb) - 1

    table = {bytes([i]): i for i in range(256)}
    next_code, width = 258, 9
    put(256, width)
    w = b""
    for byte in data:
        wc = w + bytes([byte])
        if wc in table:
            w = wc
            continue
        put(table[w], width)
        table[wc] = next_code
        next_code += 1
        if next_code == 4094:                       # table full: emit Clear and start over
            put(256, width)
            table = {bytes([i]): i for i in range(256)}
            next_code, width = 258, 9
        elif next_code == (1 << width):             # the decoder, one entry behind, switches "one code early"
            width += 1
        w = bytes([byte])
    if w:
        put(table[w], width)
        next_code += 1                              # the decoder adds an entry for this code, too
        if next_code == 4094:
            put(256, width)
            width = 9
        elif next_code == (1 << width):
            width += 1
    put(257, width)
    if nb:
        out.append((acc << (8 - nb)) & 0xFF)
    return bytes(out)


def write_tiff(path, array: np.ndarray, big_endian: bool = False, rows_per_strip: Optional[int] = None,
               compression: Optional[str] = None, predictor: bool = False, tile: Optional[Sequence[int]] = None,
               bigtiff: bool = False, planar: bool = False) -> None:
    """TIFF writer for HxW / HxWx3 / HxWx4 uint8 / uint16 frames and float32 maps or reflectance stacks.  Pillow cannot write 16-bit
    RGB; survey frames of BASELINE config 3 and mosaics of config 4 are stored with this.
    ``compression``: None, "deflate", "lzw" or "packbits"; ``predictor``: horizontal differencing in
    front of the compressor; ``tile``: (tile_length, tile_width) for a tiled layout instead of strips;
    ``bigtiff``: 64-bit offsets (files beyond 4 GB); ``planar``: PlanarConfiguration 2, one run of strips / tiles
    per sample of the pixel (band-interleaved, as GDAL writes with INTERLEAVE=BAND)."""
    import zlib
    a = np.ascontiguousarray(array)
    if a.dtype not in (np.uint8, np.uint16, np.float32) or a.ndim not in (2, 3):
        raise ValueError("write_tiff needs a uint8 / uint16 / float32 HxW or HxWxC array")
    if a.dtype == np.float32 and predictor:
        raise ValueError("the differencing predictor is for integer samples")
    if compression not in TIFF_COMPRESSION:
        raise ValueError(f"compression must be one of {sorted(k for k in TIFF_COMPRESSION if k)} or None")
    comp = TIFF_COMPRESSION[compression]
    if predictor and comp not in (5, 8):
        raise ValueError("the differencing predictor belongs to the LZW / Deflate codecs (libtiff ignores it elsewhere)")
    h, w = a.shape[:2]
    spp = 1 if a.ndim == 2 else a.shape[2]
    a3 = a.reshape(h, w, spp)
    bits = a.dtype.itemsize * 8
    e = ">" if big_endian else "<"
    file_dtype = a.dtype.newbyteorder(e)

    def encode(block: np.ndarray) -> bytes:
        """[rows, cols, spp] samples -> the bytes of one strip / tile."""
        if predictor:
            d = block.copy()
            d[:, 1:] = block[:, 1:] - block[:, :-1]          # modulo 2^bits, per sample (TIFF 6.0 section 14)
            block = d
        raw = block.astype(file_dtype).tobytes()
        if comp == 8:
            return zlib.compress(raw, 6)
        if comp == 5:
            return _lzw_encode(raw)
        if comp == 32773:
            return _packbits_encode(raw)
        return raw

    planar = bool(planar) and spp > 1
    layers = [a3[:, :, k:k + 1] for k in range(spp)] if planar else [a3]
    chunks = []
    if tile is not None:
        tl, tw = int(tile[0]), int(tile[1])
        if tl < 1 or tw < 1:
            raise ValueError("tile sizes must be positive")
        for layer in layers:
            for r in range(0, h, tl):
                for c in range(0, w, tw):
                    block = np.zeros((tl, tw, layer.shape[2]), a.dtype)
                    part = layer[r:r + tl, c:c + tw]
                    block[:part.shape[0], :part.shape[1]] = part
                    chunks.append(encode(block))
        rps = tl
    else:
        rps = h if not rows_per_strip else max(1, min(h, int(rows_per_strip)))
        for layer in layers:
            chunks += [encode(layer[r:r + rps]) for r in range(0, h, rps)]
    sizes = [len(ch) for ch in chunks]

    off_fmt, off_type = ("Q", 16) if bigtiff else ("I", 4)
    header_len, entry_len, inline = (16, 20, 8) if bigtiff else (8, 12, 4)
    entries: List[bytes] = []
    extra = b""

    def layout(offsets):
        nonlocal entries, extra
        entries, extra = [], b""
        tags = [(256, 4, [w]), (257, 4, [h]), (258, 3, [bits] * spp), (259, 3, [comp]),
                (262, 3, [2 if spp >= 3 else 1]), (277, 3, [spp]), (284, 3, [2 if planar else 1])]
        if tile is None:
            tags += [(273, off_type, offsets), (278, 4, [rps]), (279, off_type, sizes)]
        else:
            tags += [(322, 4, [tw]), (323, 4, [tl]), (324, off_type, offsets), (325, off_type, sizes)]
        if predictor:
            tags.append((317, 3, [2]))
        if a.dtype == np.float32:
            tags.append((339, 3, [3] * spp))                       # SampleFormat: IEEE floating point
        if spp == 4:
            tags.append((338, 3, [2]))
        tags.sort()
        ifd_len = (8 if bigtiff else 2) + entry_len * len(tags) + (8 if bigtiff else 4)
        extra_base = header_len + ifd_len
        for tag, typ, values in tags:
            blob = struct.pack(e + {3: "H", 4: "I", 16: "Q"}[typ] * len(values), *values)
            head = struct.pack(e + ("HHQ" if bigtiff else "HHI"), tag, typ, len(values))
            if len(blob) <= inline:
                entries.append(head + blob.ljust(inline, b"\0"))
            else:
                entries.append(head + struct.pack(e + off_fmt, extra_base + len(extra)))
                extra += blob + (b"\0" if len(blob) & 1 else b"")
        return extra_base + len(extra)

    # the chunk offsets depend on the size of the extra block: lay out twice
    data_base = layout([0] * len(chunks))
    offsets = list(np.cumsum([data_base] + sizes[:-1]).tolist())
    if layout(offsets) != data_base:
        raise AssertionError("TIFF layout did not converge")
    if not bigtiff and data_base + sum(sizes) > 0xFFFFFFFF:
        raise ValueError("the file would exceed 4 GB: pass bigtiff=True")
    with open(os.fspath(path), "wb") as fh:
        if bigtiff:
            fh.write((b"MM" if big_endian else b"II") + struct.pack(e + "HHHQ", 43, 8, 0, header_len))
            fh.write(struct.pack(e + "Q", len(entries)) + b"".join(entries) + struct.pack(e + "Q", 0))
        else:
            fh.write((b"MM" if big_endian else b"II") + struct.pack(e + "HI", 42, header_len))
            fh.write(struct.pack(e + "H", len(entries)) + b"".join(entries) + struct.pack(e + "I", 0))
        fh.write(extra)
        for ch in chunks:
            fh.write(ch)


# ------------------------------------------------------------------------------------------
# device-side decode of LZW TIFF frames
# ------------------------------------------------------------------------------------------
# LZW strips go through lars_lzw_decode_device, Deflate strips through lars_inflate_decode_device (one warp per zlib
# stream).  Both measured on B200 against the 16-thread host readers (profiles/r02_device_decode.log): LZW 88 ms
# against 250-277 ms per 16 x 12 MP of noise, Deflate 147 ms against 222 ms (37 against 64 ms on blocky content).
# PNG frames stay on the host reader: one zlib stream per image leaves the GPU with one warp per frame, and the
# device path measured 4-11x slower than 16 host threads -- it was removed.


def device_decodable(source: Source) -> bool:
    """True if :func:`decode_tiff_batch_on_device` takes this file: an LZW- or Deflate-compressed TIFF stored in
    strips of at most 1 MB decoded (what libtiff / Pillow / GDAL write by default)."""
    def probe(buf):
        info = _tiff_probe(buf)
        if info is None or info.compression not in (5, 8) or info.tile_width > 0 or info.planar_config != 1 or info.bits_per_sample > 16:
            return False
        rows = min(info.rows_per_strip, info.height)
        return rows * info.width * info.samples_per_pixel * (info.bits_per_sample // 8) <= (1 << 20)
    if isinstance(source, np.ndarray):
        return False
    try:
        if isinstance(source, (bytes, bytearray, memoryview)):
            return probe(source)
        with open(os.fspath(source), "rb") as fh, mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ) as mm:
            return probe(mm)
    except (LarsError, OSError, ValueError):
        return False


def decode_tiff_batch_on_device(sources: Sequence[Source], engine: Optional[Engine] = None, stream=None,
                                threads: Optional[int] = None, timings: Optional[dict] = None) -> DeviceFrames:
    """Equally-shaped LZW (or Deflate) TIFF frames -> a device-resident frame batch, decoded ON the GPU: the compressed
    file bytes are uploaded as they are (memory-mapped files copied by ``threads`` host threads into one pinned
    staging buffer, one H2D copy), one warp decodes each strip straight into its frame slot
    (``lars_lzw_decode_device``), a second kernel undoes the differencing predictor / big-endian samples.
    The host only parses the IFDs.  Raises ``LarsError`` for a file outside :func:`device_decodable` or a
    corrupt strip (the call synchronises once to read the strip verdict).  ``timings``: optional dict that
    receives ``kernel_ms`` (CUDA events around the two kernels) and ``h2d_bytes``."""
    eng = engine or get_engine()
    lib = eng.lib
    s = stream or eng.stream()
    if not len(sources):
        raise ValueError("no frames")
    opened = []                                   # (file handle, mmap) of path sources, closed at the end
    views: List[Optional[np.ndarray]] = []
    try:
        infos, tables = [], []
        for src in sources:
            if isinstance(src, (bytes, bytearray, memoryview)):
                raw = np.frombuffer(src, dtype=np.uint8)
            else:
                fh = open(os.fspath(src), "rb")
                opened.append(fh)
                mm = mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ)
                opened.append(mm)
                raw = np.frombuffer(mm, dtype=np.uint8)
            views.append(raw)
            info = _lib.TiffInfo()
            check(lib.lars_tiff_probe(raw.ctypes.data, raw.size, C.byref(info)), "lars_tiff_probe")
            chunks = np.zeros(info.n_strips, _lib.LZW_CHUNK_DTYPE)
            if info.compression == 8:
                check(lib.lars_tiff_deflate_chunks(raw.ctypes.data, raw.size, C.byref(info), chunks.ctypes.data, chunks.size),
                      "lars_tiff_deflate_chunks")
            else:
                check(lib.lars_tiff_lzw_chunks(raw.ctypes.data, raw.size, C.byref(info), chunks.ctypes.data, chunks.size),
                      "lars_tiff_lzw_chunks")
            infos.append(info)
            tables.append(chunks)
        first = infos[0]
        key = lambda i: (i.height, i.width, i.samples_per_pixel, i.bits_per_sample, i.predictor, i.big_endian, i.compression)
        if any(key(i) != key(first) for i in infos):
            raise ValueError("all frames of a batch must share shape, sample width, codec, predictor and byte order")
        sb = first.bits_per_sample // 8
        frames = eng.alloc_frames(len(views), first.height, first.width, first.samples_per_pixel, s, sample_bytes=sb)
        # one staging buffer: the files back to back (8-byte aligned), then the strip table of the whole batch
        starts, pos = [], 0
        for raw in views:
            starts.append(pos)
            pos += (raw.size + 7) & ~7
        n_chunks = sum(t.size for t in tables)
        table_at = pos
        total = table_at + n_chunks * _lib.LZW_CHUNK_DTYPE.itemsize
        host = torch.empty(total, dtype=torch.uint8, pin_memory=True)
        host_np = host.numpy()
        all_chunks = host_np[table_at:].view(_lib.LZW_CHUNK_DTYPE)
        k = 0
        for f, t in enumerate(tables):
            t["src_offset"] += starts[f]
            t["dst_offset"] += f * frames.stride_bytes
            all_chunks[k:k + t.size] = t
            k += t.size

        def stage(f):                             # page cache -> pinned memory; NumPy releases the GIL for the copy
            np.copyto(host_np[starts[f]:starts[f] + views[f].size], views[f])
        n_thr = default_decode_threads() if threads is None else max(1, int(threads))
        if n_thr > 1 and len(views) > 1:
            with ThreadPoolExecutor(min(n_thr, len(views))) as pool:
                list(pool.map(stage, range(len(views))))
        else:
            for f in range(len(views)):
                stage(f)
    finally:
        views.clear()                             # no view of a mapped file may outlive its map
        raw = None
        for h in reversed(opened):
            h.close()
    with torch.cuda.stream(s), torch.cuda.device(eng.device):
        dev = torch.empty(total, dtype=torch.uint8, device=eng.device)
        dev.copy_(host, non_blocking=True)
        counters = torch.zeros(2, dtype=torch.int32, device=eng.device)
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) if timings is not None else None
        if ev:
            ev[0].record(s)
        decode = lib.lars_inflate_decode_device if first.compression == 8 else lib.lars_lzw_decode_device
        check(decode(dev.data_ptr(), dev.data_ptr() + table_at, n_chunks, frames.data.data_ptr(),
                     counters.data_ptr(), s.cuda_stream), "lars_lzw_decode_device")
        check(lib.lars_tiff_post_device(frames.data.data_ptr(), frames.n_frames, frames.stride_bytes, first.height,
                                        first.width, first.samples_per_pixel, sb, first.predictor,
                                        1 if (first.big_endian and sb == 2) else 0, s.cuda_stream),
              "lars_tiff_post_device")
        if ev:
            ev[1].record(s)
        bad = int(counters[0].item())                 # synchronises the stream: the verdict of every strip
    if timings is not None:
        timings["kernel_ms"] = ev[0].elapsed_time(ev[1])
        timings["h2d_bytes"] = total
        timings["strips"] = n_chunks
    if bad:
        raise LarsError(f"{bad} compressed strip(s) of the batch are corrupt or shorter than their rows")
    return frames


def tiff_region_device_plan(raw: np.ndarray, rows: Optional[Sequence[int]] = None) -> dict:
    """Host side of decoding a row band of a (possibly tiled) LZW / Deflate TIFF on the device, no GPU involved: which
    strips / tiles touch rows ``[rows[0], rows[1])``, the ``lars_lzw_chunk`` table that decodes each of them into its
    own scratch slot, the geometry for ``lars_tiff_post_device`` (the slots are treated as frames of one chunk each) and
    the move table for ``lars_untile_device`` that places every slot's rows inside the band.  Strips are tiles as wide
    as the image.  Refuses (``LarsError``) files outside the device decoders: other codecs, planar, float, chunks
    beyond 1 MB decoded."""
    lib = _lib.load()
    info = _lib.TiffInfo()
    check(lib.lars_tiff_probe(raw.ctypes.data, raw.size, C.byref(info)), "lars_tiff_probe")
    if info.compression not in (5, 8) or info.planar_config != 1 or info.bits_per_sample > 16:
        raise LarsError("the device decoders take chunky LZW / Deflate TIFFs with 8- or 16-bit samples")
    h, w = info.height, info.width
    r0, r1 = (0, h) if rows is None else (int(rows[0]), int(rows[1]))
    if not 0 <= r0 < r1 <= h:
        raise ValueError(f"rows [{r0}, {r1}) are empty or outside the {h}-row image")
    sb = info.bits_per_sample // 8
    px = info.samples_per_pixel * sb
    tiled = info.tile_width > 0
    cw, chh = (info.tile_width, info.tile_length) if tiled else (w, info.rows_per_strip)
    across = info.tiles_across if tiled else 1
    slot_row_bytes = cw * px
    slot_bytes = chh * slot_row_bytes
    if slot_bytes > (1 << 20):
        raise LarsError("a strip / tile decodes to more than 1 MB: outside the device decoders")
    order = ">" if info.big_endian else "<"
    kinds = {3: "u2", 4: "u4", 16: "u8"}
    offs = np.frombuffer(raw, order + kinds[info.strip_offsets_type], info.n_strips, info.strip_offsets_pos).astype(np.int64)
    cnts = np.frombuffer(raw, order + kinds[info.strip_counts_type], info.n_strips, info.strip_counts_pos).astype(np.int64)
    cy0, cy1 = r0 // chh, (r1 - 1) // chh
    n = (cy1 - cy0 + 1) * across
    chunks = np.zeros(n, _lib.LZW_CHUNK_DTYPE)
    moves = np.zeros((n, 5), np.int64)
    k = 0
    for cy in range(cy0, cy1 + 1):
        valid = min(chh, h - cy * chh)
        for cx in range(across):
            idx = cy * across + cx
            need = (chh if tiled else valid) * slot_row_bytes             # tiles are stored whole, the last strip is short
            chunks[k] = (offs[idx], k * slot_bytes, cnts[idx], need)
            lo, hi = max(r0, cy * chh), min(r1, cy * chh + valid)
            cols = min(cw, w - cx * cw)
            moves[k] = (lo - cy * chh, hi - lo, lo - r0, cx * cw * px, cols * px)
            k += 1
    return {"info": info, "chunks": chunks, "moves": moves, "slot_bytes": slot_bytes, "slot_row_bytes": slot_row_bytes,
            "chunk_rows": chh, "chunk_width": cw, "rows": (r0, r1), "band_row_bytes": w * px, "sample_bytes": sb}


def decode_tiff_region_on_device(source: Source, rows: Optional[Sequence[int]] = None, engine: Optional[Engine] = None,
                                 stream=None) -> DeviceFrames:
    """A row band (default: the whole image) of an LZW / Deflate TIFF -- strips or tiles -- decoded on the GPU into a
    one-frame device batch, e.g. the band of a tiled mosaic a rank owns (``read_mosaic_band`` without the host decode).
    Only the chunks under the band are uploaded."""
    eng = engine or get_engine()
    lib = eng.lib
    s = stream or eng.stream()
    raw = np.frombuffer(source, dtype=np.uint8) if isinstance(source, (bytes, bytearray, memoryview)) \
        else np.fromfile(os.fspath(source), dtype=np.uint8)
    plan = tiff_region_device_plan(raw, rows)
    info = plan["info"]
    chunks, moves = plan["chunks"].copy(), plan["moves"]
    n = chunks.size
    # staging: the chunks' compressed bytes back to back (8-byte aligned), then the two tables
    starts = np.zeros(n, np.int64)
    pos = 0
    for k in range(n):
        starts[k] = pos
        pos += (int(chunks["src_bytes"][k]) + 7) & ~7
    table_at = pos
    moves_at = table_at + chunks.nbytes
    total = moves_at + moves.nbytes
    host = torch.empty(total, dtype=torch.uint8, pin_memory=True)
    host_np = host.numpy()
    for k in range(n):
        a, m = int(chunks["src_offset"][k]), int(chunks["src_bytes"][k])
        host_np[starts[k]:starts[k] + m] = raw[a:a + m]
    chunks["src_offset"] = starts
    host_np[table_at:moves_at] = chunks.view(np.uint8)
    host_np[moves_at:] = moves.reshape(-1).view(np.uint8)
    r0, r1 = plan["rows"]
    sb = plan["sample_bytes"]
    frames = eng.alloc_frames(1, r1 - r0, info.width, info.samples_per_pixel, s, sample_bytes=sb)
    with torch.cuda.stream(s), torch.cuda.device(eng.device):
        dev = torch.empty(total, dtype=torch.uint8, device=eng.device)
        dev.copy_(host, non_blocking=True)
        scratch = torch.empty(n * plan["slot_bytes"], dtype=torch.uint8, device=eng.device)
        counters = torch.zeros(2, dtype=torch.int32, device=eng.device)
        decode = lib.lars_inflate_decode_device if info.compression == 8 else lib.lars_lzw_decode_device
        check(decode(dev.data_ptr(), dev.data_ptr() + table_at, n, scratch.data_ptr(), counters.data_ptr(), s.cuda_stream),
              "lars_lzw_decode_device")
        check(lib.lars_tiff_post_device(scratch.data_ptr(), n, plan["slot_bytes"], plan["chunk_rows"], plan["chunk_width"],
                                        info.samples_per_pixel, sb, info.predictor, 1 if (info.big_endian and sb == 2) else 0,
                                        s.cuda_stream), "lars_tiff_post_device")
        check(lib.lars_untile_device(scratch.data_ptr(), plan["slot_bytes"], plan["slot_row_bytes"], dev.data_ptr() + moves_at,
                                     n, frames.data.data_ptr(), plan["band_row_bytes"], s.cuda_stream), "lars_untile_device")
        bad = int(counters[0].item())
    if bad:
        raise LarsError(f"{bad} compressed strip(s) / tile(s) of the band are corrupt or shorter than their rows")
    return frames


def survey_with_device_decode(sources: Sequence[Source], chunk: int = 16, engine: Optional[Engine] = None,
                              threads: Optional[int] = None, white_balance: bool = True, **fused_kw) -> dict:
    """Statistics of a survey stored as LZW TIFF frames with the decode on the GPU: chunk by chunk the compressed files
    are staged and uploaded, decoded by :func:`decode_tiff_batch_on_device`, analysed (statistics only) and folded into
    the dataset record on the device.  Same result dictionary as :meth:`SurveyPipeline.run` (single rank).  A plain
    loop -- one synchronisation per chunk for the strips' verdict, no overlap of staging and kernels yet."""
    eng = engine or get_engine()
    s = eng.stream()
    per_frame: List[np.ndarray] = []
    with torch.cuda.stream(s):
        fold = torch.zeros((2, 3, INDEX_STATS_DTYPE.itemsize), dtype=torch.uint8, device=eng.device)
    n_frames = 0
    for a in range(0, len(sources), chunk):
        part = list(sources[a:a + chunk])
        frames = decode_tiff_batch_on_device(part, eng, stream=s, threads=threads)
        res = eng.process_device(frames, outputs=("stats",), white_balance=white_balance, stream=s, **fused_kw)
        with torch.cuda.stream(s), torch.cuda.device(eng.device):
            check(eng.lib.lars_stats_merge(res.stats.data_ptr(), len(part), fold[1].data_ptr(), s.cuda_stream), "lars_stats_merge")
            check(eng.lib.lars_stats_merge(fold.data_ptr(), 2, fold[0].data_ptr(), s.cuda_stream), "lars_stats_merge")
            host = res.stats.cpu()
        s.synchronize()
        per_frame.append(host.numpy().view(INDEX_STATS_DTYPE).reshape(len(part), 3).copy())
        n_frames += len(part)
    with torch.cuda.stream(s):
        whole = fold[0].cpu()
    s.synchronize()
    records = np.concatenate(per_frame) if per_frame else np.zeros((0, 3), INDEX_STATS_DTYPE)
    bins = fused_kw.get("bins", 50)
    dataset = stats_records_to_dicts(whole.numpy().view(INDEX_STATS_DTYPE).reshape(1, 3), bins)[0]
    return {"frames": n_frames, "per_frame": records, "dataset": dataset,
            "per_frame_dicts": lambda: stats_records_to_dicts(records, bins)}


# ------------------------------------------------------------------------------------------
# streaming runner
# ------------------------------------------------------------------------------------------
class SurveyPipeline:
    """Streams equally-shaped frames through the analysis path with decode, H2D, compute and
    D2H overlapped.  One pipeline per thread / GPU; frames are independent units, so N ranks
    each run their own pipeline on their shard (``sources[rank::world]``) and the dataset-wide
    statistics meet in one all-gather (:func:`..distributed.dataset_statistics`).

    ``outputs`` selects what leaves the GPU per frame: "stats" (default, O(1) bytes per frame),
    and any of "wb", "maps", "rgb" -- those are handed, chunk by chunk, to ``on_chunk(first_frame,
    n, host)`` as views of pinned buffers that are only valid during the call.
    """

    def __init__(self, height: int, width: int, channels: int = 3, dtype=np.uint8, chunk: int = 16,
                 depth: int = 3, decode_threads: int = 8, outputs: Sequence[str] = ("stats",),
                 engine: Optional[Engine] = None, white_balance: bool = True,
                 on_chunk: Optional[Callable[[int, int, Dict[str, torch.Tensor]], None]] = None, **fused_kw):
        self.eng = engine or get_engine()
        self.h, self.w, self.c = int(height), int(width), int(channels)
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.dtype(np.uint8), np.dtype(np.uint16)):
            raise LarsError(f"dtype {self.dtype} is not supported (uint8 or uint16 frames)")
        self.sb = self.dtype.itemsize
        self.chunk, self.depth = int(chunk), max(2, int(depth))
        self.outputs = tuple(k for k in ALL_OUTPUTS if k in outputs or k == "stats")
        self.white_balance = white_balance
        self.on_chunk = on_chunk
        self.fused_kw = fused_kw
        self.decode_threads = max(1, int(decode_threads))
        # frames of a chunk decode side by side; when a chunk has fewer frames than there are decode threads
        # (large frames), the strips / tiles inside each compressed frame share the remaining threads
        self.frame_threads = max(1, self.decode_threads // max(1, min(int(chunk), self.decode_threads)))
        eng = self.eng
        self.s_in, self.s_cmp, self.s_out = (torch.cuda.Stream(device=eng.device) for _ in range(3))
        self.frame_bytes = self.h * self.w * self.c * self.sb
        self.slots = []
        for _ in range(self.depth):
            frames = eng.alloc_frames(self.chunk, self.h, self.w, self.c, self.s_cmp, sample_bytes=self.sb)
            self.slots.append({
                "host_in": torch.empty((self.chunk, self.frame_bytes), dtype=torch.uint8, pin_memory=True),
                "frames": frames,
                "res": eng.alloc_outputs(frames, self.outputs, self.s_cmp),
                "host_out": eng.alloc_host_outputs(self.chunk, self.h, self.w, self.c, self.outputs),
                "ev_in": torch.cuda.Event(), "ev_cmp": torch.cuda.Event(), "ev_out": torch.cuda.Event(),
                "pending": None,
            })
        with torch.cuda.stream(self.s_cmp):
            # [0] = running dataset record, [1] = the chunk's merged record
            self.fold = torch.zeros((2, 3, INDEX_STATS_DTYPE.itemsize), dtype=torch.uint8, device=eng.device)
        self.s_cmp.synchronize()

    # -- host side -------------------------------------------------------------------------
    def _decode_into(self, source: Source, dst_row: torch.Tensor) -> None:
        shape = (self.h, self.w, self.c)
        dst = dst_row.numpy().view(self.dtype).reshape(shape)
        arr = read_frame(source, out=None if isinstance(source, np.ndarray) else dst, threads=self.frame_threads)
        if arr.shape != shape and not (self.c == 1 and arr.shape == shape[:2]):
            raise ValueError(f"frame of shape {arr.shape} in a pipeline built for {shape}")
        if arr.dtype != self.dtype:
            raise ValueError(f"frame of dtype {arr.dtype} in a pipeline built for {self.dtype}")
        if arr.ctypes.data != dst.ctypes.data:
            np.copyto(dst, arr.reshape(shape))

    def _harvest(self, slot, per_frame: List[np.ndarray]) -> None:
        pend = slot["pending"]
        if pend is None:
            return
        slot["ev_out"].synchronize()
        first, k = pend
        rec = slot["host_out"]["stats"][:k].numpy().view(INDEX_STATS_DTYPE).reshape(k, 3)
        per_frame.append(rec.copy())
        if self.on_chunk is not None:
            self.on_chunk(first, k, {name: (t[:, :k] if name in ("maps", "rgb") else t[:k])
                                     for name, t in slot["host_out"].items()})
        slot["pending"] = None

    # -- the run ---------------------------------------------------------------------------
    def run(self, sources: Iterable[Source], group=None, distributed: Optional[bool] = None) -> dict:
        """Process every source.  Returns {"frames": n, "per_frame": [n, 3] structured records,
        "per_frame_dicts": callable -> list of dicts, "dataset": {index: {...}} over all ranks}."""
        from . import distributed as ld
        import torch.distributed as dist
        eng = self.eng
        per_frame: List[np.ndarray] = []
        with torch.cuda.stream(self.s_cmp):
            self.fold.zero_()
        ready: "queue.Queue" = queue.Queue()
        tokens: "queue.Queue" = queue.Queue()       # one token per free slot; slots are used round-robin
        for _ in range(self.depth):
            tokens.put(True)
        failure: List[BaseException] = []

        def producer():
            try:
                with ThreadPoolExecutor(self.decode_threads) as pool:
                    it = iter(sources)
                    c, first = 0, 0
                    while True:
                        batch = []
                        for src in it:
                            batch.append(src)
                            if len(batch) == self.chunk:
                                break
                        if not batch:
                            break
                        if tokens.get() is None:        # consumer gave up
                            return
                        slot = self.slots[c % self.depth]
                        futs = [pool.submit(self._decode_into, src, slot["host_in"][j]) for j, src in enumerate(batch)]
                        for f in futs:
                            f.result()
                        ready.put((c % self.depth, first, len(batch)))
                        c += 1
                        first += len(batch)
            except BaseException as exc:   # surfaced on the consumer side
                failure.append(exc)
            finally:
                ready.put(None)

        th = threading.Thread(target=producer, name="lars-ingest", daemon=True)
        th.start()
        n_frames = 0
        inflight: List[int] = []
        npx, ch = self.h * self.w, self.c
        try:
            while True:
                item = ready.get()
                if item is None:
                    break
                idx, first, k = item
                slot = self.slots[idx]
                fr: DeviceFrames = slot["frames"]
                view = DeviceFrames(fr.data[:k], fr.n_pixels, fr.channels, fr.shape, fr.sample_bytes)
                with torch.cuda.stream(self.s_in):
                    view.data[:, :self.frame_bytes].copy_(slot["host_in"][:k], non_blocking=True)
                    slot["ev_in"].record(self.s_in)
                res = slot["res"]
                sub = DeviceOutputs(frames=view,
                                    wb=None if res.wb is None else res.wb[:k],
                                    maps=None if res.maps is None else res.maps[:, :k],
                                    rgb=None if res.rgb is None else res.rgb[:, :k],
                                    stats=res.stats[:k])
                with torch.cuda.stream(self.s_cmp):
                    self.s_cmp.wait_event(slot["ev_in"])
                    eng.process_device(view, outputs=self.outputs, white_balance=self.white_balance, out=sub,
                                       stream=self.s_cmp, **self.fused_kw)
                    with torch.cuda.device(eng.device):
                        check(eng.lib.lars_stats_merge(sub.stats.data_ptr(), k, self.fold[1].data_ptr(),
                                                       self.s_cmp.cuda_stream), "lars_stats_merge")
                        check(eng.lib.lars_stats_merge(self.fold.data_ptr(), 2, self.fold[0].data_ptr(),
                                                       self.s_cmp.cuda_stream), "lars_stats_merge")
                    slot["ev_cmp"].record(self.s_cmp)
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(slot["ev_cmp"])
                    ho = slot["host_out"]
                    ho["stats"][:k].copy_(sub.stats[:k], non_blocking=True)
                    if "wb" in ho:
                        ho["wb"][:k].copy_(sub.wb[:k, :npx * ch], non_blocking=True)
                    if "maps" in ho:
                        for i in range(3):
                            ho["maps"][i, :k].copy_(sub.maps[i, :k, :npx], non_blocking=True)
                    if "rgb" in ho:
                        for i in range(3):
                            ho["rgb"][i, :k].copy_(sub.rgb[i, :k, :npx * 3], non_blocking=True)
                    slot["ev_out"].record(self.s_out)
                slot["pending"] = (first, k)
                inflight.append(idx)
                n_frames += k
                # keep depth - 1 chunks in flight on the GPU while the producer fills the remaining slot
                while len(inflight) > self.depth - 1:
                    self._harvest(self.slots[inflight.pop(0)], per_frame)
                    tokens.put(True)
            while inflight:                                  # drain in submission order
                self._harvest(self.slots[inflight.pop(0)], per_frame)
        finally:
            tokens.put(None)
            th.join()
        if failure:
            raise failure[0]
        use_dist = dist.is_initialized() and dist.get_world_size(group) > 1 if distributed is None else distributed
        with torch.cuda.stream(self.s_cmp):
            local = self.fold[0]
            whole = ld.merge_records_device(eng, ld.gather_records(local, group), self.s_cmp) if use_dist else local
            whole_h = whole.cpu()
        self.s_cmp.synchronize()
        records = np.concatenate(per_frame) if per_frame else np.zeros((0, 3), INDEX_STATS_DTYPE)
        bins = self.fused_kw.get("bins", 50)
        dataset = stats_records_to_dicts(whole_h.numpy().view(INDEX_STATS_DTYPE).reshape(1, 3), bins)[0]
        return {"frames": n_frames, "per_frame": records, "dataset": dataset,
                "per_frame_dicts": lambda: stats_records_to_dicts(records, bins)}
