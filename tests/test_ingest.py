"""Ingest: the native baseline-TIFF reader (host side of the C ABI) against Pillow and against
the arrays that were written, and the streaming SurveyPipeline against the oracle."""
import ctypes as C
import os
import warnings

import numpy as np
import pytest
from PIL import Image

from oracle import synth


def _lib():
    from lars_image_processing_b200 import _lib as L
    return L


# --------------------------------------------------------------------------------- CPU: TIFF reader
@pytest.mark.parametrize("shape", [(37, 53, 3), (64, 64, 3), (5, 7, 4), (100, 31), (300, 400, 3)])
def test_native_tiff_reader_matches_pillow_on_8bit_files(tmp_path, shape):
    from lars_image_processing_b200 import ingest
    rng = np.random.default_rng(sum(shape))
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    p = tmp_path / "a.tif"
    Image.fromarray(img).save(p)                                   # Pillow: uncompressed, multi-strip
    got = ingest.read_frame(p)
    assert got.dtype == np.uint8 and np.array_equal(got, np.array(Image.open(p))) and np.array_equal(got, img)
    assert ingest.frame_info(p) == (shape, np.dtype(np.uint8))
    # from bytes, and into a caller-supplied destination
    dst = np.empty(shape, np.uint8)
    assert ingest.read_frame(p.read_bytes(), out=dst) is not None and np.array_equal(dst, img)


@pytest.mark.parametrize("big_endian", [False, True])
@pytest.mark.parametrize("rows_per_strip", [None, 1, 7])
def test_16bit_rgb_tiff_round_trip(tmp_path, big_endian, rows_per_strip):
    """The path Pillow cannot deliver: 16-bit RGB stays 16-bit."""
    from lars_image_processing_b200 import ingest
    img = synth.vegetation_frame(5, 45, 67, np.uint16)
    p = tmp_path / "f16.tif"
    ingest.write_tiff(p, img, big_endian=big_endian, rows_per_strip=rows_per_strip)
    got = ingest.read_frame(p)
    assert got.dtype == np.uint16 and got.shape == img.shape and np.array_equal(got, img)
    # 8-bit frames written by the same writer are read identically by Pillow
    img8 = synth.vegetation_frame(6, 45, 67)
    ingest.write_tiff(p, img8, big_endian=big_endian, rows_per_strip=rows_per_strip)
    assert np.array_equal(np.array(Image.open(p)), img8) and np.array_equal(ingest.read_frame(p), img8)


def test_16bit_grayscale_tiff_agrees_with_pillow(tmp_path):
    from lars_image_processing_b200 import ingest
    img = np.random.default_rng(3).integers(0, 65536, (33, 41), dtype=np.uint16)
    p = tmp_path / "g16.tif"
    Image.fromarray(img).save(p)                                   # mode I;16
    assert np.array_equal(ingest.read_frame(p), np.array(Image.open(p)))
    ingest.write_tiff(p, img, big_endian=True)
    assert np.array_equal(np.array(Image.open(p)), img)            # Pillow reads what the writer wrote


def test_other_formats_fall_back_to_pillow(tmp_path):
    from lars_image_processing_b200 import ingest
    img = synth.vegetation_frame(8, 40, 50)
    png, jpg_tif = tmp_path / "a.png", tmp_path / "jpeg.tif"
    Image.fromarray(img).save(png)
    Image.fromarray(img).save(jpg_tif, compression="jpeg")
    assert np.array_equal(ingest.read_frame(png), img)
    # JPEG-in-TIFF: the native probe answers LARS_ERR_UNSUPPORTED and Pillow decodes
    assert ingest._tiff_probe(jpg_tif.read_bytes()) is None
    assert np.array_equal(ingest.read_frame(jpg_tif), np.array(Image.open(jpg_tif)))
    assert np.array_equal(ingest.read_frame(png.read_bytes()), img)
    assert ingest.read_frame(img) is img
    # other Pillow modes stored as TIFF: CMYK / palette indices / gray are handed over as stored by either
    # reader, bilevel goes to Pillow -- the arrays always equal np.array(Image.open(...))
    t = tmp_path / "m.tif"
    for mode, native in (("CMYK", True), ("P", True), ("L", True), ("1", False)):
        Image.fromarray(img).convert(mode).save(t)
        assert (ingest._tiff_probe(t.read_bytes()) is not None) == native, mode
        got, want = ingest.read_frame(t), np.array(Image.open(t))
        assert got.dtype == want.dtype and np.array_equal(got, want), mode


def test_corrupt_tiff_is_rejected_not_read_out_of_bounds(tmp_path):
    from lars_image_processing_b200 import ingest
    L = _lib()
    lib = L.load()
    img = synth.vegetation_frame(9, 20, 30, np.uint16)
    p = tmp_path / "x.tif"
    ingest.write_tiff(p, img)
    raw = bytearray(p.read_bytes())
    info = L.TiffInfo()
    buf = (C.c_uint8 * len(raw)).from_buffer(raw)
    assert lib.lars_tiff_probe(buf, len(raw), C.byref(info)) == 0
    assert (info.width, info.height, info.samples_per_pixel, info.bits_per_sample) == (30, 20, 3, 16)
    # truncated file: the last strip runs past the end
    assert lib.lars_tiff_probe(buf, len(raw) - 100, C.byref(info)) < 0
    assert b"strip" in lib.lars_last_error()
    # IFD offset beyond the file
    bad = bytearray(raw)
    bad[4:8] = (len(raw) + 10).to_bytes(4, "little")
    b2 = (C.c_uint8 * len(bad)).from_buffer(bad)
    assert lib.lars_tiff_probe(b2, len(bad), C.byref(info)) < 0
    # destination too small
    assert lib.lars_tiff_probe(buf, len(raw), C.byref(info)) == 0
    dst = np.empty(10, np.uint8)
    assert lib.lars_tiff_read(buf, len(raw), C.byref(info), dst.ctypes.data, dst.nbytes) < 0
    # not a TIFF at all
    junk = (C.c_uint8 * 16)(*([1] * 16))
    assert lib.lars_tiff_probe(junk, 16, C.byref(info)) < 0
    from lars_image_processing_b200._lib import LarsError
    with pytest.raises(LarsError, match="strip"):
        ingest.read_frame(bytes(raw[:len(raw) - 100]))
    p.write_bytes(bytes(raw[:len(raw) - 100]))                      # the same through a path (memory-mapped file)
    with pytest.raises(LarsError, match="strip"):
        ingest.read_frame(p)


# --------------------------------------------------------------------------------- CPU: codecs, tiles, BigTIFF
PILLOW_CODECS = {"lzw": "tiff_lzw", "deflate": "tiff_adobe_deflate", "packbits": "packbits"}


def _textured(rng, shape, dtype):
    """Half noise, half smooth ramps: both long LZW strings / PackBits runs and incompressible stretches."""
    top = np.iinfo(dtype).max
    img = rng.integers(0, top + 1, shape).astype(dtype)
    h = shape[0]
    ramp = (np.arange(shape[1]) * (top // max(shape[1], 1))).astype(dtype)
    img[h // 3: 2 * h // 3] = ramp.reshape((1, -1) + (1,) * (len(shape) - 2))
    img[2 * h // 3:] = top // 3
    return img


@pytest.mark.parametrize("codec", ["lzw", "deflate", "packbits"])
@pytest.mark.parametrize("predictor", [False, True])
def test_native_reader_matches_pillow_on_compressed_files(tmp_path, codec, predictor):
    """Files written by Pillow's libtiff: the native LZW / Deflate / PackBits decoders (+ differencing
    predictor) give what Pillow's decode of the same file gives.  300 x 400 x 3 noise overflows the LZW
    table several times per strip (Clear codes, 9 -> 12-bit widths)."""
    from lars_image_processing_b200 import ingest
    if predictor and codec == "packbits":
        pytest.skip("libtiff has no predictor in its PackBits codec")
    rng = np.random.default_rng(len(codec) + predictor)
    kw = {"tiffinfo": {317: 2}} if predictor else {}
    p = tmp_path / "c.tif"
    for shape, dtype in (((300, 400, 3), np.uint8), ((61, 47, 4), np.uint8), ((90, 130), np.uint16), ((33, 20), np.uint8)):
        img = _textured(rng, shape, dtype)
        Image.fromarray(img).save(p, compression=PILLOW_CODECS[codec], **kw)
        info = ingest._tiff_probe(p.read_bytes())
        assert info is not None and info.compression == ingest.TIFF_COMPRESSION[codec]
        assert info.predictor == (2 if predictor else 1)
        for threads in (1, 3):
            got = ingest.read_frame(p, threads=threads)
            assert got.dtype == dtype and np.array_equal(got, np.array(Image.open(p))) and np.array_equal(got, img)
        assert ingest.frame_info(p) == (shape, np.dtype(dtype))


def test_tiff_layout_sweep_codecs_tiles_bigtiff(tmp_path):
    """Writer / native reader over the whole layout space: sample width x channels x codec x predictor x
    strips / tiles x classic / BigTIFF x byte order; every file Pillow can open (8-bit or single-channel,
    not big-endian BigTIFF, which Pillow misparses) is also compared with Pillow's decode."""
    import itertools
    from lars_image_processing_b200 import ingest
    rng = np.random.default_rng(2024)
    p = tmp_path / "v.tif"
    n = n_pillow = 0
    for dtype, shape in itertools.product((np.uint8, np.uint16), ((23, 37, 3), (40, 33), (18, 50, 4))):
        img = _textured(rng, shape, dtype)
        for codec, pred, tile, big, be in itertools.product((None, "deflate", "lzw", "packbits"), (False, True),
                                                            (None, (16, 16), (32, 48)), (False, True), (False, True)):
            if pred and codec not in ("deflate", "lzw"):
                continue
            rps = None if tile else (None, 1, 7)[n % 3]
            ingest.write_tiff(p, img, big_endian=be, rows_per_strip=rps, compression=codec, predictor=pred,
                              tile=tile, bigtiff=big)
            got = ingest.read_frame(p, threads=1 + n % 3)
            assert got.dtype == dtype and got.shape == shape and np.array_equal(got, img), (dtype, shape, codec, pred, tile, big, be)
            info = ingest._tiff_probe(p.read_bytes())
            assert info.bigtiff == int(big) and info.big_endian == int(be)
            assert (info.tile_width, info.tile_length) == ((tile[1], tile[0]) if tile else (0, 0))
            n += 1
            if (dtype == np.uint8 or len(shape) == 2) and not (big and be):
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    assert np.array_equal(np.array(Image.open(p)), img), ("pillow", dtype, shape, codec, pred, tile, big, be)
                n_pillow += 1
    assert n == 432 and n_pillow == 216
    with pytest.raises(ValueError, match="predictor"):
        ingest.write_tiff(p, img, predictor=True)
    with pytest.raises(ValueError, match="compression"):
        ingest.write_tiff(p, img, compression="zstd")


def test_planar_tiff_layouts(tmp_path):
    """PlanarConfiguration 2 (band-interleaved files, GDAL's INTERLEAVE=BAND): one run of strips / tiles per sample
    of the pixel, interleaved on the way out -- every codec / predictor / strips / tiles / BigTIFF / byte order, whole
    frames and regions, compared with the arrays written and with Pillow's decode where Pillow can open the file."""
    import itertools
    from lars_image_processing_b200 import ingest
    rng = np.random.default_rng(61)
    p = tmp_path / "pl.tif"
    n = n_pillow = 0
    for dtype, shape in itertools.product((np.uint8, np.uint16), ((37, 53, 3), (20, 31, 4))):
        img = _textured(rng, shape, dtype)
        for codec, pred, tile, big, be in itertools.product((None, "deflate", "lzw", "packbits"), (False, True), (None, (16, 16)),
                                                            (False, True), (False, True)):
            if pred and codec not in ("deflate", "lzw"):
                continue
            ingest.write_tiff(p, img, big_endian=be, rows_per_strip=None if tile else 7, compression=codec, predictor=pred,
                              tile=tile, bigtiff=big, planar=True)
            info = ingest._tiff_probe(p.read_bytes())
            assert info is not None and info.planar_config == 2 and not ingest.device_decodable(p)
            got = ingest.read_frame(p, threads=1 + n % 3)
            assert got.dtype == dtype and np.array_equal(got, img), (dtype, shape, codec, pred, tile, big, be)
            assert np.array_equal(ingest.read_region(p, (5, 18), (3, 29)), img[5:18, 3:29])
            n += 1
            if dtype == np.uint8 and not (big and be):
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    assert np.array_equal(np.array(Image.open(p)), img), ("pillow", shape, codec, pred, tile, big, be)
                n_pillow += 1
    assert n == 192 and n_pillow == 72
    gray = _textured(rng, (30, 40), np.uint16)
    ingest.write_tiff(p, gray, planar=True)                        # one sample per pixel: stored chunky
    assert ingest._tiff_probe(p.read_bytes()).planar_config == 1 and np.array_equal(ingest.read_frame(p), gray)


def test_float32_tiff_maps_and_reflectance_stacks(tmp_path):
    """32-bit float TIFFs -- stored index maps (one band) and calibrated reflectance stacks (three bands, which Pillow
    cannot open at all): every codec / strips / tiles / BigTIFF / byte order / planar layout round-trips bit for bit
    (NaN and infinities included); single-band files are also compared with Pillow's decode and with Pillow-written
    files; the floating-point predictor is left to Pillow."""
    import itertools
    from lars_image_processing_b200 import ingest
    rng = np.random.default_rng(71)
    p = tmp_path / "f.tif"
    n = n_pillow = 0
    for shape in ((33, 47), (21, 30, 3), (9, 14, 4)):
        img = rng.uniform(-1, 1, shape).astype(np.float32)
        img.reshape(-1)[:6] = [np.nan, np.inf, -np.inf, 0.0, -0.0, np.float32(1e-45)]
        for codec, tile, big, be, planar in itertools.product((None, "deflate", "lzw", "packbits"), (None, (16, 16)),
                                                              (False, True), (False, True), (False, True)):
            if planar and len(shape) == 2:
                continue
            ingest.write_tiff(p, img, big_endian=be, rows_per_strip=None if tile else 5, compression=codec, tile=tile,
                              bigtiff=big, planar=planar)
            got = ingest.read_frame(p, threads=1 + n % 3)
            assert got.dtype == np.float32 and got.shape == shape
            assert np.array_equal(got.view(np.uint32), img.view(np.uint32)), (shape, codec, tile, big, be, planar)
            assert ingest.frame_info(p) == (shape, np.dtype(np.float32))
            r = ingest.read_region(p, (2, 9), (3, 13))
            assert np.array_equal(r.view(np.uint32), np.ascontiguousarray(img[2:9, 3:13]).view(np.uint32))
            n += 1
            # Pillow misparses big-endian BigTIFF and does not restore the byte order of big-endian float data that
            # went through libtiff (compressed): those layouts are pinned by the round trip only
            if len(shape) == 2 and not (big and be) and not (be and codec):
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    pil = np.array(Image.open(p))
                assert pil.dtype == np.float32 and np.array_equal(pil.view(np.uint32), img.view(np.uint32))
                n_pillow += 1
    assert n == 160 and n_pillow == 18
    ndvi = rng.uniform(-1, 1, (40, 50)).astype(np.float32)
    Image.fromarray(ndvi).save(p)                                  # Pillow's own float file (mode F)
    assert ingest._tiff_probe(p.read_bytes()) is not None and np.array_equal(ingest.read_frame(p), ndvi)
    Image.fromarray(ndvi).save(p, compression="tiff_lzw", tiffinfo={317: 3})
    assert ingest._tiff_probe(p.read_bytes()) is None              # floating-point predictor: Pillow decodes
    assert np.array_equal(ingest.read_frame(p), np.array(Image.open(p)))
    with pytest.raises(ValueError, match="predictor"):
        ingest.write_tiff(p, ndvi, compression="lzw", predictor=True)
    assert not ingest.device_decodable(p)


def test_region_reads_touch_only_their_chunks(tmp_path):
    """read_region == a NumPy crop of the whole frame for random rectangles, on strips and tiles, every
    codec, both sample widths -- and on a file whose other chunks are destroyed (only the strips / tiles
    under the rectangle are ever looked at)."""
    from lars_image_processing_b200 import ingest
    rng = np.random.default_rng(31)
    p = tmp_path / "r.tif"
    layouts = [dict(rows_per_strip=9), dict(tile=(16, 32)), dict(tile=(48, 16), compression="deflate", predictor=True),
               dict(rows_per_strip=4, compression="lzw"), dict(tile=(32, 32), compression="packbits", bigtiff=True),
               dict(rows_per_strip=11, compression="deflate", big_endian=True)]
    for k, kw in enumerate(layouts):
        dtype = np.uint16 if k % 2 else np.uint8
        shape = (70, 90, 3) if k % 3 else (70, 90)
        img = _textured(rng, shape, dtype)
        ingest.write_tiff(p, img, **kw)
        for _ in range(12):
            r0, c0 = int(rng.integers(0, 70)), int(rng.integers(0, 90))
            r1, c1 = int(rng.integers(r0 + 1, 71)), int(rng.integers(c0 + 1, 91))
            got = ingest.read_region(p, (r0, r1), (c0, c1), threads=2)
            assert got.dtype == dtype and np.array_equal(got, img[r0:r1, c0:c1]), (kw, r0, r1, c0, c1)
        assert np.array_equal(ingest.read_region(p, (5, 20)), img[5:20])                  # all columns
        dst = np.empty_like(img[10:30, 40:60])
        assert ingest.read_region(p.read_bytes(), (10, 30), (40, 60), out=dst) is not None and np.array_equal(dst, img[10:30, 40:60])
        for bad in (((0, 0), None), ((10, 5), None), ((0, 71), None), ((0, 5), (80, 91)), ((-1, 5), None)):
            with pytest.raises(ValueError):
                ingest.read_region(p, *bad)
    # rows 0..15 live in the first tile row; wipe every byte of the later tiles' data
    img = _textured(rng, (64, 64, 3), np.uint8)
    ingest.write_tiff(p, img, tile=(16, 16), compression="lzw")
    raw = bytearray(p.read_bytes())
    info = ingest._tiff_probe(bytes(raw))
    offs = np.frombuffer(bytes(raw), "<u4", info.n_strips, info.strip_offsets_pos)
    cnts = np.frombuffer(bytes(raw), "<u4", info.n_strips, info.strip_counts_pos)
    for t in range(4, info.n_strips):
        raw[offs[t]: offs[t] + cnts[t]] = b"\xff" * int(cnts[t])
    assert np.array_equal(ingest.read_region(bytes(raw), (0, 16)), img[:16])
    from lars_image_processing_b200._lib import LarsError
    with pytest.raises(LarsError, match="corrupt"):
        ingest.read_frame(bytes(raw))
    # formats Pillow decodes are cropped after the decode; arrays are cropped directly
    png = tmp_path / "a.png"
    Image.fromarray(img).save(png)
    assert np.array_equal(ingest.read_region(png, (3, 40), (7, 9)), img[3:40, 7:9])
    assert np.array_equal(ingest.read_region(img, (3, 40), (7, 9)), img[3:40, 7:9])


def test_mosaic_tiles_partition_the_image_across_ranks(tmp_path):
    """read_mosaic_tiles: the shards of all ranks are disjoint, cover the mosaic exactly once and hold the
    right pixels, whatever the file layout (file tiles need not line up with the processing tiles)."""
    from lars_image_processing_b200 import ingest
    rng = np.random.default_rng(8)
    img = _textured(rng, (96, 120, 3), np.uint16)
    p = tmp_path / "m.tif"
    for kw in (dict(tile=(32, 48), compression="deflate", predictor=True), dict(rows_per_strip=10), dict(tile=(16, 16), bigtiff=True)):
        ingest.write_tiff(p, img, **kw)
        for world in (1, 3, 8):
            seen = np.zeros(img.shape[:2], np.int32)
            for rank in range(world):
                tiles, origins = ingest.read_mosaic_tiles(p, 24, 40, rank, world, threads=2)
                assert tiles.shape == (len(origins), 24, 40, 3) and tiles.dtype == np.uint16
                for t, (r, c) in zip(tiles, origins):
                    assert np.array_equal(t, img[r:r + 24, c:c + 40])
                    seen[r:r + 24, c:c + 40] += 1
            assert (seen == 1).all()
    assert ingest.read_mosaic_tiles(img, 48, 60, 1, 2)[1] == [(48, 0), (48, 60)]
    png = tmp_path / "m.png"
    Image.fromarray((img >> 8).astype(np.uint8)).save(png)
    tiles, origins = ingest.read_mosaic_tiles(png, 48, 60, 0, 2)
    assert origins == [(0, 0), (0, 60)] and np.array_equal(tiles[1], (img >> 8).astype(np.uint8)[:48, 60:])
    with pytest.raises(ValueError, match="do not divide"):
        ingest.read_mosaic_tiles(p, 25, 40)
    # row bands need no divisor: 96 rows over 7 ranks -> 14,14,14,14,14,13,13 rows, each one frame
    rows = []
    for rank in range(7):
        band, r0 = ingest.read_mosaic_band(p, rank, 7, threads=2)
        assert band.dtype == np.uint16 and band.shape[1:] == (120, 3) and np.array_equal(band, img[r0:r0 + band.shape[0]])
        rows.append((r0, band.shape[0]))
    assert rows[0] == (0, 14) and rows[-1] == (83, 13) and sum(n for _, n in rows) == 96
    with pytest.raises(ValueError, match="ranks"):
        ingest.read_mosaic_band(p, 0, 97)
    assert ingest.largest_divisor(32768, 5000) == 4096 and ingest.largest_divisor(97, 50) == 1
    assert len(ingest.mosaic_tile_grid(32768, 32768, 4096, 4096)) == 64                  # BASELINE config 4


# --------------------------------------------------------------------------------- CPU: PNG reader
def _png_bytes(img, filters, idat=1 << 16, level=6):
    """A PNG built by hand (test infrastructure): row r is stored with filter type filters[r % len], the
    zlib stream is cut into IDAT chunks of `idat` bytes.  8- or 16-bit gray / RGB / RGBA."""
    import struct
    import zlib
    h, w = img.shape[:2]
    ch = 1 if img.ndim == 2 else img.shape[2]
    bits = img.dtype.itemsize * 8
    bpp, rb = ch * bits // 8, w * ch * bits // 8
    rows = np.frombuffer(img.astype(img.dtype.newbyteorder(">")).tobytes(), np.uint8).reshape(h, rb).astype(np.int32)
    shift = lambda v: np.concatenate([np.zeros(bpp, np.int32), v[:-bpp]]) if rb > bpp else np.zeros(rb, np.int32)
    out, prev = bytearray(), np.zeros(rb, np.int32)
    for r in range(h):
        cur, f = rows[r], filters[r % len(filters)]
        left, ul = shift(cur), shift(prev)
        if f == 0:
            x = cur
        elif f == 1:
            x = cur - left
        elif f == 2:
            x = cur - prev
        elif f == 3:
            x = cur - ((left + prev) >> 1)
        else:
            pp = left + prev - ul
            pa, pb, pc = np.abs(pp - left), np.abs(pp - prev), np.abs(pp - ul)
            x = cur - np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, prev, ul))
        out.append(f)
        out += (x & 255).astype(np.uint8).tobytes()
        prev = cur
    z = zlib.compress(bytes(out), level)
    chunk = lambda t, d: struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))
    body = b"".join(chunk(b"IDAT", z[a:a + idat]) for a in range(0, len(z), idat))
    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, bits, {1: 0, 3: 2, 4: 6}[ch], 0, 0, 0))
            + body + chunk(b"IEND", b""))


def test_native_png_reader_matches_pillow(tmp_path):
    """PNG files written by Pillow (every compression level, adaptive row filters): the native reader gives
    what np.array(Image.open(...)) gives -- 8-bit gray / RGB / RGBA and 16-bit gray."""
    from lars_image_processing_b200 import ingest
    rng = np.random.default_rng(12)
    p = tmp_path / "a.png"
    for shape, dtype in (((37, 53, 3), np.uint8), ((64, 64), np.uint8), ((50, 70, 4), np.uint8), ((33, 41), np.uint16),
                         ((1, 1, 3), np.uint8), ((5, 1), np.uint8), ((300, 400, 3), np.uint8)):
        for img in (_textured(rng, shape, dtype), rng.integers(0, np.iinfo(dtype).max + 1, shape).astype(dtype)):
            for level in (0, 1, 6, 9):
                Image.fromarray(img).save(p, compress_level=level)
                assert ingest._png_probe(p.read_bytes()) is not None        # decoded natively, not by Pillow
                got, pil = ingest.read_frame(p), np.array(Image.open(p))
                assert got.dtype == pil.dtype and got.shape == pil.shape and np.array_equal(got, pil) and np.array_equal(got, img)
                assert ingest.frame_info(p) == (shape, np.dtype(dtype))
    dst = np.empty((300, 400, 3), np.uint8)
    assert ingest.read_frame(p.read_bytes(), out=dst) is not None and np.array_equal(dst, img)
    assert np.array_equal(ingest.read_region(p, (10, 20), (5, 9)), img[10:20, 5:9])


def test_png_row_filters_idat_splits_and_16bit_rgb():
    """Hand-built PNGs: every row filter (also as the first row, where the row above is all zeros), zlib
    streams cut into many small IDAT chunks, and 16-bit RGB / RGBA, which keep their 16 bits (Pillow reduces
    16-bit RGB to 8 bits, so the written array is the reference there); 8-bit ones are also given to Pillow."""
    import io
    from lars_image_processing_b200 import ingest
    rng = np.random.default_rng(13)
    n = 0
    for shape, dtype in (((23, 31, 3), np.uint8), ((17, 9), np.uint8), ((12, 20, 4), np.uint8), ((19, 23), np.uint16),
                         ((21, 17, 3), np.uint16), ((9, 14, 4), np.uint16), ((30, 1, 3), np.uint8), ((1, 40, 3), np.uint16)):
        for img in (_textured(rng, shape, dtype), rng.integers(0, np.iinfo(dtype).max + 1, shape).astype(dtype)):
            for filters, idat in (([0], 1 << 16), ([1], 50), ([2], 1 << 16), ([3], 7), ([4], 1 << 16), ([4, 3, 2, 1, 0], 33),
                                  ([3, 4, 0, 2, 1, 4], 1 << 16)):
                blob = _png_bytes(img, filters, idat=idat)
                got = ingest.read_frame(blob)
                assert got.dtype == dtype and got.shape == shape and np.array_equal(got, img), (shape, dtype, filters, idat)
                if dtype == np.uint8 or len(shape) == 2:
                    assert np.array_equal(np.array(Image.open(io.BytesIO(blob))), img)
                n += 1
    assert n == 112


def test_png_fallbacks_and_corrupt_files(tmp_path):
    """Layouts outside the native reader go to Pillow with identical results; damaged files are rejected
    (or decode) without reading out of bounds -- 800 mutations here, 48,000 under AddressSanitizer offline."""
    import ctypes as C
    from lars_image_processing_b200 import ingest
    from lars_image_processing_b200._lib import LarsError
    L = _lib()
    lib = L.load()
    rng = np.random.default_rng(14)
    img = _textured(rng, (40, 50, 3), np.uint8)
    p = tmp_path / "f.png"
    for mode_img in (Image.fromarray(img).convert("P"), Image.fromarray(img[:, :, 0]).convert("LA"),
                     Image.fromarray(img[:, :, 0] > 100)):
        mode_img.save(p)
        assert ingest._png_probe(p.read_bytes()) is None                    # LARS_ERR_UNSUPPORTED -> Pillow
        assert np.array_equal(ingest.read_frame(p), np.array(Image.open(p)))
    good = _png_bytes(img, [4, 1, 3], idat=200)
    with pytest.raises(LarsError, match="corrupt|shorter"):
        bad = bytearray(good)
        bad[len(bad) // 2] ^= 0x55                                          # inside the zlib stream: Adler-32 / inflate fails
        ingest.read_frame(bytes(bad))
    with pytest.raises(LarsError, match="IEND|chunk"):
        ingest.read_frame(good[:len(good) - 30])
    seeds = [good, _png_bytes(_textured(rng, (19, 23), np.uint16), [0, 2, 4], idat=1 << 16),
             _png_bytes(_textured(rng, (9, 14, 4), np.uint16), [1, 3], idat=40)]
    ok = rejected = 0
    for it in range(800):
        raw = bytearray(seeds[it % 3])
        span = len(raw) if it % 3 == 0 else 64
        for _ in range(int(rng.integers(1, 4))):
            raw[int(rng.integers(0, span))] = int(rng.integers(0, 256))
        if it % 7 == 0:
            raw = raw[:int(rng.integers(1, len(raw)))]
        buf = (C.c_uint8 * len(raw)).from_buffer(raw)
        info = L.PngInfo()
        if lib.lars_png_probe(buf, len(raw), C.byref(info)) == 0:
            ok += 1
            if info.frame_bytes <= 1 << 24:
                dst = np.empty(int(info.frame_bytes), np.uint8)
                lib.lars_png_read(buf, len(raw), C.byref(info), dst.ctypes.data, dst.nbytes)
        else:
            assert lib.lars_last_error()
            rejected += 1
    assert ok + rejected == 800 and rejected > 100 and ok > 100


# --------------------------------------------------------------------------------- GPU: streaming pipeline
def _oracle(img):
    from oracle import oracle_np as o
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return o.analyze_frame(img)


@pytest.mark.gpu
def test_survey_pipeline_from_files_matches_the_oracle(engine, tmp_path):
    """PNG and TIFF files -> decode threads -> pinned ring -> GPU path -> per-frame + dataset statistics."""
    from lars_image_processing_b200 import ingest
    from oracle import oracle_np as o
    h, w, n = 96, 128, 11                                           # 11 frames, chunk 4: ragged last chunk
    frames = [synth.vegetation_frame(900 + i, h, w) for i in range(n)]
    paths = []
    for i, f in enumerate(frames):
        p = tmp_path / (f"f{i}.png" if i % 2 else f"f{i}.tif")
        Image.fromarray(f).save(p)
        paths.append(p)
    seen = {}

    def on_chunk(first, k, host):
        for j in range(k):
            seen[first + j] = (host["wb"][j].numpy().reshape(h, w, 3).copy(),
                               host["maps"][0, j].numpy().reshape(h, w).copy(),
                               host["rgb"][2, j].numpy().reshape(h, w, 3).copy())

    pipe = ingest.SurveyPipeline(h, w, chunk=4, depth=3, decode_threads=3, outputs=("stats", "wb", "maps", "rgb"),
                                 engine=engine, on_chunk=on_chunk)
    out = pipe.run(paths)
    assert out["frames"] == n and out["per_frame"].shape == (n, 3) and sorted(seen) == list(range(n))
    dicts = out["per_frame_dicts"]()
    total_hist = np.zeros(50, np.int64)
    for i, f in enumerate(frames):
        want = _oracle(f)
        assert np.array_equal(seen[i][0], want["wb"])
        assert np.array_equal(seen[i][1].view(np.uint32), want["maps"]["NDVI"].view(np.uint32))
        assert np.array_equal(seen[i][2], want["rgb"]["NDWI"])
        for t in o.INDEX_TYPES:
            assert np.array_equal(dicts[i][t]["hist"], want["stats"][t]["hist"])
            assert dicts[i][t]["count_above"] == want["stats"][t]["count_above"]
        total_hist += want["stats"]["NDVI"]["hist"]
    ds = out["dataset"]["NDVI"]
    assert ds["count"] == n * h * w and np.array_equal(ds["hist"], total_hist)
    all_ndvi = np.concatenate([_oracle(f)["maps"]["NDVI"].ravel() for f in frames]).astype(np.float64)
    assert abs(ds["mean"] - all_ndvi.mean()) <= 1e-6 * max(abs(all_ndvi.mean()), all_ndvi.std())
    assert abs(ds["std"] - all_ndvi.std()) <= 1e-6 * all_ndvi.std()
    assert ds["min"] == all_ndvi.min() and ds["max"] == all_ndvi.max()
    # a second run on the same pipeline starts from a clean dataset record
    again = pipe.run(frames[:3])
    assert again["frames"] == 3 and again["dataset"]["NDVI"]["count"] == 3 * h * w


@pytest.mark.gpu
def test_survey_pipeline_16bit_tiff_frames(engine, tmp_path):
    from lars_image_processing_b200 import ingest
    h, w, n = 64, 80, 5
    frames = [synth.vegetation_frame(950 + i, h, w, np.uint16) for i in range(n)]
    paths = []
    for i, f in enumerate(frames):
        p = tmp_path / f"s{i}.tif"
        ingest.write_tiff(p, f, big_endian=bool(i % 2), rows_per_strip=16)
        paths.append(p)
    pipe = ingest.SurveyPipeline(h, w, dtype=np.uint16, chunk=2, engine=engine)
    out = pipe.run(paths)
    dicts = out["per_frame_dicts"]()
    for i, f in enumerate(frames):
        want = _oracle(f)
        for t in ("NDVI", "GNDVI", "NDWI"):
            assert np.array_equal(dicts[i][t]["hist"], want["stats"][t]["hist"])
            assert dicts[i][t]["min"] == want["stats"][t]["min"] and dicts[i][t]["max"] == want["stats"][t]["max"]


@pytest.mark.gpu
def test_survey_pipeline_surfaces_decode_errors(engine, tmp_path):
    from lars_image_processing_b200 import ingest
    good = synth.vegetation_frame(1, 32, 48)
    pipe = ingest.SurveyPipeline(32, 48, chunk=2, engine=engine)
    with pytest.raises(ValueError):
        pipe.run([good, synth.vegetation_frame(2, 40, 48)])        # wrong shape in the stream
    assert pipe.run([good])["frames"] == 1                          # the pipeline is still usable
    assert pipe.run([])["frames"] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
def test_mosaic_file_to_tiles_to_gpu_matches_the_oracle_on_the_whole_image(engine, tmp_path, dtype):
    """BASELINE config 4 end to end in miniature: a tiled, Deflate + predictor compressed mosaic file ->
    read_mosaic_tiles (2-D processing tiles that do not line up with the file's tiles) -> one global
    white-balance histogram / stretch -> fused pass per tile -> image-wide statistics; the reassembled
    products equal the oracle's on the whole image."""
    from lars_image_processing_b200 import distributed as ld, ingest
    from lars_image_processing_b200.engine import stats_records_to_dicts
    from oracle import oracle_np as o
    img = synth.vegetation_frame(77, 256, 320, dtype)
    img[:70, :100] //= 3                                 # tiles differ: per-tile percentiles would be wrong
    p = tmp_path / "mosaic.tif"
    ingest.write_tiff(p, img, tile=(48, 112), compression="deflate", predictor=True)
    tiles, origins = ingest.read_mosaic_tiles(p, 64, 80)
    assert tiles.shape == (16, 64, 80, 3)
    dev = engine.upload(list(tiles))
    res, whole = ld.process_mosaic_tiles(engine, dev)
    out = engine.download(res)
    want = _oracle(img)
    wb = np.zeros_like(want["wb"])
    maps = {t: np.zeros((256, 320), np.float32) for t in o.INDEX_TYPES}
    for d, (r, c) in zip(out, origins):
        wb[r:r + 64, c:c + 80] = d["wb"]
        for t in o.INDEX_TYPES:
            maps[t][r:r + 64, c:c + 80] = d["maps"][t]
    assert np.array_equal(wb, want["wb"])
    st = stats_records_to_dicts(ld.records_to_numpy(whole).reshape(1, 3), 50)[0]
    for t in o.INDEX_TYPES:
        assert np.array_equal(maps[t].view(np.uint32), want["maps"][t].view(np.uint32)), t
        assert np.array_equal(st[t]["hist"], want["stats"][t]["hist"]) and st[t]["count"] == 256 * 320
        assert st[t]["count_above"] == want["stats"][t]["count_above"]
        assert st[t]["min"] == want["stats"][t]["min"] and st[t]["max"] == want["stats"][t]["max"]


@pytest.mark.gpu
def test_ragged_mosaic_bands_on_two_emulated_ranks(engine, tmp_path):
    """A mosaic whose height has no useful divisor (811 rows, prime): two 'ranks' own row bands of different
    heights (406 / 405 rows), each band is ONE frame of its own size; the white-balance counters of the two
    bands are summed (the all-reduce), both ranks build the same LUT, and the stitched products equal the
    oracle's on the whole image."""
    from lars_image_processing_b200 import ingest
    from oracle import oracle_np as o
    img = synth.vegetation_frame(91, 811, 613)
    img[:300] //= 2                                      # the bands differ: per-band percentiles would be wrong
    p = tmp_path / "ragged.tif"
    ingest.write_tiff(p, img, rows_per_strip=37, compression="lzw")
    s = engine.stream()
    bands = [ingest.read_mosaic_band(p, rank, 2) for rank in range(2)]
    assert [b.shape[0] for b, _ in bands] == [406, 405] and bands[1][1] == 406
    devs = [engine.upload([b], stream=s) for b, _ in bands]
    hists = [engine.wb_histogram(d, shared=True, stream=s) for d in devs]
    import torch
    with torch.cuda.stream(s):                           # same stream as the kernels that wrote the counters
        total = hists[0] + hists[1]                      # what dist.all_reduce(SUM) leaves on every rank
    lut, pct = engine.wb_lut(total, stream=s)
    outs = [engine.download(engine.fused(d, lut, stream=s), stream=s)[0] for d in devs]
    want = _oracle(img)
    assert np.array_equal(np.concatenate([x["wb"] for x in outs]), want["wb"])
    for t in o.INDEX_TYPES:
        got = np.concatenate([x["maps"][t] for x in outs])
        assert np.array_equal(got.view(np.uint32), want["maps"][t].view(np.uint32)), t
        assert np.array_equal(np.concatenate([x["rgb"][t] for x in outs]), want["rgb"][t]), t
        assert np.array_equal(outs[0]["stats"][t]["hist"] + outs[1]["stats"][t]["hist"], want["stats"][t]["hist"]), t


@pytest.mark.gpu
def test_lzw_tiff_frames_decoded_on_the_device(engine, tmp_path):
    """decode_tiff_batch_on_device: compressed file bytes in, frames decoded by the GPU (one warp per LZW
    strip, then predictor / byte-order kernel) -- identical to the host reader for Pillow-written 8-bit files
    (with and without predictor) and for 16-bit RGB files of either byte order; a damaged strip raises; the
    analysis of a device-decoded batch equals that of the uploaded frames."""
    import torch
    from lars_image_processing_b200 import ingest
    from lars_image_processing_b200._lib import LarsError
    rng = np.random.default_rng(41)

    def frames_to_host(dev, shape, dtype):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        torch.cuda.synchronize()
        return [dev.data[i, :n].cpu().numpy().view(dtype).reshape(shape) for i in range(dev.n_frames)]

    # Pillow / libtiff files: 8-bit RGB, several strips, table-full Clears inside the noisy strips
    imgs = [_textured(rng, (211, 333, 3), np.uint8) for _ in range(3)]
    for kw in ({}, {"tiffinfo": {317: 2}}):
        paths = []
        for i, img in enumerate(imgs):
            p = tmp_path / f"p{i}.tif"
            Image.fromarray(img).save(p, compression="tiff_lzw", **kw)
            assert ingest.device_decodable(p)
            paths.append(p)
        dev = ingest.decode_tiff_batch_on_device(paths, engine)
        for got, img, p in zip(frames_to_host(dev, img.shape, np.uint8), imgs, paths):
            assert np.array_equal(got, img) and np.array_equal(got, ingest.read_frame(p))
    # 16-bit RGB and gray, our writer: predictor on / off, both byte orders, one-row strips and large strips
    for k, (shape, kw) in enumerate((((97, 120, 3), dict(predictor=True, rows_per_strip=1)),
                                     ((97, 120, 3), dict(big_endian=True, predictor=True, rows_per_strip=16)),
                                     ((64, 50), dict(big_endian=True, rows_per_strip=7)),
                                     ((40, 33, 4), dict(rows_per_strip=40)))):
        batch = [_textured(rng, shape, np.uint16) for _ in range(2)]
        paths = []
        for i, img in enumerate(batch):
            p = tmp_path / f"w{k}_{i}.tif"
            ingest.write_tiff(p, img, compression="lzw", **kw)
            paths.append(p.read_bytes() if i else p)               # bytes and paths both work
        dev = ingest.decode_tiff_batch_on_device(paths, engine)
        assert dev.sample_bytes == 2 and dev.n_frames == 2
        for got, img in zip(frames_to_host(dev, shape, np.uint16), batch):
            assert np.array_equal(got, img), (shape, kw)
    # the analysis path runs on the decoded batch as on uploaded frames
    p = tmp_path / "a.tif"
    Image.fromarray(imgs[0]).save(p, compression="tiff_lzw")
    res_a = engine.download(engine.process_device(ingest.decode_tiff_batch_on_device([p], engine)))[0]
    res_b = engine.analyze_frame(imgs[0])
    assert np.array_equal(res_a["wb"], res_b["wb"])
    assert np.array_equal(res_a["stats"]["NDVI"]["hist"], res_b["stats"]["NDVI"]["hist"])
    # not eligible / damaged
    q = tmp_path / "z.tif"
    ingest.write_tiff(q, imgs[0], compression="packbits")
    assert not ingest.device_decodable(q) and not ingest.device_decodable(imgs[0])
    with pytest.raises(LarsError, match="LZW"):
        ingest.decode_tiff_batch_on_device([q], engine)
    raw = bytearray(p.read_bytes())
    info = ingest._tiff_probe(bytes(raw))
    off = int(np.frombuffer(bytes(raw), "<u4", info.n_strips, info.strip_offsets_pos)[3])
    raw[off: off + 4] = b"\xff" * 4                                # strip 3 now opens with code 511: no such string yet
    with pytest.raises(LarsError, match="corrupt"):
        ingest.decode_tiff_batch_on_device([bytes(raw)], engine)


@pytest.mark.gpu
def test_deflate_tiff_frames_decoded_on_the_device(engine, tmp_path):
    """Deflate TIFF strips decoded on the GPU (one warp per zlib stream) must equal the host readers' arrays
    (first hardware run: profiles/r02_ingest_gated_pytest.log)."""
    import torch
    from lars_image_processing_b200 import ingest
    rng = np.random.default_rng(43)

    def to_host(dev, shape, dtype):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        torch.cuda.synchronize()
        return [dev.data[i, :n].cpu().numpy().view(dtype).reshape(shape) for i in range(dev.n_frames)]

    imgs = [_textured(rng, (211, 333, 3), np.uint8) for _ in range(3)]
    for kw in ({}, {"tiffinfo": {317: 2}}):
        paths = []
        for i, img in enumerate(imgs):
            p = tmp_path / f"z{i}.tif"
            Image.fromarray(img).save(p, compression="tiff_adobe_deflate", **kw)
            paths.append(p)
        for got, img in zip(to_host(ingest.decode_tiff_batch_on_device(paths, engine), imgs[0].shape, np.uint8), imgs):
            assert np.array_equal(got, img)
    img16 = [_textured(rng, (97, 120, 3), np.uint16) for _ in range(2)]
    paths = []
    for i, img in enumerate(img16):
        p = tmp_path / f"z16_{i}.tif"
        ingest.write_tiff(p, img, compression="deflate", predictor=True, big_endian=bool(i), rows_per_strip=9)
        paths.append(p)
    for path, img in zip(paths, img16):                            # byte order differs: one batch each
        assert np.array_equal(to_host(ingest.decode_tiff_batch_on_device([path], engine), img.shape, np.uint16)[0], img)
        assert ingest.device_decodable(path)


@pytest.mark.gpu
def test_survey_with_device_decode_matches_the_pipeline(engine, tmp_path):
    """survey_with_device_decode (LZW files decoded on the GPU, chunk loop) gives the per-frame and dataset records of
    SurveyPipeline on the same files."""
    from lars_image_processing_b200 import ingest
    paths = []
    for i in range(7):
        p = tmp_path / f"s{i}.tif"
        Image.fromarray(synth.vegetation_frame(600 + i, 120, 160)).save(p, compression="tiff_lzw")
        paths.append(p)
    got = ingest.survey_with_device_decode(paths, chunk=3, engine=engine)
    want = ingest.SurveyPipeline(120, 160, chunk=3, engine=engine).run(paths)
    assert got["frames"] == want["frames"] == 7
    for name in got["per_frame"].dtype.names:
        assert np.array_equal(got["per_frame"][name], want["per_frame"][name]), name
    for t in ("NDVI", "GNDVI", "NDWI"):
        assert got["dataset"][t]["count"] == want["dataset"][t]["count"]
        assert np.array_equal(got["dataset"][t]["hist"], want["dataset"][t]["hist"])
        assert got["dataset"][t]["mean"] == want["dataset"][t]["mean"]


@pytest.mark.gpu
def test_mosaic_band_decoded_on_the_device(engine, tmp_path):
    """decode_tiff_region_on_device: the row band of a tiled / stripped LZW mosaic decoded on the GPU equals the host
    reader's region."""
    import torch
    from lars_image_processing_b200 import ingest
    rng = np.random.default_rng(47)
    p = tmp_path / "m.tif"
    for codec in ("lzw", "deflate"):
        for k, kw in enumerate((dict(tile=(32, 48), predictor=True), dict(rows_per_strip=9), dict(tile=(64, 64), big_endian=True))):
            dtype = np.uint16 if k % 2 else np.uint8
            img = _textured(rng, (200, 260, 3), dtype)
            ingest.write_tiff(p, img, compression=codec, **kw)
            for r0, r1 in ((0, 200), (37, 150), (190, 200)):
                dev = ingest.decode_tiff_region_on_device(p, (r0, r1), engine)
                torch.cuda.synchronize()
                n = (r1 - r0) * 260 * 3 * np.dtype(dtype).itemsize
                got = dev.data[0, :n].cpu().numpy().view(dtype).reshape(r1 - r0, 260, 3)
                assert np.array_equal(got, img[r0:r1]) and np.array_equal(got, ingest.read_region(p, (r0, r1)))


def test_tiff_round_trip_sweep(tmp_path):
    """Seeded sweep of the writer / native reader pair: shapes, sample widths, channel counts, byte
    orders and strip heights; 8-bit 1/3/4-channel and 16-bit 1-channel files are also handed to Pillow."""
    from lars_image_processing_b200 import ingest
    rng = np.random.default_rng(77)
    p = tmp_path / "s.tif"
    for case in range(40):
        h, w = int(rng.integers(1, 90)), int(rng.integers(1, 90))
        ch = (None, 3, 4)[case % 3]
        dtype = np.uint16 if case % 2 else np.uint8
        shape = (h, w) if ch is None else (h, w, ch)
        img = rng.integers(0, np.iinfo(dtype).max + 1, shape).astype(dtype)
        rps = None if case % 4 == 0 else int(rng.integers(1, h + 1))
        ingest.write_tiff(p, img, big_endian=bool(case % 5 < 2), rows_per_strip=rps)
        got = ingest.read_frame(p)
        assert got.dtype == dtype and got.shape == shape and np.array_equal(got, img), (case, shape, dtype)
        assert ingest.frame_info(p) == (shape, np.dtype(dtype))
        if dtype == np.uint8 or ch is None:
            assert np.array_equal(np.array(Image.open(p)), img), (case, "pillow")


def test_tiff_reader_survives_corrupted_files(tmp_path):
    """Robustness of the native reader: 1,200 randomly corrupted copies of valid files of every layout
    (byte flips in the header / IFD / tag values / compressed data, truncations) must each either decode
    or be rejected with an error -- never read or write out of bounds (a crash would take the test
    process down; the same driver was also run under AddressSanitizer over 768,000 mutations)."""
    import ctypes as C
    import itertools
    from lars_image_processing_b200 import ingest
    L = _lib()
    lib = L.load()
    rng = np.random.default_rng(99)
    seeds = []
    p = tmp_path / "seed.tif"
    for k, (big, dtype, codec, tile, bigtiff) in enumerate(itertools.product(
            (False, True), (np.uint8, np.uint16), (None, "lzw", "deflate", "packbits"), (None, (16, 16)), (False, True))):
        img = (rng.integers(0, 256, (19, 23, 3)) * (1 if dtype == np.uint8 else 211)).astype(dtype)
        ingest.write_tiff(p, img, big_endian=big, rows_per_strip=None if tile else 5, compression=codec,
                          predictor=bool(codec in ("lzw", "deflate") and k % 2), tile=tile, bigtiff=bigtiff,
                          planar=k % 3 == 0)
        seeds.append(p.read_bytes())
    ok = rejected = read_rejected = 0
    for it in range(1200):
        raw = bytearray(seeds[it % len(seeds)])
        span = len(raw) if it % 3 == 0 else min(len(raw), 8 + 2 + 12 * 12 + 64)
        for _ in range(int(rng.integers(1, 4))):
            raw[int(rng.integers(0, span))] = int(rng.integers(0, 256))
        if it % 5 == 0:
            raw = raw[:int(rng.integers(1, len(raw)))]
        buf = (C.c_uint8 * len(raw)).from_buffer(raw)
        info = L.TiffInfo()
        rc = lib.lars_tiff_probe(buf, len(raw), C.byref(info))
        if rc == 0:
            assert 0 < info.width and 0 < info.height
            assert info.frame_bytes == info.width * info.height * info.samples_per_pixel * (info.bits_per_sample // 8)
            if info.frame_bytes <= 1 << 24:
                dst = np.empty(int(info.frame_bytes), np.uint8)
                rc = lib.lars_tiff_read(buf, len(raw), C.byref(info), dst.ctypes.data, dst.nbytes)
                assert rc == 0 or (info.compression != 1 and lib.lars_last_error())      # only a codec may still object
                read_rejected += rc != 0
                r0, c0 = int(rng.integers(0, info.height)), int(rng.integers(0, info.width))
                lib.lars_tiff_read_region(buf, len(raw), C.byref(info), r0, info.height, c0, info.width,
                                          dst.ctypes.data, dst.nbytes, 2)
            ok += 1
        else:
            assert lib.lars_last_error()
            rejected += 1
    assert ok + rejected == 1200 and rejected > 100 and read_rejected > 10


def test_warp_lzw_decoder_equals_the_host_decoder(hostcheck):
    """lzw_warp.h (what the device-side TIFF path runs, one warp per strip) compiled for the host with its 32
    lanes run in sequence: identical output to the host decoder on valid streams of every texture (9- to
    12-bit codes, table-full Clears, KwKwK strings, chunk capacities that cut a string) and identical
    verdicts / prefixes on 3,000 corrupted or truncated streams."""
    import ctypes as C
    from lars_image_processing_b200 import ingest
    for fn in (hostcheck.hc_lzw_decode_warp, hostcheck.hc_lzw_chunk_host):
        fn.restype = C.c_uint32
        fn.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
    rng = np.random.default_rng(21)

    def both(blob, cap):
        src = np.frombuffer(blob, np.uint8)
        a, b = np.full(cap + 64, 0xAA, np.uint8), np.full(cap + 64, 0xAA, np.uint8)
        ra = hostcheck.hc_lzw_decode_warp(src.ctypes.data, src.size, a.ctypes.data, cap)
        rb = hostcheck.hc_lzw_chunk_host(src.ctypes.data, src.size, b.ctypes.data, cap)
        assert ra == rb and np.array_equal(a[:ra], b[:rb])
        assert (a[cap:] == 0xAA).all() and (b[cap:] == 0xAA).all()       # nothing written past the capacity
        if ra:
            assert (a[ra:] == 0xAA).all()                                # ... nor past what was produced
        return ra, a

    streams = []
    for n, kind in ((1, "noise"), (300, "noise"), (70000, "noise"), (70000, "ramp"), (40000, "const"), (120000, "mixed"),
                    (9000, "pairs")):
        if kind == "noise":
            d = rng.integers(0, 256, n, dtype=np.uint8)
        elif kind == "ramp":
            d = (np.arange(n) // 7 % 256).astype(np.uint8)
        elif kind == "const":
            d = np.full(n, 77, np.uint8)                                 # KwKwK on every code
        elif kind == "pairs":
            d = np.tile(np.array([5, 5, 9], np.uint8), n // 3)
        else:
            d = np.concatenate([rng.integers(0, 4, n // 2, dtype=np.uint8), (np.arange(n // 2) % 256).astype(np.uint8)])
        blob = ingest._lzw_encode(d.tobytes())
        streams.append(blob)
        for cap in (len(d), max(1, len(d) - 1), max(1, len(d) // 2), len(d) + 100):
            produced, out = both(blob, cap)
            assert produced == min(cap, len(d)) and np.array_equal(out[:produced], d[:produced]), (kind, n, cap)
    # strings defined early in an epoch and used again 20 KB later -- no Clear in between, the middle part adds few
    # table entries
    a_part = rng.integers(0, 256, 300, dtype=np.uint8)
    d = np.concatenate([a_part, ((np.arange(20000) // 50) % 2).astype(np.uint8), a_part, a_part[::-1]])
    blob = ingest._lzw_encode(d.tobytes())
    assert len(blob) < 1700                                            # the second A was coded with strings of the first (all literals: ~1,950)
    produced, out = both(blob, len(d))
    assert produced == len(d) and np.array_equal(out[:produced], d)
    rejected = 0
    for it in range(3000):
        raw = bytearray(streams[2 + it % 5])
        for _ in range(int(rng.integers(1, 4))):
            raw[int(rng.integers(0, len(raw)))] = int(rng.integers(0, 256))
        if it % 5 == 4:
            raw = raw[:int(rng.integers(1, len(raw)))]
        produced, _ = both(bytes(raw), int(rng.integers(1, 200000)))
        rejected += produced == 0
    assert rejected > 300


def test_device_decode_plan_on_the_cpu(hostcheck, tmp_path):
    """What the device-side decode is made of, without a GPU: lars_tiff_lzw_chunks (the per-strip table) plus the
    warp decoder of lzw_warp.h (run on the host through hostcheck) plus the predictor / byte-order
    step reproduce the frame for Pillow-written and self-written LZW files; files outside the device path are
    refused with LARS_ERR_UNSUPPORTED."""
    import ctypes as C
    from lars_image_processing_b200 import ingest
    L = _lib()
    lib = L.load()
    hostcheck.hc_lzw_decode_warp.restype = C.c_uint32
    hostcheck.hc_lzw_decode_warp.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
    rng = np.random.default_rng(51)
    p = tmp_path / "d.tif"

    def plan(raw):
        info = L.TiffInfo()
        assert lib.lars_tiff_probe(raw.ctypes.data, raw.size, C.byref(info)) == 0
        chunks = np.zeros(max(info.n_strips, 1), L.LZW_CHUNK_DTYPE)
        return info, chunks, lib.lars_tiff_lzw_chunks(raw.ctypes.data, raw.size, C.byref(info), chunks.ctypes.data, chunks.size)

    cases = []
    img8 = _textured(rng, (211, 333, 3), np.uint8)
    Image.fromarray(img8).save(p, compression="tiff_lzw", tiffinfo={317: 2})
    cases.append((img8, np.fromfile(p, np.uint8)))
    for shape, kw in (((97, 120, 3), dict(predictor=True, rows_per_strip=1)), ((64, 50), dict(big_endian=True, rows_per_strip=7)),
                      ((40, 33, 4), dict(big_endian=True, predictor=True, rows_per_strip=40))):
        img = _textured(rng, shape, np.uint16)
        ingest.write_tiff(p, img, compression="lzw", **kw)
        cases.append((img, np.fromfile(p, np.uint8)))
    for img, raw in cases:
        info, chunks, n = plan(raw)
        assert n == info.n_strips and ingest.device_decodable(raw.tobytes())
        sb, spp = info.bits_per_sample // 8, info.samples_per_pixel
        row_bytes = info.width * spp * sb
        assert int(chunks["dst_bytes"].sum()) == img.nbytes and chunks["dst_offset"][0] == 0
        assert np.array_equal(chunks["dst_offset"][1:], np.cumsum(chunks["dst_bytes"])[:-1])
        out = np.zeros(img.nbytes + 8, np.uint8)
        for c in chunks:
            args = (raw.ctypes.data + int(c["src_offset"]), int(c["src_bytes"]), out.ctypes.data + int(c["dst_offset"]), int(c["dst_bytes"]))
            assert hostcheck.hc_lzw_decode_warp(*args) == c["dst_bytes"]
        # what tiff_post_kernel does: byte order, then the running sum along each row per sample of the pixel
        samples = out[:img.nbytes].view(">u2" if (sb == 2 and info.big_endian) else ("<u2" if sb == 2 else np.uint8))
        rows = samples.astype(np.uint32).reshape(info.height, info.width, spp)
        if info.predictor == 2:
            rows = np.cumsum(rows, axis=1) & (0xFFFF if sb == 2 else 0xFF)
        assert np.array_equal(rows.astype(img.dtype).reshape(img.shape), img)
    # refused by the LZW chunk table: Deflate (it has its own table and IS device-decodable), tiles, uncompressed
    for kw in (dict(compression="deflate"), dict(compression="lzw", tile=(16, 16)), dict()):
        ingest.write_tiff(p, img8, **kw)
        raw = np.fromfile(p, np.uint8)
        assert plan(raw)[2] == -3 and ingest.device_decodable(p) == (kw.get("compression") == "deflate")
    big = np.zeros((600, 700, 3), np.uint8)
    ingest.write_tiff(p, big, compression="lzw")                   # one strip of 1.26 MB
    assert plan(np.fromfile(p, np.uint8))[2] == -3 and not ingest.device_decodable(p)


def test_warp_inflate_equals_zlib(hostcheck):
    """inflate_warp.h (device-side Deflate, one warp per stream) compiled for the host with its 32
    lanes run in sequence, against zlib: every compression level and strategy (stored, fixed and dynamic blocks,
    Huffman-only, RLE), textures from noise to constants (matches at every distance up to the 32 KB window, overlapping
    matches, codes longer than the 10-bit fast tables), every stream / destination alignment, capacities that cut
    the output; and 3,000 corrupted or truncated streams must be rejected or decode without touching memory past
    the capacity."""
    import ctypes as C
    import zlib
    hostcheck.hc_inflate_warp.restype = C.c_uint32
    hostcheck.hc_inflate_warp.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32]
    rng = np.random.default_rng(81)

    def run(blob, cap, skew, dskew):
        src = np.frombuffer(blob, np.uint8)
        out = np.full(cap + 72, 0xAA, np.uint8)
        n = hostcheck.hc_inflate_warp(src.ctypes.data, src.size, out.ctypes.data + dskew, cap, skew)
        assert (out[:dskew] == 0xAA).all() and (out[dskew + cap:] == 0xAA).all()
        return n, out[dskew:dskew + cap]

    datas = []
    for n, kind in ((0, "noise"), (1, "noise"), (300, "noise"), (70000, "noise"), (90000, "ramp"), (40000, "const"),
                    (150000, "mixed"), (100000, "rows"), (66000, "skewed"), (50000, "text")):
        if kind == "noise":
            d = rng.integers(0, 256, n, dtype=np.uint8)
        elif kind == "ramp":
            d = (np.arange(n) // 7 % 256).astype(np.uint8)
        elif kind == "const":
            d = np.full(n, 77, np.uint8)
        elif kind == "rows":                         # image-like: every row repeats the one 12,000 bytes above, plus noise
            row = rng.integers(0, 256, 12000, dtype=np.uint8)
            d = np.concatenate([np.where(rng.random(12000) < 0.02, rng.integers(0, 256, 12000), row).astype(np.uint8)
                                for _ in range(n // 12000 + 1)])[:n]
        elif kind == "skewed":                       # a few very frequent and many rare bytes: long Huffman codes
            d = rng.choice(256, n, p=np.r_[[0.5, 0.25, 0.125], np.full(253, 0.125 / 253)]).astype(np.uint8)
        elif kind == "text":
            d = np.frombuffer((b"white balance NDVI GNDVI NDWI " * (n // 30 + 1))[:n], np.uint8)
        else:
            d = np.concatenate([rng.integers(0, 4, n // 2, dtype=np.uint8), (np.arange(n // 2) % 256).astype(np.uint8)])
        datas.append(d)
    k = 0
    streams = []
    for d in datas:
        raw = d.tobytes()
        for level, strategy in ((0, 0), (1, 0), (6, 0), (9, 0), (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE),
                                (9, zlib.Z_FILTERED)):
            co = zlib.compressobj(level, zlib.DEFLATED, 15, 9, strategy)
            blob = co.compress(raw) + co.flush()
            streams.append(blob)
            for cap in {len(raw), max(1, len(raw) - 1), max(1, len(raw) // 2), len(raw) + 50}:
                k += 1
                n, out = run(blob, cap, k & 3, (k >> 2) & 3)
                want = min(cap, len(raw))
                if len(raw) == 0:
                    assert n == 0
                    continue
                assert n == want and np.array_equal(out[:n], d[:n]), (len(raw), level, strategy, cap)
    # a stream cut into two zlib.compress calls with a full flush in between (several blocks, byte-aligned restarts)
    co = zlib.compressobj(6)
    blob = co.compress(datas[6].tobytes()[:70000]) + co.flush(zlib.Z_FULL_FLUSH) + co.compress(datas[6].tobytes()[70000:]) + co.flush()
    n, out = run(blob, len(datas[6]), 1, 2)
    assert n == len(datas[6]) and np.array_equal(out, datas[6])
    # matches further back than the ring is trusted for (distance > 32,256: the bytes are read back from the output)
    far = np.tile(rng.integers(0, 256, 32400, dtype=np.uint8), 5)
    blob = zlib.compress(far.tobytes(), 9)
    assert len(blob) < 40000                                           # zlib did find the matches at distance 32,400
    n, out = run(blob, len(far), 3, 1)
    assert n == len(far) and np.array_equal(out, far)
    # raw garbage, wrong headers
    for bad in (b"", b"\\x78", b"\\x78\\x9c", b"\\x00\\x00\\x00\\x00", b"\\x78\\x9c\\xff\\xff\\xff\\xff", bytes(rng.integers(0, 256, 500, dtype=np.uint8))):
        n, _ = run(bad, 1000, 0, 0)
        assert n == 0 or n <= 1000
    rejected = agree = 0
    big = [s for s in streams if len(s) > 2000]
    for it in range(3000):
        raw = bytearray(big[it % len(big)])
        for _ in range(int(rng.integers(1, 4))):
            raw[int(rng.integers(0, len(raw)))] = int(rng.integers(0, 256))
        if it % 5 == 4:
            raw = raw[:int(rng.integers(1, len(raw)))]
        cap = 200000 if it % 2 else int(rng.integers(1, 200000))        # 200,000 holds every stream whole
        n, out = run(bytes(raw), cap, it & 3, (it >> 2) & 3)
        rejected += n == 0
        # whenever zlib accepts the damaged stream as a whole, the warp decoder must give the same bytes; whenever zlib
        # rejects it and the output was not cut short, so must the warp decoder (it verifies the Adler-32 trailer)
        try:
            ref = zlib.decompress(bytes(raw))
        except zlib.error:
            if cap == 200000:
                assert n == 0, it
            continue
        if n:
            agree += 1
            assert n == min(cap, len(ref)) and np.array_equal(out[:n], np.frombuffer(ref, np.uint8)[:n])
    assert rejected > 300


def test_device_region_plan_on_the_cpu(hostcheck, tmp_path):
    """Row bands of strip and tile files through the device path, emulated on the CPU: tiff_region_device_plan (chunk table
    into scratch slots, move table) + the warp decoders via hostcheck + predictor / byte order per slot + the moves
    reproduce img[r0:r1] for LZW and Deflate, strips and tiles (edge tiles narrower and shorter than a tile), 8- and
    16-bit, both byte orders."""
    import ctypes as C
    from lars_image_processing_b200 import ingest
    from lars_image_processing_b200._lib import LarsError
    hostcheck.hc_lzw_decode_warp.restype = C.c_uint32
    hostcheck.hc_lzw_decode_warp.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
    hostcheck.hc_inflate_warp.restype = C.c_uint32
    hostcheck.hc_inflate_warp.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32]
    rng = np.random.default_rng(101)
    p = tmp_path / "g.tif"
    layouts = [dict(compression="lzw", tile=(32, 48), predictor=True), dict(compression="deflate", tile=(16, 16)),
               dict(compression="lzw", rows_per_strip=9), dict(compression="deflate", rows_per_strip=13, predictor=True, big_endian=True),
               dict(compression="lzw", tile=(64, 64), big_endian=True, bigtiff=True)]
    for k, kw in enumerate(layouts):
        dtype = np.uint16 if k % 2 else np.uint8
        img = _textured(rng, (100, 130, 3), dtype)
        ingest.write_tiff(p, img, **kw)
        raw = np.fromfile(p, np.uint8)
        for r0, r1 in ((0, 100), (17, 64), (95, 100), (31, 33)):
            plan = ingest.tiff_region_device_plan(raw, (r0, r1))
            info, n = plan["info"], plan["chunks"].size
            scratch = np.zeros(n * plan["slot_bytes"] + 8, np.uint8)
            for c in plan["chunks"]:
                args = (raw.ctypes.data + int(c["src_offset"]), int(c["src_bytes"]), scratch.ctypes.data + int(c["dst_offset"]), int(c["dst_bytes"]))
                got = (hostcheck.hc_inflate_warp(*args, int(c["src_offset"]) & 3) if info.compression == 8
                       else hostcheck.hc_lzw_decode_warp(*args))
                assert got == c["dst_bytes"], (kw, r0, r1)
            sb, spp = plan["sample_bytes"], info.samples_per_pixel
            slots = scratch[:n * plan["slot_bytes"]].reshape(n, plan["chunk_rows"], plan["chunk_width"] * spp * sb)
            samples = slots.view(">u2" if (sb == 2 and info.big_endian) else ("<u2" if sb == 2 else np.uint8)).astype(np.uint32)
            samples = samples.reshape(n, plan["chunk_rows"], plan["chunk_width"], spp)
            if info.predictor == 2:                                   # what lars_tiff_post_device does to every slot
                samples = np.cumsum(samples, axis=2) & (0xFFFF if sb == 2 else 0xFF)
            slots = np.ascontiguousarray(samples.astype(img.dtype)).view(np.uint8).reshape(n, plan["chunk_rows"], -1)
            band = np.zeros((r1 - r0, plan["band_row_bytes"]), np.uint8)
            for j, (s_row, rows, d_row, d_col, nbytes) in enumerate(plan["moves"]):   # what lars_untile_device does
                band[d_row:d_row + rows, d_col:d_col + nbytes] = slots[j, s_row:s_row + rows, :nbytes]
            assert np.array_equal(band.view(img.dtype).reshape(r1 - r0, 130, 3), img[r0:r1]), (kw, r0, r1)
    for kw in (dict(), dict(compression="packbits"), dict(compression="lzw", planar=True)):
        ingest.write_tiff(p, img, **kw)
        with pytest.raises(LarsError, match="device decoders"):
            ingest.tiff_region_device_plan(np.fromfile(p, np.uint8))
    with pytest.raises(ValueError, match="rows"):
        ingest.write_tiff(p, img, compression="lzw")
        ingest.tiff_region_device_plan(np.fromfile(p, np.uint8), (5, 5))
