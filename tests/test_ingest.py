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
    png, lzw = tmp_path / "a.png", tmp_path / "lzw.tif"
    Image.fromarray(img).save(png)
    Image.fromarray(img).save(lzw, compression="tiff_lzw")
    assert np.array_equal(ingest.read_frame(png), img)
    assert np.array_equal(ingest.read_frame(lzw), img)             # compressed TIFF: LARS_ERR_UNSUPPORTED -> Pillow
    assert np.array_equal(ingest.read_frame(png.read_bytes()), img)
    assert ingest.read_frame(img) is img


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
    """Robustness of the native reader: 600 randomly corrupted copies of valid files (byte flips in the
    header / IFD / tag values, truncations) must each either decode to the right shape or be rejected with
    an error -- never read out of bounds (a crash would take the test process down)."""
    import ctypes as C
    from lars_image_processing_b200 import ingest
    L = _lib()
    lib = L.load()
    rng = np.random.default_rng(99)
    seeds = []
    for big in (False, True):
        for dtype in (np.uint8, np.uint16):
            p = tmp_path / "seed.tif"
            ingest.write_tiff(p, rng.integers(0, 256, (19, 23, 3)).astype(dtype), big_endian=big, rows_per_strip=5)
            seeds.append(p.read_bytes())
    ok = rejected = 0
    for it in range(600):
        raw = bytearray(seeds[it % len(seeds)])
        header_len = min(len(raw), 8 + 2 + 12 * 12 + 64)
        for _ in range(int(rng.integers(1, 4))):
            pos = int(rng.integers(0, header_len))
            raw[pos] = int(rng.integers(0, 256))
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
                assert lib.lars_tiff_read(buf, len(raw), C.byref(info), dst.ctypes.data, dst.nbytes) == 0
            ok += 1
        else:
            assert lib.lars_last_error()
            rejected += 1
    assert ok + rejected == 600 and rejected > 50
