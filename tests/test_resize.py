"""Lanczos resize in front of the path (preprocess_large_image, process-images.py:398-422).

The arithmetic is Pillow's (src/libImaging/Resample.c); Pillow is installed here and on the GPU
box, so both the NumPy restatement (oracle/resize_np.py) and the CUDA kernels are compared with
``PIL.Image.resize(..., LANCZOS)`` itself -- bit-exact."""
import numpy as np
import pytest
from PIL import Image

from oracle import resize_np
from oracle import synth

# (in_h, in_w, out_h, out_w)
GEOMETRIES = [(37, 53, 11, 20), (300, 400, 75, 100), (123, 77, 150, 200), (64, 64, 32, 64), (64, 64, 64, 32),
              (5, 7, 2, 3), (1, 100, 1, 10), (100, 1, 10, 1), (50, 50, 50, 50), (960, 1280, 768, 1024),
              (333, 1001, 97, 255), (200, 1000, 50, 250), (90, 404, 45, 101)]


def pil_resize(img, out_h, out_w):
    return np.array(Image.fromarray(img).resize((out_w, out_h), Image.Resampling.LANCZOS))


def stripes(h, w, c):
    img = np.zeros((h, w, c), np.uint8)
    img[::2] = 255
    img[:, ::3] = 255
    return img


# --------------------------------------------------------------------------------- CPU (oracle + host tables)
@pytest.mark.parametrize("geom", GEOMETRIES)
def test_oracle_matches_pillow(geom):
    ih, iw, oh, ow = geom
    rng = np.random.default_rng(ih * 7919 + iw)
    for img in (rng.integers(0, 256, (ih, iw, 3), dtype=np.uint8), stripes(ih, iw, 3),
                rng.integers(0, 256, (ih, iw), dtype=np.uint8)):
        assert np.array_equal(resize_np.resize_lanczos(img, ow, oh), pil_resize(img, oh, ow))


def test_oracle_preprocess_large_image_matches_pillow_chain():
    img = synth.vegetation_frame(31, 1200, 1700)
    got = resize_np.preprocess_large_image(img, 1024)
    new_w, new_h = 1024, int(1200 * (1024 / 1700))                 # process-images.py:411-416
    assert got.shape == (new_h, new_w, 3)
    assert np.array_equal(got, pil_resize(img, new_h, new_w))
    small = synth.vegetation_frame(32, 100, 200)
    assert resize_np.preprocess_large_image(small, 1024) is small  # :407-408 returns the input itself
    assert resize_np.preprocess_large_image(None) is None
    tall = synth.vegetation_frame(33, 1500, 700)
    assert resize_np.preprocess_large_image(tall, 1024).shape == (1024, int(700 * (1024 / 1500)), 3)


def _unpack_planes(words, n, groups):
    """[n][groups][3] byte planes of 4 taps -> [n][4 * groups] integer coefficients
    (k = k0 + 2^8 k1 + 2^16 k2 with k2 signed)."""
    planes = words.view(np.uint32).reshape(n, groups, 3)
    shifts = 8 * np.arange(4, dtype=np.uint32)
    b = [((planes[:, :, i, None] >> shifts) & 255).astype(np.int64).reshape(n, 4 * groups) for i in range(3)]
    b[2] = np.where(b[2] > 127, b[2] - 256, b[2])
    return b[0] + (b[1] << 8) + (b[2] << 16)


@pytest.mark.parametrize("geom", GEOMETRIES + [(3000, 4000, 768, 1024), (3648, 5472, 682, 1024),
                                               (4096, 4096, 1024, 1024)])
def test_host_coefficient_tables_match_the_oracle(geom):
    """lars_resize_tables_lanczos (host code of the C ABI, no GPU) == Resample.c restatement."""
    from lars_image_processing_b200 import _lib
    ih, iw, oh, ow = geom
    plan, t = _lib.resize_plan(ih, iw, oh, ow, 3)
    assert (plan.need_h, plan.need_v) == (int(ow != iw), int(oh != ih))
    off = 0
    for need, n_in, n_out, groups, shift in ((plan.need_h, iw, ow, plan.groups_h, 0),
                                             (plan.need_v, ih, oh, plan.groups_v, plan.row_first)):
        if not need:
            continue
        ks, b, k = resize_np.precompute_coeffs(n_in, n_out)
        assert groups == (ks + 3) // 4
        b = b.copy()
        b[:, 0] -= shift                       # vertical windows address rows of the intermediate image
        assert np.array_equal(t[off:off + 2 * n_out].reshape(n_out, 2), b)
        off += 2 * n_out
        unpacked = _unpack_planes(t[off:off + n_out * groups * 3], n_out, groups)
        off += n_out * groups * 3
        assert np.array_equal(unpacked[:, :ks], k) and not unpacked[:, ks:].any()
    if plan.mma_ksteps == 0:
        assert off * 4 == max(plan.table_bytes, 4) or (off == 0 and plan.table_bytes == 4)
    else:
        # tensor-core horizontal pass: block starts + B fragments of mma.m16n8k32 must encode the same
        # coefficients (byte planes) at the right (input pixel, output column) positions
        assert plan.mma_table_offset == (off * 4 + 7) // 8 * 8
        ks_h, b_h, k_h = resize_np.precompute_coeffs(iw, ow)
        nblocks = (ow + 7) // 8
        m0 = plan.mma_table_offset // 4
        kstart = t[m0:m0 + nblocks]
        f0 = m0 + (nblocks + 1) // 2 * 2
        frag = t[f0:f0 + nblocks * plan.mma_ksteps * 3 * 32 * 2].view(np.uint32).reshape(nblocks, plan.mma_ksteps, 3, 32, 2)
        assert (f0 + frag.size) * 4 == plan.table_bytes
        for nb in sorted({0, min(1, nblocks - 1), nblocks // 2, nblocks - 1}):
            assert kstart[nb] == (b_h[nb * 8, 0] & ~3)
            dense = np.zeros((8, plan.mma_ksteps * 32), np.int64)          # [column][k]
            for ks in range(plan.mma_ksteps):
                for lane in range(32):
                    g, t4 = lane >> 2, (lane & 3) * 4
                    for half in range(2):
                        for j in range(4):
                            b = [(int(frag[nb, ks, pl, lane, half]) >> (8 * j)) & 255 for pl in range(3)]
                            hi = b[2] - 256 if b[2] > 127 else b[2]
                            dense[g, ks * 32 + half * 16 + t4 + j] = b[0] + (b[1] << 8) + (hi << 16)
            for g in range(8):
                xo = nb * 8 + g
                want = np.zeros(plan.mma_ksteps * 32, np.int64)
                if xo < ow:
                    first, cnt = int(b_h[xo, 0]), int(b_h[xo, 1])
                    want[first - kstart[nb]:first - kstart[nb] + cnt] = k_h[xo, :cnt]
                assert np.array_equal(dense[g], want), (nb, g)
    if plan.need_h and plan.need_v:
        assert plan.row_first + plan.row_count <= ih and plan.temp_frame_bytes >= plan.row_count * ow * 3


def test_resize_plan_rejects_bad_arguments():
    import ctypes as C
    from lars_image_processing_b200 import _lib
    lib = _lib.load()
    plan = _lib.ResizePlan()
    assert lib.lars_resize_plan_lanczos(0, 10, 5, 5, 3, C.byref(plan)) < 0
    assert lib.lars_resize_plan_lanczos(10, 10, 5, 5, 2, C.byref(plan)) < 0
    assert lib.lars_resize_plan_lanczos(10, 10, 5, 5, 3, None) < 0
    assert b"channels" in lib.lars_last_error() or b"NULL" in lib.lars_last_error()
    # a window that cannot fit in shared memory is refused, not silently mis-computed
    assert lib.lars_resize_plan_lanczos(8, 2_000_000, 8, 64, 3, C.byref(plan)) < 0


# --------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("geom", GEOMETRIES + [(1500, 2000, 384, 512)])
def test_gpu_resize_is_bit_identical_to_pillow(engine, geom):
    ih, iw, oh, ow = geom
    rng = np.random.default_rng(ih * 31 + iw)
    frames = [rng.integers(0, 256, (ih, iw, 3), dtype=np.uint8), stripes(ih, iw, 3),
              synth.vegetation_frame(ih + iw, ih, iw)]
    got = engine.resize_batch(frames, oh, ow)
    for g, f in zip(got, frames):
        assert g.shape == (oh, ow, 3) and g.dtype == np.uint8
        assert np.array_equal(g, pil_resize(f, oh, ow))
    gray = rng.integers(0, 256, (ih, iw), dtype=np.uint8)
    assert np.array_equal(engine.resize_batch([gray], oh, ow)[0], pil_resize(gray, oh, ow))


@pytest.mark.gpu
def test_gpu_preprocess_large_image_drop_in(engine):
    from lars_image_processing_b200 import process_images as pi
    assert pi.preprocess_large_image(None) is None
    assert pi.preprocess_large_image(np.zeros((0, 0, 3), np.uint8)) is None
    small = synth.vegetation_frame(41, 600, 800)
    assert pi.preprocess_large_image(small) is small
    for h, w in ((3000, 4000), (2000, 1100), (1025, 1025)):
        img = synth.vegetation_frame(h + w, h, w)
        got = pi.preprocess_large_image(img, 1024)
        want = resize_np.preprocess_large_image(img, 1024)
        assert got.shape == want.shape and np.array_equal(got, want)
        t = resize_np.target_size(h, w, 1024)
        assert np.array_equal(got, pil_resize(img, t[0], t[1]))
    with pytest.raises(TypeError):
        pi.preprocess_large_image(np.zeros((2000, 2000, 3), np.uint16))


@pytest.mark.gpu
def test_gpu_analysis_after_device_side_resize(engine):
    """The app's order -- preprocess_large_image, then white balance and indices (process-images.py:1130,
    :1444-1457, :1522) -- in one trip: resize on the device, analysis on the resized frame."""
    import warnings
    from oracle import oracle_np as o
    img = synth.vegetation_frame(77, 1536, 2048)
    res = engine.analyze_batch([img], max_dimension=1024)[0]
    small = resize_np.preprocess_large_image(img, 1024)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = o.analyze_frame(small)
    assert np.array_equal(res["wb"], want["wb"])
    for t in o.INDEX_TYPES:
        assert np.array_equal(res["maps"][t].view(np.uint32), want["maps"][t].view(np.uint32))
        assert np.array_equal(res["stats"][t]["hist"], want["stats"][t]["hist"])


@pytest.mark.gpu
def test_gpu_resize_random_geometries(engine):
    """Seeded sweep over 60 random geometries (down- and up-scaling, partial 8-column blocks, widths
    that take the tensor-core path (row bytes divisible by 4) and widths that take the DP4A fallback,
    1 to 5 k-steps, single rows / columns): always bit-identical to Pillow."""
    rng = np.random.default_rng(20261018)
    for case in range(60):
        ih, iw = int(rng.integers(1, 420)), int(rng.integers(1, 420))
        if case % 3 == 0:
            iw = max(4, iw // 4 * 4)                        # aligned rows -> tensor-core horizontal pass
        if case % 7 == 0:
            oh, ow = int(rng.integers(1, 40)), int(rng.integers(1, 40))      # strong down-scaling: many k-steps
        else:
            oh, ow = int(rng.integers(1, 500)), int(rng.integers(1, 500))
        img = rng.integers(0, 256, (ih, iw, 3), dtype=np.uint8)
        got = engine.resize_batch([img, stripes(ih, iw, 3)], oh, ow)
        assert np.array_equal(got[0], pil_resize(img, oh, ow)), (case, ih, iw, oh, ow)
        assert np.array_equal(got[1], pil_resize(stripes(ih, iw, 3), oh, ow)), (case, ih, iw, oh, ow)


# --------------------------------------------------------------------------------- committed golden vectors
def _golden_cases(golden):
    g = golden["resize"]
    i = 0
    while f"in_{i}" in g:
        yield g[f"in_{i}"], g[f"out_{i}"]
        i += 1


def test_oracle_against_committed_pillow_vectors(golden):
    """tests/golden/resize.npz holds Pillow's own outputs (oracle/gen_golden_resize.py)."""
    n = 0
    for img, want in _golden_cases(golden):
        assert np.array_equal(resize_np.resize_lanczos(img, want.shape[1], want.shape[0]), want)
        n += 1
    assert n >= 5
    g = golden["resize"]
    assert np.array_equal(resize_np.preprocess_large_image(g["pre_in"], 1024), g["pre_out"])


@pytest.mark.gpu
def test_gpu_resize_against_committed_pillow_vectors(engine, golden):
    from lars_image_processing_b200 import process_images as pi
    for img, want in _golden_cases(golden):
        assert np.array_equal(engine.resize_batch([img], want.shape[0], want.shape[1])[0], want)
    g = golden["resize"]
    assert np.array_equal(pi.preprocess_large_image(g["pre_in"], 1024), g["pre_out"])


@pytest.mark.gpu
def test_gpu_rgba_frames_are_resized_with_premultiplied_alpha(engine):
    """np.array(Image.open(upload)) of an RGBA PNG reaches preprocess_large_image as HxWx4; Pillow resizes RGBA with
    premultiplied alpha (convert("RGBa") -> resize -> convert("RGBA")), and so must the drop-in -- bit for bit, for
    opaque, transparent, partly transparent and random alpha."""
    from lars_image_processing_b200 import process_images as pi
    rng = np.random.default_rng(61)
    for h, w in ((1500, 2100), (1200, 1030)):
        rgb = synth.vegetation_frame(h, h, w)
        alphas = [np.full((h, w), 255, np.uint8), np.zeros((h, w), np.uint8), rng.integers(0, 256, (h, w), dtype=np.uint8),
                  np.where(rng.random((h, w)) < 0.5, 255, rng.integers(0, 256, (h, w))).astype(np.uint8),
                  np.tile(np.arange(w, dtype=np.int64) * 255 // (w - 1), (h, 1)).astype(np.uint8)]
        for a in alphas:
            img = np.dstack([rgb, a])
            got = pi.preprocess_large_image(img, 1024)
            t = resize_np.target_size(h, w, 1024)
            want = pil_resize(img, t[0], t[1])
            assert got.shape == want.shape == (t[0], t[1], 4) and np.array_equal(got, want)
    # the device-side form used by analyze_batch(max_dimension=...) follows the same policy
    img = np.dstack([synth.vegetation_frame(7, 1300, 1100), rng.integers(0, 256, (1300, 1100), dtype=np.uint8)])
    dev = engine.upload([img])
    out = engine.resize_device(dev, 1024, 866)
    import torch
    torch.cuda.synchronize()
    got = out.data[0, :1024 * 866 * 4].cpu().numpy().reshape(1024, 866, 4)
    assert np.array_equal(got, pil_resize(img, 1024, 866))
