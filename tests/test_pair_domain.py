"""The exact formulas the kernels evaluate (pixel_math.h compiled for the host) against the
oracle, exhaustively over every (a, b) uint8 pair -- the whole domain of the fused pass."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle_np as o


def _pair_tables(hc, bins, thr):
    edges = o.histogram_edges(bins)
    n = 65536
    value = np.empty(n, np.float32)
    bin_pair = np.empty(n, np.int32)
    bin_edges = np.empty(n, np.int32)
    cmap = np.empty(n, np.int32)
    above = np.empty(n, np.uint8)
    neg = np.empty(n, np.float32)
    hc.hc_pair_tables(C.c_int(bins), C.c_void_p(edges.ctypes.data), C.c_float(thr),
                      C.c_void_p(value.ctypes.data), C.c_void_p(bin_pair.ctypes.data),
                      C.c_void_p(bin_edges.ctypes.data), C.c_void_p(cmap.ctypes.data),
                      C.c_void_p(above.ctypes.data), C.c_void_p(neg.ctypes.data))
    sh = (256, 256)
    return (value.reshape(sh), bin_pair.reshape(sh), bin_edges.reshape(sh), cmap.reshape(sh),
            above.reshape(sh).astype(bool), neg.reshape(sh))


def test_pair_values_bins_cmap_coverage(hostcheck):
    tabs = o.pair_tables(50)
    # hostcheck index [hi][lo]; oracle pair image pixel (i, j) = (R=j, G=j, N=i): NDVI = (i - j)/(i + j)
    value, bin_pair, bin_edges, cmap, above, neg = _pair_tables(hostcheck, 50, 0.2)
    want = tabs["NDVI"]
    assert np.array_equal(value.view(np.uint32), want["value"].view(np.uint32))
    assert np.array_equal(bin_pair, want["bin"])
    assert np.array_equal(bin_edges, want["bin"])
    assert np.array_equal(cmap, want["cmap"])
    assert np.array_equal(above, want["above"])
    # NDWI = 0 - GNDVI bit for bit (incl. +0 where G == N), with its own bins / cmap / coverage
    assert np.array_equal(neg.view(np.uint32), tabs["NDWI"]["value"].view(np.uint32))
    _, _, _, _, above0, _ = _pair_tables(hostcheck, 50, 0.0)
    assert np.array_equal((neg > np.float32(0.0)), tabs["NDWI"]["above"])
    assert np.array_equal(above0, tabs["GNDVI"]["value"] > np.float32(0.0))


@pytest.mark.parametrize("bins", [1, 2, 3, 5, 10, 16, 25, 32, 49, 50, 51, 63, 64])
def test_pair_bin_formula_every_supported_bin_count(hostcheck, bins):
    v = o.calculate_index(o.pair_image(), "NDVI")
    want = o.histogram_bin_by_edges(v, bins)
    _, bin_pair, bin_edges, _, _, _ = _pair_tables(hostcheck, bins, 0.2)
    assert np.array_equal(bin_pair, want)
    assert np.array_equal(bin_edges, want)
    # and for the negated map (NDWI)
    vw = o.calculate_index(o.pair_image(), "NDWI")
    half = np.float32(0.5 * bins)
    bias = np.float32(half + np.float32(2.0 ** -11))
    f = (vw.astype(np.float64) * np.float64(half) + np.float64(bias)).astype(np.float32)  # one rounding == FMA
    got = np.minimum(f.astype(np.int64), bins - 1)
    assert np.array_equal(got, o.histogram_bin_by_edges(vw, bins))


def test_generic_edge_bins_on_random_floats(hostcheck):
    rng = np.random.default_rng(3)
    x = rng.uniform(-1, 1, 200000).astype(np.float32)
    for bins in (7, 50, 64):
        edges = o.histogram_edges(bins)
        x[:bins + 1] = edges            # exactly-on-edge values
        x[bins + 1:2 * bins + 2] = np.nextafter(edges, np.float32(-2))
        x[2 * bins + 2:3 * bins + 3] = np.nextafter(edges, np.float32(2))
        xc = np.clip(x, -1, 1)
        out = np.empty(xc.size, np.int32)
        hostcheck.hc_hist_bin_edges(C.c_void_p(xc.ctypes.data), C.c_int64(xc.size), C.c_int(bins),
                                    C.c_void_p(edges.ctypes.data), C.c_void_p(out.ctypes.data))
        assert np.array_equal(np.bincount(out, minlength=bins),
                              np.histogram(xc, bins=bins, range=(-1, 1))[0])


def test_subbin_table_path_matches_numpy(hostcheck):
    """K4's fast path (4097-entry sub-bin table, literal chain behind the ambiguous entries) gives np.histogram's bin
    for every float32 tried: random values, every edge and its neighbours, every sub-bin boundary and its neighbours,
    for every bin count 1..64; and the table stays nearly unambiguous (<= 2 (bins + 1) of 4097 entries)."""
    rng = np.random.default_rng(5)
    grid = (np.arange(4097, dtype=np.float64) / 2048 - 1).astype(np.float32)
    half = (np.arange(4096, dtype=np.float64) / 2048 - 1 + 1 / 4096).astype(np.float32)      # where the rounding index ties
    around = np.concatenate([grid, np.nextafter(grid, np.float32(-2)), np.nextafter(grid, np.float32(2)),
                             np.nextafter(np.nextafter(grid, np.float32(-2)), np.float32(-2)),
                             half, np.nextafter(half, np.float32(-2)), np.nextafter(half, np.float32(2))])
    for bins in list(range(1, 65)):
        edges = o.histogram_edges(bins)
        x = np.concatenate([rng.uniform(-1, 1, 20000).astype(np.float32), around, edges,
                            np.nextafter(edges, np.float32(-2)), np.nextafter(edges, np.float32(2))])
        xc = np.ascontiguousarray(np.clip(x, -1, 1))
        out = np.empty(xc.size, np.int32)
        amb = C.c_int32(0)
        for rn in (0, 1):          # truncating index, and the conversion-free round-to-nearest index the kernel uses
            hostcheck.hc_hist_bin_subbin(C.c_void_p(xc.ctypes.data), C.c_int64(xc.size), C.c_int(bins), C.c_void_p(edges.ctypes.data),
                                         C.c_void_p(out.ctypes.data), C.byref(amb), C.c_int(rn))
            assert np.array_equal(out, o.histogram_bin_by_edges(xc, bins)), (bins, rn)
            assert np.array_equal(np.bincount(out, minlength=bins), np.histogram(xc, bins=bins, range=(-1, 1))[0]), (bins, rn)
        assert amb.value <= 2 * (bins + 1), (bins, amb.value)     # an edge on a sub-bin boundary marks both neighbours


def test_float64_edge_bins_match_numpy(hostcheck):
    """K4d's bin chain: float64 linspace edges as the kernel builds them == np.linspace, and the corrected bin ==
    np.histogram on a float64 array for on-edge / next-to-edge / random values, bins 1..64."""
    rng = np.random.default_rng(4)
    for bins in (1, 2, 7, 49, 50, 51, 63, 64):
        ref_edges = np.linspace(-1, 1, bins + 1)
        x = rng.uniform(-1, 1, 100000)
        x[:bins + 1] = ref_edges
        x[bins + 1:2 * bins + 2] = np.nextafter(ref_edges, -2.0)
        x[2 * bins + 2:3 * bins + 3] = np.nextafter(ref_edges, 2.0)
        x[3 * bins + 3:3 * bins + 3 + 5000] = np.float64(rng.uniform(-1, 1, 5000).astype(np.float32))
        xc = np.ascontiguousarray(np.clip(x, -1, 1))
        out = np.empty(xc.size, np.int32)
        edges = np.empty(bins + 1, np.float64)
        hostcheck.hc_hist_bin_edges_f64(C.c_void_p(xc.ctypes.data), C.c_int64(xc.size), C.c_int(bins),
                                        C.c_void_p(out.ctypes.data), C.c_void_p(edges.ctypes.data))
        assert np.array_equal(edges, ref_edges), bins
        assert np.array_equal(np.bincount(out, minlength=bins), np.histogram(xc, bins=bins, range=(-1, 1))[0]), bins


def test_wb_lut_entry_chain(hostcheck):
    rng = np.random.default_rng(11)
    cases = [(0.0, 255.0), (10.0, 10.0), (0.0, 0.0), (255.0, 255.0), (3.5, 3.5), (12.0, 201.0),
             (4.16, 194.56), (2.8799999999999994, 155.12)]
    for _ in range(300):
        lo = float(rng.integers(0, 200)) + float(rng.choice([0.0, rng.random()]))
        hi = lo + float(rng.integers(0, 56)) + float(rng.choice([0.0, rng.random()]))
        cases.append((lo, hi))
    for lo, hi in cases:
        out = np.empty(256, np.uint8)
        hostcheck.hc_wb_lut(C.c_double(lo), C.c_double(hi), C.c_int(256), C.c_void_p(out.ctypes.data))
        assert np.array_equal(out, o.wb_lut_from_percentiles(lo, hi, 256)), (lo, hi)
    out = np.empty(65536, np.uint8)
    hostcheck.hc_wb_lut(C.c_double(5140.25), C.c_double(51400.75), C.c_int(65536), C.c_void_p(out.ctypes.data))
    assert np.array_equal(out, o.wb_lut_from_percentiles(5140.25, 51400.75, 65536))


def test_wb_lut_entry_rgn_chain(hostcheck):
    """process-rgn.py's chain (pre-clip, float64 truncated directly) has its own table: the device entry equals the
    oracle's for integral, fractional and degenerate percentile pairs, and it really differs from the
    process-images.py chain somewhere (otherwise the separate table would be untested)."""
    rng = np.random.default_rng(12)
    cases = [(0.0, 255.0), (10.0, 10.0), (0.0, 0.0), (3.5, 3.5), (12.0, 201.0), (4.16, 194.56), (2.8799999999999994, 155.12),
             (109.0, 187.20000000000002), (61.0, 244.60000000000002)]     # the last two: the chains differ (found by search)
    for _ in range(300):
        lo = float(rng.integers(0, 200)) + float(rng.choice([0.0, rng.random()]))
        hi = lo + float(rng.integers(0, 56)) + float(rng.choice([0.0, rng.random()]))
        cases.append((lo, hi))
    for _ in range(1500):                                 # percentiles of few-valued frames: lerps on the 0.02 grid
        hist = np.zeros(256, np.int64)
        hist[rng.integers(0, 256, int(rng.integers(2, 6)))] += rng.integers(1, 40, 1)
        hist[rng.integers(0, 256)] += int(rng.integers(1, 60))
        cases.append((float(o.percentile_from_hist(hist, 0.02)), float(o.percentile_from_hist(hist, 0.98))))
    differ = 0
    for lo, hi in cases:
        out = np.empty(256, np.uint8)
        hostcheck.hc_wb_lut_rgn(C.c_double(lo), C.c_double(hi), C.c_int(256), C.c_void_p(out.ctypes.data))
        assert np.array_equal(out, o.wb_lut_rgn_from_percentiles(lo, hi, 256)), (lo, hi)
        inside = slice(int(np.ceil(lo)), int(np.floor(hi)) + 1)
        differ += int(np.any(out[inside] != o.wb_lut_from_percentiles(lo, hi, 256)[inside]))
    assert differ >= 2


def test_percentile_lerp_matches_numpy(hostcheck):
    hostcheck.hc_percentile_lerp.restype = C.c_double
    hostcheck.hc_percentile_lerp.argtypes = [C.c_double] * 3
    rng = np.random.default_rng(5)
    for _ in range(2000):
        n = int(rng.integers(2, 500))
        x = np.sort(rng.integers(0, 256, n)).astype(np.float32)
        for q in (0.02, 0.98):
            vi = np.float64(n - 1) * np.float64(q)
            lo = int(np.floor(vi))
            a, b = float(x[lo]), float(x[min(lo + 1, n - 1)])
            got = hostcheck.hc_percentile_lerp(a, b, float(vi - lo))
            assert got == float(np.percentile(x, (q * 100,))[0])


def test_float64_ndvi_and_plane_ratio(hostcheck):
    rng = np.random.default_rng(9)
    hi = rng.integers(0, 256, 5000).astype(np.float64)
    lo = rng.integers(0, 256, 5000).astype(np.float64)
    hi[:3], lo[:3] = 0, 0
    out = np.empty(5000, np.float64)
    hostcheck.hc_ratio_clip_f64(C.c_void_p(hi.ctypes.data), C.c_void_p(lo.ctypes.data), C.c_int64(5000),
                                C.c_void_p(out.ctypes.data))
    want = np.clip((hi - lo) / (hi + lo + 1e-10), -1, 1)
    assert np.array_equal(out.view(np.uint64), want.view(np.uint64))
    h32, l32 = hi.astype(np.float32), lo.astype(np.float32)
    out32 = np.empty(5000, np.float32)
    hostcheck.hc_ratio_clip_f32(C.c_void_p(h32.ctypes.data), C.c_void_p(l32.ctypes.data), C.c_int64(5000),
                                C.c_void_p(out32.ctypes.data))
    want32 = np.clip((h32 - l32) / (h32 + l32 + 1e-10), -1, 1)
    assert np.array_equal(out32.view(np.uint32), want32.view(np.uint32))


def test_bwr_colormap_range_index(hostcheck):
    x = np.linspace(-0.8, 0.8, 4001).astype(np.float32)
    out = np.empty(x.size, np.int32)
    hostcheck.hc_cmap_index_range(C.c_void_p(x.ctypes.data), C.c_int64(x.size), C.c_float(-0.5), C.c_float(0.5),
                                  C.c_void_p(out.ctypes.data))
    assert np.array_equal(out, o.colormap_index(x, -0.5, 0.5))


@pytest.mark.parametrize("bins", [1, 2, 3, 5, 7, 10, 16, 25, 32, 49, 50, 51, 63, 64])
def test_conversion_free_forms_of_the_fused_kernel(hostcheck, bins):
    """Magic-number int->float, floor-by-rounding histogram row and colormap slot (the forms the
    fused kernel evaluates) against the oracle over all pairs, for x and for 0 - x (NDWI)."""
    value = np.empty(65536, np.float32)
    row = np.empty(2 * 65536, np.int32)
    slot = np.empty(2 * 65536, np.int32)
    hostcheck.hc_pair_tables_fast(C.c_int(bins), C.c_void_p(value.ctypes.data), C.c_void_p(row.ctypes.data),
                                  C.c_void_p(slot.ctypes.data))
    v = o.calculate_index(o.pair_image(), "NDVI")
    vw = o.calculate_index(o.pair_image(), "NDWI")
    assert np.array_equal(value.reshape(256, 256).view(np.uint32), v.view(np.uint32))
    fold = lambda r: np.minimum(r, bins - 1)              # the kernel folds row `bins` into the last bin
    assert np.array_equal(fold(row[:65536].reshape(256, 256)), o.histogram_bin_by_edges(v, bins))
    assert np.array_equal(fold(row[65536:].reshape(256, 256)), o.histogram_bin_by_edges(vw, bins))
    assert row.min() >= 0 and row.max() <= bins
    assert np.array_equal(np.minimum(slot[:65536], 255).reshape(256, 256), o.colormap_index(v))
    assert np.array_equal(np.minimum(slot[65536:], 255).reshape(256, 256), o.colormap_index(vw))
    assert slot.min() >= 0 and slot.max() <= 256


def test_uint16_stretch_guess_is_within_one_step_for_every_sample_value(hostcheck):
    """The uint16 fused pass maps a sample with a float guess of the stretch plus a +-1 correction against the
    bracketing thresholds (lars_fused_kernel.cuh: stretch_u16).  That is exact iff the guess is never more than
    one step away from the reference chain (fp64 -> fp32 -> uint8, pixel_math.h: lars_wb_lut_entry).  Checked
    here for ALL 65,536 sample values under 300 percentile pairs: integral and interpolated percentiles, spans
    from 0.02 to 65,535, and p2 == p98 (one step from 0 to 255 through the inf / NaN arithmetic)."""
    import ctypes as C
    rng = np.random.default_rng(16)
    hostcheck.hc_wb_lut.argtypes = [C.c_double, C.c_double, C.c_int, C.c_void_p]
    v = np.arange(65536, dtype=np.int64)
    worst = 0
    for it in range(300):
        kind = it % 6
        a = float(rng.integers(0, 65536))
        if kind == 0:                                     # integral percentiles, any span
            lo, hi = sorted((a, float(rng.integers(0, 65536))))
        elif kind == 1:                                   # interpolated (multiples of 0.02), any span
            lo = a + round(float(rng.integers(0, 50)) * 0.02, 2)
            hi = lo + float(rng.integers(0, 3000)) + round(float(rng.integers(1, 50)) * 0.02, 2)
        elif kind == 2:                                   # the narrowest non-zero spans
            lo = a
            hi = a + round(float(rng.integers(1, 6)) * 0.02, 2)
        elif kind == 3:                                   # degenerate: p2 == p98
            lo = hi = a if it % 12 == 3 else a + round(float(rng.integers(0, 50)) * 0.02, 2)
        elif kind == 4:                                   # full range and near it
            lo, hi = float(rng.integers(0, 3)), 65535.0 - float(rng.integers(0, 3))
        else:                                             # generic doubles
            lo = float(rng.uniform(0, 60000))
            hi = lo + float(rng.uniform(0.02, 65535 - lo))
        hi = min(hi, 65535.0)
        lo = min(lo, hi)
        lut = np.empty(65536, np.uint8)
        hostcheck.hc_wb_lut(lo, hi, 65536, lut.ctypes.data)
        assert (np.diff(lut.astype(np.int32)) >= 0).all()                 # monotone: thresholds are well defined
        thr = np.full(258, 65536, np.int64)
        thr[0] = 0
        thr[1:256] = np.searchsorted(lut, np.arange(1, 256), side="left")  # smallest v with LUT(v) >= k
        # the guess parameters as wb_stretch_build_u16_kernel stores them
        if hi - lo > 0.0:
            lo_int, lo_frac, scale = int(np.floor(lo)), np.float32(lo - np.floor(lo)), np.float32(255.0 / (hi - lo))
        else:
            lo_int, lo_frac, scale = int(thr[1]) - 1, np.float32(0.5), np.float32(1048576.0)
        d = (v - lo_int).astype(np.float32)                               # exact: |v - lo_int| < 2^24
        x = (d - lo_frac).astype(np.float32)
        t = (x.astype(np.float64) * np.float64(scale) - 0.5).astype(np.float32)   # fmaf: exact product, one rounding
        t = np.minimum(np.maximum(t, np.float32(-0.5)), np.float32(254.5))
        g = np.rint(t).astype(np.int64)                                   # the magic-number add rounds to nearest even
        assert g.min() >= 0 and g.max() <= 255
        worst = max(worst, int(np.abs(g - lut).max()))
        got = g + (v >= thr[g + 1]) - (v < thr[g])
        assert np.array_equal(got, lut), (it, lo, hi)
    assert worst == 1                                                     # the correction is needed, and one step suffices


def _guided_candidates(sample_hist, q=(0.02, 0.98)):
    """NumPy restatement of wb_u16_candidates_kernel (lars_u16_kernels.cuh): per percentile the bucket of rank
    floor((m - 1) q) in the SAMPLED high-byte histogram and its non-empty neighbour on the nearer side."""
    cum = np.cumsum(sample_hist)
    m = int(cum[-1])
    picked = set()
    if m == 0:
        return picked
    for qq in q:
        r = int(np.floor((m - 1) * qq))
        b = int(np.searchsorted(cum, r, side="right"))
        below = int(cum[b - 1]) if b else 0
        cnt = int(cum[b]) - below
        nz = np.nonzero(sample_hist)[0]
        if 2 * (r - below) < cnt:
            nb = nz[nz < b]
            nb = int(nb[-1]) if nb.size else -1
        else:
            nb = nz[nz > b]
            nb = int(nb[0]) if nb.size else -1
        picked.update(x for x in (b, nb) if x >= 0)
    return picked


def test_guided_uint16_guess_covers_the_true_buckets_on_ordinary_frames():
    """Design check of the uint16 guided pass without a GPU: the high-byte histogram of every k-th 4,096-pixel work unit
    (k as u16_sample_step picks it) must put the 2 % / 98 % ranks of the FULL frame inside the <= 4 candidate buckets,
    otherwise the frame pays the level-B fallback.  Vegetation-like noise and a smooth (spatially correlated) frame,
    sizes from 0.3 to 20 MP.  (Exactness never depends on this -- test_uint16_guided_single_pass_equals_two_level
    covers frames built to defeat the guess -- only the speed does.)"""
    def step_for(units):
        s = units // 64
        return 1 if s < 1 else min(16, s)
    rng = np.random.default_rng(9)
    misses = total = 0
    for h, w in ((480, 640), (960, 1280), (3000, 4000), (3648, 5472)):
        n = h * w
        units = (n * 6 + 24575) // 24576
        step = step_for(units)
        for kind in ("noise", "smooth"):
            for c, (mu, sd) in enumerate(((23130, 8995), (28270, 8995), (38550, 11565))):
                if kind == "noise":
                    plane = np.clip(rng.normal(mu, sd, n), 0, 65535).astype(np.uint16)
                else:
                    yy = np.arange(n, dtype=np.float64) / w
                    plane = np.clip(mu + sd * np.sin(yy / (37.0 + 11 * c)) + rng.normal(0, 300, n), 0, 65535).astype(np.uint16)
                hb = plane >> 8
                # sample: pixels of every step-th unit (a unit is 4,096 consecutive pixels of the interleaved frame)
                unit_of_px = np.arange(n) // 4096
                sample_hist = np.bincount(hb[unit_of_px % step == 0], minlength=256)
                cand = _guided_candidates(sample_hist)
                full = np.cumsum(np.bincount(hb, minlength=256))
                need = set()
                for qq in (0.02, 0.98):
                    vi = (n - 1) * qq
                    for r in (int(np.floor(vi)), min(int(np.floor(vi)) + 1, n - 1)):
                        need.add(int(np.searchsorted(full, r, side="right")))
                total += 1
                misses += not need <= cand
                assert len(cand) <= 4
    assert misses == 0, (misses, total)
