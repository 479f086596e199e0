"""The NumPy oracle against the golden vectors produced by the UNMODIFIED reference functions
(oracle/gen_golden.py), and against the reference itself when /root/reference is present."""
import hashlib
import warnings

import numpy as np
import pytest

from oracle import oracle_np as o
from oracle import ref_loader, synth

INDEX_TYPES = o.INDEX_TYPES


def _frame_names(golden):
    return sorted({k.split("/")[0] for k in golden["frames"].files})


def test_golden_frames_wb_maps_stats(golden):
    g = golden["frames"]
    names = _frame_names(golden)
    assert len(names) >= 15
    for name in names:
        img = g[f"{name}/input"]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            wb_lit = o.fix_white_balance_literal(img)
            wb_hist = o.fix_white_balance_from_hist(img)
        assert np.array_equal(wb_lit, g[f"{name}/wb"]), name
        assert np.array_equal(wb_hist, g[f"{name}/wb"]), name
        for t in INDEX_TYPES:
            m = o.calculate_index(wb_lit, t)
            assert np.array_equal(m.view(np.uint32), g[f"{name}/map_{t}"].view(np.uint32)), (name, t)
            st = o.analyze_index(m, t)
            assert np.array_equal(np.array(list(st.values())), g[f"{name}/stats_{t}"]), (name, t)
            assert np.array_equal(o.index_histogram(m), g[f"{name}/hist_{t}"]), (name, t)
            assert float(np.std(m)) == float(g[f"{name}/std_{t}"])


def test_golden_percentiles_from_histogram(golden):
    g = golden["frames"]
    for name in _frame_names(golden):
        img = g[f"{name}/input"]
        pcts, _ = o.wb_luts_from_hist(o.channel_histograms(img))
        assert np.array_equal(pcts, g[f"{name}/percentiles"]), name


def test_golden_pair_domain(golden):
    g = golden["pair_domain"]
    tabs = o.pair_tables()
    for t in INDEX_TYPES:
        v = tabs[t]["value"]
        assert hashlib.sha256(v.tobytes()).digest() == g[f"sha256_{t}"].tobytes()
        assert np.array_equal(np.bincount(tabs[t]["bin"].ravel(), minlength=50), g[f"hist_{t}"])
        assert np.array_equal(v[17].view(np.uint32), g[f"row17_{t}"].view(np.uint32))
        # facts the kernels rely on: no NaN, clip is a no-op, 39,641 values
        assert not np.isnan(v).any() and np.abs(v).max() <= 1.0
        assert len(np.unique(v.view(np.uint32))) == 39641
    assert tabs["NDVI"]["value"][0, 0] == 0.0 and not np.signbit(tabs["NDVI"]["value"][0, 0])


def test_golden_variants(golden):
    g = golden["variants"]
    img = g["input"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        wb = o.fix_white_balance_literal(img)
    assert np.array_equal(wb, g["backend_wb"])
    assert np.array_equal(o.fix_white_balance_rgnir_array(img), g["backend_wb"])
    f = wb.astype(np.float32)
    for t in INDEX_TYPES:
        m = o.calculate_index_planes(f[:, :, 0], f[:, :, 1], f[:, :, 2], t)
        assert np.array_equal(m.view(np.uint32), g[f"backend_{t}"].view(np.uint32))
    nd = o.calculate_ndvi_f64(img)
    assert np.array_equal(nd.view(np.uint64), g["ndvi_f64"].view(np.uint64))
    assert np.array_equal(np.array(list(o.analyze_ndvi_statistics(nd).values())), g["ndvi_f64_stats"])
    assert np.array_equal(o.index_histogram(nd), g["ndvi_f64_hist"])


def test_histogram_bin_definition_matches_numpy():
    v = o.calculate_index(o.pair_image(), "NDVI")
    for bins in (1, 2, 3, 10, 49, 50, 51, 64, 100, 128):
        b = o.histogram_bin_by_edges(v, bins)
        assert np.array_equal(np.bincount(b.ravel(), minlength=bins),
                              np.histogram(v.ravel(), bins=bins, range=(-1, 1))[0])


def test_percentile_restatement_matches_numpy():
    rng = np.random.default_rng(7)
    lerp = 0
    for it in range(1500):
        n = int(rng.integers(1, 300))
        dom = 256 if it % 2 == 0 else 65536
        x = rng.integers(0, int(rng.integers(1, dom)) + 1, size=n)
        h = np.bincount(x, minlength=dom)
        want = np.percentile(x.astype(np.float32), (2, 98))
        for i, q in enumerate((0.02, 0.98)):
            got = o.percentile_from_hist(h, q)
            assert got == want[i] and isinstance(got, np.float64)
            lerp += got != np.floor(got)
    assert lerp > 100   # the lerp branch is really exercised


def test_colormap_restatement_sanity():
    lut = o.colormap_lut_float("RdYlGn")
    assert np.allclose(lut[128], (0.9970780469050365, 0.9987697039600154, 0.7450211457131872), atol=1e-15)
    b = o.colormap_lut("RdYlGn")
    assert tuple(b[0]) == (165, 0, 38) and tuple(b[255]) == (0, 104, 55)
    assert tuple(o.colormap_lut("RdYlBu")[255]) == (49, 54, 149)
    assert tuple(o.colormap_lut("bwr")[0]) == (0, 0, 255) and tuple(o.colormap_lut("bwr")[255]) == (255, 0, 0)


def test_colormap_index_f32_and_f64_chains_agree_on_pair_domain():
    for t in INDEX_TYPES:
        v = o.calculate_index(o.pair_image(), t)
        assert np.array_equal(o.colormap_index(v), o.colormap_index(v.astype(np.float64)))


@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present on this box")
def test_oracle_against_live_reference():
    R = ref_loader.load("process-images.py")
    frames = dict(synth.adversarial_frames())
    frames["c1_like"] = synth.vegetation_frame(1, 240, 320)
    frames["u16"] = synth.vegetation_frame(3, 60, 80, np.uint16)
    for name, img in frames.items():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = R["fix_white_balance"](img)
            assert np.array_equal(o.fix_white_balance_literal(img), ref), name
            assert np.array_equal(o.fix_white_balance_from_hist(img), ref), name
        for t in INDEX_TYPES:
            mr = R["calculate_index"](ref, t)
            assert np.array_equal(mr.view(np.uint32), o.calculate_index(ref, t).view(np.uint32))
            assert R["analyze_index"](mr, t) == o.analyze_index(mr, t)
    with pytest.raises(ValueError, match="Unknown index type: EVI"):
        R["calculate_index"](frames["c1_like"], "EVI")
    with pytest.raises(ValueError, match="Unknown index type: EVI"):
        o.calculate_index(frames["c1_like"], "EVI")
    assert R["fix_white_balance"](None) is None and o.fix_white_balance_literal(None) is None
    assert R["analyze_index"](None, "NDVI") == {} and o.analyze_index(None, "NDVI") == {}


@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present on this box")
def test_oracle_against_live_reference_random_sweep():
    """240 seeded frames against the reference's own functions: ragged shapes, uint8 / uint16, RGB / RGBA,
    dense noise and few-valued content (percentiles that fall between two distinct values, p2 == p98 with its
    inf / NaN arithmetic) -- the histogram formulation the GPU uses must reproduce the reference byte for
    byte, and the index maps and statistics computed from its output bit for bit."""
    R = ref_loader.load("process-images.py")
    rng = np.random.default_rng(2025)
    degenerate = 0
    for it in range(240):
        h, w = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        ch = 4 if it % 7 == 0 else 3
        dtype = np.uint16 if it % 3 == 0 else np.uint8
        top = np.iinfo(dtype).max
        kind = it % 4
        if kind == 0:
            img = rng.integers(0, top + 1, (h, w, ch))
        elif kind == 1:                                   # a handful of levels: lerp between distinct values
            levels = rng.integers(0, top + 1, int(rng.integers(1, 5)))
            img = levels[rng.integers(0, len(levels), (h, w, ch))]
        elif kind == 2:                                   # one channel constant: p98 == p2
            img = rng.integers(0, top + 1, (h, w, ch))
            img[:, :, int(rng.integers(0, 3))] = int(rng.integers(0, top + 1))
        else:                                             # narrow band around a random level
            c = int(rng.integers(0, top + 1))
            img = np.clip(c + rng.integers(-3, 4, (h, w, ch)), 0, top)
        img = img.astype(dtype)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = R["fix_white_balance"](img)
            got = o.fix_white_balance_from_hist(img)
        assert ref.dtype == np.uint8 and np.array_equal(got, ref), (it, img.shape, dtype, kind)
        pcts, _ = o.wb_luts_from_hist(o.channel_histograms(img))
        degenerate += int((pcts[:, 0] == pcts[:, 1]).any())
        t = INDEX_TYPES[it % 3]
        mr = R["calculate_index"](ref, t)
        assert np.array_equal(mr.view(np.uint32), o.calculate_index(got, t).view(np.uint32)), (it, t)
        assert R["analyze_index"](mr, t) == o.analyze_index(mr, t), (it, t)
    assert degenerate > 40


@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present on this box")
def test_rgnir_chain_against_live_reference(tmp_path):
    """fix_white_balance_rgnir (process-rgn.py, importable: numpy + PIL only) on PNG files of small and few-valued
    frames -- where percentiles are fractional and its float64 truncation differs from fix_white_balance's float32
    store -- against both oracle formulations; the sweep must contain frames on which the two chains differ."""
    from PIL import Image
    R = ref_loader.load("process-rgn.py")
    rng = np.random.default_rng(77)
    differ = 0
    p = tmp_path / "f.png"
    for it in range(400):
        h, w = int(rng.integers(2, 24)), int(rng.integers(2, 24))
        if it % 3 == 0:
            img = rng.integers(0, 256, (h, w, 3))
        elif it % 3 == 1:
            levels = rng.integers(0, 256, int(rng.integers(1, 6)))
            img = levels[rng.integers(0, len(levels), (h, w, 3))]
        else:
            c = int(rng.integers(0, 256))
            img = np.clip(c + rng.integers(-9, 10, (h, w, 3)), 0, 255)
        img = img.astype(np.uint8)
        Image.fromarray(img).save(p)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = R["fix_white_balance_rgnir"](str(p))
            assert np.array_equal(o.fix_white_balance_rgnir_array(img), ref), it
            assert np.array_equal(o.fix_white_balance_rgnir_from_hist(img), ref), it
            differ += int(not np.array_equal(o.fix_white_balance_literal(img), ref))
    assert differ > 0
