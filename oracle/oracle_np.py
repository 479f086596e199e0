"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the RGNir per-pixel analysis path.

A NumPy restatement of the reference's hot path (white balance -> NDVI/GNDVI/NDWI ->
statistics -> histogram -> colormap).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module; the
product package ``lars_image_processing_b200`` never does (it fails loudly without its
CUDA library instead of falling back to anything in here).

Parity status
-------------
* The reference ships NO tests, fixtures or golden vectors, so nothing of
  the reference's own pins this oracle.  It is pinned instead against the reference's
  *own functions executed unmodified* in the authoring container (``oracle/ref_loader.py``
  + ``oracle/gen_golden.py`` -> ``tests/golden/*.npz``; ``tests/test_oracle_vs_reference.py``
  re-runs the comparison whenever ``/root/reference`` is present).
* The arithmetic of the path lives in NumPy (``requirements.txt:2``, unpinned upstream);
  this oracle is pinned to **NumPy 2.3.x semantics** (NEP 50 promotion: the float64
  percentile scalars promote the white-balance expression to float64).
* Colormap RGB *bytes*: **parity unpinned** -- matplotlib (``requirements.txt:4``) is not
  installable here.  ``colormap_lut`` restates matplotlib's published
  ``LinearSegmentedColormap`` construction from the ColorBrewer anchors; the colormap
  *index* (what north_star grades bit-exact) is pinned by formula.

Two flavours of every white-balance step are kept on purpose:
``*_literal``  : calls the same NumPy routines as the reference (np.percentile, ...)
                 -- this is what the CPU baseline times;
``*_from_hist``: the histogram -> order statistic -> LUT formulation the GPU uses,
                 proved equal to the literal one in the CPU test-suite.
"""
from __future__ import annotations

import numpy as np

INDEX_TYPES = ("NDVI", "GNDVI", "NDWI")
EPSILON = 1e-10                      # process-images.py:464
HIST_BINS = 50                       # process-ndvi.py:97
HIST_RANGE = (-1, 1)                 # process-ndvi.py:97
PERCENTILES = (2, 98)                # process-images.py:437


# --------------------------------------------------------------------------------------
# white balance  (process-images.py:424-447, backend-process.py:17-26, process-rgn.py:4-49)
# --------------------------------------------------------------------------------------
def fix_white_balance_literal(img_array):
    """process-images.py:424-447 restated with the same NumPy calls and dtype flow."""
    if img_array is None or img_array.size == 0:          # :427-428
        return None
    as_f32 = img_array.astype(np.float32)                 # :431
    stretched = np.zeros_like(as_f32)                     # :432 (extra channels stay 0)
    for c in (0, 1, 2):                                   # :435
        plane = as_f32[:, :, c]
        lo, hi = np.percentile(plane, PERCENTILES)        # :437 -> float64 scalars
        # :438 -- float64 expression (NEP 50), rounded to float32 on assignment
        stretched[:, :, c] = np.clip((plane - lo) / (hi - lo) * 255, 0, 255)
    return stretched.astype(np.uint8)                     # :441 truncation


def stretch_channel_rgn(channel_f64):
    """process-rgn.py:25-33 (float64, explicit pre-clip to [p2, p98])."""
    lo, hi = np.percentile(channel_f64, PERCENTILES)
    inner = np.clip(channel_f64, lo, hi)
    return np.clip((inner - lo) / (hi - lo) * 255, 0, 255)


def fix_white_balance_rgnir_array(img_array):
    """process-rgn.py:18-44 starting from the decoded array (file decode is not on the path)."""
    as_f64 = img_array.astype(float)
    planes = [stretch_channel_rgn(as_f64[:, :, c]) for c in (0, 1, 2)]
    return np.dstack(planes).astype(np.uint8)


def channel_histograms(img_array, domain=None):
    """Per-channel value histogram of an integer HWC image -> (3, domain) int64."""
    if domain is None:
        domain = 256 if img_array.dtype == np.uint8 else 65536
    flat = img_array.reshape(-1, img_array.shape[-1])
    return np.stack([np.bincount(flat[:, c], minlength=domain) for c in (0, 1, 2)]).astype(np.int64)


def _value_at_rank(cum, rank):
    """Smallest value v with cum[v] > rank (cum = inclusive cumulative histogram)."""
    return int(np.searchsorted(cum, rank, side="right"))


def percentile_from_hist(hist, q):
    """np.percentile(x, q*100) ("linear" method) from the value histogram of x.

    Follows numpy/lib/_function_base_impl.py: virtual index ``(n-1)*q`` (:126-129),
    neighbours floor / floor+1 (:4753-4786), ``_lerp`` with the ``t >= 0.5`` rewrite
    (:4657-4678).  ``a`` and ``b`` are exact small integers so every step below is the
    same IEEE double operation NumPy performs.
    """
    hist = np.asarray(hist, dtype=np.int64)
    n = int(hist.sum())
    cum = np.cumsum(hist)
    vi = np.float64(n - 1) * np.float64(q)
    lo = int(np.floor(vi))
    if vi >= n - 1:                         # _get_indexes: above bounds -> last element
        a = b = _value_at_rank(cum, n - 1)
    else:
        a = _value_at_rank(cum, lo)
        b = _value_at_rank(cum, lo + 1)
    gamma = vi - np.float64(lo)
    a64, b64 = np.float64(a), np.float64(b)
    diff = b64 - a64
    if gamma >= 0.5:
        return b64 - diff * (np.float64(1) - gamma)
    return a64 + diff * gamma


def wb_lut_from_percentiles(lo, hi, domain=256):
    """uint8 LUT equal to process-images.py:438,441 evaluated on every possible input value.

    float64 expression -> clip -> float32 store -> uint8 truncation.  ``hi == lo`` gives
    +-inf / NaN exactly as in the reference (NaN -> 0 on x86, SURVEY.md section 7 hard part 4).
    """
    v = np.arange(domain, dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.clip((v - np.float64(lo)) / (np.float64(hi) - np.float64(lo)) * 255, 0, 255)
        t32 = t.astype(np.float32)
        out = np.where(np.isnan(t32), np.float32(0), t32).astype(np.uint8)
    return out


def wb_lut_rgn_from_percentiles(lo, hi, domain=256):
    """uint8 LUT equal to process-rgn.py:25-33, :44 evaluated on every possible sample value: pre-clip to
    [lo, hi], float64 stretch, clip, float64 -> uint8 truncation (no float32 store in between, unlike
    process-images.py:438).  ``hi == lo``: every sample clips to lo, 0/0 = NaN -> 0."""
    v = np.arange(domain, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        inner = np.clip(v, np.float64(lo), np.float64(hi))
        t = np.clip((inner - np.float64(lo)) / (np.float64(hi) - np.float64(lo)) * 255, 0, 255)
        out = np.where(np.isnan(t), 0.0, t).astype(np.uint8)
    return out


def fix_white_balance_rgnir_from_hist(img_array):
    """fix_white_balance_rgnir_array as histogram -> percentiles -> LUT -> gather (the GPU formulation)."""
    hist3 = channel_histograms(img_array)
    out = np.zeros(img_array.shape[:2] + (3,), dtype=np.uint8)
    for c in (0, 1, 2):
        lo, hi = percentile_from_hist(hist3[c], 0.02), percentile_from_hist(hist3[c], 0.98)
        out[:, :, c] = wb_lut_rgn_from_percentiles(lo, hi, hist3.shape[1])[img_array[:, :, c]]
    return out


def wb_luts_from_hist(hist3, q=(0.02, 0.98)):
    """(3, domain) histograms -> ((3,2) float64 percentiles, (3, domain) uint8 LUTs)."""
    hist3 = np.asarray(hist3)
    pcts = np.array([[percentile_from_hist(hist3[c], q[0]), percentile_from_hist(hist3[c], q[1])]
                     for c in range(3)], dtype=np.float64)
    luts = np.stack([wb_lut_from_percentiles(pcts[c, 0], pcts[c, 1], hist3.shape[1])
                     for c in range(3)])
    return pcts, luts


def fix_white_balance_from_hist(img_array):
    """Histogram -> percentiles -> LUT -> gather; the formulation the GPU path uses."""
    if img_array is None or img_array.size == 0:
        return None
    _, luts = wb_luts_from_hist(channel_histograms(img_array))
    out = np.zeros(img_array.shape, dtype=np.uint8)
    for c in (0, 1, 2):
        out[:, :, c] = luts[c][img_array[:, :, c]]
    return out


# --------------------------------------------------------------------------------------
# index maps  (process-images.py:449-490, backend-process.py:28-38, process-ndvi.py:18-31)
# --------------------------------------------------------------------------------------
def _band_pair(index_type):
    # (minuend channel, subtrahend channel): NDVI (N-R)/(N+R), GNDVI (N-G)/(N+G), NDWI (G-N)/(G+N)
    try:
        return {"NDVI": (2, 0), "GNDVI": (2, 1), "NDWI": (1, 2)}[index_type]
    except KeyError:
        raise ValueError(f"Unknown index type: {index_type}") from None   # :485


def calculate_index(img_array, index_type):
    """process-images.py:449-490: float32 normalized difference, eps=1e-10, clip to [-1,1]."""
    if img_array is None or img_array.size == 0:          # :452-453
        return None
    hi_c, lo_c = _band_pair(index_type)
    f = img_array.astype(np.float32)                      # :456
    top = f[:, :, hi_c] - f[:, :, lo_c]                   # :468/:474/:480
    bottom = f[:, :, hi_c] + f[:, :, lo_c] + EPSILON      # :469/:475/:481 (eps is weak -> float32)
    return np.clip(top / bottom, -1, 1)                   # :490


def calculate_index_planes(red, green, nir, index_type):
    """backend-process.py:28-38 (separate float32 planes)."""
    bands = {0: red, 1: green, 2: nir}
    hi_c, lo_c = _band_pair(index_type)
    return np.clip((bands[hi_c] - bands[lo_c]) / (bands[hi_c] + bands[lo_c] + EPSILON), -1, 1)


def calculate_ndvi_f64(img_array):
    """process-ndvi.py:18-31 from the decoded array: float64 NDVI on the *raw* pixels."""
    f = img_array.astype(float)
    nir, red = f[:, :, 2], f[:, :, 0]
    return np.clip((nir - red) / (nir + red + EPSILON), -1, 1)


# --------------------------------------------------------------------------------------
# statistics  (process-images.py:492-513, process-ndvi.py:50-73, :97)
# --------------------------------------------------------------------------------------
def coverage_threshold(index_type):
    return 0.0 if index_type == "NDWI" else 0.2           # :498-504


def feature_name(index_type):
    return "Water" if index_type == "NDWI" else "Vegetation"


def analyze_index(index_array, index_type):
    """process-images.py:492-513: mean / median / min / max / coverage %, exact dict keys."""
    if index_array is None or index_array.size == 0:
        return {}
    thr = coverage_threshold(index_type)
    return {
        f"Mean {index_type}": float(np.mean(index_array)),
        f"Median {index_type}": float(np.median(index_array)),
        f"Min {index_type}": float(np.min(index_array)),
        f"Max {index_type}": float(np.max(index_array)),
        f"{feature_name(index_type)} Coverage (%)": float(np.mean(index_array > thr) * 100),
    }


def analyze_ndvi_statistics(ndvi_array):
    """process-ndvi.py:50-73 (population std, coverage > 0.2)."""
    out = {
        "mean_ndvi": float(np.mean(ndvi_array)),
        "median_ndvi": float(np.median(ndvi_array)),
        "min_ndvi": float(np.min(ndvi_array)),
        "max_ndvi": float(np.max(ndvi_array)),
        "std_ndvi": float(np.std(ndvi_array)),
    }
    out["vegetation_coverage"] = float(np.sum(ndvi_array > 0.2) / ndvi_array.size * 100)
    return out


def index_histogram(index_array, bins=HIST_BINS):
    """process-ndvi.py:97 -- plt.hist(..., bins=50, range=(-1,1)) == np.histogram."""
    return np.histogram(np.asarray(index_array).ravel(), bins=bins, range=HIST_RANGE)[0]


def histogram_edges(bins=HIST_BINS, dtype=np.float32):
    """numpy/lib/_histograms_impl.py:440-447: linspace in the array's dtype."""
    return np.linspace(HIST_RANGE[0], HIST_RANGE[1], bins + 1, endpoint=True, dtype=dtype)


def histogram_bin_by_edges(values, bins=HIST_BINS):
    """Bin index as np.histogram defines it after its +-1 edge correction
    (numpy/lib/_histograms_impl.py:851-863): the last i with edges[i] <= x, last bin closed."""
    values = np.asarray(values)
    edges = histogram_edges(bins, values.dtype)
    idx = np.searchsorted(edges, values, side="right") - 1
    return np.clip(idx, 0, bins - 1)


def full_index_stats(index_array, index_type, bins=HIST_BINS):
    """Everything the fused GPU pass reports for one index map (float64 moments)."""
    x = np.asarray(index_array)
    x64 = x.astype(np.float64).ravel()
    thr = np.float32(coverage_threshold(index_type)) if x.dtype == np.float32 \
        else coverage_threshold(index_type)
    return {
        "count": int(x.size),
        "sum": float(x64.sum()),
        "sumsq": float(np.dot(x64, x64)),
        "min": float(x.min()),
        "max": float(x.max()),
        "count_above": int(np.count_nonzero(x > thr)),
        "hist": index_histogram(x, bins).astype(np.int64),
    }


# --------------------------------------------------------------------------------------
# colormap  (process-images.py:689-695 -- matplotlib restated, RGB bytes unpinned)
# --------------------------------------------------------------------------------------
_ANCHORS_8BIT = {
    # ColorBrewer 11-class diverging schemes, the data behind matplotlib's 'RdYlGn'/'RdYlBu'
    "RdYlGn": ((165, 0, 38), (215, 48, 39), (244, 109, 67), (253, 174, 97), (254, 224, 139),
               (255, 255, 191), (217, 239, 139), (166, 217, 106), (102, 189, 99), (26, 152, 80),
               (0, 104, 55)),
    "RdYlBu": ((165, 0, 38), (215, 48, 39), (244, 109, 67), (253, 174, 97), (254, 224, 144),
               (255, 255, 191), (224, 243, 248), (171, 217, 233), (116, 173, 209), (69, 117, 180),
               (49, 54, 149)),
    "bwr": ((0, 0, 255), (255, 255, 255), (255, 0, 0)),   # process-images.py:956
}


def colormap_name(index_type):
    return "RdYlBu" if index_type == "NDWI" else "RdYlGn"  # process-images.py:689-692


def colormap_lut_float(name, n=256):
    """matplotlib ``LinearSegmentedColormap.from_list(name, anchors, N=256)`` lookup table."""
    anchors = np.asarray(_ANCHORS_8BIT[name], dtype=np.float64) / 255.0
    pos = np.linspace(0.0, 1.0, len(anchors)) * (n - 1)
    xind = (n - 1) * np.linspace(0.0, 1.0, n)
    ind = np.searchsorted(pos, xind)[1:-1]
    lut = np.empty((n, 3), dtype=np.float64)
    for ch in range(3):
        y = anchors[:, ch]
        frac = (xind[1:-1] - pos[ind - 1]) / (pos[ind] - pos[ind - 1])
        lut[1:-1, ch] = frac * (y[ind] - y[ind - 1]) + y[ind - 1]
        lut[0, ch], lut[-1, ch] = y[0], y[-1]
    return np.clip(lut, 0.0, 1.0)


def colormap_lut(name, n=256):
    """(n,3) uint8 = trunc(lut*255), matplotlib's ``bytes=True`` conversion."""
    return (colormap_lut_float(name, n) * 255).astype(np.uint8)


def colormap_index(index_array, vmin=-1.0, vmax=1.0, n=256):
    """Normalize(vmin, vmax) then Colormap.__call__: int(x*N), x==1 -> N-1, clamp.

    float32 maps stay float32 through the normalisation (the in-place ``-=`` / ``/=`` of
    matplotlib's Normalize); tests prove the float64 evaluation gives the same index for
    every value a uint8 pair can produce.
    """
    x = np.array(index_array, copy=True)
    if x.dtype not in (np.float32, np.float64):
        x = x.astype(np.float64)
    x -= x.dtype.type(vmin)
    x /= x.dtype.type(vmax - vmin)
    x *= x.dtype.type(n)
    k = np.where(x < 0, -1, np.where(x >= n, n, np.trunc(x))).astype(np.int64)
    k[x == n] = n - 1
    return np.clip(k, 0, n - 1).astype(np.uint8)


def apply_colormap(index_array, index_type=None, name=None, vmin=-1.0, vmax=1.0):
    name = name or colormap_name(index_type)
    return colormap_lut(name)[colormap_index(index_array, vmin, vmax)]


# --------------------------------------------------------------------------------------
# whole path for one frame (what one "unit" of the benchmark does)
# --------------------------------------------------------------------------------------
def analyze_frame(img_array, indices=INDEX_TYPES, bins=HIST_BINS, want_rgb=True, median=False):
    """WB -> index maps -> statistics (+std, histogram) -> colormap for one HWC frame."""
    wb = fix_white_balance_literal(img_array)
    out = {"wb": wb, "maps": {}, "stats": {}, "rgb": {}}
    for name in indices:
        m = calculate_index(wb, name)
        out["maps"][name] = m
        st = full_index_stats(m, name, bins)
        if median:
            st["median"] = float(np.median(m))
        out["stats"][name] = st
        if want_rgb:
            out["rgb"][name] = apply_colormap(m, name)
    return out


def reference_cpu_path(img_array, indices=INDEX_TYPES, colormap=True, reference=None):
    """The reference's own sequence of NumPy calls for one frame -- what the CPU baseline
    times: fix_white_balance -> calculate_index xk -> analyze_index xk (+ np.std and
    np.histogram(50)) -> colormap gather.
    ``reference``: a module / namespace holding the reference's OWN ``fix_white_balance``, ``calculate_index``
    and ``analyze_index`` (oracle/_ref/process_images.py, made by oracle/build_ref.py); default: this port.
    The colormap gather is always the restatement (matplotlib is absent)."""
    wb_fn = reference.fix_white_balance if reference is not None else fix_white_balance_literal
    index_fn = reference.calculate_index if reference is not None else calculate_index
    stats_fn = reference.analyze_index if reference is not None else analyze_index
    wb = wb_fn(img_array)
    res = {}
    for name in indices:
        m = index_fn(wb, name)
        st = stats_fn(m, name)
        st["std"] = float(np.std(m))
        st["hist"] = index_histogram(m)
        if colormap:
            st["rgb"] = apply_colormap(m, name)
        res[name] = st
    return wb, res


# --------------------------------------------------------------------------------------
# exhaustive (a, b) pair tables -- every value a white-balanced uint8 frame can produce
# --------------------------------------------------------------------------------------
def pair_image():
    """256x256x3 image with pixel (i, j) = (R=j, G=j, N=i)  (SURVEY.md section 8(c) vector (1))."""
    i, j = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    return np.stack([j, j, i], axis=-1)


def pair_tables(bins=HIST_BINS):
    img = pair_image()
    tabs = {}
    for name in INDEX_TYPES:
        m = calculate_index(img, name)
        tabs[name] = {
            "value": m,
            "bin": histogram_bin_by_edges(m, bins).astype(np.uint8),
            "above": (m > np.float32(coverage_threshold(name))),
            "cmap": colormap_index(m),
        }
    return tabs
