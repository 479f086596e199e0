"""Drop-in GPU replacements for the helpers of the reference's ``process-ndvi.py``.

``calculate_ndvi`` keeps the reference's float64 flavour: NDVI of the *raw* file pixels, no
white balance, epsilon 1e-10 not absorbed (process-ndvi.py:18-31).  The matplotlib figure the
reference draws when ``visualize``/``save_path`` is given (:33-46) is figure decoration and
out of scope; with ``save_path`` the per-pixel product -- the 'RdYlGn' colormapped image at
full resolution -- is written instead.
"""
from __future__ import annotations

import os

import numpy as np

from .map_ops import colormap_map, map_statistics, ndvi_float64

__all__ = ["calculate_ndvi", "analyze_ndvi_statistics", "generate_ndvi_report"]


def calculate_ndvi(image_path, save_path=None, visualize=True):
    """process-ndvi.py:5-48 -- float64 NDVI array (-1..1) of an RGNir image file."""
    from PIL import Image
    img = np.array(Image.open(image_path))                           # :18 (decode on host)
    ndvi = ndvi_float64(img)                                         # :21-31 on the GPU
    if save_path:                                                    # :33-44 -> LUT image only
        Image.fromarray(colormap_map(ndvi.astype(np.float32), "RdYlGn", -1.0, 1.0)).save(save_path)
    return ndvi


def analyze_ndvi_statistics(ndvi_array):
    """process-ndvi.py:50-73 -- mean / median / min / max / std / vegetation coverage (>0.2).

    The statistics keep the array's dtype as NumPy does: a float64 map (``calculate_ndvi``) goes through the
    float64 kernels (K4d / K3d) -- median, min, max and the ``> 0.2`` count are exact, mean / std within 1e-12
    relative (summation order); a float32 map through K4 / K3.
    """
    arr = np.asarray(ndvi_array)
    st = map_statistics(arr, threshold=0.2, median=True)
    return {
        "mean_ndvi": st["mean"],
        "median_ndvi": st["median"],
        "min_ndvi": st["min"],
        "max_ndvi": st["max"],
        "std_ndvi": st["std"],
        "vegetation_coverage": st["coverage_pct"],                   # :69-71
    }


def generate_ndvi_report(image_path, output_dir):
    """process-ndvi.py:75-110 -- NDVI map + statistics + 50-bin histogram + text report.

    Writes ``ndvi_visualization.png`` (colormapped map), ``ndvi_histogram.npy`` (the counts the
    reference plots with ``plt.hist(bins=50, range=(-1, 1))``, :97) and ``ndvi_statistics.txt``
    (same format, :105-108).  Returns ``(ndvi_array, stats)`` like the reference.
    """
    os.makedirs(output_dir, exist_ok=True)
    ndvi_array = calculate_ndvi(image_path, os.path.join(output_dir, "ndvi_visualization.png"),
                                visualize=False)
    full = map_statistics(ndvi_array, threshold=0.2, bins=50, median=True)
    stats = {
        "mean_ndvi": full["mean"], "median_ndvi": full["median"], "min_ndvi": full["min"],
        "max_ndvi": full["max"], "std_ndvi": full["std"], "vegetation_coverage": full["coverage_pct"],
    }
    np.save(os.path.join(output_dir, "ndvi_histogram.npy"), full["hist"])
    with open(os.path.join(output_dir, "ndvi_statistics.txt"), "w") as f:
        f.write("NDVI Statistics:\n")
        for key, value in stats.items():
            f.write(f"{key}: {value:.4f}\n")
    return ndvi_array, stats
