"""Drop-in GPU replacements for the per-pixel helpers of the reference's ``process-images.py``.

Same names, argument meaning, return types and error behaviour as the reference functions
(``/root/reference/process-images.py``): ``None``/empty in -> ``None`` out (or ``{}``), unknown
index -> ``ValueError("Unknown index type: ...")``, 2-D input -> ``IndexError``.  Arrays go
in and come out as host NumPy arrays exactly as the Streamlit front end expects
(SURVEY.md section 8(b)); every operation in between runs in the sm_100a kernels.  Importing this
module does not need a GPU; calling a helper without one raises (no CPU fallback).
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from .engine import DEFAULT_BINS, INDEX_TYPES, Engine, get_engine

__all__ = ["preprocess_large_image", "fix_white_balance", "calculate_index", "analyze_index", "analyze_frame",
           "calculate_index_statistics_by_timeframe", "create_index_visualization",
           "create_change_detection_visualization", "index_change"]


def _feature_name(index_type: str) -> str:
    return "Water" if index_type == "NDWI" else "Vegetation"        # process-images.py:500-504


def _check_index_type(index_type: str) -> None:
    if index_type not in INDEX_TYPES:
        raise ValueError(f"Unknown index type: {index_type}")       # process-images.py:485


def preprocess_large_image(img_array, max_dimension=1024):
    """process-images.py:398-422 -- shrink a frame so that max(h, w) == max_dimension with Pillow's
    LANCZOS filter; frames that are already small enough are returned as they are (same object).

    GPU path: K10h + K10v, bit-identical to ``Image.fromarray(img).resize(..., LANCZOS)``.
    """
    if img_array is None or img_array.size == 0:                     # :401-402
        return None
    h, w = img_array.shape[:2]
    target = Engine.preprocess_target(h, w, max_dimension)           # :404-416
    if target is None:
        return img_array
    arr = np.asarray(img_array)
    if arr.dtype != np.uint8 or arr.ndim not in (2, 3) or (arr.ndim == 3 and arr.shape[2] not in (3, 4)):
        # Image.fromarray accepts a few more layouts (float / int32 single-band images in their own precision) and
        # rejects the rest -- uint16 RGB among them -- with this TypeError; gray, RGB and RGBA (np.array(Image.open(png)),
        # resized with premultiplied alpha like Pillow) are what the app passes
        raise TypeError(f"Cannot handle this data type: {arr.shape[2:] or (1,)}, {arr.dtype.str}")
    return get_engine().resize_batch([arr], target[0], target[1])[0]


def fix_white_balance(img_array):
    """process-images.py:424-447 -- per-channel 2nd/98th percentile stretch, uint8 out.

    GPU path: K1 histogram -> K1b percentiles + LUT -> K2 (white-balance output only).
    """
    if img_array is None or img_array.size == 0:                     # :427-428
        return None
    return get_engine().analyze_frame(img_array, outputs=("wb",))["wb"]


def calculate_index(img_array, index_type):
    """process-images.py:449-490 -- float32 NDVI / GNDVI / NDWI map of an HxWx3 frame."""
    if img_array is None or img_array.size == 0:                     # :452-453
        return None
    _check_index_type(index_type)
    if np.asarray(img_array).dtype != np.uint8:
        from .map_ops import index_generic                            # astype(float32) chain, any dtype
        return index_generic(img_array, index_type)
    res = get_engine().analyze_frame(img_array, outputs=("maps",), white_balance=False,
                                     indices=(index_type,))
    return res["maps"][index_type]


def analyze_frame(img_array, indices=INDEX_TYPES, bins: int = DEFAULT_BINS, white_balance: bool = True,
                  outputs=("wb", "maps", "rgb", "stats")) -> Optional[dict]:
    """Fused entry point (SURVEY.md section 8(b) "Ownership"): one trip to the GPU returns the
    white-balanced frame, the requested index maps, their colormapped RGB images and all
    statistics, instead of three separate helper calls."""
    if img_array is None or img_array.size == 0:
        return None
    for name in indices:
        _check_index_type(name)
    return get_engine().analyze_frame(img_array, outputs=outputs, white_balance=white_balance,
                                      indices=tuple(indices), bins=bins)


def analyze_index(index_array, index_type):
    """process-images.py:492-513 -- mean / median / min / max / coverage of an index map."""
    if index_array is None or index_array.size == 0:                 # :495-496
        return {}
    from .map_ops import map_statistics                               # GPU statistics of a float map
    thr = 0.0 if index_type == "NDWI" else 0.2                        # :498-504
    st = map_statistics(index_array, threshold=thr, median=True)
    return {
        f"Mean {index_type}": st["mean"],
        f"Median {index_type}": st["median"],
        f"Min {index_type}": st["min"],
        f"Max {index_type}": st["max"],
        f"{_feature_name(index_type)} Coverage (%)": st["coverage_pct"],
    }


def calculate_index_statistics_by_timeframe(image_data_list, index_type):
    """process-images.py:619-667 -- per-frame statistics of a time series -> DataFrame.

    All frames that still need white balance and share a shape go through the GPU as one
    batch (Pass 1 + LUT + Pass 2 with statistics only); frames with a cached
    ``'corrected_array'`` skip Pass 1 (identity LUT), as in the reference (:638-641).
    """
    import pandas as pd
    from .map_ops import frame_statistics_rows
    if image_data_list:                      # the reference only meets an unknown index inside its per-frame loop
        _check_index_type(index_type)
    rows = frame_statistics_rows(image_data_list, index_type)
    return pd.DataFrame(rows)


def create_index_visualization(index_array, index_type):
    """process-images.py:669-716 -- colour visualisation of an index map as a PIL image.

    The reference renders a decorated matplotlib figure; the per-pixel product north_star
    asks for is the colormap lookup itself ('RdYlBu' for NDWI else 'RdYlGn', vmin=-1,
    vmax=1, :689-695), returned here at full resolution as an RGB ``PIL.Image``.
    """
    if index_array is None or index_array.size == 0:                 # :672-673
        return None
    from PIL import Image
    from .map_ops import colormap_map
    cmap = "RdYlBu" if index_type == "NDWI" else "RdYlGn"
    return Image.fromarray(colormap_map(index_array, cmap, -1.0, 1.0))


def index_change(early_img, late_img, index_type, white_balance: bool = True):
    """Per-pixel products of change detection (process-images.py:885-989): index maps of both
    dates, ``late - early`` and its 'bwr' image over [-0.5, 0.5].  Frames are white-balanced
    first unless ``white_balance`` is False (the reference uses the cached 'corrected_array').
    Image registration (``align_images``, :515-565) is out of scope: pass aligned frames."""
    _check_index_type(index_type)
    from .map_ops import index_change as _change
    if white_balance:
        wb = get_engine().analyze_batch([early_img, late_img], outputs=("wb",))
        early_img, late_img = wb[0]["wb"], wb[1]["wb"]
    return _change(early_img, late_img, index_type)


def create_change_detection_visualization(image_pair, index_type):
    """process-images.py:885-989 -- returns the 'bwr' change image (PIL) instead of the reference's
    three-panel matplotlib figure; ``None`` for an invalid pair as in the reference (:888-889)."""
    if not image_pair or len(image_pair) != 2:
        return None
    from PIL import Image
    frames = []
    for d in image_pair:
        cached = d.get("corrected_array")
        frames.append(cached if cached is not None else fix_white_balance(d["array"]))
    if frames[0] is None or frames[1] is None:
        return None
    return Image.fromarray(index_change(frames[0], frames[1], index_type, white_balance=False)["rgb"])
