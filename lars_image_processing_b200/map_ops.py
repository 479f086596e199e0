"""GPU operations on caller-supplied index maps / raw frames next to the fused pass.

These back the reference helpers whose input is *not* a raw uint8 frame: ``analyze_index`` and
``analyze_ndvi_statistics`` receive a float map, ``create_index_visualization`` colours one,
``calculate_ndvi`` (process-ndvi.py) wants a float64 NDVI of the raw pixels and
``backend-process.py``'s ``calculate_index`` takes separate float32 planes.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib
from ._lib import INDEX_STATS_DTYPE, MAP_STATS_F64_DTYPE, LarsError, check
from .engine import DEFAULT_BINS, INDEX_TYPES, Engine, get_engine, stats_records_to_dicts


def _to_device_f32(eng: Engine, arr: np.ndarray, s) -> torch.Tensor:
    flat = np.ascontiguousarray(arr, dtype=np.float32).reshape(-1)
    with torch.cuda.stream(s):
        dev = torch.empty(flat.size, dtype=torch.float32, device=eng.device)
        eng._staged_h2d(dev.view(torch.uint8), flat.view(np.uint8), s)
    return dev


def _to_host(dev: torch.Tensor) -> torch.Tensor:
    """Device -> pinned host copy on the current stream (the pinned block comes from PyTorch's caching
    host allocator and returns to it when the NumPy view handed to the caller dies).  Copies into
    pageable memory run at a fraction of the PCIe rate."""
    host = torch.empty(dev.shape, dtype=dev.dtype, pin_memory=True)
    host.copy_(dev, non_blocking=True)
    return host


def device_map_statistics(eng: Engine, dev_map: torch.Tensor, n: int, threshold: float,
                          bins: int = DEFAULT_BINS, median: bool = False, stream=None):
    """Statistics (and optionally the exact median) of one device-resident float32 map.
    Returns (stats tensor [576] uint8, median tensor [3] float32 or None); nothing is synced."""
    lib = eng.lib
    s = stream or eng.stream()
    with torch.cuda.stream(s):
        stats = torch.empty(INDEX_STATS_DTYPE.itemsize, dtype=torch.uint8, device=eng.device)
        ws_bytes = int(lib.lars_map_stats_workspace_bytes(1))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=eng.device)
        med = None
        with torch.cuda.device(eng.device):
            check(lib.lars_map_stats_f32(dev_map.data_ptr(), 1, n, n, bins, float(threshold),
                                         stats.data_ptr(), ws.data_ptr(), ws_bytes, s.cuda_stream),
                  "lars_map_stats_f32")
            if median:
                med = torch.empty(3, dtype=torch.float32, device=eng.device)
                sel_bytes = int(lib.lars_select_workspace_bytes())
                sel_ws = torch.empty(sel_bytes, dtype=torch.uint8, device=eng.device)
                check(lib.lars_select_f32(dev_map.data_ptr(), n, (n - 1) // 2, n // 2, med.data_ptr(),
                                          sel_ws.data_ptr(), sel_bytes, s.cuda_stream), "lars_select_f32")
    return stats, med


def device_map_median(eng: Engine, dev_map: torch.Tensor, n: int, stream=None) -> torch.Tensor:
    """Exact np.median of one device-resident float32 map -> [3] float32 tensor
    (x[(n-1)//2], x[n//2], their float32 mean); nothing is synced."""
    lib = eng.lib
    s = stream or eng.stream()
    with torch.cuda.stream(s):
        med = torch.empty(3, dtype=torch.float32, device=eng.device)
        sel_bytes = int(lib.lars_select_workspace_bytes())
        sel_ws = torch.empty(sel_bytes, dtype=torch.uint8, device=eng.device)
        with torch.cuda.device(eng.device):
            check(lib.lars_select_f32(dev_map.data_ptr(), n, (n - 1) // 2, n // 2, med.data_ptr(),
                                      sel_ws.data_ptr(), sel_bytes, s.cuda_stream), "lars_select_f32")
    return med


def _map_statistics_f64(arr: np.ndarray, threshold: float, bins: int, median: bool) -> dict:
    """The float64 flavour (process-ndvi.py:60-71, :97): K4d + K3d keep the array's own arithmetic -- float64
    min / max / median, `x > threshold` in float64, np.histogram with float64 edges."""
    eng = get_engine()
    lib = eng.lib
    s = eng.stream()
    flat = np.ascontiguousarray(arr, dtype=np.float64).reshape(-1)
    n = flat.size
    with torch.cuda.stream(s):
        dev = torch.empty(n, dtype=torch.float64, device=eng.device)
        eng._staged_h2d(dev.view(torch.uint8), flat.view(np.uint8), s)
        stats = torch.empty(MAP_STATS_F64_DTYPE.itemsize, dtype=torch.uint8, device=eng.device)
        ws_bytes = int(lib.lars_map_stats_f64_workspace_bytes())
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=eng.device)
        med = None
        with torch.cuda.device(eng.device):
            check(lib.lars_map_stats_f64(dev.data_ptr(), n, bins, float(threshold), stats.data_ptr(), ws.data_ptr(),
                                         ws_bytes, s.cuda_stream), "lars_map_stats_f64")
            if median:
                med = torch.empty(3, dtype=torch.float64, device=eng.device)
                sel_bytes = int(lib.lars_select_f64_workspace_bytes())
                sel_ws = torch.empty(sel_bytes, dtype=torch.uint8, device=eng.device)
                check(lib.lars_select_f64(dev.data_ptr(), n, (n - 1) // 2, n // 2, med.data_ptr(), sel_ws.data_ptr(),
                                          sel_bytes, s.cuda_stream), "lars_select_f64")
        h_stats = stats.cpu()
        h_med = med.cpu() if med is not None else None
    s.synchronize()
    rec = h_stats.numpy().view(MAP_STATS_F64_DTYPE)[0]
    cnt = int(rec["count"])
    out = {
        "count": cnt, "mean": float(rec["mean"]), "std": float(rec["std"]),
        "min": float(rec["min"]), "max": float(rec["max"]),
        "count_above": int(rec["count_above"]),
        "coverage_pct": float(rec["count_above"]) / cnt * 100.0,
        "sum": float(rec["sum"]), "sumsq": float(rec["sumsq"]),
        "hist": np.array(rec["hist"][:bins], dtype=np.int64),
    }
    if h_med is not None:
        out["median"] = float(h_med[2])
        out["middle_values"] = (float(h_med[0]), float(h_med[1]))
    return out


def map_statistics(index_array, threshold: float = 0.2, bins: int = DEFAULT_BINS, median: bool = True) -> dict:
    """mean / std / min / max / coverage / histogram (/ exact median) of a host float map.  A float64 array is
    reduced in float64 (the reference's NumPy calls keep the array's dtype); everything else as float32."""
    arr = np.asarray(index_array)
    if arr.size == 0:
        return {}
    if arr.dtype == np.float64:
        return _map_statistics_f64(arr, threshold, bins, median)
    eng = get_engine()
    s = eng.stream()
    dev = _to_device_f32(eng, arr, s)
    stats, med = device_map_statistics(eng, dev, arr.size, threshold, bins, median, s)
    with torch.cuda.stream(s):
        h_stats = stats.cpu()
        h_med = med.cpu() if med is not None else None
    s.synchronize()
    rec = h_stats.numpy().view(INDEX_STATS_DTYPE)[0]
    n = int(rec["count"])
    out = {
        "count": n, "mean": float(rec["mean"]), "std": float(rec["std"]),
        "min": float(rec["min"]), "max": float(rec["max"]),
        "count_above": int(rec["count_above"]),
        "coverage_pct": float(rec["count_above"]) / n * 100.0,
        "sum": float(rec["sum"]), "sumsq": float(rec["sumsq"]),
        "hist": np.array(rec["hist"][:bins], dtype=np.int64),
    }
    if h_med is not None:
        out["median"] = float(h_med[2])
        out["middle_values"] = (float(h_med[0]), float(h_med[1]))
    return out


def colormap_map(index_array, cmap: str = "RdYlGn", vmin: float = -1.0, vmax: float = 1.0) -> np.ndarray:
    """HxW float map -> HxWx3 uint8 through the colormap LUT (Normalize(vmin, vmax))."""
    arr = np.asarray(index_array)
    eng = get_engine()
    lib = eng.lib
    s = eng.stream()
    dev = _to_device_f32(eng, arr, s)
    with torch.cuda.stream(s):
        rgb = torch.empty(arr.size * 3, dtype=torch.uint8, device=eng.device)
        with torch.cuda.device(eng.device):
            check(lib.lars_colormap_f32(dev.data_ptr(), arr.size, _lib.CMAP_IDS[cmap], float(vmin), float(vmax),
                                        rgb.data_ptr(), s.cuda_stream), "lars_colormap_f32")
        host = _to_host(rgb)
    s.synchronize()
    return host.numpy().reshape(arr.shape + (3,))


def ndvi_float64(img_array) -> np.ndarray:
    """float64 NDVI of the raw pixels (process-ndvi.py:18-31), no white balance."""
    img = np.ascontiguousarray(img_array)
    if img.ndim != 3 or img.shape[2] < 3:
        raise IndexError("expected an HxWxC frame with at least 3 channels")
    if img.dtype != np.uint8:
        raise LarsError(f"dtype {img.dtype} is not supported by the uint8 NDVI path")
    eng = get_engine()
    lib = eng.lib
    s = eng.stream()
    n = img.shape[0] * img.shape[1]
    with torch.cuda.stream(s):
        dev = torch.empty(img.size, dtype=torch.uint8, device=eng.device)
        dev.copy_(torch.from_numpy(img.reshape(-1)), non_blocking=True)
        out = torch.empty(n, dtype=torch.float64, device=eng.device)
        with torch.cuda.device(eng.device):
            check(lib.lars_ndvi_f64_u8(dev.data_ptr(), n, img.shape[2], out.data_ptr(), s.cuda_stream),
                  "lars_ndvi_f64_u8")
        host = _to_host(out)
    s.synchronize()
    return host.numpy().reshape(img.shape[:2])


def index_from_planes(hi_plane, lo_plane) -> np.ndarray:
    """clip((hi - lo) / (hi + lo + 1e-10), -1, 1) on float32 planes (backend-process.py:28-38)."""
    hi = np.asarray(hi_plane)
    lo = np.asarray(lo_plane)
    if hi.shape != lo.shape:
        raise ValueError("planes must have the same shape")
    eng = get_engine()
    lib = eng.lib
    s = eng.stream()
    dh, dl = _to_device_f32(eng, hi, s), _to_device_f32(eng, lo, s)
    with torch.cuda.stream(s):
        out = torch.empty(hi.size, dtype=torch.float32, device=eng.device)
        with torch.cuda.device(eng.device):
            check(lib.lars_index_planes_f32(dh.data_ptr(), dl.data_ptr(), hi.size, out.data_ptr(), s.cuda_stream),
                  "lars_index_planes_f32")
        host = _to_host(out)
    s.synchronize()
    return host.numpy().reshape(hi.shape)


def frame_statistics_rows(image_data_list, index_type: str) -> List[dict]:
    """Rows of calculate_index_statistics_by_timeframe (process-images.py:633-663).

    Frames are grouped by (shape, needs-white-balance) and each group goes through the fused
    pass as one batch with statistics only; the exact median comes from the radix select on
    the index map of each frame.
    """
    eng = get_engine()
    i_idx = INDEX_TYPES.index(index_type)
    feature = "Water" if index_type == "NDWI" else "Vegetation"
    rows: List[Optional[dict]] = [None] * len(image_data_list)
    groups: Dict[tuple, list] = {}
    for pos, img_data in enumerate(image_data_list):
        cached = img_data.get("corrected_array") if isinstance(img_data, dict) else None
        arr = cached if cached is not None else img_data["array"]
        arr = eng._check_frame(arr)
        groups.setdefault((arr.shape, cached is None), []).append((pos, arr))
    s = eng.stream()
    for (shape, needs_wb), members in groups.items():
        dev = eng.upload([a for _, a in members], stream=s)
        res = eng.process_device(dev, outputs=("maps", "stats"), white_balance=needs_wb,
                                 indices=(index_type,), stream=s)
        n = dev.n_pixels
        meds = []
        for k in range(len(members)):       # the fused pass already produced every other statistic
            meds.append(device_map_median(eng, res.maps[i_idx, k], n, stream=s))
        with torch.cuda.stream(s):
            h_stats = res.stats.cpu()
            h_meds = torch.stack(meds).cpu()
        s.synchronize()
        rec = h_stats.numpy().view(INDEX_STATS_DTYPE).reshape(len(members), 3)
        per = stats_records_to_dicts(rec, res.bins)
        for k, (pos, _) in enumerate(members):
            st = per[k][index_type]
            rows[pos] = {
                "Date": image_data_list[pos]["metadata"]["upload_date"],
                "Mean": st["mean"],
                "Median": float(h_meds[k, 2]),
                "Min": st["min"],
                "Max": st["max"],
                f"{feature} Coverage (%)": st["coverage_pct"],
            }
    return [r for r in rows if r is not None]


def index_generic(img_array, index_type: str) -> np.ndarray:
    """calculate_index for frames that are not uint8 (process-images.py:456-490 applies
    astype(float32) to any input).  uint16 / float32 / float64 convert on the GPU; other
    numeric dtypes are cast to float32 first exactly as the reference's first line does."""
    img = np.asarray(img_array)
    if img.ndim != 3:
        raise IndexError(f"too many indices for array: expected an HxWxC frame, got {img.ndim}-dimensional input")
    if img.shape[2] < 3:
        raise IndexError(f"index 2 is out of bounds for axis 2 with size {img.shape[2]}")
    if img.dtype.name not in ("uint16", "float32", "float64"):
        img = img.astype(np.float32)                                  # process-images.py:456
    img = np.ascontiguousarray(img)
    eng = get_engine()
    s = eng.stream()
    n = img.shape[0] * img.shape[1]
    with torch.cuda.stream(s):
        dev = torch.empty(img.nbytes, dtype=torch.uint8, device=eng.device)
        dev.copy_(torch.from_numpy(img.reshape(-1).view(np.uint8)), non_blocking=True)
        out = torch.empty(n, dtype=torch.float32, device=eng.device)
        with torch.cuda.device(eng.device):
            check(eng.lib.lars_index_hwc(dev.data_ptr(), _lib.DTYPE_IDS[img.dtype.name], n, img.shape[2],
                                         _lib.INDEX_IDS[index_type], out.data_ptr(), s.cuda_stream), "lars_index_hwc")
        host = _to_host(out)
    s.synchronize()
    return host.numpy().reshape(img.shape[:2])


def index_change(early_wb, late_wb, index_type: str, vmin: float = -0.5, vmax: float = 0.5) -> dict:
    """Change detection between two white-balanced uint8 frames of equal shape
    (process-images.py:908-923, :956): {'early', 'late', 'diff', 'rgb'}."""
    e = np.ascontiguousarray(early_wb)
    l = np.ascontiguousarray(late_wb)
    if e.shape != l.shape or e.dtype != np.uint8 or l.dtype != np.uint8 or e.ndim != 3 or e.shape[2] < 3:
        raise ValueError("change detection needs two uint8 HxWxC frames of equal shape")
    eng = get_engine()
    s = eng.stream()
    n = e.shape[0] * e.shape[1]
    with torch.cuda.stream(s):
        de = torch.empty(e.size, dtype=torch.uint8, device=eng.device)
        dl = torch.empty(l.size, dtype=torch.uint8, device=eng.device)
        de.copy_(torch.from_numpy(e.reshape(-1)), non_blocking=True)
        dl.copy_(torch.from_numpy(l.reshape(-1)), non_blocking=True)
        maps = torch.empty((3, n), dtype=torch.float32, device=eng.device)
        rgb = torch.empty(n * 3, dtype=torch.uint8, device=eng.device)
        with torch.cuda.device(eng.device):
            check(eng.lib.lars_index_change_u8(de.data_ptr(), dl.data_ptr(), n, e.shape[2], _lib.INDEX_IDS[index_type],
                                               float(vmin), float(vmax), maps[0].data_ptr(), maps[1].data_ptr(),
                                               maps[2].data_ptr(), rgb.data_ptr(), s.cuda_stream), "lars_index_change_u8")
        h_maps, h_rgb = _to_host(maps), _to_host(rgb)
    s.synchronize()
    hw = e.shape[:2]
    return {"early": h_maps[0].numpy().reshape(hw), "late": h_maps[1].numpy().reshape(hw),
            "diff": h_maps[2].numpy().reshape(hw), "rgb": h_rgb.numpy().reshape(hw + (3,))}
