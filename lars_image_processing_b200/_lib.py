"""ctypes binding of liblars_b200.so (the C ABI in include/lars_b200.h).

There is no fallback: if the CUDA library cannot be loaded the package raises.
ctypes releases the GIL for the duration of every call, so Streamlit's per-session threads
("Threading") can enter the library concurrently.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

from . import build as _build

LARS_OK = 0
NUM_INDICES = 3
MAX_BINS = 64
PIXEL_GROUP = 16
STRETCH_U16_BYTES = 2064   # sizeof(lars_stretch_u16)
CMAP_IDS = {"RdYlGn": 0, "RdYlBu": 1, "bwr": 2}
INDEX_IDS = {"NDVI": 0, "GNDVI": 1, "NDWI": 2}
WB_CHAINS = {"images": 0, "rgn": 1}     # LARS_WB_CHAIN_*: process-images.py:437-441 / process-rgn.py:25-44
DTYPE_IDS = {"uint8": 0, "uint16": 1, "float32": 2, "float64": 3}

EXPORTED_SYMBOLS = (
    "lars_init", "lars_shutdown", "lars_last_error", "lars_abi_version", "lars_sm_count",
    "lars_colormap_table", "lars_histogram_edges_f32",
    "lars_wb_hist_u8", "lars_wb_lut_build_u8", "lars_wb_lut_build_u8_chain",
    "lars_wb_peer_buffer_bytes", "lars_wb_lut_build_u8_peers",
    "lars_fused_workspace_bytes", "lars_fused_index_u8",
    "lars_map_stats_workspace_bytes", "lars_map_stats_f32", "lars_select_workspace_bytes",
    "lars_select_f32", "lars_map_stats_f64_workspace_bytes", "lars_map_stats_f64", "lars_select_f64_workspace_bytes",
    "lars_select_f64", "lars_colormap_f32", "lars_ndvi_f64_u8", "lars_index_planes_f32",
    "lars_stats_merge",
    "lars_wb_u16_workspace_bytes", "lars_wb_stretch_build_u16", "lars_wb_stretch_build_u16_staged",
    "lars_fused_index_u16",
    "lars_index_hwc", "lars_index_change_u8",
    "lars_resize_plan_lanczos", "lars_resize_tables_lanczos", "lars_resize_lanczos_u8", "lars_rgba_alpha_u8",
    "lars_tiff_probe", "lars_tiff_read", "lars_tiff_read_region", "lars_png_probe", "lars_png_read",
    "lars_tiff_lzw_chunks", "lars_lzw_decode_device", "lars_tiff_post_device",
    "lars_tiff_deflate_chunks", "lars_inflate_decode_device", "lars_untile_device",
)


class LarsError(RuntimeError):
    """A C-ABI call returned a negative status."""


class FusedArgs(C.Structure):
    """Mirror of ``lars_fused_args`` (include/lars_b200.h)."""
    _fields_ = [
        ("struct_bytes", C.c_uint32),
        ("n_frames", C.c_int32),
        ("channels", C.c_int32),
        ("bins", C.c_int32),
        ("n_pixels", C.c_int64),
        ("src", C.c_void_p),
        ("src_frame_stride", C.c_int64),
        ("wb_lut", C.c_void_p),
        ("lut_frame_stride", C.c_int64),
        ("wb_out", C.c_void_p),
        ("wb_frame_stride", C.c_int64),
        ("maps", C.c_void_p * NUM_INDICES),
        ("map_frame_stride", C.c_int64),
        ("rgb", C.c_void_p * NUM_INDICES),
        ("rgb_frame_stride", C.c_int64),
        ("cmap", C.c_int32 * NUM_INDICES),
        ("thresholds", C.c_float * NUM_INDICES),
        ("stats", C.c_void_p),
        ("workspace", C.c_void_p),
        ("workspace_bytes", C.c_size_t),
    ]


class ResizePlan(C.Structure):
    """Mirror of ``lars_resize_plan`` (include/lars_b200.h)."""
    _fields_ = [(n, C.c_int32) for n in (
        "in_h", "in_w", "out_h", "out_w", "channels", "need_h", "need_v", "ksize_h", "ksize_v",
        "row_first", "row_count", "xo_tile", "plane_words", "out_pitch", "groups_h", "groups_v",
        "mma_ksteps", "mma_plane_words")] + [
        ("mma_table_offset", C.c_uint64), ("table_bytes", C.c_uint64), ("temp_frame_bytes", C.c_uint64)]


class TiffInfo(C.Structure):
    """Mirror of ``lars_tiff_info`` (include/lars_b200.h)."""
    _fields_ = [(n, C.c_int32) for n in (
        "width", "height", "samples_per_pixel", "bits_per_sample", "big_endian", "compression",
        "planar_config", "photometric", "sample_format", "rows_per_strip", "n_strips",
        "strip_offsets_type", "strip_counts_type", "predictor")] + [
        ("strip_offsets_pos", C.c_uint64), ("strip_counts_pos", C.c_uint64), ("frame_bytes", C.c_uint64)] + [
        (n, C.c_int32) for n in ("tile_width", "tile_length", "tiles_across", "tiles_down", "bigtiff", "reserved")]


class PngInfo(C.Structure):
    """Mirror of ``lars_png_info`` (include/lars_b200.h)."""
    _fields_ = [(n, C.c_int32) for n in (
        "width", "height", "channels", "bit_depth", "color_type", "interlace", "n_idat", "reserved")] + [
        ("idat_bytes", C.c_uint64), ("frame_bytes", C.c_uint64)]


# numpy view of ``lars_lzw_chunk`` (24 bytes)
LZW_CHUNK_DTYPE = np.dtype([("src_offset", "<u8"), ("dst_offset", "<u8"), ("src_bytes", "<u4"), ("dst_bytes", "<u4")])
assert LZW_CHUNK_DTYPE.itemsize == 24

# numpy view of ``lars_index_stats`` (576 bytes)
INDEX_STATS_DTYPE = np.dtype([
    ("count", "<u8"), ("count_above", "<u8"),
    ("sum", "<f8"), ("sumsq", "<f8"), ("mean", "<f8"), ("std", "<f8"),
    ("min", "<f4"), ("max", "<f4"), ("threshold", "<f4"), ("bins", "<u4"),
    ("hist", "<u8", (MAX_BINS,)),
])
assert INDEX_STATS_DTYPE.itemsize == 576

# numpy view of ``lars_map_record_f64`` (592 bytes)
MAP_STATS_F64_DTYPE = np.dtype([
    ("count", "<u8"), ("count_above", "<u8"),
    ("sum", "<f8"), ("sumsq", "<f8"), ("mean", "<f8"), ("std", "<f8"),
    ("min", "<f8"), ("max", "<f8"), ("threshold", "<f8"), ("bins", "<u4"), ("has_nan", "<u4"),
    ("hist", "<u8", (MAX_BINS,)),
])
assert MAP_STATS_F64_DTYPE.itemsize == 592

_lock = threading.Lock()
_lib = None


def _declare(lib):
    vp, i32, i64, f64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_double, C.c_float
    lib.lars_init.argtypes = [C.c_int]
    lib.lars_init.restype = C.c_int
    lib.lars_shutdown.argtypes = []
    lib.lars_shutdown.restype = C.c_int
    lib.lars_last_error.argtypes = []
    lib.lars_last_error.restype = C.c_char_p
    lib.lars_abi_version.argtypes = []
    lib.lars_abi_version.restype = C.c_int
    lib.lars_sm_count.argtypes = []
    lib.lars_sm_count.restype = C.c_int
    lib.lars_colormap_table.argtypes = [C.c_int, vp]
    lib.lars_colormap_table.restype = C.c_int
    lib.lars_histogram_edges_f32.argtypes = [C.c_int, vp]
    lib.lars_histogram_edges_f32.restype = C.c_int
    lib.lars_wb_hist_u8.argtypes = [vp, i32, i64, i32, i64, vp, i32, vp]
    lib.lars_wb_hist_u8.restype = C.c_int
    lib.lars_wb_lut_build_u8.argtypes = [vp, i32, f64, f64, vp, vp, vp]
    lib.lars_wb_lut_build_u8.restype = C.c_int
    lib.lars_wb_lut_build_u8_chain.argtypes = [vp, i32, f64, f64, i32, vp, vp, vp]
    lib.lars_wb_lut_build_u8_chain.restype = C.c_int
    lib.lars_wb_peer_buffer_bytes.argtypes = [i32]
    lib.lars_wb_peer_buffer_bytes.restype = C.c_size_t
    lib.lars_wb_lut_build_u8_peers.argtypes = [vp, vp, i32, i32, C.c_uint32, f64, f64, i32, vp, vp, vp, vp]
    lib.lars_wb_lut_build_u8_peers.restype = C.c_int
    lib.lars_fused_workspace_bytes.argtypes = [i32]
    lib.lars_fused_workspace_bytes.restype = C.c_size_t
    lib.lars_fused_index_u8.argtypes = [C.POINTER(FusedArgs), vp]
    lib.lars_fused_index_u8.restype = C.c_int
    lib.lars_map_stats_workspace_bytes.argtypes = [i32]
    lib.lars_map_stats_workspace_bytes.restype = C.c_size_t
    lib.lars_map_stats_f32.argtypes = [vp, i32, i64, i64, i32, f32, vp, vp, C.c_size_t, vp]
    lib.lars_map_stats_f32.restype = C.c_int
    lib.lars_select_workspace_bytes.argtypes = []
    lib.lars_select_workspace_bytes.restype = C.c_size_t
    lib.lars_select_f32.argtypes = [vp, i64, C.c_uint64, C.c_uint64, vp, vp, C.c_size_t, vp]
    lib.lars_select_f32.restype = C.c_int
    lib.lars_map_stats_f64_workspace_bytes.argtypes = []
    lib.lars_map_stats_f64_workspace_bytes.restype = C.c_size_t
    lib.lars_map_stats_f64.argtypes = [vp, i64, i32, f64, vp, vp, C.c_size_t, vp]
    lib.lars_map_stats_f64.restype = C.c_int
    lib.lars_select_f64_workspace_bytes.argtypes = []
    lib.lars_select_f64_workspace_bytes.restype = C.c_size_t
    lib.lars_select_f64.argtypes = [vp, i64, C.c_uint64, C.c_uint64, vp, vp, C.c_size_t, vp]
    lib.lars_select_f64.restype = C.c_int
    lib.lars_rgba_alpha_u8.argtypes = [vp, i32, i64, i64, i32, vp]
    lib.lars_rgba_alpha_u8.restype = C.c_int
    lib.lars_colormap_f32.argtypes = [vp, i64, i32, f32, f32, vp, vp]
    lib.lars_colormap_f32.restype = C.c_int
    lib.lars_ndvi_f64_u8.argtypes = [vp, i64, i32, vp, vp]
    lib.lars_ndvi_f64_u8.restype = C.c_int
    lib.lars_index_planes_f32.argtypes = [vp, vp, i64, vp, vp]
    lib.lars_index_planes_f32.restype = C.c_int
    lib.lars_stats_merge.argtypes = [vp, i32, vp, vp]
    lib.lars_stats_merge.restype = C.c_int
    lib.lars_wb_u16_workspace_bytes.argtypes = [i32]
    lib.lars_wb_u16_workspace_bytes.restype = C.c_size_t
    lib.lars_wb_stretch_build_u16.argtypes = [vp, i32, i64, i32, i64, f64, f64, vp, vp, vp, C.c_size_t, i32, vp]
    lib.lars_wb_stretch_build_u16.restype = C.c_int
    lib.lars_wb_stretch_build_u16_staged.argtypes = [vp, i32, i64, i32, i64, f64, f64, vp, vp, vp, C.c_size_t, i32, i32, vp]
    lib.lars_wb_stretch_build_u16_staged.restype = C.c_int
    lib.lars_fused_index_u16.argtypes = [C.POINTER(FusedArgs), vp]
    lib.lars_fused_index_u16.restype = C.c_int
    lib.lars_index_hwc.argtypes = [vp, i32, i64, i32, i32, vp, vp]
    lib.lars_index_hwc.restype = C.c_int
    lib.lars_index_change_u8.argtypes = [vp, vp, i64, i32, i32, f32, f32, vp, vp, vp, vp, vp]
    lib.lars_index_change_u8.restype = C.c_int
    lib.lars_tiff_probe.argtypes = [vp, C.c_size_t, C.POINTER(TiffInfo)]
    lib.lars_tiff_probe.restype = C.c_int
    lib.lars_tiff_read.argtypes = [vp, C.c_size_t, C.POINTER(TiffInfo), vp, C.c_size_t]
    lib.lars_tiff_read.restype = C.c_int
    lib.lars_tiff_read_region.argtypes = [vp, C.c_size_t, C.POINTER(TiffInfo), i32, i32, i32, i32, vp, C.c_size_t, i32]
    lib.lars_tiff_read_region.restype = C.c_int
    lib.lars_png_probe.argtypes = [vp, C.c_size_t, C.POINTER(PngInfo)]
    lib.lars_png_probe.restype = C.c_int
    lib.lars_png_read.argtypes = [vp, C.c_size_t, C.POINTER(PngInfo), vp, C.c_size_t]
    lib.lars_png_read.restype = C.c_int
    lib.lars_tiff_lzw_chunks.argtypes = [vp, C.c_size_t, C.POINTER(TiffInfo), vp, i32]
    lib.lars_tiff_lzw_chunks.restype = C.c_int
    lib.lars_lzw_decode_device.argtypes = [vp, vp, i32, vp, vp, vp]
    lib.lars_lzw_decode_device.restype = C.c_int
    lib.lars_tiff_deflate_chunks.argtypes = [vp, C.c_size_t, C.POINTER(TiffInfo), vp, i32]
    lib.lars_tiff_deflate_chunks.restype = C.c_int
    lib.lars_inflate_decode_device.argtypes = [vp, vp, i32, vp, vp, vp]
    lib.lars_inflate_decode_device.restype = C.c_int
    lib.lars_untile_device.argtypes = [vp, i64, i64, vp, i32, vp, i64, vp]
    lib.lars_untile_device.restype = C.c_int
    lib.lars_tiff_post_device.argtypes = [vp, i32, i64, i32, i32, i32, i32, i32, i32, vp]
    lib.lars_tiff_post_device.restype = C.c_int
    lib.lars_resize_plan_lanczos.argtypes = [i32, i32, i32, i32, i32, C.POINTER(ResizePlan)]
    lib.lars_resize_plan_lanczos.restype = C.c_int
    lib.lars_resize_tables_lanczos.argtypes = [C.POINTER(ResizePlan), vp]
    lib.lars_resize_tables_lanczos.restype = C.c_int
    lib.lars_resize_lanczos_u8.argtypes = [C.POINTER(ResizePlan), vp, vp, i64, i32, vp, i64, vp, C.c_size_t, vp]
    lib.lars_resize_lanczos_u8.restype = C.c_int


def load():
    """Load (building first if the .so is absent and nvcc is available) and return the CDLL."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = _build.LIB_PATH
        if not os.path.isfile(path):
            try:
                _build.build()
            except Exception as exc:  # no CPU fallback -- fail loudly
                raise ImportError(
                    f"liblars_b200.so is missing at {path} and could not be built ({exc}); "
                    "run `python -c 'import __graft_entry__ as g; g.build()'`") from exc
        lib = C.CDLL(path)
        missing = [s for s in EXPORTED_SYMBOLS if not hasattr(lib, s)]
        if missing:
            raise ImportError(f"{path} does not export {missing}; rebuild it")
        _declare(lib)
        _lib = lib
        return lib


def check(status: int, what: str = "") -> int:
    if status < 0:
        msg = load().lars_last_error().decode("utf-8", "replace")
        raise LarsError(f"{what or 'lars call'} failed ({status}): {msg}")
    return status


def colormap_table(name: str) -> np.ndarray:
    """(256, 3) uint8 table of a built-in colormap (host-only, no GPU needed)."""
    out = np.empty((256, 3), np.uint8)
    check(load().lars_colormap_table(CMAP_IDS[name], out.ctypes.data), "lars_colormap_table")
    return out


def histogram_edges(bins: int) -> np.ndarray:
    out = np.empty(bins + 1, np.float32)
    check(load().lars_histogram_edges_f32(bins, out.ctypes.data), "lars_histogram_edges_f32")
    return out


def resize_plan(in_h: int, in_w: int, out_h: int, out_w: int, channels: int):
    """(plan, tables) of a Lanczos resize: geometry + the coefficient block as an int32 array
    (host-only, no GPU needed)."""
    lib = load()
    plan = ResizePlan()
    check(lib.lars_resize_plan_lanczos(in_h, in_w, out_h, out_w, channels, C.byref(plan)), "lars_resize_plan_lanczos")
    tables = np.zeros(int(plan.table_bytes) // 4, np.int32)
    check(lib.lars_resize_tables_lanczos(C.byref(plan), tables.ctypes.data), "lars_resize_tables_lanczos")
    return plan, tables
