"""Host side of the B200 RGNir analysis path.

PyTorch is used for plumbing only -- device / pinned memory, streams, ``torch.distributed``;
every byte of arithmetic happens in the hand-written sm_100a kernels behind the C ABI
(``include/lars_b200.h``).  There is no CPU fallback: without the CUDA library and a B200
the constructor raises.

Two levels:
  * device-resident: :class:`DeviceFrames` in, :class:`DeviceOutputs` out, nothing leaves HBM
    (what ``bench.py`` times as ``value``);
  * host arrays: :meth:`Engine.analyze_frame` / :meth:`Engine.analyze_batch` take NumPy
    frames and return NumPy results, with the H2D / D2H copies inside (what the reference's
    helper signatures need, and what ``bench.py`` times as ``e2e``).
"""
from __future__ import annotations

import ctypes as C
import threading
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import FusedArgs, INDEX_STATS_DTYPE, LarsError, PIXEL_GROUP, STRETCH_U16_BYTES, check

INDEX_TYPES = ("NDVI", "GNDVI", "NDWI")                 # process-images.py:466,472,478
DEFAULT_THRESHOLDS = (0.2, 0.2, 0.0)                    # process-images.py:498-504
DEFAULT_CMAPS = ("RdYlGn", "RdYlGn", "RdYlBu")          # process-images.py:689-692
DEFAULT_BINS = 50                                       # process-ndvi.py:97
DEFAULT_QUANTILES = (0.02, 0.98)                        # process-images.py:437
ALL_OUTPUTS = ("wb", "maps", "rgb", "stats")


def _pad_px(n_pixels: int) -> int:
    return (n_pixels + PIXEL_GROUP - 1) // PIXEL_GROUP * PIXEL_GROUP


@dataclass
class DeviceFrames:
    """A batch of equally-sized interleaved uint8 frames resident in HBM.

    ``data`` is ``[n_frames, padded_pixels * channels]`` uint8: every frame slot starts
    16-byte aligned and is padded to a whole number of 16-pixel groups.
    """
    data: torch.Tensor
    n_pixels: int
    channels: int
    shape: tuple  # (H, W)
    sample_bytes: int = 1   # 1 = uint8 samples, 2 = little-endian uint16 samples

    @property
    def n_frames(self) -> int:
        return self.data.shape[0]

    @property
    def stride_bytes(self) -> int:
        return self.data.stride(0)


@dataclass
class DeviceOutputs:
    """Device-resident results of one pass over a :class:`DeviceFrames` batch."""
    frames: DeviceFrames
    wb_hist: Optional[torch.Tensor] = None      # [F, 3, 256] int64 (bit pattern of uint64)
    wb_lut: Optional[torch.Tensor] = None       # [F, 3, 256] uint8
    wb_pct: Optional[torch.Tensor] = None       # [F, 3, 2] float64 (p2, p98)
    wb: Optional[torch.Tensor] = None           # [F, padded_px * C] uint8
    maps: Optional[torch.Tensor] = None         # [3, F, padded_px] float32
    rgb: Optional[torch.Tensor] = None          # [3, F, padded_px * 3] uint8
    stats: Optional[torch.Tensor] = None        # [F, 3, 576] uint8 (lars_index_stats records)
    bins: int = DEFAULT_BINS
    map_mask: tuple = (True, True, True)
    rgb_mask: tuple = (True, True, True)
    _keep: list = field(default_factory=list)   # workspace kept alive until the stream is done


def stats_records_to_dicts(records: np.ndarray, bins: int) -> List[Dict[str, dict]]:
    """[F, 3] structured ``lars_index_stats`` -> per-frame {index: {...}} dictionaries."""
    out = []
    for f in range(records.shape[0]):
        per = {}
        for i, name in enumerate(INDEX_TYPES):
            r = records[f, i]
            n = int(r["count"])
            per[name] = {
                "count": n,
                "count_above": int(r["count_above"]),
                "sum": float(r["sum"]),
                "sumsq": float(r["sumsq"]),
                "mean": float(r["mean"]),
                "std": float(r["std"]),
                "min": float(r["min"]),
                "max": float(r["max"]),
                "threshold": float(r["threshold"]),
                "coverage_pct": (float(r["count_above"]) / n * 100.0) if n else 0.0,
                "hist": np.array(r["hist"][:bins], dtype=np.int64),
            }
        out.append(per)
    return out


class Engine:
    """One engine per process and GPU.  Thread-safe: each calling thread gets its own stream."""

    def __init__(self, device: Optional[int] = None):
        if not torch.cuda.is_available():
            raise RuntimeError(
                "lars_image_processing_b200 needs an NVIDIA B200 (sm_100a): torch.cuda is not "
                "available and there is no CPU fallback")
        self.lib = _lib.load()
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        with torch.cuda.device(self.device):
            check(self.lib.lars_init(self.device_index), "lars_init")
            self.sm_count = check(self.lib.lars_sm_count(), "lars_sm_count")
        self._tls = threading.local()
        self._resize_plans: Dict[tuple, tuple] = {}
        self._resize_lock = threading.Lock()

    # ------------------------------------------------------------------ plumbing
    def stream(self) -> torch.cuda.Stream:
        s = getattr(self._tls, "stream", None)
        if s is None:
            s = torch.cuda.Stream(device=self.device)
            self._tls.stream = s
        return s

    def _alloc(self, shape, dtype, stream=None):
        # allocate under the stream that will use the block so the caching allocator's
        # reuse-after-free ordering is tied to that stream
        with torch.cuda.stream(stream or self.stream()):
            return torch.empty(shape, dtype=dtype, device=self.device)

    # ------------------------------------------------------------------ device-resident API
    def alloc_frames(self, n_frames: int, height: int, width: int, channels: int = 3,
                     stream=None, sample_bytes: int = 1) -> DeviceFrames:
        n_px = height * width
        data = self._alloc((n_frames, _pad_px(n_px) * channels * sample_bytes), torch.uint8, stream)
        return DeviceFrames(data, n_px, channels, (height, width), sample_bytes)

    # pageable host arrays reach the device through a per-thread pinned staging buffer filled by a few
    # copy threads (NumPy releases the GIL for large copies) while the previous piece is already on the
    # wire; a plain cudaMemcpy from pageable memory is staged by one driver thread at ~12 GB/s
    _STAGE_PIECE = 4 << 20
    _STAGE_MIN = 8 << 20

    def _staged_h2d(self, dst: torch.Tensor, src: np.ndarray, s: torch.cuda.Stream) -> None:
        """dst: 1-D uint8 device view, src: 1-D uint8 host array (pageable), same length."""
        n = src.size
        src_t = torch.from_numpy(src)
        if n < self._STAGE_MIN or src_t.is_pinned():     # small, or already page-locked (e.g. a result of ours)
            dst.copy_(src_t, non_blocking=True)
            return
        st = getattr(self._tls, "stage", None)
        if st is None or st["buf"].numel() < n:
            from concurrent.futures import ThreadPoolExecutor
            if st is not None:
                st["done"].synchronize()
            st = {"buf": torch.empty(n, dtype=torch.uint8, pin_memory=True), "done": torch.cuda.Event(),
                  "pool": st["pool"] if st is not None else ThreadPoolExecutor(4)}
            st["done"].record(s)
            self._tls.stage = st
        st["done"].synchronize()                      # the previous upload has left the staging buffer
        buf_np = st["buf"].numpy()
        piece = self._STAGE_PIECE
        futs = [st["pool"].submit(np.copyto, buf_np[a:min(n, a + piece)], src[a:min(n, a + piece)])
                for a in range(0, n, piece)]
        for k, f in enumerate(futs):
            f.result()
            a, b = k * piece, min(n, (k + 1) * piece)
            dst[a:b].copy_(st["buf"][a:b], non_blocking=True)
        st["done"].record(s)

    def upload(self, frames: Sequence[np.ndarray], stream: Optional[torch.cuda.Stream] = None) -> DeviceFrames:
        """Copy equally-shaped HWC uint8 / uint16 host frames into a padded device batch."""
        first = np.asarray(frames[0])
        h, w, c = first.shape
        s = stream or self.stream()
        sb = first.dtype.itemsize
        dev = self.alloc_frames(len(frames), h, w, c, s, sample_bytes=sb)
        nbytes = h * w * c * sb
        with torch.cuda.stream(s):
            for i, fr in enumerate(frames):
                fr = np.ascontiguousarray(fr)
                if fr.shape != first.shape or fr.dtype != first.dtype:
                    raise ValueError("all frames of a batch must share shape and dtype")
                self._staged_h2d(dev.data[i, :nbytes], fr.reshape(-1).view(np.uint8), s)
        return dev

    def wb_histogram(self, frames: DeviceFrames, shared: bool = False, stream=None) -> torch.Tensor:
        """Pass 1 (K1): [F, 3, 256] per-channel value counts (int64 view of uint64).
        ``shared=True``: the frames are tiles of one image -> a single [1, 3, 256] histogram."""
        s = stream or self.stream()
        hist = self._alloc((1 if shared else frames.n_frames, 3, 256), torch.int64, s)
        with torch.cuda.device(self.device):
            check(self.lib.lars_wb_hist_u8(frames.data.data_ptr(), frames.n_frames, frames.n_pixels,
                                           frames.channels, frames.stride_bytes, hist.data_ptr(),
                                           1 if shared else 0, s.cuda_stream), "lars_wb_hist_u8")
        return hist

    def wb_lut(self, hist: torch.Tensor, quantiles=DEFAULT_QUANTILES, stream=None, chain: str = "images"):
        """K1b: histograms -> (uint8 LUTs [S,3,256], float64 percentiles [S,3,2]).
        ``chain``: "images" = process-images.py:437-441 (float64 stretch, float32 store, truncation);
        "rgn" = process-rgn.py:25-44 (pre-clip, float64 truncated directly)."""
        s = stream or self.stream()
        n_sets = hist.shape[0]
        lut = self._alloc((n_sets, 3, 256), torch.uint8, s)
        pct = self._alloc((n_sets, 3, 2), torch.float64, s)
        with torch.cuda.device(self.device):
            check(self.lib.lars_wb_lut_build_u8_chain(hist.data_ptr(), n_sets, float(quantiles[0]),
                                                      float(quantiles[1]), _lib.WB_CHAINS[chain], lut.data_ptr(),
                                                      pct.data_ptr(), s.cuda_stream), "lars_wb_lut_build_u8_chain")
        return lut, pct

    def alloc_outputs(self, frames: DeviceFrames, outputs=ALL_OUTPUTS, stream=None) -> DeviceOutputs:
        """Pre-allocate the device buffers of a batch (reused across steps via ``out=``)."""
        s = stream or self.stream()
        F, ch, ppx = frames.n_frames, frames.channels, _pad_px(frames.n_pixels)
        res = DeviceOutputs(frames=frames)
        if "wb" in outputs:
            res.wb = self._alloc((F, ppx * ch), torch.uint8, s)
        if "maps" in outputs:
            res.maps = self._alloc((3, F, ppx), torch.float32, s)
        if "rgb" in outputs:
            res.rgb = self._alloc((3, F, ppx * 3), torch.uint8, s)
        if "stats" in outputs:
            res.stats = self._alloc((F, 3, INDEX_STATS_DTYPE.itemsize), torch.uint8, s)
        return res

    def wb_stretch_u16(self, frames: DeviceFrames, quantiles=DEFAULT_QUANTILES, shared: bool = False,
                       stream=None, hist_hook=None):
        """uint16 Pass 1: two-level radix histogram -> exact percentiles -> per-channel stretch
        thresholds.  Returns (stretch [S, 3, 2064] uint8, percentiles [S, 3, 2] float64).
        ``hist_hook(counters)`` (tile-sharded image on several GPUs) is called on the high-byte
        histogram after level A and on the low-byte histograms after level B -- int64 tensors to be
        SUM-all-reduced in place."""
        s = stream or self.stream()
        n_sets = 1 if shared else frames.n_frames
        stretch = self._alloc((n_sets, 3, STRETCH_U16_BYTES), torch.uint8, s)
        pct = self._alloc((n_sets, 3, 2), torch.float64, s)
        ws_bytes = int(self.lib.lars_wb_u16_workspace_bytes(n_sets))
        ws = self._alloc((ws_bytes,), torch.uint8, s)
        def run(stage):
            with torch.cuda.device(self.device):
                check(self.lib.lars_wb_stretch_build_u16_staged(
                    frames.data.data_ptr(), frames.n_frames, frames.n_pixels, frames.channels, frames.stride_bytes,
                    float(quantiles[0]), float(quantiles[1]), stretch.data_ptr(), pct.data_ptr(), ws.data_ptr(),
                    ws_bytes, 1 if shared else 0, stage, s.cuda_stream), "lars_wb_stretch_build_u16")

        if hist_hook is None:
            run(0)
            return stretch, pct
        hi_bytes, lo_bytes = n_sets * 3 * 256 * 8, n_sets * 3 * 4 * 256 * 8
        run(1)                                              # LARS_U16_STAGE_HIST_HI
        with torch.cuda.stream(s):
            hist_hook(ws[:hi_bytes].view(torch.int64).view(n_sets, 3, 256))
        run(2)                                              # LARS_U16_STAGE_HIST_LO
        with torch.cuda.stream(s):
            hist_hook(ws[hi_bytes:hi_bytes + lo_bytes].view(torch.int64).view(n_sets, 3, 4, 256))
        run(3)                                              # LARS_U16_STAGE_BUILD
        return stretch, pct

    def _build_fused_args(self, frames: DeviceFrames, lut, outputs, res: DeviceOutputs, s, indices=INDEX_TYPES,
                          bins: int = DEFAULT_BINS, thresholds=DEFAULT_THRESHOLDS, cmaps=DEFAULT_CMAPS,
                          rgb_indices=None) -> FusedArgs:
        """Fill a ``lars_fused_args`` block (allocating any missing output buffer of ``res``)."""
        F, npx, ch = frames.n_frames, frames.n_pixels, frames.channels
        ppx = _pad_px(npx)
        res.frames = frames
        res.bins = bins
        rgb_indices = indices if rgb_indices is None else rgb_indices
        res.map_mask = tuple(n in indices for n in INDEX_TYPES)
        res.rgb_mask = tuple(n in rgb_indices for n in INDEX_TYPES)
        a = FusedArgs()
        a.struct_bytes = C.sizeof(FusedArgs)
        a.n_frames, a.channels, a.bins, a.n_pixels = F, ch, bins, npx
        a.src, a.src_frame_stride = frames.data.data_ptr(), frames.stride_bytes
        if lut is not None:
            if lut.shape[0] not in (1, F):
                raise ValueError("lut must hold one set per frame or a single shared set")
            a.wb_lut = lut.data_ptr()
            a.lut_frame_stride = lut.stride(0) if lut.shape[0] == F else 0
        if "wb" in outputs:
            if res.wb is None:
                res.wb = self._alloc((F, ppx * ch), torch.uint8, s)
            a.wb_out, a.wb_frame_stride = res.wb.data_ptr(), res.wb.stride(0)
        if "maps" in outputs:
            if res.maps is None:
                res.maps = self._alloc((3, F, ppx), torch.float32, s)
            for i in range(3):
                a.maps[i] = res.maps[i].data_ptr() if res.map_mask[i] else None
            a.map_frame_stride = res.maps.stride(1)
        if "rgb" in outputs:
            if res.rgb is None:
                res.rgb = self._alloc((3, F, ppx * 3), torch.uint8, s)
            for i in range(3):
                a.rgb[i] = res.rgb[i].data_ptr() if res.rgb_mask[i] else None
            a.rgb_frame_stride = res.rgb.stride(1)
        for i in range(3):
            a.cmap[i] = _lib.CMAP_IDS[cmaps[i]]
            a.thresholds[i] = float(thresholds[i])
        if "stats" in outputs:
            if res.stats is None:
                res.stats = self._alloc((F, 3, INDEX_STATS_DTYPE.itemsize), torch.uint8, s)
            ws_bytes = int(self.lib.lars_fused_workspace_bytes(F))
            ws = self._alloc((ws_bytes,), torch.uint8, s)
            res._keep = [ws]
            a.stats, a.workspace, a.workspace_bytes = res.stats.data_ptr(), ws.data_ptr(), ws_bytes
        return a

    def fused(self, frames: DeviceFrames, lut: Optional[torch.Tensor], outputs=ALL_OUTPUTS,
              indices=INDEX_TYPES, bins: int = DEFAULT_BINS, thresholds=DEFAULT_THRESHOLDS,
              cmaps=DEFAULT_CMAPS, rgb_indices=None, out: Optional[DeviceOutputs] = None,
              stream=None) -> DeviceOutputs:
        """Pass 2 (K2): one read of the raw frames -> every requested product.

        ``lut`` None means identity (``calculate_index`` on an already white-balanced frame).
        ``indices`` selects which fp32 maps are written; ``rgb_indices`` (default: same)
        which colormapped images.  Statistics are always produced for all three indices
        when "stats" is requested (they share the arithmetic).
        """
        s = stream or self.stream()
        res = out or DeviceOutputs(frames=frames)
        a = self._build_fused_args(frames, lut, outputs, res, s, indices=indices, bins=bins, thresholds=thresholds,
                                   cmaps=cmaps, rgb_indices=rgb_indices)
        with torch.cuda.device(self.device):
            if frames.sample_bytes == 2:
                check(self.lib.lars_fused_index_u16(C.byref(a), s.cuda_stream), "lars_fused_index_u16")
            else:
                check(self.lib.lars_fused_index_u8(C.byref(a), s.cuda_stream), "lars_fused_index_u8")
        return res

    def process_device(self, frames: DeviceFrames, outputs=ALL_OUTPUTS, white_balance: bool = True,
                       quantiles=DEFAULT_QUANTILES, out: Optional[DeviceOutputs] = None,
                       stream=None, tiles_of_one_image: bool = False, hist_hook=None, wb_chain: str = "images",
                       **kw) -> DeviceOutputs:
        """Pass 1 + LUT + Pass 2 on a device-resident batch; nothing is synchronised.

        ``tiles_of_one_image``: the frames are tiles / row bands of ONE image (orthomosaic): the
        white-balance percentiles are global, so a single histogram is accumulated over all
        tiles and one LUT serves every tile.  ``hist_hook(hist)`` runs between Pass 1 and the
        LUT build (the multi-GPU path SUM-all-reduces the histogram there)."""
        s = stream or self.stream()
        res = out or DeviceOutputs(frames=frames)
        lut = None
        if frames.sample_bytes == 2:
            if not white_balance:
                raise LarsError("uint16 frames go through the white-balance stretch (uint8 out); "
                                "there is no identity mode for them")
            if wb_chain != "images":
                raise LarsError("the process-rgn.py chain exists for uint8 frames only (its Image.open route "
                                "cannot deliver 16-bit RGB)")
            res.wb_lut, res.wb_pct = self.wb_stretch_u16(frames, quantiles, tiles_of_one_image, s, hist_hook=hist_hook)
            return self.fused(frames, res.wb_lut, outputs=outputs, out=res, stream=s, **kw)
        if white_balance:
            res.wb_hist = self.wb_histogram(frames, shared=tiles_of_one_image, stream=s)
            if hist_hook is not None:
                with torch.cuda.stream(s):
                    hist_hook(res.wb_hist)
            res.wb_lut, res.wb_pct = self.wb_lut(res.wb_hist, quantiles, stream=s, chain=wb_chain)
            lut = res.wb_lut
        return self.fused(frames, lut, outputs=outputs, out=res, stream=s, **kw)

    # ------------------------------------------------------------------ resize in front of the path
    @staticmethod
    def preprocess_target(height: int, width: int, max_dimension: int = 1024):
        """process-images.py:404-416: (new_h, new_w), or None when the frame is small enough."""
        if max(height, width) <= max_dimension:
            return None
        if height > width:
            return max_dimension, int(width * (max_dimension / height))
        return int(height * (max_dimension / width)), max_dimension

    def _resize_plan(self, in_h, in_w, out_h, out_w, channels):
        key = (in_h, in_w, out_h, out_w, channels)
        with self._resize_lock:
            hit = self._resize_plans.get(key)
            if hit is None:
                plan, tables = _lib.resize_plan(in_h, in_w, out_h, out_w, channels)
                dev = torch.from_numpy(tables).to(self.device)      # synchronous: once per geometry
                hit = (plan, dev)
                if len(self._resize_plans) >= 64:                   # bounded: drop the oldest geometry
                    self._resize_plans.pop(next(iter(self._resize_plans)))
                self._resize_plans[key] = hit
        return hit

    def resize_device(self, frames: DeviceFrames, out_h: int, out_w: int, stream=None,
                      premultiplied_alpha: bool = True) -> DeviceFrames:
        """K10: Pillow-exact Lanczos resize of a device-resident uint8 batch (the GPU form of
        preprocess_large_image, process-images.py:398-422) -> a new padded device batch.
        4-channel frames are RGBA to Pillow, which resizes them with premultiplied alpha; with
        ``premultiplied_alpha`` (default) so does this call -- the input batch is premultiplied IN PLACE."""
        if frames.sample_bytes != 1:
            raise LarsError("resize covers uint8 frames (Pillow cannot hold multi-channel 16-bit images)")
        if out_h < 1 or out_w < 1:
            raise ValueError("height and width must be > 0")         # Pillow's message
        s = stream or self.stream()
        in_h, in_w = frames.shape
        plan, tables = self._resize_plan(in_h, in_w, out_h, out_w, frames.channels)
        tables.record_stream(s)                  # the plan cache may drop the block while this launch still reads it
        out = self.alloc_frames(frames.n_frames, out_h, out_w, frames.channels, s)
        temp_bytes = int(plan.temp_frame_bytes) * frames.n_frames
        temp = self._alloc((max(temp_bytes, 16),), torch.uint8, s)
        rgba = premultiplied_alpha and frames.channels == 4 and (plan.need_h or plan.need_v)
        with torch.cuda.device(self.device):
            if rgba:
                check(self.lib.lars_rgba_alpha_u8(frames.data.data_ptr(), frames.n_frames, frames.n_pixels,
                                                  frames.stride_bytes, 1, s.cuda_stream), "lars_rgba_alpha_u8")
            check(self.lib.lars_resize_lanczos_u8(C.byref(plan), tables.data_ptr(), frames.data.data_ptr(),
                                                  frames.stride_bytes, frames.n_frames, out.data.data_ptr(),
                                                  out.stride_bytes, temp.data_ptr(), temp_bytes, s.cuda_stream),
                  "lars_resize_lanczos_u8")
            if rgba:
                check(self.lib.lars_rgba_alpha_u8(out.data.data_ptr(), out.n_frames, out.n_pixels, out.stride_bytes, 0,
                                                  s.cuda_stream), "lars_rgba_alpha_u8")
        return out

    def resize_batch(self, frames: Sequence[np.ndarray], out_h: int, out_w: int) -> List[np.ndarray]:
        """Host uint8 frames (HxWxC or HxW) in, resized host frames out."""
        arrs = [np.asarray(f) for f in frames]
        two_d = arrs[0].ndim == 2
        if any(a.dtype != np.uint8 for a in arrs):
            raise TypeError("resize needs uint8 frames")
        s = self.stream()
        dev = self.upload([a[:, :, None] if two_d else a for a in arrs], stream=s)
        out = self.resize_device(dev, out_h, out_w, s)
        n = out_h * out_w * dev.channels
        with torch.cuda.stream(s):
            host = torch.empty((dev.n_frames, n), dtype=torch.uint8, pin_memory=True)
            host.copy_(out.data[:, :n], non_blocking=True)
        s.synchronize()
        shape = (out_h, out_w) if two_d else (out_h, out_w, dev.channels)
        return [host[i].numpy().reshape(shape).copy() for i in range(dev.n_frames)]

    # ------------------------------------------------------------------ host-array API
    def download(self, res: DeviceOutputs, stream=None, pinned: bool = True) -> List[dict]:
        """Copy the requested products of every frame back to host memory -> list of dicts."""
        s = stream or self.stream()
        fr = res.frames
        F, npx, ch = fr.n_frames, fr.n_pixels, fr.channels
        h, w = fr.shape

        def host(shape, dtype):
            return torch.empty(shape, dtype=dtype, pin_memory=pinned)

        with torch.cuda.stream(s):
            h_wb = h_maps = h_rgb = h_stats = h_pct = None
            if res.wb is not None:
                h_wb = host((F, npx * ch), torch.uint8)
                h_wb.copy_(res.wb[:, :npx * ch], non_blocking=True)
            if res.maps is not None:
                sel = [i for i in range(3) if res.map_mask[i]]
                h_maps = host((len(sel), F, npx), torch.float32)
                for k, i in enumerate(sel):
                    h_maps[k].copy_(res.maps[i, :, :npx], non_blocking=True)
            if res.rgb is not None:
                selr = [i for i in range(3) if res.rgb_mask[i]]
                h_rgb = host((len(selr), F, npx * 3), torch.uint8)
                for k, i in enumerate(selr):
                    h_rgb[k].copy_(res.rgb[i, :, :npx * 3], non_blocking=True)
            if res.stats is not None:
                h_stats = host(tuple(res.stats.shape), torch.uint8)
                h_stats.copy_(res.stats, non_blocking=True)
            if res.wb_pct is not None:
                h_pct = host(tuple(res.wb_pct.shape), torch.float64)
                h_pct.copy_(res.wb_pct, non_blocking=True)
        s.synchronize()

        stats = None
        if h_stats is not None:
            rec = h_stats.numpy().view(INDEX_STATS_DTYPE).reshape(F, 3)
            stats = stats_records_to_dicts(rec, res.bins)
        results = []
        for f in range(F):
            d: dict = {}
            if h_wb is not None:
                d["wb"] = h_wb[f].numpy().reshape(h, w, ch)
            if h_maps is not None:
                d["maps"] = {INDEX_TYPES[i]: h_maps[k, f].numpy().reshape(h, w) for k, i in enumerate(sel)}
            if h_rgb is not None:
                d["rgb"] = {INDEX_TYPES[i]: h_rgb[k, f].numpy().reshape(h, w, 3) for k, i in enumerate(selr)}
            if stats is not None:
                d["stats"] = stats[f]
            if h_pct is not None:
                d["percentiles"] = h_pct[f if h_pct.shape[0] == F else 0].numpy().copy()
            results.append(d)
        return results

    # ------------------------------------------------------------------ pinned, pipelined host API
    def alloc_host_outputs(self, n_frames: int, height: int, width: int, channels: int = 3,
                           outputs=ALL_OUTPUTS) -> Dict[str, torch.Tensor]:
        """Pinned host buffers for :meth:`run_host_batch` (allocate once, reuse every step)."""
        npx = height * width
        out: Dict[str, torch.Tensor] = {}
        if "wb" in outputs:
            out["wb"] = torch.empty((n_frames, npx * channels), dtype=torch.uint8, pin_memory=True)
        if "maps" in outputs:
            out["maps"] = torch.empty((3, n_frames, npx), dtype=torch.float32, pin_memory=True)
        if "rgb" in outputs:
            out["rgb"] = torch.empty((3, n_frames, npx * 3), dtype=torch.uint8, pin_memory=True)
        if "stats" in outputs:
            out["stats"] = torch.empty((n_frames, 3, INDEX_STATS_DTYPE.itemsize), dtype=torch.uint8,
                                       pin_memory=True)
        return out

    def run_host_batch(self, host_frames: torch.Tensor, shape, host_out: Dict[str, torch.Tensor],
                       chunk: int = 4, white_balance: bool = True, sample_bytes: int = 1, **kw) -> None:
        """Host frames in (pinned ``[F, H*W*C]`` uint8, or ``[F, H*W*C*2]`` bytes of little-endian
        uint16 samples with ``sample_bytes=2``), host results out, software-pipelined.

        The batch is cut into chunks of ``chunk`` frames; H2D of chunk c+1, the kernels of chunk
        c and D2H of chunk c-1 run concurrently on three streams over double-buffered device
        slots, so the PCIe copy engines stay busy in both directions while the SMs work.
        Returns after everything has landed in ``host_out``.
        """
        h, w, ch = shape
        F = host_frames.shape[0]
        npx = h * w
        nbytes = npx * ch * sample_bytes
        if host_frames.dtype != torch.uint8 or host_frames.shape[1] != nbytes:
            raise ValueError(f"host_frames must be uint8 [F, {nbytes}] for this shape / sample width")
        outputs = tuple(k for k in ALL_OUTPUTS if k in host_out)
        st = getattr(self._tls, "pipe", None)
        key = (chunk, h, w, ch, outputs, sample_bytes)
        if st is None or st["key"] != key:
            s_in, s_cmp, s_out = (torch.cuda.Stream(device=self.device) for _ in range(3))
            slots = []
            for _ in range(2):
                frames = self.alloc_frames(chunk, h, w, ch, s_cmp, sample_bytes=sample_bytes)
                slots.append({"frames": frames, "res": self.alloc_outputs(frames, outputs, s_cmp)})
            st = {"key": key, "streams": (s_in, s_cmp, s_out), "slots": slots}
            self._tls.pipe = st
        s_in, s_cmp, s_out = st["streams"]
        slots = st["slots"]
        n_chunks = (F + chunk - 1) // chunk
        ev_in = [torch.cuda.Event() for _ in range(n_chunks)]
        ev_cmp = [torch.cuda.Event() for _ in range(n_chunks)]
        ev_out = [torch.cuda.Event() for _ in range(n_chunks)]
        for c in range(n_chunks):
            a, b = c * chunk, min(F, (c + 1) * chunk)
            k = b - a
            slot = slots[c & 1]
            fr: DeviceFrames = slot["frames"]
            view = DeviceFrames(fr.data[:k], fr.n_pixels, fr.channels, fr.shape, fr.sample_bytes)
            with torch.cuda.stream(s_in):
                if c >= 2:
                    s_in.wait_event(ev_cmp[c - 2])          # slot's input consumed
                view.data[:, :nbytes].copy_(host_frames[a:b], non_blocking=True)
                ev_in[c].record(s_in)
            with torch.cuda.stream(s_cmp):
                s_cmp.wait_event(ev_in[c])
                if c >= 2:
                    s_cmp.wait_event(ev_out[c - 2])         # slot's outputs drained
                res = slot["res"]
                sub = DeviceOutputs(frames=view,
                                    wb=None if res.wb is None else res.wb[:k],
                                    maps=None if res.maps is None else res.maps[:, :k],
                                    rgb=None if res.rgb is None else res.rgb[:, :k],
                                    stats=None if res.stats is None else res.stats[:k])
                self.process_device(view, outputs=outputs, white_balance=white_balance, out=sub,
                                    stream=s_cmp, **kw)
                ev_cmp[c].record(s_cmp)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_cmp[c])
                if "wb" in host_out:
                    host_out["wb"][a:b].copy_(sub.wb[:k, :npx * ch], non_blocking=True)
                if "maps" in host_out:
                    for i in range(3):
                        if sub.map_mask[i]:
                            host_out["maps"][i, a:b].copy_(sub.maps[i, :k, :npx], non_blocking=True)
                if "rgb" in host_out:
                    for i in range(3):
                        if sub.rgb_mask[i]:
                            host_out["rgb"][i, a:b].copy_(sub.rgb[i, :k, :npx * 3], non_blocking=True)
                if "stats" in host_out:
                    host_out["stats"][a:b].copy_(sub.stats[:k], non_blocking=True)
                ev_out[c].record(s_out)
        s_out.synchronize()

    def run_host_mosaic(self, host_tiles: torch.Tensor, shape, host_out: Dict[str, torch.Tensor], chunk: int = 4,
                        hist_hook=None, quantiles=DEFAULT_QUANTILES, **kw) -> torch.Tensor:
        """The tiles of ONE image this rank owns (BASELINE config 4: a tile-sharded orthomosaic) from pinned host
        memory to host results.  The white-balance percentiles are global to the image
        (process-images.py:435-438), so the pass has a barrier in the middle: every tile is uploaded, ONE
        histogram is accumulated over all of them, ``hist_hook(hist)`` makes it image-wide (SUM all-reduce of
        the [1, 3, 256] counters over the ranks), one LUT is built, and only then Pass 2 runs chunk by chunk with
        the D2H of chunk c-1 overlapping the kernels of chunk c.  uint8 tiles.  Returns the per-tile statistics
        records on the device ([T, 3, 576]; merged image-wide by ``distributed.dataset_statistics``)."""
        h, w, ch = shape
        T = host_tiles.shape[0]
        npx = h * w
        nbytes = npx * ch
        if host_tiles.dtype != torch.uint8 or host_tiles.shape[1] != nbytes:
            raise ValueError(f"host_tiles must be uint8 [T, {nbytes}] for this tile shape")
        outputs = tuple(k for k in ALL_OUTPUTS if k in host_out)
        st = getattr(self._tls, "mosaic", None)
        key = (T, chunk, h, w, ch, outputs)
        if st is None or st["key"] != key:
            s_cmp, s_out = torch.cuda.Stream(device=self.device), torch.cuda.Stream(device=self.device)
            tiles = self.alloc_frames(T, h, w, ch, s_cmp)
            slots = []
            for _ in range(2):
                view = DeviceFrames(tiles.data[:chunk], npx, ch, (h, w), 1)
                slots.append(self.alloc_outputs(view, tuple(k for k in outputs if k != "stats"), s_cmp))
            stats = self._alloc((T, 3, INDEX_STATS_DTYPE.itemsize), torch.uint8, s_cmp)
            st = {"key": key, "streams": (s_cmp, s_out), "tiles": tiles, "slots": slots, "stats": stats}
            self._tls.mosaic = st
        s_cmp, s_out = st["streams"]
        tiles, slots, stats = st["tiles"], st["slots"], st["stats"]
        with torch.cuda.stream(s_cmp):
            tiles.data[:, :nbytes].copy_(host_tiles, non_blocking=True)
            hist = self.wb_histogram(tiles, shared=True, stream=s_cmp)
            if hist_hook is not None:
                hist_hook(hist)
            lut, _ = self.wb_lut(hist, quantiles, stream=s_cmp)
        n_chunks = (T + chunk - 1) // chunk
        ev_cmp = [torch.cuda.Event() for _ in range(n_chunks)]
        ev_out = [torch.cuda.Event() for _ in range(n_chunks)]
        for c in range(n_chunks):
            a, b = c * chunk, min(T, (c + 1) * chunk)
            k = b - a
            res = slots[c & 1]
            view = DeviceFrames(tiles.data[a:b], npx, ch, (h, w), 1)
            with torch.cuda.stream(s_cmp):
                if c >= 2:
                    s_cmp.wait_event(ev_out[c - 2])         # slot's outputs drained
                sub = DeviceOutputs(frames=view,
                                    wb=None if res.wb is None else res.wb[:k],
                                    maps=None if res.maps is None else res.maps[:, :k],
                                    rgb=None if res.rgb is None else res.rgb[:, :k],
                                    stats=stats[a:b] if "stats" in outputs else None)
                self.fused(view, lut, outputs=outputs, out=sub, stream=s_cmp, **kw)
                ev_cmp[c].record(s_cmp)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_cmp[c])
                if "wb" in host_out:
                    host_out["wb"][a:b].copy_(sub.wb[:k, :npx * ch], non_blocking=True)
                if "maps" in host_out:
                    for i in range(3):
                        if sub.map_mask[i]:
                            host_out["maps"][i, a:b].copy_(sub.maps[i, :k, :npx], non_blocking=True)
                if "rgb" in host_out:
                    for i in range(3):
                        if sub.rgb_mask[i]:
                            host_out["rgb"][i, a:b].copy_(sub.rgb[i, :k, :npx * 3], non_blocking=True)
                if "stats" in host_out:
                    host_out["stats"][a:b].copy_(stats[a:b], non_blocking=True)
                ev_out[c].record(s_out)
        s_out.synchronize()
        return stats

    @staticmethod
    def _check_frame(img: np.ndarray) -> np.ndarray:
        img = np.asarray(img)
        if img.ndim != 3:
            # the reference indexes img[:, :, i] and raises IndexError on 2-D input
            raise IndexError("too many indices for array: expected an HxWxC frame, "
                             f"got {img.ndim}-dimensional input")
        if img.shape[2] < 3:
            raise IndexError(f"index 2 is out of bounds for axis 2 with size {img.shape[2]}")
        if img.shape[2] > 4:
            raise LarsError(f"frames with {img.shape[2]} channels are not supported (3 or 4)")
        if img.dtype not in (np.uint8, np.uint16):
            raise LarsError(f"dtype {img.dtype} is not supported (uint8 or uint16 frames)")
        return img

    def analyze_batch(self, frames: Sequence[np.ndarray], outputs=ALL_OUTPUTS, white_balance=True,
                      max_dimension: Optional[int] = None, **kw) -> List[dict]:
        """Host frames in, host results out (H2D, [resize,] Pass 1, LUT, Pass 2, D2H, one sync).
        ``max_dimension``: apply preprocess_large_image (process-images.py:398-422) on the device
        first, as the app does before every analysis (:1130, :1444)."""
        frames = [self._check_frame(f) for f in frames]
        s = self.stream()
        dev = self.upload(frames, stream=s)
        if max_dimension is not None:
            target = self.preprocess_target(dev.shape[0], dev.shape[1], max_dimension)
            if target is not None:
                dev = self.resize_device(dev, target[0], target[1], s)
        res = self.process_device(dev, outputs=outputs, white_balance=white_balance, stream=s, **kw)
        return self.download(res, stream=s)

    def analyze_frame(self, img: np.ndarray, outputs=ALL_OUTPUTS, white_balance=True, **kw) -> dict:
        """The fused entry point for one frame: {wb, maps, rgb, stats, percentiles}."""
        return self.analyze_batch([img], outputs=outputs, white_balance=white_balance, **kw)[0]


class FramePlan:
    """Pre-bound execution plan for a fixed batch shape (SURVEY.md section 7 hard part 5: the
    many-small-frames regime is launch-bound, so nothing may be allocated or re-derived per step).

    All device buffers, the workspace and the C argument block are created once; ``run`` is then
    four C-ABI calls (Pass 1, LUT build, fused Pass 2 + finalize, optional dataset merge) and
    ``capture`` records them into a CUDA graph so that a step is a single ``replay``.
    """

    def __init__(self, engine: "Engine", frames: DeviceFrames, outputs=ALL_OUTPUTS, white_balance: bool = True,
                 quantiles=DEFAULT_QUANTILES, merge_dataset: bool = False, stream=None,
                 tiles_of_one_image: bool = False, hist_hook=None, peer_exchange=None, **fused_kw):
        """``tiles_of_one_image``: the frames are tiles of ONE image (orthomosaic, BASELINE config 4): Pass 1
        accumulates a single histogram over all of them, ``hist_hook(hist)`` runs on the plan's stream between
        Pass 1 and the LUT build (the multi-GPU path SUM-all-reduces the [1, 3, 256] counters there) and one
        LUT serves every tile (process-images.py:435-438: the percentiles are global to the image).
        ``peer_exchange`` (a ``distributed.PeerHistogramExchange``) replaces hook + LUT build by the one kernel that
        exchanges the counters over NVLink peer memory itself."""
        self.engine = engine
        self.frames = frames
        self.stream = stream or engine.stream()
        s = self.stream
        F = frames.n_frames
        self.white_balance = white_balance
        self.quantiles = (float(quantiles[0]), float(quantiles[1]))
        self.out = engine.alloc_outputs(frames, outputs, s)
        self.u16 = frames.sample_bytes == 2
        self.shared = bool(tiles_of_one_image)
        self.hist_hook = hist_hook
        self.peer_exchange = peer_exchange
        if peer_exchange is not None and not self.shared:
            raise LarsError("peer_exchange belongs to a tiles_of_one_image plan")
        if self.shared and (self.u16 or not white_balance):
            raise LarsError("FramePlan(tiles_of_one_image=True) covers white-balanced uint8 tiles; uint16 mosaics go "
                            "through Engine.process_device (staged two-level histogram)")
        n_sets = 1 if self.shared else F
        if self.u16:
            if not white_balance:
                raise LarsError("uint16 frames go through the white-balance stretch; there is no identity mode")
            self.hist = None
            self.lut = engine._alloc((F, 3, STRETCH_U16_BYTES), torch.uint8, s)       # lars_stretch_u16 records
            self.pct = engine._alloc((F, 3, 2), torch.float64, s)
            self._u16_ws_bytes = int(engine.lib.lars_wb_u16_workspace_bytes(F))
            self._u16_ws = engine._alloc((self._u16_ws_bytes,), torch.uint8, s)
        else:
            self.hist = engine._alloc((n_sets, 3, 256), torch.int64, s) if white_balance else None
            self.lut = engine._alloc((n_sets, 3, 256), torch.uint8, s) if white_balance else None
            self.pct = engine._alloc((n_sets, 3, 2), torch.float64, s) if white_balance else None
        self.out.wb_hist, self.out.wb_lut, self.out.wb_pct = self.hist, self.lut, self.pct
        self.merged = engine._alloc((3, INDEX_STATS_DTYPE.itemsize), torch.uint8, s) \
            if (merge_dataset and "stats" in outputs) else None
        # one dry run builds the argument block (and the workspace) through the normal path
        self._args = engine._build_fused_args(frames, self.lut, outputs, self.out, s, **fused_kw)
        self._stats_ptr = self._args.stats
        self.graph: Optional[torch.cuda.CUDAGraph] = None

    def rebind(self, frames: DeviceFrames, stats: Optional[torch.Tensor] = None) -> "FramePlan":
        """Point the plan at another input batch of the SAME geometry (a group of a resident survey or of a
        device ring) and, optionally, at another ``[F, 3, 576]`` destination for the statistics records; every
        other buffer (histograms, LUTs, maps, images, workspace) is reused.  Not for captured plans."""
        old = self.frames
        if (frames.n_frames, frames.n_pixels, frames.channels, frames.sample_bytes, frames.stride_bytes) != \
                (old.n_frames, old.n_pixels, old.channels, old.sample_bytes, old.stride_bytes):
            raise ValueError("rebind needs a batch of the same geometry")
        if self.graph is not None:
            raise LarsError("a captured plan cannot be rebound")
        self.frames = frames
        self.out.frames = frames
        self._args.src = frames.data.data_ptr()
        if stats is not None:
            if self._stats_ptr is None or tuple(stats.shape) != (frames.n_frames, 3, INDEX_STATS_DTYPE.itemsize):
                raise ValueError("stats must be [n_frames, 3, 576] uint8 and the plan must produce statistics")
            self._args.stats = stats.data_ptr()
            self.out.stats = stats
        return self

    def run(self, fused_events=None) -> DeviceOutputs:
        """Enqueue one step on the plan's stream (no allocation, no synchronisation).
        ``fused_events``: optional (start, end) CUDA events recorded around the fused Pass 2."""
        lib, fr, sp = self.engine.lib, self.frames, self.stream.cuda_stream
        if self.u16:
            check(lib.lars_wb_stretch_build_u16(fr.data.data_ptr(), fr.n_frames, fr.n_pixels, fr.channels, fr.stride_bytes,
                                                self.quantiles[0], self.quantiles[1], self.lut.data_ptr(),
                                                self.pct.data_ptr(), self._u16_ws.data_ptr(), self._u16_ws_bytes, 0, sp),
                  "lars_wb_stretch_build_u16")
        elif self.white_balance:
            check(lib.lars_wb_hist_u8(fr.data.data_ptr(), fr.n_frames, fr.n_pixels, fr.channels, fr.stride_bytes,
                                      self.hist.data_ptr(), 1 if self.shared else 0, sp), "lars_wb_hist_u8")
            if self.peer_exchange is not None:
                self.peer_exchange.lut_build(self.hist, self.lut, self.pct, self.quantiles, stream=self.stream)
            else:
                if self.hist_hook is not None:
                    with torch.cuda.stream(self.stream):
                        self.hist_hook(self.hist)
                check(lib.lars_wb_lut_build_u8(self.hist.data_ptr(), self.hist.shape[0], self.quantiles[0], self.quantiles[1],
                                               self.lut.data_ptr(), self.pct.data_ptr(), sp), "lars_wb_lut_build_u8")
        if fused_events is not None:
            fused_events[0].record(self.stream)
        if self.u16:
            check(lib.lars_fused_index_u16(C.byref(self._args), sp), "lars_fused_index_u16")
        else:
            check(lib.lars_fused_index_u8(C.byref(self._args), sp), "lars_fused_index_u8")
        if fused_events is not None:
            fused_events[1].record(self.stream)
        if self.merged is not None:
            check(lib.lars_stats_merge(self.out.stats.data_ptr(), fr.n_frames, self.merged.data_ptr(), sp),
                  "lars_stats_merge")
        return self.out

    def capture(self) -> "FramePlan":
        """Record ``run`` into a CUDA graph; afterwards ``replay`` costs one launch."""
        with torch.cuda.device(self.engine.device):
            self.run()                                  # warm-up outside capture (module load, attributes)
            self.stream.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self.stream):
                self.run()
            self.graph = g
        return self

    def replay(self) -> DeviceOutputs:
        if self.graph is None:
            raise LarsError("FramePlan.capture() has not been called")
        with torch.cuda.stream(self.stream):
            self.graph.replay()
        return self.out


_default_engine: Optional[Engine] = None
_default_lock = threading.Lock()


def get_engine() -> Engine:
    """Process-wide engine on the current CUDA device (created on first use)."""
    global _default_engine
    with _default_lock:
        if _default_engine is None:
            _default_engine = Engine()
        return _default_engine
