"""Ingest in front of the path and the streaming many-frame runner (SURVEY.md section 8(f) rank 4,
BASELINE configs 3 and 5).

The reference loads every frame with ``np.array(PIL.Image.open(...))`` (process-images.py:183-193;
backend-process.py:52; process-ndvi.py:18; process-rgn.py:18) inside a serial per-file loop
(backend-process.py:92-97; process-images.py:633-663).  Here:

* :func:`read_frame` keeps that behaviour for every format Pillow decodes, and adds a native
  baseline-TIFF reader (``lars_tiff_probe`` / ``lars_tiff_read``, host side of the C ABI) that
  copies the strips of a memory-mapped file straight into a pinned buffer -- including 16-bit
  RGB TIFFs, which Pillow opens as 8-bit;
* :class:`SurveyPipeline` streams any number of equally-shaped frames through the GPU path:
  host threads decode into a ring of pinned chunk buffers, one stream copies H2D, one runs
  Pass 1 + LUT + Pass 2, one copies results back, and the dataset-wide statistics are folded on
  the device chunk by chunk (``lars_stats_merge``), with the usual single all-gather across ranks
  at the end.  Nothing in here computes on the CPU: without the CUDA library it raises.
"""
from __future__ import annotations

import ctypes as C
import io
import mmap
import os
import queue
import struct
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib
from ._lib import INDEX_STATS_DTYPE, LarsError, check
from .engine import ALL_OUTPUTS, DeviceFrames, DeviceOutputs, Engine, get_engine, stats_records_to_dicts

Source = Union[str, os.PathLike, bytes, bytearray, memoryview, np.ndarray]

LARS_ERR_UNSUPPORTED = -3


# ------------------------------------------------------------------------------------------
# files -> arrays
# ------------------------------------------------------------------------------------------
def _tiff_probe(buf) -> Optional[_lib.TiffInfo]:
    """TiffInfo if the native reader handles this buffer, None if Pillow has to; raises on a
    corrupt TIFF."""
    lib = _lib.load()
    info = _lib.TiffInfo()
    view = np.frombuffer(buf, dtype=np.uint8)
    try:        # no view of a memory-mapped file may outlive this call (closing the map would fail)
        if view.size < 4 or bytes(view[:2]) not in (b"II", b"MM"):
            return None
        rc = lib.lars_tiff_probe(view.ctypes.data, view.size, C.byref(info))
    finally:
        del view
    if rc == LARS_ERR_UNSUPPORTED:
        return None
    check(rc, "lars_tiff_probe")
    return info


def _tiff_shape(info) -> tuple:
    spp = info.samples_per_pixel
    return (info.height, info.width) if spp == 1 else (info.height, info.width, spp)


def read_frame(source: Source, out: Optional[np.ndarray] = None) -> np.ndarray:
    """One frame as the array ``np.array(Image.open(source))`` would give -- except that 16-bit
    TIFFs keep their 16 bits.  ``source``: path, encoded bytes, or an array (returned as is).
    ``out``: optional destination (e.g. a row of a pinned buffer) of the right size and dtype."""
    if isinstance(source, np.ndarray):
        if out is not None:
            np.copyto(out.reshape(source.shape), source)
            return out.reshape(source.shape)
        return source
    if isinstance(source, (bytes, bytearray, memoryview)):
        return _decode_buffer(source, out)
    with open(os.fspath(source), "rb") as fh:
        size = os.fstat(fh.fileno()).st_size
        if size == 0:
            raise ValueError(f"{source}: empty file")
        with mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ) as mm:
            return _decode_buffer(mm, out)


def _decode_buffer(buf, out: Optional[np.ndarray]) -> np.ndarray:
    info = _tiff_probe(buf)
    if info is not None:
        dtype = np.uint8 if info.bits_per_sample == 8 else np.uint16
        shape = _tiff_shape(info)
        dst = out if out is not None else np.empty(shape, dtype)
        if dst.dtype != dtype or dst.size != int(np.prod(shape)) or not dst.flags.c_contiguous:
            raise ValueError(f"destination must be a contiguous {np.dtype(dtype).name} array of {shape}")
        view = np.frombuffer(buf, dtype=np.uint8)
        try:
            rc = _lib.load().lars_tiff_read(view.ctypes.data, view.size, C.byref(info), dst.ctypes.data, dst.nbytes)
        finally:
            del view
        check(rc, "lars_tiff_read")
        return dst.reshape(shape)
    from PIL import Image
    data = buf if isinstance(buf, (bytes, bytearray)) else bytes(buf)
    arr = np.array(Image.open(io.BytesIO(data)))              # process-images.py:183-193
    if out is not None:
        if out.dtype != arr.dtype or out.size != arr.size:
            raise ValueError(f"destination must hold {arr.shape} {arr.dtype}, got {out.shape} {out.dtype}")
        np.copyto(out.reshape(arr.shape), arr)
        return out.reshape(arr.shape)
    return arr


def frame_info(source: Source) -> tuple:
    """(shape, dtype) of a frame without decoding the pixels where the format allows it."""
    if isinstance(source, np.ndarray):
        return tuple(source.shape), source.dtype
    if not isinstance(source, (bytes, bytearray, memoryview)):
        with open(os.fspath(source), "rb") as fh, mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ) as mm:
            return frame_info(memoryview(mm)) if _is_tiff(mm) else _info_via_pillow(bytes(mm))
    info = _tiff_probe(source)
    if info is not None:
        return _tiff_shape(info), np.dtype(np.uint8 if info.bits_per_sample == 8 else np.uint16)
    return _info_via_pillow(source)


def _is_tiff(buf) -> bool:
    return len(buf) >= 4 and bytes(buf[:2]) in (b"II", b"MM")


def _info_via_pillow(buf) -> tuple:
    arr = _decode_buffer(buf, None)
    return tuple(arr.shape), arr.dtype


def write_tiff(path, array: np.ndarray, big_endian: bool = False, rows_per_strip: Optional[int] = None) -> None:
    """Baseline TIFF writer for HxW / HxWx3 / HxWx4 uint8 or uint16 frames (uncompressed, chunky).
    Pillow cannot write 16-bit RGB; survey frames of BASELINE config 3 are stored with this."""
    a = np.ascontiguousarray(array)
    if a.dtype not in (np.uint8, np.uint16) or a.ndim not in (2, 3):
        raise ValueError("write_tiff needs a uint8 / uint16 HxW or HxWxC array")
    h, w = a.shape[:2]
    spp = 1 if a.ndim == 2 else a.shape[2]
    bits = a.dtype.itemsize * 8
    e = ">" if big_endian else "<"
    data = a.astype(a.dtype.newbyteorder(e)).tobytes()
    rps = h if not rows_per_strip else max(1, min(h, int(rows_per_strip)))
    n_strips = (h + rps - 1) // rps
    row_bytes = w * spp * a.dtype.itemsize
    entries = []
    extra = b""
    header_len = 8
    n_tags = 10 + (1 if spp == 4 else 0)
    ifd_len = 2 + 12 * n_tags + 4
    extra_base = header_len + ifd_len

    def put_extra(blob: bytes) -> int:
        nonlocal extra
        pos = extra_base + len(extra)
        extra += blob + (b"\0" if len(blob) & 1 else b"")
        return pos

    def entry(tag, typ, values):
        fmt = {3: "H", 4: "I"}[typ]
        blob = struct.pack(e + fmt * len(values), *values)
        if len(blob) <= 4:
            entries.append(struct.pack(e + "HHI", tag, typ, len(values)) + blob.ljust(4, b"\0"))
        else:
            entries.append(struct.pack(e + "HHII", tag, typ, len(values), put_extra(blob)))

    # strip offsets depend on the size of the extra block: lay out twice
    offsets = [0] * n_strips
    for _ in range(2):
        entries, extra = [], b""
        entry(256, 4, [w])
        entry(257, 4, [h])
        entry(258, 3, [bits] * spp)
        entry(259, 3, [1])
        entry(262, 3, [2 if spp >= 3 else 1])
        entry(273, 4, offsets)
        entry(277, 3, [spp])
        entry(278, 4, [rps])
        entry(279, 4, [min(rps, h - s * rps) * row_bytes for s in range(n_strips)])
        entry(284, 3, [1])
        if spp == 4:
            entry(338, 3, [2])
        data_base = extra_base + len(extra)
        offsets = [data_base + s * rps * row_bytes for s in range(n_strips)]
    with open(os.fspath(path), "wb") as fh:
        fh.write((b"MM" if big_endian else b"II") + struct.pack(e + "HI", 42, header_len))
        fh.write(struct.pack(e + "H", len(entries)) + b"".join(entries) + struct.pack(e + "I", 0))
        fh.write(extra)
        fh.write(data)


# ------------------------------------------------------------------------------------------
# streaming runner
# ------------------------------------------------------------------------------------------
class SurveyPipeline:
    """Streams equally-shaped frames through the analysis path with decode, H2D, compute and
    D2H overlapped.  One pipeline per thread / GPU; frames are independent units, so N ranks
    each run their own pipeline on their shard (``sources[rank::world]``) and the dataset-wide
    statistics meet in one all-gather (:func:`..distributed.dataset_statistics`).

    ``outputs`` selects what leaves the GPU per frame: "stats" (default, O(1) bytes per frame),
    and any of "wb", "maps", "rgb" -- those are handed, chunk by chunk, to ``on_chunk(first_frame,
    n, host)`` as views of pinned buffers that are only valid during the call.
    """

    def __init__(self, height: int, width: int, channels: int = 3, dtype=np.uint8, chunk: int = 16,
                 depth: int = 3, decode_threads: int = 8, outputs: Sequence[str] = ("stats",),
                 engine: Optional[Engine] = None, white_balance: bool = True,
                 on_chunk: Optional[Callable[[int, int, Dict[str, torch.Tensor]], None]] = None, **fused_kw):
        self.eng = engine or get_engine()
        self.h, self.w, self.c = int(height), int(width), int(channels)
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.dtype(np.uint8), np.dtype(np.uint16)):
            raise LarsError(f"dtype {self.dtype} is not supported (uint8 or uint16 frames)")
        self.sb = self.dtype.itemsize
        self.chunk, self.depth = int(chunk), max(2, int(depth))
        self.outputs = tuple(k for k in ALL_OUTPUTS if k in outputs or k == "stats")
        self.white_balance = white_balance
        self.on_chunk = on_chunk
        self.fused_kw = fused_kw
        self.decode_threads = max(1, int(decode_threads))
        eng = self.eng
        self.s_in, self.s_cmp, self.s_out = (torch.cuda.Stream(device=eng.device) for _ in range(3))
        self.frame_bytes = self.h * self.w * self.c * self.sb
        self.slots = []
        for _ in range(self.depth):
            frames = eng.alloc_frames(self.chunk, self.h, self.w, self.c, self.s_cmp, sample_bytes=self.sb)
            self.slots.append({
                "host_in": torch.empty((self.chunk, self.frame_bytes), dtype=torch.uint8, pin_memory=True),
                "frames": frames,
                "res": eng.alloc_outputs(frames, self.outputs, self.s_cmp),
                "host_out": eng.alloc_host_outputs(self.chunk, self.h, self.w, self.c, self.outputs),
                "ev_in": torch.cuda.Event(), "ev_cmp": torch.cuda.Event(), "ev_out": torch.cuda.Event(),
                "pending": None,
            })
        with torch.cuda.stream(self.s_cmp):
            # [0] = running dataset record, [1] = the chunk's merged record
            self.fold = torch.zeros((2, 3, INDEX_STATS_DTYPE.itemsize), dtype=torch.uint8, device=eng.device)
        self.s_cmp.synchronize()

    # -- host side -------------------------------------------------------------------------
    def _decode_into(self, source: Source, dst_row: torch.Tensor) -> None:
        shape = (self.h, self.w, self.c)
        dst = dst_row.numpy().view(self.dtype).reshape(shape)
        arr = read_frame(source, out=None if isinstance(source, np.ndarray) else dst)
        if arr.shape != shape and not (self.c == 1 and arr.shape == shape[:2]):
            raise ValueError(f"frame of shape {arr.shape} in a pipeline built for {shape}")
        if arr.dtype != self.dtype:
            raise ValueError(f"frame of dtype {arr.dtype} in a pipeline built for {self.dtype}")
        if arr.ctypes.data != dst.ctypes.data:
            np.copyto(dst, arr.reshape(shape))

    def _harvest(self, slot, per_frame: List[np.ndarray]) -> None:
        pend = slot["pending"]
        if pend is None:
            return
        slot["ev_out"].synchronize()
        first, k = pend
        rec = slot["host_out"]["stats"][:k].numpy().view(INDEX_STATS_DTYPE).reshape(k, 3)
        per_frame.append(rec.copy())
        if self.on_chunk is not None:
            self.on_chunk(first, k, {name: (t[:, :k] if name in ("maps", "rgb") else t[:k])
                                     for name, t in slot["host_out"].items()})
        slot["pending"] = None

    # -- the run ---------------------------------------------------------------------------
    def run(self, sources: Iterable[Source], group=None, distributed: Optional[bool] = None) -> dict:
        """Process every source.  Returns {"frames": n, "per_frame": [n, 3] structured records,
        "per_frame_dicts": callable -> list of dicts, "dataset": {index: {...}} over all ranks}."""
        from . import distributed as ld
        import torch.distributed as dist
        eng = self.eng
        per_frame: List[np.ndarray] = []
        with torch.cuda.stream(self.s_cmp):
            self.fold.zero_()
        ready: "queue.Queue" = queue.Queue()
        tokens: "queue.Queue" = queue.Queue()       # one token per free slot; slots are used round-robin
        for _ in range(self.depth):
            tokens.put(True)
        failure: List[BaseException] = []

        def producer():
            try:
                with ThreadPoolExecutor(self.decode_threads) as pool:
                    it = iter(sources)
                    c, first = 0, 0
                    while True:
                        batch = []
                        for src in it:
                            batch.append(src)
                            if len(batch) == self.chunk:
                                break
                        if not batch:
                            break
                        if tokens.get() is None:        # consumer gave up
                            return
                        slot = self.slots[c % self.depth]
                        futs = [pool.submit(self._decode_into, src, slot["host_in"][j]) for j, src in enumerate(batch)]
                        for f in futs:
                            f.result()
                        ready.put((c % self.depth, first, len(batch)))
                        c += 1
                        first += len(batch)
            except BaseException as exc:   # surfaced on the consumer side
                failure.append(exc)
            finally:
                ready.put(None)

        th = threading.Thread(target=producer, name="lars-ingest", daemon=True)
        th.start()
        n_frames = 0
        inflight: List[int] = []
        npx, ch = self.h * self.w, self.c
        try:
            while True:
                item = ready.get()
                if item is None:
                    break
                idx, first, k = item
                slot = self.slots[idx]
                fr: DeviceFrames = slot["frames"]
                view = DeviceFrames(fr.data[:k], fr.n_pixels, fr.channels, fr.shape, fr.sample_bytes)
                with torch.cuda.stream(self.s_in):
                    view.data[:, :self.frame_bytes].copy_(slot["host_in"][:k], non_blocking=True)
                    slot["ev_in"].record(self.s_in)
                res = slot["res"]
                sub = DeviceOutputs(frames=view,
                                    wb=None if res.wb is None else res.wb[:k],
                                    maps=None if res.maps is None else res.maps[:, :k],
                                    rgb=None if res.rgb is None else res.rgb[:, :k],
                                    stats=res.stats[:k])
                with torch.cuda.stream(self.s_cmp):
                    self.s_cmp.wait_event(slot["ev_in"])
                    eng.process_device(view, outputs=self.outputs, white_balance=self.white_balance, out=sub,
                                       stream=self.s_cmp, **self.fused_kw)
                    with torch.cuda.device(eng.device):
                        check(eng.lib.lars_stats_merge(sub.stats.data_ptr(), k, self.fold[1].data_ptr(),
                                                       self.s_cmp.cuda_stream), "lars_stats_merge")
                        check(eng.lib.lars_stats_merge(self.fold.data_ptr(), 2, self.fold[0].data_ptr(),
                                                       self.s_cmp.cuda_stream), "lars_stats_merge")
                    slot["ev_cmp"].record(self.s_cmp)
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(slot["ev_cmp"])
                    ho = slot["host_out"]
                    ho["stats"][:k].copy_(sub.stats[:k], non_blocking=True)
                    if "wb" in ho:
                        ho["wb"][:k].copy_(sub.wb[:k, :npx * ch], non_blocking=True)
                    if "maps" in ho:
                        for i in range(3):
                            ho["maps"][i, :k].copy_(sub.maps[i, :k, :npx], non_blocking=True)
                    if "rgb" in ho:
                        for i in range(3):
                            ho["rgb"][i, :k].copy_(sub.rgb[i, :k, :npx * 3], non_blocking=True)
                    slot["ev_out"].record(self.s_out)
                slot["pending"] = (first, k)
                inflight.append(idx)
                n_frames += k
                # keep depth - 1 chunks in flight on the GPU while the producer fills the remaining slot
                while len(inflight) > self.depth - 1:
                    self._harvest(self.slots[inflight.pop(0)], per_frame)
                    tokens.put(True)
            while inflight:                                  # drain in submission order
                self._harvest(self.slots[inflight.pop(0)], per_frame)
        finally:
            tokens.put(None)
            th.join()
        if failure:
            raise failure[0]
        use_dist = dist.is_initialized() and dist.get_world_size(group) > 1 if distributed is None else distributed
        with torch.cuda.stream(self.s_cmp):
            local = self.fold[0]
            whole = ld.merge_records_device(eng, ld.gather_records(local, group), self.s_cmp) if use_dist else local
            whole_h = whole.cpu()
        self.s_cmp.synchronize()
        records = np.concatenate(per_frame) if per_frame else np.zeros((0, 3), INDEX_STATS_DTYPE)
        bins = self.fused_kw.get("bins", 50)
        dataset = stats_records_to_dicts(whole_h.numpy().view(INDEX_STATS_DTYPE).reshape(1, 3), bins)[0]
        return {"frames": n_frames, "per_frame": records, "dataset": dataset,
                "per_frame_dicts": lambda: stats_records_to_dicts(records, bins)}
