"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink / NVSwitch).

The path shards by independent frames: white-balance percentiles are per
frame, so no data crosses GPUs inside a frame.  The only exchange is the dataset-wide
statistics merge at the end: every rank folds its own per-frame records into 3 records
(``lars_stats_merge``), ONE all-gather moves the packed 3 x 576-byte records, and the same
kernel merges them in rank order -- deterministic, SUM and MIN/MAX in a single collective.

For one huge image sharded by tiles (config 4) the white-balance histogram is global to the
image: ``allreduce_wb_histogram`` SUM-reduces the 3 x 256 counters between Pass 1 and the
LUT build so that every rank derives the identical LUT.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from ._lib import INDEX_STATS_DTYPE, LarsError, check

RECORD_BYTES = INDEX_STATS_DTYPE.itemsize


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [begin, end) of ``n_items`` owned by ``rank`` (sizes differ by <= 1)."""
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_round_robin(n_items: int, rank: int, world: int) -> List[int]:
    """Round-robin assignment (frame i -> GPU i % world), the layout of BASELINE config 3."""
    return list(range(rank, n_items, world))


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise the default process group from torchrun's environment -> (rank, world, local_rank)."""
    import os
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        # NCCL writes its version banner / debug log to stdout by default; programs that print results
        # on stdout (bench.py: one JSON line) must not have them mixed in
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, world, local_rank


def merge_records_device(engine, records: torch.Tensor, stream=None) -> torch.Tensor:
    """[S, 3, 576] uint8 device records -> [3, 576] merged (GPU kernel, fixed order)."""
    s = stream or engine.stream()
    n_sets = records.shape[0]
    with torch.cuda.stream(s):
        out = torch.empty((3, RECORD_BYTES), dtype=torch.uint8, device=engine.device)
        with torch.cuda.device(engine.device):
            check(engine.lib.lars_stats_merge(records.data_ptr(), n_sets, out.data_ptr(), s.cuda_stream),
                  "lars_stats_merge")
    return out


def gather_records(local: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather the packed per-rank records -> [world, 3, 576] on every rank (one collective)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local.unsqueeze(0)
    local = local.contiguous()
    # concatenated along dim 0 (the layout every backend accepts), viewed as [world, ...]
    flat = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(flat, local, group=group)
    return flat.view((world,) + tuple(local.shape))


def dataset_statistics(engine, per_frame_records: torch.Tensor, group=None, stream=None) -> torch.Tensor:
    """Per-frame records of this rank -> dataset-wide records over all ranks ([3, 576] uint8)."""
    s = stream or engine.stream()
    local = merge_records_device(engine, per_frame_records, s)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    with torch.cuda.stream(s):
        gathered = gather_records(local, group)       # NCCL runs on the current (= engine) stream
    return merge_records_device(engine, gathered, s)


class AsyncDatasetStatistics:
    """The dataset-statistics exchange taken off the critical path.

    A step's result (3 x 576 bytes per rank) is not an input of the next step, so the all-gather
    and the final merge run on a side stream while the main stream already works on the next
    batch: ``submit`` folds the per-frame records into one of ``depth`` rotating local records on
    the caller's stream (a 5 us kernel), then hands that record to the side stream.  ``result``
    makes a stream wait for the most recent exchange and returns the dataset-wide records.
    Ranks no longer meet once per step, so host-side jitter of one rank does not stall the others.
    """

    def __init__(self, engine, group=None, depth: int = 2, timing: bool = False):
        """``timing``: keep a CUDA-event pair around every all-gather (``gather_us()`` reads them)."""
        self.engine, self.group, self.depth = engine, group, max(2, int(depth))
        self.timing = bool(timing)
        self.gather_events: list = []
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        dev = engine.device
        self.side = torch.cuda.Stream(device=dev)
        self.local = [torch.zeros((3, RECORD_BYTES), dtype=torch.uint8, device=dev) for _ in range(self.depth)]
        self.gathered = [torch.zeros((self.world * 3, RECORD_BYTES), dtype=torch.uint8, device=dev)
                         for _ in range(self.depth)] if self.world > 1 else None
        self.out = [torch.zeros((3, RECORD_BYTES), dtype=torch.uint8, device=dev) for _ in range(self.depth)] \
            if self.world > 1 else self.local
        self.done = [None] * self.depth
        self.count = 0
        torch.cuda.synchronize(dev)

    def submit(self, per_frame_records: torch.Tensor, stream=None) -> None:
        eng = self.engine
        s = stream or eng.stream()
        k = self.count % self.depth
        self.count += 1
        if self.done[k] is not None:
            s.wait_event(self.done[k])                    # slot k's previous exchange has been consumed
        with torch.cuda.device(eng.device):
            check(eng.lib.lars_stats_merge(per_frame_records.data_ptr(), per_frame_records.shape[0],
                                           self.local[k].data_ptr(), s.cuda_stream), "lars_stats_merge")
        ev = torch.cuda.Event()
        ev.record(s)
        if self.world == 1:
            self.done[k] = ev
            return
        self.side.wait_event(ev)
        with torch.cuda.stream(self.side):
            if self.timing:
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record(self.side)
            dist.all_gather_into_tensor(self.gathered[k], self.local[k], group=self.group)
            if self.timing:
                t1.record(self.side)
                self.gather_events.append((t0, t1))
            with torch.cuda.device(eng.device):
                check(eng.lib.lars_stats_merge(self.gathered[k].data_ptr(), self.world, self.out[k].data_ptr(),
                                               self.side.cuda_stream), "lars_stats_merge")
            done = torch.cuda.Event()
            done.record(self.side)
        self.done[k] = done

    def gather_us(self, reset: bool = True) -> List[float]:
        """Device time of every timed all-gather so far, in microseconds (synchronises the side stream)."""
        self.side.synchronize()
        out = [a.elapsed_time(b) * 1e3 for a, b in self.gather_events]
        if reset:
            self.gather_events = []
        return out

    def result(self, stream=None) -> torch.Tensor:
        """Dataset-wide records of the latest submitted step; ``stream`` waits for them."""
        if self.count == 0:
            raise RuntimeError("no statistics have been submitted")
        k = (self.count - 1) % self.depth
        (stream or self.engine.stream()).wait_event(self.done[k])
        return self.out[k]


def allreduce_wb_histogram(hist: torch.Tensor, group=None) -> torch.Tensor:
    """SUM-allreduce of the [*, 3, 256] int64 white-balance histogram (tile-sharded image)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    return hist


def process_mosaic_tiles(engine, tiles, outputs=("wb", "maps", "rgb", "stats"), group=None, stream=None, **kw):
    """One huge image (orthomosaic, BASELINE config 4) held as equally-sized tiles / row bands.

    Every rank passes the tiles it owns.  Pass 1 accumulates ONE histogram over the local tiles,
    a SUM all-reduce makes it global, every rank builds the identical LUT, Pass 2 runs per tile
    with that shared LUT, and the per-tile statistics are merged into image-wide records with
    the usual single all-gather.  Returns (DeviceOutputs of the local tiles, [3, 576] records)."""
    s = stream or engine.stream()
    res = engine.process_device(tiles, outputs=outputs, tiles_of_one_image=True,
                                hist_hook=lambda h: allreduce_wb_histogram(h, group), stream=s, **kw)
    whole = dataset_statistics(engine, res.stats, group, s) if res.stats is not None else None
    return res, whole


class PeerHistogramExchange:
    """The white-balance exchange of a tile-sharded image done by the LUT kernel itself over NVLink peer memory
    (``lars_wb_lut_build_u8_peers``) instead of an NCCL all-reduce between Pass 1 and the LUT build: 3 x 256 counters are
    pure latency, and one kernel that pushes them into every peer's buffer, waits for the peers' flags and sums in
    rank order replaces two launches and the collective's protocol.

    Every rank allocates one symmetric buffer (``torch.distributed._symmetric_memory``: CUDA VMM allocations that all
    ranks of the node map into their address space), so the kernel gets a device array of ``world`` pointers.
    Collective: construct it on every rank of the group.  Raises ``LarsError`` when symmetric memory cannot be set up
    on this box -- callers then keep the NCCL hook (``allreduce_wb_histogram``)."""

    def __init__(self, engine, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        if not (dist.is_initialized() and dist.get_world_size(group) > 1):
            raise LarsError("PeerHistogramExchange needs an initialised process group with more than one rank")
        self.engine = engine
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        nbytes = int(engine.lib.lars_wb_peer_buffer_bytes(self.world))
        try:
            with torch.cuda.device(engine.device):
                try:
                    symm_mem.set_backend("CUDA")                  # plain CUDA VMM / IPC mappings (NVSHMEM is not in this image)
                except Exception:
                    pass
                self.buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=engine.device)
                self.buf.zero_()
                torch.cuda.synchronize(engine.device)
                self.handle = symm_mem.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
            self.peer_ptrs_dev = int(self.handle.buffer_ptrs_dev)
        except Exception as exc:                                   # no VMM / fabric support, old driver, ...
            raise LarsError(f"symmetric memory is not available here: {exc}") from exc
        self.status = torch.zeros(1, dtype=torch.int32, device=engine.device)
        self.epoch = 0
        dist.barrier(group)                                        # every rank's buffer is zeroed and mapped

    def lut_build(self, hist: torch.Tensor, lut: torch.Tensor, pct: Optional[torch.Tensor], quantiles=(0.02, 0.98),
                  stream=None, chain: int = 0) -> None:
        """hist [1, 3, 256] int64: local counts in, image-wide counts out; lut [1, 3, 256] uint8, pct [1, 3, 2] float64."""
        s = stream or self.engine.stream()
        self.epoch += 1
        with torch.cuda.device(self.engine.device):
            check(self.engine.lib.lars_wb_lut_build_u8_peers(
                hist.data_ptr(), self.peer_ptrs_dev, self.rank, self.world, self.epoch & 0xFFFFFFFF or 1,
                float(quantiles[0]), float(quantiles[1]), int(chain), lut.data_ptr(),
                pct.data_ptr() if pct is not None else None, self.status.data_ptr(), s.cuda_stream),
                "lars_wb_lut_build_u8_peers")

    def timed_out(self) -> bool:
        """True if any exchange so far gave up waiting for a peer (synchronises)."""
        return bool(int(self.status.item()))


def records_to_numpy(records: torch.Tensor) -> np.ndarray:
    """Device records -> structured NumPy array.  Synchronises the device first: the records are
    usually produced on an engine stream, and ``.cpu()`` only orders against the current one."""
    if records.is_cuda:
        torch.cuda.synchronize(records.device)
    return records.cpu().numpy().view(INDEX_STATS_DTYPE).reshape(records.shape[:-1])
