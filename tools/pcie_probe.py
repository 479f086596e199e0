"""Host <-> device copy ceiling of this box with 1..N GPUs copying at once.

    python tools/pcie_probe.py                                             # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/pcie_probe.py

Every rank owns one GPU and pinned buffers; all ranks start each leg together (barrier) and the leg's time is
the maximum over ranks, so the printed aggregate is what N concurrent `run_host_batch` callers can share:
D2H alone, H2D alone, and D2H with an H2D stream running beside it (the mix of the bench's `e2e`: 24 B/px back,
3 B/px in).  Rank 0 prints one line per leg.
"""
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

n = 1 << 30
d = torch.empty(n, dtype=torch.uint8, device=dev)
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d2 = torch.empty(n // 8, dtype=torch.uint8, device=dev)
h2 = torch.empty(n // 8, dtype=torch.uint8, pin_memory=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def leg(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    t = torch.tensor([(time.perf_counter() - t0) / reps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def both():
    with torch.cuda.stream(s1):
        h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2):
        d2.copy_(h2, non_blocking=True)


a = leg(lambda: h.copy_(d, non_blocking=True))
b = leg(lambda: d.copy_(h, non_blocking=True))
c = leg(both)
if rank == 0:
    print(f"ranks {world}: D2H 1 GiB pinned per rank: {n / a / 1e9:6.1f} GB/s per rank, {world * n / a / 1e9:7.1f} GB/s aggregate")
    print(f"ranks {world}: H2D 1 GiB pinned per rank: {n / b / 1e9:6.1f} GB/s per rank, {world * n / b / 1e9:7.1f} GB/s aggregate")
    print(f"ranks {world}: D2H 1 GiB + H2D 128 MiB concurrently: {n / c / 1e9:6.1f} GB/s D2H per rank, "
          f"{world * n / c / 1e9:7.1f} GB/s D2H aggregate ({world * (n + n // 8) / c / 1e9:7.1f} GB/s both directions)", flush=True)
if world > 1:
    dist.destroy_process_group()
