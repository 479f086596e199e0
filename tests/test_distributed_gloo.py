"""World-size-2 gloo test of the multi-GPU host logic on CPU: frame sharding, the single
all-gather of packed statistics records and the rank-ordered merge.  The per-frame records are
produced here by the NumPy oracle (the CUDA kernels need a GPU); the product merges them with
lars_stats_merge on the device -- the gathering / ordering logic under test is shared."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _records_for(frames):
    import warnings
    from lars_image_processing_b200._lib import INDEX_STATS_DTYPE
    from oracle import oracle_np as o
    rec = np.zeros((len(frames), 3), INDEX_STATS_DTYPE)
    for f, img in enumerate(frames):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            res = o.analyze_frame(img, want_rgb=False)
        for i, t in enumerate(o.INDEX_TYPES):
            st = res["stats"][t]
            r = rec[f, i]
            r["count"], r["count_above"], r["sum"], r["sumsq"] = st["count"], st["count_above"], st["sum"], st["sumsq"]
            r["min"], r["max"], r["bins"], r["threshold"] = st["min"], st["max"], 50, o.coverage_threshold(t)
            r["hist"][:50] = st["hist"]
    return rec


def _merge_numpy(rec):
    """Same sums as stats_merge_kernel (the kernel folds sets lane-strided with a fixed butterfly; sums agree to rounding)."""
    out = np.zeros(3, rec.dtype)
    for i in range(3):
        for s in range(rec.shape[0]):
            r = rec[s, i]
            if r["count"] == 0:
                continue
            out[i]["count"] += r["count"]
            out[i]["count_above"] += r["count_above"]
            out[i]["sum"] += r["sum"]
            out[i]["sumsq"] += r["sumsq"]
            out[i]["hist"] += r["hist"]
            out[i]["min"] = r["min"] if out[i]["count"] == r["count"] else min(out[i]["min"], r["min"])
            out[i]["max"] = r["max"] if out[i]["count"] == r["count"] else max(out[i]["max"], r["max"])
    return out


def _worker(rank, world, port, n_frames, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from lars_image_processing_b200 import distributed as ld
    from oracle import synth
    r, w, _ = ld.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    b, e = ld.shard_range(n_frames, rank, world)
    frames = [synth.vegetation_frame(500 + i, 24, 40) for i in range(b, e)]
    local = _merge_numpy(_records_for(frames))
    packed = torch.from_numpy(local.view(np.uint8).reshape(3, -1).copy())
    gathered = ld.gather_records(packed)                       # ONE collective
    assert tuple(gathered.shape) == (world, 3, ld.RECORD_BYTES)
    merged = _merge_numpy(gathered.numpy().view(local.dtype).reshape(world, 3))
    hist = torch.from_numpy(np.full((1, 3, 256), rank + 1, np.int64))
    ld.allreduce_wb_histogram(hist)
    if rank == 0:
        q.put((merged.tobytes(), int(hist[0, 0, 0])))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_helpers():
    from lars_image_processing_b200 import distributed as ld
    for n in (0, 1, 7, 8, 1024, 100000):
        for world in (1, 2, 4, 8):
            blocks = [ld.shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in blocks]
            assert max(sizes) - min(sizes) <= 1
            rr = sorted(i for r in range(world) for i in ld.shard_round_robin(n, r, world))
            assert rr == list(range(n))


@pytest.mark.timeout(300)
def test_two_rank_gather_and_merge_matches_single_process():
    from oracle import synth
    n_frames, world = 7, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    [p.start() for p in procs]
    merged_bytes, hist_sum = q.get(timeout=240)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert hist_sum == 3                                   # 1 + 2: SUM all-reduce of the WB histogram
    frames = [synth.vegetation_frame(500 + i, 24, 40) for i in range(n_frames)]
    single = _records_for(frames)
    want = _merge_numpy(single)
    got = np.frombuffer(merged_bytes, dtype=want.dtype)
    for i in range(3):
        for k in ("count", "count_above", "min", "max"):
            assert got[i][k] == want[i][k]
        assert np.array_equal(got[i]["hist"], want[i]["hist"])
        assert abs(got[i]["sum"] - want[i]["sum"]) <= 1e-12 * max(1.0, abs(want[i]["sum"]))
        assert abs(got[i]["sumsq"] - want[i]["sumsq"]) <= 1e-12 * max(1.0, abs(want[i]["sumsq"]))
