"""The multi-GPU paths on real hardware: needs >= 2 visible B200s (skipped otherwise), one process per
GPU (all of them, up to 8), NCCL.  (1) a uint8 and a uint16 mosaic whose row bands are sharded over the ranks -- SUM all-reduce
of the white-balance counters between the stages, one all-gather of the image-wide statistics --
against the oracle on the whole image; (2) the side-stream dataset-statistics exchange; (3) a
SurveyPipeline per rank over a round-robin shard of a frame list."""
import os
import socket
import sys
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle(img):
    from oracle import oracle_np as o
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return o.analyze_frame(img)


def _mosaic(dtype):
    from oracle import synth
    img = synth.vegetation_frame(4242, 768, 640, dtype)
    img[:100] //= 4                                     # bands differ: per-rank percentiles would be wrong
    return img


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                          LOCAL_RANK=str(rank))
        import torch
        import torch.distributed as dist
        from lars_image_processing_b200 import distributed as ld, ingest
        from lars_image_processing_b200.engine import Engine, stats_records_to_dicts
        from oracle import oracle_np as o, synth
        r, w, lr = ld.init_from_env(backend="nccl")
        torch.cuda.set_device(lr)
        eng = Engine(lr)
        s = eng.stream()
        report = {}
        # ---- (1) mosaics: 12 row bands, contiguous block per rank
        for dtype in (np.uint8, np.uint16):
            img = _mosaic(dtype)
            bands = [np.ascontiguousarray(b) for b in np.split(img, 12, axis=0)]
            b0, b1 = ld.shard_range(len(bands), rank, world)
            dev = eng.upload(bands[b0:b1], stream=s)
            res, whole = ld.process_mosaic_tiles(eng, dev, stream=s)
            out = eng.download(res, stream=s)
            want = _oracle(img)
            rows = slice(b0 * 64, b1 * 64)
            ok = np.array_equal(np.concatenate([x["wb"] for x in out], axis=0), want["wb"][rows])
            for t in o.INDEX_TYPES:
                got = np.concatenate([x["maps"][t] for x in out], axis=0)
                ok &= np.array_equal(got.view(np.uint32), want["maps"][t][rows].view(np.uint32))
            st = stats_records_to_dicts(ld.records_to_numpy(whole).reshape(1, 3), 50)[0]
            for t in o.INDEX_TYPES:
                ws = want["stats"][t]
                ok &= st[t]["count"] == ws["count"] and st[t]["count_above"] == ws["count_above"]
                ok &= np.array_equal(st[t]["hist"], ws["hist"]) and st[t]["min"] == ws["min"] and st[t]["max"] == ws["max"]
            report[f"mosaic_{np.dtype(dtype).name}"] = bool(ok)
        # ---- (1b) the uint8 mosaic again through the pre-bound plan (bench.py c4 `value`) and through the pinned
        #      host path (bench.py c4 `e2e`), the histogram all-reduce as the hook of both
        from lars_image_processing_b200.engine import ALL_OUTPUTS, FramePlan
        img = _mosaic(np.uint8)
        bands = [np.ascontiguousarray(b) for b in np.split(img, 12, axis=0)]
        b0, b1 = ld.shard_range(len(bands), rank, world)
        want = _oracle(img)
        rows = slice(b0 * 64, b1 * 64)
        hook = lambda hist: dist.all_reduce(hist, op=dist.ReduceOp.SUM)
        plan = FramePlan(eng, eng.upload(bands[b0:b1], stream=s), ALL_OUTPUTS, stream=s, tiles_of_one_image=True, hist_hook=hook)
        out = eng.download(plan.run(), stream=s)
        ok = np.array_equal(np.concatenate([x["wb"] for x in out], axis=0), want["wb"][rows])
        ok &= np.array_equal(np.concatenate([x["rgb"]["NDWI"] for x in out], axis=0), want["rgb"]["NDWI"][rows])
        report["mosaic_plan"] = bool(ok)
        # ... and with the exchange done by the LUT kernel itself over NVLink peer memory (three steps: the slot sets
        # alternate by epoch parity); skipped with a note where symmetric memory cannot be set up
        try:
            peer = ld.PeerHistogramExchange(eng)
        except Exception as exc:
            peer = None
            report["peer_note"] = f"peer exchange unavailable: {exc}"[:300]
        if peer is not None:
            plan_p = FramePlan(eng, eng.upload(bands[b0:b1], stream=s), ALL_OUTPUTS, stream=s, tiles_of_one_image=True,
                               peer_exchange=peer)
            ok = True
            for _ in range(3):
                res_p = plan_p.run()
                out = eng.download(res_p, stream=s)
                ok &= np.array_equal(np.concatenate([x["wb"] for x in out], axis=0), want["wb"][rows])
                ok &= np.array_equal(np.concatenate([x["maps"]["NDVI"] for x in out], axis=0).view(np.uint32),
                                     want["maps"]["NDVI"][rows].view(np.uint32))
                whole_hist = res_p.wb_hist.cpu().numpy().reshape(3, 256)
                ok &= np.array_equal(whole_hist, o.channel_histograms(img))
            report["mosaic_peer_exchange"] = bool(ok) and not peer.timed_out()
        T = b1 - b0
        host_in = torch.empty((T, 64 * 640 * 3), dtype=torch.uint8, pin_memory=True)
        for i in range(T):
            host_in[i].copy_(torch.from_numpy(bands[b0 + i].reshape(-1)))
        host_out = eng.alloc_host_outputs(T, 64, 640, 3)
        stats = eng.run_host_mosaic(host_in, (64, 640, 3), host_out, chunk=1, hist_hook=hook)
        whole = stats_records_to_dicts(ld.records_to_numpy(ld.dataset_statistics(eng, stats)).reshape(1, 3), 50)[0]
        ok = np.array_equal(host_out["wb"].numpy().reshape(T * 64, 640, 3), want["wb"][rows])
        ok &= np.array_equal(host_out["maps"][1].numpy().reshape(T * 64, 640).view(np.uint32),
                             want["maps"]["GNDVI"][rows].view(np.uint32))
        ok &= np.array_equal(whole["GNDVI"]["hist"], want["stats"]["GNDVI"]["hist"])
        report["mosaic_host"] = bool(ok)
        # ---- (2) side-stream exchange over three steps
        ex = ld.AsyncDatasetStatistics(eng)
        frames = [synth.vegetation_frame(9000 + 10 * rank + i, 48, 64) for i in range(3)]
        dev = eng.upload(frames, stream=s)
        for _ in range(3):
            res = eng.process_device(dev, outputs=("stats",), stream=s)
            ex.submit(res.stats, stream=s)
        rec = ld.records_to_numpy(ex.result(s))
        s.synchronize()
        report["async_count"] = int(rec["count"][0]) == world * 3 * 48 * 64
        all_frames = [synth.vegetation_frame(9000 + 10 * rk + i, 48, 64) for rk in range(world) for i in range(3)]
        hist = sum(_oracle(f)["stats"]["NDVI"]["hist"] for f in all_frames)
        report["async_hist"] = bool(np.array_equal(rec["hist"][0][:50], hist))
        # ---- (3) survey pipeline per rank, round-robin shard, dataset statistics over all ranks
        survey = [synth.vegetation_frame(9500 + i, 40, 56) for i in range(11)]
        mine = [survey[i] for i in ld.shard_round_robin(len(survey), rank, world)]
        pipe = ingest.SurveyPipeline(40, 56, chunk=2, engine=eng)
        got = pipe.run(mine)
        total = sum(_oracle(f)["stats"]["GNDVI"]["hist"] for f in survey)
        report["survey_frames"] = got["frames"] == len(mine)
        report["survey_dataset"] = bool(np.array_equal(got["dataset"]["GNDVI"]["hist"], total)) and \
            got["dataset"]["GNDVI"]["count"] == len(survey) * 40 * 56
        q.put((rank, report))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as exc:                                # pragma: no cover
        import traceback
        q.put((rank, {"exception": traceback.format_exc()}))
        raise


@pytest.mark.timeout(600)
def test_mosaic_exchange_and_survey_over_nccl():
    import torch
    import torch.multiprocessing as mp
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    world = min(8, torch.cuda.device_count())      # every GPU of the box: 2 on the builder's lease, 8 on the scaling box
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    reports = dict(q.get(timeout=500) for _ in range(world))
    [p.join(timeout=120) for p in procs]
    for rank, rep in sorted(reports.items()):
        assert "exception" not in rep, rep.get("exception")
        note = rep.pop("peer_note", None)
        if note:
            print(f"rank {rank}: {note}")
        assert all(rep.values()), (rank, rep)
    assert all(p.exitcode == 0 for p in procs)
