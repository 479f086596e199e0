"""C-ABI library loads, exports every symbol include/lars_b200.h declares, struct layouts match
the ctypes mirrors, host-only helpers agree with the oracle, and the drop-in helpers keep the
reference's error conventions -- all without a GPU (no compute calls)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "lars_b200.h")


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lars_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from lars_image_processing_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 18
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/lars_b200.h but not exported"
    assert set(_lib.EXPORTED_SYMBOLS) == set(declared)
    assert lib.lars_abi_version() == 7


def test_struct_layouts_match_header(tmp_path):
    from lars_image_processing_b200 import _lib
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "lars_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu\\n", sizeof(lars_fused_args), sizeof(lars_index_stats),'
                   'offsetof(lars_fused_args, maps), offsetof(lars_fused_args, stats), offsetof(lars_index_stats, hist));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    a, b, c, d, e = map(int, subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split())
    assert a == C.sizeof(_lib.FusedArgs)
    assert b == _lib.INDEX_STATS_DTYPE.itemsize == 576
    assert c == _lib.FusedArgs.maps.offset and d == _lib.FusedArgs.stats.offset
    assert e == _lib.INDEX_STATS_DTYPE.fields["hist"][1]
    # the structs of the rows next to the path: resize plan, TIFF info, uint16 stretch record
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "lars_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(lars_resize_plan), offsetof(lars_resize_plan, table_bytes),'
                   'offsetof(lars_resize_plan, mma_table_offset), sizeof(lars_tiff_info), offsetof(lars_tiff_info, frame_bytes),'
                   'sizeof(lars_stretch_u16));'
                   'printf("%zu %zu %zu %zu\\n", offsetof(lars_tiff_info, tile_width), offsetof(lars_tiff_info, bigtiff),'
                   'sizeof(lars_png_info), offsetof(lars_png_info, frame_bytes));'
                   'printf("%zu %zu\\n", sizeof(lars_lzw_chunk), offsetof(lars_lzw_chunk, dst_bytes));'
                   'printf("%zu %zu %zu\\n", sizeof(lars_map_record_f64), offsetof(lars_map_record_f64, max), offsetof(lars_map_record_f64, hist));return 0;}\n')
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    a, b, c, d, e, f, g, h, i, j, k, l, m, n, q = map(int, subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split())
    assert m == _lib.MAP_STATS_F64_DTYPE.itemsize == 592
    assert n == _lib.MAP_STATS_F64_DTYPE.fields["max"][1] and q == _lib.MAP_STATS_F64_DTYPE.fields["hist"][1]
    assert a == C.sizeof(_lib.ResizePlan) and b == _lib.ResizePlan.table_bytes.offset
    assert c == _lib.ResizePlan.mma_table_offset.offset
    assert d == C.sizeof(_lib.TiffInfo) and e == _lib.TiffInfo.frame_bytes.offset
    assert f == _lib.STRETCH_U16_BYTES
    assert g == _lib.TiffInfo.tile_width.offset and h == _lib.TiffInfo.bigtiff.offset
    assert i == C.sizeof(_lib.PngInfo) and j == _lib.PngInfo.frame_bytes.offset
    assert k == _lib.LZW_CHUNK_DTYPE.itemsize and l == _lib.LZW_CHUNK_DTYPE.fields["dst_bytes"][1]


def test_host_tables_match_oracle():
    from lars_image_processing_b200 import _lib
    from oracle import oracle_np as o
    for name in ("RdYlGn", "RdYlBu", "bwr"):
        assert np.array_equal(_lib.colormap_table(name), o.colormap_lut(name)), name
    for bins in list(range(1, 65)):
        assert np.array_equal(_lib.histogram_edges(bins), o.histogram_edges(bins)), bins


def test_compute_entry_points_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from lars_image_processing_b200 import _lib
    from lars_image_processing_b200.engine import Engine
    lib = _lib.load()
    assert lib.lars_init(0) < 0 and b"failed" in lib.lars_last_error()
    hist = np.zeros(768, np.uint64)
    assert lib.lars_wb_hist_u8(1, 1, 16, 3, 48, hist.ctypes.data, 0, None) < 0
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine()


def test_dropin_error_conventions_need_no_gpu():
    from lars_image_processing_b200 import process_images as pi
    empty = np.zeros((0, 0, 3), np.uint8)
    assert pi.fix_white_balance(None) is None and pi.fix_white_balance(empty) is None
    assert pi.calculate_index(None, "NDVI") is None and pi.calculate_index(empty, "NDVI") is None
    assert pi.analyze_index(None, "NDVI") == {} and pi.analyze_index(np.zeros((0,)), "NDVI") == {}
    assert pi.create_index_visualization(None, "NDVI") is None
    assert pi.analyze_frame(None) is None
    with pytest.raises(ValueError, match="Unknown index type: EVI"):
        pi.calculate_index(np.zeros((4, 4, 3), np.uint8), "EVI")


def test_frame_validation_mirrors_reference_exceptions():
    from lars_image_processing_b200.engine import Engine
    from lars_image_processing_b200._lib import LarsError
    with pytest.raises(IndexError):
        Engine._check_frame(np.zeros((4, 4), np.uint8))           # reference: img[:, :, i] on 2-D
    with pytest.raises(IndexError):
        Engine._check_frame(np.zeros((4, 4, 2), np.uint8))
    with pytest.raises(LarsError):
        Engine._check_frame(np.zeros((4, 4, 3), np.float64))
    assert Engine._check_frame(np.zeros((4, 4, 4), np.uint8)).shape == (4, 4, 4)


def test_stats_record_decoding():
    from lars_image_processing_b200._lib import INDEX_STATS_DTYPE
    from lars_image_processing_b200.engine import stats_records_to_dicts
    rec = np.zeros((1, 3), INDEX_STATS_DTYPE)
    rec[0, 0]["count"], rec[0, 0]["count_above"], rec[0, 0]["mean"] = 200, 50, 0.25
    rec[0, 0]["hist"][:50] = np.arange(50)
    d = stats_records_to_dicts(rec, 50)[0]["NDVI"]
    assert d["coverage_pct"] == 25.0 and d["mean"] == 0.25 and d["hist"].sum() == np.arange(50).sum()
    assert stats_records_to_dicts(rec, 50)[0]["NDWI"]["coverage_pct"] == 0.0


def test_oracle_is_not_imported_by_the_product():
    pkg = os.path.join(ROOT, "lars_image_processing_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "/root/reference" not in text.replace("``/root/reference/process-images.py``", ""), f


def _run_bench(*args, env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          env=env, timeout=600)


def test_bench_reference_arm_contract_on_cpu():
    """`bench.py --impl reference` needs no GPU: one JSON line with the contract's keys, the CPU numbers
    described (kind / cores / sample), no device traffic claimed; ranks other than 0 print nothing."""
    import json
    r = _run_bench("--impl", "reference", "--steps", "1", "--warmup", "1", "--height", "96", "--width", "128")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mpix/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert d["metric"].startswith("RGNir Mpix/s") and "workload" in d["config"] and d["data"] == "synthetic"
    # the keys the GPU arm's config carries for the same workload (the driver compares them)
    assert d["config"]["frames_per_gpu"] == 16 and d["config"]["height"] == 96 and d["config"]["width"] == 128
    cb = d["cpu_baseline"]
    # "reference" = the reference's own functions from oracle/_ref (built where /root/reference exists), else the port
    have_ref = os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "process_images.py"))
    assert cb["kind"] == ("reference" if have_ref else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and "frames" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0
    quiet = _run_bench("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--height", "96", "--width", "128",
                       env_extra={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert quiet.returncode == 0 and quiet.stdout.strip() == ""


def test_bench_gpu_arm_fails_loudly_without_a_gpu():
    """No CPU fallback anywhere on the product path: without CUDA the GPU arm exits non-zero and prints no
    bench line (a silent fallback would put a CPU number on the record)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for the CPU-only container")
    r = _run_bench("--steps", "1", "--warmup", "1", "--height", "96", "--width", "128", "--frames", "1")
    assert r.returncode != 0
    assert not any(ln.lstrip().startswith("{") for ln in r.stdout.splitlines())
    assert "no CPU fallback" in r.stderr or "CUDA" in r.stderr or "NVIDIA" in r.stderr


def test_bench_frame_generator_is_identical_under_numpy_and_torch():
    """bench.py's counter-based generator: the host (NumPy) and device (torch) forms give the same bytes, so the
    GPU arm, the reference arm and the CPU baseline all see the same pixels (uint8 and uint16, batched fill)."""
    import torch
    import bench
    from lars_image_processing_b200.engine import DeviceFrames
    for sb in (1, 2):
        h, w = 37, 53
        npx = h * w
        ppx = (npx + 15) // 16 * 16
        fr = DeviceFrames(torch.zeros((3, ppx * 3 * sb), dtype=torch.uint8), npx, 3, (h, w), sb)
        seeds = [5, 3000, 2_000_017]
        bench.counter_fill_device(fr, seeds, batch_samples=npx * 3 * 2)
        for i, sd in enumerate(seeds):
            ref = bench.counter_frame_np(sd, h, w, sb)
            got = fr.data[i, :npx * 3 * sb].numpy().view(np.uint8 if sb == 1 else np.uint16).reshape(h, w, 3)
            assert np.array_equal(ref, got)
    big = bench.counter_frame_np(7, 300, 400, 1).reshape(-1, 3).astype(np.float64)
    assert np.allclose(big.mean(0), (90, 110, 150), atol=1.0) and np.allclose(big.std(0), (35, 35, 45), atol=1.0)


def test_reference_extract_matches_the_ast_loaded_functions():
    """oracle/_ref (oracle/build_ref.py: the reference's functions cut out of its files, text unmodified) behaves like
    the AST-loaded functions and like the port; skipped where neither the extract nor the checkout exists."""
    import warnings
    from oracle import build_ref, oracle_np as o, synth
    if build_ref.available():
        build_ref.build()
    if not os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "process_images.py")):
        pytest.skip("no oracle/_ref and no reference checkout")
    from oracle._ref import process_images as R
    for name, img in list(synth.adversarial_frames().items()) + [("veg", synth.vegetation_frame(5, 50, 70))]:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            wb = R.fix_white_balance(img)
            assert np.array_equal(wb, o.fix_white_balance_literal(img)), name
            a, b = o.reference_cpu_path(img, reference=R), o.reference_cpu_path(img)
        for t in o.INDEX_TYPES:
            assert a[1][t]["std"] == b[1][t]["std"] and np.array_equal(a[1][t]["hist"], b[1][t]["hist"])
            assert {k: v for k, v in a[1][t].items() if k not in ("hist", "rgb")} == \
                {k: v for k, v in b[1][t].items() if k not in ("hist", "rgb")}


def test_bench_workload_sharding_covers_every_frame_once():
    """bench.py's layout of the four workloads over 1 / 2 / 4 / 8 ranks (host logic, no GPU): c3 round-robin and c4 row
    bands partition the job's frames exactly, c2 and c5's device rings get disjoint seeds per rank, and the reference
    arm's config carries the same keys as the GPU arm's."""
    import types
    import bench
    for world in (1, 2, 3, 4, 8):
        for wl in ("c2", "c3", "c4", "c5"):
            a = types.SimpleNamespace(workload=wl, w=dict(bench.WORKLOADS[wl]))
            per_rank = [bench.frame_seeds(a, r, world) for r in range(world)]
            flat = [s for seeds in per_rank for s in seeds]
            assert len(set(flat)) == len(flat), (wl, world)                       # no frame on two ranks
            w = a.w
            if wl in ("c3", "c4"):
                assert sorted(flat) == [w["seed0"] + i for i in range(w["total_frames"])], (wl, world)
                assert max(map(len, per_rank)) - min(map(len, per_rank)) <= 1
            if wl == "c3":
                assert per_rank[0][:2] == [w["seed0"], w["seed0"] + world][:len(per_rank[0][:2])]   # frame i -> rank i % world
            if wl == "c4" and world > 1:
                assert per_rank[0] == list(range(w["seed0"], w["seed0"] + len(per_rank[0])))        # contiguous band
            if wl == "c2":
                assert all(len(p) == w["frames_per_gpu"] for p in per_rank)
            if wl == "c5":
                assert all(len(p) == w["ring_groups"] * w["group"] for p in per_rank)
            cfg = bench.common_config(a, world)
            assert {"workload", "height", "width", "frames_per_gpu"} <= set(cfg)
    from lars_image_processing_b200 import distributed as ld
    for n, world in ((100000, 8), (64, 8), (7, 3), (2, 4)):
        ranges = [ld.shard_range(n, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))


def test_peer_exchange_refuses_a_single_process():
    """PeerHistogramExchange is a collective over more than one rank; without a process group it raises LarsError
    (callers then keep the NCCL hook) instead of failing inside the kernel."""
    from lars_image_processing_b200 import distributed as ld
    from lars_image_processing_b200._lib import LarsError, load
    with pytest.raises(LarsError):
        ld.PeerHistogramExchange(engine=None)
    lib = load()
    assert lib.lars_wb_peer_buffer_bytes(8) >= 2 * 8 * 768 * 8 + 8 * 3 * 4 and lib.lars_wb_peer_buffer_bytes(8) % 256 == 0
    assert lib.lars_wb_peer_buffer_bytes(0) == 0
