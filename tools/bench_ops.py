"""Device-resident timing of the kernels NEXT to the fused pass (K3..K9), CUDA events on the launching
stream, inputs larger than L2 (batches of maps) so every launch streams from HBM.  Algorithmic bytes per
element are stated per line; peak = MEASURED_PEAKS.json hbm_gbs."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from lars_image_processing_b200 import _lib
from lars_image_processing_b200._lib import INDEX_STATS_DTYPE, check
from lars_image_processing_b200.engine import get_engine

eng = get_engine(); lib = eng.lib; s = eng.stream(); dev = eng.device
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    peak = 6650.0
n = 3000 * 4000
M = 8                                   # maps per timing loop: 8 x 48 MB > L2
g = torch.Generator(device=dev); g.manual_seed(1)
with torch.cuda.stream(s):
    maps = (torch.rand((M, n), generator=g, device=dev) * 2 - 1).float()
    frames = torch.randint(0, 256, (M, n * 3), generator=g, device=dev, dtype=torch.uint8)
    frames16 = torch.randint(0, 65536, (M, n * 3), generator=g, device=dev, dtype=torch.int32).to(torch.int16)
    stats = torch.empty((M, INDEX_STATS_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    ws = torch.empty(int(lib.lars_map_stats_workspace_bytes(M)), dtype=torch.uint8, device=dev)
    sel_ws = torch.empty(int(lib.lars_select_workspace_bytes()), dtype=torch.uint8, device=dev)
    med = torch.empty((M, 3), dtype=torch.float32, device=dev)
    rgb = torch.empty((M, n * 3), dtype=torch.uint8, device=dev)
    f64 = torch.empty((M, n), dtype=torch.float64, device=dev)
    out32 = torch.empty((3, M, n), dtype=torch.float32, device=dev)
sp = s.cuda_stream


def timed(label, bytes_per_el, fn, reps=5):
    with torch.cuda.device(dev):
        fn(); s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(reps):
            fn()
        e1.record(s); s.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gbs = bytes_per_el * n * M / ms / 1e6
    print(f"{label:58s} {ms * 1e3 / M:8.1f} us / 12 MP   {bytes_per_el:3d} B/el  {gbs:7.0f} GB/s  {gbs / peak * 100:5.1f} % of HBM peak", flush=True)


timed("K4 map_stats_f32 (stats + hist(50) of a float map)", 4,
      lambda: check(lib.lars_map_stats_f32(maps.data_ptr(), M, n, n, 50, 0.2, stats.data_ptr(), ws.data_ptr(), ws.numel(), sp)))
timed("K3 select_f32 (exact median, 3 radix passes)", 12,
      lambda: [check(lib.lars_select_f32(maps[i].data_ptr(), n, (n - 1) // 2, n // 2, med[i].data_ptr(), sel_ws.data_ptr(), sel_ws.numel(), sp)) for i in range(M)])
timed("K5 colormap_f32 (float map -> RGB)", 7,
      lambda: [check(lib.lars_colormap_f32(maps[i].data_ptr(), n, 0, -1.0, 1.0, rgb[i].data_ptr(), sp)) for i in range(M)])
timed("K6 ndvi_f64_u8 (raw frame -> float64 NDVI)", 11,
      lambda: [check(lib.lars_ndvi_f64_u8(frames[i].data_ptr(), n, 3, f64[i].data_ptr(), sp)) for i in range(M)])
timed("K7 index_planes_f32 (two float planes -> index)", 12,
      lambda: [check(lib.lars_index_planes_f32(maps[i].data_ptr(), maps[(i + 1) % M].data_ptr(), n, out32[0, i].data_ptr(), sp)) for i in range(M)])
timed("K8 index_hwc uint16 frame -> index", 10,
      lambda: [check(lib.lars_index_hwc(frames16[i].data_ptr(), 1, n, 3, 0, out32[0, i].data_ptr(), sp)) for i in range(M)])
timed("K9 index_change_u8 (2 frames -> 2 maps + diff + bwr RGB)", 21,
      lambda: [check(lib.lars_index_change_u8(frames[i].data_ptr(), frames[(i + 1) % M].data_ptr(), n, 3, 0, -0.5, 0.5,
                                              out32[0, i].data_ptr(), out32[1, i].data_ptr(), out32[2, i].data_ptr(), rgb[i].data_ptr(), sp)) for i in range(M)])

# ---- round 2: the float64 flavour of K4 / K3 and uint16 Pass 1 (guided single pass against the two-level form)
from lars_image_processing_b200._lib import MAP_STATS_F64_DTYPE, STRETCH_U16_BYTES
with torch.cuda.stream(s):
    f64.copy_(maps.double())
    stats64 = torch.empty((M, MAP_STATS_F64_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    ws64 = torch.empty(int(lib.lars_map_stats_f64_workspace_bytes()), dtype=torch.uint8, device=dev)
    sel64 = torch.empty(int(lib.lars_select_f64_workspace_bytes()), dtype=torch.uint8, device=dev)
    med64 = torch.empty((M, 3), dtype=torch.float64, device=dev)
    stretch = torch.empty((M, 3, STRETCH_U16_BYTES), dtype=torch.uint8, device=dev)
    pct = torch.empty((M, 3, 2), dtype=torch.float64, device=dev)
    ws16 = torch.empty(int(lib.lars_wb_u16_workspace_bytes(M)), dtype=torch.uint8, device=dev)
    # vegetation-like uint16 frames (the uniform frames16 above put ~0.4 % of the samples in every bucket)
    veg = (torch.randn((M, n, 3), generator=g, device=dev) * torch.tensor([8995., 8995., 11565.], device=dev)
           + torch.tensor([23130., 28270., 38550.], device=dev)).round_().clamp_(0, 65535).to(torch.int32).to(torch.int16).reshape(M, n * 3)
timed("K4d map_stats_f64 (stats + hist(50), float64 map)", 8,
      lambda: [check(lib.lars_map_stats_f64(f64[i].data_ptr(), n, 50, 0.2, stats64[i].data_ptr(), ws64.data_ptr(), ws64.numel(), sp)) for i in range(M)])
timed("K3d select_f64 (exact median, 6 radix passes)", 48,
      lambda: [check(lib.lars_select_f64(f64[i].data_ptr(), n, (n - 1) // 2, n // 2, med64[i].data_ptr(), sel64.data_ptr(), sel64.numel(), sp)) for i in range(M)])
for label, stage, b in (("K1u uint16 Pass 1, guided single pass (8 frames per launch)", 0, 6),
                        ("K1u uint16 Pass 1, two-level form (8 frames per launch)", 4, 12)):
    timed(label, b * 3 // 3,
          lambda: check(lib.lars_wb_stretch_build_u16_staged(veg.data_ptr(), M, n, 3, n * 6, 0.02, 0.98, stretch.data_ptr(), pct.data_ptr(),
                                                             ws16.data_ptr(), ws16.numel(), 0, stage, sp)))
