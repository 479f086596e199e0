"""Latency of the reference-facing drop-in helpers (host NumPy arrays in, host arrays out, pageable
memory, one frame per call -- exactly how the Streamlit app calls them) next to the NumPy oracle
port of the same helper on one host core.  BASELINE config 2 frame (4000x3000 uint8) by default."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from PIL import Image
from oracle import oracle_np as o, resize_np, synth
from lars_image_processing_b200 import process_images as pi, process_ndvi as pn

h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3000, 4000)
img = synth.vegetation_frame(2, h, w)


def timed(fn, reps):
    out = fn()                                     # warm-up (plans, pools, page faults of the first call)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        ts.append(time.perf_counter() - t0)
    return sorted(ts)[len(ts) // 2] * 1e3, out     # median


rows = []
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    g_ms, wb = timed(lambda: pi.fix_white_balance(img), 5)
    c_ms, wb_ref = timed(lambda: o.fix_white_balance_literal(img), 1)
    rows.append(("fix_white_balance(img)", g_ms, c_ms, np.array_equal(wb, wb_ref)))
    g_ms, ndvi = timed(lambda: pi.calculate_index(wb, "NDVI"), 5)
    c_ms, ndvi_ref = timed(lambda: o.calculate_index(wb_ref, "NDVI"), 1)
    rows.append(("calculate_index(wb, 'NDVI')", g_ms, c_ms, np.array_equal(ndvi.view(np.uint32), ndvi_ref.view(np.uint32))))
    g_ms, st = timed(lambda: pi.analyze_index(ndvi, "NDVI"), 5)
    c_ms, st_ref = timed(lambda: o.analyze_index(ndvi_ref, "NDVI"), 1)
    rows.append(("analyze_index(ndvi, 'NDVI')  [incl. exact median]", g_ms, c_ms,
                 st["Median NDVI"] == st_ref["Median NDVI"] and st["Min NDVI"] == st_ref["Min NDVI"]))
    g_ms, viz = timed(lambda: pi.create_index_visualization(ndvi, "NDVI"), 5)
    c_ms, viz_ref = timed(lambda: o.apply_colormap(ndvi_ref, "NDVI"), 1)
    rows.append(("create_index_visualization(ndvi)  [colormap only]", g_ms, c_ms, np.array_equal(np.array(viz), viz_ref)))
    g_ms, s6 = timed(lambda: pn.analyze_ndvi_statistics(ndvi), 5)
    c_ms, s6_ref = timed(lambda: o.analyze_ndvi_statistics(ndvi_ref), 1)
    rows.append(("analyze_ndvi_statistics(ndvi)", g_ms, c_ms, s6["median_ndvi"] == s6_ref["median_ndvi"]))
    g_ms, small = timed(lambda: pi.preprocess_large_image(img, 1024), 5)
    c_ms, small_ref = timed(lambda: np.array(Image.fromarray(img).resize((small.shape[1], small.shape[0]),
                                                                          Image.Resampling.LANCZOS)), 1)
    rows.append(("preprocess_large_image(img, 1024)  [vs Pillow]", g_ms, c_ms, np.array_equal(small, small_ref)))
    g_ms, fr = timed(lambda: pi.analyze_frame(img), 5)
    c_ms, fr_ref = timed(lambda: o.reference_cpu_path(img), 1)
    rows.append(("analyze_frame(img): WB + 3 maps + 3 RGB + stats, one trip", g_ms, c_ms, np.array_equal(fr["wb"], wb_ref)))
    g_ms, fr2 = timed(lambda: pi.analyze_frame(img, outputs=("stats",)), 5)
    rows.append(("analyze_frame(img, outputs=('stats',))", g_ms, float("nan"), True))

print(f"frame {w}x{h} uint8, host arrays in pageable memory; host has {os.cpu_count()} cores")
print(f"{'helper':60s} {'B200 ms':>9s} {'NumPy ms':>9s} {'ratio':>7s}  parity")
for name, g, c, ok in rows:
    print(f"{name:60s} {g:9.2f} {c:9.1f} {c / g:7.1f}  {'bit-exact' if ok else 'MISMATCH'}")
