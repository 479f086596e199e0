"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the UNMODIFIED reference.

Run in the authoring container (where /root/reference exists):

    python -m oracle.gen_golden

Every output below is produced by the reference's own functions (loaded by
``oracle/ref_loader.py``) -- fix_white_balance / calculate_index / analyze_index from
process-images.py, calculate_index / fix_white_balance from backend-process.py,
analyze_ndvi_statistics from process-ndvi.py -- plus the NumPy calls the reference makes inline
(np.std, np.histogram(bins=50, range=(-1, 1)), process-ndvi.py:65,:97).  The fixtures travel
to the GPU box; /root/reference does not.
"""
from __future__ import annotations

import hashlib
import os
import warnings

import numpy as np

from . import ref_loader, synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
INDEX_TYPES = ("NDVI", "GNDVI", "NDWI")


def _frames():
    frames = dict(synth.adversarial_frames())
    frames["veg_48x64"] = synth.vegetation_frame(101, 48, 64)
    frames["veg_33x65"] = synth.vegetation_frame(102, 33, 65)
    frames["smooth_40x56"] = synth.smooth_frame(103, 40, 56)
    frames["veg_u16_24x32"] = synth.vegetation_frame(104, 24, 32, np.uint16)
    return frames


def frame_golden(img, R):
    out = {"input": img}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        wb = R["fix_white_balance"](img)
    out["wb"] = wb
    for name in INDEX_TYPES:
        m = R["calculate_index"](wb, name)
        st = R["analyze_index"](m, name)
        out[f"map_{name}"] = m
        out[f"stats_{name}"] = np.array(list(st.values()), dtype=np.float64)   # mean, median, min, max, coverage
        out[f"std_{name}"] = np.float64(np.std(m))
        out[f"hist_{name}"] = np.histogram(m.flatten(), bins=50, range=(-1, 1))[0].astype(np.int64)
    out["percentiles"] = np.array([np.percentile(img[:, :, c].astype(np.float32), (2, 98)) for c in range(3)])
    return out


def main():
    if not ref_loader.available():
        raise SystemExit("reference not present; golden vectors can only be generated in the authoring container")
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    R = ref_loader.load("process-images.py")
    Rb = ref_loader.load("backend-process.py")
    Rn = ref_loader.load("process-ndvi.py")

    pack = {}
    for fname, img in _frames().items():
        for k, v in frame_golden(img, R).items():
            pack[f"{fname}/{k}"] = v
    np.savez_compressed(os.path.join(GOLDEN_DIR, "frames.npz"), **pack)

    # exhaustive (a, b) pair domain: digests of the reference's own outputs
    i, j = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    pair_img = np.stack([j, j, i], axis=-1)
    pair = {}
    for name in INDEX_TYPES:
        m = R["calculate_index"](pair_img, name)
        pair[f"sha256_{name}"] = np.frombuffer(hashlib.sha256(m.tobytes()).digest(), dtype=np.uint8)
        pair[f"hist_{name}"] = np.histogram(m.flatten(), bins=50, range=(-1, 1))[0].astype(np.int64)
        pair[f"diag_{name}"] = m[np.arange(256), (np.arange(256) * 7 + 3) % 256].copy()
        pair[f"row17_{name}"] = m[17].copy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "pair_domain.npz"), **pair)

    # backend-process.py / process-ndvi.py variants on one frame
    from PIL import Image
    img = synth.vegetation_frame(105, 36, 52)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        wb_pil = Rb["fix_white_balance"](Image.fromarray(img))
    wb = np.array(wb_pil)
    f32 = wb.astype(np.float32)
    var = {"input": img, "backend_wb": wb}
    for name in INDEX_TYPES:
        var[f"backend_{name}"] = Rb["calculate_index"](f32[:, :, 0].copy(), f32[:, :, 1].copy(), f32[:, :, 2].copy(), name)
    raw = img.astype(float)
    ndvi64 = np.clip((raw[:, :, 2] - raw[:, :, 0]) / (raw[:, :, 2] + raw[:, :, 0] + 1e-10), -1, 1)  # process-ndvi.py:18-31
    var["ndvi_f64"] = ndvi64
    st = Rn["analyze_ndvi_statistics"](ndvi64)
    var["ndvi_f64_stats"] = np.array(list(st.values()), dtype=np.float64)
    var["ndvi_f64_hist"] = np.histogram(ndvi64.flatten(), bins=50, range=(-1, 1))[0].astype(np.int64)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "variants.npz"), **var)
    for f in sorted(os.listdir(GOLDEN_DIR)):
        print(f, os.path.getsize(os.path.join(GOLDEN_DIR, f)))


if __name__ == "__main__":
    main()
