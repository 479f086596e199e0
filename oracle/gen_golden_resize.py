"""TEST INFRASTRUCTURE -- writes tests/golden/resize.npz: inputs and the outputs of Pillow's own
``Image.resize(..., Image.Resampling.LANCZOS)`` (the third-party arithmetic behind
preprocess_large_image, process-images.py:398-422) for a handful of small geometries, so the parity of
the restatement and of the CUDA kernels is pinned by committed vectors and not only by whatever Pillow
happens to be installed where the tests run.  Generated with Pillow 12.2.0:

    python oracle/gen_golden_resize.py
"""
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402

CASES = [(96, 128, 37, 50), (64, 200, 128, 80), (150, 90, 40, 200), (33, 1028, 20, 257), (257, 64, 64, 64)]


def main():
    out = {}
    rng = np.random.default_rng(31337)
    for i, (ih, iw, oh, ow) in enumerate(CASES):
        img = synth.vegetation_frame(50 + i, ih, iw) if i % 2 == 0 else rng.integers(0, 256, (ih, iw, 3), dtype=np.uint8)
        out[f"in_{i}"] = img
        out[f"out_{i}"] = np.array(Image.fromarray(img).resize((ow, oh), Image.Resampling.LANCZOS))
    big = synth.vegetation_frame(60, 48, 1100)                         # preprocess_large_image itself: a thin strip wider than 1024
    out["pre_in"] = big
    new_w, new_h = 1024, int(48 * (1024 / 1100))
    out["pre_out"] = np.array(Image.fromarray(big).resize((new_w, new_h), Image.Resampling.LANCZOS))
    import PIL
    out["pillow_version"] = np.array(PIL.__version__)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "resize.npz"), **out)
    print("wrote tests/golden/resize.npz with Pillow", PIL.__version__)


if __name__ == "__main__":
    main()
