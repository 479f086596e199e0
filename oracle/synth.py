"""TEST INFRASTRUCTURE ONLY -- seeded synthetic RGNir frames.

"Vegetation-like" channels R~N(90,35), G~N(110,35), NIR~N(150,45) clipped to the dtype
range, HWC interleaved.  Plus the adversarial frames the parity tests need: constant
channel, all-zero, two-level, lerp (percentile falls between two distinct values), RGBA,
1x1, odd pixel counts.
"""
from __future__ import annotations

import numpy as np

_MEAN = (90.0, 110.0, 150.0)
_STD = (35.0, 35.0, 45.0)


def vegetation_frame(seed, height, width, dtype=np.uint8, channels=3):
    rng = np.random.default_rng(seed)
    scale = 1.0 if dtype == np.uint8 else 257.0
    top = 255 if dtype == np.uint8 else 65535
    out = np.empty((height, width, channels), dtype=dtype)
    for c in range(channels):
        mu, sd = (_MEAN[c], _STD[c]) if c < 3 else (200.0, 20.0)
        plane = rng.normal(mu * scale, sd * scale, size=(height, width))
        out[:, :, c] = np.clip(np.rint(plane), 0, top).astype(dtype)
    return out


def smooth_frame(seed, height, width):
    """Low-frequency frame with long runs of equal values (worst case for atomics)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float64)
    out = np.empty((height, width, 3), dtype=np.uint8)
    for c in range(3):
        ph = rng.uniform(0, 6.28, size=2)
        plane = _MEAN[c] + _STD[c] * (np.sin(xx / (37.0 + 11 * c) + ph[0]) + np.cos(yy / (53.0 - 7 * c) + ph[1]))
        out[:, :, c] = np.clip(np.rint(plane), 0, 255).astype(np.uint8)
    return out


def adversarial_frames():
    """name -> small uint8 frame exercising one corner of the path each."""
    rng = np.random.default_rng(1234)
    frames = {}
    frames["all_zero"] = np.zeros((8, 8, 3), np.uint8)
    frames["all_255"] = np.full((8, 8, 3), 255, np.uint8)
    const = vegetation_frame(11, 16, 24)
    const[:, :, 1] = 77                                  # p98 == p2 on one channel
    frames["constant_channel"] = const
    two = np.where(rng.random((20, 30, 3)) < 0.5, 10, 200).astype(np.uint8)
    frames["two_level"] = two
    # lerp: n = 101 -> virtual index 2.0 / 98.0 exact; n = 50 -> 0.98 / 48.02 fractional
    lerp = np.zeros((5, 10, 3), np.uint8)
    lerp[..., 0] = np.arange(50, dtype=np.uint8).reshape(5, 10) * 5
    lerp[..., 1] = (np.arange(50, dtype=np.uint8).reshape(5, 10) % 7) * 30
    lerp[..., 2] = 255 - np.arange(50, dtype=np.uint8).reshape(5, 10) * 3
    frames["lerp_n50"] = lerp
    frames["one_pixel"] = np.array([[[3, 200, 90]]], np.uint8)
    frames["two_pixels"] = np.array([[[3, 200, 90], [250, 1, 91]]], np.uint8)
    frames["odd_count"] = vegetation_frame(12, 7, 13)     # 91 px, not a multiple of 4/16
    frames["prime_count"] = vegetation_frame(13, 1, 1021)
    frames["rgba"] = vegetation_frame(14, 9, 17, channels=4)
    frames["tile_plus_one"] = vegetation_frame(15, 1, 2049)
    frames["narrow_range"] = (vegetation_frame(16, 32, 32) // 64 + 100).astype(np.uint8)
    sat = vegetation_frame(17, 32, 48)
    sat[:8] = 255
    sat[-8:] = 0
    frames["saturated_bands"] = sat
    return frames
