"""TEST INFRASTRUCTURE -- CPU restatement of the resize step in front of the analysis path.

Reference: ``preprocess_large_image(img_array, max_dimension=1024)`` process-images.py:398-422:
``Image.fromarray(img).resize((new_w, new_h), Image.Resampling.LANCZOS)``.  The arithmetic lives
in a third-party dependency, **Pillow** (requirements.txt, unpinned upstream; this container has
Pillow 12.2.0): ``src/libImaging/Resample.c`` -- separable two-pass convolution (horizontal pass
first, uint8 intermediate image, then vertical), coefficients computed in double precision from
the windowed sinc (support 3, stretched by the down-scale factor), normalised, rounded to 22-bit
fixed point; each output sample is ``clip8((2^21 + sum(pixel * coef)) >> 22)``.

Pinned: Pillow itself is installed here and on the GPU box, so ``tests/test_resize.py`` compares
this restatement with ``PIL.Image.resize`` directly (bit-exact on every case), and the CUDA
kernels with both.  Only tests, ``smoke()`` and bench CPU legs may import this module.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2          # Resample.c: fixed-point fraction bits for 8-bit channels
LANCZOS_SUPPORT = 3.0


def _sinc(x: float) -> float:
    if x == 0.0:
        return 1.0
    x = x * math.pi
    return math.sin(x) / x


def _lanczos(x: float) -> float:
    if -3.0 <= x < 3.0:
        return _sinc(x) * _sinc(x / 3.0)
    return 0.0


def precompute_coeffs(in_size: int, out_size: int):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the full-image box (0, in_size).
    Returns (ksize, bounds[out_size, 2] int32 (first, count), kk[out_size, ksize] int32)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = LANCZOS_SUPPORT * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_lanczos((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            if v < 0:
                kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS))
            else:
                kk[xx, x] = int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _clip8(acc: np.ndarray) -> np.ndarray:
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def _pass(img: np.ndarray, bounds: np.ndarray, kk: np.ndarray, axis: int) -> np.ndarray:
    """One separable pass along ``axis`` (1 = horizontal, 0 = vertical) of an HxWxC uint8 image."""
    src = img.astype(np.int64)
    out_size = bounds.shape[0]
    shape = list(img.shape)
    shape[axis] = out_size
    out = np.empty(shape, np.uint8)
    for o in range(out_size):
        first, count = int(bounds[o, 0]), int(bounds[o, 1])
        k = kk[o, :count].astype(np.int64)
        if axis == 1:
            acc = (src[:, first:first + count, :] * k[None, :, None]).sum(axis=1)
            out[:, o, :] = _clip8(acc + (1 << (PRECISION_BITS - 1)))
        else:
            acc = (src[first:first + count, :, :] * k[:, None, None]).sum(axis=0)
            out[o, :, :] = _clip8(acc + (1 << (PRECISION_BITS - 1)))
    return out


def resize_lanczos(img: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """``Image.fromarray(img).resize((out_w, out_h), LANCZOS)`` for HxWxC (or HxW) uint8."""
    img = np.asarray(img)
    squeeze = img.ndim == 2
    if squeeze:
        img = img[:, :, None]
    h, w = img.shape[:2]
    cur = img
    if out_w != w:
        _, bh, kh = precompute_coeffs(w, out_w)
    if out_h != h:
        _, bv, kv = precompute_coeffs(h, out_h)
    if out_w != w:
        if out_h != h:
            # Resample.c: the horizontal pass only produces the source rows the vertical pass uses
            first = int(bv[0, 0])
            last = int(bv[-1, 0] + bv[-1, 1])
            cur = cur[first:last]
            bv = bv.copy()
            bv[:, 0] -= first
        cur = _pass(cur, bh, kh, axis=1)
    if out_h != h:
        cur = _pass(cur, bv, kv, axis=0)
    if out_w == w and out_h == h:
        cur = cur.copy()
    return cur[:, :, 0] if squeeze else cur


def target_size(h: int, w: int, max_dimension: int = 1024):
    """process-images.py:404-416 -- returns None when no resize happens, else (new_h, new_w)."""
    if max(h, w) <= max_dimension:
        return None
    if h > w:
        return max_dimension, int(w * (max_dimension / h))
    return int(h * (max_dimension / w)), max_dimension


def preprocess_large_image(img_array, max_dimension: int = 1024):
    """process-images.py:398-422."""
    if img_array is None or img_array.size == 0:
        return None
    h, w = img_array.shape[:2]
    t = target_size(h, w, max_dimension)
    if t is None:
        return img_array
    return resize_lanczos(img_array, t[1], t[0])
