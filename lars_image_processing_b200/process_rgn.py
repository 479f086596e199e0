"""Drop-in GPU replacement for ``fix_white_balance_rgnir`` of the reference's ``process-rgn.py``.

The file variant computes the percentile stretch in float64 with an explicit pre-clip to
[p2, p98] (process-rgn.py:25-33) and truncates the float64 value to uint8 (:44) -- unlike
``fix_white_balance`` of process-images.py, which stores the stretch into a float32 array first
(:438) and truncates that.  The two differ by one step wherever the float64 value lands a hair
below an integer (fractional percentiles: small or few-valued frames), so this entry point has
its own table (``LARS_WB_CHAIN_RGN`` of K1b); K1 and K2 are shared.
"""
from __future__ import annotations

import numpy as np

from .engine import get_engine

__all__ = ["fix_white_balance_rgnir"]


def fix_white_balance_rgnir(image_path, save_path=None):
    """process-rgn.py:4-49 -- returns the corrected array only when ``save_path`` is None."""
    from PIL import Image
    img = np.array(Image.open(image_path))                           # :18
    corrected = get_engine().analyze_frame(img, outputs=("wb",), wb_chain="rgn")["wb"][:, :, :3]
    corrected = np.ascontiguousarray(corrected)                      # :41 dstack of 3 channels
    if save_path:                                                    # :46-49
        Image.fromarray(corrected).save(save_path)
        return None
    return corrected
