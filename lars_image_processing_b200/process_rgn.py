"""Drop-in GPU replacement for ``fix_white_balance_rgnir`` of the reference's ``process-rgn.py``.

The file variant computes the same percentile stretch in float64 with an explicit pre-clip to
[p2, p98] (process-rgn.py:25-33); the uint8 result is identical to ``fix_white_balance``
(SURVEY.md section 8(a) row a9), so the same K1 / K1b / K2 kernels serve it.
"""
from __future__ import annotations

import numpy as np

from .engine import get_engine

__all__ = ["fix_white_balance_rgnir"]


def fix_white_balance_rgnir(image_path, save_path=None):
    """process-rgn.py:4-49 -- returns the corrected array only when ``save_path`` is None."""
    from PIL import Image
    img = np.array(Image.open(image_path))                           # :18
    corrected = get_engine().analyze_frame(img, outputs=("wb",))["wb"][:, :, :3]
    corrected = np.ascontiguousarray(corrected)                      # :41 dstack of 3 channels
    if save_path:                                                    # :46-49
        Image.fromarray(corrected).save(save_path)
        return None
    return corrected
