"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference functions.

The reference modules cannot be imported here (streamlit / matplotlib / pymongo /
scikit-image are not installed), but its pure-numpy helpers can be
run unmodified: each file is parsed with ``ast``, the wanted ``FunctionDef`` nodes are
compiled one by one and executed into a namespace that only holds ``np``, ``Image``,
``io`` and ``gc``.  No reference source is copied into this repository; the functions are
read from ``/root/reference`` at call time, which exists only in the authoring container.

Used by ``oracle/gen_golden.py`` (to produce ``tests/golden/*.npz``) and by the
``-m "not gpu"`` tests that pin ``oracle/oracle_np.py`` against the real reference when
``/root/reference`` is present.  Nothing under ``lars_image_processing_b200/`` imports it.
"""
from __future__ import annotations

import ast
import gc
import io
import os

import numpy as np

try:  # Pillow is only needed by the file-path variants
    from PIL import Image
except Exception:  # pragma: no cover
    Image = None

REFERENCE_ROOT = os.environ.get("LARS_REFERENCE_ROOT", "/root/reference")

# file -> function names that are pure numpy / PIL
_WANTED = {
    "process-images.py": ("preprocess_large_image", "fix_white_balance", "calculate_index",
                          "analyze_index"),
    "process-ndvi.py": ("calculate_ndvi", "analyze_ndvi_statistics"),
    "process-rgn.py": ("fix_white_balance_rgnir",),
    "backend-process.py": ("fix_white_balance", "calculate_index"),
}


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "process-images.py"))


def load(filename: str) -> dict:
    """Return {name: function} for the pure helpers of one reference file."""
    path = os.path.join(REFERENCE_ROOT, filename)
    with open(path, "r", encoding="utf-8") as fh:
        tree = ast.parse(fh.read(), filename=path)
    ns = {"np": np, "Image": Image, "io": io, "gc": gc}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in _WANTED[filename]:
            mod = ast.Module(body=[node], type_ignores=[])
            exec(compile(mod, path, "exec"), ns)  # noqa: S102 - reference code, read-only
    return {k: ns[k] for k in _WANTED[filename] if k in ns}


_cache: dict = {}


def ref(filename: str, name: str):
    if filename not in _cache:
        _cache[filename] = load(filename)
    return _cache[filename][name]
