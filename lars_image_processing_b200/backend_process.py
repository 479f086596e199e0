"""Drop-in GPU replacements for the helpers of the reference's ``backend-process.py``.

Same signatures: ``fix_white_balance(img)`` is PIL in / PIL out (:17-26),
``calculate_index(red, green, nir, index_type)`` takes separate float32 planes (:28-38),
``process_image`` / ``batch_process`` walk files (:49-97).  The per-index matplotlib figure
(:40-47) is replaced by the full-resolution colormapped image (the per-pixel product).
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

from .engine import INDEX_TYPES, get_engine
from .map_ops import index_from_planes

__all__ = ["fix_white_balance", "calculate_index", "create_index_visualization", "process_image",
           "batch_process", "INPUT_DIR", "OUTPUT_DIR"]

# backend-process.py:8-15 (the reference hard-codes author-local paths; callers set these)
INPUT_DIR = "."
OUTPUT_DIR = "./processed"
PROCESS_WB = False
PROCESS_NDVI = False
PROCESS_GNDVI = False
PROCESS_NDWI = True


def fix_white_balance(img):
    """backend-process.py:17-26 -- PIL image in, white-balanced PIL image out."""
    from PIL import Image
    arr = np.array(img)
    out = get_engine().analyze_frame(arr, outputs=("wb",))["wb"]
    return Image.fromarray(np.ascontiguousarray(out[:, :, :3]) if out.shape[2] == 3 else out)


def calculate_index(red, green, nir, index_type):
    """backend-process.py:28-38 -- normalized difference of two float32 planes, clipped."""
    if index_type == "NDVI":
        return index_from_planes(nir, red)
    if index_type == "GNDVI":
        return index_from_planes(nir, green)
    if index_type == "NDWI":
        return index_from_planes(green, nir)
    # the reference falls through to an unbound local here (:37)
    raise UnboundLocalError("cannot access local variable 'index' where it is not associated with a value")


def create_index_visualization(index_array, index_type, output_path):
    """backend-process.py:40-47 -- colormapped index image written to ``output_path``."""
    from PIL import Image
    from .map_ops import colormap_map
    cmap = "RdYlBu" if index_type == "NDWI" else "RdYlGn"
    Image.fromarray(colormap_map(index_array, cmap, -1.0, 1.0)).save(output_path)


def process_image(image_path, output_dir, process_wb=False, indices=None):
    """backend-process.py:49-73 -- one file: WB (+ save) and the selected index images.

    One fused GPU pass produces the white-balanced frame and every requested colormapped
    index image (the reference recomputes float planes and a figure per index).
    """
    from PIL import Image
    output_dir = Path(output_dir)
    img_name = Path(image_path).stem
    img = np.array(Image.open(image_path))
    requested = list(indices or ())
    wanted = tuple(i for i in requested if i in INDEX_TYPES)
    outputs = ("wb", "rgb") if wanted else ("wb",)
    res = get_engine().analyze_frame(img, outputs=outputs, indices=wanted or INDEX_TYPES)
    if process_wb:
        (output_dir / "white_balanced").mkdir(parents=True, exist_ok=True)
        # an RGBA file stays RGBA with alpha 0, as Image.fromarray of the reference's 4-channel result does (:21-26)
        Image.fromarray(np.ascontiguousarray(res["wb"])).save(output_dir / "white_balanced" / f"{img_name}_wb.tif")
    for index_type in requested:                      # in the caller's order, like the reference's loop (:67-72)
        (output_dir / index_type).mkdir(parents=True, exist_ok=True)
        if index_type not in INDEX_TYPES:
            # the reference's calculate_index falls through to an unbound local (:37) after the directory was made
            # and the indices before this one were written
            raise UnboundLocalError("cannot access local variable 'index' where it is not associated with a value")
        Image.fromarray(res["rgb"][index_type]).save(
            output_dir / index_type / f"{img_name}_{index_type.lower()}.png")


def batch_process():
    """backend-process.py:75-97 -- every image file of INPUT_DIR through process_image."""
    input_path, output_path = Path(INPUT_DIR), Path(OUTPUT_DIR)
    indices = [n for n, on in (("NDVI", PROCESS_NDVI), ("GNDVI", PROCESS_GNDVI), ("NDWI", PROCESS_NDWI)) if on]
    extensions = {".tif", ".tiff", ".png", ".jpg", ".jpeg"}
    image_files = [f for f in input_path.glob("*") if f.suffix.lower() in extensions]
    total = len(image_files)
    for idx, image_file in enumerate(image_files, 1):
        try:
            print(f"Processing {idx}/{total}: {image_file.name}")
            process_image(image_file, output_path, PROCESS_WB, indices if indices else None)
        except Exception as e:  # the reference prints and continues (:96-97)
            print(f"Error processing {image_file.name}: {str(e)}")
