"""lars_image_processing_b200 -- B200-native RGNir per-pixel analysis path.

White balance -> NDVI / GNDVI / NDWI -> statistics / histogram -> colormap, as hand-written
sm_100a CUDA kernels behind a C ABI (``include/lars_b200.h``), re-exposed under the
reference's own helper names:

    from lars_image_processing_b200 import process_images as pi
    wb = pi.fix_white_balance(img)            # process-images.py:424
    ndvi = pi.calculate_index(wb, "NDVI")     # process-images.py:449
    stats = pi.analyze_index(ndvi, "NDVI")    # process-images.py:492
    everything = pi.analyze_frame(img)        # fused: one trip to the GPU

Importing the package needs neither a GPU nor the compiled library; the first call that
computes anything does, and raises if they are missing (there is no CPU fallback).
"""
__version__ = "0.1.0"

INDEX_TYPES = ("NDVI", "GNDVI", "NDWI")
