"""Kernel time of the device LZW decode on 16 copies of one 12 MP noise frame (Pillow-written LZW TIFF);
run with and without LARS_LZW_VARIANT=2 to compare the two kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from PIL import Image
from lars_image_processing_b200 import ingest
from lars_image_processing_b200.engine import get_engine
rng = np.random.default_rng(0)
img = np.clip(rng.normal(110.0, 40.0, (3000, 4000, 3)), 0, 255).astype(np.uint8)
p = "/dev/shm/lars_lzw_quick.tif" if os.path.isdir("/dev/shm") else "/tmp/lars_lzw_quick.tif"
Image.fromarray(img).save(p, compression="tiff_lzw")
eng, tm = get_engine(), {}
for _ in range(3):
    ingest.decode_tiff_batch_on_device([p] * 16, eng, threads=16, timings=tm)
    print(f"variant {os.environ.get('LARS_LZW_VARIANT', '1')}: 16 x 12 MP, {tm['strips']} strips, kernels {tm['kernel_ms']:.1f} ms "
          f"= {16 * img.size / tm['kernel_ms'] / 1e6:.1f} GB/s decoded", flush=True)
os.remove(p)
