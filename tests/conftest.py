import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def hostcheck():
    """pixel_math.h compiled for the host (test infrastructure, see tests/hostcheck)."""
    src = os.path.join(ROOT, "tests", "hostcheck", "hostcheck.cpp")
    out_dir = os.path.join(ROOT, "tests", "hostcheck", "_build")
    os.makedirs(out_dir, exist_ok=True)
    lib_path = os.path.join(out_dir, "libhostcheck.so")
    csrc = os.path.join(ROOT, "lars_image_processing_b200", "csrc")
    deps = [src] + [os.path.join(csrc, h) for h in ("pixel_math.h", "lzw_warp.h", "inflate_warp.h", "tiff_host.h")]
    if not os.path.isfile(lib_path) or os.path.getmtime(lib_path) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-pthread", "-shared", "-o", lib_path, src,
                        "-ldl"], check=True)
    return ctypes.CDLL(lib_path)


@pytest.fixture(scope="session")
def golden():
    d = os.path.join(ROOT, "tests", "golden")
    return {name[:-4]: np.load(os.path.join(d, name)) for name in os.listdir(d) if name.endswith(".npz")}


@pytest.fixture(scope="session")
def engine():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from lars_image_processing_b200.engine import get_engine
    return get_engine()
