"""Build liblars_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.environ.get("LARS_B200_LIB") or os.path.join(PKG_DIR, "liblars_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # exact arithmetic: no fast-math anywhere; host tables must not contract to FMA either
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2,-pthread",
    "-shared",
    "-ldl",         # the TIFF reader binds the system zlib at run time (dlopen), nothing is linked against it
]


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(nvcc):
        raise RuntimeError("nvcc not found; cannot build liblars_b200.so")
    return nvcc


def sources():
    return [os.path.join(CSRC, "lars_capi.cu")]


def _stale() -> bool:
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(PKG_DIR, "..", "include", "lars_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps if os.path.isfile(d))


def build(force: bool = False, verbose: bool = False, defines=(), out_path: str = None) -> str:
    """Compile the CUDA library if missing or older than its sources; return its path.
    ``defines`` / ``out_path`` build tuning variants next to the default library."""
    if out_path is None:
        out_path = LIB_PATH
        if not force and not _stale():
            return LIB_PATH
    cmd = [find_nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-o", out_path, *sources()]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    return out_path


if __name__ == "__main__":
    print(build(force=True, verbose=True))
