"""In-tree nvcc build of the CUDA extension for sm_100a (no torch types cross the ABI, so plain nvcc)."""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(CSRC, "libsag_b200.so")
SOURCES = ["sag_kernels.cu"]
DEPS = ["sag_core.cuh", "sag_layout.h", os.path.join("..", "..", "include", "sag_b200.h"),
        os.path.join("..", "..", "include", "sag_detmath.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-fmad=false", "-lineinfo", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC"]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found")
    return p


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(os.path.join(CSRC, f)) <= t for f in SOURCES + DEPS)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + \
        [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.check_call(cmd)
    return LIB
