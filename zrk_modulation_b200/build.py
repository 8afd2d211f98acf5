"""In-tree nvcc build of libzrk_b200.so for sm_100a (the product has no other backend)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(CSRC, "libzrk_b200.so")
SOURCES = ["zrk_host.cu"]
DEPS = ["zrk_host.cu", "zrk_tick.cuh", "zrk_layout.h", os.path.join("..", "..", "include", "zrk_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-fmad=false",  # no implicit FMA contraction: fp64 results must match the reference's numpy
              "-cudart", "shared",  # link libcudart.so (torch loads the same one) instead of embedding the static runtime
              "-Xcompiler", "-fPIC", "-shared"]


def nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if force or stale():
        cuda_lib = os.path.join(os.path.dirname(os.path.dirname(os.path.realpath(nvcc()))), "lib64")
        # (a client that is not a Python process with torch loaded finds the shared runtime next to the toolkit)
        cmd = [nvcc()] + NVCC_FLAGS + ["-Xlinker", f"-rpath={cuda_lib}"] + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SOURCES
        subprocess.check_call(cmd, cwd=CSRC)
    return OUT


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv))
