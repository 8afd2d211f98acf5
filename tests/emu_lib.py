"""Builds and loads tests/emu/_build/libzrk_emu.so: the product sources compiled for the CPU on
the fiber SIMT emulator.  TEST INFRASTRUCTURE ONLY — used by `-m "not gpu"` tests to cover the
host-side logic and the kernel's control flow without a GPU; never used by the package."""
import ctypes as C
import os
import subprocess

from zrk_modulation_b200 import _cabi

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")
_lib = None


def load():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", HERE, "-s"])
        _lib = _cabi.declare(C.CDLL(os.path.join(HERE, "_build", "libzrk_emu.so")))
    return _lib
