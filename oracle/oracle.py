"""ctypes binding of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  Never imported by zrk_modulation_b200.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from zrk_modulation_b200 import _cabi
from zrk_modulation_b200.scenario import Batch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libzrk_oracle.so")
ZO_NUMPY_POW, ZO_KEEP_MSGS = 1, 2


class ZoRun(C.Structure):
    _fields_ = [("msgs", C.POINTER(_cabi.ZrkMsg)), ("n_msgs", C.c_size_t), ("cap_msgs", C.c_size_t),
                ("ids", C.POINTER(C.c_int32)), ("n_ids", C.c_size_t), ("cap_ids", C.c_size_t),
                ("fault_code", C.c_int32), ("fault_tick", C.c_int32),
                ("counts", C.c_int64 * _cabi.ZRK_MSG_NTYPES)]


_lib = None


def build(force: bool = False) -> str:
    src = [os.path.join(HERE, "zrk_oracle.c"), os.path.join(HERE, "zrk_oracle.h"),
           os.path.join(HERE, "..", "include", "zrk_b200.h")]
    if force or not os.path.exists(LIB) or any(
            os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(LIB) for s in src):
        subprocess.check_call(["make", "-C", HERE, "-B", "-s"])
    return LIB


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(LIB)
        lib.zrk_oracle_run.argtypes = [C.POINTER(_cabi.ZrkBatchDesc), C.c_int, C.c_int, C.c_int,
                                       C.POINTER(ZoRun), C.POINTER(_cabi.ZrkStateView)]
        lib.zrk_oracle_run.restype = C.c_int
        lib.zrk_oracle_free.argtypes = [C.POINTER(ZoRun)]
        lib.zrk_oracle_free.restype = None
        lib.zrk_oracle_run_batch.argtypes = [C.POINTER(_cabi.ZrkBatchDesc), C.c_int, C.c_int, C.c_int,
                                             C.c_int, C.c_int, C.POINTER(C.c_int64),
                                             C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        lib.zrk_oracle_run_batch.restype = C.c_int
        _lib = lib
    return _lib


def run(batch: Batch, scenario: int, n_ticks: int, numpy_pow: bool = False, state: bool = True):
    """Run one scenario; returns dict(msgs=structured array, ids, fault_code, fault_tick, counts,
    state=dict of arrays shaped for a batch of one)."""
    lib = load()
    d = batch.desc()
    r = ZoRun()
    arrays, view = (None, None)
    if state:
        B, NS, R, M, L, Cc = batch.dims()
        arrays, view = _cabi.alloc_state((1, NS, R, M, L, Cc))
    flags = ZO_KEEP_MSGS | (ZO_NUMPY_POW if numpy_pow else 0)
    rc = lib.zrk_oracle_run(C.byref(d), scenario, n_ticks, flags, C.byref(r),
                            C.byref(view) if state else None)
    if rc != 0:
        raise RuntimeError(f"zrk_oracle_run rc={rc}")
    msgs = np.zeros(r.n_msgs, dtype=_cabi.MSG_DTYPE)
    if r.n_msgs:
        C.memmove(msgs.ctypes.data, r.msgs, r.n_msgs * 48)
    ids = np.zeros(r.n_ids, dtype=np.int32)
    if r.n_ids:
        C.memmove(ids.ctypes.data, r.ids, r.n_ids * 4)
    out = dict(msgs=msgs, ids=ids, fault_code=int(r.fault_code), fault_tick=int(r.fault_tick),
               counts=np.array(list(r.counts), dtype=np.int64), state=arrays)
    lib.zrk_oracle_free(C.byref(r))
    return out


def run_batch(batch: Batch, s0: int, s1: int, n_ticks: int, n_threads: int = 1,
              numpy_pow: bool = False):
    """Counters only, threaded over scenarios (the CPU baseline)."""
    lib = load()
    d = batch.desc()
    n = s1 - s0
    counts = np.zeros((n, _cabi.ZRK_MSG_NTYPES), dtype=np.int64)
    fc = np.zeros(n, dtype=np.int32)
    ft = np.zeros(n, dtype=np.int32)
    rc = lib.zrk_oracle_run_batch(C.byref(d), s0, s1, n_ticks, ZO_NUMPY_POW if numpy_pow else 0,
                                  n_threads, counts.ctypes.data_as(C.POINTER(C.c_int64)),
                                  fc.ctypes.data_as(C.POINTER(C.c_int32)),
                                  ft.ctypes.data_as(C.POINTER(C.c_int32)))
    if rc != 0:
        raise RuntimeError(f"zrk_oracle_run_batch rc={rc}")
    return counts, fc, ft
