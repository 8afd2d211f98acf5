"""Thin numpy wrapper over the C ABI (include/zrk_b200.h): one Engine = one zrk_handle."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _cabi
from .scenario import Batch


class ZrkError(RuntimeError):
    def __init__(self, rc: int, msg: str):
        super().__init__(f"zrk error {rc}: {msg}")
        self.rc = rc


class Engine:
    """Device state of a batch of scenarios plus the tick loop.

    `lib` defaults to the CUDA library (libzrk_b200.so); there is no CPU fallback.  With `waves` > 1
    the constructor returns while the batch is still being copied (keep the Batch alive and
    unchanged until the first sync / read; an invalid scenario is reported there)."""

    def __init__(self, batch: Batch, log_types: int = _cabi.LOG_EVENTS, log_capacity: int = 0,
                 mask_ticks: int = 0, ticks_per_launch: int = 0, device: int = 0, exact_only: bool = False,
                 waves: int = 0, certificate: bool = True, lib=None, wave_rounds: int = 0):
        self.lib = lib if lib is not None else _cabi.load()
        self.batch = batch
        self.device = device
        self._h = C.c_void_p()
        opt = _cabi.ZrkOptions()
        opt.device = device
        opt.log_types = log_types
        opt.log_capacity = log_capacity
        opt.mask_ticks = mask_ticks
        opt.ticks_per_launch = ticks_per_launch
        opt.flags = (_cabi.OPT_EXACT_ONLY if exact_only else 0) | (0 if certificate else _cabi.OPT_NO_CERTIFICATE)
        opt.wave_rounds = wave_rounds  # rounds of the first step issued wave by wave (see zrk_options.wave_rounds)
        opt.waves = waves  # > 1: upload / init / step in scenario waves on separate streams (see zrk_options.waves)
        self.log_types = log_types
        desc = batch.desc()
        self._desc = desc  # (waves: the arrays it points to are read until the first sync)
        rc = self.lib.zrk_create(C.byref(desc), C.byref(opt), C.byref(self._h))
        if rc != 0:
            msg = self.lib.zrk_last_error(None)
            self._h = C.c_void_p()
            raise ZrkError(rc, msg.decode() if msg else "")

    # -- lifecycle
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.zrk_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != 0:
            msg = self.lib.zrk_last_error(self._h)
            raise ZrkError(rc, msg.decode() if msg else "")

    # -- the tick
    def step(self, n_ticks: int = 1):
        self._ck(self.lib.zrk_step(self._h, int(n_ticks)))

    def sync(self):
        self._ck(self.lib.zrk_sync(self._h))

    def time(self):
        t, ms = C.c_int64(), C.c_int64()
        self._ck(self.lib.zrk_get_time(self._h, C.byref(t), C.byref(ms)))
        return t.value, ms.value

    # -- reads
    def status(self):
        B = self.batch.n_scenarios
        fc, ft, nr = (np.zeros(B, dtype=np.int32) for _ in range(3))
        p = lambda a: a.ctypes.data_as(_cabi._pi32)
        self._ck(self.lib.zrk_read_status(self._h, p(fc), p(ft), p(nr)))
        return fc, ft, nr

    def counts(self) -> np.ndarray:
        B = self.batch.n_scenarios
        c = np.zeros((B, _cabi.ZRK_MSG_NTYPES), dtype=np.int64)
        self._ck(self.lib.zrk_read_counts(self._h, c.ctypes.data_as(C.POINTER(C.c_int64))))
        return c

    def totals(self) -> np.ndarray:
        """Batch totals reduced on the device (int64[NTYPES + 4]): messages per type, then scenarios faulted /
        AttributeError / KeyError / in all — the vector distributed.local_stats builds from counts() and status()."""
        t = np.zeros(_cabi.ZRK_MSG_NTYPES + 4, dtype=np.int64)
        self._ck(self.lib.zrk_read_totals(self._h, t.ctypes.data_as(C.POINTER(C.c_int64))))
        return t

    def alive_slot_ticks(self) -> np.ndarray:
        """Per scenario: alive air objects summed over the ticks run (the total length of the
        active_objects lists AirEnv.step published)."""
        a = np.zeros(self.batch.n_scenarios, dtype=np.int64)
        self._ck(self.lib.zrk_read_activity(self._h, a.ctypes.data_as(C.POINTER(C.c_int64))))
        return a

    def certified_ticks(self) -> np.ndarray:
        """Per scenario: ticks whose CCP step skipped link_object's scan because the association
        certificate proved its outcome (a diagnostic; results do not depend on it)."""
        a = np.zeros(self.batch.n_scenarios, dtype=np.int64)
        self._ck(self.lib.zrk_read_certified(self._h, a.ctypes.data_as(C.POINTER(C.c_int64))))
        return a

    def log_capacity(self) -> int:
        return int(self.lib.zrk_log_capacity(self._h))

    def records(self, s0: int = 0, s1: Optional[int] = None):
        """Explicit records of scenarios [s0, s1): (structured array [n, cap], counts [n])."""
        s1 = self.batch.n_scenarios if s1 is None else s1
        cap = self.log_capacity()
        recs = np.zeros((s1 - s0, cap), dtype=_cabi.MSG_DTYPE)
        n = np.zeros(s1 - s0, dtype=np.int32)
        self._ck(self.lib.zrk_read_records(self._h, s0, s1, recs.ctypes.data_as(C.POINTER(_cabi.ZrkMsg)),
                                           n.ctypes.data_as(_cabi._pi32)))
        return recs, n

    def messages(self, scenario: int, tick0: int = 0, tick1: Optional[int] = None):
        """Full message list of one scenario (needs LOG_FULL): (msgs structured array, id pool)."""
        if tick1 is None:
            tick1 = self.time()[0] + 1
        n, nid = C.c_size_t(), C.c_size_t()
        rc = self.lib.zrk_read_messages(self._h, scenario, tick0, tick1, None, 0, C.byref(n), None, 0, C.byref(nid))
        if rc not in (0, _cabi.ERR_CAPACITY) or (rc == _cabi.ERR_CAPACITY and n.value == 0 and nid.value == 0):
            self._ck(rc)
        msgs = np.zeros(max(1, n.value), dtype=_cabi.MSG_DTYPE)
        ids = np.zeros(max(1, nid.value), dtype=np.int32)
        self._ck(self.lib.zrk_read_messages(self._h, scenario, tick0, tick1,
                                            msgs.ctypes.data_as(C.POINTER(_cabi.ZrkMsg)), msgs.shape[0], C.byref(n),
                                            ids.ctypes.data_as(_cabi._pi32), ids.shape[0], C.byref(nid)))
        return msgs[:n.value], ids[:nid.value]

    def state(self):
        arrays, view = _cabi.alloc_state(self.batch.dims())
        self._ck(self.lib.zrk_read_state(self._h, C.byref(view)))
        return arrays

    # -- checkpoint / resume
    def save_state(self) -> np.ndarray:
        """The complete device state as an opaque uint8 array (zrk_save_state)."""
        n = C.c_size_t()
        self._ck(self.lib.zrk_state_bytes(self._h, C.byref(n)))
        buf = np.empty(n.value, dtype=np.uint8)
        self._ck(self.lib.zrk_save_state(self._h, buf.ctypes.data_as(C.c_void_p), n.value))
        return buf

    def load_state(self, buf: np.ndarray) -> None:
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        self._ck(self.lib.zrk_load_state(self._h, buf.ctypes.data_as(C.c_void_p), buf.nbytes))

    def outcomes(self, s0: int = 0, s1: Optional[int] = None) -> dict:
        """Per-scenario outcome table from the event log and counters (the structured replacement
        of the reference's text log, main.py:21-28): launches, cancelled launches, targets
        destroyed, self-detonations, first launch / last detonation tick, fault."""
        A = _cabi
        s1 = self.batch.n_scenarios if s1 is None else s1
        recs, n = self.records(s0, s1)
        live = np.arange(recs.shape[1])[None, :] < n[:, None]
        ty, tick, flag = recs["type"], recs["tick"].astype(np.int64), recs["flag"]
        det = live & (ty == A.MSG_MISSILE_DETONATE)
        ok = live & (ty == A.MSG_LAUNCH_SUCCESSFUL)
        big = np.iinfo(np.int64).max
        first_launch = np.where(ok, tick, big).min(axis=1)
        last_det = np.where(det, tick, -1).max(axis=1)
        fc, ft, _ = self.status()
        return {"launches": ok.sum(axis=1), "cancelled": (live & (ty == A.MSG_LAUNCH_CANCELLED)).sum(axis=1),
                "targets_destroyed": (det & (flag == 0)).sum(axis=1), "self_detonations": (det & (flag != 0)).sum(axis=1),
                "first_launch_tick": np.where(first_launch == big, -1, first_launch), "last_detonation_tick": last_det,
                "fault_code": fc[s0:s1] & 63, "fault_tick": ft[s0:s1], "messages": self.counts()[s0:s1].sum(axis=1)}

    def launch_stats(self):
        n, ms = C.c_int64(), C.c_double()
        self._ck(self.lib.zrk_launch_stats(self._h, C.byref(n), C.byref(ms)))
        return n.value, ms.value

    def torch_view(self, name: str, dtype=None):
        """Zero-copy torch tensor over a device buffer of the handle ("summary": int32
        [B, 16 + 18] header words + per-type message counters; "dyn", "rec", "masks": raw state).
        The view is borrowed: valid until the engine is closed."""
        import torch
        ptr, nbytes = self.device_ptr(name)
        if name == "summary":
            shape, typestr, td = (self.batch.n_scenarios, 16 + _cabi.ZRK_MSG_NTYPES), "<i4", torch.int32
        else:
            shape, typestr, td = (nbytes,), "|u1", torch.uint8

        class _Buf:
            __cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}

        holder = _Buf()
        t = torch.as_tensor(holder, device=torch.device("cuda", self.device))
        t._zrk_owner = (holder, self)
        return t if dtype is None else t.view(dtype)

    def device_ptr(self, name: str):
        p, sz = C.c_void_p(), C.c_size_t()
        self._ck(self.lib.zrk_device_ptr(self._h, name.encode(), C.byref(p), C.byref(sz)))
        return p.value, sz.value


class ShardedEngine:
    """One Engine per GPU of this process, each owning a contiguous block of the batch's scenarios
    (distributed.shard_range): the single-process counterpart of the one-rank-per-GPU layout.  Scenarios never
    interact (each is its own Manager object graph, main.py:35-149), so the devices exchange nothing: a step is
    issued to all of them at once (one host thread per device: zrk_step blocks while it alternates its kernels)
    and per-scenario reads go to the device that owns the scenario; batch totals are summed on the host."""

    def __init__(self, batch: Batch, devices, threads: bool = True, **kw):
        from concurrent.futures import ThreadPoolExecutor
        from .distributed import shard_range
        self.batch = batch
        self.devices = list(devices)
        if not self.devices:
            raise ValueError("no devices")
        n = len(self.devices)
        self.ranges = [shard_range(batch.n_scenarios, r, n) for r in range(n)]
        # threads=False issues the calls one device after the other (a library that is not thread-safe)
        self._pool = ThreadPoolExecutor(max_workers=n if threads else 1)
        self.log_types = kw.get("log_types", _cabi.LOG_EVENTS)
        self.engines = list(self._pool.map(
            lambda a: Engine(batch.shard(*a[1]), device=a[0], **kw) if a[1][1] > a[1][0] else None,
            zip(self.devices, self.ranges)))

    def _each(self, fn):
        live = [e for e in self.engines if e is not None]
        return list(self._pool.map(fn, live))

    def _owner(self, s: int):
        for e, (s0, s1) in zip(self.engines, self.ranges):
            if s0 <= s < s1:
                return e, s - s0
        raise IndexError(s)

    def close(self):
        self._each(lambda e: e.close())
        self._pool.shutdown(wait=True)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def step(self, n_ticks: int = 1):
        self._each(lambda e: e.step(n_ticks))

    def sync(self):
        self._each(lambda e: e.sync())

    def time(self):
        return next(e for e in self.engines if e is not None).time()

    def _cat(self, parts):
        return np.concatenate(parts, axis=0)

    def status(self):
        parts = self._each(lambda e: e.status())
        return tuple(self._cat([p[i] for p in parts]) for i in range(3))

    def counts(self) -> np.ndarray:
        return self._cat(self._each(lambda e: e.counts()))

    def totals(self) -> np.ndarray:
        return np.sum(self._each(lambda e: e.totals()), axis=0)

    def alive_slot_ticks(self) -> np.ndarray:
        return self._cat(self._each(lambda e: e.alive_slot_ticks()))

    def certified_ticks(self) -> np.ndarray:
        return self._cat(self._each(lambda e: e.certified_ticks()))

    def state(self):
        parts = self._each(lambda e: e.state())
        return {k: self._cat([p[k] for p in parts]) for k in parts[0]}

    def messages(self, scenario: int, tick0: int = 0, tick1: Optional[int] = None):
        e, i = self._owner(scenario)
        return e.messages(i, tick0, tick1)

    def records(self, s0: int = 0, s1: Optional[int] = None):
        s1 = self.batch.n_scenarios if s1 is None else s1
        recs, ns = [], []
        for e, (a, b) in zip(self.engines, self.ranges):
            lo, hi = max(a, s0), min(b, s1)
            if e is not None and hi > lo:
                r, n = e.records(lo - a, hi - a)
                recs.append(r)
                ns.append(n)
        cap = max(r.shape[1] for r in recs)
        recs = [np.pad(r, ((0, 0), (0, cap - r.shape[1]))) for r in recs]
        return self._cat(recs), self._cat(ns)

    def launch_stats(self):
        parts = self._each(lambda e: e.launch_stats())
        return sum(p[0] for p in parts), max(p[1] for p in parts)
