"""Host-side mirror of the reference's drop-in seam: `Timer` (modules/Timer.py), `BaseModel`
(modules/BaseModel.py:7-44) and `Manager` (modules/Manager.py:9-140), driving the CUDA tick for
B scenarios at once.

    manager = run_simulation_from_config("simulation_config.yaml")      # main.py:151-174
    total = sum(len(m) for m in manager.messages.values())              # main.py:171
    manager.give_messages_by_type(MessageType.DRAW_OBJECTS, step_time=t)  # UI/PolygonEditor.py:634

`BatchedManager` keeps the method names and semantics of `Manager` for one selected scenario of
the batch (scenario 0 unless `manager.scenario(i)` is used); `run_simulation(end_time)` runs the
whole batch on the GPU.  `BatchedTick` is a `BaseModel` whose `step()` advances every scenario
one tick, so the batched engine can also be registered as a module of a stock `Manager`.
"""
from __future__ import annotations

from abc import ABCMeta, abstractmethod
from collections.abc import Mapping
from typing import Any, Dict, List, Optional, Sequence, Union

import numpy as np

from . import _cabi
from .engine import Engine
from .messages import (AirObjectView, BaseMessage, MessageType, MESSAGE_TYPES, TARGET_TYPES, _TrajectoryView,
                       materialise)
from .scenario import Batch, compile_batch, load_config


class Timer:
    """modules/Timer.py:5-30 (integer milliseconds)."""

    def __init__(self):
        self.__dt = 1  # SIMULATION_STEP (constants.py:23)
        self.__time = 0

    def set_time(self, time: int) -> None:
        self.__time = time

    def set_dt(self, dt: int) -> None:
        self.__dt = dt

    def get_time(self) -> int:
        return self.__time

    def update_time(self) -> None:
        self.__time += self.__dt

    def get_dt(self) -> float:
        return self.__dt


class BaseModel(metaclass=ABCMeta):
    """modules/BaseModel.py:7-44."""

    @abstractmethod
    def __init__(self, manager, id: int, pos: np.ndarray) -> None:
        self._manager = manager
        self.__id = id
        self.__pos = pos

    @property
    def id(self) -> int:
        return self.__id

    @property
    def pos(self) -> np.ndarray:
        return self.__pos

    @pos.setter
    def pos(self, new_pos: np.ndarray) -> None:
        self.__pos = new_pos

    @abstractmethod
    def step(self) -> None:
        pass


class LazyMessages(Mapping):
    """`Manager.messages`: {step time (ms): [messages]} of one scenario, materialised from the
    device log on access (needs a full log)."""

    def __init__(self, owner: "BatchedManager", scenario: int):
        self._o = owner
        self._s = scenario
        self._cache: Dict[int, List[BaseMessage]] = {}
        self._extra: Dict[int, List[BaseMessage]] = {}

    def _n_ticks(self) -> int:
        return self._o._ticks_with_messages(self._s)

    def __len__(self) -> int:
        return len(self._keys())

    def _keys(self):
        dt = self._o.batch.dt_ms
        ks = [t * dt for t in range(self._n_ticks())]
        return ks + sorted(k for k in self._extra if k not in set(ks))

    def __iter__(self):
        return iter(self._keys())

    def __contains__(self, t) -> bool:
        dt = self._o.batch.dt_ms
        return (isinstance(t, int) and t >= 0 and t % dt == 0 and t // dt < self._n_ticks()) or t in self._extra

    def __getitem__(self, t: int) -> List[BaseMessage]:
        if t not in self:
            raise KeyError(t)
        dt = self._o.batch.dt_ms
        out: List[BaseMessage] = []
        if t % dt == 0 and t // dt < self._n_ticks():
            if t not in self._cache:  # one read of the device log fills every tick
                for k, lst in enumerate(self._o._materialise(self._s, 0, self._n_ticks())):
                    self._cache[k * dt] = lst
            out = self._cache[t]
        if t in self._extra:  # Manager.add_message by host code: appended, then the stable sort by -relevance
            out = sorted(out + self._extra[t], key=lambda x: -x.relevance)
        return out

    def add(self, msg: BaseMessage, t: int) -> None:
        self._extra.setdefault(t, []).append(msg)

    def invalidate(self) -> None:
        self._cache.clear()


class ScenarioView:
    """The `Manager` interface for one scenario of a batch."""

    def __init__(self, owner: "BatchedManager", scenario: int):
        self._owner = owner
        self.index = scenario
        self.time = owner.time
        self.messages = LazyMessages(owner, scenario)
        self.modules = owner.modules

    def add_message(self, msg: BaseMessage, step_time: Optional[int] = None) -> None:
        """Manager.add_message (Manager.py:41-58).  Host-injected messages appear in the log;
        the device tick does not read them."""
        if step_time is None:
            step_time = self.time.get_time()
        if msg.send_time is None:
            msg.send_time = step_time
        self.messages.add(msg, step_time)

    def give_messages(self, step_time: Optional[int] = None) -> List[BaseMessage]:
        if step_time is None:
            step_time = self.time.get_time()
        return self.messages[step_time] if step_time in self.messages else []

    def give_messages_by_id(self, receiver_id: int, step_time: Optional[int] = None) -> List[BaseMessage]:
        return [m for m in self.give_messages(step_time) if m.receiver_id == receiver_id]

    def give_messages_by_type(self, msg_type: MessageType, receiver_id: Optional[int] = None,
                              step_time: Optional[int] = None) -> List[BaseMessage]:
        return [m for m in self.give_messages(step_time)
                if m.type == msg_type and (receiver_id is None or m.receiver_id == receiver_id)]

    @property
    def fault(self):
        """(code, tick) of the exception the reference would have raised in this scenario, or None."""
        return self._owner.fault_of(self.index)

    def message_counts(self) -> Dict[MessageType, int]:
        c = self._owner.counts()[self.index]
        return {t: int(c[i]) for i, t in enumerate(MESSAGE_TYPES)}


class BatchedManager(ScenarioView):
    """Drop-in for `Manager` (Manager.py:9-140) over a batch of scenarios on one GPU."""

    def __init__(self, batch: Batch, log: str = "auto", device: int = 0, ticks_per_launch: int = 0,
                 max_ticks: Optional[int] = None, lib=None, devices=None):
        self.batch = batch
        self.time = Timer()
        self.time.set_dt(batch.dt_ms)
        self.modules: List = []
        # log: "rebuilt" (what "auto" picks) keeps event records + per-tick masks + the track journal — 8 + 4 R bytes
        # per scenario-tick, any batch size — and Manager.messages is rebuilt from them on access, message for
        # message what the reference stores (Manager.py:41-58); "full" stores every message as a record (small
        # batches only); "events" keeps no masks: counts, status and event records, no Manager.messages.
        if log == "auto":
            log = "rebuilt"
        if max_ticks is None:
            max_ticks = max(1, int(max(batch.n_ticks(s) for s in range(batch.n_scenarios))))
        self.max_ticks = max_ticks
        self.log = log
        # devices: None = the one GPU `device`; a list of ordinals or "all" = this process drives every one of them,
        # each owning a contiguous block of scenarios (ShardedEngine) — main.run_simulation_from_config is one
        # process too (main.py:151-174); the one-rank-per-GPU layout of bench.py needs no change here
        if devices == "all":
            import torch
            devices = list(range(torch.cuda.device_count()))
        if log == "full":
            cap = (3 * batch.n_slots + 16 + 4 * batch.max_radars) * max_ticks + 64
            kw = dict(log_types=_cabi.LOG_FULL, log_capacity=cap, mask_ticks=max_ticks + 1)
        elif log == "rebuilt":
            kw = dict(log_types=_cabi.LOG_EVENTS | _cabi.LOG_MASKS, mask_ticks=max_ticks + 1)
        else:
            kw = dict(log_types=_cabi.LOG_EVENTS)
        if devices is not None and len(devices) > 1:
            from .engine import ShardedEngine
            self.engine = ShardedEngine(batch, devices, threads=lib is None, ticks_per_launch=ticks_per_launch, lib=lib, **kw)
        else:
            self.engine = Engine(batch, ticks_per_launch=ticks_per_launch, device=devices[0] if devices else device, lib=lib, **kw)
        self._status = None
        self._counts = None
        self._objects: Dict[int, Dict[int, AirObjectView]] = {}
        self._views: Dict[int, ScenarioView] = {}
        ScenarioView.__init__(self, self, 0)

    # ---- Manager API (Manager.py:16-39)
    def add_module(self, module) -> None:
        if module not in self.modules:
            self.modules.append(module)

    def remove_module(self, module_id: int) -> bool:
        for i, module in enumerate(self.modules):
            if hasattr(module, 'id') and module.id == module_id:
                self.modules.pop(i)
                return True
        return False

    def get_module_by_id(self, module_id: int):
        for module in self.modules:
            if hasattr(module, 'id') and module.id == module_id:
                return module
        return None

    def run_simulation(self, end_time: int) -> None:
        """Manager.run_simulation (Manager.py:111-140): `while time < end_time`, on the GPU."""
        now, dt = self.time.get_time(), self.batch.dt_ms
        if end_time > now:
            self.step_ticks(-(-(end_time - now) // dt))

    # ---- batch API
    def step_ticks(self, n: int) -> None:
        self.engine.step(n)
        self.time.set_time(self.time.get_time() + n * self.batch.dt_ms)
        self._status = self._counts = None
        self._objects.clear()
        self.messages.invalidate()
        for v in self._views.values():
            v.messages.invalidate()

    def scenario(self, i: int) -> ScenarioView:
        if not 0 <= i < self.batch.n_scenarios:
            raise IndexError(i)
        if i == 0:
            return self
        if i not in self._views:
            self._views[i] = ScenarioView(self, i)
        return self._views[i]

    def status(self):
        if self._status is None:
            self._status = self.engine.status()
        return self._status

    def counts(self) -> np.ndarray:
        if self._counts is None:
            self._counts = self.engine.counts()
        return self._counts

    def fault_of(self, s: int):
        fc, ft, _ = self.status()
        code = int(fc[s]) & 63
        return (code, int(ft[s])) if code else None

    def total_messages(self, s: Optional[int] = None) -> int:
        c = self.counts()
        return int(c.sum() if s is None else c[s].sum())

    def close(self) -> None:
        self.engine.close()

    # ---- log materialisation
    def _ticks_with_messages(self, s: int) -> int:
        fc, ft, _ = self.status()
        done = self.time.get_time() // self.batch.dt_ms
        if int(fc[s]) & 63:  # the reference aborted in tick ft: its messages exist, nothing later
            return int(ft[s]) + 1
        if self.batch.tick_limit is not None:  # this scenario's own `while time < end_time`
            done = min(done, int(self.batch.tick_limit[s]))
        return int(done)

    def _object_views(self, s: int) -> Dict[int, AirObjectView]:
        if s in self._objects:
            return self._objects[s]
        b = self.batch
        st = self.engine.state()
        T = b.max_targets
        views: Dict[int, AirObjectView] = {}
        for i in range(int(b.n_targets[s])):
            views[int(b.target_id[s, i])] = AirObjectView(
                int(b.target_id[s, i]), False, TARGET_TYPES[int(b.target_type[s, i])], st["pos"][s, i].copy(),
                _TrajectoryView(b.target_vel[s, i], b.target_pos[s, i], 0.0))
        for m in range(int(b.n_missiles[s])):
            views[int(b.missile_id[s, m])] = AirObjectView(
                int(b.missile_id[s, m]), True, None, st["pos"][s, T + m].copy(),
                _TrajectoryView(st["missile_vel"][s, m], b.launcher_pos[s, int(b.missile_launcher[s, m])],
                                float(st["missile_t0"][s, m])))
        self._objects[s] = views
        return views

    def _materialise(self, s: int, tick0: int, tick1: int) -> List[List[BaseMessage]]:
        if self.log == "events":
            raise RuntimeError("this BatchedManager keeps only the event log; create it with log='rebuilt' (the default) "
                               "or log='full' to materialise Manager.messages")
        msgs, ids = self.engine.messages(s, tick0, tick1)
        objs = self._object_views(s)
        dt = self.batch.dt_ms
        out: List[List[BaseMessage]] = [[] for _ in range(tick1 - tick0)]
        for r in msgs:
            out[int(r["tick"]) - tick0].append(materialise(r, ids, int(r["tick"]) * dt, objs))
        return out


class BatchedTick(BaseModel):
    """A `BaseModel` whose `step()` advances every scenario of a `BatchedManager` by one tick.
    Registered in a stock reference `Manager` it is stepped once per tick (after the priority
    modules, Manager.py:123-127), which keeps the batch in lockstep with that manager's clock."""

    def __init__(self, manager, id: int, batched: BatchedManager, pos: Optional[np.ndarray] = None):
        super().__init__(manager, id, np.zeros(3) if pos is None else pos)
        self.batched = batched

    def step(self) -> None:
        self.batched.step_ticks(1)


def create_objects_from_config(config: Union[Dict[str, Any], Sequence[Dict[str, Any]]], noise="auto",
                               noise_seed: int = 12345, **kw):
    """main.create_objects_from_config (main.py:35-149) for one config or a list of configs:
    returns (BatchedManager, {id: description}) — the second item lists the scenario-0 module ids
    the reference would have put in `objects_by_id`.

    `noise` replaces the reference's unseeded np.random.normal(0, 5) radar noise (Radar.py:138-142):
    "auto" draws a shared N(0, 5^2) table [64, R, slots, 3] from `noise_seed`, an array is used as
    given ([W,R,slots,3] or [B,W,R,slots,3]), None disables the noise."""
    batch = compile_batch(config)
    if isinstance(noise, str) and noise == "auto":
        from .scenario import make_noise
        noise = make_noise(64, batch.max_radars, batch.n_slots, seed=noise_seed)
    batch.set_noise(noise)
    mgr = BatchedManager(batch, **kw)
    ids: Dict[int, str] = {int(batch.airenv_id[0]): "AirEnv"}
    for r in range(int(batch.n_radars[0])):
        ids[int(batch.radar_id[0, r])] = "SectorRadar"
    for l in range(int(batch.n_launchers[0])):
        ids[int(batch.launcher_id[0, l])] = "MissileLauncher"
    if int(batch.ccp_id[0]) != _cabi.ZRK_NONE:
        ids[int(batch.ccp_id[0])] = "CombatControlPoint"
    return mgr, ids


def run_simulation_from_config(config_path: Union[str, Dict[str, Any]], noise="auto", **kw) -> BatchedManager:
    """main.run_simulation_from_config (main.py:151-174)."""
    config = load_config(config_path) if isinstance(config_path, str) else config_path
    manager, _ = create_objects_from_config(config, noise=noise, **kw)
    manager.run_simulation(config['simulation']['duration'])
    return manager
