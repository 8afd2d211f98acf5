"""The Python drop-in boundary (zrk_modulation_b200.manager / .messages) mirrors the reference's
Manager / BaseModel / message interface: the call patterns of main.py:151-174 and
UI/PolygonEditor.py:627-669 work unchanged and give the reference's results (golden fixtures).

`-m "not gpu"`: through the emulated build (host logic only).  `-m gpu`: through libzrk_b200.so."""
import numpy as np
import pytest

import parity
from refharness import canonicalise
from zrk_modulation_b200 import (BaseModel, BatchedManager, BatchedTick, MessageType, Timer, compile_batch,
                                 run_simulation_from_config, synthetic_batch)
from zrk_modulation_b200 import messages as M


def _lib(request):
    if request.node.get_closest_marker("gpu"):
        return None
    import emu_lib
    return emu_lib.load()


def _check_against_golden(name, lib):
    g = parity.load_golden(name)
    manager = run_simulation_from_config(g["cfg"], noise=g["noise"], lib=lib)
    dt = g["cfg"]["simulation"]["time_step"]
    msgs, ids, counts = canonicalise(manager.messages, dt)
    assert parity.diff_msgs(g["msgs"], g["ids"], msgs, ids, xyz_rtol=1e-12) == ""
    assert np.array_equal(counts, g["counts"])
    # main.py:171
    assert sum(len(m) for m in manager.messages.values()) == len(g["msgs"]) == manager.total_messages(0)
    f = manager.fault
    assert (f or (0, 0)) == (g["fault_code"], g["fault_tick"])
    return manager, g


CASES = ["shipped_simulation_config", "shipped_config", "mixed_06", "mid_00", "dense_02"]


@pytest.mark.parametrize("name", CASES)
def test_messages_match_reference_emulated(request, name):
    _check_against_golden(name, _lib(request))


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES + ["mixed_04", "tiny_04"])
def test_messages_match_reference_cuda(request, name):
    _check_against_golden(name, _lib(request))


def _ui_replay(manager):
    """The replay loop of UI/PolygonEditor.py:627-669."""
    frames = []
    for t in sorted(manager.messages.keys()):
        frame = []
        for msg in manager.give_messages_by_type(MessageType.DRAW_OBJECTS, step_time=t):
            frame.append((msg.obj_id, tuple(msg.coordinates), msg.is_visible_by_radar, msg.target_type))
        frames.append(frame)
    return frames


def test_ui_call_pattern_emulated(request):
    manager, g = _check_against_golden("shipped_simulation_config", _lib(request))
    frames = _ui_replay(manager)
    assert len(frames) == 200 and sum(len(f) for f in frames) == 42
    first = manager.give_messages_by_type(MessageType.DRAW_OBJECTS, step_time=0)[0]
    assert isinstance(first, M.CPPDrawerObjectsMessage) and first.receiver_id == M.MANAGER_ID
    assert first.target_type is M.TargetType.AIR_PLANE and first.coordinates.shape == (3,)
    # Manager.give_messages_by_id: the launcher's mailbox of tick 2 holds the launch command
    cmd = [m for m in manager.give_messages_by_id(3, step_time=400) if m.type == MessageType.LAUNCH_COMMAND]
    assert len(cmd) == 1 and isinstance(cmd[0], M.CPPLaunchMissileRequestMessage) and cmd[0].target.id == 40
    d = [m for t in manager.messages for m in manager.messages[t] if m.type == MessageType.DESTROYED_MISSILE]
    assert all(isinstance(m.missile_id, tuple) for m in d)  # the reference's 1-tuple (Messages.py:160)
    assert repr(first).startswith("CPPDrawerObjectsMessage: send_time=0, type=MessageType.DRAW_OBJECTS")


def test_manager_api_surface(request):
    lib = _lib(request)
    b = synthetic_batch(3, 6, n_radars=1, n_launchers=1, missiles_per_launcher=2, noise_period=8)
    b.durations[:] = 20 * b.dt_ms
    mgr = BatchedManager(b, log="full", lib=lib)
    assert isinstance(mgr.time, Timer) and mgr.time.get_dt() == 200 and mgr.time.get_time() == 0
    assert mgr.give_messages() == [] and len(mgr.messages) == 0

    class Probe(BaseModel):
        def __init__(self, manager, id):
            super().__init__(manager, id, np.zeros(3))

        def step(self):
            pass

    p = Probe(mgr, 77)
    mgr.add_module(p)
    mgr.add_module(p)
    assert mgr.modules == [p] and mgr.get_module_by_id(77) is p and mgr.get_module_by_id(5) is None
    assert mgr.remove_module(77) and not mgr.remove_module(77)
    mgr.run_simulation(10 * 200)
    assert mgr.time.get_time() == 2000 and sorted(mgr.messages.keys()) == [200 * i for i in range(10)]
    mgr.run_simulation(1900)  # already past: no-op, like `while time < end_time`
    assert mgr.time.get_time() == 2000
    # per-scenario views share the clock
    v = mgr.scenario(2)
    assert v.time is mgr.time and len(v.messages) == 10
    assert v.message_counts()[MessageType.ACTIVE_OBJECTS] == 10
    # host-side add_message lands in the log of that tick, sorted by relevance
    mgr.add_message(M.MissileCountRequestMessage(sender_id=1, receiver_id=2), step_time=400)
    assert mgr.messages[400][0].relevance == 3 and mgr.messages[400][0].send_time == 400
    with pytest.raises(KeyError):
        mgr.messages[123]
    mgr.close()


def test_batched_tick_is_a_basemodel(request):
    """BatchedTick.step() advances the whole batch one tick: usable as a module of a Manager loop."""
    lib = _lib(request)
    g = parity.load_golden("shipped_simulation_config")
    b = compile_batch(g["cfg"], noise=g["noise"])
    batched = BatchedManager(b, log="full", lib=lib)

    class StockLikeManager:  # the loop of Manager.run_simulation (Manager.py:116-140)
        def __init__(self):
            self.time = Timer()
            self.time.set_dt(200)
            self.modules = []

        def run_simulation(self, end_time):
            while self.time.get_time() < end_time:
                for m in self.modules:
                    m.step()
                self.time.update_time()

    host = StockLikeManager()
    tick = BatchedTick(host, 4242, batched)
    assert isinstance(tick, BaseModel) and tick.id == 4242
    host.modules.append(tick)
    host.run_simulation(g["cfg"]["simulation"]["duration"])
    msgs, ids, _ = canonicalise(batched.messages, 200)
    assert parity.diff_msgs(g["msgs"], g["ids"], msgs, ids, xyz_rtol=1e-12) == ""
    batched.close()


def test_events_only_manager_refuses_messages(request):
    b = synthetic_batch(2, 4, n_radars=1, noise_period=4)
    mgr = BatchedManager(b, log="events", lib=_lib(request))
    mgr.step_ticks(3)
    assert mgr.total_messages() > 0
    with pytest.raises(RuntimeError):
        mgr.messages[0]
    mgr.close()


@pytest.mark.reference
def test_message_classes_match_the_reference_signatures():
    """Every message class of modules/Messages.py exists in the mirror with the same constructor
    parameters / defaults and the same __repr__ (build container only)."""
    import inspect
    import refharness as RH
    if not RH.available():
        pytest.skip("reference tree not present")
    RH.ref()
    from modules import Messages as R
    names = [n for n in dir(R) if isinstance(getattr(R, n), type) and issubclass(getattr(R, n), R.BaseMessage)
             and getattr(R, n) is not R.BaseMessage]
    assert len(names) == 17

    def shape(cls):
        return [(p.name, p.default) for p in inspect.signature(cls.__init__).parameters.values()]

    for n in names:
        assert shape(getattr(R, n)) == shape(getattr(M, n)), n
    assert [t.name for t in R.MessageType] == [t.name for t in M.MessageType]
    assert [t.value for t in R.MessageType] == [t.value for t in M.MessageType]
    for cls_name, kw in [("DestroyedMissileId", dict(sender_id=5, missile_id=10, time=200, receiver_id=0, self_detonation=True)),
                         ("MissileCountResponseMessage", dict(sender_id=3, count=4, time=200, receiver_id=0)),
                         ("MissileDetonateMessage", dict(sender_id=11, target_id=None, self_detonation=True)),
                         ("MissileCountRequestMessage", dict(sender_id=0, time=0, receiver_id=3))]:
        a, b = getattr(R, cls_name)(**kw), getattr(M, cls_name)(**kw)
        assert repr(a) == repr(b) and vars(a).keys() == vars(b).keys()
