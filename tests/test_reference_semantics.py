"""Behavioural facts of the reference that a builder cannot guess from the class names
(SURVEY §5.9), checked one by one on hand-made scenarios — in the spirit of the reference's own
tests/ml_test.py (launcher behaviour), but against real runs instead of mocks.  Each case runs
the oracle and the tick kernel (emulated build here; `-m gpu` repeats it on the B200)."""
import numpy as np
import pytest

from oracle import oracle as O
from zrk_modulation_b200 import LOG_FULL, compile_batch
from zrk_modulation_b200 import _cabi as A
from zrk_modulation_b200.engine import Engine


def base_cfg(ticks=40, dt=200):
    return {
        "simulation": {"time_step": dt, "duration": dt * ticks},
        "air_environment": {"id": 1, "position": [0.0, 0.0, 0.0], "targets": [
            {"id": 100, "type": "AIR_PLANE", "position": [5000.0, 5000.0, 2000.0], "velocity": [100.0, 0.0, 0.0]}]},
        "combat_control_point": {"id": 0, "missile_launcher_ids": [3], "radar_ids": [5]},
        "missile_launchers": [{"id": 3, "position": [500.0, 500.0, 0.0], "max_missiles": 5, "missiles": [
            {"id": 10, "velocity": 600, "life_time": 60, "explosion_radius": 200},
            {"id": 11, "velocity": 600, "life_time": 60, "explosion_radius": 200}]}],
        "radars": [{"id": 5, "position": [0.0, 0.0, 20.0], "azimuth_start": 0.0, "elevation_start": 0.0,
                    "max_distance": 100000.0, "azimuth_range": 360.0, "elevation_range": 180.0,
                    "azimuth_speed": 0.0, "elevation_speed": 0.0, "scan_mode": "horizontal"}]}


def run_both(cfg, request, noise_seed=None):
    """Messages of the scenario from the oracle and from the tick kernel (must be identical)."""
    import parity
    lib = None
    if not request.node.get_closest_marker("gpu"):
        import emu_lib
        lib = emu_lib.load()
    b = compile_batch(cfg)
    if noise_seed is not None:
        from zrk_modulation_b200 import make_noise
        b.set_noise(make_noise(8, b.max_radars, b.n_slots, seed=noise_seed))
    nt = b.n_ticks(0)
    r = O.run(b, 0, nt)
    import refharness as RH
    if RH.available():  # build container: the unmodified reference agrees with the oracle on this case too
        ref = RH.run_reference(cfg, b.noise)
        rp = O.run(b, 0, nt, numpy_pow=True)
        assert parity.diff_msgs(ref["msgs"], ref["ids"], rp["msgs"], rp["ids"]) == ""
    with Engine(b, log_types=LOG_FULL, log_capacity=(3 * b.n_slots + 24) * nt + 64, mask_ticks=nt + 1, lib=lib) as eng:
        eng.step(nt)
        msgs, ids = eng.messages(0, 0, nt)
        st = eng.state()
    assert parity.diff_msgs(r["msgs"], r["ids"], msgs, ids) == ""
    return msgs, ids, st


def of_type(msgs, t):
    return msgs[msgs["type"] == t]


CASES = []


def case(f):
    CASES.append(f)
    return f


@case
def launch_pipeline_latencies(request):
    """tick 0 detect + count requests (relevance 3, sorted first), tick 1 answers +
    LAUNCH_COMMAND, tick 2 LAUNCH_MISSILE + LAUNCH_SUCCESSFUL, tick 3 LAUNCHED + NEW_MISSILE,
    tick 4 first MISSILE_POS."""
    msgs, _, _ = run_both(base_cfg(), request)
    assert msgs[0]["type"] == A.MSG_MISSILE_COUNT_REQUEST and msgs[0]["relevance"] == 3 and msgs[0]["tick"] == 0
    first = lambda t: int(of_type(msgs, t)["tick"][0])
    assert first(A.MSG_MISSILE_COUNT_RESPONSE) == 1 and first(A.MSG_LAUNCH_COMMAND) == 1
    assert first(A.MSG_LAUNCH_MISSILE) == 2 and first(A.MSG_LAUNCH_SUCCESSFUL) == 2
    assert first(A.MSG_LAUNCHED_MISSILE) == 3 and first(A.MSG_NEW_MISSILE) == 3
    assert first(A.MSG_MISSILE_POS) == 4


@case
def launcher_pops_the_last_missile(request):
    """MissileLauncher.py:70: with missiles [10, 11], 11 fires first."""
    msgs, _, _ = run_both(base_cfg(), request)
    assert int(of_type(msgs, A.MSG_LAUNCH_MISSILE)["receiver"][0]) == 11
    assert len(of_type(msgs, A.MSG_LAUNCH_COMMAND)) == 1  # one target, followed after the first command


@case
def max_missiles_drops_the_surplus(request):
    """MissileLauncher.add_missile silently drops missiles beyond max_missiles (:43-48)."""
    cfg = base_cfg()
    cfg["missile_launchers"][0]["max_missiles"] = 1
    msgs, _, _ = run_both(cfg, request)
    assert int(of_type(msgs, A.MSG_MISSILE_COUNT_RESPONSE)["b"][0]) == 1
    assert int(of_type(msgs, A.MSG_LAUNCH_MISSILE)["receiver"][0]) == 10


@case
def too_far_target_cancels_and_missile_returns(request):
    """Missile.py:92-94 -> LAUNCH_CANCELLED; the launcher puts the missile back (:125-128); the CCP
    has already counted the launch and never retries."""
    cfg = base_cfg()
    cfg["missile_launchers"][0]["missiles"] = [{"id": 10, "velocity": 300, "life_time": 3, "explosion_radius": 50}]
    msgs, _, st = run_both(cfg, request)
    c = of_type(msgs, A.MSG_LAUNCH_CANCELLED)
    assert len(c) == 1 and int(c["b"][0]) == A_CANCEL_TOO_FAR and len(of_type(msgs, A.MSG_LAUNCH_SUCCESSFUL)) == 0
    assert st["launcher_count"][0, 0] == 1 and st["ccp_launched"][0, 0] == 1
    assert st["ccp_t_following"][0, 0] == 1 and st["ccp_t_missile"][0, 0] == A.ZRK_NONE
    assert len(of_type(msgs, A.MSG_LAUNCH_COMMAND)) == 1


A_CANCEL_TOO_FAR = 4


@case
def stationary_target_is_a_new_target_every_tick(request):
    """velocity 0 -> AirObject.velocity is NaN (AirObject.py:35): ranges collapse to [0, 0], so with
    radar noise the target is NEW_TARGET on every tick, and every launch at it is cancelled
    (SURVEY §5.9#6)."""
    cfg = base_cfg(ticks=12)
    cfg["air_environment"]["targets"][0]["velocity"] = [0.0, 0.0, 0.0]
    msgs, _, st = run_both(cfg, request, noise_seed=1)
    assert len(of_type(msgs, A.MSG_LAUNCH_COMMAND)) == 2       # once per capacity unit
    c = of_type(msgs, A.MSG_LAUNCH_CANCELLED)
    assert len(c) == 2 and set(c["b"]) == {3}                   # "Interception times are not positive"
    assert st["ccp_n_targets"][0] == 1                          # same key overwritten, never a second entry


@case
def self_detonation_timer_rounding(request):
    """life_time 35 s at dt 200 ms expires after 176 decrements, not 175."""
    cfg = base_cfg(ticks=200)
    # a 1 m fuse: the missile flies past its target between two ticks and runs out its timer
    cfg["missile_launchers"][0]["missiles"] = [{"id": 10, "velocity": 600, "life_time": 35, "explosion_radius": 1}]
    msgs, _, _ = run_both(cfg, request)
    det = of_type(msgs, A.MSG_MISSILE_DETONATE)
    pos = of_type(msgs, A.MSG_MISSILE_POS)
    assert len(det) == 1 and det["flag"][0] == 1 and det["b"][0] == A.ZRK_NONE
    assert int(det["tick"][0]) - int(pos["tick"][0]) + 1 == 176


@case
def destroyed_target_stays_a_ghost_track(request):
    """DestroyedMissileId.missile_id is a 1-tuple, so the CCP never deletes tracks:
    after the hit the target and missile entries are still there."""
    msgs, _, st = run_both(base_cfg(ticks=120), request)
    det = of_type(msgs, A.MSG_MISSILE_DETONATE)
    assert len(det) == 1 and det["flag"][0] == 0 and det["b"][0] == 100
    k = int(det["tick"][0])
    assert list(of_type(msgs, A.MSG_DESTROYED_MISSILE)["tick"]) == [k + 1]
    assert st["ccp_n_targets"][0] == 1 and st["ccp_n_missiles"][0] == 1
    act = of_type(msgs, A.MSG_ACTIVE_OBJECTS)
    assert act["b"][act["tick"] == k][0] == 2 and act["b"][act["tick"] == k + 1][0] == 0


@case
def sector_sweep_alternates(request):
    """simulation_config.yaml's radar (range 180, speed 180): the sector alternates 0, 180, 0 ...
    so detections happen on even ticks only."""
    cfg = base_cfg(ticks=8)
    cfg["radars"][0].update(azimuth_range=180.0, azimuth_speed=180.0, elevation_range=90.0, max_distance=20000.0)
    cfg["missile_launchers"] = []
    cfg["combat_control_point"]["missile_launcher_ids"] = []
    msgs, _, _ = run_both(cfg, request)
    found = of_type(msgs, A.MSG_FOUND_OBJECTS)
    assert list(found["b"]) == [1, 0, 1, 0, 1, 0, 1, 0]


@case
def no_ccp_no_radar_still_ticks(request):
    """Without combat_control_point / radars the loop still runs AirEnv (main.py:104-105)."""
    cfg = base_cfg(ticks=5)
    del cfg["combat_control_point"]
    cfg["radars"] = []
    msgs, _, _ = run_both(cfg, request)
    assert list(msgs["type"]) == [A.MSG_ACTIVE_OBJECTS] * 5


@pytest.mark.parametrize("fn", CASES, ids=lambda f: f.__name__)
def test_semantics_emulated(request, fn):
    fn(request)


@pytest.mark.gpu
@pytest.mark.parametrize("fn", CASES, ids=lambda f: f.__name__)
def test_semantics_cuda(request, fn):
    fn(request)
