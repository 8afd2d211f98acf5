"""The log that scales: event records + per-tick masks + track journal only (ZRK_LOG_EVENTS | ZRK_LOG_MASKS),
with MISSILE_POS / CCP_UPDATE_TARGET / DRAW_OBJECTS rebuilt on the host by zrk_read_messages
(Manager.messages, Manager.py:41-58; CCP.py:237-284, :350-357; Missile.py:183-184).  In this mode the
association certificate, the quiet streak and (warp teams) two-kernel stepping all run, so these tests are
also the message-level parity tests of those paths.  Bit-exact against the reference's golden fixtures
(tolerance: none on decisions / ids / order; rtol 1e-12 on launch-solution payloads, see test_gpu_parity)
and against the oracle (tolerance: none)."""
import numpy as np
import pytest

import parity
from enginecheck import check_batch_full
from oracle import oracle as O
from scenarios import SPECIAL_KINDS, random_config, special_config
from zrk_modulation_b200 import LOG_EVENTS, LOG_MASKS, compile_batch, make_noise, synthetic_batch
from zrk_modulation_b200.engine import Engine


def _lib(gpu):
    if gpu:
        return None
    import emu_lib
    return emu_lib.load()


def _golden(name, lib):
    g = parity.load_golden(name)
    b = compile_batch(g["cfg"], noise=g["noise"])
    nt = b.n_ticks(0)
    with Engine(b, log_types=LOG_EVENTS | LOG_MASKS, log_capacity=64 * b.n_slots + 512, mask_ticks=nt + 1, lib=lib) as eng:
        eng.step(nt)
        msgs, ids = eng.messages(0, 0, nt)
        fc, ft, _ = eng.status()
        counts = eng.counts()
    assert parity.diff_msgs(g["msgs"], g["ids"], msgs, ids, xyz_rtol=1e-12) == ""
    assert (int(fc[0]) & 63, int(ft[0])) == (g["fault_code"], g["fault_tick"]), g["fault_where"]
    assert np.array_equal(counts[0], g["counts"])
    check_batch_full(b, nt, lib=lib, rebuilt=True)  # and bit-exact against the oracle, state included


def _families(lib, seeds):
    for family in ("mixed", "dense", "mid", "tiny"):
        for seed in seeds:
            cfg = random_config(seed, family)
            b = compile_batch(cfg)
            b.set_noise(make_noise(8, b.max_radars, b.n_slots, seed=seed))
            check_batch_full(b, b.n_ticks(0), lib=lib, rebuilt=True)


@pytest.mark.parametrize("name", parity.golden_names())
def test_golden_fixture_rebuilt_log_emulated(name):
    _golden(name, _lib(False))


def test_random_scenarios_rebuilt_log_emulated():
    _families(_lib(False), range(700, 703))


def test_synthetic_families_rebuilt_log_emulated():
    lib = _lib(False)
    check_batch_full(synthetic_batch(6, 24, n_radars=1, n_launchers=2, missiles_per_launcher=4, seed0=900), 250, lib=lib, rebuilt=True)
    check_batch_full(synthetic_batch(6, 24, n_radars=2, n_launchers=2, missiles_per_launcher=4, seed0=910, box_km=5.0), 200, lib=lib,
                     rebuilt=True)
    check_batch_full(synthetic_batch(2, 64, n_radars=2, seed0=920), 100, lib=lib, rebuilt=True)  # a team larger than a warp
    # a window into the run (the dicts are replayed from tick 0 all the same)
    b = synthetic_batch(2, 24, n_radars=1, n_launchers=2, missiles_per_launcher=4, seed0=930)
    with Engine(b, log_types=LOG_EVENTS | LOG_MASKS, mask_ticks=260, lib=lib) as eng:
        eng.step(250)
        whole, wids = eng.messages(1, 0, 250)
        part, pids = eng.messages(1, 100, 140)
    sel = whole[(whole["tick"] >= 100) & (whole["tick"] < 140)].copy()
    # (the id pool of the partial read starts at the window: compare the lists through their own pools)
    assert len(sel) == len(part)
    for x, y in zip(sel, part):
        assert (x["tick"], x["type"], x["sender"], x["receiver"], x["b"]) == (y["tick"], y["type"], y["sender"], y["receiver"], y["b"])
        if x["type"] in parity.LIST_TYPES:
            assert np.array_equal(wids[x["a"]:x["a"] + x["b"]], pids[y["a"]:y["a"] + y["b"]])
        else:
            assert x["a"] == y["a"] and parity.same_f64(x["xyz"], y["xyz"])


def test_events_only_handle_refuses_messages():
    from zrk_modulation_b200.engine import ZrkError
    b = synthetic_batch(1, 4, n_radars=1, noise_period=2)
    with Engine(b, lib=_lib(False)) as eng:
        eng.step(3)
        with pytest.raises(ZrkError):
            eng.messages(0, 0, 3)


# ---------------------------------------------------------------- on the B200

@pytest.mark.gpu
@pytest.mark.parametrize("name", parity.golden_names())
def test_golden_fixture_rebuilt_log_cuda(name):
    _golden(name, None)


@pytest.mark.gpu
def test_random_scenarios_rebuilt_log_cuda():
    _families(None, range(720, 730))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", SPECIAL_KINDS)
def test_corner_cases_rebuilt_log_cuda(kind):
    b = compile_batch(special_config(kind))
    b.set_noise(make_noise(8, b.max_radars, b.n_slots, seed=5))
    check_batch_full(b, b.n_ticks(0), rebuilt=True)


@pytest.mark.gpu
def test_every_message_of_every_scenario_of_a_4096_scenario_batch():
    """What the events-only tests cannot see: WHICH objects each radar saw, every tick, in every scenario.
    The whole rebuilt log of all 4096 scenarios (c5 shape, 120 ticks: launches, missiles in flight,
    detonations) against the oracle's, message for message."""
    B, ticks = 4096, 120
    b = synthetic_batch(B, 24, n_radars=1, n_launchers=2, missiles_per_launcher=4, seed0=40000)
    with Engine(b, log_types=LOG_EVENTS | LOG_MASKS, mask_ticks=ticks + 1) as eng:
        eng.step(ticks)
        fc, ft, _ = eng.status()
        counts = eng.counts()
        cert = eng.certified_ticks()
        for s in range(B):
            r = O.run(b, s, ticks, state=False)
            msgs, ids = eng.messages(s, 0, ticks)
            d = parity.diff_msgs(r["msgs"], r["ids"], msgs, ids)
            assert d == "", f"scenario {s}: {d}"
            assert (r["fault_code"], r["fault_tick"]) == (int(fc[s]) & 63, int(ft[s]))
            assert np.array_equal(counts[s], r["counts"])
    assert cert.sum() > 0.7 * B * ticks  # the certified paths carried the run


@pytest.mark.gpu
def test_large_teams_rebuilt_log_cuda():
    check_batch_full(synthetic_batch(8, 64, n_radars=2, seed0=950), 300, rebuilt=True)
    check_batch_full(synthetic_batch(2, 256, n_radars=2, seed0=960), 60, rebuilt=True)
