"""Host-side logic on the CPU: the product sources (zrk_host.cu + zrk_tick.cuh) compiled for the
fiber SIMT emulator in tests/emu, compared with the oracle.  Covers state build, the C ABI reads,
the message-log reconstruction and the per-tick control flow without a GPU.  The emulated build
is test infrastructure only."""
import numpy as np
import pytest

import emu_lib
import parity
from enginecheck import check_batch_counts, check_batch_full
from scenarios import SPECIAL_KINDS, random_config, special_config
from zrk_modulation_b200 import compile_batch, make_noise, synthetic_batch
from zrk_modulation_b200.engine import Engine, ZrkError


@pytest.fixture(scope="module")
def lib():
    return emu_lib.load()


@pytest.mark.parametrize("name", [n for n in parity.golden_names() if n.startswith("shipped") or n.endswith(("0", "4", "6"))])
def test_emulated_tick_on_golden_scenarios(lib, name):
    g = parity.load_golden(name)
    b = compile_batch(g["cfg"], noise=g["noise"])
    check_batch_full(b, b.n_ticks(0), lib=lib)


@pytest.mark.parametrize("family,seed,tpl", [("mixed", 300, 1), ("dense", 301, 5), ("mid", 302, 0), ("tiny", 303, 3),
                                             ("dense", 304, 0), ("mixed", 305, 16)])
def test_emulated_tick_random(lib, family, seed, tpl):
    cfg = random_config(seed, family)
    b = compile_batch(cfg)
    b.set_noise(make_noise(8, b.max_radars, b.n_slots, seed=seed))
    check_batch_full(b, b.n_ticks(0), lib=lib, ticks_per_launch=tpl)


@pytest.mark.parametrize("kind", SPECIAL_KINDS)
@pytest.mark.parametrize("noise", [False, True])
def test_emulated_tick_corner_cases(lib, kind, noise):
    cfg = special_config(kind)
    b = compile_batch(cfg)
    if noise:
        b.set_noise(make_noise(8, b.max_radars, b.n_slots, seed=5))
    check_batch_full(b, b.n_ticks(0), lib=lib)


def test_ragged_batch_and_chunked_steps(lib):
    cfgs = [random_config(400 + i, "mixed") for i in range(5)]
    for c in cfgs:
        c["simulation"]["time_step"] = 200
    b = compile_batch(cfgs)
    b.set_noise(make_noise(8, b.max_radars, b.n_slots, seed=1))
    check_batch_full(b, 40, lib=lib, step_chunks=[1, 2, 30, 7])


def test_ragged_durations_stop_each_scenario_at_its_end(lib):
    cfgs = [random_config(420 + i, "mixed") for i in range(4)]
    for i, c in enumerate(cfgs):
        c["simulation"] = {"time_step": 200, "duration": 200 * (10 + 7 * i)}
    b = compile_batch(cfgs)
    assert list(b.tick_limit) == [10, 17, 24, 31]
    b.set_noise(make_noise(8, b.max_radars, b.n_slots, seed=2))
    check_batch_full(b, 31, lib=lib)
    from zrk_modulation_b200 import BatchedManager, MessageType
    mgr = BatchedManager(b, log="full", lib=lib)
    mgr.run_simulation(31 * 200)
    for i, limit in enumerate([10, 17, 24, 31]):
        v = mgr.scenario(i)
        want = limit if not v.fault else v.fault[1] + 1
        assert len(v.messages) == want and v.message_counts()[MessageType.ACTIVE_OBJECTS] == want
    mgr.close()


def test_exact_only_kernel_agrees(lib):
    cfg = random_config(410, "dense")
    b = compile_batch(cfg)
    b.set_noise(make_noise(8, b.max_radars, b.n_slots, seed=3))
    check_batch_full(b, b.n_ticks(0), lib=lib, exact_only=True)


def test_team_of_64_and_128(lib):
    b = synthetic_batch(2, 40, n_radars=2, n_launchers=2, missiles_per_launcher=3, noise_period=8)
    check_batch_full(b, 30, lib=lib)
    b = synthetic_batch(1, 100, n_radars=1, n_launchers=1, missiles_per_launcher=4, noise_period=8, box_km=8.0)
    check_batch_full(b, 12, lib=lib)


def test_team_of_1024_fits_shared_memory(lib):
    b = synthetic_batch(1, 600, n_radars=1, noise_period=4)
    check_batch_full(b, 3, lib=lib)


def test_events_mode_counts(lib):
    b = synthetic_batch(8, 24, n_radars=1, n_launchers=2, missiles_per_launcher=4, noise_period=8)
    check_batch_counts(b, 40, lib=lib, n_threads=2, sample=4)


def test_create_rejects_bad_input(lib):
    b = synthetic_batch(2, 4, n_radars=1)
    b.target_id[1, 1] = b.target_id[1, 0]
    with pytest.raises(ZrkError) as e:
        Engine(b, lib=lib)
    assert e.value.rc == -5
    b = synthetic_batch(1, 1100, n_radars=1, noise_period=0)
    with pytest.raises(ZrkError) as e:
        Engine(b, lib=lib)
    assert e.value.rc == -5


def test_read_messages_needs_full_log(lib):
    b = synthetic_batch(1, 4, n_radars=1, noise_period=4)
    with Engine(b, lib=lib) as eng:
        eng.step(2)
        with pytest.raises(ZrkError):
            eng.messages(0)


def test_c_abi_example_on_the_emulated_build(lib, tmp_path):
    """tests/c_abi/abi_example.c (plain C against include/zrk_b200.h) against the emulated build."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "abi_example")
    subprocess.check_call(["gcc", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "c_abi", "abi_example.c"),
                           "-o", exe, "-ldl"])
    out = subprocess.run([exe, os.path.join(emu_lib.HERE, "_build", "libzrk_emu.so")], capture_output=True, text=True,
                         check=True).stdout.splitlines()
    g = parity.load_golden("shipped_simulation_config_nonoise")
    assert [int(x) for x in out[0].split()] == list(g["counts"]) and out[1] == "total 685"
    assert out[2:] == ["detonate tick 11 missile 3005 target 40", "detonate tick 13 missile 4005 target 39"]


def test_handles_in_a_loop_reuse_host_buffers(lib):
    """Create / step / read / destroy in a loop with changing batch sizes and wave counts: the
    page-locked read-back buffers are handed from handle to handle; results never change."""
    ref = {}
    for it in range(16):
        B = [3, 130, 260, 17][it % 4]
        b = synthetic_batch(B, 6, n_radars=1, n_launchers=1, missiles_per_launcher=2, noise_period=8, box_km=6.0)
        eng = Engine(b, waves=[0, 2, 3, 0, 5][it % 5], lib=lib)
        eng.step(12)
        got = (eng.counts(), eng.alive_slot_ticks(), eng.status()[0])
        if B in ref:
            assert all(np.array_equal(x, y) for x, y in zip(ref[B], got))
        else:
            ref[B] = got
        if it % 3:
            eng.close()   # the others are left to the garbage collector


@pytest.mark.parametrize("shape", [(2, 200, 3, 2, 10, 20, 12.0), (2, 70, 4, 3, 6, 30, 6.0), (1, 400, 2, 2, 20, 12, 15.0)],
                         ids=["256x3radars", "128x4radars", "512x2radars"])
def test_large_teams_odd_and_even_radar_counts(lib, shape):
    """Teams larger than a warp rank two radars per barrier: odd and even radar counts, launchers
    and missiles in flight, dense boxes (conflicting claims)."""
    B, T, R, L, mpl, ticks, box = shape
    b = synthetic_batch(B, T, n_radars=R, n_launchers=L, missiles_per_launcher=mpl, seed0=7 * T, noise_period=8, box_km=box)
    check_batch_full(b, ticks, lib=lib, ticks_per_launch=3)


def test_many_steps_of_rounds_and_waves(lib):
    """Two-kernel stepping over more zrk_step calls than there are per-round counters (pending, work queue: 1024 slots,
    6 per step), plain and created in scenario waves (only the first round is forked): counts and faults against
    the oracle."""
    from oracle import oracle as O
    b = synthetic_batch(8, 24, n_radars=1, n_launchers=2, missiles_per_launcher=4, seed0=77000)
    steps, per = 175, 8
    oc, ofc, oft = O.run_batch(b, 0, b.n_scenarios, steps * per, n_threads=2)
    for waves, wr in ((0, 0), (2, 1), (4, 2)):
        with Engine(b, waves=waves, wave_rounds=wr, lib=lib) as eng:
            for _ in range(steps):
                eng.step(per)
            fc, ft, nr = eng.status()
            counts = eng.counts()
        assert np.array_equal(fc & 63, ofc) and np.array_equal(ft, oft), f"faults differ (waves={waves}, wave_rounds={wr})"
        assert np.array_equal(counts, oc), f"message counts differ (waves={waves}, wave_rounds={wr})"
