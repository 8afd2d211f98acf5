"""Parity of the CUDA path (libzrk_b200.so on a B200) with the oracle and with the reference's
golden fixtures, through the C ABI.  fp64 throughout: messages, payload bits, faults and final
state must be identical (tolerance: none), except against the *reference* fixtures where numpy's
libm pow() in the launch solve allows the last bits of a missile velocity to differ
(rtol 1e-12 on payloads there; decisions and ids exact)."""
import numpy as np
import pytest

import parity
from enginecheck import check_batch_counts, check_batch_full
from oracle import oracle as O
from scenarios import SPECIAL_KINDS, random_config, special_config
from zrk_modulation_b200 import LOG_FULL, compile_batch, make_noise, synthetic_batch
from zrk_modulation_b200.engine import Engine

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", parity.golden_names())
def test_golden_fixture_through_cuda(name):
    g = parity.load_golden(name)
    b = compile_batch(g["cfg"], noise=g["noise"])
    nt = b.n_ticks(0)
    with Engine(b, log_types=LOG_FULL, log_capacity=(3 * b.n_slots + 24) * nt + 64, mask_ticks=nt + 1) as eng:
        eng.step(nt)
        msgs, ids = eng.messages(0, 0, nt)
        fc, ft, _ = eng.status()
        counts = eng.counts()
    assert parity.diff_msgs(g["msgs"], g["ids"], msgs, ids, xyz_rtol=1e-12) == ""
    assert (int(fc[0]) & 63, int(ft[0])) == (g["fault_code"], g["fault_tick"]), g["fault_where"]
    assert np.array_equal(counts[0], g["counts"])
    check_batch_full(b, nt)  # and bit-exact against the oracle


@pytest.mark.parametrize("family", ["mixed", "dense", "mid", "tiny"])
@pytest.mark.parametrize("tpl", [0, 1])
def test_random_scenarios_full_log(family, tpl):
    for seed in range(500, 512):
        cfg = random_config(seed, family)
        b = compile_batch(cfg)
        b.set_noise(make_noise(8, b.max_radars, b.n_slots, seed=seed))
        check_batch_full(b, b.n_ticks(0), ticks_per_launch=tpl)


@pytest.mark.parametrize("kind", SPECIAL_KINDS)
@pytest.mark.parametrize("noise", [False, True])
@pytest.mark.parametrize("exact_only", [False, True])
def test_corner_cases(kind, noise, exact_only):
    """Exact distance ties, near-ties a few ulps apart, coordinates beyond the fp32 pre-filter's
    range, objects exactly on sector edges / at the radar."""
    cfg = special_config(kind)
    b = compile_batch(cfg)
    if noise:
        b.set_noise(make_noise(8, b.max_radars, b.n_slots, seed=5))
    check_batch_full(b, b.n_ticks(0), exact_only=exact_only)


def test_ragged_batch_chunked():
    cfgs = [random_config(600 + i, f) for i, f in enumerate(["mixed", "dense", "mid", "tiny"] * 8)]
    for c in cfgs:
        c["simulation"]["time_step"] = 200
    b = compile_batch(cfgs)
    b.set_noise(make_noise(8, b.max_radars, b.n_slots, seed=11))
    check_batch_full(b, 60, step_chunks=[1, 2, 3, 50, 4])


def test_exact_only_kernel():
    for seed in (700, 701, 702):
        cfg = random_config(seed, "dense")
        b = compile_batch(cfg)
        b.set_noise(make_noise(8, b.max_radars, b.n_slots, seed=seed))
        check_batch_full(b, b.n_ticks(0), exact_only=True)


@pytest.mark.parametrize("n_targets,radars,launchers,mpl,ticks,box", [
    (24, 1, 2, 4, 250, 20.0),    # BASELINE config 5 shape (N = 32 slots, missiles in flight)
    (64, 1, 0, 0, 60, 20.0),     # config 3 shape
    (56, 2, 2, 4, 80, 12.0),     # team of 64 with launches, two radars (one sweeping)
    (120, 2, 1, 8, 40, 20.0),    # team of 128
    (256, 2, 0, 0, 12, 20.0),    # config 4 shape (team of 256)
    (300, 1, 2, 8, 10, 20.0),    # team of 512
    (200, 3, 2, 10, 20, 12.0),   # team of 256, three radars (two radars per barrier + an odd one)
    (70, 4, 3, 6, 30, 6.0),      # team of 128, four radars, dense box
])
def test_synthetic_families_full_log(n_targets, radars, launchers, mpl, ticks, box):
    b = synthetic_batch(6, n_targets, n_radars=radars, n_launchers=launchers, missiles_per_launcher=mpl,
                        seed0=1000, noise_period=16, box_km=box)
    check_batch_full(b, ticks)


def test_config2_shape_1000_objects():
    """BASELINE config 2: one scenario, one radar, 1000 air objects — SURVEY §8d's comparison plan:
    50 ticks with everything visible, then 1000 ticks of a 30-degree sweeping sector, every
    message against the oracle."""
    b = synthetic_batch(1, 1000, n_radars=1, seed0=7, noise_period=8)
    check_batch_full(b, 50)
    b = synthetic_batch(1, 1000, n_radars=1, seed0=8, noise_period=8)
    b.radar_par[0, 0] = [0.0, 0.0, 30000.0, 30.0, 90.0, 30.0, 0.0, 0.0]
    check_batch_full(b, 1000)


def test_config2_full_length_self_consistency():
    """Config 2 at its full 10 000 ticks (no oracle: size-independent properties): one launch per
    256 ticks == one launch per tick-chunk of odd sizes; every tick emits ACTIVE/ALL/FOUND once."""
    from zrk_modulation_b200 import _cabi as A
    b = synthetic_batch(1, 1000, n_radars=1, seed0=9, noise_period=16)
    b.radar_par[0, 0] = [0.0, 0.0, 30000.0, 30.0, 90.0, 30.0, 0.0, 0.0]
    out = []
    for tpl, chunks in ((0, [10000]), (37, [1234, 4321, 4445])):
        with Engine(b, ticks_per_launch=tpl) as eng:
            for c in chunks:
                eng.step(c)
            out.append((eng.counts(), eng.status(), eng.state()))
    (c0, s0, st0), (c1, s1, st1) = out
    assert np.array_equal(c0, c1) and all(np.array_equal(x, y) for x, y in zip(s0, s1))
    assert parity.same_f64(st0["pos"], st1["pos"]) and np.array_equal(st0["ccp_t_upd"], st1["ccp_t_upd"])
    done = 10000 if not (s0[0][0] & 63) else int(s0[1][0]) + 1
    assert c0[0, A.MSG_ACTIVE_OBJECTS] == c0[0, A.MSG_ALL_OBJECTS] == c0[0, A.MSG_FOUND_OBJECTS] == done


def test_events_mode_large_batch():
    """Events-only log at a batch that fills the GPU: counts, faults and sampled event records."""
    b = synthetic_batch(4096, 24, n_radars=1, n_launchers=2, missiles_per_launcher=4, seed0=5000)
    counts = check_batch_counts(b, 250, n_threads=8, sample=32)
    from zrk_modulation_b200 import _cabi as A
    assert counts[:, A.MSG_LAUNCH_COMMAND].sum() > 0 and counts[:, A.MSG_MISSILE_DETONATE].sum() > 0


@pytest.mark.parametrize("n,seed0", [(2048, 10000), (16384, 50000)])
def test_thousands_of_random_scenarios_in_one_ragged_batch(n, seed0):
    """Thousands of random scenarios of all four families (ragged counts, launches, cancels,
    sweeps, both fault kinds) as ONE batch: per-scenario message counts, fault code / tick and
    sampled event records against the oracle."""
    fams = ["mixed", "dense", "mid", "tiny"]
    cfgs = []
    for i in range(n):
        c = random_config(seed0 + i, fams[i % 4])
        c["simulation"] = {"time_step": 200, "duration": 200 * 90}
        cfgs.append(c)
    b = compile_batch(cfgs)
    b.set_noise(make_noise(16, b.max_radars, b.n_slots, seed=77))
    counts = check_batch_counts(b, 90, n_threads=8, sample=48)
    from zrk_modulation_b200 import _cabi as A
    assert counts[:, A.MSG_LAUNCH_SUCCESSFUL].sum() > 500 and counts[:, A.MSG_LAUNCH_CANCELLED].sum() > 0


def test_config3_counts():
    b = synthetic_batch(4096, 64, n_radars=1, seed0=9000)
    check_batch_counts(b, 40, n_threads=8, sample=8)


def test_properties_at_scale():
    """Size-independent properties on a large batch (no oracle): step(a)+step(b) == step(a+b);
    ticks_per_launch does not change results; replicated scenarios give identical results."""
    b = synthetic_batch(8192, 24, n_radars=1, n_launchers=2, missiles_per_launcher=4, seed0=123)
    # make the second half a replica of the first
    for name in ("target_pos", "target_vel"):
        getattr(b, name)[4096:] = getattr(b, name)[:4096]
    res = []
    for tpl, chunks in ((0, [120]), (1, [120]), (16, [50, 70])):
        with Engine(b, ticks_per_launch=tpl) as eng:
            for c in chunks:
                eng.step(c)
            res.append((eng.counts(), eng.status(), eng.state()))
    c0, s0, st0 = res[0]
    for c, s, st in res[1:]:
        assert np.array_equal(c0, c)
        assert all(np.array_equal(x, y) for x, y in zip(s0, s))
        assert parity.same_f64(st0["pos"], st["pos"]) and np.array_equal(st0["ccp_t_key"], st["ccp_t_key"])
    assert np.array_equal(c0[:4096], c0[4096:])
    assert parity.same_f64(st0["pos"][:4096], st0["pos"][4096:])
    # conservation: every detonation removes its missile; launched = detonated + in flight
    from zrk_modulation_b200 import _cabi as A
    assert (c0[:, A.MSG_NEW_MISSILE] == c0[:, A.MSG_LAUNCH_SUCCESSFUL]).all()
    assert (c0[:, A.MSG_MISSILE_DETONATE] <= c0[:, A.MSG_NEW_MISSILE]).all()
    ok = (s0[0] & 63) == 0
    assert (c0[ok, A.MSG_ACTIVE_OBJECTS] == 120).all() and (c0[~ok, A.MSG_ACTIVE_OBJECTS] == s0[1][~ok] + 1).all()


def test_c_abi_from_plain_c(tmp_path):
    """The boundary is a C ABI: a plain C program (tests/c_abi/abi_example.c, no Python, no torch)
    builds simulation_config.yaml by hand, runs it through libzrk_b200.so and must print the
    reference's known answers (no radar noise: 685 messages, detonations at ticks 11 and 13)."""
    import os
    import subprocess
    from zrk_modulation_b200 import _cabi
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "abi_example")
    subprocess.check_call(["gcc", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "c_abi", "abi_example.c"),
                           "-o", exe, "-ldl"])
    out = subprocess.run([exe, _cabi.lib_path()], capture_output=True, text=True, check=True).stdout.splitlines()
    g = parity.load_golden("shipped_simulation_config_nonoise")
    assert [int(x) for x in out[0].split()] == list(g["counts"])
    assert out[1] == "total 685"
    assert out[2:] == ["detonate tick 11 missile 3005 target 40", "detonate tick 13 missile 4005 target 39"]


@pytest.mark.gpu
def test_many_steps_of_rounds_and_waves():
    """Two-kernel stepping over many zrk_step calls: the per-round counters (pending, work queue) are used up and
    reset on the way (1024 slots, 6 per step), and a handle created in scenario waves forks only its first round —
    counts and faults of every scenario against the oracle either way."""
    b = synthetic_batch(512, 24, n_radars=1, n_launchers=2, missiles_per_launcher=4, seed0=77000)
    steps, per = 200, 8
    oc, ofc, oft = O.run_batch(b, 0, b.n_scenarios, steps * per, n_threads=8)
    for waves, wr in ((0, 0), (4, 1), (8, 2)):
        with Engine(b, waves=waves, wave_rounds=wr) as eng:
            for _ in range(steps):
                eng.step(per)
            fc, ft, nr = eng.status()
            counts = eng.counts()
            n_launch, ms = eng.launch_stats()
        assert np.array_equal(fc & 63, ofc) and np.array_equal(ft, oft), f"faults differ (waves={waves}, wave_rounds={wr})"
        assert np.array_equal(counts, oc), f"message counts differ (waves={waves}, wave_rounds={wr})"
        assert n_launch >= 2 * 6 * steps  # (rounds of two launches, not one launch per step)
