"""SURVEY §8(f) "next" rows built on top of the path: checkpoint / resume of the device state,
the per-scenario outcome table, and the Monte-Carlo perturbation family.
`-m "not gpu"` through the emulated build, `-m gpu` through libzrk_b200.so."""
import numpy as np
import pytest

import parity
from zrk_modulation_b200 import LOG_FULL, compile_batch, perturbed_batch, synthetic_batch
from zrk_modulation_b200 import _cabi as A
from zrk_modulation_b200.engine import Engine, ZrkError


def _lib(request):
    if request.node.get_closest_marker("gpu"):
        return None
    import emu_lib
    return emu_lib.load()


def _checkpoint_roundtrip(lib):
    b = synthetic_batch(6, 12, n_radars=1, n_launchers=2, missiles_per_launcher=2, noise_period=8)
    kw = dict(log_types=LOG_FULL, log_capacity=4000, mask_ticks=64, lib=lib)
    with Engine(b, **kw) as ref:
        ref.step(40)
        want = (ref.counts(), ref.state(), [ref.messages(s, 0, 40) for s in range(6)])
    with Engine(b, **kw) as e1:
        e1.step(17)
        blob = e1.save_state()
    with Engine(b, **kw) as e2:
        e2.load_state(blob)
        assert e2.time()[0] == 17
        e2.step(23)
        got = (e2.counts(), e2.state(), [e2.messages(s, 0, 40) for s in range(6)])
        with pytest.raises(ZrkError):
            e2.load_state(blob[:100])
    assert np.array_equal(want[0], got[0])
    for k in want[1]:
        assert parity.same_f64(want[1][k], got[1][k]) if want[1][k].dtype == np.float64 else np.array_equal(want[1][k], got[1][k])
    for (m0, i0), (m1, i1) in zip(want[2], got[2]):
        assert parity.diff_msgs(m0, i0, m1, i1) == ""
    b2 = synthetic_batch(5, 12, n_radars=1, n_launchers=2, missiles_per_launcher=2, noise_period=8)
    with Engine(b2, **kw) as e3, pytest.raises(ZrkError):
        e3.load_state(blob)


def _outcomes(lib):
    g = parity.load_golden("shipped_simulation_config")
    b = compile_batch(g["cfg"], noise=g["noise"])
    with Engine(b, lib=lib) as eng:
        eng.step(b.n_ticks(0))
        o = eng.outcomes()
    assert o["launches"][0] == 2 and o["targets_destroyed"][0] == 2 and o["self_detonations"][0] == 0
    assert o["first_launch_tick"][0] == 3 and o["last_detonation_tick"][0] == 13 and o["fault_code"][0] == 0
    assert o["messages"][0] == 685


def _perturbed(lib):
    g = parity.load_golden("shipped_simulation_config")
    b = perturbed_batch(g["cfg"], 16, pos_sigma=50.0, vel_sigma=2.0, seed=3, noise=g["noise"])
    assert b.n_scenarios == 16 and np.array_equal(b.target_pos[0], compile_batch(g["cfg"]).target_pos[0])
    assert not np.array_equal(b.target_pos[1], b.target_pos[0])
    with Engine(b, lib=lib) as eng:
        eng.step(200)
        c = eng.counts()
    assert np.array_equal(c[0], g["counts"])          # scenario 0 is the original
    assert (c[:, A.MSG_ACTIVE_OBJECTS] == 200).all()


def test_checkpoint_roundtrip_emulated(request):
    _checkpoint_roundtrip(_lib(request))


def test_outcomes_emulated(request):
    _outcomes(_lib(request))


def test_perturbed_batch_emulated(request):
    _perturbed(_lib(request))


@pytest.mark.gpu
def test_checkpoint_roundtrip_cuda(request):
    _checkpoint_roundtrip(_lib(request))


@pytest.mark.gpu
def test_outcomes_and_perturbed_cuda(request):
    _outcomes(_lib(request))
    _perturbed(_lib(request))


@pytest.mark.gpu
def test_zero_copy_torch_view_of_the_summary(request):
    """PyTorch as plumbing: the per-scenario counters as a CUDA tensor without a host hop
    (what a NCCL statistics all-reduce consumes)."""
    import torch
    b = synthetic_batch(64, 12, n_radars=1, n_launchers=2, missiles_per_launcher=2, noise_period=8)
    with Engine(b) as eng:
        eng.step(50)
        t = eng.torch_view("summary")
        assert t.is_cuda and tuple(t.shape) == (64, 34) and t.dtype == torch.int32
        assert np.array_equal(t[:, 16:].cpu().numpy().astype(np.int64), eng.counts())
        assert int(t[:, 0].min()) <= 50  # H_TICK
        raw = eng.torch_view("dyn")
        assert raw.dtype == torch.uint8 and raw.numel() == eng.device_ptr("dyn")[1]


def test_checkpoint_of_another_batch_is_refused(request):
    """zrk_load_state checks the ABI version, the state layout and a checksum of the scenario parameters
    (same shape, different radar -> refused)."""
    import emu_lib
    from zrk_modulation_b200.engine import Engine, ZrkError
    lib = emu_lib.load()
    a = synthetic_batch(3, 6, n_radars=1, n_launchers=1, missiles_per_launcher=2, seed0=5)
    b = synthetic_batch(3, 6, n_radars=1, n_launchers=1, missiles_per_launcher=2, seed0=5)
    b.radar_par[:, 0, 2] = 12345.0  # another max_distance
    with Engine(a, lib=lib) as ea, Engine(b, lib=lib) as eb:
        ea.step(10)
        buf = ea.save_state()
        with pytest.raises(ZrkError):
            eb.load_state(buf)
        ea.load_state(buf)  # its own checkpoint is fine
        bad = buf.copy()
        bad[56:60] = 99  # the ABI version field (after magic, tick, three sizes, four ints)
        with pytest.raises(ZrkError):
            ea.load_state(bad)
