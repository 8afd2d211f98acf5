"""Error behaviour of the C ABI: no exception crosses it — return codes + zrk_last_error — and
nothing is silently truncated (log overflow, mask-history ring, bad descriptors).
`-m "not gpu"` through the emulated build, `-m gpu` through libzrk_b200.so."""
import ctypes as C

import numpy as np
import pytest

from zrk_modulation_b200 import LOG_FULL, _cabi, compile_batch, make_noise, synthetic_batch
from zrk_modulation_b200.engine import Engine, ZrkError
from scenarios import random_config


def _lib(request):
    if request.node.get_closest_marker("gpu"):
        return _cabi.load()
    import emu_lib
    return emu_lib.load()


def _log_overflow(lib):
    b = synthetic_batch(2, 8, n_radars=1, noise_period=4)
    with Engine(b, log_types=LOG_FULL, log_capacity=40, mask_ticks=32, lib=lib) as eng:
        eng.step(20)
        fc, ft, nr = eng.status()
        assert (fc & _cabi.FAULT_LOG_OVERFLOW).all() and (nr > 40).all()   # the cursor keeps counting
        assert (eng.counts()[:, _cabi.MSG_ACTIVE_OBJECTS] == 20).all()      # the simulation itself is unaffected
        with pytest.raises(ZrkError) as e:
            eng.messages(0)
        assert e.value.rc == _cabi.ERR_CAPACITY and "log_capacity" in str(e.value)


def _mask_ring(lib):
    b = synthetic_batch(1, 4, n_radars=1, noise_period=4)
    with Engine(b, log_types=LOG_FULL, log_capacity=2000, mask_ticks=8, lib=lib) as eng:
        eng.step(20)
        with pytest.raises(ZrkError) as e:
            eng.messages(0, 0, 20)
        assert e.value.rc == _cabi.ERR_CAPACITY and "mask_ticks" in str(e.value)
        msgs, ids = eng.messages(0, 14, 20)   # the last 8 ticks are still in the ring
        assert sorted(set(msgs["tick"])) == list(range(14, 20))


def _bad_descriptors(lib):
    b = synthetic_batch(1, 4, n_radars=1, noise_period=4)
    d = b.desc()
    h = C.c_void_p()
    d.abi_version = 99
    assert lib.zrk_create(C.byref(d), None, C.byref(h)) == _cabi.ERR_ARG and b"abi_version" in lib.zrk_last_error(None)
    d = b.desc()
    d.target_pos = None
    assert lib.zrk_create(C.byref(d), None, C.byref(h)) == _cabi.ERR_ARG and b"missing array" in lib.zrk_last_error(None)
    d = b.desc()
    d.dt_ms = 0
    assert lib.zrk_create(C.byref(d), None, C.byref(h)) == _cabi.ERR_ARG
    d = b.desc()
    d.noise_objects = 1   # noise table too small for the slots
    assert lib.zrk_create(C.byref(d), None, C.byref(h)) == _cabi.ERR_ARG
    assert lib.zrk_create(None, None, C.byref(h)) == _cabi.ERR_ARG
    b2 = synthetic_batch(2, 4, n_radars=1, n_launchers=1, missiles_per_launcher=1, noise_period=4)
    b2.missile_launcher[1, 0] = 7   # launcher index out of range
    with pytest.raises(ZrkError) as e:
        Engine(b2, lib=lib)
    assert e.value.rc == _cabi.ERR_ARG
    b3 = synthetic_batch(2, 4, n_radars=1, n_launchers=2, missiles_per_launcher=1, noise_period=4)
    b3.launcher_id[0, 1] = b3.launcher_id[0, 0]
    with pytest.raises(ZrkError) as e:
        Engine(b3, lib=lib)
    assert e.value.rc == _cabi.ERR_UNSUPPORTED
    assert lib.zrk_step(None, 1) == _cabi.ERR_ARG and lib.zrk_sync(None) == _cabi.ERR_ARG


def _ragged_durations(lib):
    from enginecheck import check_batch_full
    cfgs = [random_config(430 + i, "mixed") for i in range(6)]
    for i, c in enumerate(cfgs):
        c["simulation"] = {"time_step": 200, "duration": 200 * (5 + 9 * i)}
    b = compile_batch(cfgs)
    b.set_noise(make_noise(8, b.max_radars, b.n_slots, seed=2))
    check_batch_full(b, 50, lib=lib if lib is not _cabi._LIB else None, step_chunks=[3, 40, 7])


def _waves_agree(lib):
    """zrk_options.waves: scenario ranges on separate streams give the very same run."""
    b = synthetic_batch(300, 6, n_radars=1, n_launchers=1, missiles_per_launcher=2, noise_period=8, box_km=6.0)
    outs = []
    for waves in (0, 3, 64):   # 300 scenarios -> ranges of 128, 128, 44; 64 asks for more waves than there are ranges
        with Engine(b, waves=waves, ticks_per_launch=7, lib=lib) as eng:
            eng.step(5)
            eng.step(25)
            fc, ft, nr = eng.status()
            recs, n = eng.records()
            live = np.arange(recs.shape[1])[None, :] < n[:, None]
            outs.append((eng.counts(), fc, ft, nr, np.where(live, recs["type"], -1), eng.state()))
    ref = outs[0]
    assert ref[0][:, _cabi.MSG_LAUNCH_MISSILE].sum() > 0   # the runs are not trivial
    for o in outs[1:]:
        for a, c in zip(ref[:5], o[:5]):
            assert np.array_equal(a, c)
        for k in ref[5]:
            assert np.array_equal(ref[5][k], o[5][k], equal_nan=True), k


def _waves_defer_validation(lib):
    """With waves zrk_create does not wait for the init kernel: its verdict arrives with the first
    synchronising call and stays."""
    b = synthetic_batch(260, 4, n_radars=1, n_launchers=2, missiles_per_launcher=1, noise_period=4)
    b.launcher_id[200, 1] = b.launcher_id[200, 0]          # in the second wave
    with pytest.raises(ZrkError):
        Engine(b, lib=lib)                                  # without waves: at create
    eng = Engine(b, waves=2, lib=lib)                       # with waves: accepted ...
    eng.step(3)
    for call in (eng.sync, eng.counts, eng.status, lambda: eng.step(1)):
        with pytest.raises(ZrkError) as e:                  # ... and reported from then on
            call()
        assert e.value.rc == _cabi.ERR_UNSUPPORTED and "unique" in str(e.value)
    eng.close()


CASES = [_log_overflow, _mask_ring, _bad_descriptors, _ragged_durations, _waves_agree, _waves_defer_validation]


@pytest.mark.parametrize("fn", CASES, ids=lambda f: f.__name__.strip("_"))
def test_error_paths_emulated(request, fn):
    fn(_lib(request))


@pytest.mark.gpu
@pytest.mark.parametrize("fn", CASES, ids=lambda f: f.__name__.strip("_"))
def test_error_paths_cuda(request, fn):
    fn(_lib(request))
