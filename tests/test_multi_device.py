"""A single process driving several GPUs (BatchedManager(devices=...), ShardedEngine): the reference's
caller is one process (main.run_simulation_from_config, main.py:151-174).  Each device owns a contiguous
block of scenarios; results are those of the one-device run, scenario for scenario."""
import numpy as np
import pytest

from zrk_modulation_b200 import synthetic_batch
from zrk_modulation_b200.manager import BatchedManager


def _same(m1, mN, B, sample):
    assert np.array_equal(m1.counts(), mN.counts())
    f1, fN = m1.status(), mN.status()
    assert np.array_equal(f1[0], fN[0]) and np.array_equal(f1[1], fN[1])
    assert np.array_equal(m1.engine.totals(), mN.engine.totals())
    for s in sample:
        a, b = m1.scenario(s).messages, mN.scenario(s).messages
        assert sorted(a.keys()) == sorted(b.keys())
        for t in a.keys():
            assert [repr(x) for x in a[t]] == [repr(x) for x in b[t]], (s, t)


def test_sharded_manager_matches_single_emulated():
    import emu_lib
    lib = emu_lib.load()
    B = 11  # (not a multiple of the device count: blocks of 4, 4, 3)
    b = synthetic_batch(B, 24, n_radars=1, n_launchers=2, missiles_per_launcher=4, seed0=21)
    b.durations[:] = 80 * 200
    m1 = BatchedManager(b, lib=lib)
    mN = BatchedManager(b, lib=lib, devices=[0, 0, 0])  # three handles (one device in the emulated build)
    m1.run_simulation(80 * 200)
    mN.run_simulation(80 * 200)
    _same(m1, mN, B, [0, 3, 4, 10])
    m1.close()
    mN.close()


@pytest.mark.gpu
def test_sharded_manager_on_all_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    B = 4096 + 5
    b = synthetic_batch(B, 24, n_radars=1, n_launchers=2, missiles_per_launcher=4, seed0=31)
    b.durations[:] = 120 * 200
    m1 = BatchedManager(b)
    mN = BatchedManager(b, devices="all")
    m1.run_simulation(120 * 200)
    mN.run_simulation(120 * 200)
    _same(m1, mN, B, [0, 1, B // 2, B - 1])
    m1.close()
    mN.close()


@pytest.mark.gpu
def test_two_handles_on_one_gpu_threads():
    """The threaded path of ShardedEngine (one host thread per handle) on a single GPU."""
    B = 515
    b = synthetic_batch(B, 24, n_radars=1, n_launchers=2, missiles_per_launcher=4, seed0=41)
    b.durations[:] = 100 * 200
    m1 = BatchedManager(b)
    mN = BatchedManager(b, devices=[0, 0])
    m1.run_simulation(100 * 200)
    mN.run_simulation(100 * 200)
    _same(m1, mN, B, [0, 257, 258, B - 1])
    m1.close()
    mN.close()
