"""The oracle (oracle/zrk_oracle.c) against what the unmodified reference produced
(tests/golden/*.npz, made by tests/golden/make_golden.py): every message, every fp64 payload bit,
every fault, the final state.  This is what pins the oracle."""
import numpy as np
import pytest

import parity
from oracle import oracle as O
from zrk_modulation_b200 import compile_batch


@pytest.mark.parametrize("name", parity.golden_names())
def test_oracle_matches_reference_fixture(name):
    g = parity.load_golden(name)
    b = compile_batch(g["cfg"], noise=g["noise"])
    r = O.run(b, 0, b.n_ticks(0), numpy_pow=True)
    assert parity.diff_msgs(g["msgs"], g["ids"], r["msgs"], r["ids"]) == ""
    assert (r["fault_code"], r["fault_tick"]) == (g["fault_code"], g["fault_tick"]), g["fault_where"]
    assert np.array_equal(r["counts"], g["counts"])
    assert parity.diff_state(g["state"], r["state"], known=g["state_known"]) == ""


@pytest.mark.parametrize("name", parity.golden_names())
def test_correctly_rounded_mode_same_decisions(name):
    """Without the libm pow() of numpy's `**` (what the CUDA path computes) the log has the same
    messages; fp64 payloads may differ in the last bits of a launch solution."""
    g = parity.load_golden(name)
    b = compile_batch(g["cfg"], noise=g["noise"])
    r = O.run(b, 0, b.n_ticks(0), numpy_pow=False)
    assert parity.diff_msgs(g["msgs"], g["ids"], r["msgs"], r["ids"], xyz_rtol=1e-12) == ""
    assert (r["fault_code"], r["fault_tick"]) == (g["fault_code"], g["fault_tick"])


def test_known_answers_of_shipped_configs():
    """seed-independent known answers of the shipped configs."""
    from zrk_modulation_b200 import _cabi as A
    g = parity.load_golden("shipped_simulation_config")
    assert len(g["msgs"]) == 685
    c = g["counts"]
    assert c[A.MSG_ACTIVE_OBJECTS] == c[A.MSG_ALL_OBJECTS] == c[A.MSG_FOUND_OBJECTS] == 200
    assert c[A.MSG_DRAW_OBJECTS] == 42 and c[A.MSG_MISSILE_POS] == 16 and c[A.MSG_CCP_UPDATE_TARGET] == 9
    det = g["msgs"][g["msgs"]["type"] == A.MSG_MISSILE_DETONATE]
    assert list(det["tick"]) == [11, 13]
    g = parity.load_golden("shipped_config")
    assert len(g["msgs"]) == 2888
    det = g["msgs"][g["msgs"]["type"] == A.MSG_MISSILE_DETONATE]
    assert list(det["tick"]) == [74, 146] and list(det["a"]) == [10, 11]
    assert len(parity.load_golden("shipped_simulation_config_copy")["msgs"]) == 800


def test_timer_rounding_kat():
    """life_time 35 s at dt 200 ms expires after 176 decrements, not 175."""
    period, n = 35.0, 0
    while True:
        period -= 200 / 1000
        n += 1
        if period <= 0:
            break
    assert n == 176
