"""Live cross-check of the oracle against the reference imported from /root/reference (build
container only; skipped elsewhere — the committed fixtures cover the GPU box)."""
import numpy as np
import pytest

import parity
import refharness as RH
from oracle import oracle as O
from scenarios import random_config
from zrk_modulation_b200 import compile_batch, make_noise

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not RH.available(), reason="reference tree not present")]


@pytest.mark.parametrize("family,seed", [(f, s) for f in ("mixed", "dense", "mid", "tiny") for s in range(100, 112)])
def test_oracle_vs_live_reference(family, seed):
    cfg = random_config(seed, family)
    b = compile_batch(cfg)
    noise = make_noise(8, b.max_radars, b.n_slots, seed=seed)
    ref = RH.run_reference(cfg, noise)
    b.set_noise(noise)
    r = O.run(b, 0, b.n_ticks(0), numpy_pow=True)
    assert parity.diff_msgs(ref["msgs"], ref["ids"], r["msgs"], r["ids"]) == ""
    assert (r["fault_code"], r["fault_tick"]) == (ref["fault_code"], ref["fault_tick"]), ref["fault_where"]
    assert parity.diff_state(ref["state"], r["state"], known=ref["state_known"]) == ""
