"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Each fixture is one scenario: its config (reference YAML schema, as JSON), the injected radar
noise table, and what the reference produced: the canonical message log (zrk_msg records +
id pool of the list-carrying messages), per-type message counts, the fault (exception) it hit,
and the final model state.  tests/test_oracle_golden.py replays them through the oracle.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

import refharness as RH  # noqa: E402
from scenarios import SPECIAL_KINDS, random_config, special_config  # noqa: E402
from zrk_modulation_b200 import compile_batch, load_config, make_noise  # noqa: E402


def save(name, cfg, noise):
    ref = RH.run_reference(cfg, noise)
    out = dict(cfg=json.dumps(cfg), msgs=ref["msgs"], ids=ref["ids"], counts=ref["counts"],
               fault=np.array([ref["fault_code"], ref["fault_tick"]]), fault_where=ref["fault_where"],
               state_known=ref["state_known"], has_noise=noise is not None,
               noise=noise if noise is not None else np.zeros(0))
    for k, v in ref["state"].items():
        out["st_" + k] = v
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: {len(ref['msgs'])} msgs, fault {ref['fault_where'] or '-'}")


def main():
    for fn in ("config.yaml", "simulation_config.yaml", "simulation_config copy.yaml"):
        cfg = load_config(os.path.join(RH.REFERENCE_DIR, fn))
        b = compile_batch(cfg)
        tag = fn.replace(".yaml", "").replace(" ", "_")
        save(f"shipped_{tag}", cfg, make_noise(8, b.max_radars, b.n_slots, seed=7))
        save(f"shipped_{tag}_nonoise", cfg, None)
    # random families; seeds picked to include launches, detonations, cancels and both faults
    for fam, seeds in (("mixed", range(0, 10)), ("dense", range(0, 8)), ("mid", range(0, 8)), ("tiny", range(0, 6))):
        for seed in seeds:
            cfg = random_config(seed, fam)
            b = compile_batch(cfg)
            noise = make_noise(8, b.max_radars, b.n_slots, seed=seed) if seed % 5 else None
            save(f"{fam}_{seed:02d}", cfg, noise)


def special(only=None):
    for kind in SPECIAL_KINDS:
        if only and kind not in only:
            continue
        cfg = special_config(kind)
        b = compile_batch(cfg)
        save(f"special_{kind}", cfg, make_noise(8, b.max_radars, b.n_slots, seed=5))
        save(f"special_{kind}_nonoise", cfg, None)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "special":
        special(sys.argv[2:])  # optionally only the named kinds
    else:
        main()
        special()
