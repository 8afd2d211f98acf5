#!/usr/bin/env python
"""Offline sweep: random scenarios of the four families of tests/scenarios.py through the emulated build of the
product sources (tests/emu), against the oracle — every message, every fp64 payload bit, faults, final state.
Each scenario is run twice: with every message stored as a record (ZRK_LOG_FULL: general tick only) and with the
log that scales (events + masks + track journal, messages rebuilt on the host: certificate, quiet streak and
two-kernel stepping active), at a ticks-per-launch setting that rotates through 0 / 1 / 7 / 64.

    python tests/emu_sweep.py [scenarios per family] [first seed] > profiles/r02_emu_sweep.log

TEST INFRASTRUCTURE (uses the emulator and the oracle); not part of the pytest suite because of its length.
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import emu_lib  # noqa: E402
from enginecheck import check_batch_full  # noqa: E402
from scenarios import random_config  # noqa: E402
from zrk_modulation_b200 import compile_batch, make_noise  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
    lib = emu_lib.load()
    t0 = time.time()
    total = bad = 0
    for family in ("mixed", "dense", "mid", "tiny"):
        fam_bad = 0
        for i in range(n):
            seed = seed0 + i
            cfg = random_config(seed, family)
            b = compile_batch(cfg)
            b.set_noise(make_noise(8, b.max_radars, b.n_slots, seed=seed))
            tpl = (0, 1, 7, 64)[i % 4]
            for rebuilt in (False, True):
                total += 1
                try:
                    check_batch_full(b, b.n_ticks(0), lib=lib, ticks_per_launch=tpl, rebuilt=rebuilt)
                except AssertionError as e:
                    fam_bad += 1
                    print(f"DIFF family={family} seed={seed} tpl={tpl} rebuilt={rebuilt}: {str(e)[:300]}")
        bad += fam_bad
        print(f"{family}: {n} scenarios x 2 log modes, {fam_bad} differences", flush=True)
    print(f"sweep: {total} runs, {bad} differences, seeds {seed0}..{seed0 + n - 1}, {time.time() - t0:.0f} s")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
