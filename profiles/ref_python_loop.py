#!/usr/bin/env python
"""Times the UNMODIFIED reference's own tick loop, Manager.run_simulation (modules/Manager.py:111-140),
on every host core, over the same synthetic scenario families bench.py runs on the GPU.

    python profiles/ref_python_loop.py [--out profiles/ref_python_loop.json]

The reference is pure Python and does not travel to the GPU box, so this runs where /root/reference
(or $ZRK_REFERENCE_DIR) exists — the build container — and bench.py cites the JSON it writes
(`cpu_baseline.python_loop`); when the reference IS present bench.py calls `measure()` live instead.
Only SectorRadar.smooth_objects is replaced (noise from the shared table instead of
the unseeded global RNG) and logging is disabled; one OS process per core, each running its share of
the scenarios; the timed region is run_simulation alone (object construction is outside it).
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import multiprocessing as mp
import os
import platform
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SHAPES = {  # name: (targets, radars, launchers, missiles per launcher, ticks timed)
    "c5": (24, 1, 2, 4, 250),
    "c3": (64, 1, 0, 0, 60),
    "c4": (256, 2, 0, 0, 6),
}


def _one(args):
    name, seed, ticks = args
    import numpy as np
    import refharness
    from zrk_modulation_b200 import synthetic_batch
    from zrk_modulation_b200.scenario import batch_to_config
    T, R, L, mpl, _ = SHAPES[name]
    b = synthetic_batch(1, T, n_radars=R, n_launchers=L, missiles_per_launcher=mpl, seed0=seed)
    dt = int(b.dt_ms)
    b.durations[:] = ticks * dt
    cfg = batch_to_config(b, 0)
    noise = b.noise
    M = refharness.ref()
    from modules.Radar import SectorRadar
    mgr, _objs = M.create_objects_from_config(cfg)
    radars = [m for m in mgr.modules if isinstance(m, SectorRadar)]
    ridx = {id(r): i for i, r in enumerate(radars)}
    static = {int(b.target_id[0, i]): i for i in range(T)}
    static.update({int(b.missile_id[0, m]): T + m for m in range(L * mpl)})

    def smooth(self, objects):
        tick = self._manager.time.get_time() // dt
        for o in objects:
            o.pos += noise[tick % noise.shape[0], ridx[id(self)], static[o.id]]

    orig = SectorRadar.smooth_objects
    SectorRadar.smooth_objects = smooth
    fault = ""
    t0 = time.perf_counter()
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            mgr.run_simulation(ticks * dt)
    except Exception as e:  # the reference aborts on its own faults: count the ticks it ran
        fault = type(e).__name__
    dt_s = time.perf_counter() - t0
    SectorRadar.smooth_objects = orig
    ran = mgr.time.get_time() // dt + (1 if fault else 0)
    return dt_s, int(ran), (T + L * mpl), fault


def measure(name="c5", per_core=1, cores=None, ticks=None):
    """object-updates/s of the reference loop on `cores` processes over per_core x cores scenarios."""
    cores = cores or os.cpu_count() or 1
    T, R, L, mpl, dticks = SHAPES[name]
    ticks = ticks or dticks
    jobs = [(name, 100000 + i, ticks) for i in range(per_core * cores)]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_one, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    upd = sum(r[1] * r[2] for r in res)
    busy = sum(r[0] for r in res)
    return {"workload": name, "scenarios": len(jobs), "ticks": ticks, "slots": T + L * mpl, "cores": cores,
            "object_updates_per_s": upd / wall, "object_updates_per_s_per_core": upd / busy,
            "seconds_per_scenario": busy / len(jobs), "wall_s": wall, "faulted": sum(1 for r in res if r[3]),
            "what": "Manager.run_simulation of the unmodified reference (noise patch only, logging disabled), "
                    "one process per core; timed region = run_simulation"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "ref_python_loop.json"))
    ap.add_argument("--per-core", type=int, default=2)
    args = ap.parse_args()
    import refharness
    if not refharness.available():
        raise SystemExit("the reference is not present (set ZRK_REFERENCE_DIR)")
    import numpy
    out = {"host": {"cpu_count": os.cpu_count(), "machine": platform.machine(), "python": platform.python_version(),
                    "numpy": numpy.__version__, "where": "build container (no GPU)"},
           "runs": [measure("c5", args.per_core), measure("c3", 1), measure("c4", 1)]}
    json.dump(out, open(args.out, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
