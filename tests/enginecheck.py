"""Run a Batch through an Engine (CUDA library, or the emulated build for host-logic tests) and
compare every scenario with the oracle: full message log, counts, fault status, final state."""
import numpy as np

import parity
from oracle import oracle as O
from zrk_modulation_b200 import LOG_FULL
from zrk_modulation_b200.engine import Engine


def check_batch_full(batch, n_ticks, lib=None, ticks_per_launch=0, exact_only=False, scenarios=None,
                     step_chunks=None, rebuilt=False):
    """Bit-exact comparison of all messages + state of the given scenarios (default: all).
    rebuilt: the log that scales — event records + masks + track journal only (the quiet streak and the
    certificate run), bulk messages rebuilt on the host by zrk_read_messages."""
    from zrk_modulation_b200 import LOG_EVENTS, LOG_MASKS
    B = batch.n_scenarios
    scenarios = range(B) if scenarios is None else scenarios
    # (rebuilt: events + track journal; dense scenarios change tracks all the time, give them room)
    cap = 64 * batch.n_slots + 512 if rebuilt else (3 * batch.n_slots + 16 + 4 * batch.max_radars) * n_ticks + 64
    with Engine(batch, log_types=(LOG_EVENTS | LOG_MASKS) if rebuilt else LOG_FULL, log_capacity=cap, mask_ticks=n_ticks + 1,
                ticks_per_launch=ticks_per_launch, exact_only=exact_only, lib=lib) as eng:
        if step_chunks:
            for k in step_chunks:
                eng.step(k)
            assert sum(step_chunks) == n_ticks
        else:
            eng.step(n_ticks)
        eng.sync()
        fc, ft, nr = eng.status()
        counts = eng.counts()
        alive = eng.alive_slot_ticks()
        st = eng.state()
        from zrk_modulation_b200 import _cabi as A
        for s in scenarios:
            r = O.run(batch, s, n_ticks)
            act = r["msgs"][r["msgs"]["type"] == A.MSG_ACTIVE_OBJECTS]
            assert int(alive[s]) == int(act["b"].sum()), f"scenario {s}: alive slot-ticks"
            msgs, ids = eng.messages(s, 0, n_ticks)
            d = parity.diff_msgs(r["msgs"], r["ids"], msgs, ids)
            assert d == "", f"scenario {s}: {d}"
            assert (r["fault_code"], r["fault_tick"]) == (int(fc[s]) & 63, int(ft[s])), f"scenario {s} fault"
            assert np.array_equal(counts[s], r["counts"]), f"scenario {s} counts"
            d = parity.diff_state(r["state"], st, 0, s)
            assert d == "", f"scenario {s}: {d}"


def check_batch_counts(batch, n_ticks, lib=None, ticks_per_launch=0, n_threads=4, sample=None):
    """Counts / faults / event records of every scenario against the oracle (events mode)."""
    B = batch.n_scenarios
    with Engine(batch, ticks_per_launch=ticks_per_launch, lib=lib) as eng:
        eng.step(n_ticks)
        eng.sync()
        fc, ft, nr = eng.status()
        counts = eng.counts()
        recs, nrec = eng.records(0, B) if B <= 4096 else (None, None)
    oc, ofc, oft = O.run_batch(batch, 0, B, n_ticks, n_threads=n_threads)
    assert np.array_equal(fc & 63, ofc), "fault codes differ"
    assert np.array_equal(ft, oft), "fault ticks differ"
    assert np.array_equal(counts, oc), "message counts differ"
    assert not (fc & 64).any(), "event log overflow"
    if recs is not None:
        rng = np.random.default_rng(0)
        pick = rng.choice(B, size=min(B, sample or 16), replace=False)
        from zrk_modulation_b200 import _cabi as A
        bulk = [A.MSG_ALL_OBJECTS, A.MSG_FOUND_OBJECTS, A.MSG_ACTIVE_OBJECTS, A.MSG_CCP_UPDATE_TARGET,
                A.MSG_DRAW_OBJECTS, A.MSG_MISSILE_POS]
        for s in pick:
            r = O.run(batch, int(s), n_ticks, state=False)
            ev = r["msgs"][~np.isin(r["msgs"]["type"], bulk)]
            # the device appends in emission order; the reference sorts tick 0 by relevance
            got = recs[s, :nrec[s]]
            key = lambda m: np.lexsort((np.arange(len(m)), -m["relevance"].astype(int), m["tick"]))
            got = got[key(got)]
            d = parity.diff_msgs(ev, r["ids"], got, r["ids"])
            assert d == "", f"scenario {s} events: {d}"
    return counts
