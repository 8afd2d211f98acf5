"""Comparison helpers shared by the parity tests.  Everything is compared bit for bit: integer
fields with ==, fp64 payloads and state through their uint64 bit patterns (NaN == NaN)."""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LIST_TYPES = (5, 6, 7)  # ALL_OBJECTS, FOUND_OBJECTS, ACTIVE_OBJECTS


def same_f64(x, y):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    if x.shape != y.shape:
        return False
    eq = (x.view(np.uint64) == y.view(np.uint64)) | (np.isnan(x) & np.isnan(y)) | ((x == 0) & (y == 0))
    return bool(eq.all())


def diff_msgs(am, aids, bm, bids, xyz_rtol=None):
    """Return '' if the two canonical logs are identical, else a description of the first
    difference.  With xyz_rtol the fp64 payloads are compared with that relative tolerance."""
    if len(am) != len(bm):
        n = min(len(am), len(bm))
        extra = am[n] if len(am) > n else bm[n]
        return f"length {len(am)} != {len(bm)}; first extra record {extra}"
    for f in ("tick", "type", "relevance", "flag", "sender", "receiver", "b"):
        bad = np.nonzero(am[f] != bm[f])[0]
        if len(bad):
            return f"field {f} differs at {bad[0]}: {am[bad[0]]} vs {bm[bad[0]]}"
    islist = np.isin(am["type"], LIST_TYPES)
    bad = np.nonzero((am["a"] != bm["a"]) & ~islist)[0]
    if len(bad):
        return f"field a differs at {bad[0]}: {am[bad[0]]} vs {bm[bad[0]]}"
    for i in np.nonzero(islist)[0]:
        x = aids[am["a"][i]:am["a"][i] + am["b"][i]]
        y = bids[bm["a"][i]:bm["a"][i] + bm["b"][i]]
        if not np.array_equal(x, y):
            return f"object list differs at {i} (tick {am['tick'][i]}, type {am['type'][i]}): {x} vs {y}"
    if xyz_rtol is None:
        if not same_f64(am["xyz"], bm["xyz"]):
            bad = np.nonzero((am["xyz"].view(np.uint64) != bm["xyz"].view(np.uint64)).any(axis=1))[0]
            return f"xyz bits differ at {bad[0]}: {am[bad[0]]} vs {bm[bad[0]]}"
    elif not np.allclose(am["xyz"], bm["xyz"], rtol=xyz_rtol, atol=0, equal_nan=True):
        return "xyz differs beyond tolerance"
    return ""


def diff_state(a, b, sa=0, sb=0, known=None, skip=()):
    """Compare two zrk_state_view dicts (scenario sa of a against sb of b)."""
    for name in a:
        if name in skip:
            continue
        x, y = a[name][sa], b[name][sb]
        if known is not None and name in ("pos", "prev_pos", "has_prev"):
            x, y = x[known], y[known]
        if name == "slot_obj":  # identity of a removed slot is only known to the oracle / device
            m = a["slot_alive"][sa].astype(bool) & b["slot_alive"][sb].astype(bool)
            x, y = x[m], y[m]
        ok = same_f64(x, y) if x.dtype == np.float64 else np.array_equal(x, y)
        if not ok:
            return f"state '{name}' differs:\n{x}\nvs\n{y}"
    return ""


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    g = dict(cfg=json.loads(str(z["cfg"])), msgs=z["msgs"], ids=z["ids"], counts=z["counts"],
             fault_code=int(z["fault"][0]), fault_tick=int(z["fault"][1]), fault_where=str(z["fault_where"]),
             state_known=z["state_known"], noise=z["noise"] if bool(z["has_noise"]) else None)
    g["state"] = {k[3:]: z[k] for k in z.files if k.startswith("st_")}
    return g
