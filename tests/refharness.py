"""Runs the UNMODIFIED reference (/root/reference) as the ground truth and canonicalises its
message log into zrk_msg records (include/zrk_b200.h).

Only usable in the build container (the reference does not travel to the GPU box): used by
tests/golden/make_golden.py to produce the committed fixtures and by tests marked `reference`.
The single patch applied is SURVEY Appendix C's: SectorRadar.smooth_objects (Radar.py:138-142)
adds noise[tick % W, radar_index, static_index] instead of drawing from the unseeded global RNG.
"""
from __future__ import annotations

import contextlib
import io
import logging
import os
import sys
import traceback
import warnings
from typing import Any, Dict, Optional

import numpy as np

REFERENCE_DIR = os.environ.get("ZRK_REFERENCE_DIR", "/root/reference")

from zrk_modulation_b200 import _cabi
from zrk_modulation_b200.scenario import compile_batch

_ref = None

CANCEL_REASONS = [
    "No interception possible",           # Missile.py:73
    "Interception impossible in the future",  # Missile.py:79
    "No real interception time",          # Missile.py:83
    "Interception times are not positive",  # Missile.py:90
    "Target is too far",                  # Missile.py:94
]


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "main.py"))


def ref():
    """Import the reference's main module read-only (no bytecode written)."""
    global _ref
    if _ref is None:
        if not available():
            raise RuntimeError("reference not present")
        sys.dont_write_bytecode = True
        if REFERENCE_DIR not in sys.path:
            sys.path.insert(0, REFERENCE_DIR)
        logging.disable(logging.CRITICAL)
        warnings.filterwarnings("ignore")
        import main as M  # noqa
        _ref = M
    return _ref


def canonicalise(messages, dt):
    """{step time: [message objects]} -> (zrk_msg records, id pool, per-type counts).  Works on the
    reference's message objects and on the mirror's (same attribute names)."""
    NONE = _cabi.ZRK_NONE
    nn = lambda v: NONE if v is None else int(v)
    recs, ids = [], []
    counts = np.zeros(_cabi.ZRK_MSG_NTYPES, dtype=np.int64)
    for t in sorted(messages):
        for m in messages[t]:
            ty = _cabi.MSG_NAMES.index(m.type.name)
            counts[ty] += 1
            a = b = NONE
            flag = 0
            xyz = (0.0, 0.0, 0.0)
            if ty == _cabi.MSG_MISSILE_POS:
                a = m.missile_id
            elif ty == _cabi.MSG_MISSILE_DETONATE:
                a, b, flag = m.missile_id, nn(m.target_id), int(bool(m.self_detonation))
            elif ty in (_cabi.MSG_ACTIVE_OBJECTS, _cabi.MSG_ALL_OBJECTS, _cabi.MSG_FOUND_OBJECTS):
                lst = (m.active_objects if ty == _cabi.MSG_ACTIVE_OBJECTS else
                       m.objects if ty == _cabi.MSG_ALL_OBJECTS else m.visible_objects)
                a, b = len(ids), len(lst)
                ids.extend(int(o.id) for o in lst)
            elif ty == _cabi.MSG_DESTROYED_MISSILE:
                assert isinstance(m.missile_id, tuple) and len(m.missile_id) == 1
                a, flag = int(m.missile_id[0]), int(bool(m.self_detonation))
            elif ty == _cabi.MSG_LAUNCH_MISSILE:
                a = int(m.target.id)
            elif ty == _cabi.MSG_LAUNCH_SUCCESSFUL:
                a, b = int(m.missile.id), int(m.target_id)
                xyz = tuple(float(x) for x in m.missile.trajectory.velocity)
                assert m.launch_time == t / 1000
            elif ty == _cabi.MSG_LAUNCH_CANCELLED:
                a = int(m.missile.id)
                b = [i for i, r in enumerate(CANCEL_REASONS) if m.reason.startswith(r)][0]
            elif ty == _cabi.MSG_LAUNCHED_MISSILE:
                a, b = int(m.missile.id), int(m.target_id)
            elif ty == _cabi.MSG_NEW_MISSILE:
                a = int(m.missile.id)
            elif ty == _cabi.MSG_MISSILE_COUNT_RESPONSE:
                b = int(m.count)
            elif ty == _cabi.MSG_LAUNCH_COMMAND:
                a, b = int(m.target.id), int(m.radar_id)
                xyz = tuple(float(x) for x in m.target_position)
            elif ty == _cabi.MSG_CCP_UPDATE_TARGET:
                a, b = int(m.target.id), nn(m.missile_id)
            elif ty == _cabi.MSG_DRAW_OBJECTS:
                a = int(m.obj_id)
                b = 3 if isinstance(m.target_type, str) else list(type(m.target_type)).index(m.target_type)
                flag = int(bool(m.is_visible_by_radar))
                xyz = tuple(float(x) for x in m.coordinates)
            elif ty == _cabi.MSG_MISSILE_COUNT_REQUEST:
                pass
            else:
                raise AssertionError(f"unexpected message type {m.type}")
            assert m.send_time == t
            recs.append((t // dt, ty, int(m.relevance), flag, 0, int(m.sender_id), nn(m.receiver_id),
                         int(a), int(b), xyz))
    msgs = np.array(recs, dtype=_cabi.MSG_DTYPE) if recs else np.zeros(0, dtype=_cabi.MSG_DTYPE)
    return msgs, np.array(ids, dtype=np.int32), counts


def run_reference(cfg: Dict[str, Any], noise: Optional[np.ndarray], n_ticks: Optional[int] = None):
    """Run one scenario through the reference.  Returns dict(msgs, ids, fault_code, fault_tick,
    fault_where, counts, state) in the oracle's output format."""
    M = ref()
    from modules.Radar import SectorRadar
    from modules.utils import Target
    from modules.Missile import Missile
    from modules.MissileLauncher import MissileLauncher
    from modules.CCP import CombatControlPoint
    from modules.constants import MessageType

    batch = compile_batch(cfg)
    T = batch.max_targets
    static = {}
    for i in range(int(batch.n_targets[0])):
        static[int(batch.target_id[0, i])] = i
    for m in range(int(batch.n_missiles[0])):
        static[int(batch.missile_id[0, m])] = T + m
    mgr, objs = M.create_objects_from_config(cfg)
    dt = cfg["simulation"]["time_step"]
    radars = [m for m in mgr.modules if isinstance(m, SectorRadar)]
    ridx = {id(r): i for i, r in enumerate(radars)}

    def smooth(self, objects):
        if noise is None:
            return
        tick = self._manager.time.get_time() // dt
        for o in objects:
            o.pos += noise[tick % noise.shape[0], ridx[id(self)], static[o.id]]

    all_objs = {}
    for o in mgr.modules[0]._AirEnv__objects:
        all_objs[static[o.id]] = o
    for l in mgr.modules:
        if isinstance(l, MissileLauncher):
            for mi in l.missiles:
                all_objs[static[mi.id]] = mi
    orig = SectorRadar.smooth_objects
    SectorRadar.smooth_objects = smooth
    duration = cfg["simulation"]["duration"] if n_ticks is None else n_ticks * dt
    fault_code, fault_tick, fault_where = 0, 0, ""
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            mgr.run_simulation(duration)
    except Exception as e:  # the reference aborts; we record where
        tb = traceback.extract_tb(e.__traceback__)[-1]
        fault_where = f"{type(e).__name__}@{os.path.basename(tb.filename)}:{tb.lineno}"
        fault_tick = mgr.time.get_time() // dt
        if isinstance(e, AttributeError) and tb.lineno == 259 and tb.filename.endswith("CCP.py"):
            fault_code = _cabi.FAULT_ATTRIBUTE_ERROR_CCP259
        elif isinstance(e, KeyError) and tb.lineno == 107 and tb.filename.endswith("CCP.py"):
            fault_code = _cabi.FAULT_KEY_ERROR_CCP107
        else:
            fault_code = 99
    finally:
        SectorRadar.smooth_objects = orig

    msgs, ids, counts = canonicalise(mgr.messages, dt)

    # final state in the zrk_state_view layout (batch of one)
    NONE = _cabi.ZRK_NONE
    nn = lambda v: NONE if v is None else int(v)
    B, NS, R, Mx, L, Cc = batch.dims()
    st, _ = _cabi.alloc_state((1, NS, R, Mx, L, Cc))
    env = mgr.modules[0]
    slots = env._AirEnv__objects
    st["slot_obj"][:] = -1
    st["missile_target"][:] = -1
    for k, o in enumerate(slots):
        st["slot_alive"][0, k] = o is not None
        # identity of removed slots is not recoverable from the reference list; filled by caller
        if o is not None:
            st["slot_obj"][0, k] = static[o.id]
    air = {}
    for k, o in enumerate(slots):
        if o is not None:
            air[static[o.id]] = o
    launchers = [m for m in mgr.modules if isinstance(m, MissileLauncher)]
    ccps = [m for m in mgr.modules if isinstance(m, CombatControlPoint)]

    def note(o):
        if o is not None and hasattr(o, "id") and o.id in static:
            all_objs[static[o.id]] = o

    for o in slots:
        note(o)
    for l in launchers:
        for mi in l.missiles + l.launched_missiles:
            note(mi)
    for c in ccps:
        for e in c._target_dict.values():
            note(e.target)
        for e in c._missile_dict.values():
            note(e.missile)
            note(getattr(e.missile, "target", None))
    for si, o in all_objs.items():
        st["pos"][0, si] = np.asarray(o.pos, dtype=np.float64)
        pp = getattr(o, "prev_pos", None)
        st["has_prev"][0, si] = pp is not None
        if pp is not None:
            st["prev_pos"][0, si] = pp
        if isinstance(o, Missile):
            m = si - T
            st["missile_status"][0, m] = {"ready": 0, "active": 1, "detonated": 2}[o.status]
            st["missile_period"][0, m] = o.detonate_period
            st["missile_vel"][0, m] = o.trajectory.velocity
            st["missile_t0"][0, m] = o.trajectory.start_time
            st["missile_target"][0, m] = -1 if o.target is None else static[o.target.id]
    known = np.zeros(NS, dtype=bool)
    known[list(all_objs)] = True
    for i, r in enumerate(radars):
        st["radar_az"][0, i], st["radar_el"][0, i] = r.current_azimuth, r.current_elevation
    for i, l in enumerate(launchers):
        st["launcher_count"][0, i] = len(l.missiles)
    if ccps:
        c = ccps[0]
        st["ccp_n_targets"][0], st["ccp_n_missiles"][0] = len(c._target_dict), len(c._missile_dict)
        st["ccp_t_ref"][:], st["ccp_m_ref"][:] = -1, -1
        st["ccp_t_missile"][:], st["ccp_m_target"][:] = NONE, NONE
        for e, (k, v) in enumerate(c._target_dict.items()):
            st["ccp_t_key"][0, e], st["ccp_t_ref"][0, e] = k, static[v.target.id]
            st["ccp_t_upd"][0, e] = int(round(v.upd_time * 1000)) // dt
            st["ccp_t_missile"][0, e] = nn(v.missile_id)
            st["ccp_t_following"][0, e] = bool(v.following)
        for e, (k, v) in enumerate(c._missile_dict.items()):
            st["ccp_m_key"][0, e], st["ccp_m_ref"][0, e] = k, static[v.missile.id]
            st["ccp_m_upd"][0, e] = int(round(v.upd_time * 1000)) // dt
            st["ccp_m_target"][0, e] = nn(v.target_id)
        for i, k in enumerate(c.missile_launcher_coords):
            st["ccp_capacity"][0, i] = c.missile_launcher_capacity.get(k, 0)
            st["ccp_launched"][0, i] = c.missile_launcher_launched.get(k, 0)
    else:
        st["ccp_t_ref"][:], st["ccp_m_ref"][:] = -1, -1
        st["ccp_t_missile"][:], st["ccp_m_target"][:] = NONE, NONE
    return dict(msgs=msgs, ids=ids, fault_code=fault_code,
                fault_tick=int(fault_tick), fault_where=fault_where, counts=counts, state=st,
                state_known=known, batch=batch)
