"""Mirror of the reference's message vocabulary: modules/constants.py, modules/BaseMessage.py and
modules/Messages.py — same class names, constructor signatures, attributes and __repr__ — so that
code written against the reference's `Manager.messages` (main.py:171, UI/PolygonEditor.py:627-669)
reads the batched engine's log unchanged.

Messages are materialised lazily from the device log (zrk_msg records).  Where the reference
stores references to live model objects (`target`, `missile`, object lists), the mirror stores
`AirObjectView`s: same `.id` / `.pos` / `.type` / `.trajectory.velocity` attributes, backed by
the engine's state.
"""
from __future__ import annotations

import inspect
from abc import ABCMeta
from enum import Enum
from typing import List, Optional

import numpy as np

from . import _cabi


class MessageType(Enum):  # modules/constants.py:3-21
    LAUNCH_MISSILE = "launch_missile"
    LAUNCHED_MISSILE = "launched_missile"
    LAUNCH_COMMAND = "launch_command"
    MISSILE_COUNT_REQUEST = "missile_count_request"
    MISSILE_COUNT_RESPONSE = "missile_count_response"
    ALL_OBJECTS = "all_objects"
    FOUND_OBJECTS = "found_objects"
    ACTIVE_OBJECTS = "active_objects"
    CCP_UPDATE_TARGET = 'ccp_upd_target'
    UPDATE_TARGET = 'upd_target'
    MISSILE_GET_HIT = 'missile_get_hit'
    DESTROYED_MISSILE = 'destroyed_missile'
    DRAW_OBJECTS = 'draw_objects'
    MISSILE_POS = 'missile_pos'
    MISSILE_DETONATE = 'missile_detonate'
    NEW_MISSILE = 'new_missile'
    LAUNCH_CANCELLED = 'launch_cancelled'
    LAUNCH_SUCCESSFUL = 'launch_successful'


MESSAGE_TYPES = list(MessageType)
assert [t.name for t in MESSAGE_TYPES] == _cabi.MSG_NAMES

# modules/constants.py:23-38
SIMULATION_STEP = 1
MISSILE_VELOCITY_MODULE = 1600
MISSILE_DETONATE_PERIOD = 120
MISSILE_DETONATE_RADIUS = MISSILE_VELOCITY_MODULE / 1000 * SIMULATION_STEP
MIN_DIST_DETECTION = 30
MAX_DIST_DETECTION = 50000
POSSIBLE_TARGET_RADIUS = 100
MISSILE_TYPE_DRAWER = 0
TARGET_TYPE_DRAWER = 1
CCP_ID = 0
DRAWER_ID = 1
MANAGER_ID = 2


class TargetType(Enum):  # modules/utils.py:6-9
    AIR_PLANE = "самолет"
    HELICOPTER = "вертолет"
    ANOTHER = "другое"


TARGET_TYPES = list(TargetType)
ZUR = 'ЗУР'  # CCP.py:247, :400

CANCEL_REASONS = [  # Missile.py:73-94
    "No interception possible: target and interceptor are stationary relative or parallel.",
    "Interception impossible in the future: computed time t <= 0.",
    "No real interception time: target is too fast or out of range.",
    "Interception times are not positive; interception not possible in future.",
    "Target is too far for this rocket (detonation_period over limited)",
]


class _TrajectoryView:
    def __init__(self, velocity, start_pos, start_time):
        self.velocity = np.asarray(velocity, dtype=np.float64)
        self.start_pos = np.asarray(start_pos, dtype=np.float64)
        self.start_time = start_time


class AirObjectView:
    """Stand-in for a reference Target / Missile object carried inside a message."""

    def __init__(self, id: int, is_missile: bool, type=None, pos=None, trajectory=None):
        self.id = id
        self.is_missile = is_missile
        if not is_missile:
            self.type = type
        self.pos = pos
        self.trajectory = trajectory

    def __repr__(self):
        kind = "Missile" if self.is_missile else "Target"
        return f"{kind}(id={self.id}, pos={self.pos})"


class BaseMessage(metaclass=ABCMeta):  # modules/BaseMessage.py:4-23
    def __init__(self, type: MessageType, sender_id: int, receiver_id: int = None, send_time: int = None,
                 relevance: int = 1) -> None:
        self.relevance = relevance
        self.send_time = send_time
        self.type = type
        self.sender_id = sender_id
        self.receiver_id = receiver_id

    def __repr__(self) -> str:
        return (f"{self.__class__.__name__}: send_time={self.send_time}, type={self.type}, "
                f"sender_id={self.sender_id}, receiver_id={self.receiver_id}, relevance={self.relevance}")


_REQ = inspect.Parameter.empty


def _message_class(name, mtype, params, payload=(), tail=None, where=""):
    """Build one message class of the reference's vocabulary from its description.

    params:  constructor parameters in the reference's order, (name, default) with _REQ = required;
             `sender_id`, `receiver_id`, `time` (-> send_time) and `relevance` feed BaseMessage.
    payload: (attribute, function of the bound arguments) pairs set on the instance.
    tail:    function of the instance giving what the reference's __repr__ appends."""
    sig = inspect.Signature([inspect.Parameter("self", inspect.Parameter.POSITIONAL_OR_KEYWORD)] +
                            [inspect.Parameter(n, inspect.Parameter.POSITIONAL_OR_KEYWORD, default=d) for n, d in params])

    def __init__(self, *args, **kwargs):
        v = sig.bind(self, *args, **kwargs)
        v.apply_defaults()
        v = v.arguments
        BaseMessage.__init__(self, type=mtype, sender_id=v["sender_id"], receiver_id=v.get("receiver_id"),
                             send_time=v.get("time"), relevance=v.get("relevance", 1))
        for attr, fn in payload:
            setattr(self, attr, fn(v))

    def __repr__(self):
        return BaseMessage.__repr__(self) + (tail(self) if tail else "")

    __init__.__signature__ = sig
    return type(name, (BaseMessage,), {"__init__": __init__, "__repr__": __repr__, "__module__": __name__,
                                       "__doc__": f"{mtype.name} message ({where})"})


def _arg(name):
    return lambda v: v[name]


def _ids(objs):
    return [o.id for o in objs] if objs else []


_T = MessageType
LaunchMissileMessage = _message_class(
    "LaunchMissileMessage", _T.LAUNCH_MISSILE, [("sender_id", _REQ), ("receiver_id", None), ("time", None), ("target", None)],
    [("target", _arg("target"))], where="MissileLauncher -> Missile, Messages.py:14")
CPPLaunchMissileRequestMessage = _message_class(
    "CPPLaunchMissileRequestMessage", _T.LAUNCH_COMMAND,
    [("sender_id", _REQ), ("target", _REQ), ("target_position", _REQ), ("radar_id", _REQ), ("time", None),
     ("receiver_id", None), ("relevance", 1)],
    [("target", _arg("target")), ("target_position", _arg("target_position")), ("radar_id", _arg("radar_id"))],
    lambda m: f", target={m.target}, target_position={m.target_position}, radar_id={m.radar_id}",
    where="CCP -> MissileLauncher, Messages.py:27")
LaunchedMissileMessage = _message_class(
    "LaunchedMissileMessage", _T.LAUNCHED_MISSILE,
    [("sender_id", _REQ), ("missile", _REQ), ("target_id", _REQ), ("time", None), ("receiver_id", None)],
    [("missile", _arg("missile")), ("target_id", _arg("target_id"))],
    lambda m: f", missile.id={m.missile.id if m.missile else None}, target_id={m.target_id}",
    where="MissileLauncher -> CCP, Messages.py:43")
MissileCountRequestMessage = _message_class(
    "MissileCountRequestMessage", _T.MISSILE_COUNT_REQUEST,
    [("sender_id", _REQ), ("time", None), ("receiver_id", None), ("relevance", 3)],
    where="CCP -> MissileLauncher, relevance 3, Messages.py:58")
MissileCountResponseMessage = _message_class(
    "MissileCountResponseMessage", _T.MISSILE_COUNT_RESPONSE,
    [("sender_id", _REQ), ("count", _REQ), ("time", None), ("receiver_id", None)],
    [("count", _arg("count"))], lambda m: f", count={m.count}", where="MissileLauncher -> CCP, Messages.py:70")
AllObjectsMessage = _message_class(
    "AllObjectsMessage", _T.ALL_OBJECTS, [("sender_id", _REQ), ("objects", _REQ), ("time", None), ("receiver_id", None)],
    [("objects", _arg("objects"))], where="Radar -> CCP, Messages.py:84")
FoundObjectsMessage = _message_class(
    "FoundObjectsMessage", _T.FOUND_OBJECTS,
    [("sender_id", _REQ), ("visible_objects", _REQ), ("time", None), ("receiver_id", None)],
    [("visible_objects", _arg("visible_objects"))],
    lambda m: f", visible_objects.ids={_ids(m.visible_objects)}, visible_objects.count={len(_ids(m.visible_objects))}",
    where="Radar -> CCP, Messages.py:94")
CPPUpdateTargetRadarMessage = _message_class(
    "CPPUpdateTargetRadarMessage", _T.CCP_UPDATE_TARGET,
    [("sender_id", _REQ), ("target", _REQ), ("missile_id", _REQ), ("time", None), ("receiver_id", None)],
    [("target", _arg("target")), ("missile_id", _arg("missile_id"))],
    lambda m: f", target.id={m.target.id if m.target else None}, missile_id={m.missile_id}",
    where="CCP -> Radar, read by nobody, Messages.py:109")
ActiveObjectsMessage = _message_class(
    "ActiveObjectsMessage", _T.ACTIVE_OBJECTS,
    [("sender_id", _REQ), ("active_objects", _REQ), ("time", None), ("receiver_id", None)],
    [("active_objects", _arg("active_objects"))],
    lambda m: f", active_objects.ids={_ids(m.active_objects)}, active_objects.count={len(_ids(m.active_objects))}",
    where="AirEnv -> Radar, Messages.py:124")
UpdateTargetPosition = _message_class(
    "UpdateTargetPosition", _T.UPDATE_TARGET, [("sender_id", _REQ), ("upd_object", _REQ), ("time", None), ("receiver_id", None)],
    [("upd_object", _arg("upd_object"))], where="Radar -> Missile, never emitted (SURVEY 5.9#2), Messages.py:139")
DestroyedMissileId = _message_class(
    "DestroyedMissileId", _T.DESTROYED_MISSILE,
    [("sender_id", _REQ), ("missile_id", _REQ), ("time", None), ("receiver_id", None), ("self_detonation", False)],
    # the reference's trailing comma makes this a 1-tuple (Messages.py:160)
    [("missile_id", lambda v: (v["missile_id"],)), ("self_detonation", _arg("self_detonation"))],
    lambda m: f", missile_id={m.missile_id}", where="Radar -> CCP, Messages.py:153")
MissileDetonateMessage = _message_class(
    "MissileDetonateMessage", _T.MISSILE_DETONATE, [("sender_id", _REQ), ("target_id", None), ("self_detonation", False)],
    [("missile_id", _arg("sender_id")), ("target_id", _arg("target_id")), ("self_detonation", _arg("self_detonation"))],
    lambda m: f", missile_id={m.missile_id}, target_id={m.target_id}", where="Missile -> AirEnv, Radar, Messages.py:168")
MissilePosMessage = _message_class(
    "MissilePosMessage", _T.MISSILE_POS, [("sender_id", _REQ)], [("missile_id", _arg("sender_id"))],
    lambda m: f", missile_id={m.missile_id}", where="Missile -> AirEnv, Messages.py:184")
MissileSuccessfulLaunchMessage = _message_class(
    "MissileSuccessfulLaunchMessage", _T.LAUNCH_SUCCESSFUL,
    [("sender_id", _REQ), ("launch_time", _REQ), ("target", _REQ), ("missile", _REQ), ("receiver_id", _REQ)],
    [("missile", _arg("missile")), ("launch_time", _arg("launch_time")), ("target_id", lambda v: v["target"].id)],
    lambda m: f", missile_id={m.missile.id}, target_id={m.target_id}, launch_time={m.launch_time}",
    where="Missile -> MissileLauncher, Messages.py:198")
MissileLaunchCancelledMessage = _message_class(
    "MissileLaunchCancelledMessage", _T.LAUNCH_CANCELLED,
    [("sender_id", _REQ), ("reason", _REQ), ("missile", _REQ), ("receiver_id", _REQ)],
    [("missile", _arg("missile")), ("reason", _arg("reason"))],
    lambda m: f", missile_id={m.missile.id}, reason=\"{m.reason}\"", where="Missile -> MissileLauncher, Messages.py:214")
CPPDrawerObjectsMessage = _message_class(
    "CPPDrawerObjectsMessage", _T.DRAW_OBJECTS,
    [("sender_id", _REQ), ("obj_id", _REQ), ("target_type", _REQ), ("coordinates", _REQ), ("is_visible_by_radar", _REQ),
     ("time", None), ("receiver_id", None)],
    [("obj_id", _arg("obj_id")), ("target_type", _arg("target_type")), ("coordinates", _arg("coordinates")),
     ("is_visible_by_radar", _arg("is_visible_by_radar"))],
    lambda m: (f", obj_id={m.obj_id}, target_type={m.target_type}, coordinates={m.coordinates}, "
               f"is_visible_by_radar={m.is_visible_by_radar}"), where="CCP -> GUI, Messages.py:229")
MissileToAirEnvMessage = _message_class(
    "MissileToAirEnvMessage", _T.NEW_MISSILE, [("sender_id", _REQ), ("missile", _REQ), ("time", None), ("receiver_id", None)],
    [("missile", _arg("missile"))], lambda m: f", missile.id={m.missile.id if m.missile else None}",
    where="MissileLauncher -> AirEnv, Messages.py:245")


def _none(v: int) -> Optional[int]:
    return None if v == _cabi.ZRK_NONE else int(v)


def materialise(rec, ids, time_ms: int, objects) -> BaseMessage:
    """One zrk_msg record -> a message object of the reference's class.  `objects` maps an
    air-object id to its AirObjectView; `ids` is the id pool of list-carrying messages."""
    A = _cabi
    ty = int(rec["type"])
    s, r = int(rec["sender"]), _none(int(rec["receiver"]))
    a, b = int(rec["a"]), int(rec["b"])
    obj = lambda i: objects[i] if i in objects else AirObjectView(i, False)
    if ty == A.MSG_MISSILE_POS:
        m = MissilePosMessage(sender_id=s)
    elif ty == A.MSG_MISSILE_DETONATE:
        m = MissileDetonateMessage(sender_id=s, target_id=_none(b), self_detonation=bool(rec["flag"]))
    elif ty == A.MSG_ACTIVE_OBJECTS:
        m = ActiveObjectsMessage(sender_id=s, active_objects=[obj(int(i)) for i in ids[a:a + b]])
    elif ty == A.MSG_ALL_OBJECTS:
        m = AllObjectsMessage(time=time_ms, sender_id=s, receiver_id=r, objects=[obj(int(i)) for i in ids[a:a + b]])
    elif ty == A.MSG_FOUND_OBJECTS:
        m = FoundObjectsMessage(time=time_ms, sender_id=s, receiver_id=r, visible_objects=[obj(int(i)) for i in ids[a:a + b]])
    elif ty == A.MSG_DESTROYED_MISSILE:
        m = DestroyedMissileId(time=time_ms, sender_id=s, receiver_id=r, missile_id=a, self_detonation=bool(rec["flag"]))
    elif ty == A.MSG_LAUNCH_MISSILE:
        m = LaunchMissileMessage(receiver_id=r, sender_id=s, target=obj(a))
    elif ty == A.MSG_LAUNCH_SUCCESSFUL:
        m = MissileSuccessfulLaunchMessage(sender_id=s, launch_time=time_ms / 1000, target=obj(b), missile=obj(a), receiver_id=r)
    elif ty == A.MSG_LAUNCH_CANCELLED:
        m = MissileLaunchCancelledMessage(sender_id=s, reason=CANCEL_REASONS[b], missile=obj(a), receiver_id=r)
    elif ty == A.MSG_LAUNCHED_MISSILE:
        m = LaunchedMissileMessage(sender_id=s, receiver_id=r, missile=obj(a), target_id=b)
    elif ty == A.MSG_NEW_MISSILE:
        m = MissileToAirEnvMessage(sender_id=s, missile=obj(a))
    elif ty == A.MSG_MISSILE_COUNT_REQUEST:
        m = MissileCountRequestMessage(time=time_ms, sender_id=s, receiver_id=r)
    elif ty == A.MSG_MISSILE_COUNT_RESPONSE:
        m = MissileCountResponseMessage(time=time_ms, sender_id=s, receiver_id=r, count=b)
    elif ty == A.MSG_LAUNCH_COMMAND:
        m = CPPLaunchMissileRequestMessage(time=time_ms, sender_id=s, receiver_id=r, target=obj(a),
                                           target_position=np.array(rec["xyz"], dtype=np.float64), radar_id=b)
    elif ty == A.MSG_CCP_UPDATE_TARGET:
        m = CPPUpdateTargetRadarMessage(time=time_ms, sender_id=s, receiver_id=r, target=obj(a), missile_id=_none(b))
    elif ty == A.MSG_DRAW_OBJECTS:
        m = CPPDrawerObjectsMessage(time=time_ms, sender_id=s, receiver_id=r, obj_id=a,
                                    target_type=ZUR if b == 3 else TARGET_TYPES[b],
                                    coordinates=np.array(rec["xyz"], dtype=np.float64),
                                    is_visible_by_radar=bool(rec["flag"]))
    else:
        raise ValueError(f"unknown message type {ty}")
    m.send_time = time_ms  # Manager.add_message fills it (Manager.py:53-54)
    return m
