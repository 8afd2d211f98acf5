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


class LaunchMissileMessage(BaseMessage):  # Messages.py:14
    def __init__(self, sender_id: int, receiver_id: int = None, time: int = None, target=None):
        super().__init__(type=MessageType.LAUNCH_MISSILE, send_time=time, sender_id=sender_id, receiver_id=receiver_id)
        self.target = target


class CPPLaunchMissileRequestMessage(BaseMessage):  # Messages.py:27
    def __init__(self, sender_id: int, target, target_position: np.ndarray, radar_id: int, time: int = None,
                 receiver_id: int = None, relevance: int = 1):
        super().__init__(type=MessageType.LAUNCH_COMMAND, send_time=time, sender_id=sender_id,
                         receiver_id=receiver_id, relevance=relevance)
        self.target = target
        self.target_position = target_position
        self.radar_id = radar_id

    def __repr__(self) -> str:
        return f"{super().__repr__()}, target={self.target}, target_position={self.target_position}, radar_id={self.radar_id}"


class LaunchedMissileMessage(BaseMessage):  # Messages.py:43
    def __init__(self, sender_id: int, missile, target_id: int, time: int = None, receiver_id: int = None):
        super().__init__(type=MessageType.LAUNCHED_MISSILE, send_time=time, sender_id=sender_id, receiver_id=receiver_id)
        self.missile = missile
        self.target_id = target_id

    def __repr__(self) -> str:
        return f"{super().__repr__()}, missile.id={self.missile.id if self.missile else None}, target_id={self.target_id}"


class MissileCountRequestMessage(BaseMessage):  # Messages.py:58
    def __init__(self, sender_id: int, time: int = None, receiver_id: int = None, relevance: int = 3):
        super().__init__(type=MessageType.MISSILE_COUNT_REQUEST, send_time=time, sender_id=sender_id,
                         receiver_id=receiver_id, relevance=relevance)


class MissileCountResponseMessage(BaseMessage):  # Messages.py:70
    def __init__(self, sender_id: int, count: int, time: int = None, receiver_id: int = None):
        super().__init__(type=MessageType.MISSILE_COUNT_RESPONSE, send_time=time, sender_id=sender_id, receiver_id=receiver_id)
        self.count = count

    def __repr__(self) -> str:
        return f"{super().__repr__()}, count={self.count}"


class AllObjectsMessage(BaseMessage):  # Messages.py:84
    def __init__(self, sender_id: int, objects: List, time: int = None, receiver_id: int = None):
        super().__init__(type=MessageType.ALL_OBJECTS, send_time=time, sender_id=sender_id, receiver_id=receiver_id)
        self.objects = objects


class FoundObjectsMessage(BaseMessage):  # Messages.py:94
    def __init__(self, sender_id: int, visible_objects: List, time: int = None, receiver_id: int = None):
        super().__init__(type=MessageType.FOUND_OBJECTS, send_time=time, sender_id=sender_id, receiver_id=receiver_id)
        self.visible_objects = visible_objects

    def __repr__(self) -> str:
        ids = [obj.id for obj in self.visible_objects] if self.visible_objects else []
        return f"{super().__repr__()}, visible_objects.ids={ids}, visible_objects.count={len(ids)}"


class CPPUpdateTargetRadarMessage(BaseMessage):  # Messages.py:109
    def __init__(self, sender_id: int, target, missile_id: int, time: int = None, receiver_id: int = None):
        super().__init__(type=MessageType.CCP_UPDATE_TARGET, send_time=time, sender_id=sender_id, receiver_id=receiver_id)
        self.target = target
        self.missile_id = missile_id

    def __repr__(self) -> str:
        return f"{super().__repr__()}, target.id={self.target.id if self.target else None}, missile_id={self.missile_id}"


class ActiveObjectsMessage(BaseMessage):  # Messages.py:124
    def __init__(self, sender_id: int, active_objects: List, time: int = None, receiver_id: int = None):
        super().__init__(type=MessageType.ACTIVE_OBJECTS, send_time=time, sender_id=sender_id, receiver_id=receiver_id)
        self.active_objects = active_objects

    def __repr__(self) -> str:
        ids = [obj.id for obj in self.active_objects] if self.active_objects else []
        return f"{super().__repr__()}, active_objects.ids={ids}, active_objects.count={len(ids)}"


class UpdateTargetPosition(BaseMessage):  # Messages.py:139 (never emitted: SURVEY 5.9#2)
    def __init__(self, sender_id: int, upd_object, time: int = None, receiver_id: int = None):
        super().__init__(type=MessageType.UPDATE_TARGET, send_time=time, sender_id=sender_id, receiver_id=receiver_id)
        self.upd_object = upd_object


class DestroyedMissileId(BaseMessage):  # Messages.py:153
    def __init__(self, sender_id: int, missile_id: int, time: int = None, receiver_id: int = None,
                 self_detonation: bool = False):
        super().__init__(type=MessageType.DESTROYED_MISSILE, send_time=time, sender_id=sender_id, receiver_id=receiver_id)
        self.missile_id = missile_id,  # the reference's trailing comma: a 1-tuple (Messages.py:160)
        self.self_detonation = self_detonation

    def __repr__(self) -> str:
        return f"{super().__repr__()}, missile_id={self.missile_id}"


class MissileDetonateMessage(BaseMessage):  # Messages.py:168
    def __init__(self, sender_id: int, target_id: int = None, self_detonation: bool = False):
        super().__init__(type=MessageType.MISSILE_DETONATE, sender_id=sender_id)
        self.missile_id = sender_id
        self.target_id = target_id
        self.self_detonation = self_detonation

    def __repr__(self) -> str:
        return f"{super().__repr__()}, missile_id={self.missile_id}, target_id={self.target_id}"


class MissilePosMessage(BaseMessage):  # Messages.py:184
    def __init__(self, sender_id: int):
        super().__init__(type=MessageType.MISSILE_POS, sender_id=sender_id)
        self.missile_id = sender_id

    def __repr__(self) -> str:
        return f"{super().__repr__()}, missile_id={self.missile_id}"


class MissileSuccessfulLaunchMessage(BaseMessage):  # Messages.py:198
    def __init__(self, sender_id: int, launch_time: float, target, missile, receiver_id: int):
        super().__init__(type=MessageType.LAUNCH_SUCCESSFUL, sender_id=sender_id, receiver_id=receiver_id)
        self.missile = missile
        self.launch_time = launch_time
        self.target_id = target.id

    def __repr__(self) -> str:
        return f"{super().__repr__()}, missile_id={self.missile.id}, target_id={self.target_id}, launch_time={self.launch_time}"


class MissileLaunchCancelledMessage(BaseMessage):  # Messages.py:214
    def __init__(self, sender_id: int, reason: str, missile, receiver_id: int):
        super().__init__(type=MessageType.LAUNCH_CANCELLED, sender_id=sender_id, receiver_id=receiver_id)
        self.missile = missile
        self.reason = reason

    def __repr__(self) -> str:
        return f"{super().__repr__()}, missile_id={self.missile.id}, reason=\"{self.reason}\""


class CPPDrawerObjectsMessage(BaseMessage):  # Messages.py:229
    def __init__(self, sender_id: int, obj_id: int, target_type, coordinates: np.ndarray, is_visible_by_radar: bool,
                 time: int = None, receiver_id: int = None):
        super().__init__(type=MessageType.DRAW_OBJECTS, send_time=time, sender_id=sender_id, receiver_id=receiver_id)
        self.obj_id = obj_id
        self.target_type = target_type
        self.coordinates = coordinates
        self.is_visible_by_radar = is_visible_by_radar

    def __repr__(self) -> str:
        return (f"{super().__repr__()}, obj_id={self.obj_id}, target_type={self.target_type}, "
                f"coordinates={self.coordinates}, is_visible_by_radar={self.is_visible_by_radar}")


class MissileToAirEnvMessage(BaseMessage):  # Messages.py:245
    def __init__(self, sender_id: int, missile, time: int = None, receiver_id: int = None):
        super().__init__(type=MessageType.NEW_MISSILE, send_time=time, sender_id=sender_id, receiver_id=receiver_id)
        self.missile = missile

    def __repr__(self) -> str:
        return f"{super().__repr__()}, missile.id={self.missile.id if self.missile else None}"


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
