"""ctypes mirror of include/zrk_b200.h and the loader of the CUDA library.

The library is the product: there is no CPU fallback.  `load()` raises if
`libzrk_b200.so` has not been built (``python -c "import __graft_entry__ as g; g.build()"``).
"""
from __future__ import annotations

import ctypes as C
import os

ZRK_ABI_VERSION = 1
ZRK_NONE = -(2 ** 31)
ZRK_MSG_NTYPES = 18

# zrk_msg_type, declaration order of modules/constants.py:3-21
(MSG_LAUNCH_MISSILE, MSG_LAUNCHED_MISSILE, MSG_LAUNCH_COMMAND, MSG_MISSILE_COUNT_REQUEST,
 MSG_MISSILE_COUNT_RESPONSE, MSG_ALL_OBJECTS, MSG_FOUND_OBJECTS, MSG_ACTIVE_OBJECTS,
 MSG_CCP_UPDATE_TARGET, MSG_UPDATE_TARGET, MSG_MISSILE_GET_HIT, MSG_DESTROYED_MISSILE,
 MSG_DRAW_OBJECTS, MSG_MISSILE_POS, MSG_MISSILE_DETONATE, MSG_NEW_MISSILE, MSG_LAUNCH_CANCELLED,
 MSG_LAUNCH_SUCCESSFUL) = range(18)

MSG_NAMES = ["LAUNCH_MISSILE", "LAUNCHED_MISSILE", "LAUNCH_COMMAND", "MISSILE_COUNT_REQUEST",
             "MISSILE_COUNT_RESPONSE", "ALL_OBJECTS", "FOUND_OBJECTS", "ACTIVE_OBJECTS",
             "CCP_UPDATE_TARGET", "UPDATE_TARGET", "MISSILE_GET_HIT", "DESTROYED_MISSILE",
             "DRAW_OBJECTS", "MISSILE_POS", "MISSILE_DETONATE", "NEW_MISSILE", "LAUNCH_CANCELLED",
             "LAUNCH_SUCCESSFUL"]


def _bit(t):
    return 1 << t


LOG_EVENTS = (_bit(MSG_LAUNCH_MISSILE) | _bit(MSG_LAUNCHED_MISSILE) | _bit(MSG_LAUNCH_COMMAND) |
              _bit(MSG_MISSILE_COUNT_REQUEST) | _bit(MSG_MISSILE_COUNT_RESPONSE) |
              _bit(MSG_DESTROYED_MISSILE) | _bit(MSG_MISSILE_DETONATE) | _bit(MSG_NEW_MISSILE) |
              _bit(MSG_LAUNCH_CANCELLED) | _bit(MSG_LAUNCH_SUCCESSFUL))
LOG_BULK = _bit(MSG_CCP_UPDATE_TARGET) | _bit(MSG_DRAW_OBJECTS) | _bit(MSG_MISSILE_POS)
LOG_MASKS = 0x80000000
LOG_FULL = LOG_EVENTS | LOG_BULK | LOG_MASKS

FAULT_NONE, FAULT_ATTRIBUTE_ERROR_CCP259, FAULT_KEY_ERROR_CCP107 = 0, 1, 2
FAULT_LOG_OVERFLOW = 64

OK, ERR_ARG, ERR_CUDA, ERR_NO_DEVICE, ERR_NOMEM, ERR_UNSUPPORTED, ERR_CAPACITY = 0, -1, -2, -3, -4, -5, -6

_pi32 = C.POINTER(C.c_int32)
_pf64 = C.POINTER(C.c_double)
_pu8 = C.POINTER(C.c_uint8)


class ZrkMsg(C.Structure):
    _fields_ = [("tick", C.c_int32), ("type", C.c_uint8), ("relevance", C.c_uint8),
                ("flag", C.c_uint8), ("pad", C.c_uint8), ("sender", C.c_int32),
                ("receiver", C.c_int32), ("a", C.c_int32), ("b", C.c_int32),
                ("xyz", C.c_double * 3)]


assert C.sizeof(ZrkMsg) == 48

# numpy view of zrk_msg
MSG_DTYPE = [("tick", "<i4"), ("type", "u1"), ("relevance", "u1"), ("flag", "u1"), ("pad", "u1"),
             ("sender", "<i4"), ("receiver", "<i4"), ("a", "<i4"), ("b", "<i4"), ("xyz", "<f8", (3,))]


class ZrkBatchDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("n_scenarios", C.c_int32), ("max_targets", C.c_int32),
        ("max_radars", C.c_int32), ("max_launchers", C.c_int32), ("max_missiles", C.c_int32),
        ("max_ccp_launchers", C.c_int32), ("dt_ms", C.c_int32),
        ("n_targets", _pi32), ("n_radars", _pi32), ("n_launchers", _pi32), ("n_missiles", _pi32),
        ("n_ccp_launchers", _pi32), ("airenv_id", _pi32), ("ccp_id", _pi32), ("tick_limit", _pi32),
        ("target_id", _pi32), ("target_type", _pi32), ("target_pos", _pf64), ("target_vel", _pf64),
        ("radar_id", _pi32), ("radar_pos", _pf64), ("radar_par", _pf64),
        ("launcher_id", _pi32), ("launcher_pos", _pf64),
        ("missile_id", _pi32), ("missile_launcher", _pi32), ("missile_v0", _pf64),
        ("missile_v0sq", _pf64), ("missile_radius", _pf64), ("missile_life", _pf64),
        ("ccp_launcher_id", _pi32), ("ccp_launcher_pos", _pf64),
        ("noise", _pf64), ("noise_period", C.c_int32), ("noise_radars", C.c_int32),
        ("noise_objects", C.c_int32), ("noise_scenario_stride", C.c_int64),
    ]


OPT_EXACT_ONLY = 1
OPT_NO_CERTIFICATE = 2


class ZrkOptions(C.Structure):
    _fields_ = [("device", C.c_int32), ("log_types", C.c_uint32), ("log_capacity", C.c_int32),
                ("mask_ticks", C.c_int32), ("ticks_per_launch", C.c_int32),
                ("flags", C.c_uint32), ("waves", C.c_int32), ("wave_rounds", C.c_int32)]


STATE_FIELDS = [
    ("slot_obj", "i4", "NS"), ("slot_alive", "u1", "NS"),
    ("pos", "f8", "NS3"), ("prev_pos", "f8", "NS3"), ("has_prev", "u1", "NS"),
    ("radar_az", "f8", "R"), ("radar_el", "f8", "R"),
    ("missile_status", "u1", "M"), ("missile_period", "f8", "M"), ("missile_vel", "f8", "M3"),
    ("missile_t0", "f8", "M"), ("missile_target", "i4", "M"),
    ("launcher_count", "i4", "L"),
    ("ccp_n_targets", "i4", "1"), ("ccp_n_missiles", "i4", "1"),
    ("ccp_t_key", "i4", "NS"), ("ccp_t_ref", "i4", "NS"), ("ccp_t_upd", "i4", "NS"),
    ("ccp_t_missile", "i4", "NS"), ("ccp_t_following", "u1", "NS"),
    ("ccp_m_key", "i4", "NS"), ("ccp_m_ref", "i4", "NS"), ("ccp_m_upd", "i4", "NS"),
    ("ccp_m_target", "i4", "NS"),
    ("ccp_capacity", "i4", "C"), ("ccp_launched", "i4", "C"),
]
_CT = {"i4": _pi32, "f8": _pf64, "u1": _pu8}


class ZrkStateView(C.Structure):
    _fields_ = [(name, _CT[dt]) for name, dt, _ in STATE_FIELDS]


def state_shapes(batch_dims):
    """Shapes of the zrk_state_view arrays for given (B, NS, R, M, L, C)."""
    B, NS, R, M, L, Cc = batch_dims
    dims = {"NS": (B, NS), "NS3": (B, NS, 3), "R": (B, R), "M": (B, M), "M3": (B, M, 3),
            "L": (B, L), "1": (B,), "C": (B, Cc)}
    return {name: (dims[k], dt) for name, dt, k in STATE_FIELDS}


def alloc_state(batch_dims):
    """Allocate numpy buffers + the ctypes view pointing at them."""
    import numpy as np
    arrays, view = {}, ZrkStateView()
    for name, (shape, dt) in state_shapes(batch_dims).items():
        shape = tuple(max(1, s) for s in shape)
        a = np.zeros(shape, dtype=dt)
        arrays[name] = a
        setattr(view, name, a.ctypes.data_as(_CT[dt]))
    return arrays, view


_LIB = None
LIB_NAME = "libzrk_b200.so"


def lib_path() -> str:
    if os.environ.get("ZRK_B200_LIB"):  # development: A/B a differently tuned build of the same sources
        return os.environ["ZRK_B200_LIB"]
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", LIB_NAME)


def load():
    """dlopen libzrk_b200.so and declare the prototypes of include/zrk_b200.h."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build the CUDA library first (__graft_entry__.build()). "
            "zrk_modulation_b200 has no CPU fallback.")
    lib = declare(C.CDLL(path))
    if lib.zrk_abi_version() != ZRK_ABI_VERSION:
        raise RuntimeError("libzrk_b200.so ABI version mismatch; rebuild")
    _LIB = lib
    return lib


def declare(lib):
    """Attach the prototypes of include/zrk_b200.h to a loaded library."""
    H = C.c_void_p
    lib.zrk_create.argtypes = [C.POINTER(ZrkBatchDesc), C.POINTER(ZrkOptions), C.POINTER(H)]
    lib.zrk_step.argtypes = [H, C.c_int]
    lib.zrk_sync.argtypes = [H]
    lib.zrk_get_time.argtypes = [H, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.zrk_read_status.argtypes = [H, _pi32, _pi32, _pi32]
    lib.zrk_read_counts.argtypes = [H, C.POINTER(C.c_int64)]
    lib.zrk_read_activity.argtypes = [H, C.POINTER(C.c_int64)]
    lib.zrk_read_totals.argtypes = [H, C.POINTER(C.c_int64)]
    lib.zrk_read_certified.argtypes = [H, C.POINTER(C.c_int64)]
    lib.zrk_read_records.argtypes = [H, C.c_int, C.c_int, C.POINTER(ZrkMsg), _pi32]
    lib.zrk_log_capacity.argtypes = [H]
    lib.zrk_read_messages.argtypes = [H, C.c_int, C.c_int, C.c_int, C.POINTER(ZrkMsg), C.c_size_t,
                                      C.POINTER(C.c_size_t), _pi32, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.zrk_read_state.argtypes = [H, C.POINTER(ZrkStateView)]
    lib.zrk_state_bytes.argtypes = [H, C.POINTER(C.c_size_t)]
    lib.zrk_save_state.argtypes = [H, C.c_void_p, C.c_size_t]
    lib.zrk_load_state.argtypes = [H, C.c_void_p, C.c_size_t]
    lib.zrk_device_ptr.argtypes = [H, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    lib.zrk_launch_stats.argtypes = [H, C.POINTER(C.c_int64), C.POINTER(C.c_double)]
    lib.zrk_last_error.argtypes = [H]
    lib.zrk_last_error.restype = C.c_char_p
    lib.zrk_destroy.argtypes = [H]
    lib.zrk_destroy.restype = None
    lib.zrk_abi_version.argtypes = []
    lib.zrk_release_cached_memory.argtypes = []
    for f in ("zrk_create", "zrk_step", "zrk_sync", "zrk_get_time", "zrk_read_status",
              "zrk_read_counts", "zrk_read_totals", "zrk_read_activity", "zrk_read_certified", "zrk_read_records", "zrk_log_capacity", "zrk_read_messages",
              "zrk_read_state", "zrk_device_ptr", "zrk_launch_stats", "zrk_abi_version", "zrk_state_bytes",
              "zrk_save_state", "zrk_load_state", "zrk_release_cached_memory"):
        getattr(lib, f).restype = C.c_int
    return lib


EXPORTED_SYMBOLS = ["zrk_create", "zrk_step", "zrk_sync", "zrk_get_time", "zrk_read_status",
                    "zrk_read_counts", "zrk_read_totals", "zrk_read_activity", "zrk_read_certified", "zrk_read_records", "zrk_log_capacity", "zrk_read_messages",
                    "zrk_read_state", "zrk_device_ptr", "zrk_launch_stats", "zrk_last_error",
                    "zrk_destroy", "zrk_abi_version", "zrk_state_bytes", "zrk_save_state", "zrk_load_state",
                    "zrk_release_cached_memory"]
