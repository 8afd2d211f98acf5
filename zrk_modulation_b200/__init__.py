"""zrk_modulation_b200 — B200-native batched tick of Ollegorii/ZRK_modulation.

The hot path (Manager.run_simulation, modules/Manager.py:111-140, and every step() it calls)
runs as hand-written sm_100a CUDA behind the C ABI of include/zrk_b200.h.  This package is the
Python host side: the scenario compiler (reference YAML schema) and the mirror of the
reference's BaseModel / Manager / message interface.  There is no CPU fallback.
"""
from ._cabi import (ZRK_NONE, LOG_EVENTS, LOG_BULK, LOG_MASKS, LOG_FULL, MSG_NAMES)  # noqa: F401
from .scenario import (Batch, compile_batch, load_config, make_noise, synthetic_batch,  # noqa: F401
                       batch_to_config, perturbed_batch)

from .messages import MessageType, TargetType  # noqa: F401,E402
from .manager import (BaseModel, BatchedManager, BatchedTick, Timer, create_objects_from_config,  # noqa: F401,E402
                      run_simulation_from_config)
from .engine import Engine, ZrkError  # noqa: F401,E402

__all__ = ["MessageType", "TargetType", "BaseModel", "BatchedManager", "BatchedTick", "Timer",
           "create_objects_from_config", "run_simulation_from_config", "Engine", "ZrkError","Batch", "compile_batch", "load_config", "make_noise", "synthetic_batch",
           "batch_to_config", "perturbed_batch", "ZRK_NONE", "LOG_EVENTS", "LOG_BULK", "LOG_MASKS", "LOG_FULL",
           "MSG_NAMES"]
