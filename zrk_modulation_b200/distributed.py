"""Multi-GPU plumbing: scenarios are independent (each is its own Manager object graph,
main.py:35-149), so a batch shards into contiguous blocks of scenarios, one block per rank, with
no per-tick exchange.  The only collective is the final all-reduce of the statistics vector
(NCCL on GPUs; gloo in the CPU tests)."""
from __future__ import annotations

from typing import Tuple

import numpy as np

from . import _cabi

STATS_LEN = _cabi.ZRK_MSG_NTYPES + 4  # per-type message totals, faulted, attribute-faults, key-faults, scenarios


def shard_range(n_scenarios: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [s0, s1) of scenarios owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_scenarios, world)
    s0 = rank * base + min(rank, rem)
    return s0, s0 + base + (1 if rank < rem else 0)


def local_stats(counts: np.ndarray, fault_code: np.ndarray) -> np.ndarray:
    """This rank's statistics vector (int64[STATS_LEN])."""
    fc = fault_code & 63
    return np.concatenate([counts.sum(axis=0), [(fc != 0).sum(), (fc == _cabi.FAULT_ATTRIBUTE_ERROR_CCP259).sum(),
                                                (fc == _cabi.FAULT_KEY_ERROR_CCP107).sum(), counts.shape[0]]]).astype(np.int64)


def local_stats_device(summary):
    """The same statistics vector computed on the GPU from Engine.torch_view("summary")
    (int32 [B, 16 + 18]: header words, then the per-type counters): no host hop before NCCL."""
    import torch
    from .engine import Engine  # noqa: F401  (documentation pointer only)
    H_FAULT = 5  # zrk_layout.h
    counts = summary[:, 16:].to(torch.int64).sum(dim=0)
    fc = summary[:, H_FAULT] & 63
    extra = torch.stack([(fc != 0).sum(), (fc == _cabi.FAULT_ATTRIBUTE_ERROR_CCP259).sum(),
                         (fc == _cabi.FAULT_KEY_ERROR_CCP107).sum(),
                         torch.tensor(summary.shape[0], device=summary.device)]).to(torch.int64)
    return torch.cat([counts, extra])


def gather_stats(stats, device=None) -> np.ndarray:
    """Sum the statistics vectors of all ranks (torch.distributed must be initialised when
    world_size > 1)."""
    import torch
    import torch.distributed as dist
    t = stats.clone() if torch.is_tensor(stats) else torch.from_numpy(np.array(stats, copy=True))
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t.cpu().numpy()
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def describe_stats(stats: np.ndarray) -> dict:
    n = _cabi.ZRK_MSG_NTYPES
    return {"messages_total": int(stats[:n].sum()),
            "messages": {name: int(stats[i]) for i, name in enumerate(_cabi.MSG_NAMES) if stats[i]},
            "faulted_scenarios": int(stats[n]), "attribute_error_ccp259": int(stats[n + 1]),
            "key_error_ccp107": int(stats[n + 2]), "scenarios": int(stats[n + 3])}
