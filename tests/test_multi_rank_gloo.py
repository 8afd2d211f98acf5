"""The N>1 path on the CPU: world_size 2 over gloo.  Each rank owns a contiguous block of the
scenarios, runs them (through the emulated build: host logic only) with no per-tick exchange, and
the final statistics all-reduce gives the single-rank totals."""
import os
import socket
import sys

import numpy as np
import pytest

from zrk_modulation_b200.distributed import describe_stats, gather_stats, local_stats, shard_range

HERE = os.path.dirname(os.path.abspath(__file__))
N_SCEN, TICKS = 6, 30


def test_shard_range_partitions():
    for n in (1, 5, 8, 131072, 1048576):
        for w in (1, 2, 3, 4, 8):
            blocks = [shard_range(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _batch(s0, s1):
    from zrk_modulation_b200 import synthetic_batch
    return synthetic_batch(s1 - s0, 10, n_radars=1, n_launchers=2, missiles_per_launcher=2, seed0=s0, noise_period=8)


def _run(batch):
    import emu_lib
    from zrk_modulation_b200.engine import Engine
    with Engine(batch, lib=emu_lib.load()) as eng:
        eng.step(TICKS)
        fc, _, _ = eng.status()
        return local_stats(eng.counts(), fc)


def _worker(rank, world, port, out):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s0, s1 = shard_range(N_SCEN, rank, world)
    total = gather_stats(_run(_batch(s0, s1)))
    dist.barrier()
    if rank == 0:
        np.save(out, total)
    dist.destroy_process_group()


def test_two_ranks_equal_one(tmp_path):
    import emu_lib
    emu_lib.load()  # build once before forking workers
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "stats.npy")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    two = np.load(out)
    one = _run(_batch(0, N_SCEN))
    assert np.array_equal(one, two)
    d = describe_stats(two)
    assert d["scenarios"] == N_SCEN and d["messages"]["ACTIVE_OBJECTS"] <= N_SCEN * TICKS
