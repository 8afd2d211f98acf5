#!/usr/bin/env python
"""bench.py — object-updates/s of the batched tick on N B200s (one process per GPU).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference tick on the host CPU cores

A "step" is one pass of the hot path over one batch: every scenario of this rank's shard is
advanced `ticks` ticks (Manager.run_simulation, Manager.py:111-140) from its initial state.
Workload (default `c5`): BASELINE.json config 5 sharded — 131072 scenarios per GPU x 32 object
slots (24 targets + 2 launchers x 4 missiles, all of which fly), 1 radar, 250 ticks; at 8 GPUs
this is exactly "1M scenarios x 32 objects".  Scenarios are independent, so ranks exchange
nothing per tick; the only collective is the final statistics all-reduce (NCCL).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (scenarios per GPU, targets, radars, launchers, missiles per launcher, ticks)
    "c5": (131072, 24, 1, 2, 4, 250),
    "c3": (4096, 64, 1, 0, 0, 1000),
    "c4": (8192, 256, 2, 0, 0, 200),
    "c2": (1, 1000, 1, 0, 0, 10000),
    "tiny": (256, 24, 1, 2, 4, 50),
}
BYTES_PER_OBJECT_UPDATE = 128  # fp64 layout, one tick per pass (DESIGN.md "Algorithmic bytes")


def describe(name, n_gpus):
    B, T, R, L, mpl, ticks = WORKLOADS[name]
    return {"workload": f"{name}: {B} scenarios/GPU x {T + L * mpl} object slots ({T} targets + {L}x{mpl} missiles), "
                        f"{R} radar(s), {ticks} ticks, dt=200ms; BASELINE config "
                        f"{ {'c5': 5, 'c3': 3, 'c4': 4, 'c2': 2}.get(name, 0) } shape",
            "scenarios_per_gpu": B, "object_slots": T + L * mpl, "ticks": ticks, "radars": R,
            "total_scenarios": B * n_gpus, "parallelism": f"scenario-sharded x{n_gpus}, no per-tick exchange",
            "cache": "state per GPU larger than L2 (inputs larger than L2)" if B * (T + L * mpl) * 128 > 126e6
                     else "state fits L2; L2 flushed between timed steps"}


def make_batch(name, rank, pinned=False):
    from zrk_modulation_b200 import synthetic_batch
    B, T, R, L, mpl, ticks = WORKLOADS[name]
    b = synthetic_batch(B, T, n_radars=R, n_launchers=L, missiles_per_launcher=mpl, seed0=rank * B)
    if pinned:  # page-locked host buffers: the H2D copies of zrk_create run at full PCIe rate
        import torch
        b._pins = []
        for f, v in list(vars(b).items()):
            if isinstance(v, np.ndarray) and v.size and f != "durations":
                t = torch.from_numpy(np.ascontiguousarray(v)).pin_memory()
                setattr(b, f, t.numpy())
                b._pins.append(t)
    return b, ticks


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm = [float(r[1]) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.strip().lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _profile_summaries(workload, ticks_per_launch):
    """Committed ncu summaries (profiles/*_summary.json) of this workload, newest kernel version first."""
    import glob
    out = []
    for p in glob.glob(os.path.join(ROOT, "profiles", "*_summary.json")):
        try:
            d = json.load(open(p))
        except ValueError:
            continue
        if d.get("workload") == workload:  # (the newest kernel version profiled on this workload)
            out.append((d.get("seq", 0), p, d))
    return [(p, d) for _, p, d in sorted(out, key=lambda t: -t[0])]


def profiled(workload, ticks_per_launch, key, what):
    """`key` of the newest committed ncu summary of this workload as a labelled dict, else None."""
    for p, d in _profile_summaries(workload, ticks_per_launch):
        if key in d:
            return {"what": what, key: d[key], "source": os.path.relpath(p, ROOT)}
    return None


def profiled_traffic(workload, ticks_per_launch):
    """dram__bytes_read.sum + dram__bytes_write.sum of one tick-kernel launch, from the newest committed
    `ncu --set full` summary of the same workload (profiles/*_summary.json), else None."""
    for _, d in _profile_summaries(workload, ticks_per_launch):
        if d.get("dram_bytes_per_launch") is not None:
            return d["dram_bytes_per_launch"]
    return None


def cpu_reference_throughput(name, n_scen_per_core, cores, ticks=None):
    """The reference tick on the host: oracle/zrk_oracle.c (the C restatement of the reference,
    pinned bit-exactly to it by tests/golden) on `cores` threads over a bounded sample of the
    same synthetic scenarios.  Returns object-updates/s."""
    from oracle import oracle as O
    from zrk_modulation_b200 import synthetic_batch
    B, T, R, L, mpl, wticks = WORKLOADS[name]
    ticks = ticks or wticks
    n = max(1, min(B, n_scen_per_core * cores))
    b = synthetic_batch(n, T, n_radars=R, n_launchers=L, missiles_per_launcher=mpl, seed0=0)
    O.load()
    t0 = time.perf_counter()
    O.run_batch(b, 0, n, ticks, n_threads=cores)
    dt = time.perf_counter() - t0
    return n * (T + L * mpl) * ticks / dt, n, ticks, dt


def python_loop(name):
    """The reference's OWN Python tick loop (Manager.run_simulation, Manager.py:111-140) on all host cores.
    Live when the reference tree is present (the build container; $ZRK_REFERENCE_DIR), else the committed
    measurement profiles/ref_python_loop.json (made by profiles/ref_python_loop.py) is cited: the reference
    is pure Python and does not travel to the GPU box."""
    shape = name if name in ("c5", "c3", "c4") else "c5"
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        sys.path.insert(0, os.path.join(ROOT, "profiles"))
        import refharness
        if refharness.available():
            import ref_python_loop
            r = ref_python_loop.measure(shape, per_core=1)
            return {"value": r["object_updates_per_s"], "unit": "object-updates/s", "cores": r["cores"], "kind": "reference",
                    "per_core": r["object_updates_per_s_per_core"], "measured": "live, this host",
                    "sample": f"{r['scenarios']} scenarios x {r['slots']} slots x {r['ticks']} ticks, one process per core"}
    except Exception as e:  # the live leg is best effort
        print(f"python_loop: live measurement failed: {e}", file=sys.stderr)
    p = os.path.join(ROOT, "profiles", "ref_python_loop.json")
    if os.path.exists(p):
        d = json.load(open(p))
        for r in d["runs"]:
            if r["workload"] == shape:
                return {"value": r["object_updates_per_s"], "unit": "object-updates/s", "cores": r["cores"], "kind": "reference",
                        "per_core": r["object_updates_per_s_per_core"],
                        "measured": f"cited from profiles/ref_python_loop.json ({d['host']['where']}, "
                                    f"{d['host']['cpu_count']} cores): the reference is pure Python and is not on this box",
                        "sample": f"{r['scenarios']} scenarios x {r['slots']} slots x {r['ticks']} ticks, one process per core"}
    return None


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    B, T, R, L, mpl, ticks = WORKLOADS[args.workload]
    slots = T + L * mpl
    per_core = max(1, int(2.0e6 / (slots * min(ticks, 250) * (1 + slots / 32.0))))
    sample_ticks = min(ticks, 250)
    vals, times = [], []
    for i in range(args.warmup + args.steps):
        v, n, tk, dt = cpu_reference_throughput(args.workload, per_core, cores, sample_ticks)
        if i >= args.warmup:
            vals.append(v)
            times.append(dt)
    value = float(np.mean(vals))
    line = {"impl": "reference", "metric": "object-updates/sec", "value": value, "unit": "object-updates/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(times) * 1e3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": describe(args.workload, args.gpus),
            "cpu_baseline": {"value": value, "unit": "object-updates/s", "cores": cores, "kind": "port",
                             "sample": f"{n} scenarios x {slots} slots x {tk} ticks per step, C restatement of the "
                                       f"reference tick (oracle/zrk_oracle.c) on {cores} host threads"},
            "e2e": {"value": value, "unit": "object-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "scenario_ticks_per_s": value / slots}
    line["cpu_baseline"]["python_loop"] = python_loop(args.workload)
    print(json.dumps(line), flush=True)


def bind_host_thread(local_rank, local_world):
    """Pins this rank's host thread to its own share of the cores next to its GPU (NVML's CPU affinity of the device =
    the GPU's NUMA node), so that eight ranks feeding eight GPUs neither migrate nor share cores while they enqueue
    the copies and launches of a step.  Returns what was done (for the bench line), or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        near = sorted(c for c in os.sched_getaffinity(0) if (mask[c // 64] >> (c % 64)) & 1)
        if not near:
            return None
        # the ranks whose GPUs share this core set split it evenly
        same = [r for r in range(local_world)
                if list(pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(r), words)) == list(mask)]
        per = max(1, len(near) // max(1, len(same)))
        k = same.index(local_rank) if local_rank in same else 0
        mine = near[k * per:(k + 1) * per] or near
        os.sched_setaffinity(0, mine)
        return {"cores": [mine[0], mine[-1]], "n": len(mine), "of_gpu_node": len(near)}
    except Exception as e:  # no NVML, fewer devices than ranks, a restricted cpuset ...
        return {"error": str(e)[:80]}


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from zrk_modulation_b200 import _cabi
    from zrk_modulation_b200.engine import Engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = local_rank
    host_binding = bind_host_thread(local_rank, world) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    _cabi.load()
    batch, ticks = make_batch(args.workload, rank + args.shard_offset, pinned=True)
    if args.ticks:
        ticks = args.ticks
    B, NS = batch.n_scenarios, batch.n_slots
    units = B * NS * ticks
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # the log the run keeps: event records + per-tick alive / visibility masks + track journal — enough to rebuild
    # Manager.messages of any scenario (zrk_read_messages); --log events keeps the event records only
    def log_kw(n_ticks):
        if args.log == "events":
            return dict(log_types=_cabi.LOG_EVENTS)
        return dict(log_types=_cabi.LOG_EVENTS | _cabi.LOG_MASKS, mask_ticks=n_ticks + 1)

    # ---- device-resident pass: state already in HBM when the timed region starts
    kern_ms, launches = [], 0
    sampler = None
    for i in range(args.warmup + args.steps):
        eng = Engine(batch, device=dev, ticks_per_launch=args.ticks_per_launch, certificate=not args.no_certificate, **log_kw(ticks))
        eng.sync()
        flush.fill_(i & 0xff)  # evict the freshly written state from L2
        if i == 0 and rank == 0:  # same load during warm-up and timed steps: sample across both
            sampler = ClockSampler(dev)
        barrier()
        n0, ms0 = eng.launch_stats()
        eng.step(ticks)
        eng.sync()
        n1, ms1 = eng.launch_stats()  # CUDA events on the launch stream, per launch
        barrier()
        if i >= args.warmup:
            kern_ms.append(ms1 - ms0)
            launches += n1 - n0
        if i == args.warmup + args.steps - 1:
            from zrk_modulation_b200.distributed import local_stats_device
            dev_stats = local_stats_device(eng.torch_view("summary")).clone()  # stays on the GPU for NCCL
            counts = eng.counts()
            fc, ft, nr = eng.status()
            alive_ticks = int(eng.alive_slot_ticks().sum())
            # a scenario the reference would have aborted freezes at its fault tick: only the
            # ticks it really ran count as work
            ran_ticks = np.where((fc & 63) != 0, np.minimum(ft.astype(np.int64) + 1, ticks), ticks)
            units_done = int(NS * ran_ticks.sum())
            cert_ticks = int(eng.certified_ticks().sum())
        eng.close()
    clocks = sampler.stop() if sampler else None
    t_dev = torch.tensor([sum(kern_ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    total_ms = float(t_dev.item())

    # ---- end to end through the C ABI with host buffers: create (H2D) + step + read-back (D2H)
    h2d = sum(getattr(batch, f).nbytes for f in (
        "n_targets", "n_radars", "n_launchers", "n_missiles", "n_ccp_launchers", "airenv_id", "ccp_id", "target_id",
        "target_type", "target_pos", "target_vel", "radar_id", "radar_pos", "radar_par", "launcher_id", "launcher_pos",
        "missile_id", "missile_launcher", "missile_v0", "missile_v0sq", "missile_radius", "missile_life",
        "ccp_launcher_id", "ccp_launcher_pos")) + (batch.noise.nbytes if batch.noise is not None else 0)
    d2h = (_cabi.ZRK_MSG_NTYPES + 4) * 8

    # (many ranks share the host's PCIe fabric: their uploads take longer than the first round, two rounds hide them —
    # measured: 8 GPUs 35.5 -> 34.5 ms, 4 GPUs 33.6 -> 33.3 ms; on 1 or 2 GPUs the smaller launches of a second forked round cost ~1.5 ms)
    wave_rounds = args.wave_rounds or (2 if world >= 4 else 1)

    def e2e_once(src, result="totals"):
        """create (H2D) + step + result read-back (D2H) + destroy; returns (seconds, result)."""
        barrier()
        t0 = time.perf_counter()
        eng = Engine(src, device=dev, ticks_per_launch=args.ticks_per_launch, waves=args.waves, wave_rounds=wave_rounds, certificate=not args.no_certificate,
                     **log_kw(ticks))
        ta = time.perf_counter()
        eng.step(ticks)
        tb = time.perf_counter()
        if result == "totals":
            out = eng.totals()           # batch totals reduced on the device: 22 words
        elif result == "counters":
            out = eng.counts()           # per-scenario message counters (B x 34 words)
        else:
            out = (eng.counts(), eng.records(0, B)[1])  # ... and every scenario's event records
        tc = time.perf_counter()
        eng.close()                      # zrk_destroy is part of the step
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        e2e_parts.append((ta - t0, tb - ta, tc - tb, t1 - tc))  # host-side time in create / step / read / destroy
        barrier()
        return t1 - t0, out

    e2e_parts = []

    e2e_s = []
    for i in range(args.warmup + args.steps):
        dt, tot = e2e_once(batch)
        if i >= args.warmup:
            e2e_s.append(dt)
    e2e_breakdown = [round(float(np.mean([p[k] for p in e2e_parts[args.warmup:]])) * 1e3, 3) for k in range(4)]
    t_e2e = torch.tensor([sum(e2e_s)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    from zrk_modulation_b200.distributed import local_stats as _ls
    assert np.array_equal(tot, _ls(counts, fc)), "end-to-end run disagrees with the device-resident run"
    # less friendly variants of the same step, one run each (after a warm-up): inputs in pageable memory;
    # per-scenario counters as the result; counters + every scenario's event records
    variants = {}
    if not args.no_variants:
        pageable, _ = make_batch(args.workload, rank + args.shard_offset, pinned=False)
        rec_bytes = None
        for name, src, res in (("pageable_inputs", pageable, "totals"), ("per_scenario_counters", batch, "counters"),
                               ("counters_and_event_records", batch, "records")):
            e2e_once(src, res)
            dt, out = e2e_once(src, res)
            tv = torch.tensor([dt], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tv, op=dist.ReduceOp.MAX)
            variants[name] = {"seconds": float(tv.item()),
                              "d2h_bytes_per_step": {"totals": d2h, "counters": B * (16 + _cabi.ZRK_MSG_NTYPES) * 4,
                                                     "records": B * (16 + _cabi.ZRK_MSG_NTYPES) * 4 * 2 + B * 48 * 70}[res]}
        del pageable

    # ---- final statistics gather: the one collective of the path (NCCL all-reduce over NVLink)
    from zrk_modulation_b200.distributed import describe_stats, gather_stats, local_stats
    assert np.array_equal(dev_stats.cpu().numpy(), local_stats(counts, fc)), "device-side statistics differ"
    stats = gather_stats(dev_stats, device=torch.device("cuda", dev))
    alive_all, units_all = (int(x) for x in gather_stats(torch.tensor([alive_ticks, units_done], dtype=torch.int64),
                                                        device=torch.device("cuda", dev)))

    # ---- what the masks cost: the same device-resident step keeping the event records only
    events_only = None
    if not args.no_variants and args.log == "masks":
        ms_l = []
        for i in range(3):
            eng = Engine(batch, device=dev, ticks_per_launch=args.ticks_per_launch, certificate=not args.no_certificate,
                         log_types=_cabi.LOG_EVENTS)
            eng.sync()
            flush.fill_(i & 0xff)
            barrier()
            n0, m0 = eng.launch_stats()
            eng.step(ticks)
            eng.sync()
            n1, m1 = eng.launch_stats()
            barrier()
            if i >= 1:
                ms_l.append(m1 - m0)
            eng.close()
        tv = torch.tensor([sum(ms_l)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tv, op=dist.ReduceOp.MAX)
        events_only = {"ms_per_step": float(tv.item()) / 2, "steps": 2, "warmup": 1}

    # ---- the other BASELINE shapes, short device-resident legs (config 3, config 4; config 2 does not shard: N = 1 only)
    others = {}
    if not args.no_variants and args.workload == "c5":
        for name in ("c3", "c4") + (("c2",) if world == 1 else ()):
            ob, oticks = make_batch(name, rank, pinned=False)
            oB, oNS = ob.n_scenarios, ob.n_slots
            ms_l, osteps = [], 2
            for i in range(1 + osteps):
                eng = Engine(ob, device=dev, certificate=not args.no_certificate, **log_kw(oticks))
                eng.sync()
                flush.fill_(i & 0xff)
                barrier()
                n0, m0 = eng.launch_stats()
                eng.step(oticks)
                eng.sync()
                n1, m1 = eng.launch_stats()
                barrier()
                if i >= 1:
                    ms_l.append(m1 - m0)
                if i == osteps:
                    ofc, oft, _ = eng.status()
                    ocert = int(eng.certified_ticks().sum())
                eng.close()
            ran = np.where((ofc & 63) != 0, np.minimum(oft.astype(np.int64) + 1, oticks), oticks)
            tv = torch.tensor([sum(ms_l), float(oNS * ran.sum())], dtype=torch.float64, device="cuda")
            if world > 1:
                tmax = tv[:1].clone()
                dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
                dist.all_reduce(tv[1:], op=dist.ReduceOp.SUM)
                tv[0] = tmax[0]
            oms, ounits = float(tv[0].item()), float(tv[1].item())
            oval = ounits * osteps / (oms * 1e-3)
            peak, _ = measured_peak_gbs()
            others[name] = {"config": describe(name, world), "value": oval, "unit": "object-updates/s", "steps": osteps, "warmup": 1,
                            "ms_per_step": oms / osteps, "scenario_ticks_per_s": oval / oNS,
                            "certified_tick_fraction": ocert / float(oB * oticks),
                            "roofline": {"bound": "hbm", "achieved": oval / world * BYTES_PER_OBJECT_UPDATE / 1e9, "peak": peak,
                                         "unit": "GB/s", "frac": oval / world * BYTES_PER_OBJECT_UPDATE / 1e9 / peak},
                            "faulted_scenarios_rank0": int(((ofc & 63) != 0).sum())}
            del ob

    if rank == 0:
        value = units_all * args.steps / (total_ms * 1e-3)
        e2e_value = units_all * args.steps / float(t_e2e.item())
        peak, peak_src = measured_peak_gbs()
        per_launch_ms = sum(kern_ms) / max(1, launches)
        ticks_per_launch = ticks * args.steps / max(1, launches)
        achieved = (units_all / world) * (ticks_per_launch / ticks) * BYTES_PER_OBJECT_UPDATE / (per_launch_ms * 1e-3) / 1e9
        line = {
            "metric": "object-updates/sec", "value": value, "unit": "object-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(describe(args.workload, world), ticks_per_launch=ticks_per_launch),
            "scenario_ticks_per_s": value / NS,
            "active_slot_fraction": alive_all / float(units * world),  # alive objects / (scenarios x slots x ticks), SURVEY 8d
            "object_updates_per_step": units_all,  # what `value` counts: slots x ticks really run (faulted scenarios stop early)
            "object_updates_not_run": units * world - units_all,
            "certified_tick_fraction": cert_ticks / float(B * ticks),  # rank 0's shard: ticks whose link_object scan was proven away
            "e2e": {"value": e2e_value, "unit": "object-updates/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": float(t_e2e.item()) / args.steps * 1e3,
                    "waves": args.waves, "wave_rounds": wave_rounds, "host_binding": host_binding,
                    "host_ms": dict(zip(("create", "step", "read", "destroy"), e2e_breakdown)),  # rank 0, where the host thread spends the step
                    "timed_region": "zrk_create from pinned host arrays (H2D) + zrk_step + zrk_read_totals (D2H) + zrk_destroy",
                    "variants": {k: dict(v, value=units_all / v["seconds"], ms_per_step=v["seconds"] * 1e3)
                                 for k, v in variants.items()}},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": profiled_traffic(args.workload, ticks_per_launch), "peak_source": peak_src,
                         "kernel": "zrk_streak_kernel + zrk_tick_kernel (two-kernel stepping: all launches of the step)",
                         "bytes_per_object_update": BYTES_PER_OBJECT_UPDATE,
                         "real_bound": profiled(args.workload, ticks_per_launch, "issue_active_pct",
                                                "instruction issue / dependent latency, not HBM"),
                         "note": "algorithmic bytes of a one-tick-per-pass fp64 layout x object-updates per launch; "
                                 "a launch fuses ticks_per_launch ticks with state in shared memory, so real DRAM "
                                 "traffic is far lower (see profiles/)"},
            "stats": describe_stats(stats),
        }
        # Where the kernel really lives: instruction issue.  Warp instructions per scenario-tick (ncu, committed
        # summary of this workload) x scenario-ticks/s per GPU against SMs x 4 schedulers x the SM clock under load.
        wi = profiled(args.workload, ticks_per_launch, "warp_instructions_per_scenario_tick", "ncu smsp__inst_executed.sum / scenario-ticks")
        if wi and clocks and clocks.get("sm_mhz"):
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            peak_issue = sms * 4 * clocks["sm_mhz"] * 1e6
            ach_issue = wi["warp_instructions_per_scenario_tick"] * (value / NS / world)
            line["issue_roofline"] = {"bound": "issue", "achieved": ach_issue / 1e9, "peak": peak_issue / 1e9, "unit": "G warp-instr/s",
                                      "frac": ach_issue / peak_issue, "warp_instructions_per_scenario_tick": wi["warp_instructions_per_scenario_tick"],
                                      "source": wi["source"], "sms": sms, "sm_mhz": clocks["sm_mhz"]}
        R_ = WORKLOADS[args.workload][2]
        line["log"] = {"mode": args.log,
                       "kept": "event records + per-tick alive / per-radar visibility masks + track journal: Manager.messages of any "
                               "scenario is rebuilt from them (zrk_read_messages)" if args.log == "masks" else "event records only",
                       "mask_bytes_per_step": int(B * ticks * (1 + max(1, R_)) * ((NS + 31) // 32) * 4) if args.log == "masks" else 0}
        if events_only:
            events_only["value"] = units_all / (events_only["ms_per_step"] * 1e-3)
            line["log"]["without_masks"] = events_only
        if others:
            line["other_workloads"] = others
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            v, n, tk, dt = cpu_reference_throughput(args.workload, 8192, cores, min(ticks, 250))  # (c5 on 16 threads: the whole shard, ~10 s)
            line["cpu_baseline"] = {"value": v, "unit": "object-updates/s", "cores": cores, "kind": "port",
                                    "sample": f"{n} scenarios x {NS} slots x {tk} ticks in {dt:.2f}s, "
                                              f"oracle/zrk_oracle.c on {cores} host threads",
                                    "python_loop": python_loop(args.workload)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--ticks", type=int, default=0)
    ap.add_argument("--ticks-per-launch", type=int, default=0)
    ap.add_argument("--waves", type=int, default=8,
                    help="scenario waves of the end-to-end leg (zrk_options.waves): the H2D copy of one wave overlaps the kernels of the others")
    ap.add_argument("--wave-rounds", type=int, default=0, help="zrk_options.wave_rounds of the end-to-end leg (0: 2 on 4 or more GPUs, 1 on fewer)")
    ap.add_argument("--log", default="masks", choices=["masks", "events"],
                    help="masks (default): event records + per-tick masks + track journal, every message rebuildable; events: event records only")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true", help="skip the extra end-to-end variants and the other workloads")
    ap.add_argument("--no-certificate", action="store_true", help="A/B: always run CCP.link_object's scan (ZRK_OPT_NO_CERTIFICATE)")
    ap.add_argument("--shard-offset", type=int, default=0, help="run the scenario shard of rank + this (to look at another rank's data on one GPU)")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
