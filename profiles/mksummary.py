#!/usr/bin/env python
"""Builds profiles/<tag>_<workload>_summary.json from an ncu launch list of one bench step (launchsum.py's sums plus
the launch-time-weighted issue utilisation), the file bench.py reads `roofline.traffic` and the issue roofline's
instruction count from.

    python profiles/mksummary.py gpurun_out/r02f_launches_c5.csv c5 32768000 21 > profiles/r02f_c5_summary.json
    (arguments: launch list, workload, scenario-ticks per step, seq — bench.py takes the largest seq of a workload)
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import launchsum  # noqa: E402


def main():
    path, workload, sticks, seq = sys.argv[1], sys.argv[2], float(sys.argv[3]), int(sys.argv[4])
    data = launchsum.load(path)
    agg, done = {}, False
    wsum = tsum = 0.0
    for i in sorted(data):
        d = data[i]
        k = "streak" if "streak" in d["k"] else "tick" if "tick" in d["k"] else "init" if "init" in d["k"] else "totals/summary"
        if k == "totals/summary":
            done = True
        if done and k in ("streak", "tick", "init"):
            continue  # (launches after the first read-back belong to bench.py's later legs)
        a = agg.setdefault(k, {"launches": 0, "ms": 0.0, "warp_inst": 0.0, "dram_bytes": 0.0})
        a["launches"] += 1
        a["ms"] += d.get("gpu__time_duration.sum", 0)
        a["warp_inst"] += d.get("smsp__inst_executed.sum", 0)
        a["dram_bytes"] += d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
        if k in ("streak", "tick"):
            t = d.get("gpu__time_duration.sum", 0)
            wsum += t * d.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0)
            tsum += t
    step = {k: sum(a[k] for n, a in agg.items() if n in ("streak", "tick")) for k in ("launches", "ms", "warp_inst", "dram_bytes")}
    out = {"round": 2, "seq": seq, "workload": workload,
           "kernel": "zrk_streak_kernel + zrk_tick_kernel (two-kernel stepping)" if "streak" in agg else "zrk_tick_kernel",
           "command": "ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,"
                      "dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:zrk_ --csv python bench.py --workload "
                      f"{workload} --steps 1 --warmup 0 --no-cpu-baseline --no-variants (scratch/r2c_prof.sh)",
           "launches_per_step": int(step["launches"]), "gpu_time_ms_sum": round(step["ms"], 3),
           "dram_bytes_per_step": int(step["dram_bytes"]), "dram_bytes_per_launch": int(step["dram_bytes"]),
           "warp_instructions_per_scenario_tick": round(step["warp_inst"] / sticks, 1),
           "issue_active_pct": round(wsum / tsum, 1) if tsum else None,
           "per_kernel": agg,
           "note": "sums over all tick / streak launches of one bench step (serialised, cold-cache ncu timings); "
                   "dram_bytes_per_launch = per step, kept under the old key for bench.py; issue_active_pct = launch-time-weighted mean",
           "source": os.path.basename(path)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
