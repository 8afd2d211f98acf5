#!/usr/bin/env python
"""Summarise an ncu report of the tick kernel: headline metrics + instruction share per kernel
region (by CUDA source line, from `ncu --page source --print-source cuda,sass`).

    python profiles/analyze.py gpurun_out/prof.ncu-rep [scenario_ticks]
"""
import csv
import io
import subprocess
import sys

import os

SRC = os.environ.get("ZRK_ANALYZE_SRC", "zrk_modulation_b200/csrc/zrk_tick.cuh")  # the source the profiled library was built from
MARKS = [("helpers (dot3, pymod, tsec)", None), ("radar_sees", "SectorRadar.find_visible_objects for one object"),
         ("radar_move", "SectorRadar.move_to_next_sector_circular"), ("team primitives", "team primitives"),
         ("put_rec", "record log"), ("missile_solve", "Missile._calculate_trajectory_params"),
         ("views", "shared-memory views"), ("calc_range/link_exact", "CCP.link_object"),
         ("link_eval", "State of the squared-domain scan"), ("link_d2 (fp64 pair distance)", "double link_d2("), ("link scan (fp32 filter + loop)", "fp32 pre-filter of the scan"),
         ("association certificate", "---- association certificate"), ("streak: rival check, sweep", "---- quiet streak"),
         ("streak: lean loop, one tick per pass", "struct StreakLane {"), ("streak: lean loop, two ticks per pass", "// The lean loop for the commonest shape"),
         ("streak: frozen slots, event records", "// AirEnv.py:33-40 inside a streak"),
         ("streak: entry (state from the blob)", "__device__ __forceinline__ int quiet_streak("),
         ("streak: careful tick", "const double now_next = tick_seconds(P, tick + 1);"), ("streak: commit, write-back", "// ---- commit"),
         ("serial_section", "lane 0 alone: launchers"), ("kernel prologue", "---- the kernel"),
         ("tick head", "for (int it = 0; it < n_ticks; it++)"), ("AirEnv.step", "AirEnv.step (AirEnv.py:26-53)"),
         ("Missile.step", "Missile.step, status 'active'"), ("radar loop", "SectorRadar.step for every radar"),
         ("sector sweep + serial call", "lane 0: sector sweep"), ("CCP link arrays + order", "CombatControlPoint.step, detections"),
         ("CCP rounds", "int lo_rank = 0;"), ("send_objects_to_GUI", "send_objects_to_GUI (CCP.py:237-284)"),
         ("tick end", "if (nrec > P.log_cap)"), ("streak kernel wrapper", "---- the streak kernel")]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    sticks = float(sys.argv[2]) if len(sys.argv) > 2 else None
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    m = dict(zip(hdr, vals))
    want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
            "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
    u = dict(zip(hdr, units))
    for k in want:
        if k in m:
            print(f"{k} = {m[k]} {u.get(k, '')}")
    if sticks and "smsp__inst_executed.sum" in m:
        print(f"warp instructions per scenario-tick = {float(m['smsp__inst_executed.sum'].replace(',', '')) / sticks:.0f}")
    src = open(SRC).read().splitlines()

    def find(pat):
        for i, l in enumerate(src):
            if pat in l:
                return i + 1
        return None

    marks = [(n, 1 if p is None else find(p)) for n, p in MARKS]
    marks = [(n, l) for n, l in marks if l] + [("end", len(src) + 1)]
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--print-source", "cuda,sass", "--csv"))))
    h = next(r for r in rows if len(r) > 10 and r[0] == "Line No")
    ii, isamp = h.index("Instructions Executed"), h.index("# Samples")
    agg, cur = {}, None
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        if len(r) > 10 and r[0] not in ("", "Line No") and r[ii] not in ("-", ""):
            try:
                k = (cur, int(r[0]))
                a = agg.get(k, (0, 0))
                agg[k] = (a[0] + int(r[ii]), a[1] + (int(r[isamp]) if r[isamp] not in ("-", "") else 0))
            except ValueError:
                pass
    tot = sum(a[0] for a in agg.values()) or 1
    ts = sum(a[1] for a in agg.values()) or 1
    print("\nregion                                  inst%   stall-samples%")
    for (name, a), (_, b) in zip(marks[:-1], marks[1:]):
        i = sum(v[0] for (f, l), v in agg.items() if f == "zrk_tick.cuh" and a <= l < b)
        s = sum(v[1] for (f, l), v in agg.items() if f == "zrk_tick.cuh" and a <= l < b)
        print(f"{name:38s} {100 * i / tot:6.1f} {100 * s / ts:8.1f}")
    i = sum(v[0] for (f, l), v in agg.items() if f != "zrk_tick.cuh")
    print(f"{'other files (intrinsics, libm)':38s} {100 * i / tot:6.1f}")


if __name__ == "__main__":
    main()
