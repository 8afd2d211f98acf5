#!/usr/bin/env python
"""Sums an ncu launch list (`--metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_*`) of one
bench step per kernel: launches, time, warp instructions, DRAM bytes.

    python profiles/launchsum.py gpurun_out/r02_launches_c5.csv [scenario_ticks]
"""
import csv
import io
import json
import sys


def load(path):
    rows = [l for l in open(path) if l.startswith('"')]
    r = list(csv.reader(io.StringIO("".join(rows))))
    h = r[0]
    ik, im, iv, iid = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
    iu = h.index("Metric Unit")
    data = {}
    for x in r[1:]:
        d = data.setdefault(int(x[iid]), {"k": x[ik]})
        v = float(x[iv].replace(",", ""))
        u = x[iu]
        if x[im].startswith("dram__bytes"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        if x[im] == "gpu__time_duration.sum":
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(u, 1e-6)
        d[x[im]] = v
    return data


def main():
    data = load(sys.argv[1])
    sticks = float(sys.argv[2]) if len(sys.argv) > 2 else None
    agg = {}
    step_done = False
    for i in sorted(data):
        d = data[i]
        k = "streak" if "streak" in d["k"] else "tick" if "tick" in d["k"] else "init" if "init" in d["k"] else \
            "totals/summary"
        if k == "totals/summary":
            step_done = True
        if step_done and k in ("streak", "tick", "init"):
            continue  # (launches after the first read-back belong to bench.py's later legs)
        a = agg.setdefault(k, {"launches": 0, "ms": 0.0, "warp_inst": 0.0, "dram_bytes": 0.0})
        a["launches"] += 1
        a["ms"] += d.get("gpu__time_duration.sum", 0)
        a["warp_inst"] += d.get("smsp__inst_executed.sum", 0)
        a["dram_bytes"] += d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
    out = {"per_kernel": agg}
    step = {k: sum(a[k] for n, a in agg.items() if n in ("streak", "tick")) for k in ("launches", "ms", "warp_inst", "dram_bytes")}
    out["step"] = step
    if sticks:
        out["warp_instructions_per_scenario_tick"] = step["warp_inst"] / sticks
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
