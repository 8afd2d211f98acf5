#!/usr/bin/env python
"""Top source lines of an ncu report by executed warp instructions (uses the source embedded in
the report, so it stays right after the tree has moved on).

    python profiles/toplines.py gpurun_out/prof.ncu-rep [n_lines] [scenario_ticks]
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    sticks = float(sys.argv[3]) if len(sys.argv) > 3 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = next(r for r in rows if len(r) > 10 and r[0] == "Line No")
    ii, isamp, isrc = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
    agg, text, cur = {}, {}, None
    for r in rows:
        if len(r) == 2 and r[0] in ("File Path", "File Name"):
            cur = r[1].split("/")[-1]
        if len(r) > 10 and r[0] not in ("", "Line No"):
            try:
                k = (cur, int(r[0]))
            except ValueError:
                continue
            if r[isrc] and k not in text:
                text[k] = r[isrc]
            if r[ii] not in ("-", ""):
                a = agg.get(k, (0, 0))
                agg[k] = (a[0] + int(r[ii]), a[1] + (int(r[isamp]) if r[isamp] not in ("-", "") else 0))
    tot = sum(a[0] for a in agg.values()) or 1
    ts = sum(a[1] for a in agg.values()) or 1
    print(f"{'file:line':24s} {'inst%':>6s} {'samp%':>6s} {'instr/scen-tick':>15s}  source")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
        per = f"{a[0] / sticks:15.1f}" if sticks else " " * 15
        print(f"{k[0] + ':' + str(k[1]):24s} {100 * a[0] / tot:6.2f} {100 * a[1] / ts:6.2f} {per}  {text.get(k, '').strip()[:100]}")


if __name__ == "__main__":
    main()
