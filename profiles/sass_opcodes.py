#!/usr/bin/env python
"""Opcode histogram of the shipped library, per kernel, from `cuobjdump -sass` — the evidence behind
DESIGN.md's statements about the generated code (bulk-TMA state staging, packed fp32x2 scan, no
tensor-core or tile-TMA instructions because the path has neither a contraction nor a tile).

    python profiles/sass_opcodes.py > profiles/sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "zrk_modulation_b200", "csrc", "libzrk_b200.so")
NOTE = {"UBLKCP": "cp.async.bulk (non-tensor TMA): state blob staging", "SYNCS": "mbarrier ops",
        "FFMA2": "packed fp32x2 FMA (association scan)", "FADD2": "packed fp32x2 add", "FMUL2": "packed fp32x2 mul",
        "DFMA": "fp64 FMA (np.dot chains)", "DADD": "fp64 add", "DMUL": "fp64 mul", "SHFL": "warp shuffle",
        "VOTE": "ballot", "REDUX": "warp reduce", "MATCH": "match.any", "BAR": "CTA barrier", "UTMALDG": "tile TMA load",
        "UTCMMA": "tcgen05.mma", "HMMA": "mma.sync", "LDS": "shared load", "STS": "shared store", "LDG": "global load",
        "STG": "global store", "ATOMS": "shared atomic", "ATOMG": "global atomic", "RED": "global reduction", "MUFU": "SFU"}
WATCH = ["UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2", "DFMA", "DADD", "DMUL", "SHFL", "VOTE", "REDUX", "MATCH", "BAR", "LDS", "STS",
         "LDG", "STG", "ATOMS", "ATOMG", "RED", "MUFU", "UTMALDG", "UTCMMA", "HMMA"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    parts = re.split(r"\n\s+Function : ", txt)
    total = collections.Counter()
    print(f"# {os.path.relpath(LIB, ROOT)}: SASS opcodes per kernel (cuobjdump -sass, sm_100a)\n")
    for part in parts[1:]:
        name = part.split("\n")[0].strip()
        ops = collections.Counter()
        n = 0
        for m in re.finditer(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(?:\.[A-Z0-9_.]+)?\s", part):
            ops[m.group(1)] += 1
            n += 1
        total.update(ops)
        demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
        short = re.sub(r"\(.*", "", demangled)
        print(f"{short}: {n} instructions ({n * 16 / 1024:.0f} KB)")
        print("   " + "  ".join(f"{k} {ops[k]}" for k in WATCH if ops[k]))
    print("\n# whole library")
    for k in WATCH:
        print(f"{k:8s} {total[k]:7d}  {NOTE.get(k, '')}")
    print("\n(UTMALDG / UTCMMA / HMMA = 0: no tile movement, no contraction on this path.)")


if __name__ == "__main__":
    main()
