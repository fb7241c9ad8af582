#!/usr/bin/env python3
"""Per-instruction view of an .ncu-rep (source page): executed counts, lanes, top stall, opcode mix.

usage: tools/ncu_source.py x.ncu-rep [min_exec_fraction]  -> prints the hot SASS with counts
"""
import csv
import io
import subprocess
import sys
from collections import Counter


def load(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    # first row: kernel name; second: header
    hdr = rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:] if len(r) == len(hdr)]


def main():
    rep = sys.argv[1]
    frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.2
    rows = load(rep)
    mx = max(int(r["Instructions Executed"]) for r in rows)
    tot = sum(int(r["Instructions Executed"]) for r in rows)
    mix = Counter()
    for n, r in enumerate(rows):
        ex = int(r["Instructions Executed"])
        op = r["Source"].split()[0] if not r["Source"].lstrip().startswith("@") else r["Source"].split()[1]
        mix[op.split(".")[0]] += ex
        if ex >= frac * mx:
            stalls = {k[6:]: int(v) for k, v in r.items() if k.startswith("stall_") and "(Not" not in k and v not in ("", "0")}
            top = sorted(stalls.items(), key=lambda kv: -kv[1])[:2]
            print(f"{n:5d} {ex / mx:6.2f} {float(r['Avg. Predicated-On Threads Executed'] or 0):5.1f} {r['# Samples']:>6s} "
                  f"{r['Source'][:80]:80s} {top}")
    print(f"total warp-inst {tot}, hottest {mx}")
    for op, c in mix.most_common(40):
        print(f"  {op:12s} {c / tot * 100:6.2f}%  {c / mx:8.2f} per hottest-iteration")


if __name__ == "__main__":
    main()
