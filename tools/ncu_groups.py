#!/usr/bin/env python3
"""Share of the executed warp instructions per REGION of a kernel (pop, push, variate, event branches, ...) from an
.ncu-rep captured with --import-source on: tools/ncu_lines.py's per-line counts summed over line ranges.

usage: tools/ncu_groups.py x.ncu-rep mmc|aloha
The ranges name the source as it was when the capture was taken (they are given per capture below); this is how the
"variate block = 25 % of M/M/c's warp instructions at 5 lanes" figures of DESIGN.md 5.3 were read."""
import re
import subprocess
import sys
from pathlib import Path

# (file, first line, last line, region) -- profiles/r02_mmc_narrow_v1 (gpurun_out/r02ac/mmc_narrow.ncu-rep)
RANGES = {
    "mmc": [("models_mmc.cu", 0, 377, "claim / loop head"), ("models_mmc.cu", 378, 399, "root decode, customer word, pop call"),
            ("models_mmc.cu", 400, 493, "event branches"), ("models_mmc.cu", 494, 507, "variate glue"),
            ("models_mmc.cu", 508, 526, "push-turn glue"), ("models_mmc.cu", 527, 9999, "exit check / results"),
            ("engine32.cuh", 520, 560, "path-parallel push"), ("engine32.cuh", 561, 610, "pop sift-down"),
            ("engine32.cuh", 480, 499, "re-insertion loop"), ("engine32.cuh", 464, 468, "at / later / time_of"),
            ("philox.hpp", 0, 9999, "variate math"), ("variates.hpp", 0, 9999, "variate math"), ("trace_hash.hpp", 0, 9999, "digest")],
    "aloha": [("models_aloha.cu", 0, 552, "claim / loop head"), ("models_aloha.cu", 553, 640, "event body (decode, masks, push choice, exit check)"),
              ("models_aloha.cu", 641, 9999, "results"), ("engine32.cuh", 520, 560, "path-parallel push / re-insertion"),
              ("engine32.cuh", 561, 610, "pop sift-down"), ("engine32.cuh", 464, 468, "at / later / time_of"),
              ("philox.hpp", 0, 9999, "variate math"), ("variates.hpp", 0, 9999, "variate math"), ("trace_hash.hpp", 0, 9999, "digest")],
}


def main():
    rep, which = sys.argv[1], sys.argv[2]
    tool = Path(__file__).resolve().parent / "ncu_lines.py"
    text = subprocess.run([sys.executable, str(tool), rep, "100000"], capture_output=True, text=True, check=True).stdout
    groups = {}
    for line in text.splitlines():
        m = re.match(r"\s*([\d.]+)%\s+lanes\s+([\d.]+)\s+(\S+):(\d+)", line)
        if not m:
            continue
        share, lanes, name, number = float(m.group(1)), float(m.group(2)), m.group(3), int(m.group(4))
        region = name
        for f, lo, hi, label in RANGES[which]:
            if f == name and lo <= number <= hi:
                region = label
                break
        g = groups.setdefault(region, [0.0, 0.0])
        g[0] += share
        g[1] += share * lanes
    for region, (share, weighted) in sorted(groups.items(), key=lambda kv: -kv[1][0]):
        if share > 0:
            print(f"{share:6.2f} %  lanes {weighted / share:5.1f}  {region}")


if __name__ == "__main__":
    main()
