#!/usr/bin/env python3
"""Executed warp instructions per CUDA source line from an .ncu-rep captured with --import-source on.

usage: tools/ncu_lines.py x.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    lines = []
    path, hdr = None, None
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            path, hdr = r[1], None
            continue
        if len(r) == 2 and r[0] == "Function Name":
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d["Address"] != "-":      # a SASS row under its source line
            continue
        try:
            lines.append((int(d["Instructions Executed"]), int(d["Thread Instructions Executed"]), path.split("/")[-1],
                          d["Line No"], r[1]))
        except ValueError:
            pass
    total = sum(x[0] for x in lines)
    print(f"total warp instructions {total}")
    for ex, th, f, ln, src in sorted(lines, reverse=True)[:top_n]:
        print(f"{100.0 * ex / total:5.2f}%  lanes {th / ex if ex else 0:4.1f}  {f}:{ln:<5} {src.strip()[:100]}")


if __name__ == "__main__":
    main()
