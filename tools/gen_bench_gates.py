#!/usr/bin/env python3
"""Parity-gate constants for bench.py: checksums of the first K replications of every shard a rank can own, for every
BASELINE config, computed by the CPU oracle (oracle/port; pinned to the reference engine by the CPU test suite).

    python tools/gen_bench_gates.py            -> tests/golden/bench_gates.json

bench.py fetches its result columns once before timing, sums rows [0, K) of its shard and refuses to time anything
unless {n_events, end_time, trace_hash (wrapping)} equal the committed sums -- it never executes the oracle for this.
M/M/1 additionally has the sums over ALL replications of each 1 Mi shard (tools/gen_full_parity.py)."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT / "tools"))


def main():
    import oracles
    from libcxxdes_b200 import abi
    import bench_workloads as W
    port = oracles.port_oracle()
    threads = os.cpu_count() or 1
    out = {"generator": "tools/gen_bench_gates.py", "oracle": "oracle/port", "seed": W.SEED, "windows": {}}
    for key, w in W.workloads().items():
        t0 = time.perf_counter()
        rows = {}
        for first in W.gate_firsts(w):
            c = port.run(w.model, w.params, abi.Clock.of(*w.clock), W.SEED, first, w.gate_window, threads=threads)
            assert not c.status.any()
            rows[str(first)] = {"n": w.gate_window,
                                "n_events_sum": int(c.n_events.sum(dtype=np.uint64)),
                                "end_time_sum": int(c.end_time.sum()),
                                "trace_hash_sum": int(c.trace_hash.sum(dtype=np.uint64))}
        out["windows"][key] = rows
        print(f"{key}: {len(rows)} windows of {w.gate_window} replications in {time.perf_counter() - t0:.1f} s", flush=True)
    path = ROOT / "tests" / "golden" / "bench_gates.json"
    path.write_text(json.dumps(out, indent=1) + "\n")
    print(f"wrote {path}")


if __name__ == "__main__":
    main()
