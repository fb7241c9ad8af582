#!/usr/bin/env python3
"""Generate tests/golden/*.json from the UNMODIFIED reference engine (oracle/_ref).

Run in the build container (where /root/reference exists):  python tools/gen_golden.py
The fixtures pin the restated oracle and the CUDA path on machines without the reference.
Each case stores per-replication columns (hex for 64-bit words) and, for a few replications,
the full popped-token trace.
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import oracles  # noqa: E402
from libcxxdes_b200 import abi  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"


def columns_to_json(c):
    return {
        "end_time": c.end_time.tolist(),
        "n_events": c.n_events.tolist(),
        "trace_hash": [f"{int(x):016x}" for x in c.trace_hash],
        "f64_bits": [[f"{int(x):016x}" for x in a.view(np.uint64)] for a in c.f64],
        "i64": [a.tolist() for a in c.i64],
    }


def mm1_cases(ref):
    cases = []
    grid = [
        # lambda, mu, n_packets, seed, first_rep, n_reps, until
        (5.0, 10.0, 10000, 0x5EED, 0, 24, -1),          # BASELINE config 2 shape
        (5.0, 10.0, 100000, 0, 0, 1, -1),               # BASELINE config 1 (single CPU run)
        (0.5, 10.0, 2000, 0x5EED, 100, 32, -1),
        (9.0, 10.0, 2000, 0x5EED, 1 << 32, 32, -1),     # replication ids above 2^32
        (9.9, 10.0, 3000, 0xDEADBEEFCAFEF00D, 7, 16, -1),
        (12.0, 10.0, 500, 1, 0, 16, -1),                # unstable queue
        (5.0, 10.0, 0, 3, 0, 4, -1),
        (5.0, 10.0, 1, 3, 0, 4, -1),
        (5.0, 10.0, 2, 3, 0, 8, -1),
        (5.0, 10.0, 3, 3, 0, 8, -1),
        (5.0, 10.0, 1000, 0x5EED, 0, 32, 50000),        # run_until(50 s) stops mid-run
        (5.0, 10.0, 50, 0x5EED, 0, 16, 10 ** 9),        # run_until far beyond the end
    ]
    for lam, mu, n, seed, first, reps, until in grid:
        p = abi.MM1Params(lam, mu, n)
        cols = ref.run(abi.MODEL_MM1, p, oracles.MM1_CLOCK, seed, first, reps, threads=8, until=until)
        case = {"lambda": lam, "mu": mu, "n_packets": n, "seed": seed, "first_replication": first,
                "n_replications": reps, "run_until": until, "columns": columns_to_json(cols)}
        cases.append(case)
    # full traces of two small replications
    traces = []
    for lam, n, seed, rep in [(5.0, 6, 0x5EED, 0), (9.0, 12, 0x5EED, 5)]:
        p = abi.MM1Params(lam, 10.0, n)
        total, tr = ref.trace(abi.MODEL_MM1, p, oracles.MM1_CLOCK, seed, rep)
        traces.append({"lambda": lam, "mu": 10.0, "n_packets": n, "seed": seed, "replication": rep,
                       "n_events": total, "trace": tr})
    return {"model": "mm1", "clock": {"unit": [1, 0], "precision": [1, 1]}, "cases": cases,
            "traces": traces}


def main():
    ref = oracles.ref_oracle()
    if ref is None:
        raise SystemExit("oracle/_ref is not built: run in the container that has /root/reference")
    GOLDEN.mkdir(parents=True, exist_ok=True)
    out = mm1_cases(ref)
    (GOLDEN / "mm1_reference.json").write_text(json.dumps(out, indent=0, separators=(",", ":")))
    print("wrote", GOLDEN / "mm1_reference.json")


if __name__ == "__main__":
    main()
