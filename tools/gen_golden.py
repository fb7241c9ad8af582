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


CLOCK_S_MS = {"unit": [1, 0], "precision": [1, 1]}
CLOCK_S_S = {"unit": [1, 0], "precision": [1, 0]}


def generic_cases(ref, model_id, params_cls, clock, grid, trace_grid):
    """grid: list of (params tuple, seed, first_rep, n_reps, until)."""
    clk = abi.Clock.of(tuple(clock["unit"]), tuple(clock["precision"]))
    cases, traces = [], []
    for fields, seed, first, reps, until in grid:
        p = params_cls(*fields)
        cols = ref.run(model_id, p, clk, seed, first, reps, threads=8, until=until)
        cases.append({"params": list(fields), "seed": seed, "first_replication": first,
                      "n_replications": reps, "run_until": until, "columns": columns_to_json(cols)})
    for fields, seed, rep in trace_grid:
        p = params_cls(*fields)
        total, tr = ref.trace(model_id, p, clk, seed, rep, max_events=20000)
        traces.append({"params": list(fields), "seed": seed, "replication": rep, "n_events": total, "trace": tr})
    return {"clock": clock, "cases": cases, "traces": traces}


def aloha_cases(ref):
    grid = [
        ((64, 50, 0.1, 3.0, 1.0, 1000.0, 0), 0x5EED, 0, 50, -1),      # BASELINE config 3: the 50-point sweep
        ((3, 0, 2.0, 2.0, 1.0, 1000.0, 2), 0x5EED, 0, 8, -1),         # 21-event identity
        ((16, 0, 1.0, 1.0, 1.0, 1000.0, 150), 0x5EED, 40, 16, -1),
        ((64, 0, 3.0, 3.0, 1.0, 1000.0, 100), 1, 0, 8, -1),            # heavy load, many key ties
        ((5, 0, 1.0, 1.0, 1.0, 1000.0, 0), 0x5EED, 0, 4, -1),          # zero packets
        ((1, 0, 0.5, 0.5, 1.0, 1000.0, 20), 0x5EED, 0, 4, -1),         # a single station never collides
        ((16, 0, 1.0, 1.0, 1.0, 1000.0, 100), 7, 0, 16, 30000),        # run_until(30 s)
    ]
    traces = [((3, 0, 2.0, 2.0, 1.0, 1000.0, 2), 0x5EED, 0), ((8, 0, 2.0, 2.0, 1.0, 1000.0, 5), 0x5EED, 3)]
    out = generic_cases(ref, abi.MODEL_ALOHA, abi.AlohaParams, CLOCK_S_MS, grid, traces)
    out["model"] = "aloha"
    return out


def mmc_cases(ref):
    grid = [
        ((9.0, 1.0, 0.5, 8, 0, 1000), 0x5EED, 0, 16, -1),              # BASELINE config 4 shape
        ((6.0, 1.0, 0.5, 8, 0, 400), 0x5EED, 5, 16, -1),
        ((12.0, 1.0, 0.5, 8, 0, 400), 0x5EED, 0, 16, -1),
        ((2.0, 1.0, 0.5, 1, 0, 300), 2, 0, 16, -1),                    # c = 1
        ((9.0, 1.0, 0.0, 8, 0, 200), 2, 0, 8, -1),                     # zero patience: timeout ties with the acquisition
        ((9.0, 1.0, 0.5, 8, 0, 0), 2, 0, 4, -1),                       # no customers
        ((9.0, 1.0, 0.5, 8, 0, 300), 3, 0, 16, 20000),                 # run_until(20 s)
    ]
    traces = [((9.0, 1.0, 0.5, 1, 0, 3), 0x5EED, 0), ((9.0, 1.0, 0.5, 2, 0, 12), 0x5EED, 1)]
    out = generic_cases(ref, abi.MODEL_MMC, abi.MMCParams, CLOCK_S_MS, grid, traces)
    out["model"] = "mmc"
    return out


def archsim_cases(ref):
    grid = [
        ((1, 10, 16, 32, 1, 0, 1024), 0x5EED, 0, 16, -1),              # BASELINE config 5 shape (sequential driver)
        ((1, 10, 16, 32, 4, 0, 256), 0x5EED, 0, 16, -1),               # 4 requesters: mutex contention
        ((1, 10, 4, 8, 4, 0, 300), 0x5EED, 100, 16, -1),
        ((0, 0, 4, 8, 4, 0, 100), 0x5EED, 0, 16, -1),                  # zero latencies: everything ties at t = 0
        ((1, 10, 2, 64, 3, 0, 100), 9, 0, 8, -1),
        ((1, 10, 16, 32, 2, 0, 0), 9, 0, 4, -1),                       # no loads
        ((1, 10, 16, 32, 4, 0, 200), 1, 0, 16, 500),                   # run_until(500)
    ]
    traces = [((1, 10, 16, 32, 2, 0, 2), 0x5EED, 0), ((1, 10, 2, 4, 3, 0, 4), 0x5EED, 2)]
    out = generic_cases(ref, abi.MODEL_ARCHSIM, abi.ArchSimParams, CLOCK_S_S, grid, traces)
    out["model"] = "archsim"
    return out


def kat_cases(ref):
    import kat_programs as K
    clk = abi.Clock.of((1, 0), (1, 0))
    cases = []
    for k, fn in sorted(K.SCENARIOS.items()):
        _, until = fn()
        kp = K.KatParams(k, 0)
        cols = ref.run(K.ORACLE_MODEL_KAT, kp, clk, 1, 0, 1, until=until)
        total, tr = ref.trace(K.ORACLE_MODEL_KAT, kp, clk, 1, 0, until=until, max_events=1000)
        cases.append({"scenario": k, "name": fn.__name__, "run_until": until, "n_replications": 1,
                      "columns": columns_to_json(cols), "n_events": total, "trace": tr})
    return {"model": "kat", "clock": CLOCK_S_S, "cases": cases}


def program_cases(ref):
    """Random process programs (tests/test_programs_fuzz.py random_program, by seed) run by the reference engine
    itself through oracle/ref/program_runner.hpp: run() columns and the first events of replication 0."""
    import test_programs_fuzz as F
    cases = []
    for case in F.CASES:
        prog = F.random_program(case)
        clk = abi.Clock.of(*F.UNIT)
        cols = ref.run(abi.MODEL_PROGRAM, prog.to_params(), clk, 1000 + case, 0, 8, threads=4)
        until = int(cols.end_time.max()) // 2
        part = ref.run(abi.MODEL_PROGRAM, prog.to_params(), clk, 1000 + case, 0, 8, threads=4, until=until)
        total, tr = ref.trace(abi.MODEL_PROGRAM, prog.to_params(), clk, 1000 + case, 0, max_events=64)
        cases.append({"program_seed": case, "seed": 1000 + case, "n_replications": 8, "columns": columns_to_json(cols),
                      "run_until": until, "columns_until": columns_to_json(part), "n_events": total, "trace": tr})
    return {"model": "program", "clock": CLOCK_S_MS, "cases": cases}


def main():
    ref = oracles.ref_oracle()
    if ref is None:
        raise SystemExit("oracle/_ref is not built: run in the container that has /root/reference")
    GOLDEN.mkdir(parents=True, exist_ok=True)
    out = mm1_cases(ref)
    (GOLDEN / "mm1_reference.json").write_text(json.dumps(out, indent=0, separators=(",", ":")))
    print("wrote", GOLDEN / "mm1_reference.json")
    for name, fn in (("aloha", aloha_cases), ("mmc", mmc_cases), ("archsim", archsim_cases), ("kat", kat_cases),
                     ("programs_random", program_cases)):
        (GOLDEN / f"{name}_reference.json").write_text(json.dumps(fn(ref), indent=0, separators=(",", ":")))
        print("wrote", GOLDEN / f"{name}_reference.json")


if __name__ == "__main__":
    main()
