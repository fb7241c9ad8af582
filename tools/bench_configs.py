#!/usr/bin/env python3
"""One JSON line per BASELINE.json config (and for M/M/1 written as a process program): events/s of the CUDA path
at the config's full size, CUDA events on the launching stream, best of `runs`; next to it the reference engine
(oracle/_ref, one replication per host thread) on a bounded sample of the same workload.

  tools/bench_configs.py [runs] [cpu_seconds]      -> profiles/r01_bench_configs.jsonl is this tool's output
bench.py stays the contract's one-line bench of configs[1]; this is the table behind DESIGN.md 5.3 / 5.4."""
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import libcxxdes_b200 as L  # noqa: E402
from libcxxdes_b200 import abi  # noqa: E402
import kat_programs  # noqa: E402
import oracles  # noqa: E402

S_MS = ((1, L.SECONDS), (1, L.MILLISECONDS))
S_S = ((1, L.SECONDS), (1, L.SECONDS))
CONFIGS = [
    ("configs[1] M/M/1, 1Mi replications x 10k packets", abi.MODEL_MM1, L.MM1Params(5.0, 10.0, 10000), 1 << 20, S_MS, 2000),
    ("configs[2] ALOHA, 64 stations, 256k replications, load sweep", abi.MODEL_ALOHA,
     L.AlohaParams(64, 50, 0.1, 3.0, 1.0, 1000.0, 0), 1 << 18, S_MS, 400),
    ("configs[3] M/M/c balking, c=8, 1Mi replications x 1000 customers", abi.MODEL_MMC,
     L.MMCParams(9.0, 1.0, 0.5, 8, 0, 1000), 1 << 20, S_MS, 8000),
    ("configs[4] memory hierarchy, sequential driver, 4Mi replications x 1024 loads", abi.MODEL_ARCHSIM,
     L.ArchSimParams(1, 10, 16, 32, 1, 0, 1024), 1 << 22, S_S, 20000),
    ("memory hierarchy, 4 requesters x 256 loads, 4Mi replications", abi.MODEL_ARCHSIM,
     L.ArchSimParams(1, 10, 16, 32, 4, 0, 256), 1 << 22, S_S, 20000),
    ("M/M/1 as a process program (interpreter kernels), 1Mi replications x 1000 packets", abi.MODEL_PROGRAM,
     kat_programs.mm1_program(5.0, 10.0, 1000), 1 << 20, S_MS, 20000),
]


def main():
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    cpu_seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 4.0
    ref = oracles.ref_oracle()
    threads = os.cpu_count() or 1
    for name, model, params, reps, clock, cpu_reps in CONFIGS:
        ens = L.Ensemble(params, reps, seed=0x5EED)
        ens.time_unit(*clock[0]).time_precision(*clock[1])
        best = None
        for _ in range(runs + 1):                     # first run is the warm-up
            ens.run().sync()
            ms = ens.last_run_ms()
            best = ms if best is None or _ == 1 else min(best, ms)
        acc, _ = ens.reduce()
        launches = ens.launch_count()
        ens.close()
        line = {"workload": name, "replications": reps, "events": int(acc.n_events), "replication_errors": int(acc.n_errors),
                "kernel_ms": round(best, 3), "events_per_s": acc.n_events / (best * 1e-3), "n_gpus": 1}
        if ref is not None:
            p = params.to_params() if hasattr(params, "to_params") else params
            t0 = time.perf_counter()
            cols = ref.run(model, p, abi.Clock.of(*clock), 0x5EED, 0, cpu_reps, threads=threads)
            dt = time.perf_counter() - t0
            if dt < cpu_seconds / 4:                   # sample too small to time: scale it once
                cpu_reps = int(cpu_reps * cpu_seconds / max(dt, 1e-3))
                t0 = time.perf_counter()
                cols = ref.run(model, p, abi.Clock.of(*clock), 0x5EED, 0, cpu_reps, threads=threads)
                dt = time.perf_counter() - t0
            ev = int(cols.n_events.sum())
            line["cpu_reference"] = {"events_per_s": ev / dt, "threads": threads, "kind": "reference engine (oracle/_ref)",
                                     "sample": f"{cpu_reps} replications, {ev} events in {dt:.2f} s"}
            line["gpu_over_cpu"] = line["events_per_s"] / (ev / dt)
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
