#!/usr/bin/env python3
"""Run one ensemble of any model through the C ABI and print events/s (also the ncu target).

  tools/profile_model.py aloha|mmc|archsim|archsim4|mm1|mm1prog [replications] [runs] [cpu_replications]
Configurations are the BASELINE.json configs (3, 4, 5 and 2); mm1prog is M/M/1 (1000 packets) as a process
program on the interpreter kernel.  With cpu_replications > 0 the reference engine (oracle/_ref, else the
restated oracle) is timed on that many replications on all host threads, for the speed-up column."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import libcxxdes_b200 as L  # noqa: E402
from libcxxdes_b200 import abi  # noqa: E402

CONFIGS = {
    "mm1": (L.MM1Params(5.0, 10.0, 10000), 1 << 20, ((1, L.SECONDS), (1, L.MILLISECONDS))),
    "aloha": (L.AlohaParams(64, 50, 0.1, 3.0, 1.0, 1000.0, 0), 1 << 18, ((1, L.SECONDS), (1, L.MILLISECONDS))),
    "mmc": (L.MMCParams(9.0, 1.0, 0.5, 8, 0, 1000), 1 << 20, ((1, L.SECONDS), (1, L.MILLISECONDS))),
    "archsim": (L.ArchSimParams(1, 10, 16, 32, 1, 0, 1024), 1 << 22, ((1, L.SECONDS), (1, L.SECONDS))),
    "archsim4": (L.ArchSimParams(1, 10, 16, 32, 4, 0, 256), 1 << 22, ((1, L.SECONDS), (1, L.SECONDS))),
}

name = sys.argv[1]
if name == "mm1prog":
    import kat_programs
    CONFIGS[name] = (kat_programs.mm1_program(5.0, 10.0, 1000), 1 << 18, ((1, L.SECONDS), (1, L.MILLISECONDS)))
params, reps, clock = CONFIGS[name]
if len(sys.argv) > 2:
    reps = int(sys.argv[2])
runs = int(sys.argv[3]) if len(sys.argv) > 3 else 2
import os
flags = int(os.environ.get("CXB_FLAGS", "0"))   # e.g. 4 = CXXDES_B200_FLAG_COOP_HEAP
ens = L.Ensemble(params, reps, seed=0x5EED, flags=flags)
ens.time_unit(*clock[0]).time_precision(*clock[1])
for _ in range(runs):
    ens.run().sync()
    acc, _ = ens.reduce()
    ms = ens.last_run_ms()
    print(f"{name} [{ens.kernel_name()}]: {reps} replications, {acc.n_events} events, {ms:.3f} ms, {acc.n_events / ms / 1e6:.2f} Gev/s, "
          f"errors {acc.n_errors}", flush=True)
ens.close()

cpu_reps = int(sys.argv[4]) if len(sys.argv) > 4 else 0
if cpu_reps > 0:
    import os
    import time
    import oracles
    MODEL = {"mm1": abi.MODEL_MM1, "aloha": abi.MODEL_ALOHA, "mmc": abi.MODEL_MMC, "archsim": abi.MODEL_ARCHSIM,
             "archsim4": abi.MODEL_ARCHSIM, "mm1prog": abi.MODEL_PROGRAM}[name]
    oracle = oracles.ref_oracle() if name != "mm1prog" else None
    kind = "reference engine"
    if oracle is None:
        oracle, kind = oracles.port_oracle(), "restated oracle"
    p = params.to_params() if hasattr(params, "to_params") else params
    threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    cols = oracle.run(MODEL, p, abi.Clock.of(*clock), 0x5EED, 0, cpu_reps, threads=threads)
    dt = time.perf_counter() - t0
    ev = int(cols.n_events.sum())
    print(f"{name} cpu ({kind}, {threads} threads): {cpu_reps} replications, {ev} events, {dt:.2f} s, "
          f"{ev / dt / 1e6:.2f} Mev/s -> GPU/CPU = {acc.n_events / ms * 1e3 / (ev / dt):.0f}x", flush=True)
