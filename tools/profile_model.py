#!/usr/bin/env python3
"""Run one ensemble of any model through the C ABI and print events/s (also the ncu target).

  tools/profile_model.py aloha|mmc|archsim|archsim4|mm1 [replications] [runs]
Configurations are the BASELINE.json configs (3, 4, 5 and 2)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import libcxxdes_b200 as L  # noqa: E402

CONFIGS = {
    "mm1": (L.MM1Params(5.0, 10.0, 10000), 1 << 20, ((1, L.SECONDS), (1, L.MILLISECONDS))),
    "aloha": (L.AlohaParams(64, 50, 0.1, 3.0, 1.0, 1000.0, 0), 1 << 18, ((1, L.SECONDS), (1, L.MILLISECONDS))),
    "mmc": (L.MMCParams(9.0, 1.0, 0.5, 8, 0, 1000), 1 << 20, ((1, L.SECONDS), (1, L.MILLISECONDS))),
    "archsim": (L.ArchSimParams(1, 10, 16, 32, 1, 0, 1024), 1 << 22, ((1, L.SECONDS), (1, L.SECONDS))),
    "archsim4": (L.ArchSimParams(1, 10, 16, 32, 4, 0, 256), 1 << 22, ((1, L.SECONDS), (1, L.SECONDS))),
}

name = sys.argv[1]
params, reps, clock = CONFIGS[name]
if len(sys.argv) > 2:
    reps = int(sys.argv[2])
runs = int(sys.argv[3]) if len(sys.argv) > 3 else 2
ens = L.Ensemble(params, reps, seed=0x5EED)
ens.time_unit(*clock[0]).time_precision(*clock[1])
for _ in range(runs):
    ens.run().sync()
    acc, _ = ens.reduce()
    ms = ens.last_run_ms()
    print(f"{name}: {reps} replications, {acc.n_events} events, {ms:.3f} ms, {acc.n_events / ms / 1e6:.2f} Gev/s, "
          f"errors {acc.n_errors}", flush=True)
ens.close()
