#!/usr/bin/env python3
"""One M/M/1 ensemble run through the C ABI, for `ncu` (tools/profile_mm1.py [replications] [packets] [runs])."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import libcxxdes_b200 as L  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
packets = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
runs = int(sys.argv[3]) if len(sys.argv) > 3 else 2
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
ens = L.Ensemble(L.MM1Params(5.0, 10.0, packets), reps, seed=0x5EED, flags=flags)
ens.time_unit(1, L.SECONDS).time_precision(1, L.MILLISECONDS)
for _ in range(runs):
    ens.run().sync()
    acc, _ = ens.reduce()
    print(f"{acc.n_events} events, {ens.last_run_ms():.3f} ms, {acc.n_events / ens.last_run_ms() / 1e6:.2f} Gev/s, errors {acc.n_errors}")
ens.close()
