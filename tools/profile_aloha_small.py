#!/usr/bin/env python3
"""ALOHA at a tenth of the BASELINE packet count (one wave of replications): the ncu target for the ALOHA kernel."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import libcxxdes_b200 as L  # noqa: E402
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 384
ens = L.Ensemble(L.AlohaParams(64, 50, 0.1, 3.0, 1.0, 100.0, 0), reps, seed=0x5EED)
ens.time_unit(1, L.SECONDS).time_precision(1, L.MILLISECONDS)
ens.run().sync()
acc, _ = ens.reduce()
print(f"aloha small: {reps} replications, {acc.n_events} events, {ens.last_run_ms():.3f} ms, {acc.n_events / ens.last_run_ms() / 1e6:.2f} Gev/s")
ens.close()
