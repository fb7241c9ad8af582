#!/usr/bin/env python3
"""Kernel-time sweep of the M/M/1 ensemble over loads (device events around the run).

usage: tools/sweep_mm1.py [replications] [packets]   -> one line per (geometry, lambda)
"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import libcxxdes_b200 as L  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
packets = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
geoms = [0]   # (one launch geometry is built; earlier sweeps compared several)
lams = [float(x) for x in os.environ.get("SWEEP_LAMBDAS", "1,5").split(",")]
for geom in geoms:
    for lam in lams:
        ens = L.Ensemble(L.MM1Params(lam, 10.0, packets), reps, seed=0x5EED)
        ens.time_unit(1, L.SECONDS).time_precision(1, L.MILLISECONDS)
        best = None
        for _ in range(3):
            ens.run().sync()
            ms = ens.last_run_ms()
            best = ms if best is None else min(best, ms)
        acc, _ = ens.reduce()
        print(f"geometry {geom} lambda {lam}: {best:.3f} ms, {acc.n_events / best / 1e6:.2f} Gev/s, errors {acc.n_errors}", flush=True)
        ens.close()
