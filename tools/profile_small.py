#!/usr/bin/env python3
"""One wave of replications of a model at a reduced size: the ncu target for the slower kernels
(ncu replays a kernel ~40 times; the full BASELINE sizes take minutes per replay).
usage: tools/profile_small.py aloha|mmc|archsim4"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import libcxxdes_b200 as L  # noqa: E402
name = sys.argv[1]
cfg = {"aloha": (L.AlohaParams(64, 50, 0.1, 3.0, 1.0, 100.0, 0), 148 * 384, ((1, L.SECONDS), (1, L.MILLISECONDS))),
       "mmc": (L.MMCParams(9.0, 1.0, 0.5, 8, 0, 300), 148 * 352, ((1, L.SECONDS), (1, L.MILLISECONDS))),
       "archsim4": (L.ArchSimParams(1, 10, 16, 32, 4, 0, 64), 148 * 384, ((1, L.SECONDS), (1, L.SECONDS)))}[name]
ens = L.Ensemble(cfg[0], cfg[1], seed=0x5EED)
ens.time_unit(*cfg[2][0]).time_precision(*cfg[2][1])
ens.run().sync()
acc, _ = ens.reduce()
print(f"{name} small: {cfg[1]} replications, {acc.n_events} events, {ens.last_run_ms():.3f} ms, {acc.n_events / ens.last_run_ms() / 1e6:.2f} Gev/s")
ens.close()
