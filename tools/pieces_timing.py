#!/usr/bin/env python3
"""run() against the same run cut into run_for pieces, for every model: the pieces continue from saved state, so
their total is the cost of the run plus the state traffic -- not the sum of replays from tick 0.
usage: tools/pieces_timing.py [n_pieces]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import libcxxdes_b200 as L  # noqa: E402
import kat_programs  # noqa: E402

S_MS = ((1, L.SECONDS), (1, L.MILLISECONDS))
S_S = ((1, L.SECONDS), (1, L.SECONDS))
CASES = [("mm1", L.MM1Params(5.0, 10.0, 10000), 1 << 18, S_MS), ("aloha", L.AlohaParams(64, 50, 0.1, 3.0, 1.0, 1000.0, 0), 1 << 16, S_MS),
         ("mmc", L.MMCParams(9.0, 1.0, 0.5, 8, 0, 1000), 1 << 18, S_MS), ("archsim", L.ArchSimParams(1, 10, 16, 32, 1, 0, 1024), 1 << 20, S_S),
         ("mm1 program", kat_programs.mm1_program(5.0, 10.0, 1000), 1 << 18, S_MS)]
n_pieces = int(sys.argv[1]) if len(sys.argv) > 1 else 8
for name, params, reps, clock in CASES:
    ens = L.Ensemble(params, reps, seed=1)
    ens.time_unit(*clock[0]).time_precision(*clock[1])
    ens.run().sync()
    ens.run().sync()
    full = ens.last_run_ms()
    whole, _ = ens.reduce()
    end = int(ens.fetch().end_time.max())
    ens.reset()
    total = 0.0
    for k in range(n_pieces):
        ens.run_for(end // n_pieces + 1).sync()
        total += ens.last_run_ms()
    acc, _ = ens.reduce()
    assert acc.n_events == whole.n_events and acc.trace_hash_sum == whole.trace_hash_sum, name
    print(f"{name}: {reps} replications, run() {full:.1f} ms, {n_pieces} run_for pieces {total:.1f} ms in total (same digest sum)", flush=True)
    ens.close()
