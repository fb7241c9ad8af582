"""Randomised GPU parity sweeps (fixed seeds) for ALOHA, M/M/c and the memory-hierarchy model against the
restated oracle: the compact 32-bit kernels, their hand-over to the 64-bit kernels (fine time precision
pushes the clock past 2^32 ticks), horizons, shard offsets."""
import numpy as np
import pytest

import golden_util
import oracles
import libcxxdes_b200 as L
from libcxxdes_b200 import abi

pytestmark = pytest.mark.gpu

CLOCKS = [((1, L.SECONDS), (1, L.MILLISECONDS)), ((1, L.SECONDS), (1, L.MICROSECONDS)), ((2, L.SECONDS), (5, L.MILLISECONDS)),
          ((1, L.SECONDS), (1, L.NANOSECONDS))]


def run_gpu(params, n_reps, clock, seed, first, until):
    ens = L.Ensemble(params, n_reps, seed=seed, first_replication=first)
    ens.time_unit(*clock[0]).time_precision(*clock[1])
    if until < 0:
        ens.run()
    else:
        ens.run_until(until)
    cols = ens.fetch(strict=False)
    ens.close()
    return cols


def check(model, params, rng, what, n_reps=160, capacity_may_overflow=False):
    clock = CLOCKS[int(rng.integers(len(CLOCKS)))]
    if model == abi.MODEL_ARCHSIM:
        clock = ((1, L.SECONDS), (1, L.SECONDS))     # integer latencies: the example's clock
    seed, first = int(rng.integers(1 << 40)), int(rng.integers(1 << 34))
    oclock = abi.Clock.of(*clock)
    full = oracles.port_oracle().run(model, params, oclock, seed, first, 8, threads=4)
    until = -1 if rng.random() < 0.5 else int(rng.uniform(0.0, 1.2) * float(full.end_time.max()))
    got = run_gpu(params, n_reps, clock, seed, first, until)
    want = oracles.port_oracle().run(model, params, oclock, seed, first, n_reps, threads=oracles.os.cpu_count(), until=until)
    bad = (got.status & ~np.uint32(abi.REP_UNFINISHED)) != 0
    if bad.any():
        # the only acceptable flags: a bounded device container was too small for this replication
        # (reference containers are unbounded); the row is then reported as an error, never compared
        assert capacity_may_overflow, what
        assert ((got.status[bad] & ~np.uint32(abi.REP_UNFINISHED | abi.REP_HEAP_OVERFLOW | abi.REP_WAITER_OVERFLOW)) == 0).all()
        keep = ~bad
        for name in ("end_time", "n_events", "trace_hash"):
            assert np.array_equal(getattr(got, name)[keep], getattr(want, name)[keep]), f"{what}: {name}"
        return
    golden_util.assert_columns_equal(got, want, what=f"{what} clock={clock} until={until}")


@pytest.mark.parametrize("case", range(14))
def test_aloha_fuzz(built, case):
    rng = np.random.default_rng(7000 + case)
    stations = int(rng.choice([1, 2, 3, 16, 40, 64]))
    sweep = int(rng.choice([0, 0, 7]))
    lam0, lam1 = float(rng.uniform(0.1, 3.0)), float(rng.uniform(0.1, 3.0))
    pps = int(rng.choice([1, 2, 9, 40]))
    frame = float(rng.choice([1.0, 0.25, 0.0]))
    p = abi.AlohaParams(stations, sweep, lam0, lam1, frame, float(rng.choice([10.0, 30.0])), pps)
    check(abi.MODEL_ALOHA, p, rng, f"aloha N={stations} sweep={sweep} lam={lam0:.3f}..{lam1:.3f} pps={pps} frame={frame}")


@pytest.mark.parametrize("case", range(14))
def test_mmc_fuzz(built, case):
    rng = np.random.default_rng(8000 + case)
    lam = float(rng.choice([0.5, 3.0, 9.0, 14.0]) * rng.uniform(0.9, 1.1))
    patience = float(rng.choice([0.0, 0.05, 0.5, 3.0]))
    servers = int(rng.choice([1, 2, 8]))
    n = int(rng.choice([1, 2, 30, 400]))
    p = abi.MMCParams(lam, 1.0, patience, servers, 0, n)
    # lambda * patience customers wait at once, each with a patience token and a wake-up token per release: beyond ~25
    # the 96-entry heap of the 64-bit kernel fills up, and so does its wait list when one server faces many arrivals
    # per service -- those replications take the last pass (heap and wait list in global memory, sized for the
    # model's worst case) and must equal the oracle like all others: no status flag is acceptable
    # (the test files that ran on a B200 before that pass existed run without it, conftest.py: there the flags are accepted)
    import os
    check(abi.MODEL_MMC, p, rng, f"mmc lam={lam:.3f} patience={patience} c={servers} n={n}",
          capacity_may_overflow=os.environ.get("CXXDES_B200_DEEP_BYTES") == "0" and lam * patience > 25.0)


@pytest.mark.parametrize("case", range(12))
def test_archsim_fuzz(built, case):
    rng = np.random.default_rng(9000 + case)
    p = abi.ArchSimParams(int(rng.integers(0, 4)), int(rng.integers(0, 12)), int(rng.choice([1, 3, 16])),
                          int(rng.choice([1, 8, 32, 200])), int(rng.choice([1, 2, 4, 7])), 0, int(rng.choice([1, 10, 150])))
    check(abi.MODEL_ARCHSIM, p, rng, f"archsim {tuple(getattr(p, f[0]) for f in p._fields_)}")
