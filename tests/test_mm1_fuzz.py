"""Randomised parity sweeps for the M/M/1 path (fixed seeds: deterministic).

CPU: the host-compiled lane (tests/native) against the oracle.  GPU: the C ABI against the oracle.
Parameters are drawn to hit the corners: tiny models, heavy load, horizons at arbitrary ticks, run_for
pieces, small caller-imposed rings, fine time precision (wide-clock mode), non-unit tick multipliers."""
import numpy as np
import pytest

import golden_util
import oracles
import libcxxdes_b200 as L
from libcxxdes_b200 import abi
from test_mm1_lane_host import lane_lib, masked_equal, run_lanes   # noqa: F401  (fixture + helpers)

CLOCKS = [
    (((1, L.SECONDS), (1, L.MILLISECONDS)), (1000, 1, 1e-3)),
    (((1, L.SECONDS), (1, L.MICROSECONDS)), (1000000, 1, 1e-6)),
    (((2, L.SECONDS), (5, L.MILLISECONDS)), (400, 5, 1e-3)),
    (((1, L.MILLISECONDS), (1, L.MICROSECONDS)), (1000, 1, 1e-6)),
]


def draw_case(rng):
    lam = float(np.round(rng.choice([0.3, 2.0, 5.0, 7.5, 9.0, 9.7, 11.0]) * rng.uniform(0.9, 1.1), 3))
    n = int(rng.choice([2, 3, 5, 17, 200, 1500, 4000]))
    clock_idx = int(rng.integers(len(CLOCKS)))
    tpu = CLOCKS[clock_idx][1][0]
    end_guess = n / lam * tpu
    until = -1 if rng.random() < 0.5 else int(rng.uniform(0, 1.3) * end_guess)
    pieces = 0 if rng.random() < 0.5 else int(rng.integers(1, 6))
    return lam, n, clock_idx, until, pieces, int(rng.integers(1 << 30))


@pytest.mark.parametrize("case", range(60))
def test_lane_fuzz(lane_lib, port, case):   # noqa: F811
    rng = np.random.default_rng(1000 + case)
    lam, n, ci, until, pieces, seed = draw_case(rng)
    (unit, prec), folded = CLOCKS[ci]
    p = abi.MM1Params(lam, 10.0, n)
    spill = 1024 if lam > 6.0 or rng.random() < 0.3 else 0
    wide = 1 if folded[0] >= 1000000 or rng.random() < 0.3 else 0
    piece = 0 if not pieces else max(1, int(n / lam * folded[0] / pieces))
    got, _ = run_lanes(lane_lib, p, 40, seed=seed, until=until, piece=piece, global_slots=spill, clock=folded, wide=wide,
                       event_slots=int(rng.integers(1, 7)))
    want = port.run(abi.MODEL_MM1, p, abi.Clock.of(unit, prec), seed, 0, 40, threads=4, until=until)
    ok = masked_equal(got, want, f"lam={lam} n={n} clock={ci} until={until} piece={piece} spill={spill} wide={wide}")
    # only "redo me on the 64-bit kernel" flags are acceptable
    assert ((got.status[~ok] & (abi.REP_QUEUE_OVERFLOW | abi.REP_TIME_RANGE)) != 0).all()
    if spill and not wide:
        assert ok.all()


@pytest.mark.gpu
@pytest.mark.parametrize("case", range(40))
def test_gpu_fuzz(port, case):
    rng = np.random.default_rng(5000 + case)
    lam, n, ci, until, pieces, seed = draw_case(rng)
    (unit, prec), folded = CLOCKS[ci]
    p = abi.MM1Params(lam, 10.0, n)
    n_reps = 300
    cap = int(rng.choice([0, 0, 0, 3, 6]))
    first = int(rng.integers(0, 1 << 33))
    ens = L.Ensemble(p, n_reps, seed=seed, first_replication=first, queue_capacity=cap)
    ens.time_unit(*unit).time_precision(*prec)
    if pieces and until > 0:
        step = max(1, until // pieces)
        for _ in range(pieces):
            ens.run_for(step)
        until = step * pieces
    elif until >= 0:
        ens.run_until(until)
    else:
        ens.run()
    got = ens.fetch(strict=False)
    ens.close()
    want = port.run(abi.MODEL_MM1, p, abi.Clock.of(unit, prec), seed, first, n_reps, threads=oracles.os.cpu_count(),
                    until=until)
    assert not (got.status & ~np.uint32(abi.REP_UNFINISHED)).any()
    golden_util.assert_columns_equal(got, want, what=f"lam={lam} n={n} clock={ci} until={until} pieces={pieces} cap={cap}")
