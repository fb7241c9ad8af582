"""GPU parity tests (B200): the CUDA path through the C ABI vs the CPU oracle and the reference's
golden vectors.  Integer time, event counts and trace digests must be bit-exact; the double
statistics are compared as bit patterns too (tolerance 0 -- stricter than the 1e-12 the task allows)."""
import numpy as np
import pytest

import golden_util
import oracles
import libcxxdes_b200 as L
from libcxxdes_b200 import abi

pytestmark = pytest.mark.gpu


def run_gpu(params, n_reps, seed=0x5EED, first=0, until=-1, flags=0, queue_capacity=0, clock=None, strict=True):
    ens = L.Ensemble(params, n_reps, seed=seed, first_replication=first, flags=flags, queue_capacity=queue_capacity)
    unit, prec = clock or ((1, L.SECONDS), (1, L.MILLISECONDS))
    ens.time_unit(*unit).time_precision(*prec)
    if until < 0:
        ens.run()
    else:
        ens.run_until(until)
    cols = ens.fetch(strict=strict)
    acc, _ = ens.reduce()
    ens.close()
    return cols, acc


@pytest.mark.parametrize("flags", [0, abi.FLAG_GENERIC_ENGINE], ids=["fused", "engine"])
@pytest.mark.parametrize("idx", range(12))
def test_gpu_matches_reference_golden(built, idx, flags):
    g = golden_util.load("mm1_reference.json")
    case = g["cases"][idx]
    p = abi.MM1Params(case["lambda"], case["mu"], case["n_packets"])
    got, _ = run_gpu(p, case["n_replications"], seed=case["seed"], first=case["first_replication"],
                     until=case["run_until"], flags=flags)
    golden_util.assert_columns_equal(got, golden_util.golden_columns(case), what=f"golden case {idx}")


@pytest.mark.parametrize("flags", [0, abi.FLAG_GENERIC_ENGINE], ids=["fused", "engine"])
@pytest.mark.parametrize("lam,n_packets,n_reps", [(5.0, 2000, 4096), (0.5, 1000, 2048), (9.0, 1500, 2048),
                                                  (9.9, 2000, 1024), (12.0, 600, 512)])
def test_gpu_matches_oracle(port, lam, n_packets, n_reps, flags):
    p = abi.MM1Params(lam, 10.0, n_packets)
    got, acc = run_gpu(p, n_reps, flags=flags)
    want = port.run(abi.MODEL_MM1, p, oracles.MM1_CLOCK, 0x5EED, 0, n_reps, threads=oracles.os.cpu_count())
    assert not got.status.any()
    golden_util.assert_columns_equal(got, want, what=f"lambda={lam}")
    # accumulators: integer sums exact, checksum of checksums, FP sums to 1e-12 relative
    assert acc.n_replications == n_reps and acc.n_errors == 0
    assert acc.n_events == int(want.n_events.sum())
    assert acc.end_time_sum == int(want.end_time.sum())
    assert acc.trace_hash_sum == int(want.trace_hash.sum(dtype=np.uint64))
    assert acc.i64_sum[0] == int(want.i64[0].sum())
    assert abs(acc.f64_sum[0] - want.f64[0].sum()) <= 1e-12 * abs(want.f64[0].sum())
    assert abs(acc.f64_sumsq[0] - (want.f64[0] ** 2).sum()) <= 1e-12 * (want.f64[0] ** 2).sum()


def test_config1_single_long_replication(port):
    """BASELINE config 1 through the GPU path: examples/producer_consumer.cpp with n_packets = 100000,
    avg_latency near 1/(mu - lambda) = 0.2 s and bit-equal to the CPU run."""
    p = abi.MM1Params(5.0, 10.0, 100000)
    got, _ = run_gpu(p, 3, seed=0)
    want = port.run(abi.MODEL_MM1, p, oracles.MM1_CLOCK, 0, 0, 3, threads=3)
    assert not got.status.any()
    golden_util.assert_columns_equal(got, want, what="config 1")
    assert abs(got.f64[0][0] - 0.2) < 0.006


def test_heavy_load_full_length_uses_the_second_ring_level(port):
    """rho = 0.95 at the full 10000 packets: queues of 50-150 packets live in the global-memory
    level of the ring for most of the run."""
    p = abi.MM1Params(9.5, 10.0, 10000)
    got, _ = run_gpu(p, 192, seed=4)
    want = port.run(abi.MODEL_MM1, p, oracles.MM1_CLOCK, 4, 0, 192, threads=oracles.os.cpu_count())
    assert (want.i64[1] > 40).mean() > 0.9          # every queue leaves the 31-packet window
    assert not got.status.any()
    golden_util.assert_columns_equal(got, want, what="rho 0.95")


@pytest.mark.parametrize("lam,n_packets", [(5.0, 12000), (9.0, 14000)])
def test_fine_time_precision_stays_on_the_fast_kernel(port, lam, n_packets):
    """time_unit 1 s with time_precision 1 us: the clock passes 2^30 ticks, the generate-ahead kernel
    keeps its 32-bit times relative to a per-lane epoch (MODE_WIDE); run(), a horizon beyond 2^30 and
    run_for pieces across the moves of the epoch all equal the oracle."""
    clock = ((1, L.SECONDS), (1, L.MICROSECONDS))
    oclock = abi.Clock.of(*clock)
    p = abi.MM1Params(lam, 10.0, n_packets)
    got, _ = run_gpu(p, 256, seed=23, clock=clock)
    want = port.run(abi.MODEL_MM1, p, oclock, 23, 0, 256, threads=oracles.os.cpu_count())
    assert want.end_time.max() > 2 ** 30 and not got.status.any()
    golden_util.assert_columns_equal(got, want, what="us precision, run()")
    got, _ = run_gpu(p, 256, seed=23, clock=clock, until=1_500_000_000, strict=False)
    want = port.run(abi.MODEL_MM1, p, oclock, 23, 0, 256, threads=oracles.os.cpu_count(), until=1_500_000_000)
    golden_util.assert_columns_equal(got, want, what="us precision, run_until(1.5e9)")
    ens = L.Ensemble(p, 256, seed=23)
    ens.time_unit(*clock[0]).time_precision(*clock[1])
    for _ in range(4):
        ens.run_for(400_000_000)
    got = ens.fetch(strict=False)
    ens.close()
    want = port.run(abi.MODEL_MM1, p, oclock, 23, 0, 256, threads=oracles.os.cpu_count(), until=1_600_000_000)
    golden_util.assert_columns_equal(got, want, what="us precision, 4 x run_for(4e8)")


def test_small_ring_takes_the_retry_path(port):
    """A 4-entry shared ring overflows in most replications: they are re-run by the global-ring
    kernel and must still match the oracle bit for bit."""
    p = abi.MM1Params(8.0, 10.0, 1200)
    got, _ = run_gpu(p, 3000, queue_capacity=4)
    want = port.run(abi.MODEL_MM1, p, oracles.MM1_CLOCK, 0x5EED, 0, 3000, threads=oracles.os.cpu_count())
    assert (want.i64[1] > 4).mean() > 0.5          # the scenario really overflows
    assert not got.status.any()
    golden_util.assert_columns_equal(got, want, what="retry path")


def test_tick_range_guard_falls_back_to_64bit(port):
    """Very slow arrivals push the clock past 2^32 ticks; the 32-bit fast kernel must hand those
    replications to the 64-bit kernel."""
    p = abi.MM1Params(0.00002, 0.001, 400)
    got, _ = run_gpu(p, 512)
    want = port.run(abi.MODEL_MM1, p, oracles.MM1_CLOCK, 0x5EED, 0, 512, threads=oracles.os.cpu_count())
    assert want.end_time.max() > 2 ** 32
    assert not got.status.any()
    golden_util.assert_columns_equal(got, want, what="64-bit ticks")


def test_sharding_is_invisible(port):
    """Replication r of a shard starting at `first` equals replication first+r of one big run."""
    p = abi.MM1Params(5.0, 10.0, 500)
    whole, _ = run_gpu(p, 1024)
    for first, n in [(0, 300), (300, 500), (800, 224)]:
        part, _ = run_gpu(p, n, first=first)
        assert np.array_equal(part.trace_hash, whole.trace_hash[first:first + n])
        assert np.array_equal(part.end_time, whole.end_time[first:first + n])


@pytest.mark.parametrize("until", [0, 1, 999, 20000, 10 ** 7])
def test_run_until_matches_oracle(port, until):
    p = abi.MM1Params(5.0, 10.0, 800)
    got, _ = run_gpu(p, 256, seed=5, until=until, strict=False)
    want = port.run(abi.MODEL_MM1, p, oracles.MM1_CLOCK, 5, 0, 256, threads=4, until=until)
    golden_util.assert_columns_equal(got, want, what=f"until={until}")


def test_run_for_accumulates(port):
    """run_for(5 s) twice == run_until(10 s) (tests/process.test.cpp:127-147 pattern)."""
    p = abi.MM1Params(5.0, 10.0, 800)
    ens = L.Ensemble(p, 128, seed=5)
    ens.time_unit(1, L.SECONDS).time_precision(1, L.MILLISECONDS)
    ens.run_for(5000)
    a = ens.fetch(strict=False)
    assert (a.end_time == 5000).all()
    ens.run_for(5000)
    b = ens.fetch(strict=False)
    ens.close()
    want = port.run(abi.MODEL_MM1, p, oracles.MM1_CLOCK, 5, 0, 128, threads=4, until=10000)
    golden_util.assert_columns_equal(b, want, what="run_for x2")


@pytest.mark.parametrize("lam,n_packets,piece,n_pieces,kw", [
    (5.0, 800, 3000, 12, {}),                       # steady state cut 12 times
    (5.0, 800, 1, 40, {}),                          # horizons inside the start-up events
    (9.0, 1500, 20000, 6, {}),                      # spilling ring: entries beyond the window are drawn again
    (8.0, 1200, 15000, 8, {"queue_capacity": 4}),   # replications handed to the fallback kernel stay there
    (5.0, 800, 3000, 5, {"flags": abi.FLAG_MM1_LOCKSTEP})])   # kernels without saved state replay from tick 0
def test_run_for_in_pieces_matches_oracle(port, lam, n_packets, piece, n_pieces, kw):
    """run_for(piece) x n (environment.ipp:209-214, tests/process.test.cpp:127-147 pattern): after every
    piece the columns equal the oracle's run_until(k * piece); then run() finishes the replications."""
    p = abi.MM1Params(lam, 10.0, n_packets)
    n_reps = 512
    ens = L.Ensemble(p, n_reps, seed=31, **kw)
    ens.time_unit(1, L.SECONDS).time_precision(1, L.MILLISECONDS)
    for k in range(1, n_pieces + 1):
        ens.run_for(piece)
        if k in (1, 2, n_pieces // 2, n_pieces):
            got = ens.fetch(strict=False)
            want = port.run(abi.MODEL_MM1, p, oracles.MM1_CLOCK, 31, 0, n_reps, threads=oracles.os.cpu_count(),
                            until=k * piece)
            golden_util.assert_columns_equal(got, want, what=f"after piece {k}")
    ens.run()
    got = ens.fetch()
    ens.close()
    want = port.run(abi.MODEL_MM1, p, oracles.MM1_CLOCK, 31, 0, n_reps, threads=oracles.os.cpu_count())
    assert not got.status.any()
    golden_util.assert_columns_equal(got, want, what="run() after the pieces")


def test_time_series_from_pieces(port):
    """stats.time_series: per-interval statistics from run_for pieces equal the differences of the oracle's
    run_until columns; batch means with the transient dropped land on the M/M/1 latency 1/(mu - lambda)."""
    from libcxxdes_b200 import stats
    p = abi.MM1Params(5.0, 10.0, 4000)
    n_reps, piece, n_pieces = 2048, 50000, 10
    ens = L.Ensemble(p, n_reps, seed=13)
    ens.time_unit(1, L.SECONDS).time_precision(1, L.MILLISECONDS)
    rows = stats.time_series(ens, piece, n_pieces)
    ens.close()
    prev = None
    for j in (1, 2, 7):
        want = port.run(abi.MODEL_MM1, p, oracles.MM1_CLOCK, 13, 0, n_reps, threads=oracles.os.cpu_count(), until=j * piece)
        cum_events = sum(r["events"] for r in rows[:j])
        cum_count = sum(r["d_count"] for r in rows[:j])
        assert cum_events == int(want.n_events.sum()) and cum_count == int(want.i64[0].sum())
        assert abs(sum(r["d_stat"] for r in rows[:j]) - want.f64[1].sum()) <= 1e-9 * want.f64[1].sum()
    mean, half = stats.batch_means(rows, warmup=1)
    assert abs(mean - 0.2) < 0.01 and half < 0.01


def test_finished_replications_follow_later_horizons(port):
    """run_until(t) sets now_ = max(now_, t) even when no event is left (environment.ipp:194): a replication
    that finished in an earlier piece still reports the later horizon as its end time."""
    p = abi.MM1Params(9.0, 10.0, 3)
    ens = L.Ensemble(p, 64, seed=2)
    ens.time_unit(1, L.SECONDS).time_precision(1, L.MILLISECONDS)
    ens.run_for(5000)
    ens.run_for(5000)
    got = ens.fetch(strict=False)
    ens.close()
    want = port.run(abi.MODEL_MM1, p, oracles.MM1_CLOCK, 2, 0, 64, threads=4, until=10000)
    assert (want.end_time == 10000).all()
    golden_util.assert_columns_equal(got, want, what="finished earlier")


def test_reset_forgets_saved_state(port):
    p = abi.MM1Params(5.0, 10.0, 600)
    ens = L.Ensemble(p, 128, seed=8)
    ens.time_unit(1, L.SECONDS).time_precision(1, L.MILLISECONDS)
    ens.run_for(4000)
    ens.reset()
    ens.run_for(1000)
    got = ens.fetch(strict=False)
    ens.close()
    want = port.run(abi.MODEL_MM1, p, oracles.MM1_CLOCK, 8, 0, 128, threads=4, until=1000)
    golden_util.assert_columns_equal(got, want, what="after reset")


def test_other_clocks(port):
    p = abi.MM1Params(5.0, 10.0, 300)
    for unit, prec in [((1, L.MILLISECONDS), (1, L.MICROSECONDS)), ((2, L.SECONDS), (5, L.MILLISECONDS))]:
        got, _ = run_gpu(p, 128, seed=9, clock=(unit, prec))
        want = port.run(abi.MODEL_MM1, p, abi.Clock.of(unit, prec), 9, 0, 128)
        golden_util.assert_columns_equal(got, want, what=f"clock {unit}/{prec}")


def test_full_size_properties():
    """BASELINE config 2 shape at reduced replication count but full packet count: size-independent
    properties (event-count identity, all packets consumed, latency near theory, determinism)."""
    n_packets, n_reps = 10000, 16384
    p = abi.MM1Params(5.0, 10.0, n_packets)
    a, acc = run_gpu(p, n_reps)
    assert not a.status.any()
    assert (a.i64[0] == n_packets).all()
    assert (a.n_events >= 2 * n_packets + 6).all() and (a.n_events <= 3 * n_packets + 6).all()
    assert abs(a.f64[0].mean() - 0.2) < 0.004          # 1/(mu-lambda)
    assert abs(a.n_events.mean() / n_packets - 2.5) < 0.01
    b, acc2 = run_gpu(p, n_reps)                       # idempotent: same seeds, same bits
    assert np.array_equal(a.trace_hash, b.trace_hash)
    assert acc.trace_hash_sum == acc2.trace_hash_sum and acc.f64_sum[0] == acc2.f64_sum[0]
    assert len(np.unique(a.trace_hash)) == n_reps      # replications are independent streams


def test_errors_are_reported():
    with pytest.raises(L.CxxdesError):
        L.Ensemble(L.MM1Params(-1.0, 10.0, 10), 4).run()
    with pytest.raises(L.CxxdesError):
        L.Ensemble(L.MM1Params(5.0, 10.0, 10), 0).run()
    ens = L.Ensemble(L.MM1Params(5.0, 10.0, 10), 4)
    ens._ensure()
    with pytest.raises(L.CxxdesError):
        ens.fetch()                                    # fetch before run
    ens.close()


# ---------------------------------------------------------------------------------------------
# BASELINE configs[1] EXACTLY as benchmarked: lambda = 5, mu = 10, 10 000 packets, seed 0x5EED, 1 ms precision,
# 1 048 576 replications -- i.e. with more replications than resident lanes (work stealing, the second-pass
# list and its helper all engaged), not a reduced copy of it.
# ---------------------------------------------------------------------------------------------
def _window_vs_oracle(port, p, cols, base, first, n, seed, what):
    want = port.run(abi.MODEL_MM1, p, oracles.MM1_CLOCK, seed, first, n, threads=oracles.os.cpu_count())
    lo = first - base
    got = abi.HostColumns(n)
    got.end_time[:] = cols.end_time[lo:lo + n]
    got.n_events[:] = cols.n_events[lo:lo + n]
    got.trace_hash[:] = cols.trace_hash[lo:lo + n]
    for k in range(abi.N_F64):
        got.f64[k][:] = cols.f64[k][lo:lo + n]
    for k in range(abi.N_I64):
        got.i64[k][:] = cols.i64[k][lo:lo + n]
    golden_util.assert_columns_equal(got, want, what=what)


def test_headline_config_full_size_against_the_oracle(port):
    """All 1 048 576 replications of the headline run: (1) the device accumulators equal the checksum of checksums
    the CPU oracle produced for every replication (tests/golden/mm1_config2_full.json, tools/gen_full_parity.py);
    (2) so does every one of its 64 blocks of 16 384 replications, from the fetched columns; (3) 5 632 replications
    in windows of global ids -- the first lanes' first replications, the ids around 113 664 where lanes start to
    steal their second replication, the longest queue of the shard, the last ids -- are compared column by column,
    bit for bit, with the oracle run here."""
    full = golden_util.load("mm1_config2_full.json")
    n, seed = full["n_replications"], full["seed"]
    p = abi.MM1Params(full["lambda"], full["mu"], full["n_packets"])
    ens = L.Ensemble(p, n, seed=seed)
    ens.time_unit(1, L.SECONDS).time_precision(1, L.MILLISECONDS)
    ens.run()
    cols = ens.fetch()
    acc, _ = ens.reduce()
    ens.close()
    assert not cols.status.any() and acc.n_errors == 0
    assert acc.n_events == full["n_events_sum"]
    assert acc.end_time_sum == full["end_time_sum"]
    assert acc.trace_hash_sum == full["trace_hash_sum"]
    assert acc.i64_sum[0] == full["packets_sum"] and acc.i64_sum[1] == full["max_queue_sum"]
    for b in full["blocks"]:
        s = slice(b["first"], b["first"] + b["n"])
        assert int(cols.n_events[s].sum(dtype=np.uint64)) == b["n_events_sum"], f"block at {b['first']}: n_events"
        assert int(cols.end_time[s].sum()) == b["end_time_sum"], f"block at {b['first']}: end_time"
        assert int(cols.trace_hash[s].sum(dtype=np.uint64)) == b["trace_hash_sum"], f"block at {b['first']}: digest"
    for q in full["long_queue"]:
        assert int(cols.i64[1][q["replication"]]) == q["max_queue"]
    longest = max(full["long_queue"], key=lambda q: q["max_queue"])["replication"]
    for first, cnt in [(0, 1536), (113664 - 512, 1536), (longest - 256, 512), (n // 2, 1024), (n - 1024, 1024)]:
        _window_vs_oracle(port, p, cols, 0, first, cnt, seed, what=f"headline window at {first}")


def test_headline_second_pass_runs_next_to_the_first(port):
    """The replications of the headline ensemble whose queue outgrows the 31-packet window (ids 3 850 044 and
    4 042 079 of the 1 Mi..4 Mi shard a rank 3 of the weak-scaling bench owns: max queue 32 and 33) are redone with
    the two-level ring -- by the helper grid, while the first pass runs -- and equal the oracle bit for bit."""
    p = abi.MM1Params(5.0, 10.0, 10000)
    first, n = 3_800_000, 262_144          # 2.3 rounds of the 113 664 resident lanes
    ens = L.Ensemble(p, n, seed=0x5EED, first_replication=first)
    ens.time_unit(1, L.SECONDS).time_precision(1, L.MILLISECONDS)
    ens.run()
    cols = ens.fetch()
    counts = ens.pass_counts()
    ens.close()
    assert not cols.status.any()
    assert counts["second_pass"] >= 2 and counts["fallback"] == 0
    assert counts["helper"] == counts["second_pass"]       # nothing was left for after the first pass
    for rep in (3_850_044, 4_042_079):
        _window_vs_oracle(port, p, cols, first, rep - 64, 128, 0x5EED, what=f"around replication {rep}")
        assert cols.i64[1][rep - first] >= 32
    _window_vs_oracle(port, p, cols, first, first, 256, 0x5EED, what="first ids of the shard")


def test_second_pass_list_longer_than_the_helper(port):
    """Many replications outgrow a small window: the helper takes what it can while the first pass runs and the
    pass that follows takes the rest of the same list; every replication is done exactly once."""
    p = abi.MM1Params(7.0, 10.0, 3000)
    n = 150_000
    ens = L.Ensemble(p, n, seed=77, queue_capacity=16)
    ens.time_unit(1, L.SECONDS).time_precision(1, L.MILLISECONDS)
    ens.run()
    cols = ens.fetch()
    counts = ens.pass_counts()
    acc, _ = ens.reduce()
    ens.close()
    assert counts["second_pass"] > 1000 and counts["helper"] >= 1
    assert not cols.status.any() and acc.n_errors == 0
    assert (cols.i64[0] == 3000).all()
    want = port.run(abi.MODEL_MM1, p, oracles.MM1_CLOCK, 77, 0, 4096, threads=oracles.os.cpu_count())
    assert (want.i64[1] > 16).sum() > 20
    _window_vs_oracle(port, p, cols, 0, 0, 4096, 77, what="small window")
    _window_vs_oracle(port, p, cols, 0, n - 2048, 2048, 77, what="small window, last ids")


def test_pieces_on_an_ensemble_larger_than_the_grid(port):
    """run_until(past the end), run_for, run() on more replications than resident lanes: every lane of a warp can
    claim a replication that needs nothing in this piece (finished earlier) while the work list still holds
    replications further on -- the warp must keep claiming.  Compared with one plain run() and the oracle."""
    p = abi.MM1Params(5.0, 10.0, 12)
    n = 1_000_000                                      # more than 8 claims x the 113 664 resident lanes
    ref = L.Ensemble(p, n, seed=3)
    ref.time_unit(1, L.SECONDS).time_precision(1, L.MILLISECONDS)
    want = ref.run().fetch()
    ref.close()
    ens = L.Ensemble(p, n, seed=3)
    ens.time_unit(1, L.SECONDS).time_precision(1, L.MILLISECONDS)
    ens.run_until(2400)                                # about half of the replications are finished by now
    half = ens.fetch(strict=False)
    assert 0.2 < (half.status & abi.REP_UNFINISHED).astype(bool).mean() < 0.8
    ens.run_until(10 ** 6)                             # past every end: the rest finishes, clocks go to 10^6
    a = ens.fetch(strict=False)
    assert not (a.status & abi.REP_UNFINISHED).any() and (a.end_time == 10 ** 6).all()
    assert np.array_equal(a.trace_hash, want.trace_hash) and np.array_equal(a.n_events, want.n_events)
    ens.run_for(500)
    b = ens.fetch(strict=False)
    assert (b.end_time == 10 ** 6 + 500).all() and np.array_equal(b.trace_hash, want.trace_hash)
    ens.run()
    c = ens.fetch()
    acc, _ = ens.reduce()
    ens.close()
    assert np.array_equal(c.trace_hash, want.trace_hash) and (c.end_time == 10 ** 6 + 500).all()
    assert acc.n_errors == 0
    _window_vs_oracle(port, p, want, 0, n - 4096, 4096, 3, what="last ids")


def test_clock_follows_run_until_after_the_end(port):
    """run(); run_until(t); run_for(dt): nothing is left to dispatch, but now_ = max(now_, t) and run_for adds to
    each replication's own clock (environment.ipp:190-196,205-209)."""
    p = abi.MM1Params(5.0, 10.0, 30)
    ens = L.Ensemble(p, 300, seed=4)
    ens.time_unit(1, L.SECONDS).time_precision(1, L.MILLISECONDS)
    natural = ens.run().fetch().end_time.copy()
    t = int(np.median(natural))
    after = ens.run_until(t).fetch().end_time
    assert np.array_equal(after, np.maximum(natural, t))
    later = ens.run_for(250).fetch().end_time
    assert np.array_equal(later, np.maximum(natural, t) + 250)
    with pytest.raises(L.CxxdesError):
        L.Ensemble(p, 4).run_until(5).run_for(2 ** 63 - 3)
    ens.close()


def test_unfinished_replications_are_not_errors():
    p = abi.MM1Params(5.0, 10.0, 500)
    ens = L.Ensemble(p, 256, seed=6)
    ens.time_unit(1, L.SECONDS).time_precision(1, L.MILLISECONDS)
    ens.run_until(2000)
    cols = ens.fetch(strict=False)
    acc, _ = ens.reduce()
    ens.close()
    assert (cols.status & abi.REP_UNFINISHED).all() and acc.n_errors == 0
