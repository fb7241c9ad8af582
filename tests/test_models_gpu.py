"""GPU parity tests (B200) for ALOHA, M/M/c with balking and the memory-hierarchy model: the CUDA
path through the C ABI vs the reference's golden vectors and the CPU oracle, bit-exact."""
import os

import numpy as np
import pytest

import golden_util
import oracles
import libcxxdes_b200 as L
from libcxxdes_b200 import abi
from model_cases import MODELS, N_GOLDEN_CASES, clock_of, clock_pairs

pytestmark = pytest.mark.gpu


def run_gpu(params, n_reps, clock, seed=0x5EED, first=0, until=-1, strict=True):
    ens = L.Ensemble(params, n_reps, seed=seed, first_replication=first)
    ens.time_unit(*clock[0]).time_precision(*clock[1])
    if until < 0:
        ens.run()
    else:
        ens.run_until(until)
    cols = ens.fetch(strict=strict)
    acc, _ = ens.reduce()
    ens.close()
    return cols, acc


@pytest.mark.parametrize("idx", range(N_GOLDEN_CASES))
@pytest.mark.parametrize("name", sorted(MODELS))
def test_gpu_matches_reference_golden(built, name, idx):
    model, cls, fname = MODELS[name]
    g = golden_util.load(fname)
    case = g["cases"][idx]
    got, _ = run_gpu(cls(*case["params"]), case["n_replications"], clock_pairs(g), seed=case["seed"],
                     first=case["first_replication"], until=case["run_until"], strict=False)
    assert not (got.status & ~np.uint32(abi.REP_UNFINISHED)).any()
    golden_util.assert_columns_equal(got, golden_util.golden_columns(case), what=f"{name} golden {idx}")


@pytest.mark.parametrize("name,fields,n_reps", [
    ("aloha", (64, 50, 0.1, 3.0, 1.0, 100.0, 0), 2000),      # the offered-load sweep, 64 stations
    ("aloha", (24, 0, 1.5, 1.5, 1.0, 1000.0, 60), 1024),
    ("aloha", (64, 0, 3.0, 3.0, 1.0, 1000.0, 40), 512),      # heavy load: many key ties in a 64-entry heap
    ("mmc", (9.0, 1.0, 0.5, 8, 0, 400), 1024),
    ("mmc", (12.0, 1.0, 0.5, 8, 0, 300), 1024),
    ("mmc", (3.0, 1.0, 0.2, 2, 0, 300), 512),
    ("archsim", (1, 10, 16, 32, 1, 0, 1024), 1024),          # the shipped sequential driver
    ("archsim", (1, 10, 16, 32, 4, 0, 256), 1024),           # 4 requesters: mutex contention
    ("archsim", (0, 0, 4, 8, 4, 0, 100), 512),               # zero latencies: everything ties
    ("archsim", (2, 7, 3, 16, 5, 0, 120), 512)])
def test_gpu_matches_oracle(port, name, fields, n_reps):
    model, cls, fname = MODELS[name]
    g = golden_util.load(fname)
    p = cls(*fields)
    got, acc = run_gpu(p, n_reps, clock_pairs(g), seed=0xFEED, first=17)
    want = port.run(model, p, clock_of(g), 0xFEED, 17, n_reps, threads=os.cpu_count())
    assert not got.status.any()
    golden_util.assert_columns_equal(got, want, what=f"{name} {fields}")
    assert acc.n_events == int(want.n_events.sum()) and acc.n_errors == 0
    assert acc.trace_hash_sum == int(want.trace_hash.sum(dtype=np.uint64))
    assert acc.i64_sum[0] == int(want.i64[0].sum()) and acc.i64_sum[1] == int(want.i64[1].sum())
    assert abs(acc.f64_sum[0] - want.f64[0].sum()) <= 1e-12 * max(1.0, abs(want.f64[0].sum()))


def test_aloha_full_size_properties(built):
    """BASELINE config 3 shape at reduced replication count: event identity per sweep step and the
    pure-ALOHA throughput curve S = G exp(-2G)."""
    p = abi.AlohaParams(64, 50, 0.1, 3.0, 1.0, 1000.0, 0)
    n_reps = 1000
    got, _ = run_gpu(p, n_reps, ((1, L.SECONDS), (1, L.MILLISECONDS)))
    assert not got.status.any()
    pps = got.i64[1]
    assert (got.n_events == (2 * 64 * pps + 2 * 64 + 3).astype(np.uint64)).all()
    for step in (2, 10, 25, 49):
        sel = np.arange(n_reps) % 50 == step
        G = got.f64[1][sel][0]
        assert abs(got.f64[0][sel].mean() - G * np.exp(-2 * G)) < 0.02


def test_mmc_conservation(built):
    p = abi.MMCParams(9.0, 1.0, 0.5, 8, 0, 1000)
    got, _ = run_gpu(p, 4096, ((1, L.SECONDS), (1, L.MILLISECONDS)))
    assert not got.status.any()
    assert (got.i64[0] + got.i64[1] == 1000).all()


def test_invalid_model_parameters(built):
    with pytest.raises(L.CxxdesError):
        L.Ensemble(abi.AlohaParams(65, 0, 1.0, 1.0, 1.0, 1000.0, 10), 4).run()
    with pytest.raises(L.CxxdesError):
        L.Ensemble(abi.MMCParams(9.0, 1.0, 0.5, 0, 0, 10), 4).run()
    with pytest.raises(L.CxxdesError):
        L.Ensemble(abi.ArchSimParams(1, 10, 16, 32, 17, 0, 10), 4).run()


@pytest.mark.parametrize("name,fields", [
    ("aloha", (24, 50, 0.1, 3.0, 1.0, 100.0, 0)),
    ("mmc", (9.0, 1.0, 0.5, 8, 0, 120)),
    ("archsim", (1, 10, 16, 32, 4, 0, 64))])
def test_run_until_pieces_then_run(port, name, fields):
    """The kernels that replay a replication from tick 0: run_until(t1), run_until(t2), ... equal the oracle's
    run_until at every horizon, and a final run() leaves the clock at max(natural end, last horizon) --
    run_until had left now_ there (environment.ipp:194) and run() only moves it forward."""
    model, cls, fname = MODELS[name]
    g = golden_util.load(fname)
    p = cls(*fields)
    n_reps = 600
    clock = clock_pairs(g)
    ends = port.run(model, p, clock_of(g), 77, 0, n_reps, threads=os.cpu_count()).end_time
    horizons = (int(ends.min()) // 2, int(np.quantile(ends, 0.3)), int(np.quantile(ends, 0.7)))   # the last two straddle the ends
    ens = L.Ensemble(p, n_reps, seed=77)
    ens.time_unit(*clock[0]).time_precision(*clock[1])
    for t in horizons:
        ens.run_until(t)
        got = ens.fetch(strict=False)
        want = port.run(model, p, clock_of(g), 77, 0, n_reps, threads=os.cpu_count(), until=t)
        assert not (got.status & ~np.uint32(abi.REP_UNFINISHED)).any()
        golden_util.assert_columns_equal(got, want, what=f"{name} run_until({t})")
    got = ens.run().fetch()
    want = port.run(model, p, clock_of(g), 77, 0, n_reps, threads=os.cpu_count())
    assert (want.end_time < horizons[-1]).any() and (want.end_time > horizons[0]).any()   # the case straddles the ends
    want.end_time[:] = np.maximum(want.end_time, horizons[-1])
    golden_util.assert_columns_equal(got, want, what=f"{name} run() after run_until")
    ens.close()


def test_aloha_run_for_in_many_pieces_continues_from_saved_state(port):
    """run_for(piece) x 12 on the compact ALOHA kernel: every replication that stops at the horizon is saved (heap,
    station masks, counters: 16 + 2 (N + 2) words) and the next launch continues from it (environment.ipp:190-214);
    columns equal the oracle's run_until(k * piece) after every piece, run() finishes the rest."""
    model, cls, fname = MODELS["aloha"]
    g = golden_util.load(fname)
    p = cls(24, 50, 0.1, 3.0, 1.0, 60.0, 0)
    n_reps = 500
    clock = clock_pairs(g)
    ends = port.run(model, p, clock_of(g), 9, 0, n_reps, threads=os.cpu_count()).end_time
    piece = int(ends.max()) // 11
    ens = L.Ensemble(p, n_reps, seed=9)
    ens.time_unit(*clock[0]).time_precision(*clock[1])
    for k in range(1, 13):
        ens.run_for(piece)
        got = ens.fetch(strict=False)
        want = port.run(model, p, clock_of(g), 9, 0, n_reps, threads=os.cpu_count(), until=k * piece)
        assert not (got.status & ~np.uint32(abi.REP_UNFINISHED)).any()
        golden_util.assert_columns_equal(got, want, what=f"aloha after piece {k}")
    assert not ens.fetch(strict=False).status.any()          # 12 pieces reach past the longest replication
    ens.reset()
    ens.run_for(piece)
    want = port.run(model, p, clock_of(g), 9, 0, n_reps, threads=os.cpu_count(), until=piece)
    golden_util.assert_columns_equal(ens.fetch(strict=False), want, what="aloha after reset")
    ens.close()


@pytest.mark.parametrize("fields", [(1, 10, 16, 32, 4, 0, 96), (1, 10, 16, 32, 1, 0, 300), (0, 0, 4, 8, 3, 0, 60)])
def test_archsim_run_for_in_many_pieces_continues_from_saved_state(port, fields):
    """run_for(piece) x 12 on the memory-hierarchy kernel (one and several requesters, zero latencies): a replication
    that stops at the horizon is saved -- heap, block store, per-requester words, both mutex wait lists -- and the
    next launch continues from it; columns equal the oracle's run_until(k * piece) after every piece."""
    model, cls, fname = MODELS["archsim"]
    g = golden_util.load(fname)
    p = cls(*fields)
    n_reps = 400
    clock = clock_pairs(g)
    ends = port.run(model, p, clock_of(g), 4, 0, n_reps, threads=os.cpu_count()).end_time
    piece = max(1, int(ends.max()) // 11)
    ens = L.Ensemble(p, n_reps, seed=4)
    ens.time_unit(*clock[0]).time_precision(*clock[1])
    for k in range(1, 13):
        ens.run_for(piece)
        got = ens.fetch(strict=False)
        want = port.run(model, p, clock_of(g), 4, 0, n_reps, threads=os.cpu_count(), until=k * piece)
        assert not (got.status & ~np.uint32(abi.REP_UNFINISHED)).any()
        golden_util.assert_columns_equal(got, want, what=f"archsim {fields} after piece {k}")
    ens.run()
    want = port.run(model, p, clock_of(g), 4, 0, n_reps, threads=os.cpu_count())
    want.end_time[:] = np.maximum(want.end_time, 12 * piece)
    golden_util.assert_columns_equal(ens.fetch(), want, what=f"archsim {fields}: run() after the pieces")
    ens.close()


@pytest.mark.parametrize("fields", [(9.0, 1.0, 0.5, 8, 0, 150), (3.0, 1.0, 0.2, 2, 0, 120)])
def test_mmc_run_for_in_many_pieces_continues_from_saved_state(port, fields):
    """run_for(piece) x 12 on the compact M/M/c kernel: a replication that stops at the horizon is saved -- heap,
    semaphore and its wait list, counters, the words of the customers created so far -- and the next launch
    continues from it; columns equal the oracle's run_until(k * piece) after every piece, run() finishes the rest."""
    model, cls, fname = MODELS["mmc"]
    g = golden_util.load(fname)
    p = cls(*fields)
    n_reps = 400
    clock = clock_pairs(g)
    ends = port.run(model, p, clock_of(g), 6, 0, n_reps, threads=os.cpu_count()).end_time
    piece = max(1, int(ends.max()) // 11)
    ens = L.Ensemble(p, n_reps, seed=6)
    ens.time_unit(*clock[0]).time_precision(*clock[1])
    for k in range(1, 13):
        ens.run_for(piece)
        got = ens.fetch(strict=False)
        want = port.run(model, p, clock_of(g), 6, 0, n_reps, threads=os.cpu_count(), until=k * piece)
        assert not (got.status & ~np.uint32(abi.REP_UNFINISHED)).any()
        golden_util.assert_columns_equal(got, want, what=f"mmc {fields} after piece {k}")
    ens.run()
    want = port.run(model, p, clock_of(g), 6, 0, n_reps, threads=os.cpu_count())
    want.end_time[:] = np.maximum(want.end_time, 12 * piece)
    golden_util.assert_columns_equal(ens.fetch(), want, what=f"mmc {fields}: run() after the pieces")
    ens.close()


@pytest.mark.parametrize("name,fields", [
    ("aloha", (32, 50, 0.1, 3.0, 1.0, 80.0, 0)),
    ("mmc", (9.0, 1.0, 0.5, 8, 0, 150)),
    ("archsim", (1, 10, 16, 32, 4, 0, 96))])
def test_sharding_is_invisible_for_every_model(built, name, fields):
    """Replications carry their GLOBAL id in the Philox key (and in the load sweep), so splitting an ensemble into
    shards -- one per GPU -- changes nothing: shard columns concatenate to the unsplit ensemble's, bit for bit."""
    model, cls, fname = MODELS[name]
    g = golden_util.load(fname)
    p = cls(*fields)
    whole, acc = run_gpu(p, 700, clock_pairs(g), seed=12, first=5)
    parts = [run_gpu(p, n, clock_pairs(g), seed=12, first=5 + first) for first, n in ((0, 123), (123, 400), (523, 177))]
    for column in ("end_time", "n_events", "trace_hash"):
        assert np.array_equal(np.concatenate([getattr(c, column) for c, _ in parts]), getattr(whole, column)), column
    for k in range(abi.N_F64):
        assert np.array_equal(np.concatenate([c.f64[k] for c, _ in parts]).view(np.uint64), whole.f64[k].view(np.uint64))
    assert sum(a.n_events for _, a in parts) == acc.n_events
    assert sum(a.trace_hash_sum for _, a in parts) % (1 << 64) == acc.trace_hash_sum


def test_archsim_full_size_properties(built):
    """BASELINE config 5 at its full size on one GPU (4 Mi replications x 1024 loads, the shipped sequential driver):
    every load is a hit or a miss, a miss costs mem1 + mem2 latency and a hit mem2's, and the event count is a
    function of the two (a hit takes 5 events, a miss 10; 5 more frame the replication)."""
    p = abi.ArchSimParams(1, 10, 16, 32, 1, 0, 1024)
    got, acc = run_gpu(p, 1 << 22, ((1, L.SECONDS), (1, L.SECONDS)))
    assert not got.status.any()
    hits, misses = got.i64[0], got.i64[1]
    assert ((hits + misses) == 1024).all()
    assert (got.f64[0] == (hits * 1 + misses * 11).astype(np.float64)).all()       # sum of load latencies in ticks
    assert (got.end_time == hits * 1 + misses * 11).all()                          # one requester: loads back to back
    assert (got.n_events.astype(np.int64) == 5 * hits + 10 * misses + 5).all()
    assert acc.n_events == int(got.n_events.sum())
    # addresses are uniform over 32 with a 16-block LRU store: the hit rate is near (but below) one half
    assert 0.40 < hits.mean() / 1024 < 0.5


@pytest.mark.parametrize("name,fields", [
    ("aloha", (4, 0, 2.0, 2.0, 1.0, 1000.0, 3)),       # 56 832 resident lanes
    ("mmc", (9.0, 1.0, 0.5, 8, 0, 6))])                 # 52 096 resident lanes
def test_pieces_on_an_ensemble_larger_than_the_grid(port, name, fields):
    """run_until / run_for / run() in pieces on several times more replications than resident lanes: in a later piece
    a lane can claim a replication that needs nothing (finished earlier, or nothing due) while the work list still
    holds replications further on; a warp whose 32 lanes all drew such replications must keep claiming, or the tail of
    the ensemble is never visited (finished rows miss the later horizon, running rows never finish)."""
    model, cls, fname = MODELS[name]
    g = golden_util.load(fname)
    p = cls(*fields)
    n = 400_000
    clock = clock_pairs(g)
    want, _ = run_gpu(p, n, clock, seed=21)
    ends = np.sort(want.end_time)
    t_half, t_most, t_all = int(ends[n // 2]), int(ends[(n * 97) // 100]), int(ends[-1]) + 10
    ens = L.Ensemble(p, n, seed=21)
    ens.time_unit(*clock[0]).time_precision(*clock[1])
    ens.run_until(t_half)
    ens.run_until(t_most)                       # 97 % are finished: most claims of the next pieces find nothing to do
    ens.run_for(1)
    got = ens.run_until(t_all).fetch(strict=False)
    assert not got.status.any() and (got.end_time == t_all).all()
    assert np.array_equal(got.trace_hash, want.trace_hash) and np.array_equal(got.n_events, want.n_events)
    ens.run_for(7)
    got = ens.run().fetch()
    ens.close()
    assert (got.end_time == t_all + 7).all() and np.array_equal(got.trace_hash, want.trace_hash)
    tail = port.run(model, p, clock_of(g), 21, n - 2048, 2048, threads=os.cpu_count())
    assert np.array_equal(tail.trace_hash, want.trace_hash[n - 2048:])
    assert np.array_equal(tail.end_time, want.end_time[n - 2048:])
