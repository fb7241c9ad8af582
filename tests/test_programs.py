"""Process programs (generic lowering): the reference's own known-answer scenarios
(tests/controlflow.test.cpp, tests/process.test.cpp, examples/{clocks,resource,mutex,semaphore,queue,event}.cpp)
run natively on the reference engine == as process programs on the CPU interpreter == on the CUDA interpreter."""
import ctypes as C
import os

import numpy as np
import pytest

import golden_util
import kat_programs as K
import oracles
import libcxxdes_b200 as L
from libcxxdes_b200 import abi

CLK = abi.Clock.of((1, 0), (1, 0))


def _golden(k):
    g = golden_util.load("kat_reference.json")
    return next(c for c in g["cases"] if c["scenario"] == k)


@pytest.mark.parametrize("k", sorted(K.SCENARIOS))
def test_cpu_interpreter_matches_reference_golden(port, k):
    prog, until = K.SCENARIOS[k]()
    pp = prog.to_params()
    case = _golden(k)
    got = port.run(abi.MODEL_PROGRAM, pp, CLK, 1, 0, 1, until=until)
    golden_util.assert_columns_equal(got, golden_util.golden_columns(case), what=case["name"])
    total, events = port.trace(abi.MODEL_PROGRAM, pp, CLK, 1, 0, until=until, max_events=1000)
    assert total == case["n_events"] and [list(e) for e in events] == case["trace"]
    assert got.end_time[0] == K.EXPECTED_END[k]


@pytest.mark.parametrize("k", sorted(K.SCENARIOS))
def test_cpu_interpreter_matches_reference_live(port, ref, k):
    prog, until = K.SCENARIOS[k]()
    a = port.run(abi.MODEL_PROGRAM, prog.to_params(), CLK, 7, 3, 4, until=until)
    b = ref.run(K.ORACLE_MODEL_KAT, K.KatParams(k, 0), CLK, 7, 3, 4, until=until)
    golden_util.assert_columns_equal(a, b, what=K.SCENARIOS[k].__name__)


@pytest.mark.parametrize("k", sorted(K.SCENARIOS))
def test_reference_engine_runs_the_program_itself(ref, k):
    """oracle/ref/program_runner.hpp executes the PROGRAM with the reference's own awaitables; the result equals
    the scenario written natively on the reference engine (the golden vectors), trace included."""
    prog, until = K.SCENARIOS[k]()
    case = _golden(k)
    got = ref.run(abi.MODEL_PROGRAM, prog.to_params(), CLK, 1, 0, 1, until=until)
    golden_util.assert_columns_equal(got, golden_util.golden_columns(case), what=case["name"])
    total, events = ref.trace(abi.MODEL_PROGRAM, prog.to_params(), CLK, 1, 0, until=until, max_events=1000)
    assert total == case["n_events"] and [list(e) for e in events] == case["trace"]


def test_known_event_counts(port):
    """examples/resource.cpp = 16 events, examples/mutex.cpp = 38 events."""
    for k, n in ((6, 16), (7, 38)):
        prog, until = K.SCENARIOS[k]()
        assert port.run(abi.MODEL_PROGRAM, prog.to_params(), CLK, 1, 0, 1, until=until).n_events[0] == n


def test_mm1_program_equals_hand_lowered_model(port):
    for lam, n in ((5.0, 300), (9.0, 200), (0.5, 100)):
        a = port.run(abi.MODEL_PROGRAM, K.mm1_program(lam, 10.0, n).to_params(), oracles.MM1_CLOCK, 0x5EED, 0, 32)
        b = port.run(abi.MODEL_MM1, abi.MM1Params(lam, 10.0, n), oracles.MM1_CLOCK, 0x5EED, 0, 32)
        assert np.array_equal(a.trace_hash, b.trace_hash) and np.array_equal(a.end_time, b.end_time)
        assert np.array_equal(a.n_events, b.n_events) and np.array_equal(a.i64[0], b.i64[0])
        assert np.array_equal(a.f64[1].view(np.uint64), b.f64[1].view(np.uint64))


SWEEP = [1.0, 3.0, 5.0, 7.0, 9.0]


def test_parameter_sweep_inside_one_ensemble(port, ref):
    """Program.sweep: replication r runs load point (first_replication + r) % steps -- the loop over lambda of
    examples/producer_consumer.cpp:67-73 as one ensemble.  Reference engine == CPU interpreter, every point equals
    the single-rate program of that lambda, and the split into shards does not matter."""
    prog = K.mm1_program(SWEEP, 10.0, 200)
    a = port.run(abi.MODEL_PROGRAM, prog.to_params(), oracles.MM1_CLOCK, 11, 3, 30, threads=4)
    b = ref.run(abi.MODEL_PROGRAM, prog.to_params(), oracles.MM1_CLOCK, 11, 3, 30, threads=4)
    golden_util.assert_columns_equal(a, b, what="sweep: reference engine")
    for r in (0, 1, 7, 29):
        lam = SWEEP[(3 + r) % len(SWEEP)]
        one = port.run(abi.MODEL_PROGRAM, K.mm1_program(lam, 10.0, 200).to_params(), oracles.MM1_CLOCK, 11, 3 + r, 1)
        assert one.trace_hash[0] == a.trace_hash[r] and one.end_time[0] == a.end_time[r]


@pytest.mark.gpu
def test_gpu_parameter_sweep_matches_oracle_and_theory(port):
    prog = K.mm1_program(SWEEP, 10.0, 3000)
    n_reps = 2000
    for no_fast in (False, True):
        if no_fast:
            os.environ["CXXDES_B200_PRG_NO_FAST"] = "1"
        try:
            ens = L.Ensemble(prog, n_reps, seed=11, first_replication=3)
            ens.time_unit(1, L.SECONDS).time_precision(1, L.MILLISECONDS)
            got = ens.run().fetch()
            ens.close()
        finally:
            os.environ.pop("CXXDES_B200_PRG_NO_FAST", None)
        want = port.run(abi.MODEL_PROGRAM, prog.to_params(), oracles.MM1_CLOCK, 11, 3, n_reps, threads=os.cpu_count())
        golden_util.assert_columns_equal(got, want, what=f"sweep (full-size kernel only: {no_fast})")
    for step, lam in enumerate(SWEEP[:4]):              # rho <= 0.7: steady state within 3000 packets
        sel = (3 + np.arange(n_reps)) % len(SWEEP) == step
        latency = got.f64[1][sel].mean() / 3000
        assert abs(latency - 1.0 / (10.0 - lam)) < 0.08 / (10.0 - lam)


def test_if_else_between_awaits_reference_engine_equals_cpu_interpreter(port, ref):
    """M/M/1/K: `if (q.size() < k) co_await q.put(..); else ++dropped;` -- the branch reads the real queue on the
    reference engine (oracle/ref/program_runner.hpp) and the interpreters' own state elsewhere."""
    prog = K.mm1k_program(8.0, 10.0, 400, 4)
    a = port.run(abi.MODEL_PROGRAM, prog.to_params(), oracles.MM1_CLOCK, 21, 0, 40, threads=4)
    b = ref.run(abi.MODEL_PROGRAM, prog.to_params(), oracles.MM1_CLOCK, 21, 0, 40, threads=4)
    golden_util.assert_columns_equal(a, b, what="M/M/1/K")
    assert (a.i64[1] > 0).all() and ((a.i64[0] + a.i64[1]) == 400).all()      # some dropped; served + dropped = offered


@pytest.mark.gpu
def test_gpu_finite_buffer_queue_matches_oracle_and_theory(port):
    lam, mu, k, n = 8.0, 10.0, 4, 2500
    prog = K.mm1k_program(lam, mu, n, k)
    n_reps = 1500
    ens = L.Ensemble(prog, n_reps, seed=21)
    ens.time_unit(1, L.SECONDS).time_precision(1, L.MILLISECONDS)
    got = ens.run().fetch()
    ens.close()
    want = port.run(abi.MODEL_PROGRAM, prog.to_params(), oracles.MM1_CLOCK, 21, 0, n_reps, threads=os.cpu_count())
    golden_util.assert_columns_equal(got, want, what="M/M/1/K")
    rho, cap = lam / mu, k + 1                                  # k waiting + one in service
    p_loss = (1 - rho) * rho ** cap / (1 - rho ** (cap + 1))    # M/M/1/N blocking probability (PASTA)
    assert abs(got.i64[1].sum() / (n_reps * n) - p_loss) < 0.004


def test_consumer_loop_as_the_example_writes_it(port, ref):
    """`while (true) { ...; if (n == n_packets) co_return; ... }` (producer_consumer.cpp:42-53): same event trace as
    the hand-lowered M/M/1 model, on the CPU interpreter and on the reference engine."""
    prog = K.mm1_program_as_written(5.0, 10.0, 250)
    a = port.run(abi.MODEL_PROGRAM, prog.to_params(), oracles.MM1_CLOCK, 0x5EED, 0, 24, threads=4)
    b = port.run(abi.MODEL_MM1, abi.MM1Params(5.0, 10.0, 250), oracles.MM1_CLOCK, 0x5EED, 0, 24, threads=4)
    c = ref.run(abi.MODEL_PROGRAM, prog.to_params(), oracles.MM1_CLOCK, 0x5EED, 0, 24, threads=4)
    for other in (b, c):
        assert np.array_equal(a.trace_hash, other.trace_hash) and np.array_equal(a.end_time, other.end_time)
        assert np.array_equal(a.n_events, other.n_events) and np.array_equal(a.i64[0], other.i64[0])
        assert np.array_equal(a.f64[1].view(np.uint64), other.f64[1].view(np.uint64))


@pytest.mark.gpu
def test_gpu_consumer_loop_as_the_example_writes_it(port):
    prog = K.mm1_program_as_written(5.0, 10.0, 600)
    got = _run_gpu(prog, 512, clock=((1, L.SECONDS), (1, L.MILLISECONDS)), seed=0x5EED)
    want = port.run(abi.MODEL_MM1, abi.MM1Params(5.0, 10.0, 600), oracles.MM1_CLOCK, 0x5EED, 0, 512, threads=os.cpu_count())
    assert not got.status.any()
    assert np.array_equal(got.trace_hash, want.trace_hash) and np.array_equal(got.end_time, want.end_time)
    assert np.array_equal(got.f64[1].view(np.uint64), want.f64[1].view(np.uint64))


def test_builder_rejects_incomplete_programs():
    prog = L.Program()
    prog.process(0, root=True)
    with pytest.raises(ValueError):
        prog.to_params()                      # no body
    prog2 = L.Program()
    p = prog2.process(0)
    with prog2.body(p) as b:
        b.delay(1)
    with pytest.raises(ValueError):
        prog2.to_params()                     # no root


# ------------------------------------------------------------------------------------------ GPU
def _run_gpu(prog, n_reps, clock=((1, 0), (1, 0)), seed=1, first=0, until=-1):
    ens = L.Ensemble(prog, n_reps, seed=seed, first_replication=first)
    ens.time_unit(*clock[0]).time_precision(*clock[1])
    if until < 0:
        ens.run()
    else:
        ens.run_until(until)
    cols = ens.fetch(strict=False)
    ens.close()
    return cols


@pytest.mark.gpu
@pytest.mark.parametrize("k", sorted(K.SCENARIOS))
def test_gpu_interpreter_matches_reference_golden(built, k):
    prog, until = K.SCENARIOS[k]()
    case = _golden(k)
    got = _run_gpu(prog, 64, until=until)
    assert not (got.status & ~np.uint32(abi.REP_UNFINISHED)).any()
    want = golden_util.golden_columns(case)
    for r in (0, 17, 63):                     # the scenarios are deterministic: every replication is the same
        assert got.n_events[r] == want.n_events[0] and got.end_time[r] == want.end_time[0]
        assert got.trace_hash[r] == want.trace_hash[0]
        assert got.i64[0][r] == want.i64[0][0] and got.i64[1][r] == want.i64[1][0]
    assert got.end_time[0] == K.EXPECTED_END[k]


@pytest.mark.gpu
@pytest.mark.parametrize("lam,n", [(5.0, 500), (9.0, 300)])
def test_gpu_mm1_program_matches_oracle_and_fused_kernel(port, lam, n):
    """The same model three ways on the device side: process program on the interpreter kernel,
    hand-written fused kernel, and both against the CPU oracle."""
    n_reps = 512
    prog = K.mm1_program(lam, 10.0, n)
    got = _run_gpu(prog, n_reps, clock=((1, L.SECONDS), (1, L.MILLISECONDS)), seed=0x5EED)
    assert not got.status.any()
    want = port.run(abi.MODEL_MM1, abi.MM1Params(lam, 10.0, n), oracles.MM1_CLOCK, 0x5EED, 0, n_reps,
                    threads=os.cpu_count())
    assert np.array_equal(got.trace_hash, want.trace_hash) and np.array_equal(got.end_time, want.end_time)
    assert np.array_equal(got.n_events, want.n_events) and np.array_equal(got.i64[0], want.i64[0])
    assert np.array_equal(got.f64[1].view(np.uint64), want.f64[1].view(np.uint64))
    ens = L.Ensemble(abi.MM1Params(lam, 10.0, n), n_reps, seed=0x5EED)
    ens.time_unit(1, L.SECONDS).time_precision(1, L.MILLISECONDS)
    fused = ens.run().fetch()
    ens.close()
    assert np.array_equal(fused.trace_hash, got.trace_hash)


def _wide_program():
    """One process, all_of over 24 timeouts inside a loop: needs 24 heap entries, more than the
    first-pass heap of 2 * n_procs + 8 = 10."""
    prog = L.Program()
    main = prog.process(0, root=True)
    with prog.body(main) as b:
        with b.loop(5):
            b.all_of(*[L.Delay(3 + (7 * j) % 11) for j in range(24)])
            b.mark(1, 0)
    return prog


@pytest.mark.gpu
def test_gpu_small_heap_overflow_takes_the_retry_pass(port):
    prog = _wide_program()
    got = _run_gpu(prog, 300)
    want = port.run(abi.MODEL_PROGRAM, prog.to_params(), abi.Clock.of((1, 0), (1, 0)), 1, 0, 300, threads=4)
    assert not got.status.any()
    golden_util.assert_columns_equal(got, want, what="wide all_of")


@pytest.mark.gpu
def test_gpu_rejects_bad_programs(built):
    prog = L.Program()
    p = prog.process(0, root=True)
    with prog.body(p) as b:
        b.await_(5)                            # process 5 does not exist
    with pytest.raises(L.CxxdesError):
        L.Ensemble(prog, 4).run()


# ---------------------------------------------------------------- run_until / run_for in pieces
def _rich_program():
    """Every kind of state a replication can be saved with: semaphore waiters, a bounded queue with a blocked
    producer, event waiters with latencies, armed and stale any_of handlers, loops, registers, draws."""
    prog = L.Program()
    res = prog.resource(2)
    q = prog.queue(2)
    evt = prog.event()
    main = prog.process(0, root=True)
    r_fast, r_slow = prog.rate(0.2), prog.rate(0.05)
    members = []
    for k in range(4):                                   # workers contending for the resource
        p = prog.process(1 + k, priority=k - 1)
        with prog.body(p) as b:
            with b.loop(40):
                b.delay_exp(r_fast); b.down(res); b.set_reg_now(); b.delay_exp(r_slow); b.stat_ticks(0); b.up(res); b.count(0)
        members.append(p)
    ticker = prog.process(10)
    with prog.body(ticker) as b:
        with b.loop(60):
            b.delay(7); b.wake(evt)
    members.append(ticker)
    for k in range(2):                                   # event waiters, one with a wake-up latency
        p = prog.process(11 + k)
        with prog.body(p) as b:
            with b.loop(25):
                b.wait(evt, latency=3 * k, priority=-5 if k else None); b.mark(10, 1 + k)
        members.append(p)
    producer, consumer = prog.process(20), prog.process(21)
    with prog.body(producer) as b:
        with b.loop(50):
            b.put(q); b.delay_exp(r_fast)
    with prog.body(consumer) as b:
        with b.loop(50):
            b.pop(q); b.delay_exp(r_slow); b.stat_seconds(1)
    members += [producer, consumer]
    watch = prog.process(30)
    with prog.body(watch) as b:
        with b.loop(20):
            b.any_of(L.Delay(13), L.Delay(29)); b.mark(10, 3)
    members.append(watch)
    with prog.body(main) as b:
        b.all_of(*members)
    return prog


def _stale_timer_program():
    """any_of(delay(1), delay(1000)) in a loop leaves 30 stale tokens scheduled: more than the shared-memory
    interpreter's heap holds, so every replication is handed to the full-size kernel -- and stays there."""
    prog = L.Program()
    main = prog.process(0, root=True)
    with prog.body(main) as b:
        with b.loop(30):
            b.any_of(L.Delay(1), L.Delay(1000)); b.mark(1, 0)
    return prog


PIECE_CASES = {
    "mm1": (lambda: K.mm1_program(5.0, 10.0, 600), ((1, L.SECONDS), (1, L.MILLISECONDS)), 9000, 8),
    "mm1_heavy": (lambda: K.mm1_program(9.0, 10.0, 500), ((1, L.SECONDS), (1, L.MILLISECONDS)), 5000, 12),
    "rich": (_rich_program, ((1, 0), (1, 0)), 61, 9),
    "rich_fine": (_rich_program, ((1, L.SECONDS), (1, L.MILLISECONDS)), 97000, 7),
    "stale_timers": (_stale_timer_program, ((1, 0), (1, 0)), 7, 6),
    "clocks": (lambda: K.SCENARIOS[5]()[0], ((1, 0), (1, 0)), 3, 7),
}


@pytest.mark.parametrize("name", sorted(PIECE_CASES))
def test_cpu_interpreter_pieces_are_prefixes(port, name):
    """Oracle-side sanity for the cases below: run_until(t) columns are consistent across horizons
    (events and clock never decrease) and the rich program exercises what it claims to."""
    make, clock, piece, n_pieces = PIECE_CASES[name]
    prog = make()
    prev = None
    for k in (1, n_pieces):
        cols = port.run(abi.MODEL_PROGRAM, prog.to_params(), abi.Clock.of(*clock), 31, 0, 16, until=k * piece)
        assert (cols.end_time == k * piece).all() or name != "clocks"
        if prev is not None:
            assert (cols.n_events >= prev.n_events).all()
        prev = cols
    if name != "clocks":                                   # the last horizon still cuts replications short
        full = port.run(abi.MODEL_PROGRAM, prog.to_params(), abi.Clock.of(*clock), 31, 0, 16)
        assert (full.n_events > prev.n_events).any() and not full.status.any()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(PIECE_CASES))
def test_gpu_program_run_for_in_pieces_matches_oracle(port, name):
    """run_for(piece) x n on the interpreter kernels: the shared-memory interpreter saves every replication that
    stops at the horizon and the next launch continues from it (environment.ipp:190-214).  After every piece
    the columns equal the oracle's run_until(k * piece); run() then finishes the replications."""
    make, clock, piece, n_pieces = PIECE_CASES[name]
    prog = make()
    n_reps = 700
    ens = L.Ensemble(prog, n_reps, seed=31)
    ens.time_unit(*clock[0]).time_precision(*clock[1])
    for k in range(1, n_pieces + 1):
        ens.run_for(piece)
        got = ens.fetch(strict=False)
        want = port.run(abi.MODEL_PROGRAM, prog.to_params(), abi.Clock.of(*clock), 31, 0, n_reps, threads=os.cpu_count(),
                        until=k * piece)
        assert not (got.status & ~np.uint32(abi.REP_UNFINISHED)).any()
        golden_util.assert_columns_equal(got, want, what=f"{name} after piece {k}")
    if name != "clocks":                                   # the clocks never stop
        ens.run()
        got = ens.fetch()
        want = port.run(abi.MODEL_PROGRAM, prog.to_params(), abi.Clock.of(*clock), 31, 0, n_reps, threads=os.cpu_count())
        assert not got.status.any()
        # a replication that had finished before the last horizon T was left at now_ = T by run_until (environment.ipp:194)
        want.end_time[:] = np.maximum(want.end_time, n_pieces * piece)
        golden_util.assert_columns_equal(got, want, what=f"{name}: run() after the pieces")
        ens.run_for(1000)                                  # exhausted: nothing left to dispatch, run_until(now() + 1000)
        want.end_time[:] += 1000                           # only moves every replication's clock (environment.ipp:190-209)
        golden_util.assert_columns_equal(ens.fetch(), want, what=f"{name}: run_for after exhaustion")
    ens.close()


@pytest.mark.gpu
def test_gpu_program_pieces_finished_replications_follow_the_horizon_and_reset_starts_over(port):
    prog = K.mm1_program(5.0, 10.0, 40)                    # ends near tick 8000; later horizons lie beyond that
    clock = ((1, L.SECONDS), (1, L.MILLISECONDS))
    n_reps = 300
    ens = L.Ensemble(prog, n_reps, seed=5)
    ens.time_unit(*clock[0]).time_precision(*clock[1])
    for t in (3000, 9000, 20000, 50000):
        ens.run_until(t)
        want = port.run(abi.MODEL_PROGRAM, prog.to_params(), abi.Clock.of(*clock), 5, 0, n_reps, until=t)
        golden_util.assert_columns_equal(ens.fetch(strict=False), want, what=f"run_until({t})")
    assert (ens.fetch(strict=False).end_time == 50000).all()
    ens.reset()                                            # environment::reset: replications start over
    ens.run_until(3000)
    want = port.run(abi.MODEL_PROGRAM, prog.to_params(), abi.Clock.of(*clock), 5, 0, n_reps, until=3000)
    golden_util.assert_columns_equal(ens.fetch(strict=False), want, what="after reset")
    ens.close()


@pytest.mark.gpu
def test_gpu_program_deep_queues_on_request(port, monkeypatch):
    """A heavily loaded queue outgrows the default 128 items per sync::queue (reported per replication, never a
    different trace); cxxdes_b200_config.queue_capacity gives the shared-memory interpreter deeper queues.
    (CXXDES_B200_DEEP_BYTES=0: without the interpreter's last pass, which would redo those replications --
    tests/test_written_after_the_last_gpu_run.py.)"""
    monkeypatch.setenv("CXXDES_B200_DEEP_BYTES", "0")
    prog = K.mm1_program(9.9, 10.0, 6000)
    clock = ((1, L.SECONDS), (1, L.MILLISECONDS))
    n_reps = 256
    want = port.run(abi.MODEL_PROGRAM, prog.to_params(), abi.Clock.of(*clock), 3, 0, n_reps, threads=os.cpu_count())
    ens = L.Ensemble(prog, n_reps, seed=3)
    ens.time_unit(*clock[0]).time_precision(*clock[1])
    got = ens.run().fetch(strict=False)
    ens.close()
    overflowed = (got.status & abi.REP_QUEUE_OVERFLOW) != 0
    assert overflowed.any() and not (got.status & ~np.uint32(abi.REP_QUEUE_OVERFLOW)).any()
    assert np.array_equal(got.trace_hash[~overflowed], want.trace_hash[~overflowed])
    ens = L.Ensemble(prog, n_reps, seed=3, queue_capacity=2000)
    ens.time_unit(*clock[0]).time_precision(*clock[1])
    got = ens.run().fetch()
    assert not got.status.any()
    golden_util.assert_columns_equal(got, want, what="deep queues")
    ens.reset()
    ens.run_for(200000)
    ens.run_for(200000)
    part = port.run(abi.MODEL_PROGRAM, prog.to_params(), abi.Clock.of(*clock), 3, 0, n_reps, threads=os.cpu_count(), until=400000)
    golden_util.assert_columns_equal(ens.fetch(strict=False), part, what="deep queues, two pieces")
    ens.close()


# ------------------------------------------------------------------ validation without a GPU
def _raw(ops, procs=((0, 0),), roots=(0,), **counts):
    """A hand-made program: ops as (code, a, b, imm), procs as (entry, ordinal)."""
    ops_arr = (abi.Op * len(ops))(*[abi.Op(*o) for o in ops])
    procs_arr = (abi.Process * len(procs))(*[abi.Process(e, o, abi.INHERIT64, 0, 0, abi.INHERIT64) for e, o in procs])
    roots_arr = (C.c_uint32 * len(roots))(*roots)
    one = (C.c_uint64 * 1)(0)
    rates = (C.c_double * 1)(1.0)
    pp = abi.ProgramParams(ops_arr, procs_arr, roots_arr, one, one, one, rates, len(ops), len(procs), len(roots),
                           counts.get("n_queues", 0), counts.get("n_sems", 0), counts.get("n_mutexes", 0),
                           counts.get("n_events", 0), counts.get("n_rates", 0), 0, 0)
    pp._keep = (ops_arr, procs_arr, roots_arr, one, rates)
    return pp


def test_check_program_without_a_gpu(cuda_lib):
    """cxxdes_b200_check_program: what cxxdes_b200_create would reject, with the reason, no device needed."""
    import ctypes
    R, DELAY, ALL, CH_DELAY, CH_ALL, IF, JUMP, QPUT = (abi.OP_RETURN, abi.OP_DELAY, abi.OP_ALL_OF, abi.OP_CHILD_DELAY,
                                                      abi.OP_CHILD_ALL_OF, abi.OP_IF, abi.OP_JUMP, abi.OP_Q_PUT)
    LOOP, END = abi.OP_LOOP, abi.OP_END_LOOP
    D = (DELAY, 0, abi.INHERIT32, 1)
    ok = [
        [(DELAY, 0, abi.INHERIT32, 5), (R, 0, 0, 0)],
        [(ALL, 2, abi.INHERIT32, 0), (CH_ALL, 2, abi.INHERIT32, 0), (CH_DELAY, 0, abi.INHERIT32, 1), (CH_DELAY, 0, abi.INHERIT32, 2),
         (CH_DELAY, 0, abi.INHERIT32, 3), (R, 0, 0, 0)],
        [(IF, abi.SRC_NOW << 8 | abi.CMP["<"] << 12, 2, 10), (DELAY, 0, abi.INHERIT32, 1), (R, 0, 0, 0)],
        # for (3) { for (2) { delay } if (now < 10) delay }   -- the IF lands on the outer END_LOOP, inside its own body
        [(LOOP, 0, 7, 3), (LOOP, 0, 4, 2), D, (END, 0, 2, 0), (IF, abi.SRC_NOW << 8 | abi.CMP["<"] << 12, 6, 10), D,
         (END, 0, 1, 0), (R, 0, 0, 0)],
        [(LOOP, 0, 3, 0), D, (END, 0, 1, 0), (R, 0, 0, 0)],          # zero iterations
    ]
    bad = {
        "last op must be RETURN": [(DELAY, 0, abi.INHERIT32, 5)],
        "bad composition": [(ALL, 70, abi.INHERIT32, 0), (CH_DELAY, 0, abi.INHERIT32, 1), (R, 0, 0, 0)],
        "CHILD_": [(ALL, 3, abi.INHERIT32, 0), (CH_DELAY, 0, abi.INHERIT32, 1), (R, 0, 0, 0)],
        "must lie ahead": [(JUMP, 0, 0, 0), (R, 0, 0, 0)],
        "IF operand": [(IF, 7 << 8, 1, 0), (R, 0, 0, 0)],
        "queue operand": [(QPUT, 0, 0, 0), (R, 0, 0, 0)],
        "unknown op": [(99, 0, 0, 0), (R, 0, 0, 0)],
        # loop structure (the interpreters keep two counters per process and trust the pairing)
        "END_LOOP without": [D, (END, 0, 0, 0), (R, 0, 0, 0)],
        "LOOP without": [(LOOP, 0, 2, 3), D, (R, 0, 0, 0)],
        "do not pair up": [(LOOP, 0, 2, 3), D, (END, 0, 0, 0), (R, 0, 0, 0)],
        "two deep": [(LOOP, 0, 7, 2), (LOOP, 0, 6, 2), (LOOP, 0, 5, 2), D, (END, 0, 3, 0), (END, 0, 2, 0), (END, 0, 1, 0), (R, 0, 0, 0)],
        "LOOP count": [(LOOP, 0, 3, 1 << 32), D, (END, 0, 1, 0), (R, 0, 0, 0)],
        "[0, 2^32)": [(LOOP, 0, 3, -1), D, (END, 0, 1, 0), (R, 0, 0, 0)],
        "stay inside its loop": [(LOOP, 0, 4, 3), (JUMP, 0, 4, 0), D, (END, 0, 1, 0), D, (R, 0, 0, 0)],            # out of the body
        "inside its loop body": [(JUMP, 0, 2, 0), (LOOP, 0, 4, 3), D, (END, 0, 2, 0), (R, 0, 0, 0)],                # into the body
        # a time the 47-bit keys cannot hold (and `now + t` must not overflow)
        "below 2^47": [(DELAY, 0, abi.INHERIT32, 1 << 47), (R, 0, 0, 0)],
        "2^47 ticks": [(ALL, 1, abi.INHERIT32, 0), (CH_DELAY, 0, abi.INHERIT32, (1 << 63) - 1), (R, 0, 0, 0)],
    }
    for ops in ok:
        assert cuda_lib.cxxdes_b200_check_program(ctypes.byref(_raw(ops))) == abi.OK, cuda_lib.cxxdes_b200_last_error()
    for reason, ops in bad.items():
        assert cuda_lib.cxxdes_b200_check_program(ctypes.byref(_raw(ops))) == abi.ERR_INVALID
        assert reason.encode() in cuda_lib.cxxdes_b200_last_error(), (reason, cuda_lib.cxxdes_b200_last_error())
    assert cuda_lib.cxxdes_b200_check_program(None) == abi.ERR_INVALID
    # a second body that would start inside the first one's loop
    inside = _raw([(LOOP, 0, 3, 2), D, (END, 0, 1, 0), (R, 0, 0, 0)], procs=((0, 0), (1, 1)))
    assert cuda_lib.cxxdes_b200_check_program(ctypes.byref(inside)) == abi.ERR_INVALID
    assert b"starts inside a loop" in cuda_lib.cxxdes_b200_last_error()
    for k in sorted(K.SCENARIOS):                       # everything the builders emit passes, and Program.check() says so
        K.SCENARIOS[k]()[0].check()
    with pytest.raises(L.CxxdesError):
        prog = L.Program()
        p = prog.process(0, root=True)
        with prog.body(p) as b:
            b.await_(9)
        prog.check()


# ------------------------------------------------------------------ 65 processes: ALOHA's co_main + 64 stations
def _aloha_pair(n_stations=64, pps=20, lam=1.5):
    rate = float(np.float32(lam) / np.float32(n_stations))     # cfg_.lambda / cfg_.num_stations in float (aloha.cpp:52)
    prog = K.aloha_shaped_program(n_stations, pps, 1000, rate)
    return prog, abi.AlohaParams(n_stations, 0, lam, lam, 1.0, 1000.0, pps)


def test_aloha_as_a_program_on_the_cpu(port, ref):
    """examples/aloha.cpp:39-71 (64 stations, all_of.range) as a process program: the CPU interpreter, the reference
    engine executing the program with its own awaitables, and the ALOHA model of the oracle dispatch the same tokens."""
    prog, ap = _aloha_pair()
    prog.check()
    a = port.run(abi.MODEL_PROGRAM, prog.to_params(), oracles.MM1_CLOCK, 7, 3, 6, threads=4)
    b = port.run(abi.MODEL_ALOHA, ap, oracles.MM1_CLOCK, 7, 3, 6, threads=4)
    c = ref.run(abi.MODEL_PROGRAM, prog.to_params(), oracles.MM1_CLOCK, 7, 3, 6, threads=4)
    for other in (b, c):
        assert np.array_equal(a.n_events, other.n_events) and np.array_equal(a.end_time, other.end_time)
        assert np.array_equal(a.trace_hash, other.trace_hash)
    assert (a.n_events == 2 * 64 * 20 + 2 * 64 + 3).all()        # 2 N pps + 2 N + 3


@pytest.mark.gpu
def test_gpu_aloha_as_a_program_matches_the_aloha_kernel(port):
    """The same program on the device interpreter (72 processes, 80 heap entries, 64 children per composition in its own tables) against
    the hand-written ALOHA kernel: event counts, end times and trace digests equal, bit for bit."""
    prog, ap = _aloha_pair(pps=40)
    clock = ((1, L.SECONDS), (1, L.MILLISECONDS))
    n = 600
    ens = L.Ensemble(prog, n, seed=9, first_replication=5)
    ens.time_unit(*clock[0]).time_precision(*clock[1])
    got = ens.run().fetch()
    ens.close()
    ens = L.Ensemble(ap, n, seed=9, first_replication=5)
    ens.time_unit(*clock[0]).time_precision(*clock[1])
    model = ens.run().fetch()
    name = ens.kernel_name()
    ens.close()
    assert name == "aloha_compact_kernel<false>" and not got.status.any() and not model.status.any()
    assert np.array_equal(got.n_events, model.n_events) and np.array_equal(got.end_time, model.end_time)
    assert np.array_equal(got.trace_hash, model.trace_hash)
    want = port.run(abi.MODEL_PROGRAM, prog.to_params(), oracles.MM1_CLOCK, 9, 5, 64, threads=os.cpu_count())
    assert np.array_equal(got.trace_hash[:64], want.trace_hash)
