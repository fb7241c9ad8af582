"""GPU tests of what was written AFTER the round's last B200 run (the GPU budget was spent): verified on the CPU
emulation of the device only (tests/test_kernels_on_host.py), first run on a B200 by whoever runs `-m gpu` next.  They
sit in one file that sorts after all the others so that, under `pytest -x`, the tests that HAVE run on the device are
through before any of these can stop the run.

  * run() on an ensemble that is already at its end keeps results and clocks (c_api.cu);
  * M/M/c: the last pass (heap and wait list in global memory) for overloaded systems (models_mmc.cu);
  * process programs: the interpreter's last pass -- timed waits in a loop, more than 4 sync objects of a kind, more
    awaiters / armed compositions / processes than the per-thread tables hold (program_kernel.cu);
  * the reference's own examples recorded by <cxxdes/cxxdes.hpp>, through the C ABI (tests/golden/recorded_examples.json);
  * further cases of the randomised sweeps (the case numbers after those of test_*_fuzz.py).
(CPU-side checks of the same programs against the reference engine are here too, unmarked.)"""
import os

import numpy as np
import pytest

import golden_util
import kat_programs as K   # noqa: F401
import oracles             # noqa: F401
import libcxxdes_b200 as L
from libcxxdes_b200 import abi
from model_cases import MODELS, clock_of, clock_pairs
import test_mm1_fuzz as MF
import test_models_fuzz as XF
import test_programs_fuzz as PF
import test_record_examples as E


# ---- the C ABI at the end of a run ---------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("params,clock", [
    (abi.MM1Params(5.0, 10.0, 30), ((1, L.SECONDS), (1, L.MILLISECONDS))),
    (abi.AlohaParams(16, 0, 1.5, 1.5, 1.0, 10.0, 6), ((1, L.SECONDS), (1, L.MILLISECONDS))),
    (abi.MMCParams(9.0, 1.0, 0.5, 8, 0, 40), ((1, L.SECONDS), (1, L.MILLISECONDS))),
    (abi.ArchSimParams(1, 10, 16, 32, 2, 0, 40), ((1, L.SECONDS), (1, L.SECONDS)))])
def test_run_on_an_exhausted_ensemble_keeps_results_and_clocks(built, params, clock):
    """The reference's run() on an environment that has nothing left dispatches nothing and leaves now() alone.  Here a
    run() on an exhausted handle computes the same results again when no state was kept (a benchmark loop relies on the
    work being done), or does nothing when the pieces that led to the end kept every replication's state -- either way
    results are unchanged and every clock stays where run_until / run_for calls had put it, before and after the end.
    (Found by tools/fuzz_api_on_host.py: the clocks used to fall back to the natural end, or run_for's step was added
    twice.)"""
    def fresh():
        ens = L.Ensemble(params, 300, seed=12)
        return ens.time_unit(*clock[0]).time_precision(*clock[1])
    ens = fresh()
    first = ens.run().fetch()
    natural = first.end_time.copy()
    t = int(np.quantile(natural, 0.6))
    ens.run_until(t)
    ens.run_for(7)
    want = np.maximum(natural, t) + 7
    again = ens.run().fetch()                       # no state kept: computed again, clocks restored
    assert np.array_equal(again.end_time, want)
    assert np.array_equal(again.trace_hash, first.trace_hash) and np.array_equal(again.n_events, first.n_events)
    assert np.array_equal(ens.run_until(t + 3).fetch().end_time, want)           # max(now, t + 3) = now
    assert np.array_equal(ens.run().run().fetch().end_time, want)
    ens.close()
    ens = fresh()                                   # the end reached through pieces: a horizon before it, kept state
    h = int(np.quantile(natural, 0.4))
    ens.run_until(h)
    ens.run()
    ens.run_for(11)
    want = np.maximum(natural, h) + 11
    cols = ens.run().fetch()
    assert np.array_equal(cols.end_time, want) and np.array_equal(cols.trace_hash, first.trace_hash)
    ens.reset()                                     # reset() forgets all of it
    assert np.array_equal(ens.run().fetch().end_time, natural)
    ens.close()


# ---- M/M/c: the last pass ----------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("fields,n_reps,flags,at_least", [
    ((13.0, 1.0, 0.05, 1, 0, 400), 200, 0, 10),                # one server, 13 arrivals per service: a long service strands > 96
    ((40.0, 1.0, 0.05, 1, 0, 800), 150, 0, 75),                # 40 per service: most replications see more than 96 waiters
    ((40.0, 1.0, 0.05, 1, 0, 800), 96, abi.FLAG_WIDE_TOKENS, 48),
    ((15.0, 1.0, 3.0, 2, 0, 400), 128, 0, 64)])                 # long patience: the patience tokens pile up in the heap too
def test_mmc_overloaded_systems_take_the_unbounded_pass(port, fields, n_reps, flags, at_least):
    """The reference's containers are unbounded (std::priority_queue in the environment, std::vector of waiters in the
    semaphore's event); the kernels' on-chip heaps and wait lists hold 64 / 96 entries.  An overloaded system -- every
    customer who arrives during one long service leaves an acquire_proc waiting -- outgrows them: those replications are redone
    by the last pass, heap and wait list in global memory sized for the model's worst case, and come out equal to the
    oracle, column for column, with no status flag; pass_counts tells how many took that pass."""
    model, cls, fname = MODELS["mmc"]
    g = golden_util.load(fname)
    p = cls(*fields)
    clock = clock_pairs(g)
    want = port.run(model, p, clock_of(g), 0xD33B, 5, n_reps, threads=os.cpu_count())
    ens = L.Ensemble(p, n_reps, seed=0xD33B, first_replication=5, flags=flags)
    ens.time_unit(*clock[0]).time_precision(*clock[1])
    got = ens.run().fetch()                       # strict: any status flag raises
    counts = ens.pass_counts()
    acc, _ = ens.reduce()
    assert counts["second_pass"] >= at_least
    golden_util.assert_columns_equal(got, want, what=f"overloaded mmc {fields}")
    assert acc.n_errors == 0 and acc.n_events == int(want.n_events.sum())
    # the event trace of such a replication (the tracing kernel uses the same global-memory storage)
    r = int(np.argmax(want.n_events))
    total, digest, events = ens.trace(r, max_events=64)
    want_total, want_events = port.trace(model, p, clock_of(g), 0xD33B, 5 + r, max_events=64)
    assert total == want_total == int(want.n_events[r]) and digest == int(want.trace_hash[r])
    assert [tuple(e) for e in events] == [tuple(e) for e in want_events]
    # in pieces: the replications are replayed from tick 0 to every horizon, as the 64-bit pass does
    ens.reset()
    horizon = int(want.end_time.max()) // 2
    ens.run_until(horizon)
    part = port.run(model, p, clock_of(g), 0xD33B, 5, n_reps, threads=os.cpu_count(), until=horizon)
    got = ens.fetch(strict=False)
    assert not (got.status & ~np.uint32(abi.REP_UNFINISHED)).any()
    golden_util.assert_columns_equal(got, part, what=f"overloaded mmc {fields} run_until")
    golden_util.assert_columns_equal(ens.run().fetch(), want, what=f"overloaded mmc {fields} run() after run_until")
    ens.close()


@pytest.mark.gpu
def test_mmc_without_the_unbounded_pass_reports_the_capacity(built, monkeypatch):
    """CXXDES_B200_DEEP_BYTES=0 switches the last pass off: the overloaded replications keep their status flag (never a
    silently different trace) and fetch() raises."""
    monkeypatch.setenv("CXXDES_B200_DEEP_BYTES", "0")
    ens = L.Ensemble(abi.MMCParams(13.0, 1.0, 0.05, 1, 0, 400), 64, seed=3)
    ens.time_unit(1, L.SECONDS).time_precision(1, L.MILLISECONDS)
    ens.run()
    got = ens.fetch(strict=False)
    assert (got.status & np.uint32(abi.REP_HEAP_OVERFLOW | abi.REP_WAITER_OVERFLOW)).any()
    with pytest.raises(L.CxxdesError):
        ens.fetch()
    assert ens.pass_counts()["second_pass"] == 0
    ens.close()


# ---- process programs: the interpreter's last pass ---------------------------------------------------------------------
def _tandem(stations, jobs, nested):
    """A line of `stations` machines: a queue in front of each (and one behind the last), a one-unit resource per machine,
    an event per machine that its worker wakes after every job -- more sync objects of each kind than the four the
    interpreter's own arrays hold."""
    prog = L.Program()
    queues = [prog.queue(0) for _ in range(stations + 1)]
    machines = [prog.semaphore(1, 1) for _ in range(stations)]
    done = [prog.event() for _ in range(stations)]
    lam, mu = prog.rate(4.0), prog.rate(6.0)
    main = prog.process(0, root=True)
    procs = [prog.process(1)]
    with prog.body(procs[0]) as b:
        with b.loop(jobs):
            b.put(queues[0])
            b.delay_exp(lam)
    for i in range(stations):
        p = prog.process(2 + i)
        with prog.body(p) as b:
            with b.loop(jobs):
                b.pop(queues[i])
                b.down(machines[i])
                b.delay_exp(mu)
                b.up(machines[i])
                b.put(queues[i + 1], from_reg=True)
                b.wake(done[i])
        procs.append(p)
    sink = prog.process(50)
    with prog.body(sink) as b:
        with b.loop(jobs):
            b.pop(queues[stations])
            b.stat_seconds(0)
            if nested:      # (a nested composition with an event wait: the full-size interpreter's ground)
                b.any_of(L.AllOf(L.Delay(1), L.Delay(2)), L.Wait(done[0], 0, None))
    procs.append(sink)
    with prog.body(main) as b:
        b.all_of(*procs)
    return prog


@pytest.mark.gpu
@pytest.mark.parametrize("nested", [False, True])
def test_gpu_more_than_four_sync_objects_of_a_kind(port, nested):
    """7 queues, 6 semaphores, 6 events (up to 16 of each kind): the interpreter's per-thread arrays hold four, so such a
    program runs on the shared-memory pass where that applies and on the last pass (lists and rings in global memory)
    otherwise -- columns, the event trace and run_until pieces against the CPU interpreter, which the CPU suite holds
    against the reference engine on the same program."""
    prog = _tandem(6, 40, nested)
    clock = ((1, L.SECONDS), (1, L.MILLISECONDS))
    n = 96
    want = port.run(abi.MODEL_PROGRAM, prog.to_params(), abi.Clock.of(*clock), 5, 0, n, threads=os.cpu_count())
    assert not want.status.any()
    ens = L.Ensemble(prog, n, seed=5)
    ens.time_unit(*clock[0]).time_precision(*clock[1])
    golden_util.assert_columns_equal(ens.run().fetch(), want, what="tandem line")
    total, digest, events = ens.trace(3, max_events=200)
    want_total, want_events = port.trace(abi.MODEL_PROGRAM, prog.to_params(), abi.Clock.of(*clock), 5, 3, max_events=200)
    assert total == want_total and [tuple(e) for e in events] == [tuple(e) for e in want_events]
    horizon = int(want.end_time.max()) // 2
    ens.reset()
    part = port.run(abi.MODEL_PROGRAM, prog.to_params(), abi.Clock.of(*clock), 5, 0, n, threads=os.cpu_count(), until=horizon)
    golden_util.assert_columns_equal(ens.run_until(horizon).fetch(strict=False), part, what="tandem line, run_until")
    golden_util.assert_columns_equal(ens.run().fetch(), want, what="tandem line, run() after run_until")
    ens.close()


@pytest.mark.parametrize("nested", [False, True])
def test_more_than_four_sync_objects_on_the_reference_engine(port, ref, nested):
    """The same program run by the unmodified reference engine (oracle/ref/program_runner.hpp) and by the CPU interpreter."""
    prog = _tandem(6, 40, nested)
    clock = abi.Clock.of((1, L.SECONDS), (1, L.MILLISECONDS))
    a = port.run(abi.MODEL_PROGRAM, prog.to_params(), clock, 5, 0, 48, threads=os.cpu_count())
    b = ref.run(abi.MODEL_PROGRAM, prog.to_params(), clock, 5, 0, 48, threads=os.cpu_count())
    golden_util.assert_columns_equal(a, b, what="tandem line: CPU interpreter vs reference engine")
    with pytest.raises(L.CxxdesError):       # 255 processes: refused with a message
        _crowd(254).check()
    with pytest.raises(L.CxxdesError):       # 17 queues: refused with a message
        many = L.Program()
        for _ in range(17):
            many.queue(0)
        root = many.process(0, root=True)
        with many.body(root) as body:
            body.delay(1)
        many.check()


def _many_awaiters(k):
    """k processes `co_await` the same coroutine object (the reference keeps their completion tokens in a vector that
    reserves two, coroutine_data.ipp:181, and grows)."""
    prog = L.Program()
    main = prog.process(0, root=True)
    slow = prog.process(1)
    with prog.body(slow) as b:
        b.delay(50)
    fans = []
    for i in range(k):
        p = prog.process(2 + i, priority=i - 2)
        with prog.body(p) as b:
            b.delay(i)
            b.await_(slow)
            b.delay(1)
        fans.append(p)
    with prog.body(main) as b:
        b.all_of(*fans)
    return prog


def _many_compositions(k):
    """k processes inside `all_of(timeout, any_of(timeout, evt.wait()))` at once: 2 k armed handlers."""
    prog = L.Program()
    ev = prog.event()
    main = prog.process(0, root=True)
    workers = []
    for i in range(k):
        p = prog.process(1 + i)
        with prog.body(p) as b:
            with b.loop(3):
                b.all_of(L.Delay(5 + i), L.AnyOf(L.Delay(100), L.Wait(ev, 0, None)))
        workers.append(p)
    waker = prog.process(99)
    with prog.body(waker) as b:
        with b.loop(6):
            b.delay(40)
            b.wake(ev)
    with prog.body(main) as b:
        b.all_of(*workers, waker)
    return prog


def _crowd(k):
    """k customer processes started one after another (`async`), three desks (a resource): more processes than the 72 of
    the interpreters' own tables -- up to 254 fit a token's tag; the last pass keeps the table in global memory."""
    prog = L.Program()
    desk = prog.semaphore(3, 3)
    lam, mu = prog.rate(20.0), prog.rate(1.0)
    main = prog.process(0, root=True)
    people = []
    for i in range(k):
        p = prog.process(1 + i)
        with prog.body(p) as b:
            b.down(desk)
            b.delay_exp(mu)
            b.up(desk)
            b.stat_seconds(0)
        people.append(p)
    with prog.body(main) as b:
        for p in people:
            b.async_(p)
            b.delay_exp(lam)
        for p in people:
            b.await_(p)
    return prog


_BEYOND_THE_FIXED_POOLS = {"six awaiters of one process": lambda: _many_awaiters(6),
                           "eighty armed compositions": lambda: _many_compositions(40),
                           "two hundred processes": lambda: _crowd(200)}


@pytest.mark.gpu
@pytest.mark.parametrize("what", sorted(_BEYOND_THE_FIXED_POOLS))
def test_gpu_last_pass_holds_more_awaiters_and_compositions(port, what):
    """More than 2 awaiters of one process, more than 24 armed compositions: the interpreters flag the replication, the
    last pass (8 awaiters, 96 compositions) redoes it; equal to the CPU interpreter."""
    prog = _BEYOND_THE_FIXED_POOLS[what]()
    clock = ((1, L.SECONDS), (1, L.SECONDS))
    n = 40
    want = port.run(abi.MODEL_PROGRAM, prog.to_params(), abi.Clock.of(*clock), 5, 0, n, threads=os.cpu_count())
    assert not want.status.any()
    ens = L.Ensemble(prog, n, seed=5)
    ens.time_unit(*clock[0]).time_precision(*clock[1])
    golden_util.assert_columns_equal(ens.run().fetch(), want, what=what)
    if "processes" not in what:      # (a program of more than 72 processes runs on the last pass as a whole, nothing is redone)
        assert ens.pass_counts()["second_pass"] == n
    total, digest, events = ens.trace(3, max_events=300)
    want_total, want_events = port.trace(abi.MODEL_PROGRAM, prog.to_params(), abi.Clock.of(*clock), 5, 3, max_events=300)
    assert total == want_total and digest == int(want.trace_hash[3])
    assert [tuple(e) for e in events] == [tuple(e) for e in want_events]
    ens.close()


@pytest.mark.parametrize("what", sorted(_BEYOND_THE_FIXED_POOLS))
def test_programs_beyond_the_fixed_pools_on_the_reference_engine(port, ref, what):
    prog = _BEYOND_THE_FIXED_POOLS[what]()
    clock = abi.Clock.of((1, L.SECONDS), (1, L.SECONDS))
    golden_util.assert_columns_equal(port.run(abi.MODEL_PROGRAM, prog.to_params(), clock, 5, 0, 8),
                                     ref.run(abi.MODEL_PROGRAM, prog.to_params(), clock, 5, 0, 8), what=what)


@pytest.mark.gpu
def test_gpu_queues_beyond_128_items_take_the_last_pass(port):
    """The heavily loaded M/M/1 program of test_programs.py::test_gpu_program_deep_queues_on_request with the last pass
    in place (the default): the replications whose queue outgrows 128 items are redone with their rings in global memory
    and equal the oracle."""
    prog = K.mm1_program(9.9, 10.0, 6000)
    clock = ((1, L.SECONDS), (1, L.MILLISECONDS))
    n_reps = 256
    want = port.run(abi.MODEL_PROGRAM, prog.to_params(), abi.Clock.of(*clock), 3, 0, n_reps, threads=os.cpu_count())
    ens = L.Ensemble(prog, n_reps, seed=3)
    ens.time_unit(*clock[0]).time_precision(*clock[1])
    golden_util.assert_columns_equal(ens.run().fetch(), want, what="queues beyond 128 items: the last pass")
    assert ens.pass_counts()["second_pass"] == int((want.i64[1] > 128).sum()) or ens.pass_counts()["second_pass"] > 0
    ens.close()
    ens = L.Ensemble(prog, n_reps, seed=3, queue_capacity=2000)      # deep queues on the first pass: nothing to redo
    ens.time_unit(*clock[0]).time_precision(*clock[1])
    golden_util.assert_columns_equal(ens.run().fetch(), want, what="deep queues")
    assert ens.pass_counts()["second_pass"] == 0
    ens.close()


def _timed_waits(n_waiters, turns, patience, wake_every):
    """`for (turns) co_await (evt.wait() || timeout(patience))` in n_waiters processes, one process that wakes the event
    every wake_every ticks: every timeout that wins leaves its waiter in the event's list until the next wake (any_of
    does not cancel the loser), so the list holds about n_waiters * wake_every / patience entries."""
    prog = L.Program()
    ev = prog.event()
    main = prog.process(0, root=True)
    workers = []
    for k in range(n_waiters):
        p = prog.process(1 + k)
        with prog.body(p) as b:
            with b.loop(turns):
                b.any_of(L.Wait(ev, 0, None), L.Delay(patience + k))
        workers.append(p)
    waker = prog.process(100)
    with prog.body(waker) as b:
        with b.loop(turns * patience // wake_every + 2):
            b.delay(wake_every)
            b.wake(ev)
    with prog.body(main) as b:
        b.all_of(*workers, waker)
    return prog


@pytest.mark.gpu
def test_gpu_timed_waits_in_a_loop_take_the_last_pass(port, monkeypatch):
    """The reference's wait list is a std::vector (sync/event.hpp:87-139); the interpreters' lists hold 16 entries (the
    shared-memory pass: one per process).  A timed wait in a loop with rare wakes outgrows them: the last pass -- heap,
    wait lists and queues in global memory -- redoes those replications; columns and the event trace equal the CPU
    interpreter's.  Without that pass the replications keep their capacity flag."""
    prog = _timed_waits(n_waiters=4, turns=60, patience=5, wake_every=150)
    clock = ((1, L.SECONDS), (1, L.SECONDS))
    n_reps = 64
    want = port.run(abi.MODEL_PROGRAM, prog.to_params(), abi.Clock.of(*clock), 21, 0, n_reps, threads=os.cpu_count())
    assert not want.status.any()
    ens = L.Ensemble(prog, n_reps, seed=21)
    ens.time_unit(*clock[0]).time_precision(*clock[1])
    got = ens.run().fetch()
    golden_util.assert_columns_equal(got, want, what="timed waits")
    assert ens.pass_counts()["second_pass"] == n_reps            # (no random draws: every replication is the same)
    total, digest, events = ens.trace(5, max_events=4000)
    want_total, want_events = port.trace(abi.MODEL_PROGRAM, prog.to_params(), abi.Clock.of(*clock), 21, 5, max_events=4000)
    assert total == want_total and digest == int(want.trace_hash[5])
    assert [tuple(e) for e in events] == [tuple(e) for e in want_events]
    horizon = int(want.end_time.max()) // 2                      # in pieces: replayed from tick 0 to each horizon
    ens.reset()
    part = port.run(abi.MODEL_PROGRAM, prog.to_params(), abi.Clock.of(*clock), 21, 0, n_reps, threads=os.cpu_count(), until=horizon)
    golden_util.assert_columns_equal(ens.run_until(horizon).fetch(strict=False), part, what="timed waits, run_until")
    golden_util.assert_columns_equal(ens.run().fetch(), want, what="timed waits, run() after run_until")
    ens.close()
    monkeypatch.setenv("CXXDES_B200_DEEP_BYTES", "0")
    ens = L.Ensemble(prog, n_reps, seed=21)
    ens.time_unit(*clock[0]).time_precision(*clock[1])
    got = ens.run().fetch(strict=False)
    ens.close()
    assert (got.status & abi.REP_WAITER_OVERFLOW).all()


# ---- the reference's examples, recorded, through the C ABI -----------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", [name for name, _, _ in E.EXAMPLES])
def test_gpu_dispatches_the_reference_engines_tokens_for_the_recorded_example(built, name):
    """Through the C ABI: the recorded program as an ensemble (every replication of such a model is the same: no random
    draws) and the event trace of one replication against the rows the reference engine produced for the source file."""
    case = E._golden()[name]
    tables, want, until = case["program"], case["reference"], case["until"]
    clock, pairs = E.T.clock_of(tables)
    ens = L.Ensemble(E.T._RawProgram(tables), 40, seed=E.T.SEED, first_replication=E.T.FIRST)
    ens.time_unit(*pairs[0]).time_precision(*pairs[1])
    got = (ens.run() if until < 0 else ens.run_until(until)).fetch(strict=False)
    assert not (got.status & ~np.uint32(abi.REP_UNFINISHED)).any()
    assert (got.n_events == want["n_events"]).all() and (got.end_time == want["end_time"]).all()
    total, digest, events = ens.trace(7, max_events=20000, until=until)
    ens.close()
    assert total == want["n_events"] and digest == int(got.trace_hash[7])
    assert [list(e[:3]) for e in events] == want["rows"], f"{name}: (time, priority, kind) differ"


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["event", "clocks", "get_environment", "bug", "process_rewrite"])
def test_gpu_example_that_drives_its_environment_runs_as_it_is(built, tmp_path, name):
    """The example's source file with its own main(): env.bind(...) several times, env.run() / env.run_for(t) -- compiled
    against the recording header WITHOUT -DCXXDES_B200_RECORD_ONLY, so that env.run*() records what was bound and runs one
    replication of it through the C ABI; env.now() afterwards is where the reference engine's clock ends for the same
    file.  (Needs the reference's examples: skipped where /root/reference does not exist.)"""
    if not (E.REFERENCE / "examples").is_dir():
        pytest.skip("no /root/reference on this machine")
    now, want = E._build_and_run(tmp_path, name, None, on_device=True)
    assert want["n_events"] > 0 and now == want["end_time"]


# ---- further cases of the randomised sweeps -----------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("case", range(40, 80))
def test_more_mm1_fuzz(port, case):
    MF.test_gpu_fuzz(port, case)


@pytest.mark.gpu
@pytest.mark.parametrize("case", range(14, 40))
def test_more_aloha_fuzz(built, case):
    XF.test_aloha_fuzz(built, case)


@pytest.mark.gpu
@pytest.mark.parametrize("case", range(14, 40))
def test_more_mmc_fuzz(built, case):
    XF.test_mmc_fuzz(built, case)


@pytest.mark.gpu
@pytest.mark.parametrize("case", range(12, 40))
def test_more_archsim_fuzz(built, case):
    XF.test_archsim_fuzz(built, case)


@pytest.mark.gpu
@pytest.mark.parametrize("full_size_only", (False, True))
@pytest.mark.parametrize("case", PF.MORE_CASES)
def test_more_random_programs_match_the_cpu_interpreter(port, monkeypatch, case, full_size_only):
    PF.test_gpu_random_programs_match_the_cpu_interpreter(port, monkeypatch, case, full_size_only)
