"""Randomised process programs, three ways: the UNMODIFIED REFERENCE ENGINE executing each operation with its own
awaitables (oracle/ref/program_runner.hpp; committed as tests/golden/programs_random_reference.json, and live
where oracle/_ref exists) == the restated CPU interpreter of the oracle == the CUDA interpreters (shared-memory
first pass + full-size retry kernel, and the full-size kernel alone), bit for bit, run() and run_for pieces.  The generator only emits programs every side can finish: each loop body contains a delay, a process is
awaited at most once, and nothing depends on capacities the device bounds differently from the oracle."""
import os
import random

import numpy as np
import pytest

import golden_util
import libcxxdes_b200 as L
from libcxxdes_b200 import abi

UNIT = ((1, L.SECONDS), (1, L.MILLISECONDS))


def random_program(seed):
    rng = random.Random(seed)
    prog = L.Program()
    queues = [prog.queue(rng.choice((0, 0, 1, 2, 3))) for _ in range(rng.randint(0, 2))]
    sems = [prog.semaphore(rng.randint(0, 2), rng.choice((2, 3, 2 ** 64 - 1))) for _ in range(rng.randint(0, 2))]
    mutexes = [prog.mutex() for _ in range(rng.randint(0, 2))]
    events = [prog.event() for _ in range(rng.randint(0, 2))]
    rates = [prog.rate(rng.choice((0.5, 2.0, 11.0))) for _ in range(2)]
    main = prog.process(0, root=True)
    n_workers = rng.randint(2, 7)
    next_ordinal = [100]

    def prio():
        return rng.choice((None, None, None, -3, -1, 0, 2, 5))

    def helper(depth):
        """A process awaited exactly once by its creator: a short body of timers."""
        p = prog.process(next_ordinal[0], priority=prio(), latency=rng.randint(0, 3), return_latency=rng.randint(0, 3),
                         return_priority=prio())
        next_ordinal[0] += 1
        with prog.body(p) as b:
            for _ in range(rng.randint(1, 3)):
                timer(b)
                if rng.random() < 0.3:
                    b.mark(10, rng.randint(0, 9))
        return p

    def timer(b):
        k = rng.random()
        if k < 0.4:
            b.delay(rng.randint(0, 40), priority=prio())
        elif k < 0.75:
            b.delay_exp(rng.choice(rates))
        elif k < 0.9:
            b.delay_real(rng.choice((0.001, 0.0125, 0.3)))
        else:
            b.until(rng.randint(0, 3000), priority=prio())

    budget = [21 - n_workers]      # processes left for helpers (24 in total)

    def statement(b, depth):
        k = rng.random()
        if k < 0.30:
            timer(b)
        elif k < 0.38 and queues:
            b.put(rng.choice(queues), from_reg=rng.random() < 0.3)
        elif k < 0.46 and queues:
            b.pop(rng.choice(queues))
        elif k < 0.52 and sems:
            b.down(rng.choice(sems))
        elif k < 0.58 and sems:
            b.up(rng.choice(sems))
        elif k < 0.66 and mutexes:
            m = rng.choice(mutexes)
            b.acquire(m); timer(b); b.release(m)
        elif k < 0.71 and events:
            b.wait(rng.choice(events), latency=rng.randint(0, 5), priority=prio())
        elif k < 0.77 and events:
            b.wake(rng.choice(events))
        elif k < 0.86:
            children = [L.Delay(rng.randint(0, 60), priority=prio()) for _ in range(rng.randint(1, 3))]
            if rng.random() < 0.3:
                children.append(L.Until(rng.randint(0, 2000)))
            if events and rng.random() < 0.3:          # evt.wait() || timeout(t)
                children.insert(rng.randint(0, len(children)), L.Wait(rng.choice(events), rng.randint(0, 4), prio()))
            if rng.random() < 0.35:      # a nested composition: (a && b) || c, a && b && c
                inner = [L.Delay(rng.randint(0, 60), priority=prio()) for _ in range(rng.randint(1, 3))]
                if rng.random() < 0.3:
                    inner.append(L.AnyOf(L.Delay(rng.randint(0, 30)), L.Until(rng.randint(0, 1500))))
                nested = (L.AllOf if rng.random() < 0.5 else L.AnyOf)(*inner, priority=prio() if rng.random() < 0.3 else None,
                                                                      return_latency=rng.choice((0, 0, 2)))
                children.insert(rng.randint(0, len(children)), nested)
            if budget[0] > 0 and depth == 0 and rng.random() < 0.5:
                budget[0] -= 1
                children.insert(rng.randint(0, len(children)), helper(depth))
            (b.all_of if rng.random() < 0.5 else b.any_of)(*children, priority=prio() if rng.random() < 0.4 else None,
                                                             return_latency=rng.choice((0, 0, 0, 1, 4)))
        elif k < 0.90 and budget[0] > 0 and depth == 0:
            budget[0] -= 1
            h = helper(depth)
            (b.await_ if rng.random() < 0.6 else b.async_)(h)
        elif k < 0.905 and depth < 2:
            # if / else on the model's own state, with statements (awaits included) inside
            kinds = ["counter", "now", "register"] + (["queue_size"] if queues else []) + (["sem_value"] if sems else []) + \
                    (["mutex"] if mutexes else [])
            src = rng.choice(kinds)
            index = {"counter": rng.randint(0, 1), "queue_size": rng.choice(queues) if queues else 0,
                     "sem_value": rng.choice(sems) if sems else 0, "mutex": rng.choice(mutexes) if mutexes else 0}.get(src, 0)
            value = {"now": rng.randint(0, 4000), "register": rng.randint(0, 4000), "counter": rng.randint(0, 6),
                     "mutex": rng.randint(0, 1)}.get(src, rng.randint(0, 3))
            with b.if_(src, index, rng.choice(("==", "!=", "<", "<=", ">", ">=")), value):
                for _ in range(rng.randint(1, 2)):
                    statement(b, 2)
            if rng.random() < 0.5:
                with b.else_():
                    statement(b, 2)
        elif k < 0.915:
            b.lazy_timeout(rng.randint(0, 50)); timer(b); b.await_lazy(priority=prio())
        elif k < 0.93:
            b.set_reg_now()
        elif k < 0.96:
            b.stat_ticks(rng.randint(0, 1)) if rng.random() < 0.5 else b.stat_seconds(rng.randint(0, 1))
        elif k < 0.98:
            b.count(rng.randint(0, 1), rng.randint(1, 3))
        else:
            b.mark(rng.choice((1, 10, 1000)), rng.randint(0, 9))

    workers = []
    for wkr in range(n_workers):
        p = prog.process(1 + wkr, priority=prio())
        with prog.body(p) as b:
            for _ in range(rng.randint(1, 4)):
                if rng.random() < 0.6:
                    with b.loop(rng.randint(1, 12)):
                        b.delay(rng.randint(1, 9))          # every loop body suspends at least once
                        for _ in range(rng.randint(1, 4)):
                            statement(b, 1)                 # helpers only outside loops: awaited once
                else:
                    statement(b, 0)
        workers.append(p)
    with prog.body(main) as b:
        rng.shuffle(workers)
        if rng.random() < 0.8:
            b.all_of(*workers)
        else:
            b.any_of(*workers)
            b.mark(1, 0)
    return prog


CASES = list(range(40))            # with committed vectors from the reference engine (programs_random_reference.json)
MORE_CASES = list(range(40, 100))  # live only: reference engine == CPU interpreter here, CPU interpreter == CUDA on the device


def _compare(got, want, what):
    """Replications the device bounded (heap / wait list / queue capacity) are excluded; the rest is bit-exact."""
    bad = (got.status & ~np.uint32(abi.REP_UNFINISHED)) != 0
    assert ((got.status[bad] & ~np.uint32(abi.REP_UNFINISHED | abi.REP_HEAP_OVERFLOW | abi.REP_WAITER_OVERFLOW |
                                           abi.REP_QUEUE_OVERFLOW)) == 0).all(), f"{what}: unexpected status"
    ok = ~bad
    assert ok.sum() * 2 >= len(ok), f"{what}: too many bounded replications"
    for name in ("n_events", "end_time", "trace_hash"):
        assert np.array_equal(getattr(got, name)[ok], getattr(want, name)[ok]), f"{what}: {name} differs"
    for k in range(abi.N_F64):
        assert np.array_equal(got.f64[k].view(np.uint64)[ok], want.f64[k].view(np.uint64)[ok]), f"{what}: f64[{k}]"
    for k in range(abi.N_I64):
        assert np.array_equal(got.i64[k][ok], want.i64[k][ok]), f"{what}: i64[{k}]"


@pytest.mark.parametrize("case", CASES[:12])
def test_generator_makes_programs_the_oracle_finishes(port, case):
    prog = random_program(case)
    cols = port.run(abi.MODEL_PROGRAM, prog.to_params(), abi.Clock.of(*UNIT), 7, 0, 8)
    assert not cols.status.any() and (cols.n_events > 0).all()
    again = port.run(abi.MODEL_PROGRAM, prog.to_params(), abi.Clock.of(*UNIT), 7, 0, 8)
    golden_util.assert_columns_equal(cols, again, what="determinism")


def _golden_case(case):
    g = golden_util.load("programs_random_reference.json")
    return next(c for c in g["cases"] if c["program_seed"] == case)


@pytest.mark.parametrize("case", CASES)
def test_cpu_interpreter_matches_reference_engine_golden(port, case):
    """The committed vectors were produced by the reference engine running the program itself."""
    prog = random_program(case)
    c = _golden_case(case)
    clk = abi.Clock.of(*UNIT)
    got = port.run(abi.MODEL_PROGRAM, prog.to_params(), clk, c["seed"], 0, c["n_replications"])
    golden_util.assert_columns_equal(got, golden_util.golden_columns(c), what=f"program {case}")
    part = port.run(abi.MODEL_PROGRAM, prog.to_params(), clk, c["seed"], 0, c["n_replications"], until=c["run_until"])
    golden_util.assert_columns_equal(part, golden_util.golden_columns({**c, "columns": c["columns_until"]}),
                                     what=f"program {case} run_until")
    total, events = port.trace(abi.MODEL_PROGRAM, prog.to_params(), clk, c["seed"], 0, max_events=64)
    assert total == c["n_events"] and [list(e) for e in events] == c["trace"]


@pytest.mark.parametrize("case", CASES + MORE_CASES)
def test_cpu_interpreter_matches_reference_engine_live(port, ref, case):
    prog = random_program(case)
    clk = abi.Clock.of(*UNIT)
    a = port.run(abi.MODEL_PROGRAM, prog.to_params(), clk, 77 + case, 5, 24, threads=4)
    b = ref.run(abi.MODEL_PROGRAM, prog.to_params(), clk, 77 + case, 5, 24, threads=4)
    golden_util.assert_columns_equal(a, b, what=f"program {case}")
    for frac in (0.25, 0.7):
        until = int(a.end_time.max() * frac)
        golden_util.assert_columns_equal(
            port.run(abi.MODEL_PROGRAM, prog.to_params(), clk, 77 + case, 5, 24, threads=4, until=until),
            ref.run(abi.MODEL_PROGRAM, prog.to_params(), clk, 77 + case, 5, 24, threads=4, until=until),
            what=f"program {case} run_until({until})")


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_gpu_random_programs_match_the_reference_engine_golden(built, case):
    prog = random_program(case)
    c = _golden_case(case)
    ens = L.Ensemble(prog, c["n_replications"], seed=c["seed"])
    ens.time_unit(*UNIT[0]).time_precision(*UNIT[1])
    _compare(ens.run().fetch(strict=False), golden_util.golden_columns(c), f"program {case} vs reference engine")
    ens.reset()
    ens.run_until(c["run_until"])
    _compare(ens.fetch(strict=False), golden_util.golden_columns({**c, "columns": c["columns_until"]}),
             f"program {case} run_until vs reference engine")
    ens.close()


@pytest.mark.gpu
@pytest.mark.parametrize("full_size_only", (False, True))
@pytest.mark.parametrize("case", CASES)
def test_gpu_random_programs_match_the_cpu_interpreter(port, monkeypatch, case, full_size_only):
    if full_size_only:
        monkeypatch.setenv("CXXDES_B200_PRG_NO_FAST", "1")   # read at create: program_kernel.cu alone
    prog = random_program(case)
    n_reps = 96
    ens = L.Ensemble(prog, n_reps, seed=1000 + case)
    ens.time_unit(*UNIT[0]).time_precision(*UNIT[1])
    got = ens.run().fetch(strict=False)
    want = port.run(abi.MODEL_PROGRAM, prog.to_params(), abi.Clock.of(*UNIT), 1000 + case, 0, n_reps, threads=os.cpu_count())
    _compare(got, want, f"program {case}")
    # the same program in run_for pieces (the stateful instantiation), then to the end
    end = int(want.end_time.max())
    piece = max(1, end // 5)
    ens.reset()
    for k in range(1, 4):
        ens.run_for(piece)
        part = port.run(abi.MODEL_PROGRAM, prog.to_params(), abi.Clock.of(*UNIT), 1000 + case, 0, n_reps,
                        threads=os.cpu_count(), until=k * piece)
        _compare(ens.fetch(strict=False), part, f"program {case} after piece {k}")
    got = ens.run().fetch(strict=False)
    want.end_time[:] = np.maximum(want.end_time, 3 * piece)
    _compare(got, want, f"program {case}: run() after the pieces")
    ens.close()
