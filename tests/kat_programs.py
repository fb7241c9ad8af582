"""The known-answer scenarios of oracle/ref/kat_models.hpp as process programs (same numbering).

Each function returns (Program, run_until) -- run_until < 0 means run to exhaustion.  Ordinals match
the names given with tracer::named() on the reference side; a `_Co_with` helper coroutine has the
ordinal of its parent | 0x8000 (how the reference harness reports unnamed library coroutines)."""
import ctypes as C

from libcxxdes_b200 import AllOf, AnyOf, Delay, Program, Until, Wait

ORACLE_MODEL_KAT = 100


class KatParams(C.Structure):
    _fields_ = [("scenario", C.c_int32), ("reserved", C.c_int32)]


def compositions():          # 1: tests/controlflow.test.cpp:60-72,126-137
    prog = Program()
    main = prog.process(0, root=True)
    with prog.body(main) as b:
        b.all_of(Delay(10), Delay(20)); b.mark(1, 0)
        b.any_of(Delay(10), Delay(20)); b.mark(1, 0)
        b.all_of(Until(0), Delay(5)); b.mark(1, 0)
    return prog, -1


def latencies():             # 2: tests/process.test.cpp:81-105
    prog = Program()
    main = prog.process(0, root=True)
    f = prog.process(1, latency=6, return_latency=8)
    with prog.body(f) as b:
        b.mark(1, 0); b.delay(5); b.mark(1, 0)
    with prog.body(main) as b:
        b.await_(f); b.mark(1, 0)
    return prog, -1


def out_of_scope():          # 3: tests/process.test.cpp:25-48
    prog = Program()
    main = prog.process(0, root=True)
    foo = prog.process(1)
    with prog.body(foo) as b:
        b.delay(10); b.mark(1, 0)
    with prog.body(main) as b:
        b.any_of(foo, Delay(1)); b.mark(1, 0)
    return prog, -1


def priorities():            # 4: tests/process.test.cpp:149-187
    prog = Program()
    main = prog.process(0, root=True)
    spawners = [prog.process(1, priority=100), prog.process(2, priority=101), prog.process(3, priority=102)]
    reset = prog.process(4, priority=0)
    foos = {}
    for s, base in zip(spawners, (10, 20, 30)):
        foos[s] = [prog.process(base + k) for k in range(3)]
        for k, f in enumerate(foos[s]):
            with prog.body(f) as b:
                b.delay(k + 1); b.mark(1000, base + k)
    for s in spawners:
        with prog.body(s) as b:
            for f in foos[s]:
                b.async_(f)
    with prog.body(reset) as b:
        with b.loop(4):
            b.delay(1)
    with prog.body(main) as b:
        b.all_of(*spawners, reset)
    return prog, -1


def clocks():                # 5: examples/clocks.cpp, env.run_for(20)
    prog = Program()
    for ordinal, period in ((1, 2), (2, 3), (3, 7)):
        p = prog.process(ordinal, root=True)
        with prog.body(p) as b:
            with b.loop(1 << 30):
                b.mark(10, ordinal); b.delay(period)
    return prog, 20


def resource():              # 6: examples/resource.cpp
    prog = Program()
    res = prog.resource(3)
    main = prog.process(0, root=True)
    ps = [prog.process(1), prog.process(2), prog.process(3), prog.process(4, priority=10000)]
    for idx, (p, duration) in enumerate(zip(ps, (4, 10, 2, 10))):
        with prog.body(p) as b:
            b.down(res); b.mark(10, idx); b.delay(duration); b.up(res); b.mark(10, idx)
    with prog.body(main) as b:
        b.all_of(*ps)
    return prog, -1


def mutex():                 # 7: examples/mutex.cpp
    prog = Program()
    m = prog.mutex()
    main = prog.process(0, root=True)
    ps = []
    for idx, prio in ((1, 5), (2, 3), (3, 2), (4, 4), (5, 1)):
        p = prog.process(idx, priority=prio)
        helper = prog.process(0x8000 | idx)        # _Co_with(m) { ... }: co_with.ipp:27-35
        with prog.body(helper) as b:
            b.acquire(m); b.mark(10, idx); b.delay(5); b.mark(10, idx); b.release(m)
        with prog.body(p) as b:
            b.await_(helper)
        ps.append(p)
    with prog.body(main) as b:
        b.all_of(*ps)
    return prog, -1


def semaphore():             # 8: examples/semaphore.cpp (flattened)
    prog = Program()
    q = prog.semaphore(1)
    main = prog.process(0, root=True)
    p1, p2, p3 = prog.process(1), prog.process(2), prog.process(3)
    with prog.body(p1) as b:
        b.down(q); b.down(q); b.down(q); b.mark(1, 0)
    with prog.body(p2) as b:
        b.delay(5); b.up(q)
    with prog.body(p3) as b:
        b.delay(10); b.up(q)
    with prog.body(main) as b:
        b.all_of(p1, p2, p3)
    return prog, -1


def queue():                 # 9: examples/queue.cpp
    prog = Program()
    q = prog.queue(1)
    main = prog.process(0, root=True)
    p1, p2 = prog.process(1), prog.process(2)
    with prog.body(p1) as b:
        b.pop(q); b.mark(10, 1); b.delay(7); b.pop(q); b.mark(10, 2)
    with prog.body(p2) as b:
        b.delay(5); b.put(q); b.put(q); b.mark(10, 0)
    with prog.body(main) as b:
        b.all_of(p1, p2)
    return prog, -1


def event():                 # 10: examples/event.cpp
    prog = Program()
    evt = prog.event()
    main = prog.process(0, root=True)
    p1, p2, p3, p4 = (prog.process(k) for k in (1, 2, 3, 4))
    with prog.body(p1) as b:
        b.wait(evt, 2, -100); b.mark(10, 1)
    with prog.body(p2) as b:
        b.delay(5); b.wake(evt); b.mark(10, 2)
    with prog.body(p3) as b:
        b.wait(evt, 8); b.mark(10, 3); b.until(500); b.wake(evt); b.until(600); b.wake(evt)
    with prog.body(p4) as b:
        b.delay(20); b.wait(evt); b.mark(10, 4); b.wait(evt, 2); b.mark(10, 4)
    with prog.body(main) as b:
        b.all_of(p1, p2, p3, p4)
    return prog, -1


def debug():                 # 11: examples/debug.cpp
    prog = Program()
    main = prog.process(0, root=True)
    p1, p2 = prog.process(1), prog.process(2)
    with prog.body(p1) as b:
        b.delay(5); b.mark(10, 1); b.delay(5); b.mark(10, 1)
    with prog.body(p2) as b:
        b.delay(2); b.mark(10, 2); b.delay(3, priority=-1); b.mark(10, 2)
    with prog.body(main) as b:
        b.all_of(p1, p2)
    return prog, -1


def raii_trick():            # 12: examples/raii_trick.cpp
    prog = Program()
    res = prog.resource(1)
    main = prog.process(0, root=True)
    ps = []
    for ordinal, (ident, wait, prio) in enumerate(((0, 10, 0), (1, 5, -1), (2, 8, -2), (3, 8, -3)), start=1):
        p = prog.process(ordinal, priority=prio)
        helper = prog.process(0x8000 | ordinal)    # _Co_with(res) { ... }: co_with.ipp:27-35
        with prog.body(helper) as b:
            b.down(res); b.delay(wait); b.mark(10, ident); b.up(res)
        with prog.body(p) as b:
            b.await_(helper)
        ps.append(p)
    with prog.body(main) as b:
        b.all_of(*ps)
    return prog, -1


def get_environment():       # 13: examples/get_environment.cpp
    prog = Program()
    p1 = prog.process(1, root=True)
    p2 = prog.process(2, priority=-10000, root=True)
    with prog.body(p1) as b:
        b.mark(10, 1); b.mark(10, 1); b.delay(4); b.mark(10, 1); b.delay(4); b.mark(10, 1)
    with prog.body(p2) as b:
        with b.loop(4):
            b.mark(10, 2); b.delay(0)              # yield() == delay(0), timeout.ipp:180-182
    return prog, -1


def lazy_timeout():          # 14: examples/lazy_timeout.cpp
    prog = Program()
    main = prog.process(0, root=True)
    with prog.body(main) as b:
        b.lazy_timeout(10); b.await_lazy(); b.mark(1, 0)                       # auto t = co_await lazy_timeout(10); co_await t;
        b.lazy_timeout(10); b.delay(20); b.mark(1, 0); b.await_lazy(); b.mark(1, 0)   # ... co_await timeout(20); co_await t;
    return prog, -1


def sequential():             # 15: sequential(a, b), the comma operator, a sequential as a child of all_of
    prog = Program()
    main = prog.process(0, root=True)
    s1 = prog.sequential(0, Delay(3), Delay(4))
    s2 = prog.sequential(0, Delay(1), Delay(2))
    s3 = prog.sequential(0, Delay(5), Delay(5))
    with prog.body(main) as b:
        b.await_(s1); b.mark(1, 0)
        b.await_(s2); b.mark(1, 0)
        b.all_of(s3, Delay(7)); b.mark(1, 0)
    return prog, -1


def any_of_example():         # 16: examples/any_of.cpp -- nested compositions, a chain that nests
    prog = Program()
    main = prog.process(0, root=True)
    with prog.body(main) as b:
        b.any_of(AllOf(Delay(1000), Delay(5)), AllOf(Delay(100), Delay(1))); b.mark(1, 0)
        b.all_of(Delay(10), Delay(20)); b.mark(1, 0)
        b.any_of(Delay(10), Delay(20)); b.mark(1, 0)
        b.all_of(AllOf(Delay(1), Delay(2)), Delay(3)); b.mark(1, 0)        # delay(1) && delay(2) && delay(3)
    return prog, -1


def wait_or_timeout():        # 17: examples/complicated.cpp p1 / p2 -- evt.wait() || timeout(t)
    prog = Program()
    evt = prog.event()
    main = prog.process(0, root=True)
    p1, p2 = prog.process(1), prog.process(2)
    with prog.body(p1) as b:
        b.delay(5); b.wake(evt); b.delay(7); b.wake(evt); b.delay(8); b.wake(evt)
    with prog.body(p2) as b:
        b.any_of(Wait(evt), Delay(8)); b.mark(1, 0)
        b.any_of(Wait(evt), Delay(3)); b.mark(1, 0)
        b.delay(4); b.mark(1, 0)
        b.all_of(Wait(evt, 2), Delay(1)); b.mark(1, 0)
    with prog.body(main) as b:
        b.all_of(p1, p2)
    return prog, -1


def complicated():            # 18: examples/complicated.cpp, examples 1, 2, 3 (p4(4)) and 5
    prog = Program()
    evt = prog.event()
    main = prog.process(0, root=True)
    p1, p2, p3, p5 = prog.process(1), prog.process(2), prog.process(3), prog.process(5)
    p0a, p0b = prog.process(10), prog.process(11)
    for p0 in (p0a, p0b):
        with prog.body(p0) as b:
            b.delay(5)
    with prog.body(p1) as b:
        b.delay(5); b.wake(evt)
    seq2 = prog.sequential(2, Delay(5), Delay(5))
    with prog.body(p2) as b:
        b.mark(1, 0); b.any_of(Wait(evt), Delay(8)); b.mark(1, 0)
        b.await_(p0a); b.mark(1, 0)
        b.await_(seq2); b.mark(1, 0)
    seq3 = prog.sequential(3, Delay(5), Delay(10), p0b)
    with prog.body(p3) as b:
        b.await_(seq3); b.mark(1, 0)
    p4 = [prog.process(20 + k) for k in range(5)]             # p4(k): `if (k == 0) co_return; ...; co_await p4(k - 1);`
    with prog.body(p4[0]) as b:
        pass
    for k in range(1, 5):
        with prog.body(p4[k]) as b:
            b.delay(5); b.mark(1, 0); b.await_(p4[k - 1])
    with prog.body(p5) as b:
        b.all_of(Delay(5), Delay(10), Delay(20)); b.mark(1, 0)
        b.any_of(Delay(5), Delay(10), Delay(20)); b.mark(1, 0)
    with prog.body(main) as b:
        b.all_of(p1, p2); b.await_(p3); b.await_(p4[4]); b.await_(p5)
    return prog, -1


SCENARIOS = {1: compositions, 2: latencies, 3: out_of_scope, 4: priorities, 5: clocks, 6: resource, 7: mutex,
             8: semaphore, 9: queue, 10: event, 11: debug, 12: raii_trick, 13: get_environment, 14: lazy_timeout, 15: sequential, 16: any_of_example, 17: wait_or_timeout, 18: complicated}
# end time of every scenario on the reference engine (the reference's tests / examples document the
# observation times; scenario 1 ends at 40 because the loser of any_of(timeout(10), timeout(20)) still fires)
EXPECTED_END = {1: 40, 2: 19, 3: 10, 4: 4, 5: 20, 6: 12, 7: 25, 8: 10, 9: 12, 10: 602, 11: 10, 12: 31, 13: 8, 14: 30, 15: 20, 16: 1000, 17: 22, 18: 100}


def mm1_program(lam, mu, n_packets):
    """examples/producer_consumer.cpp:31-58 as a process program (same ordinals / streams as the
    hand-written kernels: co_main 0, producer 1, consumer 2); f64[1] = total_latency, i64[0] = packets consumed.
    `lam` may be a sequence: the load sweep of the example's main() inside one ensemble (Program.sweep)."""
    prog = Program()
    if hasattr(lam, "__len__"):
        prog.sweep(len(lam))
    q = prog.queue()
    main = prog.process(0, root=True)
    prod, cons = prog.process(1), prog.process(2)
    r_lam, r_mu = prog.rate(lam), prog.rate(mu)
    with prog.body(prod) as b:
        with b.loop(n_packets):
            b.put(q); b.delay_exp(r_lam)
    with prog.body(cons) as b:
        with b.loop(n_packets - 1):
            b.pop(q); b.count(0); b.delay_exp(r_mu); b.stat_seconds(1)
        b.pop(q); b.count(0)
    with prog.body(main) as b:
        b.all_of(prod, cons)
    return prog


def aloha_shaped_program(n_stations, packets_per_station, frame_ticks, rate):
    """The processes of examples/aloha.cpp:39-71 as a process program: co_main awaits all_of.range over N stations, a
    station loops `co_await timeout(frame); co_await timeout(exponential(rate))`.  The station set and the collision
    flags (aloha.cpp:54-59,73-78) are plain data that schedules nothing, so the TOKENS -- times, order, digest -- are
    those of the ALOHA model kernels with the same seed: ordinals 0 (co_main) and 1..N double as Philox stream ids."""
    prog = Program()
    main = prog.process(0, root=True)
    stations = [prog.process(1 + s) for s in range(n_stations)]
    r = prog.rate(rate)
    for st in stations:
        with prog.body(st) as b:
            with b.loop(packets_per_station):
                b.delay(frame_ticks); b.delay_exp(r)
    with prog.body(main) as b:
        b.all_of(*stations)
    return prog


def mm1k_program(lam, mu, n_packets, k):
    """M/M/1/K loss system: the producer drops a packet when the buffer already holds k (the textbook use of an `if`
    between awaits: `if (q.size() < k) co_await q.put(...); else ++dropped;`).  i64[0] = served, i64[1] = dropped."""
    prog = Program()
    q = prog.queue()
    main = prog.process(0, root=True)
    prod, cons = prog.process(1), prog.process(2)
    r_lam, r_mu = prog.rate(lam), prog.rate(mu)
    with prog.body(prod) as b:
        with b.loop(n_packets):
            with b.if_("queue_size", q, "<", k):
                b.put(q)
            with b.else_():
                b.count(1)
            b.delay_exp(r_lam)
    with prog.body(cons) as b:
        with b.loop(1 << 30):                      # serves until the run ends
            b.pop(q); b.delay_exp(r_mu); b.count(0); b.stat_seconds(1)
    with prog.body(main) as b:
        b.async_(cons)
        b.await_(prod)
    return prog


def mm1_program_as_written(lam, mu, n_packets):
    """examples/producer_consumer.cpp:40-58 with the consumer's control flow as the example writes it:
    `while (true) { x = co_await q.pop(); ++n; if (n == n_packets) co_return; co_await env.timeout(mu()); total += .. }`
    -- same trace as mm1_program(), whose consumer loop is unrolled instead."""
    prog = Program()
    q = prog.queue()
    main = prog.process(0, root=True)
    prod, cons = prog.process(1), prog.process(2)
    r_lam, r_mu = prog.rate(lam), prog.rate(mu)
    with prog.body(prod) as b:
        with b.loop(n_packets):
            b.put(q); b.delay_exp(r_lam)
    with prog.body(cons) as b:
        with b.loop(1 << 30):                               # while (true)
            b.pop(q); b.count(0)
            with b.if_("counter", 0, "==", n_packets):
                b.return_()
            b.delay_exp(r_mu); b.stat_seconds(1)
    with prog.body(main) as b:
        b.all_of(prod, cons)
    return prog
