"""Builder for process programs (cxxdes_b200_program): the generic lowering of coroutine bodies.

Each method appends the lowering of one reference awaitable to the body being defined, so a model
reads like its `CXXDES_SIMULATION` source:

    prog = Program()
    q = prog.queue()
    main, prod, cons = prog.process(0), prog.process(1), prog.process(2)
    with prog.body(prod) as b:                    # coroutine<> producer()
        with b.loop(n_packets):                   #   for (i = 0; i < n_packets; ++i) {
            b.put(q)                              #     co_await q.put(now_seconds());
            b.delay_exp(prog.rate(lam))           #     co_await env.timeout(lambda()); }
    ...
    with prog.body(main) as b:
        b.all_of(prod, cons)                      # co_await (producer() && consumer());
    Ensemble(prog, n_replications).run()

The same program object is interpreted by the CUDA kernel (csrc/program_kernel.cu) and, in the
tests, by the CPU oracle's independent interpreter (oracle/port/program_interp.hpp).
"""
import ctypes as C
import struct

from . import abi


class Delay:
    """delay(t) / timeout(t) as a child of all_of / any_of (timeout.ipp:106-143)."""

    def __init__(self, ticks, priority=None):
        self.ticks, self.priority = int(ticks), priority


class Until:
    """until(t) as a child of all_of / any_of."""

    def __init__(self, ticks, priority=None):
        self.ticks, self.priority = int(ticks), priority


class AllOf:
    """all_of(...) used as a CHILD of another composition -- `(a && b) || c`, and what the reference makes of a chain
    `a && b && c` = all_of(all_of(a, b), c) (compositions.ipp).  Children: Delay / Until / process ids / AllOf / AnyOf."""
    code = abi.OP_CHILD_ALL_OF

    def __init__(self, *children, priority=None, return_latency=0):
        self.children, self.priority, self.return_latency = children, priority, int(return_latency)


class AnyOf(AllOf):
    """any_of(...) used as a child of another composition."""
    code = abi.OP_CHILD_ANY_OF


class Wait:
    """evt.wait(latency, priority) as a child of all_of / any_of: `evt.wait() || timeout(t)`, waiting for a signal with
    a time limit (sync/event.hpp:136-139)."""

    def __init__(self, event, latency=0, priority=None):
        self.event, self.latency, self.priority = int(event), int(latency), priority


def _prio32(priority):
    return abi.INHERIT32 if priority is None else int(priority)


class Body:
    def __init__(self, prog, proc):
        self.prog, self.proc = prog, proc
        self._loops = []

    def __enter__(self):
        self.prog._procs[self.proc]["entry"] = len(self.prog._ops)
        return self

    def __exit__(self, exc_type, *_):
        if exc_type is None:
            if self._loops:
                raise ValueError("unterminated loop")
            self._emit(abi.OP_RETURN)

    def _emit(self, code, a=0, b=0, imm=0):
        self.prog._ops.append([code, a, b, imm])
        return len(self.prog._ops) - 1

    # --- time ---
    def delay(self, ticks, priority=None):       # co_await delay(t) / timeout(t); delay(0) == yield()
        self._emit(abi.OP_DELAY, 0, _prio32(priority), int(ticks))

    def until(self, ticks, priority=None):       # co_await until(t)
        self._emit(abi.OP_UNTIL, 0, _prio32(priority), int(ticks))

    instant = until                              # `instant` and `until` are one functor (timeout.ipp:101)

    def lazy_timeout(self, ticks):               # auto t = co_await lazy_timeout(ticks): reg = now() + ticks, no event
        self._emit(abi.OP_LAZY_TIMEOUT, 0, 0, int(ticks))

    def await_lazy(self, priority=None):         # co_await t  (the instant the lazy_timeout returned)
        self._emit(abi.OP_UNTIL_REG, 0, _prio32(priority), 0)

    def delay_exp(self, rate_index):             # co_await env.timeout(exponential(rate))
        self._emit(abi.OP_DELAY_EXP, rate_index, abi.INHERIT32, 0)

    def delay_real(self, x):                     # co_await env.timeout(double x)
        self._emit(abi.OP_DELAY_REAL, 0, abi.INHERIT32, struct.unpack("<q", struct.pack("<d", float(x)))[0])

    # --- processes ---
    def fresh(self, proc):                       # a new coroutine object from the same coroutine function
        self._emit(abi.OP_FRESH, proc)

    def await_(self, proc):                      # co_await p
        self._emit(abi.OP_AWAIT, proc)

    def async_(self, proc):                      # co_await async(p)
        self._emit(abi.OP_ASYNC, proc)

    def _composition(self, code, children, priority=None, return_latency=0):
        self._emit(code, len(children), _prio32(priority), int(return_latency))
        for c in children:
            if isinstance(c, AllOf):                 # nested: its children follow in place (pre-order)
                self._composition(c.code, c.children, c.priority, c.return_latency)
            elif isinstance(c, Wait):
                self._emit(abi.OP_CHILD_EV_WAIT, c.event, _prio32(c.priority), c.latency)
            elif isinstance(c, Delay):
                self._emit(abi.OP_CHILD_DELAY, 0, _prio32(c.priority), c.ticks)
            elif isinstance(c, Until):
                self._emit(abi.OP_CHILD_UNTIL, 0, _prio32(c.priority), c.ticks)
            else:
                self._emit(abi.OP_CHILD_PROC, int(c))

    # priority / return_latency: all_of(...).priority(p).return_latency(l) (any_of.ipp:30-38) -- the priority the
    # children inherit, and a delay added to the token that resumes the awaiting process
    def all_of(self, *children, priority=None, return_latency=0):   # co_await all_of(...) / (a && b)
        self._composition(abi.OP_ALL_OF, children, priority, return_latency)

    def any_of(self, *children, priority=None, return_latency=0):   # co_await any_of(...) / (a || b)
        self._composition(abi.OP_ANY_OF, children, priority, return_latency)

    # --- sync primitives ---
    def put(self, queue, from_reg=False):        # co_await q.put(now) (or the process register)
        self._emit(abi.OP_Q_PUT, queue, 1 if from_reg else 0)

    def pop(self, queue):                        # reg = co_await q.pop()
        self._emit(abi.OP_Q_POP, queue)

    def down(self, sem):                         # co_await sem.down()   / resource.acquire()
        self._emit(abi.OP_SEM_DOWN, sem)

    def up(self, sem):                           # co_await sem.up()     / handle.release()
        self._emit(abi.OP_SEM_UP, sem)

    def acquire(self, mutex):                    # co_await mtx.acquire()
        self._emit(abi.OP_MTX_ACQUIRE, mutex)

    def release(self, mutex):                    # co_await handle.release()
        self._emit(abi.OP_MTX_RELEASE, mutex)

    def wait(self, event, latency=0, priority=None):   # co_await evt.wait(latency, priority)
        self._emit(abi.OP_EV_WAIT, event, _prio32(priority), int(latency))

    def wake(self, event):                       # co_await evt.wake()
        self._emit(abi.OP_EV_WAKE, event)

    # --- control / statistics ---
    def loop(self, count):
        body = self

        class _Loop:
            def __enter__(self_inner):
                body._loops.append(body._emit(abi.OP_LOOP, 0, 0, int(count)))

            def __exit__(self_inner, exc_type, *_):
                if exc_type is None:
                    head = body._loops.pop()
                    end = body._emit(abi.OP_END_LOOP, 0, head + 1, 0)
                    body.prog._ops[head][2] = end + 1

        return _Loop()

    # --- structured if / else on the model's own state (plain reads, no events) ---
    def if_(self, source, index, cmp, value):
        """`if (<source> <cmp> value) { ... }`: source is "counter" (i64[index]), "queue_size", "sem_value", "now",
        "register" or "mutex" (is_acquired); use as `with b.if_("queue_size", q, "<", 5): ...`, optionally followed
        at once by `with b.else_(): ...`."""
        src = {"counter": abi.SRC_COUNTER, "queue_size": abi.SRC_QUEUE_SIZE, "sem_value": abi.SRC_SEM_VALUE,
               "now": abi.SRC_NOW, "register": abi.SRC_REGISTER, "mutex": abi.SRC_MUTEX}[source]
        body = self

        class _If:
            def __enter__(self_inner):
                self_inner.at = body._emit(abi.OP_IF, int(index) | (src << 8) | (abi.CMP[cmp] << 12), 0, int(value))

            def __exit__(self_inner, exc_type, *_):
                if exc_type is None:
                    body.prog._ops[self_inner.at][2] = len(body.prog._ops)
                    body._last_if = self_inner.at
        return _If()

    def else_(self):
        body = self
        if_at = getattr(self, "_last_if", None)
        if if_at is None or self.prog._ops[if_at][2] != len(self.prog._ops):
            raise ValueError("else_() must follow its if_() block at once")

        class _Else:
            def __enter__(self_inner):
                self_inner.at = body._emit(abi.OP_JUMP)           # the then-block ends by jumping over the else-block
                body.prog._ops[if_at][2] = len(body.prog._ops)
                body._last_if = None

            def __exit__(self_inner, exc_type, *_):
                if exc_type is None:
                    body.prog._ops[self_inner.at][2] = len(body.prog._ops)
        return _Else()

    def return_(self):                           # co_return in the middle of a body (e.g. inside an if_ inside a loop)
        self._emit(abi.OP_RETURN)

    def set_reg_now(self):                       # reg = now()
        self._emit(abi.OP_SET_REG_NOW)

    def stat_seconds(self, index):               # f64[index] += now_seconds() - seconds(reg)
        self._emit(abi.OP_STAT_SECONDS, index)

    def stat_ticks(self, index):                 # f64[index] += now() - reg
        self._emit(abi.OP_STAT_TICKS, index)

    def count(self, index, amount=1):            # i64[index] += amount
        self._emit(abi.OP_COUNT, index, 0, int(amount))

    def mark(self, scale, tag):                  # observation: fold now() * scale + tag into i64[0], i64[1]
        self._emit(abi.OP_MARK, 0, int(tag), int(scale))


class Program:
    def __init__(self):
        self._ops = []
        self._procs = []
        self._roots = []
        self._queue_max = []
        self._sems = []
        self._n_mutexes = 0
        self._n_events = 0
        self._rates = []
        self._sweep_steps = 0
        self._keep = None

    def process(self, ordinal, priority=None, latency=0, return_latency=0, return_priority=None, root=False):
        """A coroutine<> object: `f().priority(p).latency(l).return_latency(r)` (coroutine.ipp:106-176)."""
        self._procs.append({"entry": None, "ordinal": int(ordinal),
                            "priority": abi.INHERIT64 if priority is None else int(priority),
                            "latency": int(latency), "return_latency": int(return_latency),
                            "return_priority": abi.INHERIT64 if return_priority is None else int(return_priority)})
        pid = len(self._procs) - 1
        if root:
            self._roots.append(pid)
        return pid

    def sequential(self, parent_ordinal, *steps, priority=None):
        """sequential(a, b, ...) / the comma operator: an unnamed coroutine that awaits its arguments in order
        (sequential.ipp:1-20).  Returns the process; await it, or use it as a child of all_of / any_of.
        `steps` are Delay / Until objects or process ids; the helper is named `parent_ordinal | 0x8000`, the way
        the oracle harness reports library-created coroutines."""
        p = self.process(int(parent_ordinal) | 0x8000, priority=priority)
        with self.body(p) as b:
            for st in steps:
                if isinstance(st, Delay):
                    b.delay(st.ticks, st.priority)
                elif isinstance(st, Until):
                    b.until(st.ticks, st.priority)
                else:
                    b.await_(int(st))
        return p

    def root(self, proc):
        """env.bind(proc) -- simulation::bind binds co_main; several roots are allowed (examples/clocks.cpp)."""
        self._roots.append(proc)

    def body(self, proc):
        return Body(self, proc)

    def queue(self, max_size=0):
        self._queue_max.append(int(max_size))
        return len(self._queue_max) - 1

    def semaphore(self, value=0, max_value=2 ** 64 - 1):
        self._sems.append((int(value), int(max_value)))
        return len(self._sems) - 1

    def resource(self, count):                   # sync::resource{count} == semaphore{count, count}
        return self.semaphore(count, count)

    def mutex(self):
        self._n_mutexes += 1
        return self._n_mutexes - 1

    def event(self):
        self._n_events += 1
        return self._n_events - 1

    def sweep(self, steps):
        """A parameter sweep inside one ensemble -- the loop over load points in the examples' main()
        (producer_consumer.cpp:67-73, aloha.cpp:93-103): replication r uses point r % steps of every rate given
        as a sequence of `steps` values."""
        self._sweep_steps = int(steps)

    def rate(self, value):
        """One rate, or one per sweep point (a sequence of sweep() values)."""
        self._rates.append([float(v) for v in value] if hasattr(value, "__len__") else float(value))
        return len(self._rates) - 1

    def check(self):
        """Validate against the interpreters' rules without a GPU (cxxdes_b200_check_program); raises CxxdesError."""
        from . import ensemble
        lib = ensemble.load_library()
        pp = self.to_params()
        rc = lib.cxxdes_b200_check_program(C.byref(pp))
        if rc != abi.OK:
            raise ensemble.CxxdesError(rc, lib.cxxdes_b200_last_error().decode())
        return self

    def to_params(self):
        for i, p in enumerate(self._procs):
            if p["entry"] is None:
                raise ValueError(f"process {i} has no body")
        if not self._roots:
            raise ValueError("no root process (env.bind)")
        ops = (abi.Op * len(self._ops))(*[abi.Op(*o) for o in self._ops])
        procs = (abi.Process * len(self._procs))(*[
            abi.Process(p["entry"], p["ordinal"], p["priority"], p["latency"], p["return_latency"],
                        p["return_priority"]) for p in self._procs])
        roots = (C.c_uint32 * len(self._roots))(*self._roots)
        qmax = (C.c_uint64 * max(1, len(self._queue_max)))(*self._queue_max)
        sinit = (C.c_uint64 * max(1, len(self._sems)))(*[s[0] for s in self._sems])
        smax = (C.c_uint64 * max(1, len(self._sems)))(*[s[1] for s in self._sems])
        rows = max(1, self._sweep_steps)
        table = []
        for step in range(rows):
            for r in self._rates:
                if isinstance(r, list):
                    if len(r) != rows:
                        raise ValueError("a swept rate needs one value per sweep point")
                    table.append(r[step])
                else:
                    table.append(r)
        rates = (C.c_double * max(1, len(table)))(*table)
        self._keep = (ops, procs, roots, qmax, sinit, smax, rates)
        return abi.ProgramParams(ops, procs, roots, qmax, sinit, smax, rates, len(self._ops), len(self._procs),
                                 len(self._roots), len(self._queue_max), len(self._sems), self._n_mutexes,
                                 self._n_events, len(self._rates), self._sweep_steps, 0)
