// oracle/ref/kat_models.hpp -- TEST INFRASTRUCTURE (oracle), not product code.
//
// Known-answer scenarios written natively on the UNMODIFIED reference engine.  They are the
// reference's own tests and examples (cited per scenario), flattened where the original nests
// compositions (`a && b && c` is all_of(all_of(a, b), c) in the reference; the process-program
// interpreter takes flat child lists), with every process object named for the trace digest.
// The same scenarios exist as process programs in tests/kat_programs.py; the parity tests
// compare: reference engine (this file)  ==  CPU interpreter  ==  CUDA interpreter.
#pragma once
#include "harness.hpp"

#include <cxxdes/sync/event.hpp>
#include <cxxdes/sync/mutex.hpp>
#include <cxxdes/sync/queue.hpp>
#include <cxxdes/sync/resource.hpp>
#include <cxxdes/sync/semaphore.hpp>

namespace ref_oracle {

constexpr int ORACLE_MODEL_KAT = 100;  // oracle-only model id; params = kat_params

struct kat_params {
    int32_t scenario;
    int32_t reserved;
};

// 1 -- tests/controlflow.test.cpp:60-72,126-137: all_of / any_of of timeouts, ready child of all_of
CXXDES_SIMULATION(kat_compositions) {
    tracer &tr;
    std::vector<time_integral> marks;
    explicit kat_compositions(tracer &t) : tr{t} { bind(); }
    coroutine<> body() {
        co_await all_of(timeout(10), timeout(20));
        marks.push_back(now());  // 20
        co_await any_of(timeout(10), timeout(20));
        marks.push_back(now());  // 30
        co_await all_of(until(0), delay(5));
        marks.push_back(now());  // 35
    }
    coroutine<> co_main() { return tr.named(body(), 0); }
};

// 2 -- tests/process.test.cpp:81-105: start latency 6 + body 5 + return latency 8 = 19
CXXDES_SIMULATION(kat_latencies) {
    tracer &tr;
    std::vector<time_integral> marks;
    explicit kat_latencies(tracer &t) : tr{t} { bind(); }
    unique_coroutine<int> f() {
        marks.push_back(now());  // 6
        co_await delay(5);
        marks.push_back(now());  // 11
        co_return 5;
    }
    coroutine<> body() {
        co_await tr.named(f(), 1).latency(6).return_latency(8);
        marks.push_back(now());  // 19
    }
    coroutine<> co_main() { return tr.named(body(), 0); }
};

// 3 -- tests/process.test.cpp:25-48: any_of does not cancel the loser; foo() still runs to t = 10
CXXDES_SIMULATION(kat_out_of_scope) {
    tracer &tr;
    std::vector<time_integral> marks;
    explicit kat_out_of_scope(tracer &t) : tr{t} { bind(); }
    coroutine<int> foo() {
        co_await delay(10);
        marks.push_back(now());  // 10
        co_return 10;
    }
    coroutine<> body() {
        {
            auto p = tr.named(foo(), 1);
            co_await any_of(p, delay(1));
            marks.push_back(now());  // 1
        }
    }
    coroutine<> co_main() { return tr.named(body(), 0); }
};

// 4 -- tests/process.test.cpp:149-187: priority ordering at equal time (counter 100 / 101 / 102)
CXXDES_SIMULATION(kat_priorities) {
    tracer &tr;
    std::vector<time_integral> marks;  // now * 1000 + foo ordinal, in execution order
    int counter = 100;
    explicit kat_priorities(tracer &t) : tr{t} { bind(); }
    coroutine<void> foo(int t, int expected, int id) {
        co_await delay(t);
        if (now() != t || counter != expected) marks.push_back(-1);  // the EXPECT_EQs of the reference test
        marks.push_back(now() * 1000 + id);
        counter = counter + 1;
    }
    coroutine<void> spawn_foos(int expected, uint32_t base) {
        co_await async(tr.named(foo(1, expected, (int)base + 0), base + 0));
        co_await async(tr.named(foo(2, expected, (int)base + 1), base + 1));
        co_await async(tr.named(foo(3, expected, (int)base + 2), base + 2));
    }
    coroutine<void> reset_counter() {
        for (int i = 0; i < 4; ++i) {
            counter = 100;
            co_await delay(1);
        }
    }
    coroutine<void> body() {
        co_await all_of(tr.named(spawn_foos(100, 10), 1).priority(100), tr.named(spawn_foos(101, 20), 2).priority(101),
                        tr.named(spawn_foos(102, 30), 3).priority(102), tr.named(reset_counter(), 4).priority(0));
    }
    coroutine<> co_main() { return tr.named(body(), 0); }
};

// 5 -- examples/clocks.cpp: three root processes, periods 2, 3 and 7; heap tie order at t = 6, 12, 18
struct kat_clocks {
    environment env;
    tracer &tr;
    std::vector<time_integral> marks;  // now * 10 + clock id
    explicit kat_clocks(tracer &t) : tr{t} {
        env.bind(tr.named(clock(2, 1), 1));
        env.bind(tr.named(clock(3, 2), 2));
        env.bind(tr.named(clock(7, 3), 3));
    }
    coroutine<> clock(time_integral period, int id) {
        while (true) {
            marks.push_back(env.now() * 10 + id);
            co_await timeout(period);
        }
    }
    time_integral now() const { return env.now(); }
};

// 6 -- examples/resource.cpp: resource{3}, durations 4 / 10 / 2 / 10, #3 with priority 10000; #3 done at 12
CXXDES_SIMULATION(kat_resource) {
    tracer &tr;
    std::vector<time_integral> marks;
    cxxdes::sync::resource resource{3};
    explicit kat_resource(tracer &t) : tr{t} { bind(); }
    coroutine<> p(int id, time_integral duration) {
        auto handle = co_await resource.acquire();
        marks.push_back(now() * 10 + id);
        co_await timeout(duration);
        co_await handle.release();
        marks.push_back(now() * 10 + id);
    }
    coroutine<> body() {
        co_await all_of(tr.named(p(0, 4), 1), tr.named(p(1, 10), 2), tr.named(p(2, 2), 3),
                        tr.named(p(3, 10), 4).priority(10000));
    }
    coroutine<> co_main() { return tr.named(body(), 0); }
};

// 7 -- examples/mutex.cpp: five processes, _Co_with(m) { delay(5) }, priorities 5 3 2 4 1
CXXDES_SIMULATION(kat_mutex) {
    tracer &tr;
    std::vector<time_integral> marks;
    cxxdes::sync::mutex m;
    explicit kat_mutex(tracer &t) : tr{t} { bind(); }
    coroutine<> p(int idx, time_integral t) {
        _Co_with(m) {
            marks.push_back(now() * 10 + idx);
            co_await delay(t);
            marks.push_back(now() * 10 + idx);
        };
    }
    coroutine<> body() {
        co_await all_of(tr.named(p(1, 5), 1).priority(5), tr.named(p(2, 5), 2).priority(3), tr.named(p(3, 5), 3).priority(2),
                        tr.named(p(4, 5), 4).priority(4), tr.named(p(5, 5), 5).priority(1));
    }
    coroutine<> co_main() { return tr.named(body(), 0); }
};

// 8 -- examples/semaphore.cpp (flattened): p1 takes three units, p2 / p3 give one at 5 and 10
CXXDES_SIMULATION(kat_semaphore) {
    tracer &tr;
    std::vector<time_integral> marks;
    cxxdes::sync::semaphore<> q{1};
    explicit kat_semaphore(tracer &t) : tr{t} { bind(); }
    coroutine<> p1() {
        co_await q.down();
        co_await q.down();
        co_await q.down();
        marks.push_back(now());  // 10
    }
    coroutine<> p2() {
        co_await timeout(5);
        co_await q.up();
    }
    coroutine<> p3() {
        co_await timeout(10);
        co_await q.up();
    }
    coroutine<> body() { co_await all_of(tr.named(p1(), 1), tr.named(p2(), 2), tr.named(p3(), 3)); }
    coroutine<> co_main() { return tr.named(body(), 0); }
};

// 9 -- examples/queue.cpp: bounded queue{1}: the second put blocks until the pop at 12
CXXDES_SIMULATION(kat_queue) {
    tracer &tr;
    std::vector<time_integral> marks;
    cxxdes::sync::queue<int> q{1};
    explicit kat_queue(tracer &t) : tr{t} { bind(); }
    coroutine<> p1() {
        auto r1 = co_await q.pop();
        marks.push_back(now() * 10 + r1);  // 5, r = 1
        co_await timeout(7);
        auto r2 = co_await q.pop();
        marks.push_back(now() * 10 + r2);  // 12, r = 2
    }
    coroutine<> p2() {
        co_await timeout(5);
        co_await q.put(1);
        co_await q.put(2);
        marks.push_back(now() * 10);  // 5
    }
    coroutine<> body() { co_await all_of(tr.named(p1(), 1), tr.named(p2(), 2)); }
    coroutine<> co_main() { return tr.named(body(), 0); }
};

// 10 -- examples/event.cpp: event.wait(latency, priority) / wake
CXXDES_SIMULATION(kat_event) {
    tracer &tr;
    std::vector<time_integral> marks;
    cxxdes::sync::event evt;
    explicit kat_event(tracer &t) : tr{t} { bind(); }
    coroutine<> p1() {
        co_await evt.wait(2, -100);
        marks.push_back(now() * 10 + 1);  // 7
    }
    coroutine<> p2() {
        co_await timeout(5);
        co_await evt.wake();
        marks.push_back(now() * 10 + 2);  // 5
    }
    coroutine<> p3() {
        co_await evt.wait(8);
        marks.push_back(now() * 10 + 3);  // 13
        co_await until(500);
        co_await evt.wake();
        co_await until(600);
        co_await evt.wake();
    }
    coroutine<> p4() {
        co_await timeout(20);
        co_await evt.wait();
        marks.push_back(now() * 10 + 4);  // 500
        co_await evt.wait(2);
        marks.push_back(now() * 10 + 4);  // 602
    }
    coroutine<> body() { co_await all_of(tr.named(p1(), 1), tr.named(p2(), 2), tr.named(p3(), 3), tr.named(p4(), 4)); }
    coroutine<> co_main() { return tr.named(body(), 0); }
};

// 11 -- examples/debug.cpp: at t = 5 the timeout with explicit priority -1 pops before the one with 0
CXXDES_SIMULATION(kat_debug) {
    tracer &tr;
    std::vector<time_integral> marks;
    explicit kat_debug(tracer &t) : tr{t} { bind(); }
    coroutine<> p1() {
        co_await timeout(5);
        marks.push_back(now() * 10 + 1);  // 5
        co_await timeout(5);
        marks.push_back(now() * 10 + 1);  // 10
    }
    coroutine<> p2() {
        co_await timeout(2);
        marks.push_back(now() * 10 + 2);  // 2
        co_await timeout(3, -1);
        marks.push_back(now() * 10 + 2);  // 5, before p1's
    }
    coroutine<> body() { co_await all_of(tr.named(p1(), 1), tr.named(p2(), 2)); }
    coroutine<> co_main() { return tr.named(body(), 0); }
};

// 12 -- examples/raii_trick.cpp: resource{1}, _Co_with(res) { timeout(wait) }, priorities 0 -1 -2 -3:
//       ids 3, 2, 1, 0 finish at 8, 16, 21, 31
CXXDES_SIMULATION(kat_raii_trick) {
    tracer &tr;
    std::vector<time_integral> marks;
    cxxdes::sync::resource res{1};
    explicit kat_raii_trick(tracer &t) : tr{t} { bind(); }
    coroutine<> p(int id, time_integral wait_time) {
        _Co_with(res) {
            co_await timeout(wait_time);
            marks.push_back(now() * 10 + id);
        };
    }
    coroutine<> body() {
        co_await all_of(tr.named(p(0, 10), 1).priority(0), tr.named(p(1, 5), 2).priority(-1),
                        tr.named(p(2, 8), 3).priority(-2), tr.named(p(3, 8), 4).priority(-3));
    }
    coroutine<> co_main() { return tr.named(body(), 0); }
};

// 13 -- examples/get_environment.cpp: two roots bound with env.bind, the second with priority -10000
//       runs its four yield() rounds before the first one's body starts; this_environment() is no event
struct kat_get_environment {
    environment env;
    tracer &tr;
    std::vector<time_integral> marks;
    explicit kat_get_environment(tracer &t) : tr{t} {
        env.bind(tr.named(p1(), 1));
        env.bind(tr.named(p2(), 2).priority(-10000));
    }
    coroutine<> p1() {
        marks.push_back(env.now() * 10 + 1);
        auto e = co_await this_environment();  // does not yield control
        marks.push_back(e->now() * 10 + 1);
        co_await timeout(4);
        marks.push_back(e->now() * 10 + 1);  // 4
        co_await e->timeout(4);
        marks.push_back(e->now() * 10 + 1);  // 8
    }
    coroutine<> p2() {
        for (int k = 0; k < 4; ++k) {
            marks.push_back(env.now() * 10 + 2);
            co_await yield();
        }
    }
    time_integral now() const { return env.now(); }
};

// 14 -- examples/lazy_timeout.cpp: a lazy_timeout fixes its deadline when created; awaiting it after
//       the deadline does not suspend: 10, 30, 30
CXXDES_SIMULATION(kat_lazy_timeout) {
    tracer &tr;
    std::vector<time_integral> marks;
    explicit kat_lazy_timeout(tracer &t) : tr{t} { bind(); }
    coroutine<> body() {
        {
            auto t = co_await lazy_timeout(10);
            co_await t;
            marks.push_back(now());  // 10
        }
        {
            auto t = co_await lazy_timeout(10);
            co_await timeout(20);
            marks.push_back(now());  // 30
            co_await t;
            marks.push_back(now());  // 30
        }
    }
    coroutine<> co_main() { return tr.named(body(), 0); }
};

// 15 -- sequential composition: sequential(a, b) and the comma operator are an unnamed coroutine that awaits its
//       arguments in order (sequential.ipp:1-20, compositions.ipp:25-27): 7, 10, 20
CXXDES_SIMULATION(kat_sequential) {
    tracer &tr;
    std::vector<time_integral> marks;
    explicit kat_sequential(tracer &t) : tr{t} { bind(); }
    coroutine<> body() {
        co_await sequential(delay(3), delay(4));
        marks.push_back(now());  // 7
        co_await (delay(1), delay(2));
        marks.push_back(now());  // 10
        co_await all_of(sequential(delay(5), delay(5)), delay(7));
        marks.push_back(now());  // 20
    }
    coroutine<> co_main() { return tr.named(body(), 0); }
};

// 16 -- examples/any_of.cpp: nested compositions -- (a && b) || (c && d), and a chain a && b && c, which the
//       operators build as all_of(all_of(a, b), c): 100, 120, 130, 133
CXXDES_SIMULATION(kat_any_of_example) {
    tracer &tr;
    std::vector<time_integral> marks;
    explicit kat_any_of_example(tracer &t) : tr{t} { bind(); }
    coroutine<> body() {
        co_await ((timeout(1000) && timeout(5)) || (timeout(100) && timeout(1)));
        marks.push_back(now());  // 100
        co_await all_of(timeout(10), timeout(20));
        marks.push_back(now());  // 120
        co_await any_of(timeout(10), timeout(20));
        marks.push_back(now());  // 130
        co_await (delay(1) && delay(2) && delay(3));
        marks.push_back(now());  // 133
    }
    coroutine<> co_main() { return tr.named(body(), 0); }
};

// 17 -- examples/complicated.cpp p1 / p2: waiting for a signal with a time limit, `evt.wait() || timeout(t)`.
//       First the wake wins (5), then the timeout does (8) and the waiter left behind is woken for nothing (12),
//       then a wait with latency inside all_of: 12, woken at 20 + 2 = 22
CXXDES_SIMULATION(kat_wait_or_timeout) {
    tracer &tr;
    cxxdes::sync::event evt;
    std::vector<time_integral> marks;
    explicit kat_wait_or_timeout(tracer &t) : tr{t} { bind(); }
    coroutine<> p1() {
        co_await timeout(5);
        co_await evt.wake();
        co_await timeout(7);
        co_await evt.wake();      // 12: wakes the waiter the second any_of left behind
        co_await timeout(8);
        co_await evt.wake();      // 20
    }
    coroutine<> p2() {
        co_await (evt.wait() || timeout(8));
        marks.push_back(now());   // 5
        co_await (evt.wait() || timeout(3));
        marks.push_back(now());   // 8
        co_await timeout(4);
        marks.push_back(now());   // 12
        co_await all_of(evt.wait(2), timeout(1));
        marks.push_back(now());   // 22
    }
    coroutine<> body() { co_await all_of(tr.named(p1(), 1), tr.named(p2(), 2)); }
    coroutine<> co_main() { return tr.named(body(), 0); }
};

// 18 -- examples/complicated.cpp, examples 1, 2, 3 and 5: a wait with a time limit, a process awaited in the middle
//       of a body, sequential(...) with a process among its arguments, a coroutine that awaits itself recursively
//       (p4(4) here instead of p4(20)), and the dynamic all_of.range / any_of.range
CXXDES_SIMULATION(kat_complicated) {
    tracer &tr;
    cxxdes::sync::event evt;
    std::vector<time_integral> marks;
    explicit kat_complicated(tracer &t) : tr{t} { bind(); }
    coroutine<> p0() { co_await timeout(5); }
    coroutine<> p1() {
        co_await timeout(5);
        co_await evt.wake();
    }
    coroutine<> p2() {
        marks.push_back(now());
        co_await (evt.wait() || timeout(8));
        marks.push_back(now());
        co_await tr.named(p0(), 10);
        marks.push_back(now());
        co_await sequential(timeout(5), timeout(5));
        marks.push_back(now());
    }
    coroutine<> p3() {
        co_await sequential(timeout(5), timeout(10), tr.named(p0(), 11));
        marks.push_back(now());
    }
    coroutine<> p4(unsigned k) {
        if (k == 0) co_return;
        co_await timeout(5);
        marks.push_back(now());
        co_await tr.named(p4(k - 1), 20 + k - 1);
    }
    coroutine<> p5() {
        auto to_wait = std::vector{timeout(5), timeout(10), timeout(20)};
        co_await all_of.range(to_wait.begin(), to_wait.end());
        marks.push_back(now());
        co_await any_of.range(to_wait.begin(), to_wait.end());
        marks.push_back(now());
    }
    coroutine<> body() {
        co_await all_of(tr.named(p1(), 1), tr.named(p2(), 2));
        co_await tr.named(p3(), 3);
        co_await tr.named(p4(4), 24);
        co_await tr.named(p5(), 5);
    }
    coroutine<> co_main() { return tr.named(body(), 0); }
};

// Every scenario folds its marks into i64[0] (count) and i64[1] (order-sensitive checksum) so the
// application-level observations are compared too, not only the token trace.
inline void fold_marks(const std::vector<time_integral> &marks, rep_result &r) {
    r.i64[0] = (int64_t)marks.size();
    uint64_t h = 0;
    for (auto m : marks) h = h * 1000003ULL + (uint64_t)m;
    r.i64[1] = (int64_t)h;
}

template <typename Sim>
inline rep_result run_kat_sim(tracer &tr, int64_t until) {
    rep_result r;
    {
        Sim sim{tr};
        run_traced(sim.env, tr, until);
        r.end_time = sim.now();
        fold_marks(sim.marks, r);
    }
    return r;
}

inline rep_result run_kat(const void *params, int64_t until, tracer &tr, std::vector<time_integral> *marks_out) {
    const auto &p = *static_cast<const kat_params *>(params);
    switch (p.scenario) {
        case 1: return run_kat_sim<kat_compositions>(tr, until);
        case 2: return run_kat_sim<kat_latencies>(tr, until);
        case 3: return run_kat_sim<kat_out_of_scope>(tr, until);
        case 4: return run_kat_sim<kat_priorities>(tr, until);
        case 5: return run_kat_sim<kat_clocks>(tr, until);
        case 6: return run_kat_sim<kat_resource>(tr, until);
        case 7: return run_kat_sim<kat_mutex>(tr, until);
        case 8: return run_kat_sim<kat_semaphore>(tr, until);
        case 9: return run_kat_sim<kat_queue>(tr, until);
        case 10: return run_kat_sim<kat_event>(tr, until);
        case 11: return run_kat_sim<kat_debug>(tr, until);
        case 12: return run_kat_sim<kat_raii_trick>(tr, until);
        case 13: return run_kat_sim<kat_get_environment>(tr, until);
        case 14: return run_kat_sim<kat_lazy_timeout>(tr, until);
        case 15: return run_kat_sim<kat_sequential>(tr, until);
        case 16: return run_kat_sim<kat_any_of_example>(tr, until);
        case 17: return run_kat_sim<kat_wait_or_timeout>(tr, until);
        case 18: return run_kat_sim<kat_complicated>(tr, until);
        default: throw std::runtime_error("ref_oracle: unknown KAT scenario");
    }
}

}  // namespace ref_oracle
