// Builds a few models with the host C++ builder (include/cxxdes_b200/program.hpp) and prints the resulting
// cxxdes_b200_program tables as JSON; tests/test_host_cpp.py compares them field by field with what the
// Python builder makes of the same models (tests/kat_programs.py) -- whose programs are pinned on the
// reference engine.  No simulation runs here.
#include <cinttypes>
#include <cstdio>

#include "cxxdes_b200/program.hpp"

using namespace cxxdes::b200;

static void dump(const char *name, const program &prog, bool last = false) {
    cxxdes_b200_program p = prog.c_struct();
    std::printf("\"%s\": {\"ops\": [", name);
    for (std::uint32_t i = 0; i < p.n_ops; ++i)
        std::printf("%s[%u, %u, %d, %" PRId64 "]", i ? ", " : "", p.ops[i].code, p.ops[i].a, p.ops[i].b, p.ops[i].imm);
    std::printf("], \"procs\": [");
    for (std::uint32_t i = 0; i < p.n_procs; ++i)
        std::printf("%s[%u, %u, %" PRId64 ", %" PRId64 ", %" PRId64 ", %" PRId64 "]", i ? ", " : "", p.procs[i].entry,
                    p.procs[i].ordinal, p.procs[i].priority, p.procs[i].latency, p.procs[i].return_latency,
                    p.procs[i].return_priority);
    std::printf("], \"roots\": [");
    for (std::uint32_t i = 0; i < p.n_roots; ++i) std::printf("%s%u", i ? ", " : "", p.roots[i]);
    std::printf("], \"queue_max\": [");
    for (std::uint32_t i = 0; i < p.n_queues; ++i) std::printf("%s%" PRIu64, i ? ", " : "", p.queue_max[i]);
    std::printf("], \"sems\": [");
    for (std::uint32_t i = 0; i < p.n_sems; ++i)
        std::printf("%s[%" PRIu64 ", %" PRIu64 "]", i ? ", " : "", p.sem_init[i], p.sem_max[i]);
    std::printf("], \"rates\": [");
    for (std::uint32_t i = 0; i < p.n_rates * (p.sweep_steps ? p.sweep_steps : 1); ++i)
        std::printf("%s%.17g", i ? ", " : "", p.rates[i]);
    std::printf("], \"sweep_steps\": %u, \"n_mutexes\": %u, \"n_events\": %u}%s\n", p.sweep_steps, p.n_mutexes, p.n_events,
                last ? "" : ",");
}

// examples/producer_consumer.cpp:31-58
static program mm1(double lam, double mu, std::int64_t n_packets) {
    program prog;
    queue q = prog.add_queue();
    process co_main = prog.add_process(0), producer = prog.add_process(1), consumer = prog.add_process(2);
    prog.bind(co_main);
    prog.body_of(producer, [&](body &co) {
        co.loop(n_packets, [&] {
            co.await(q.put());
            co.await(env_timeout(exponential{lam}));
        });
    });
    prog.body_of(consumer, [&](body &co) {
        co.loop(n_packets - 1, [&] {
            co.await(q.pop()).count(0);
            co.await(env_timeout(exponential{mu})).stat_seconds(1);
        });
        co.await(q.pop()).count(0);
    });
    prog.body_of(co_main, [&](body &co) { co.await(producer && consumer); });
    return prog;
}

// the load sweep of examples/producer_consumer.cpp:61-75 inside one ensemble: lambda differs per sweep point
static program mm1_sweep(const std::vector<double> &lambdas, double mu, std::int64_t n_packets) {
    program prog;
    prog.sweep((std::uint32_t)lambdas.size());
    queue q = prog.add_queue();
    process co_main = prog.add_process(0), producer = prog.add_process(1), consumer = prog.add_process(2);
    prog.bind(co_main);
    prog.body_of(producer, [&](body &co) {
        co.loop(n_packets, [&] {
            co.await(q.put());
            co.await(env_timeout(exponential{lambdas}));
        });
    });
    prog.body_of(consumer, [&](body &co) {
        co.loop(n_packets - 1, [&] {
            co.await(q.pop()).count(0);
            co.await(env_timeout(exponential{mu})).stat_seconds(1);
        });
        co.await(q.pop()).count(0);
    });
    prog.body_of(co_main, [&](body &co) { co.await(producer && consumer); });
    return prog;
}

// M/M/1/K: if (q.size() < k) co_await q.put(..); else ++dropped;
static program mm1k(double lam, double mu, std::int64_t n_packets, std::int64_t k) {
    program prog;
    queue q = prog.add_queue();
    process co_main = prog.add_process(0), producer = prog.add_process(1), consumer = prog.add_process(2);
    prog.bind(co_main);
    prog.body_of(producer, [&](body &co) {
        co.loop(n_packets, [&] {
            co.if_then_else(q.size() < k, [&] { co.await(q.put()); }, [&] { co.count(1); });
            co.await(env_timeout(exponential{lam}));
        });
    });
    prog.body_of(consumer, [&](body &co) {
        co.loop(1 << 30, [&] {
            co.await(q.pop());
            co.await(env_timeout(exponential{mu})).count(0).stat_seconds(1);
        });
    });
    prog.body_of(co_main, [&](body &co) {
        co.await(async(consumer));
        co.await(producer);
    });
    return prog;
}

// tests/controlflow.test.cpp:60-72,126-137
static program compositions() {
    program prog;
    process co_main = prog.add_process(0);
    prog.bind(co_main);
    prog.body_of(co_main, [&](body &co) {
        co.await(all_of(delay(10), delay(20))).mark(1, 0);
        co.await(timeout(10) || timeout(20)).mark(1, 0);
        co.await(until(0) && delay(5)).mark(1, 0);
    });
    return prog;
}

// tests/process.test.cpp:81-105
static program latencies() {
    program prog;
    process co_main = prog.add_process(0);
    prog.bind(co_main);
    program::options o;
    o.latency = 6;
    o.return_latency = 8;
    process f = prog.add_process(1, o);
    prog.body_of(f, [&](body &co) { co.mark(1, 0).await(delay(5)).mark(1, 0); });
    prog.body_of(co_main, [&](body &co) { co.await(f).mark(1, 0); });
    return prog;
}

// tests/process.test.cpp:149-187
static program priorities() {
    program prog;
    process co_main = prog.add_process(0);
    prog.bind(co_main);
    std::vector<process> spawners = {prog.add_process(1, 100), prog.add_process(2, 101), prog.add_process(3, 102)};
    process reset = prog.add_process(4, 0);
    std::vector<std::vector<process>> foos;
    const int bases[3] = {10, 20, 30};
    for (int s = 0; s < 3; ++s) {
        foos.emplace_back();
        for (int k = 0; k < 3; ++k) foos[s].push_back(prog.add_process(bases[s] + k));
        for (int k = 0; k < 3; ++k)
            prog.body_of(foos[s][k], [&](body &co) { co.await(delay(k + 1)).mark(1000, bases[s] + k); });
    }
    for (int s = 0; s < 3; ++s)
        prog.body_of(spawners[s], [&](body &co) {
            for (process f : foos[s]) co.await(async(f));
        });
    prog.body_of(reset, [&](body &co) { co.loop(4, [&] { co.await(delay(1)); }); });
    std::vector<process> all = spawners;
    all.push_back(reset);
    prog.body_of(co_main, [&](body &co) { co.await(all_of_range(all)); });
    return prog;
}

// examples/resource.cpp
static program resource_example() {
    program prog;
    resource res = prog.add_resource(3);
    process co_main = prog.add_process(0);
    prog.bind(co_main);
    std::vector<process> ps = {prog.add_process(1), prog.add_process(2), prog.add_process(3), prog.add_process(4, 10000)};
    const int durations[4] = {4, 10, 2, 10};
    for (int i = 0; i < 4; ++i)
        prog.body_of(ps[i], [&](body &co) {
            co.await(res.acquire()).mark(10, i);
            co.await(delay(durations[i]));
            co.await(res.release()).mark(10, i);
        });
    prog.body_of(co_main, [&](body &co) { co.await(all_of_range(ps)); });
    return prog;
}

// examples/mutex.cpp: _Co_with(m) { ... } is an extra coroutine (co_with.ipp:27-35)
static program mutex_example() {
    program prog;
    mutex m = prog.add_mutex();
    process co_main = prog.add_process(0);
    prog.bind(co_main);
    std::vector<process> ps;
    const int idx[5] = {1, 2, 3, 4, 5}, prio[5] = {5, 3, 2, 4, 1};
    for (int i = 0; i < 5; ++i) {
        process p = prog.add_process(idx[i], prio[i]);
        process helper = prog.add_process(0x8000u | idx[i]);
        prog.body_of(helper, [&](body &co) {
            co.await(m.acquire()).mark(10, idx[i]);
            co.await(delay(5)).mark(10, idx[i]);
            co.await(m.release());
        });
        prog.body_of(p, [&](body &co) { co.await(helper); });
        ps.push_back(p);
    }
    prog.body_of(co_main, [&](body &co) { co.await(all_of_range(ps)); });
    return prog;
}

// examples/queue.cpp (bounded queue: the second put blocks)
static program queue_example() {
    program prog;
    queue q = prog.add_queue(1);
    process co_main = prog.add_process(0);
    prog.bind(co_main);
    process p1 = prog.add_process(1), p2 = prog.add_process(2);
    prog.body_of(p1, [&](body &co) {
        co.await(q.pop()).mark(10, 1);
        co.await(delay(7));
        co.await(q.pop()).mark(10, 2);
    });
    prog.body_of(p2, [&](body &co) {
        co.await(delay(5));
        co.await(q.put());
        co.await(q.put()).mark(10, 0);
    });
    prog.body_of(co_main, [&](body &co) { co.await(p1 && p2); });
    return prog;
}

// examples/event.cpp
static program event_example() {
    program prog;
    event evt = prog.add_event();
    process co_main = prog.add_process(0);
    prog.bind(co_main);
    process p1 = prog.add_process(1), p2 = prog.add_process(2), p3 = prog.add_process(3), p4 = prog.add_process(4);
    prog.body_of(p1, [&](body &co) { co.await(evt.wait(2, -100)).mark(10, 1); });
    prog.body_of(p2, [&](body &co) {
        co.await(delay(5));
        co.await(evt.wake()).mark(10, 2);
    });
    prog.body_of(p3, [&](body &co) {
        co.await(evt.wait(8)).mark(10, 3);
        co.await(until(500));
        co.await(evt.wake());
        co.await(until(600));
        co.await(evt.wake());
    });
    prog.body_of(p4, [&](body &co) {
        co.await(delay(20));
        co.await(evt.wait()).mark(10, 4);
        co.await(evt.wait(2)).mark(10, 4);
    });
    prog.body_of(co_main, [&](body &co) { co.await(all_of(p1, p2, p3, p4)); });
    return prog;
}

// examples/debug.cpp: explicit priority on one timeout
static program debug_example() {
    program prog;
    process co_main = prog.add_process(0);
    prog.bind(co_main);
    process p1 = prog.add_process(1), p2 = prog.add_process(2);
    prog.body_of(p1, [&](body &co) {
        co.await(delay(5)).mark(10, 1);
        co.await(delay(5)).mark(10, 1);
    });
    prog.body_of(p2, [&](body &co) {
        co.await(delay(2)).mark(10, 2);
        co.await(delay(3).priority(-1)).mark(10, 2);
    });
    prog.body_of(co_main, [&](body &co) { co.await(p1 && p2); });
    return prog;
}

// examples/clocks.cpp: three processes bound to the environment, no co_main
static program clocks_example() {
    program prog;
    const int ordinal[3] = {1, 2, 3}, period[3] = {2, 3, 7};
    for (int i = 0; i < 3; ++i) {
        process p = prog.add_process(ordinal[i]);
        prog.bind(p);
        prog.body_of(p, [&](body &co) {
            co.loop(1 << 30, [&] {
                co.mark(10, ordinal[i]);
                co.await(delay(period[i]));
            });
        });
    }
    return prog;
}

// examples/lazy_timeout.cpp
static program lazy_timeout_example() {
    program prog;
    process co_main = prog.add_process(0);
    prog.bind(co_main);
    prog.body_of(co_main, [&](body &co) {
        co.await(lazy_timeout(10));       // auto t = co_await lazy_timeout(10_x);
        co.await(lazy_instant).mark(1, 0);  // co_await t;
        co.await(lazy_timeout(10));
        co.await(timeout(20)).mark(1, 0);
        co.await(lazy_instant).mark(1, 0);
    });
    return prog;
}

// examples/any_of.cpp: nested compositions, chains that nest
static program any_of_example() {
    program prog;
    process co_main = prog.add_process(0);
    prog.bind(co_main);
    prog.body_of(co_main, [&](body &co) {
        co.await((timeout(1000) && timeout(5)) || (timeout(100) && timeout(1))).mark(1, 0);
        co.await(all_of(timeout(10), timeout(20))).mark(1, 0);
        co.await(any_of(timeout(10), timeout(20))).mark(1, 0);
        co.await(delay(1) && delay(2) && delay(3)).mark(1, 0);          // all_of(all_of(1, 2), 3)
    });
    return prog;
}

// examples/complicated.cpp p1 / p2: evt.wait() || timeout(t)
static program wait_or_timeout_example() {
    program prog;
    event evt = prog.add_event();
    process co_main = prog.add_process(0);
    prog.bind(co_main);
    process p1 = prog.add_process(1), p2 = prog.add_process(2);
    prog.body_of(p1, [&](body &co) {
        co.await(delay(5)).await(evt.wake()).await(delay(7)).await(evt.wake()).await(delay(8)).await(evt.wake());
    });
    prog.body_of(p2, [&](body &co) {
        co.await(evt.wait() || timeout(8)).mark(1, 0);
        co.await(evt.wait() || timeout(3)).mark(1, 0);
        co.await(delay(4)).mark(1, 0);
        co.await(all_of(evt.wait(2), timeout(1))).mark(1, 0);
    });
    prog.body_of(co_main, [&](body &co) { co.await(p1 && p2); });
    return prog;
}

// sequential(a, b), the comma operator, a sequential as a child of all_of (sequential.ipp:1-20)
static program sequential_example() {
    program prog;
    process co_main = prog.add_process(0);
    prog.bind(co_main);
    process s1 = prog.sequential(0, {delay(3), delay(4)});
    process s2 = prog.sequential(0, {delay(1), delay(2)});
    process s3 = prog.sequential(0, {delay(5), delay(5)});
    prog.body_of(co_main, [&](body &co) {
        co.await(s1).mark(1, 0);
        co.await(s2).mark(1, 0);
        co.await(all_of(s3, delay(7))).mark(1, 0);
    });
    return prog;
}

int main() {
    std::printf("{\n");
    dump("mm1", mm1(5.0, 10.0, 300));
    dump("mm1k", mm1k(8.0, 10.0, 400, 4));
    dump("mm1_sweep", mm1_sweep({1.0, 3.0, 5.0, 7.0, 9.0}, 10.0, 300));
    dump("1", compositions());
    dump("2", latencies());
    dump("4", priorities());
    dump("5", clocks_example());
    dump("6", resource_example());
    dump("7", mutex_example());
    dump("9", queue_example());
    dump("10", event_example());
    dump("11", debug_example());
    dump("14", lazy_timeout_example());
    dump("15", sequential_example());
    dump("16", any_of_example());
    dump("17", wait_or_timeout_example(), true);
    std::printf("}\n");
    // error behaviour: the reference throws std::runtime_error on misuse (environment.ipp:44-45)
    int caught = 0;
    try {
        program bad;
        bad.add_process(0);
        bad.c_struct();
    } catch (std::runtime_error const &) { ++caught; }
    try {
        program bad;
        process p = bad.add_process(0);
        bad.body_of(p, [](body &co) { co.await(delay(1)); });
        bad.c_struct();   // no root
    } catch (std::runtime_error const &) { ++caught; }
    return caught == 2 ? 0 : 1;
}
