// tests/native/recorded_models.hpp -- libcxxdes models as a user writes them: real C++20 coroutine bodies on the
// reference's public API (CXXDES_SIMULATION, coroutine<>, timeout / delay / yield, all_of / any_of / && / ||, async,
// sync::queue / resource / mutex, .priority() / .return_latency(), env.time_unit / time_precision / timeout,
// _Co_with, sequential / the comma operator, subroutine<>).
// This ONE source text is compiled twice by tests/test_record.py:
//   * against the reference's <cxxdes/cxxdes.hpp> (record_reference.cpp): the unmodified engine runs it;
//   * against include/cxxdes_b200/record/cxxdes/cxxdes.hpp (record_dump.cpp): the bodies are recorded into process
//     programs, which the interpreters (CPU oracle, CUDA kernels) then run.
// Both must dispatch the same tokens.  Bodies follow the reference's examples (resource.cpp, mutex.cpp,
// producer_consumer.cpp, clocks.cpp) and tests (controlflow.test.cpp), without their prints.
#pragma once
#include <cstddef>

#include <cxxdes/cxxdes.hpp>

#include "cxxdes_b200/random_variable.hpp"

namespace recorded {

using namespace cxxdes::core;

// examples/resource.cpp:6-31
CXXDES_SIMULATION(resource_example) {
    cxxdes::sync::resource resource{3};

    coroutine<> p(int id, time_integral duration) {
        (void)id;
        auto handle = co_await resource.acquire();
        co_await timeout(duration);
        co_await handle.release();
    }

    coroutine<> co_main() {
        co_await all_of(p(0, 4), p(1, 10), p(2, 2), p(3, 10).priority(10000 /* make it last to acquire the resource */));
        co_return;
    }
};

// examples/mutex.cpp:8-31 with the acquire / release written out (no _Co_with)
CXXDES_SIMULATION(mutex_example) {
    cxxdes::sync::mutex m;

    coroutine<> p(time_integral t) {
        auto handle = co_await m.acquire();
        co_await delay(t);
        co_await handle.release();
    }

    coroutine<> co_main() {
        co_await all_of(p(5).priority(5), p(5).priority(3), p(5).priority(2), p(5).priority(4), p(5).priority(1));
    }
};

// examples/producer_consumer.cpp:9-59 (the statistics in plain C++ left out: a recording cannot see them)
CXXDES_SIMULATION(producer_consumer) {
    producer_consumer(double lambda_, double mu_, std::size_t n_packets_)
        : n_packets{n_packets_}, lambda{env, 1, lambda_}, mu{env, 2, mu_} {
        using namespace cxxdes::core::time_ops;
        env.time_unit(1_s);
        env.time_precision(1_ms);
    }

    cxxdes::sync::queue<double> q;
    std::size_t n_packets;
    cxxdes_b200::exponential_rv lambda;
    cxxdes_b200::exponential_rv mu;

    coroutine<> producer() {
        for (std::size_t i = 0; i < n_packets; ++i) {
            co_await q.put(now_seconds());
            co_await env.timeout(lambda());
        }
    }

    coroutine<> consumer() {
        std::size_t n = 0;
        while (true) {
            auto x = co_await q.pop();
            (void)x;
            ++n;
            if (n == n_packets) co_return;
            co_await env.timeout(mu());
        }
    }

    coroutine<> co_main() { co_await (producer() && consumer()); }
};

// examples/clocks.cpp:20-25: bodies that never end (stopped by run_until)
CXXDES_SIMULATION(clocks) {
    coroutine<> clock(time_integral period) {
        while (true) co_await timeout(period);
    }
    coroutine<> co_main() { co_await all_of(clock(2), clock(3), clock(7)); }
};

// return values, latencies, async, any_of with a timeout, && of timeouts with a priority, a counted loop whose bound is
// a value another process returned (the same in every replication)
CXXDES_SIMULATION(mixed) {
    coroutine<int> worker(time_integral t) {
        co_await delay(t);
        co_return (int)t;
    }
    coroutine<> side() {
        co_await delay(3);
        co_await yield();
    }
    coroutine<> co_main() {
        int r = co_await worker(5).return_latency(2);
        co_await async(side());
        co_await any_of(worker(4), timeout(9));
        co_await (timeout(2) && timeout(3, -1));
        for (int k = 0; k < r; ++k) {
            co_await yield();
            co_await delay(1);
        }
        co_await until(40);
    }
};

// examples/mutex.cpp:8-31 and examples/raii_trick.cpp:8-30 AS WRITTEN: `_Co_with(x) { ... };` -- a coroutine of the library
// that acquires, runs the body (a subroutine) and releases (co_with.ipp:27-35)
CXXDES_SIMULATION(co_with_example) {
    cxxdes::sync::mutex m;
    cxxdes::sync::resource res{1};

    coroutine<> p(int idx, time_integral t) {
        (void)idx;
        _Co_with(m) {
            co_await delay(t);
        };
    }
    coroutine<> r(int id, time_integral wait_time) {
        (void)id;
        _Co_with(res) {
            co_await timeout(wait_time);
        };
    }

    coroutine<> co_main() {
        co_await all_of(p(1, 5).priority(5), p(2, 5).priority(3), p(3, 5).priority(2), p(4, 5).priority(4), p(5, 5).priority(1));
        co_await all_of(r(0, 10).priority(0), r(1, 5).priority(-1), r(2, 8).priority(-2), r(3, 8).priority(-3));
    }
};

// sequential(...), the comma operator, a sequential as a child of all_of, subroutine<T> with a return value
// (sequential.ipp, compositions.ipp:24-27, subroutine.ipp; tests/controlflow.test.cpp, examples/subroutine.cpp)
CXXDES_SIMULATION(sequential_example) {
    subroutine<int> twice(time_integral t) {
        co_await delay(t);
        co_await delay(t);
        co_return 2;
    }
    subroutine<> nested(time_integral t) {
        int n = co_await twice(t);
        for (int k = 0; k < n; ++k) co_await yield();
    }
    coroutine<> worker(time_integral t) {
        co_await nested(t);
        co_await delay(1);
    }
    coroutine<> co_main() {
        co_await sequential(delay(3), delay(4));
        co_await (delay(1), delay(2));
        co_await all_of(sequential(delay(5), worker(2)), delay(7));
        co_await sequential(worker(1), timeout(2, 3), worker(3).priority(-2));
        co_await (worker(1), delay(2), yield());
    }
};

}  // namespace recorded
