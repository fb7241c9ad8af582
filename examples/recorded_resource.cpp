// examples/recorded_resource.cpp -- the reference's examples/resource.cpp, coroutine bodies untouched, on the B200.
//
//   g++ -std=c++20 -O2 -I include/cxxdes_b200/record -I include examples/recorded_resource.cpp
//       -L libcxxdes_b200/_lib -lcxxdes_b200 -Wl,-rpath,$PWD/libcxxdes_b200/_lib -o recorded_resource
//
// `-I include/cxxdes_b200/record` makes <cxxdes/cxxdes.hpp> the RECORDING implementation of the libcxxdes API: the
// bodies below run once on the host, every co_await is noted, and sim.ensemble(R) hands the resulting process program
// to the ensemble engine.  Compiled against the reference's own headers instead, the same model runs on its CPU engine
// (tests/native/record_reference.cpp does that for the parity tests).
#include <cstdio>
#include <cstdlib>

#include <cxxdes/cxxdes.hpp>

using namespace cxxdes::core;

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
        // in the end, coroutine #3 finishes at 12.
        co_return;
    }
};

int main(int argc, char **argv) {
    const std::uint64_t n = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 4096;
    try {
        resource_example sim;
        auto ens = sim.ensemble(n);
        ens->run();
        const auto acc = ens->reduce();
        std::printf("replications %llu, events %llu, errors %llu, now() of replication 0: %lld\n",
                    (unsigned long long)acc.n_replications, (unsigned long long)acc.n_events, (unsigned long long)acc.n_errors,
                    (long long)ens->now()[0]);
        return (acc.n_errors == 0 && ens->now()[0] == 12 && acc.n_events == 16 * n) ? 0 : 1;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 2;
    }
}
