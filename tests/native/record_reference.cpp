// tests/native/record_reference.cpp -- TEST INFRASTRUCTURE.  The models of recorded_models.hpp compiled against the
// UNMODIFIED reference headers (-I /root/reference/include) and run on the reference engine, observed from outside
// through its public API (oracle/ref/harness.hpp: env.next_event() before every env.step()).  Prints, per model and
// replication, the event count, the final clock and the (time, priority, kind) of every popped token as JSON.
#include <cinttypes>
#include <cstdio>

#include "harness.hpp"
#include "recorded_models.hpp"

// `bind` is protected in simulation<>; the harness needs to bind co_main without running
template <typename Sim>
struct opened : Sim {
    using Sim::Sim;
    void bind_now() { this->bind(); }
};

template <typename Sim, typename... Args>
static void run(const char *name, int n_reps, int64_t until, bool last, Args... args) {
    std::printf("\"%s\": [", name);
    for (int r = 0; r < n_reps; ++r) {
        cxxdes_b200::rv_context::seed() = 7;
        cxxdes_b200::rv_context::replication() = 3 + (uint64_t)r;
        opened<Sim> sim{args...};
        ref_oracle::tracer tr;
        std::vector<ref_oracle::trace_event> rows;
        tr.dump = &rows;
        tr.dump_limit = 4000;
        sim.bind_now();
        ref_oracle::run_traced(sim.env, tr, until);
        std::printf("%s{\"n_events\": %" PRIu64 ", \"end_time\": %" PRIdMAX ", \"rows\": [", r ? ", " : "", tr.n_events,
                    (std::intmax_t)sim.env.now());
        for (size_t k = 0; k < rows.size(); ++k)
            std::printf("%s[%" PRId64 ", %" PRId64 ", %u]", k ? ", " : "", rows[k].time, rows[k].priority, rows[k].kind);
        std::printf("]}");
    }
    std::printf("]%s\n", last ? "" : ",");
}

int main() {
    std::printf("{\n");
    run<recorded::resource_example>("resource", 1, -1, false);
    run<recorded::mutex_example>("mutex", 1, -1, false);
    run<recorded::producer_consumer>("producer_consumer", 4, -1, false, 5.0, 10.0, (std::size_t)400);
    run<recorded::clocks>("clocks", 1, 20, false);
    run<recorded::mixed>("mixed", 1, -1, false);
    run<recorded::co_with_example>("co_with", 1, -1, false);
    run<recorded::sequential_example>("sequential", 1, -1, true);
    std::printf("}\n");
    return 0;
}
