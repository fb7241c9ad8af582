// tests/native/example_reference.cpp -- TEST INFRASTRUCTURE.  The same example source file (-DEXAMPLE_FILE, -DEXAMPLE_SIM;
// main() renamed) compiled against the UNMODIFIED reference headers and run on the reference engine, observed through its
// public API (oracle/ref/harness.hpp: env.next_event() before every env.step()).  Writes the event count, the final clock
// and the (time, priority, kind) of every popped token as JSON to argv[1].
#include <cinttypes>
#include <cstdio>

#include "harness.hpp"

#define main cxxdes_example_main
#include EXAMPLE_FILE
#undef main

// `bind` is protected in simulation<>; the harness needs to bind co_main without running
template <typename Sim>
struct opened : Sim {
    using Sim::Sim;
    void bind_now() { this->bind(); }
};

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    opened<EXAMPLE_SIM> sim;
    ref_oracle::tracer tr;
    std::vector<ref_oracle::trace_event> rows;
    tr.dump = &rows;
    tr.dump_limit = 20000;
    sim.bind_now();
    ref_oracle::run_traced(sim.env, tr, -1);
    std::FILE *out = std::fopen(argv[1], "w");
    if (!out) return 2;
    std::fprintf(out, "{\"n_events\": %" PRIu64 ", \"end_time\": %" PRIdMAX ", \"rows\": [", tr.n_events, (std::intmax_t)sim.env.now());
    for (size_t k = 0; k < rows.size(); ++k)
        std::fprintf(out, "%s[%" PRId64 ", %" PRId64 ", %u]", k ? ", " : "", rows[k].time, rows[k].priority, rows[k].kind);
    std::fprintf(out, "]}\n");
    std::fclose(out);
    return 0;
}
