// tests/native/example_reference.cpp -- TEST INFRASTRUCTURE.  The same example source file (-DEXAMPLE_FILE; main() renamed)
// compiled against the UNMODIFIED reference headers and run on the reference engine, observed through its public API
// (oracle/ref/harness.hpp: env.next_event() before every env.step()).  Writes the event count, the final clock and the
// (time, priority, kind) of every popped token as JSON to argv[1].  Shapes as in example_recorder.cpp:
//   -DEXAMPLE_SIM=class [-DEXAMPLE_ARGS=...] [-DEXAMPLE_UNTIL=t]: the harness binds co_main and steps the environment;
//   -DEXAMPLE_MAIN: the example's own main() runs; the word `environment` in the example's text names a class derived
//       from the reference's whose run() / run_until() / run_for() step the base environment one observed event at a
//       time (nothing else differs: the base class does all the work).
#include <cinttypes>
#include <cstdio>

#include "harness.hpp"

static ref_oracle::tracer g_tracer;
static std::vector<ref_oracle::trace_event> g_rows;
static std::intmax_t g_end_time = 0;

#if defined(EXAMPLE_MAIN)
namespace cxxdes::core {      // (next to the class it stands in for: examples name it `environment`, `des::environment`, ...)
struct observed_environment : environment {
    using environment::environment;
    void run() { observe(-1); }
    void run_until(time_integral t) { observe(t); }
    void run_for(time_integral t) { observe(now() + t); }

private:
    void observe(std::int64_t until) {
        ref_oracle::run_traced(*this, g_tracer, until);
        g_end_time = now();
    }
};
}  // namespace cxxdes::core
#define environment observed_environment
#endif

#if !defined(EXAMPLE_MAIN)
#define main cxxdes_example_main
#endif
#include EXAMPLE_FILE
#undef main
#undef environment

#include <cstdlib>

static int write_rows(const char *path) {
    std::FILE *out = std::fopen(path, "w");
    if (!out) return 2;
    std::fprintf(out, "{\"n_events\": %" PRIu64 ", \"end_time\": %" PRIdMAX ", \"rows\": [", g_tracer.n_events, g_end_time);
    for (size_t k = 0; k < g_rows.size(); ++k)
        std::fprintf(out, "%s[%" PRId64 ", %" PRId64 ", %u]", k ? ", " : "", g_rows[k].time, g_rows[k].priority, g_rows[k].kind);
    std::fprintf(out, "]}\n");
    std::fclose(out);
    return 0;
}

#if defined(EXAMPLE_MAIN)
// the example's own main() is THE main; what it dispatched is written when the process exits ($EXAMPLE_OUT)
static void write_at_exit() {
    if (const char *path = std::getenv("EXAMPLE_OUT")) write_rows(path);
}
static const int registered_at_start = (g_tracer.dump = &g_rows, g_tracer.dump_limit = 20000, std::atexit(write_at_exit));
#else
// `bind` is protected in simulation<>; the harness needs to bind co_main without running
template <typename Sim>
struct opened : Sim {
    using Sim::Sim;
    void bind_now() { this->bind(); }
};

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    g_tracer.dump = &g_rows;
    g_tracer.dump_limit = 20000;
#if defined(EXAMPLE_ARGS)
    opened<EXAMPLE_SIM> sim{EXAMPLE_ARGS};
#else
    opened<EXAMPLE_SIM> sim;
#endif
    sim.bind_now();
#if defined(EXAMPLE_UNTIL)
    ref_oracle::run_traced(sim.env, g_tracer, EXAMPLE_UNTIL);
#else
    ref_oracle::run_traced(sim.env, g_tracer, -1);
#endif
    g_end_time = sim.env.now();
    return write_rows(argv[1]);
}
#endif
