// tests/native/record_dump.cpp -- TEST INFRASTRUCTURE.  Compiled with -DCXXDES_B200_RECORDING and
// -I include/cxxdes_b200/record: the models of recorded_models.hpp, real coroutine bodies, are RECORDED into process
// programs and printed as JSON (same shape as program_builder_dump.cpp).  No simulation runs here.
#include <cinttypes>
#include <cstdio>

#include "recorded_models.hpp"

using cxxdes::b200::program;

static void dump(const char *name, const program &prog, cxxdes::b200::time unit, cxxdes::b200::time prec, bool last = false) {
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
    for (std::uint32_t i = 0; i < p.n_rates; ++i) std::printf("%s%.17g", i ? ", " : "", p.rates[i]);
    std::printf("], \"n_mutexes\": %u, \"n_events\": %u, \"clock\": [[%" PRIdMAX ", %d], [%" PRIdMAX ", %d]]}%s\n", p.n_mutexes,
                p.n_events, unit.t, (int)unit.u, prec.t, (int)prec.u, last ? "" : ",");
}

template <typename Sim>
static void record(const char *name, Sim &&sim, bool last = false) {
    const program prog = sim.record();
    dump(name, prog, sim.env.time_unit(), sim.env.time_precision(), last);
}

int main() {
    std::printf("{\n");
    record("resource", recorded::resource_example{});
    record("mutex", recorded::mutex_example{});
    record("producer_consumer", recorded::producer_consumer{5.0, 10.0, 400});
    record("clocks", recorded::clocks{});
    record("mixed", recorded::mixed{});
    record("co_with", recorded::co_with_example{});
    record("sequential", recorded::sequential_example{}, true);
    std::printf("}\n");
    return 0;
}
