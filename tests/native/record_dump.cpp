// tests/native/record_dump.cpp -- TEST INFRASTRUCTURE.  Compiled with -DCXXDES_B200_RECORDING and
// -I include/cxxdes_b200/record: the models of recorded_models.hpp, real coroutine bodies, are RECORDED into process
// programs and printed as JSON (same shape as program_builder_dump.cpp).  No simulation runs here.
#include <cinttypes>
#include <cstdio>

#include "program_dump.hpp"
#include "recorded_models.hpp"

using cxxdes::b200::program;

static void dump(const char *name, const program &prog, cxxdes::b200::time unit, cxxdes::b200::time prec, bool last = false) {
    dump_program(stdout, name, prog, unit, prec, last);
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
