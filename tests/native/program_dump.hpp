// tests/native/program_dump.hpp -- TEST INFRASTRUCTURE.  A process program (cxxdes_b200/program.hpp) as JSON tables, the
// shape tests/test_record.py::program_params reads (same as program_builder_dump.cpp).
#pragma once
#include <cinttypes>
#include <cstdio>

#include "cxxdes_b200/program.hpp"

static void dump_program(std::FILE *out, const char *name, const cxxdes::b200::program &prog, cxxdes::b200::time unit,
                         cxxdes::b200::time prec, bool last = false) {

    cxxdes_b200_program p = prog.c_struct();
    std::fprintf(out, "\"%s\": {\"ops\": [", name);
    for (std::uint32_t i = 0; i < p.n_ops; ++i)
        std::fprintf(out, "%s[%u, %u, %d, %" PRId64 "]", i ? ", " : "", p.ops[i].code, p.ops[i].a, p.ops[i].b, p.ops[i].imm);
    std::fprintf(out, "], \"procs\": [");
    for (std::uint32_t i = 0; i < p.n_procs; ++i)
        std::fprintf(out, "%s[%u, %u, %" PRId64 ", %" PRId64 ", %" PRId64 ", %" PRId64 "]", i ? ", " : "", p.procs[i].entry,
                    p.procs[i].ordinal, p.procs[i].priority, p.procs[i].latency, p.procs[i].return_latency,
                    p.procs[i].return_priority);
    std::fprintf(out, "], \"roots\": [");
    for (std::uint32_t i = 0; i < p.n_roots; ++i) std::fprintf(out, "%s%u", i ? ", " : "", p.roots[i]);
    std::fprintf(out, "], \"queue_max\": [");
    for (std::uint32_t i = 0; i < p.n_queues; ++i) std::fprintf(out, "%s%" PRIu64, i ? ", " : "", p.queue_max[i]);
    std::fprintf(out, "], \"sems\": [");
    for (std::uint32_t i = 0; i < p.n_sems; ++i)
        std::fprintf(out, "%s[%" PRIu64 ", %" PRIu64 "]", i ? ", " : "", p.sem_init[i], p.sem_max[i]);
    std::fprintf(out, "], \"rates\": [");
    for (std::uint32_t i = 0; i < p.n_rates; ++i) std::fprintf(out, "%s%.17g", i ? ", " : "", p.rates[i]);
    std::fprintf(out, "], \"n_mutexes\": %u, \"n_events\": %u, \"clock\": [[%" PRIdMAX ", %d], [%" PRIdMAX ", %d]]}%s\n", p.n_mutexes,
                p.n_events, unit.t, (int)unit.u, prec.t, (int)prec.u, last ? "" : ",");
}
