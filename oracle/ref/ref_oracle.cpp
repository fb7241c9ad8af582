// oracle/ref/ref_oracle.cpp -- TEST INFRASTRUCTURE (oracle), not product code.
//
// Primary oracle: the config models of BASELINE.json written against the
// UNMODIFIED reference engine (include path /root/reference/include; nothing is
// copied into this repository), with the example RNG
// (examples/random_variable.hpp) replaced by the shared Philox stream.
// Built by oracle/Makefile into oracle/_ref/libcxxdes_ref.so; only tests,
// __graft_entry__.smoke() and bench.py's reference/cpu_baseline legs load it.
#include <atomic>
#include <cstring>
#include <thread>

#include "harness.hpp"
#include "models.hpp"

using namespace ref_oracle;

namespace {

template <typename F>
void parallel_for(uint64_t n, int n_threads, F &&body) {
    if (n_threads <= 1) {
        for (uint64_t i = 0; i < n; ++i) body(i);
        return;
    }
    std::atomic<uint64_t> next{0};
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t)
        pool.emplace_back([&] {
            for (;;) {
                uint64_t i = next.fetch_add(1, std::memory_order_relaxed);
                if (i >= n) break;
                body(i);
            }
        });
    for (auto &th : pool) th.join();
}

void store(const cxxdes_b200_columns *out, uint64_t i, const rep_result &r) {
    if (out->end_time) out->end_time[i] = r.end_time;
    if (out->n_events) out->n_events[i] = r.n_events;
    if (out->trace_hash) out->trace_hash[i] = r.trace_hash;
    if (out->status) out->status[i] = r.status;
    for (int k = 0; k < CXXDES_B200_N_F64; ++k)
        if (out->f64[k]) out->f64[k][i] = r.f64[k];
    for (int k = 0; k < CXXDES_B200_N_I64; ++k)
        if (out->i64[k]) out->i64[k][i] = r.i64[k];
}

}  // namespace

extern "C" {

// Run replications [first_rep, first_rep + n_reps) of `model` on the reference
// engine.  run_until < 0: env.run(); otherwise env.run_until(run_until).
int cxxdes_ref_run(int model, const void *params, const cxxdes_b200_clock *clock, uint64_t seed,
                   uint64_t first_rep, uint64_t n_reps, int n_threads, int64_t run_until,
                   const cxxdes_b200_columns *out) {
    try {
        parallel_for(n_reps, n_threads, [&](uint64_t i) {
            rep_result r = run_model(model, params, *clock, seed, first_rep + i, run_until,
                                     nullptr, 0);
            store(out, i, r);
        });
    } catch (std::exception const &) {
        return 1;
    }
    return 0;
}

// Dump the first `max_events` popped tokens of one replication; returns the
// replication's total event count.
uint64_t cxxdes_ref_trace(int model, const void *params, const cxxdes_b200_clock *clock,
                          uint64_t seed, uint64_t rep, int64_t run_until, uint64_t max_events,
                          int64_t *time, int64_t *priority, uint32_t *kind, uint32_t *proc,
                          uint64_t *n_dumped) {
    std::vector<trace_event> dump;
    rep_result r = run_model(model, params, *clock, seed, rep, run_until, &dump, max_events);
    for (size_t k = 0; k < dump.size(); ++k) {
        time[k] = dump[k].time;
        priority[k] = dump[k].priority;
        kind[k] = dump[k].kind;
        proc[k] = dump[k].proc;
    }
    *n_dumped = dump.size();
    return r.n_events;
}

int cxxdes_ref_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
