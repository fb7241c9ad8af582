// oracle/ref/models.hpp -- TEST INFRASTRUCTURE (oracle), not product code.
//
// The BASELINE.json config models as user code on the reference engine.  Each
// model keeps the process bodies of the reference example it comes from and
// changes only (1) the random source: shared Philox stream instead of
// std::mt19937_64 seeded from std::random_device
// (examples/random_variable.hpp:17,33-37) and (2) every process object is passed
// through tracer::named() when created so the harness can name it in the digest.
#pragma once
#include "harness.hpp"

#include <cxxdes/sync/queue.hpp>

namespace ref_oracle {

using namespace cxxdes::core::time_ops;

struct rep_result {
    int64_t end_time = 0;
    uint64_t n_events = 0;
    uint64_t trace_hash = 0;
    uint32_t status = 0;
    double f64[CXXDES_B200_N_F64] = {0, 0};
    int64_t i64[CXXDES_B200_N_I64] = {0, 0};
};

// Shared exponential variate: draw number `n` of stream `stream`.
struct exp_stream {
    cxxdes_b200::stream_addr addr;
    cxxdes_b200::divisor rate;
    uint64_t n = 0;
    exp_stream(uint64_t seed, uint64_t rep, uint32_t stream, double lambda)
        : addr{seed, rep, stream}, rate{cxxdes_b200::make_divisor(lambda)} {}
    double operator()() {
        return cxxdes_b200::exponential(addr, n++, rate, cxxdes_b200::LOG_TABLE_HOST);
    }
};

// ---------------------------------------------------------------------------
// M/M/1: examples/producer_consumer.cpp:9-59, process bodies verbatim in
// behaviour; stream ids: producer = 1, consumer = 2.
// ---------------------------------------------------------------------------
CXXDES_SIMULATION(mm1_model) {
    mm1_model(cxxdes_b200_mm1_params const &p, cxxdes_b200_clock const &c, uint64_t seed,
              uint64_t rep, tracer &tr)
        : n_packets{p.n_packets}, lambda{seed, rep, 1, p.lambda}, mu{seed, rep, 2, p.mu}, tr{tr} {
        env.time_unit(to_time(c.unit_t, c.unit_u));
        env.time_precision(to_time(c.precision_t, c.precision_u));
        bind();
    }

    cxxdes::sync::queue<double> q;
    std::size_t n_packets;
    double total_latency = 0;
    exp_stream lambda;
    exp_stream mu;
    double avg_latency = 0.0;
    std::size_t consumed = 0;
    std::size_t max_queue = 0;
    tracer &tr;

    coroutine<> producer() {
        for (std::size_t i = 0; i < n_packets; ++i) {
            co_await q.put(now_seconds());
            if (q.size() > max_queue) max_queue = q.size();
            co_await env.timeout(lambda());
        }
    }

    coroutine<> consumer() {
        std::size_t n = 0;
        while (true) {
            auto x = co_await q.pop();
            ++n;
            consumed = n;
            if (n == n_packets) {
                avg_latency = total_latency / n_packets;
                co_return;
            }
            co_await env.timeout(mu());
            total_latency += (now_seconds() - x);
        }
    }

    coroutine<> main_body() {
        co_await (tr.named(producer(), 1) && tr.named(consumer(), 2));
    }

    coroutine<> co_main() { return tr.named(main_body(), 0); }
};

inline rep_result run_mm1(const void *params, cxxdes_b200_clock const &clock, uint64_t seed,
                          uint64_t rep, int64_t until, tracer &tr) {
    auto const &p = *static_cast<cxxdes_b200_mm1_params const *>(params);
    mm1_model sim{p, clock, seed, rep, tr};
    run_traced(sim.env, tr, until);
    rep_result r;
    r.end_time = sim.now();
    r.f64[0] = sim.avg_latency;
    r.f64[1] = sim.total_latency;
    r.i64[0] = (int64_t)sim.consumed;
    r.i64[1] = (int64_t)sim.max_queue;
    return r;
}

inline rep_result run_model(int model, const void *params, cxxdes_b200_clock const &clock,
                            uint64_t seed, uint64_t rep, int64_t until,
                            std::vector<trace_event> *dump, size_t dump_limit) {
    tracer tr;
    tr.dump = dump;
    tr.dump_limit = dump_limit;
    rep_result r;
    switch (model) {
        case CXXDES_B200_MODEL_MM1: r = run_mm1(params, clock, seed, rep, until, tr); break;
        default: throw std::runtime_error("ref_oracle: unknown model");
    }
    r.n_events = tr.n_events;
    r.trace_hash = tr.hash;
    return r;
}

}  // namespace ref_oracle
