// oracle/port/models.hpp -- TEST INFRASTRUCTURE (oracle), not product code.
//
// The BASELINE.json config models lowered by hand to resume-point state machines
// on the restated engine (engine.hpp).  Each `resume` follows the process body of
// the cited reference example statement by statement; pushes happen in the same
// program order as in the coroutine.
#pragma once
#include <deque>

#include "engine.hpp"

namespace port_oracle {

struct rep_result {
    int64_t end_time = 0;
    uint64_t n_events = 0;
    uint64_t trace_hash = 0;
    uint32_t status = 0;
    double f64[CXXDES_B200_N_F64] = {0, 0};
    int64_t i64[CXXDES_B200_N_I64] = {0, 0};
    uint64_t peak_heap = 0;
};

// time_unit().count(time_precision()): misc/time.hpp:75-86.
inline int64_t time_count(int64_t t, int u, int64_t prec_t, int prec_u) {
    static const int64_t conv[5] = {1, 1000, 1000000, 1000000000, 1000000000000};
    if (u < prec_u) return conv[prec_u - u] * t / prec_t;
    return (t / prec_t) / conv[u - prec_u];
}

inline cxxdes_b200::clock_config fold_clock(const cxxdes_b200_clock &c) {
    cxxdes_b200::clock_config k;
    k.ticks_per_unit = time_count(c.unit_t, c.unit_u, c.precision_t, c.precision_u);
    k.tick_mult = c.precision_t;
    static const double conv[5] = {1, 1e-3, 1e-6, 1e-9, 1e-12};
    k.seconds_per_tick_unit = conv[c.precision_u];
    return k;
}

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
// M/M/1 -- examples/producer_consumer.cpp:9-59.
// ---------------------------------------------------------------------------
struct mm1_model {
    engine eng;
    cxxdes_b200::clock_config clk;
    uint64_t n_packets;
    exp_stream lambda, mu;
    std::deque<double> q;  // sync::queue<double>::q_, queue.hpp:98
    event q_event;         // sync::queue<double>::event_, queue.hpp:100
    int32_t p_main, p_prod, p_cons;
    uint64_t i = 0;        // producer loop index, producer_consumer.cpp:32
    uint64_t n = 0;        // consumer count, :39
    double x = 0;          // popped timestamp, :42
    double total_latency = 0, avg_latency = 0;
    uint64_t max_queue = 0;

    mm1_model(const cxxdes_b200_mm1_params &p, const cxxdes_b200_clock &c, uint64_t seed,
              uint64_t rep)
        : clk(fold_clock(c)), n_packets(p.n_packets), lambda(seed, rep, 1, p.lambda),
          mu(seed, rep, 2, p.mu) {
        p_main = eng.add_process(0);
        p_prod = eng.add_process(1);
        p_cons = eng.add_process(2);
        eng.bind_root(p_main);  // simulation::bind, simulation.hpp:76-81
    }

    double now_seconds() const { return cxxdes_b200::sim_to_seconds(eng.now, clk); }

    void resume(int32_t p) {
        if (p == p_main) resume_main();
        else if (p == p_prod) resume_producer();
        else resume_consumer();
    }

    // co_main: co_await (producer() && consumer());  producer_consumer.cpp:56-58
    void resume_main() {
        process &me = eng.procs[p_main];
        if (me.pc == 0) {
            // all_of::await_bind binds children in argument order (any_of.ipp:41-48)
            eng.bind(p_prod, me.priority);
            eng.bind(p_cons, me.priority);
            // await_ready: neither complete -> remaining = 2 (any_of.ipp:50-64)
            // await_suspend: output token + shared handler on both completion tokens (:66-84)
            int32_t h = eng.add_handler(true, 2, 2, token{0, me.priority, p_main, NO_HANDLER});
            eng.await_completion(p_prod, p_main, h);
            eng.await_completion(p_cons, p_main, h);
            me.pc = 1;
        } else {
            eng.do_return(p_main);
        }
    }

    // producer: producer_consumer.cpp:31-36
    void resume_producer() {
        if (i < n_packets) {
            // q.put(now_seconds()): unbounded -> can_put; emplace; wake (queue.hpp:48-53)
            q.push_back(now_seconds());
            eng.wake(q_event);
            if (q.size() > max_queue) max_queue = q.size();
            // env.timeout(lambda()): timeout.ipp:184-187
            eng.timeout(p_prod, cxxdes_b200::real_to_sim(lambda(), clk));
            ++i;
        } else {
            eng.do_return(p_prod);
        }
    }

    // consumer: producer_consumer.cpp:38-54
    void resume_consumer() {
        process &me = eng.procs[p_cons];
        if (me.pc == 2) total_latency += (now_seconds() - x);  // :52
        // q.pop(): while (!can_pop) wait; front; pop; wake (queue.hpp:58-65)
        if (q.empty()) {
            eng.wait(q_event, p_cons);
            me.pc = 1;
            return;
        }
        x = q.front();
        q.pop_front();
        eng.wake(q_event);
        ++n;
        if (n == n_packets) {
            avg_latency = total_latency / (double)n_packets;  // :46
            eng.do_return(p_cons);
            return;
        }
        eng.timeout(p_cons, cxxdes_b200::real_to_sim(mu(), clk));  // :50
        me.pc = 2;
    }

    rep_result result() const {
        rep_result r;
        r.f64[0] = avg_latency;
        r.f64[1] = total_latency;
        r.i64[0] = (int64_t)n;
        r.i64[1] = (int64_t)max_queue;
        return r;
    }
};

rep_result run_model(int model, const void *params, const cxxdes_b200_clock &clock, uint64_t seed,
                     uint64_t rep, int64_t until, std::vector<trace_event> *dump,
                     size_t dump_limit);

}  // namespace port_oracle
