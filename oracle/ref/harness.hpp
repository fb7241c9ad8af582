// oracle/ref/harness.hpp -- TEST INFRASTRUCTURE (oracle), not product code.
//
// Runs models on the UNMODIFIED reference engine (headers under
// /root/reference/include, compiled where they lie) and observes the dispatch
// loop from outside through public API only:
//     while (token *t = env.next_event()) { observe(t); env.step(); }
// (environment.ipp:101-106 next_event, :117-146 step).  Process ordinals are
// attached to each process when its coroutine object is created (public
// accessor coroutine::coro_data(), coroutine.ipp:76-78), which schedules nothing.
#pragma once
#include <cxxdes/cxxdes.hpp>

#include <cstdint>
#include <unordered_map>
#include <vector>

#include "cxxdes_b200.h"
#include "cxxdes_b200/shared/philox.hpp"
#include "cxxdes_b200/shared/trace_hash.hpp"
#include "cxxdes_b200/shared/variates.hpp"

namespace ref_oracle {

using namespace cxxdes::core;

struct rep_result {
    int64_t end_time = 0;
    uint64_t n_events = 0;
    uint64_t trace_hash = 0;
    uint32_t status = 0;
    double f64[CXXDES_B200_N_F64] = {0, 0};
    int64_t i64[CXXDES_B200_N_I64] = {0, 0};
};

struct trace_event {
    int64_t time;
    int64_t priority;
    uint32_t kind;
    uint32_t proc;
};

struct tracer {
    std::unordered_map<const void *, uint32_t> ids;
    uint64_t hash = cxxdes_b200::TRACE_HASH_INIT;
    uint64_t n_events = 0;
    std::vector<trace_event> *dump = nullptr;
    size_t dump_limit = 0;

    // Registered coroutine_data objects are pinned (intrusive reference) for the
    // tracer's lifetime so an address is never reused by a later, unregistered
    // helper coroutine.
    std::vector<cxxdes::memory::ptr<const coroutine_data>> pins;

    void reg(const coroutine_data *coro, uint32_t id) {
        ids[coro] = id;
        pins.emplace_back(coro);
    }

    // Name a process before it is started, keeping the value category of the
    // expression (so `named(producer(), 1) && named(consumer(), 2)` builds the
    // same all_of as `producer() && consumer()`).
    template <typename Co>
    Co &&named(Co &&c, uint32_t id) {
        reg(c.coro_data(), id);
        return std::forward<Co>(c);
    }

    // Library-created helper coroutines (_Co_with, sequential) cannot register
    // themselves; they are identified through their parent link
    // (coroutine_data.ipp:41-49) and reported as `parent | 0x8000`.
    uint32_t lookup(coroutine_data *cd) {
        if (!cd) return cxxdes_b200::PROC_NONE;
        auto it = ids.find(cd);
        if (it != ids.end()) return it->second;
        coroutine_data *p = cd->parent().get();
        for (int depth = 0; p && depth < 16; ++depth) {
            auto jt = ids.find(p);
            if (jt != ids.end()) return jt->second | 0x8000u;
            p = p->parent().get();
        }
        return 0x7FFFu;
    }

    void observe(token *t) {
        uint32_t kind = t->handler ? cxxdes_b200::TOKEN_HANDLER
                                   : (t->coro_data ? cxxdes_b200::TOKEN_RESUME
                                                   : cxxdes_b200::TOKEN_NOOP);
        uint32_t proc = lookup(t->coro_data.get());
        hash = cxxdes_b200::trace_hash_step(hash, t->time, t->priority, kind, proc);
        ++n_events;
        if (dump && dump->size() < dump_limit)
            dump->push_back(trace_event{t->time, t->priority, kind, proc});
    }
};

inline cxxdes::core::time to_time(int64_t t, int32_t u) {
    return cxxdes::core::time{(std::intmax_t)t, (cxxdes::core::time_units)u};
}

// env.run() with observation; `until < 0` means run to exhaustion, otherwise
// environment::run_until(until) semantics (environment.ipp:190-196).
inline void run_traced(environment &env, tracer &tr, int64_t until) {
    if (until < 0) {
        while (token *t = env.next_event()) {
            tr.observe(t);
            env.step();
        }
    } else {
        while (env.next_event() && env.next_event()->time <= until) {
            tr.observe(env.next_event());
            env.step();
        }
        env.run_until((time_integral)until);  // no events left <= until: only advances now_
    }
}

inline cxxdes_b200::clock_config fold_clock(const cxxdes_b200_clock &c) {
    cxxdes_b200::clock_config k;
    auto unit = to_time(c.unit_t, c.unit_u);
    auto prec = to_time(c.precision_t, c.precision_u);
    k.ticks_per_unit = (int64_t)unit.count(prec);
    k.tick_mult = c.precision_t;
    static const double conv[5] = {1, 1e-3, 1e-6, 1e-9, 1e-12};
    k.seconds_per_tick_unit = conv[c.precision_u];
    return k;
}

}  // namespace ref_oracle
