// oracle/ref/program_runner.hpp -- TEST INFRASTRUCTURE (oracle), not product code.
//
// Runs ANY process program (cxxdes_b200_program) on the UNMODIFIED reference engine: every operation is
// executed by co_awaiting the real reference awaitable it stands for -- delay/until/lazy_timeout
// (timeout.ipp), coroutine objects with their options (coroutine.ipp:106-176), async (async.ipp),
// all_of / any_of (any_of.ipp), sync::queue / semaphore / mutex / event (sync/*.hpp).  Nothing of the
// interpreters' bookkeeping is reproduced here: the token traffic is whatever the library does.  This pins the
// semantics of the CUDA interpreters (and of the restated CPU interpreter) on the reference for arbitrary --
// in the tests: randomly generated -- programs, not only for the hand-written scenarios of kat_models.hpp:
//     reference engine (this file)  ==  oracle/port/program_interp.hpp  ==  program_fast.cu / program_kernel.cu
#pragma once
#include "harness.hpp"

#include <cxxdes/sync/event.hpp>
#include <cxxdes/sync/mutex.hpp>
#include <cxxdes/sync/queue.hpp>
#include <cxxdes/sync/semaphore.hpp>

#include <deque>
#include <memory>
#include <optional>
#include <variant>

namespace ref_oracle {

// One child of a composition whose arity and child kinds are only known at run time: a user-level awaitable
// (it satisfies the library's `awaitable` concept, awaitable.ipp:14-27) that forwards every call to the real
// awaitable it holds, so `all_of.range(first, last)` (any_of.ipp:200-213) can take a mixed list.
struct nested_child;   // a composition used as a child: the library's own all_of / any_of object over any_child's

struct any_child {
    using delay_type = decltype(delay(time_integral{}, priority_type{}));
    using until_type = decltype(until(time_integral{}, priority_type{}));
    using wait_type = decltype(std::declval<cxxdes::sync::event &>().wait(time_integral{}, priority_type{}));
    std::variant<coroutine<>, delay_type, until_type, std::shared_ptr<nested_child>, wait_type> v;

    void await_bind(environment *env, priority_type priority);
    bool await_ready();
    void await_suspend(coroutine_data_ptr coro_data);
    token *await_token();
    void await_resume() { await_resume(no_return_value_tag{}); }
    void await_resume(no_return_value_tag);
};

struct nested_child {
    using all_type = decltype(all_of.range(std::declval<std::vector<any_child>::iterator>(),
                                           std::declval<std::vector<any_child>::iterator>()));
    using any_type = decltype(any_of.range(std::declval<std::vector<any_child>::iterator>(),
                                           std::declval<std::vector<any_child>::iterator>()));
    std::variant<all_type, any_type> comp;
};

template <typename F>
inline auto visit_child(any_child &c, F &&f) {
    return std::visit(
        [&](auto &a) {
            if constexpr (std::is_same_v<std::remove_reference_t<decltype(a)>, std::shared_ptr<nested_child>>)
                return std::visit([&](auto &inner) { return f(inner); }, a->comp);
            else
                return f(a);
        },
        c.v);
}
inline void any_child::await_bind(environment *env, priority_type priority) {
    visit_child(*this, [&](auto &a) { a.await_bind(env, priority); });
}
inline bool any_child::await_ready() {
    return visit_child(*this, [](auto &a) -> bool { return a.await_ready(); });
}
inline void any_child::await_suspend(coroutine_data_ptr coro_data) {
    visit_child(*this, [&](auto &a) { a.await_suspend(coro_data); });
}
inline token *any_child::await_token() {
    return visit_child(*this, [](auto &a) -> token * { return a.await_token(); });
}
inline void any_child::await_resume(no_return_value_tag) {
    visit_child(*this, [](auto &a) { a.await_resume(no_return_value_tag{}); });
}
static_assert(awaitable<any_child>);

struct program_sim {
    environment env;
    tracer &tr;
    const cxxdes_b200_program &prg;
    cxxdes_b200::clock_config clk;
    uint64_t seed, rep;

    std::deque<cxxdes::sync::queue<int64_t>> queues;
    std::deque<cxxdes::sync::semaphore<uint64_t>> sems;
    std::deque<cxxdes::sync::mutex> mutexes;
    std::deque<std::optional<cxxdes::sync::mutex::handle>> handles;
    std::deque<cxxdes::sync::event> events;

    struct proc_state {
        std::optional<coroutine<>> co;
        int64_t reg = 0;
        uint64_t draws = 0;
    };
    std::vector<proc_state> procs;
    std::vector<cxxdes_b200::divisor> rates;
    double f64[CXXDES_B200_N_F64] = {0, 0};
    int64_t i64[CXXDES_B200_N_I64] = {0, 0};

    program_sim(const cxxdes_b200_program &p, cxxdes_b200_clock const &c, uint64_t seed_, uint64_t rep_, tracer &t)
        : tr{t}, prg{p}, clk{fold_clock(c)}, seed{seed_}, rep{rep_}, procs(p.n_procs) {
        env.time_unit(to_time(c.unit_t, c.unit_u));
        env.time_precision(to_time(c.precision_t, c.precision_u));
        for (uint32_t q = 0; q < p.n_queues; ++q) queues.emplace_back((std::size_t)p.queue_max[q]);
        for (uint32_t s = 0; s < p.n_sems; ++s) sems.emplace_back(p.sem_init[s], p.sem_max[s]);
        for (uint32_t m = 0; m < p.n_mutexes; ++m) {
            mutexes.emplace_back();
            handles.emplace_back();
        }
        for (uint32_t e = 0; e < p.n_events; ++e) events.emplace_back();
        const uint64_t row = p.sweep_steps ? rep_ % p.sweep_steps : 0;   // parameter sweep: this replication's rates
        for (uint32_t r = 0; r < p.n_rates; ++r) rates.push_back(cxxdes_b200::make_divisor(p.rates[row * p.n_rates + r]));
        for (uint32_t k = 0; k < p.n_procs; ++k) fresh(k);
        for (uint32_t r = 0; r < p.n_roots; ++r) env.bind(*procs[p.roots[r]].co);   // coroutine.ipp:352-357
    }

    static priority_type prio_of(int32_t b) {
        return b == CXXDES_B200_INHERIT ? priority_consts::inherit : (priority_type)b;
    }

    // a new coroutine object from the coroutine function, with its options, named for the digest
    void fresh(uint32_t p) {
        const cxxdes_b200_process &d = prg.procs[p];
        coroutine<> co = body(p);
        if (d.priority != INT64_MAX) co.priority((priority_type)d.priority);
        co.latency((time_integral)d.latency);
        co.return_latency((time_integral)d.return_latency);
        if (d.return_priority != INT64_MAX) co.return_priority((priority_type)d.return_priority);
        tr.reg(co.coro_data(), d.ordinal);
        procs[p].co.reset();
        procs[p].co.emplace(std::move(co));
    }

    // the children of a composition, built from the operations that follow it in pre-order; a nested composition
    // becomes the library's own all_of / any_of object (with its options) wrapped as a child
    std::vector<any_child> children_of(const cxxdes_b200_op &head, uint32_t &pos) {
        std::vector<any_child> children;
        for (uint32_t c = 0; c < head.a; ++c) {
            const cxxdes_b200_op ch = prg.ops[pos++];
            if (ch.code == CXXDES_B200_OP_CHILD_PROC) children.push_back(any_child{*procs[ch.a].co});
            else if (ch.code == CXXDES_B200_OP_CHILD_DELAY)
                children.push_back(any_child{delay((time_integral)ch.imm, prio_of(ch.b))});
            else if (ch.code == CXXDES_B200_OP_CHILD_UNTIL)
                children.push_back(any_child{until((time_integral)ch.imm, prio_of(ch.b))});
            else if (ch.code == CXXDES_B200_OP_CHILD_EV_WAIT)
                children.push_back(any_child{events[ch.a].wait((time_integral)ch.imm, prio_of(ch.b))});
            else {
                std::vector<any_child> inner = children_of(ch, pos);
                std::shared_ptr<nested_child> nested;
                if (ch.code == CXXDES_B200_OP_CHILD_ALL_OF) {
                    auto comp = all_of.range(inner.begin(), inner.end());
                    if (ch.b != CXXDES_B200_INHERIT) comp.priority((priority_type)ch.b);
                    comp.return_latency((time_integral)ch.imm);
                    nested.reset(new nested_child{std::move(comp)});
                } else {
                    auto comp = any_of.range(inner.begin(), inner.end());
                    if (ch.b != CXXDES_B200_INHERIT) comp.priority((priority_type)ch.b);
                    comp.return_latency((time_integral)ch.imm);
                    nested.reset(new nested_child{std::move(comp)});
                }
                children.push_back(any_child{nested});
            }
        }
        return children;
    }

    coroutine<> body(uint32_t p) {
        uint32_t pc = prg.procs[p].entry;
        int64_t loops[8];
        int depth = 0;
        for (;;) {
            const cxxdes_b200_op op = prg.ops[pc];
            proc_state &me = procs[p];
            switch (op.code) {
                case CXXDES_B200_OP_RETURN:
                    co_return;
                case CXXDES_B200_OP_DELAY:
                    co_await delay((time_integral)op.imm, prio_of(op.b));
                    break;
                case CXXDES_B200_OP_UNTIL:
                    co_await until((time_integral)op.imm, prio_of(op.b));
                    break;
                case CXXDES_B200_OP_DELAY_EXP: {
                    cxxdes_b200::stream_addr addr{seed, rep, prg.procs[p].ordinal};
                    const double v = cxxdes_b200::exponential(addr, me.draws++, rates[op.a], cxxdes_b200::LOG_TABLE_HOST);
                    co_await env.timeout(v);
                    break;
                }
                case CXXDES_B200_OP_DELAY_REAL: {
                    double v;
                    std::memcpy(&v, &op.imm, sizeof v);
                    co_await env.timeout(v);
                    break;
                }
                case CXXDES_B200_OP_FRESH:
                    fresh(op.a);
                    break;
                case CXXDES_B200_OP_AWAIT:
                    co_await *procs[op.a].co;
                    break;
                case CXXDES_B200_OP_ASYNC:
                    co_await async(*procs[op.a].co);
                    break;
                case CXXDES_B200_OP_ALL_OF:
                case CXXDES_B200_OP_ANY_OF: {
                    uint32_t pos = pc + 1;
                    std::vector<any_child> children = children_of(op, pos);
                    if (op.code == CXXDES_B200_OP_ALL_OF) {
                        auto comp = all_of.range(children.begin(), children.end());
                        if (op.b != CXXDES_B200_INHERIT) comp.priority((priority_type)op.b);
                        comp.return_latency((time_integral)op.imm);
                        co_await std::move(comp);
                    } else {
                        auto comp = any_of.range(children.begin(), children.end());
                        if (op.b != CXXDES_B200_INHERIT) comp.priority((priority_type)op.b);
                        comp.return_latency((time_integral)op.imm);
                        co_await std::move(comp);
                    }
                    pc = pos - 1;
                    break;
                }
                case CXXDES_B200_OP_Q_PUT:
                    co_await queues[op.a].put(op.b ? me.reg : (int64_t)env.now());
                    break;
                case CXXDES_B200_OP_Q_POP: {
                    const int64_t v = co_await queues[op.a].pop();
                    procs[p].reg = v;
                    break;
                }
                case CXXDES_B200_OP_SEM_DOWN:
                    co_await sems[op.a].down();
                    break;
                case CXXDES_B200_OP_SEM_UP:
                    co_await sems[op.a].up();
                    break;
                case CXXDES_B200_OP_MTX_ACQUIRE: {
                    auto h = co_await mutexes[op.a].acquire();
                    handles[op.a].emplace(std::move(h));
                    break;
                }
                case CXXDES_B200_OP_MTX_RELEASE: {
                    auto h = std::move(*handles[op.a]);
                    handles[op.a].reset();
                    co_await h.release();
                    break;
                }
                case CXXDES_B200_OP_EV_WAIT:
                    co_await events[op.a].wait((time_integral)op.imm, prio_of(op.b));
                    break;
                case CXXDES_B200_OP_EV_WAKE:
                    co_await events[op.a].wake();
                    break;
                case CXXDES_B200_OP_LOOP:
                    if (op.imm <= 0) {
                        pc = (uint32_t)op.b;
                        continue;
                    }
                    loops[depth++] = op.imm;
                    break;
                case CXXDES_B200_OP_END_LOOP:
                    if (--loops[depth - 1] > 0) {
                        pc = (uint32_t)op.b;
                        continue;
                    }
                    --depth;
                    break;
                case CXXDES_B200_OP_SET_REG_NOW:
                    me.reg = (int64_t)env.now();
                    break;
                case CXXDES_B200_OP_STAT_SECONDS:
                    f64[op.a] += env.now_seconds() - cxxdes_b200::sim_to_seconds(me.reg, clk);
                    break;
                case CXXDES_B200_OP_STAT_TICKS:
                    f64[op.a] += (double)((int64_t)env.now() - me.reg);
                    break;
                case CXXDES_B200_OP_COUNT:
                    i64[op.a] += op.imm;
                    break;
                case CXXDES_B200_OP_MARK:
                    ++i64[0];
                    i64[1] = (int64_t)((uint64_t)i64[1] * 1000003ULL + (uint64_t)((int64_t)env.now() * op.imm + (int64_t)op.b));
                    break;
                case CXXDES_B200_OP_IF: {   // plain C++ between the awaits, on the library's own accessors
                    const uint32_t idx = op.a & 0xFFu, src = (op.a >> 8) & 0xFu, cmp = (op.a >> 12) & 0x7u;
                    int64_t v = 0;
                    switch (src) {
                        case CXXDES_B200_SRC_COUNTER: v = i64[idx]; break;
                        case CXXDES_B200_SRC_QUEUE_SIZE: v = (int64_t)queues[idx].size(); break;
                        case CXXDES_B200_SRC_SEM_VALUE: v = (int64_t)sems[idx].value(); break;
                        case CXXDES_B200_SRC_NOW: v = (int64_t)env.now(); break;
                        case CXXDES_B200_SRC_REGISTER: v = me.reg; break;
                        default: v = mutexes[idx].is_acquired() ? 1 : 0; break;
                    }
                    const bool t = cmp == CXXDES_B200_CMP_EQ ? v == op.imm : cmp == CXXDES_B200_CMP_NE ? v != op.imm
                                 : cmp == CXXDES_B200_CMP_LT ? v < op.imm : cmp == CXXDES_B200_CMP_LE ? v <= op.imm
                                 : cmp == CXXDES_B200_CMP_GT ? v > op.imm : v >= op.imm;
                    if (!t) {
                        pc = (uint32_t)op.b;
                        continue;
                    }
                    break;
                }
                case CXXDES_B200_OP_JUMP:
                    pc = (uint32_t)op.b;
                    continue;
                case CXXDES_B200_OP_LAZY_TIMEOUT: {
                    auto t = co_await lazy_timeout((time_integral)op.imm);   // no event; t is an `instant`
                    procs[p].reg = (int64_t)t.time();
                    break;
                }
                case CXXDES_B200_OP_UNTIL_REG:
                    co_await instant((time_integral)me.reg, prio_of(op.b));
                    break;
                default:
                    throw std::runtime_error("program_runner: unknown op");
            }
            ++pc;
        }
    }
};

inline rep_result run_program(const void *params, cxxdes_b200_clock const &clock, uint64_t seed, uint64_t rep,
                              int64_t until, tracer &tr) {
    auto const &p = *static_cast<cxxdes_b200_program const *>(params);
    program_sim sim{p, clock, seed, rep, tr};
    run_traced(sim.env, tr, until);
    rep_result r;
    r.end_time = (int64_t)sim.env.now();
    for (int k = 0; k < CXXDES_B200_N_F64; ++k) r.f64[k] = sim.f64[k];
    for (int k = 0; k < CXXDES_B200_N_I64; ++k) r.i64[k] = sim.i64[k];
    return r;
}

}  // namespace ref_oracle
