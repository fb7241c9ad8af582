// oracle/port/program_interp.hpp -- TEST INFRASTRUCTURE (oracle), not product code.
//
// CPU interpreter of process programs (cxxdes_b200_program) on the restated engine.  Written
// independently of the CUDA interpreter (csrc/program_kernel.cu): unbounded std::vector containers,
// one handler object per composition, no pools.  Both are checked against the same scenarios
// written natively on the reference engine (oracle/ref/kat_models.hpp).
#pragma once
#include <deque>

#include "engine.hpp"
#include "models.hpp"

namespace port_oracle {

struct program_model {
    engine eng;
    cxxdes_b200::clock_config clk;
    const cxxdes_b200_program &prg;
    cxxdes_b200::stream_addr addr;
    uint64_t rate_row;   // parameter sweep: this replication's row of prg.rates (cxxdes_b200.h, sweep_steps)

    struct runtime {
        uint32_t pc = 0;
        uint32_t draws = 0;
        int64_t reg = 0;
        std::vector<uint32_t> loops;
    };
    std::vector<runtime> rt;
    std::vector<std::deque<int64_t>> queues;
    std::vector<event> q_event, s_event, m_event, e_event;
    std::vector<uint64_t> sem_value;
    std::vector<char> mtx_owned;
    double f64[CXXDES_B200_N_F64] = {0, 0};
    int64_t i64[CXXDES_B200_N_I64] = {0, 0};
    uint32_t status = 0;

    program_model(const cxxdes_b200_program &p, const cxxdes_b200_clock &c, uint64_t seed, uint64_t rep)
        : clk(fold_clock(c)), prg(p), addr{seed, rep, 0}, rate_row(p.sweep_steps ? rep % p.sweep_steps : 0) {
        rt.resize(p.n_procs);
        for (uint32_t i = 0; i < p.n_procs; ++i) {
            eng.add_process(p.procs[i].ordinal, p.procs[i].priority, p.procs[i].latency);
            rt[i].pc = p.procs[i].entry;
        }
        queues.resize(p.n_queues);
        q_event.resize(p.n_queues);
        s_event.resize(p.n_sems);
        m_event.resize(p.n_mutexes);
        e_event.resize(p.n_events);
        for (uint32_t i = 0; i < p.n_sems; ++i) sem_value.push_back(p.sem_init[i]);
        mtx_owned.assign(p.n_mutexes, 0);
        // environment::bind for every root (coroutine.ipp:352-357)
        for (uint32_t r = 0; r < p.n_roots; ++r) eng.bind_root((int32_t)p.roots[r]);
    }

    int64_t prio_of(const cxxdes_b200_op &op) const {
        return op.b == CXXDES_B200_INHERIT ? PRIO_INHERIT : (int64_t)op.b;
    }

    // a coroutine handle's return_latency / return_priority (coroutine.ipp:141-176) go into the
    // completion token created when it is awaited (coroutine.ipp:194-207)
    void await_completion(uint32_t child, int32_t awaiter, int32_t handler) {
        eng.await_completion((int32_t)child, awaiter, handler, prg.procs[child].return_latency,
                             prg.procs[child].return_priority);
    }

    void fresh(uint32_t p) {
        process &pr = eng.procs[p];
        pr.priority = prg.procs[p].priority;
        pr.bound = false;
        pr.complete = false;
        pr.completion.clear();
        rt[p].pc = prg.procs[p].entry;
        rt[p].loops.clear();
    }

    // all_of / any_of, possibly nested (any_of.ipp:41-92).  Restated recursively, the way the reference's awaitables
    // call each other: await_bind, await_ready, await_suspend of a composition call the same function of every
    // child.  `pos` walks the children, which follow the head in pre-order.
    struct comp_node {
        bool is_all = false, ready = false;
        size_t total = 0, remaining = 0;
        int64_t prio = 0, latency = 0;
        std::vector<int> kids;        // >= 0: index into nodes, < 0: ~op position of a leaf
        std::vector<bool> kid_ready;
    };
    std::vector<comp_node> nodes;

    int parse(const cxxdes_b200_op &head, bool is_all, int64_t inherited, uint32_t &pos) {
        const int id = (int)nodes.size();
        nodes.emplace_back();
        nodes[id].is_all = is_all;
        nodes[id].total = head.a;
        nodes[id].latency = head.imm;
        int64_t pr = prio_of(head);
        nodes[id].prio = pr == PRIO_INHERIT ? inherited : pr;     // await_bind: priority_ = inherited (:44-45)
        for (uint32_t c = 0; c < head.a; ++c) {
            const cxxdes_b200_op &ch = prg.ops[pos];
            if (ch.code == CXXDES_B200_OP_CHILD_ALL_OF || ch.code == CXXDES_B200_OP_CHILD_ANY_OF) {
                ++pos;
                const int kid = parse(ch, ch.code == CXXDES_B200_OP_CHILD_ALL_OF, nodes[id].prio, pos);
                nodes[id].kids.push_back(kid);
            } else {
                if (ch.code == CXXDES_B200_OP_CHILD_PROC) eng.bind(ch.a, nodes[id].prio);   // await_bind of a coroutine child
                nodes[id].kids.push_back(~(int)pos);
                ++pos;
            }
        }
        return id;
    }

    bool ready_of(int id) {   // await_ready (:50-64)
        comp_node &n = nodes[id];
        n.remaining = n.total;
        n.kid_ready.clear();
        for (int k : n.kids) {
            bool r;
            if (k >= 0) r = ready_of(k);
            else {
                const cxxdes_b200_op &ch = prg.ops[~k];
                r = ch.code == CXXDES_B200_OP_CHILD_PROC ? (bool)eng.procs[ch.a].complete
                    : ch.code == CXXDES_B200_OP_CHILD_UNTIL ? eng.now >= ch.imm : false;
            }
            nodes[id].kid_ready.push_back(r);
            if (r) --nodes[id].remaining;
        }
        comp_node &m = nodes[id];
        m.ready = m.is_all ? m.remaining == 0 : m.remaining < m.total;
        return m.ready;
    }

    void suspend(int id, int32_t self, int32_t out_handler) {   // await_suspend (:66-84)
        // the output token {return latency, priority, self}; its handler is the enclosing composition's (:79-80)
        const int32_t h = eng.add_handler(nodes[id].is_all, nodes[id].total, nodes[id].remaining,
                                          token{nodes[id].latency, nodes[id].prio, self, out_handler});
        for (size_t c = 0; c < nodes[id].kids.size(); ++c) {
            if (nodes[id].kid_ready[c]) continue;
            const int k = nodes[id].kids[c];
            if (k >= 0) {
                suspend(k, self, h);
                continue;
            }
            const cxxdes_b200_op &ch = prg.ops[~k];
            if (ch.code == CXXDES_B200_OP_CHILD_PROC) await_completion(ch.a, self, h);
            else {
                int64_t pr = prio_of(ch);
                if (pr == PRIO_INHERIT) pr = nodes[id].prio;
                if (ch.code == CXXDES_B200_OP_CHILD_EV_WAIT)   // wait_awaitable::await_suspend; the token gets the handler
                    e_event[ch.a].waiters.push_back(token{ch.imm, pr, self, h});
                else
                    eng.schedule(token{ch.code == CXXDES_B200_OP_CHILD_DELAY ? eng.now + ch.imm : ch.imm, pr, self, h});
            }
        }
    }

    // returns true when the awaiting process suspends; `pos` ends behind the last child
    bool composition(int32_t self, const cxxdes_b200_op &head, uint32_t &pos) {
        nodes.clear();
        const int root = parse(head, head.code == CXXDES_B200_OP_ALL_OF, eng.procs[self].priority, pos);
        if (ready_of(root)) return false;
        suspend(root, self, NO_HANDLER);
        return true;
    }

    void resume(int32_t p) {
        runtime &r = rt[p];
        process &me = eng.procs[p];
        for (;;) {
            const cxxdes_b200_op &op = prg.ops[r.pc];
            switch (op.code) {
                case CXXDES_B200_OP_RETURN:
                    eng.do_return(p);
                    return;
                case CXXDES_B200_OP_DELAY:
                    eng.timeout(p, op.imm, prio_of(op));
                    ++r.pc;
                    return;
                case CXXDES_B200_OP_UNTIL: {
                    ++r.pc;
                    if (eng.now >= op.imm) break;
                    int64_t pr = prio_of(op);
                    if (pr == PRIO_INHERIT) pr = me.priority;
                    eng.schedule(token{op.imm, pr, p, NO_HANDLER});
                    return;
                }
                case CXXDES_B200_OP_DELAY_EXP: {
                    addr.stream = prg.procs[p].ordinal;
                    double v = cxxdes_b200::exponential(addr, (uint64_t)r.draws++, cxxdes_b200::make_divisor(prg.rates[rate_row * prg.n_rates + op.a]),
                                                        cxxdes_b200::LOG_TABLE_HOST);
                    eng.timeout(p, cxxdes_b200::real_to_sim(v, clk), prio_of(op));
                    ++r.pc;
                    return;
                }
                case CXXDES_B200_OP_DELAY_REAL:
                    eng.timeout(p, cxxdes_b200::real_to_sim(cxxdes_b200::bits_as_double((uint64_t)op.imm), clk), prio_of(op));
                    ++r.pc;
                    return;
                case CXXDES_B200_OP_FRESH:
                    fresh(op.a);
                    ++r.pc;
                    break;
                case CXXDES_B200_OP_AWAIT:
                    eng.bind(op.a, me.priority);
                    ++r.pc;
                    if (eng.procs[op.a].complete) break;
                    await_completion(op.a, p, NO_HANDLER);
                    return;
                case CXXDES_B200_OP_ASYNC:
                    eng.bind(op.a, me.priority);
                    if (!eng.procs[op.a].complete) await_completion(op.a, NO_TARGET, NO_HANDLER);
                    ++r.pc;
                    break;
                case CXXDES_B200_OP_ALL_OF:
                case CXXDES_B200_OP_ANY_OF: {
                    uint32_t pos = r.pc + 1;
                    const bool suspended = composition(p, op, pos);
                    rt[p].pc = pos;
                    if (suspended) return;
                    break;
                }
                case CXXDES_B200_OP_Q_PUT: {
                    uint64_t max_size = prg.queue_max[op.a];
                    if (max_size != 0 && queues[op.a].size() >= max_size) {
                        eng.wait(q_event[op.a], p);
                        return;
                    }
                    queues[op.a].push_back(op.b ? r.reg : eng.now);
                    eng.wake(q_event[op.a]);
                    ++r.pc;
                    break;
                }
                case CXXDES_B200_OP_Q_POP:
                    if (queues[op.a].empty()) {
                        eng.wait(q_event[op.a], p);
                        return;
                    }
                    r.reg = queues[op.a].front();
                    queues[op.a].pop_front();
                    eng.wake(q_event[op.a]);
                    ++r.pc;
                    break;
                case CXXDES_B200_OP_SEM_DOWN:
                    if (sem_value[op.a] == 0) {
                        eng.wait(s_event[op.a], p);
                        return;
                    }
                    --sem_value[op.a];
                    eng.wake(s_event[op.a]);
                    ++r.pc;
                    break;
                case CXXDES_B200_OP_SEM_UP:
                    if (sem_value[op.a] >= prg.sem_max[op.a]) {
                        eng.wait(s_event[op.a], p);
                        return;
                    }
                    ++sem_value[op.a];
                    eng.wake(s_event[op.a]);
                    ++r.pc;
                    break;
                case CXXDES_B200_OP_MTX_ACQUIRE:
                    if (mtx_owned[op.a]) {
                        eng.wait(m_event[op.a], p);
                        return;
                    }
                    mtx_owned[op.a] = 1;
                    ++r.pc;
                    break;
                case CXXDES_B200_OP_MTX_RELEASE:
                    mtx_owned[op.a] = 0;
                    eng.wake(m_event[op.a]);
                    ++r.pc;
                    break;
                case CXXDES_B200_OP_EV_WAIT:
                    eng.wait(e_event[op.a], p, op.imm, prio_of(op));
                    ++r.pc;
                    return;
                case CXXDES_B200_OP_EV_WAKE:
                    eng.wake(e_event[op.a]);
                    ++r.pc;
                    break;
                case CXXDES_B200_OP_LOOP:
                    if (op.imm <= 0) r.pc = (uint32_t)op.b;
                    else {
                        r.loops.push_back((uint32_t)op.imm);
                        ++r.pc;
                    }
                    break;
                case CXXDES_B200_OP_END_LOOP:
                    if (--r.loops.back() > 0) r.pc = (uint32_t)op.b;
                    else {
                        r.loops.pop_back();
                        ++r.pc;
                    }
                    break;
                case CXXDES_B200_OP_SET_REG_NOW:
                    r.reg = eng.now;
                    ++r.pc;
                    break;
                case CXXDES_B200_OP_LAZY_TIMEOUT:   // lazy_timeout: await_bind fixes tsim = now + t, await_ready is
                    r.reg = eng.now + op.imm;       // true (timeout.ipp:106-163): no token
                    ++r.pc;
                    break;
                case CXXDES_B200_OP_UNTIL_REG: {    // co_await instant(tsim, priority): timeout_base, timeout.ipp:14-32
                    ++r.pc;
                    if (eng.now >= r.reg) break;
                    int64_t pr = prio_of(op);
                    if (pr == PRIO_INHERIT) pr = me.priority;
                    eng.schedule(token{r.reg, pr, p, NO_HANDLER});
                    return;
                }
                case CXXDES_B200_OP_STAT_SECONDS:
                    f64[op.a] += cxxdes_b200::sim_to_seconds(eng.now, clk) - cxxdes_b200::sim_to_seconds(r.reg, clk);
                    ++r.pc;
                    break;
                case CXXDES_B200_OP_STAT_TICKS:
                    f64[op.a] += (double)(eng.now - r.reg);
                    ++r.pc;
                    break;
                case CXXDES_B200_OP_COUNT:
                    i64[op.a] += op.imm;
                    ++r.pc;
                    break;
                case CXXDES_B200_OP_IF: {   // plain C++ between the awaits: no token
                    const uint32_t idx = op.a & 0xFFu, src = (op.a >> 8) & 0xFu, cmp = (op.a >> 12) & 0x7u;
                    int64_t v = 0;
                    switch (src) {
                        case CXXDES_B200_SRC_COUNTER: v = i64[idx]; break;
                        case CXXDES_B200_SRC_QUEUE_SIZE: v = (int64_t)queues[idx].size(); break;
                        case CXXDES_B200_SRC_SEM_VALUE: v = (int64_t)sem_value[idx]; break;
                        case CXXDES_B200_SRC_NOW: v = eng.now; break;
                        case CXXDES_B200_SRC_REGISTER: v = r.reg; break;
                        default: v = mtx_owned[idx] ? 1 : 0; break;
                    }
                    const bool t = cmp == CXXDES_B200_CMP_EQ ? v == op.imm : cmp == CXXDES_B200_CMP_NE ? v != op.imm
                                 : cmp == CXXDES_B200_CMP_LT ? v < op.imm : cmp == CXXDES_B200_CMP_LE ? v <= op.imm
                                 : cmp == CXXDES_B200_CMP_GT ? v > op.imm : v >= op.imm;
                    r.pc = t ? r.pc + 1 : (uint32_t)op.b;
                    break;
                }
                case CXXDES_B200_OP_JUMP:
                    r.pc = (uint32_t)op.b;
                    break;
                case CXXDES_B200_OP_MARK:
                    ++i64[0];
                    i64[1] = (int64_t)((uint64_t)i64[1] * 1000003ULL + (uint64_t)(eng.now * op.imm + (int64_t)op.b));
                    ++r.pc;
                    break;
                default:
                    throw std::runtime_error("program_interp: unknown op");
            }
        }
    }

    rep_result result() const {
        rep_result r;
        for (int k = 0; k < CXXDES_B200_N_F64; ++k) r.f64[k] = f64[k];
        for (int k = 0; k < CXXDES_B200_N_I64; ++k) r.i64[k] = i64[k];
        r.status = status;
        return r;
    }
};

}  // namespace port_oracle
