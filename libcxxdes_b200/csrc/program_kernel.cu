// program_kernel.cu -- generic lowering: one kernel interprets a process program
// (cxxdes_b200_program, include/cxxdes_b200.h) for every replication.
//
// This is the device counterpart of the reference's process runtime for models that have no
// hand-written kernel: each operation reproduces the token traffic of one reference awaitable
//  on top of the same engine the model kernels use (device/engine.cuh: libstdc++-exact
// heap, clock, trace digest).  Thread per replication; the event heap is in shared memory, the
// (small) process table, sync objects and handler pool are per-thread local arrays.
#include <new>
#include <cstdlib>
#include <type_traits>

#include "device/models_common.cuh"
#include "device/program_tags.cuh"
#include "device/trace.cuh"
#include "kernels.h"

namespace cxb {

namespace {

constexpr int PRG_THREADS = 128;
constexpr int PRG_HEAP = 80;
constexpr int PRG_MAX_PROCS = 72;   // examples/aloha.cpp as a program: co_main + 64 stations
constexpr int PRG_MAX_NODES = 8;  // compositions in one co_await (the outer one and those nested in it)
constexpr int PRG_HANDLERS = 24;  // live any_of/all_of handlers per replication (pool, by serial); the shared-memory first pass holds 8
constexpr int PRG_MAX_OBJ = 4;    // queues / semaphores / mutexes / events of each kind (the per-thread arrays of this kernel)
constexpr int PRG_DEEP_MAX_PROCS = 254;  // what the 8 bits of a token's tag name (device/program_tags.cuh); the last pass keeps the table in global memory
constexpr int PRG_DEEP_MAX_OBJ = 16;   // ... of its last pass, whose lists and rings live in global memory: what a program may have
constexpr int PRG_WAITERS = 16;
constexpr int PRG_QUEUE_CAP = 128;
constexpr int64_t INHERIT64 = INT64_MAX;

template <int AWAITERS>
struct proc_rt_n {
    int64_t prio;
    int64_t reg;
    uint32_t pc;
    uint32_t draws;
    uint32_t loop[2];
    uint32_t ctag[AWAITERS];      // completion tokens of the processes that await this one (coroutine.ipp:194-207)
    int64_t cprio[AWAITERS];
    int64_t clat[AWAITERS];
    uint8_t loop_depth;
    uint8_t n_comp;
    bool bound;
    bool complete;
};
constexpr int PRG_AWAITERS = 2;        // as coroutine_data.ipp:181 reserves; the last pass holds 8
constexpr int PRG_DEEP_AWAITERS = 8;
constexpr int PRG_DEEP_HANDLERS = 96;  // armed compositions of the last pass (24 above, 8 on the shared-memory pass)

// DEEP (the last pass, below): the same lists and rings in global memory, with the capacities of program_deep.
template <bool DEEP>
struct waiters_rt {  // sync::event: FIFO of {latency, priority, process}  (event.hpp:87-139)
    uint32_t who[PRG_WAITERS];
    uint32_t hser[PRG_WAITERS];   // handler serial + 1 when the wait is a child of any_of / all_of, else 0
    int32_t lat[PRG_WAITERS];
    int16_t prio[PRG_WAITERS];
    int n;
    __device__ __forceinline__ int capacity() const { return PRG_WAITERS; }
};
template <>
struct waiters_rt<true> {
    uint32_t *who, *hser;
    int32_t *lat;
    int16_t *prio;
    int n, cap;
    __device__ __forceinline__ int capacity() const { return cap; }
};

struct handler_rt {  // any_all_helper::custom_handler (any_of.ipp:3-27)
    int64_t out_latency;     // .return_latency(): added to the satisfying child token's time
    uint32_t out_serial_plus1;   // a nested composition: its output token goes to the parent's handler (0: resumes the process)
    uint32_t serial_plus1;
    uint32_t total, remaining;
    bool is_all, armed;
};

template <bool DEEP>
struct queue_rt {
    int64_t item[PRG_QUEUE_CAP];
    uint32_t head, count;
    __device__ __forceinline__ uint32_t capacity() const { return PRG_QUEUE_CAP; }
};
template <>
struct queue_rt<true> {
    int64_t *item;
    uint32_t head, count, cap;
    __device__ __forceinline__ uint32_t capacity() const { return cap; }
};

__device__ __forceinline__ bool branch_taken(uint32_t cmp, int64_t v, int64_t imm) {
    return cmp == CXXDES_B200_CMP_EQ ? v == imm : cmp == CXXDES_B200_CMP_NE ? v != imm : cmp == CXXDES_B200_CMP_LT ? v < imm
         : cmp == CXXDES_B200_CMP_LE ? v <= imm : cmp == CXXDES_B200_CMP_GT ? v > imm : v >= imm;
}

template <bool DEEP>
struct interp {
    static constexpr int MAX_OBJ = DEEP ? PRG_DEEP_MAX_OBJ : PRG_MAX_OBJ;
    static constexpr int AWAITERS = DEEP ? PRG_DEEP_AWAITERS : PRG_AWAITERS;
    static constexpr int HANDLERS = DEEP ? PRG_DEEP_HANDLERS : PRG_HANDLERS;
    using proc_rt = proc_rt_n<AWAITERS>;
    environment env;
    const device_program &prg;
    const shard &sh;
    const log_entry *table;
    std::conditional_t<DEEP, proc_rt *, proc_rt[PRG_MAX_PROCS]> procs;
    waiters_rt<DEEP> q_wait[MAX_OBJ], s_wait[MAX_OBJ], m_wait[MAX_OBJ], e_wait[MAX_OBJ];
    queue_rt<DEEP> queues[MAX_OBJ];
    uint64_t sem_value[MAX_OBJ];
    bool mtx_owned[MAX_OBJ];
    handler_rt handlers[HANDLERS];
    uint32_t next_serial;
    double f64[CXXDES_B200_N_F64];
    int64_t i64[CXXDES_B200_N_I64];
    stream_addr addr;
    uint32_t rate_base;   // parameter sweep: first rate of this replication's row

    __device__ interp(const device_program &p, const shard &s, const log_entry *t) : prg(p), sh(s), table(t) {}

    // DEEP: wait lists and queue rings of this thread in its slab of global memory (program_deep_plan lays it out the
    // same way): only the objects the program has get storage
    __device__ void attach(unsigned char *g, const program_deep &d) {
        if constexpr (DEEP) {
            procs = (proc_rt *)g;
            g += sizeof(proc_rt) * (size_t)prg.n_procs;
            auto lists = [&](waiters_rt<true> *w, uint32_t count) {
                for (uint32_t o = 0; o < (uint32_t)MAX_OBJ; ++o) {
                    const size_t W = o < count ? (size_t)d.waiter_cap : 0;
                    w[o].who = (uint32_t *)g;   g += 4 * W;
                    w[o].hser = (uint32_t *)g;  g += 4 * W;
                    w[o].lat = (int32_t *)g;    g += 4 * W;
                    w[o].prio = (int16_t *)g;   g += 2 * W;
                    w[o].cap = (int)W;
                }
            };
            lists(q_wait, prg.n_queues);
            lists(s_wait, prg.n_sems);
            lists(m_wait, prg.n_mutexes);
            lists(e_wait, prg.n_events);
            for (uint32_t o = 0; o < (uint32_t)MAX_OBJ; ++o) {
                queues[o].item = (int64_t *)g;
                queues[o].cap = o < prg.n_queues ? (uint32_t)d.queue_cap : 0u;
                g += 8 * (size_t)queues[o].cap;
            }
        }
    }

    __device__ void fresh(uint32_t p) {
        proc_rt &r = procs[p];
        r.prio = prg.procs[p].priority;
        r.pc = prg.procs[p].entry;
        r.loop_depth = 0;
        r.n_comp = 0;
        r.bound = false;
        r.complete = false;
    }

    // coroutine_data::bind_, environment.ipp:281-307
    __device__ void bind(uint32_t p, int64_t inherited) {
        proc_rt &r = procs[p];
        if (r.bound) return;
        r.bound = true;
        if (r.prio == INHERIT64) r.prio = inherited;
        env.schedule(env.now + prg.procs[p].latency, r.prio, ptag(p));
    }

    // coroutine::await_suspend, coroutine.ipp:194-207
    __device__ void register_completion(uint32_t child, uint32_t tag) {
        proc_rt &c = procs[child];
        if (c.n_comp >= AWAITERS) {
            env.status |= CXXDES_B200_REP_WAITER_OVERFLOW;
            return;
        }
        const int k = c.n_comp++;
        c.ctag[k] = tag;
        const int64_t rp = prg.procs[child].return_priority;
        c.cprio[k] = rp == INHERIT64 ? c.prio : rp;
        c.clat[k] = prg.procs[child].return_latency;
    }

    // coroutine_data_::do_return, environment.ipp:321-348
    __device__ void do_return(uint32_t p) {
        proc_rt &r = procs[p];
        for (int k = 0; k < r.n_comp; ++k) env.schedule(r.clat[k] + env.now, r.cprio[k], r.ctag[k]);
        r.n_comp = 0;
        r.complete = true;
    }

    __device__ void wait_on(waiters_rt<DEEP> &w, uint32_t p, int64_t latency, int64_t prio, uint32_t handler_serial_plus1 = 0) {
        if (w.n >= w.capacity() || prio < -32768 || prio > 32767 || latency < 0 || latency > INT32_MAX) {
            env.status |= CXXDES_B200_REP_WAITER_OVERFLOW;
            return;
        }
        w.who[w.n] = p;
        w.hser[w.n] = handler_serial_plus1;
        w.lat[w.n] = (int32_t)latency;
        w.prio[w.n] = (int16_t)prio;
        ++w.n;
    }

    // wake_awaitable::await_ready, event.hpp:125-134
    __device__ void wake(waiters_rt<DEEP> &w) {
        for (int k = 0; k < w.n; ++k) env.schedule(env.now + w.lat[k], (int64_t)w.prio[k], ptag(w.who[k], w.hser[k]));
        w.n = 0;
    }

    __device__ int64_t op_priority(const cxxdes_b200_op &op, const proc_rt &self) const {
        return op.b == CXXDES_B200_INHERIT ? self.prio : (int64_t)op.b;
    }

    // any_of / all_of, possibly nested: any_of.ipp:41-92.  The children of the composition at `head` follow it in
    // pre-order; a nested composition is a child that has children of its own.  Returns true when the awaiting
    // process suspends; `span` = number of child operations (so the caller can step over them).
    __device__ bool composition(uint32_t self_id, const cxxdes_b200_op &head, uint32_t first_child, uint32_t &span) {
        proc_rt &self = procs[self_id];
        struct node_rt {
            int64_t prio;            // what the node's children inherit
            uint32_t serial_plus1;   // its handler, once suspended
            int16_t parent;
            uint8_t total, n_ready, left;
            bool is_all, ready, suspended;
        } nodes[PRG_MAX_NODES];
        int n_nodes = 1;
        nodes[0].prio = head.b == CXXDES_B200_INHERIT ? self.prio : (int64_t)head.b;
        nodes[0].parent = -1;
        nodes[0].total = nodes[0].left = (uint8_t)head.a;
        nodes[0].n_ready = 0;
        nodes[0].is_all = head.code == CXXDES_B200_OP_ALL_OF;
        // 1. await_bind + the ready flags of the leaves, children in argument order (pre-order walk)
        uint64_t leaf_ready = 0;
        uint32_t pos = first_child;
        for (int cur = 0;;) {
            while (cur >= 0 && nodes[cur].left == 0) cur = nodes[cur].parent;
            if (cur < 0) break;
            --nodes[cur].left;
            const cxxdes_b200_op &ch = prg.ops[pos];
            if (ch.code == CXXDES_B200_OP_CHILD_ALL_OF || ch.code == CXXDES_B200_OP_CHILD_ANY_OF) {
                if (n_nodes >= PRG_MAX_NODES) {
                    env.status |= CXXDES_B200_REP_BAD_PROGRAM;
                    span = 0;
                    return true;
                }
                node_rt &nd = nodes[n_nodes];
                nd.prio = ch.b == CXXDES_B200_INHERIT ? nodes[cur].prio : (int64_t)ch.b;   // priority_ = inherited
                nd.parent = (int16_t)cur;
                nd.total = nd.left = (uint8_t)ch.a;
                nd.n_ready = 0;
                nd.is_all = ch.code == CXXDES_B200_OP_CHILD_ALL_OF;
                cur = n_nodes++;
            } else {
                bool ready = false;
                if (ch.code == CXXDES_B200_OP_CHILD_PROC) {
                    bind(ch.a, nodes[cur].prio);
                    ready = procs[ch.a].complete;
                } else if (ch.code == CXXDES_B200_OP_CHILD_UNTIL) {
                    ready = env.now >= ch.imm;
                }
                if (ready) {
                    leaf_ready |= 1ULL << ((pos - first_child) & 63u);
                    ++nodes[cur].n_ready;
                }
            }
            ++pos;
        }
        span = pos - first_child;
        // 2. await_ready of every composition, innermost first (children have larger numbers than their parent)
        for (int i = n_nodes - 1; i >= 0; --i) {
            const uint32_t remaining = (uint32_t)nodes[i].total - nodes[i].n_ready;
            nodes[i].ready = nodes[i].is_all ? remaining == 0 : remaining < nodes[i].total;
            if (nodes[i].ready && nodes[i].parent >= 0) ++nodes[nodes[i].parent].n_ready;
        }
        if (nodes[0].ready) return false;
        // 3. await_suspend, outermost first: a handler per non-ready composition (its output token goes to the
        //    parent's handler, or resumes the process), then its non-ready children in order
        auto make_handler = [&](int i, const cxxdes_b200_op &op) -> bool {
            const uint32_t serial = next_serial++;
            if (serial + 1 > PTAG_MAX_SERIAL) {
                env.status |= CXXDES_B200_REP_BAD_PROGRAM;
                return false;
            }
            // storage: any handler that has fired (its late tokens are recognised by serial, device/program_tags.cuh)
            int slot = -1;
            for (int k = 0; k < HANDLERS; ++k)
                if (!handlers[k].armed) {
                    slot = k;
                    break;
                }
            if (slot < 0) {   // more live compositions than the pool holds
                env.status |= CXXDES_B200_REP_WAITER_OVERFLOW;
                return false;
            }
            handler_rt &h = handlers[slot];
            h.serial_plus1 = serial + 1;
            h.total = nodes[i].total;
            h.remaining = (uint32_t)nodes[i].total - nodes[i].n_ready;
            h.is_all = nodes[i].is_all;
            h.armed = true;
            h.out_latency = op.imm;
            h.out_serial_plus1 = nodes[i].parent >= 0 ? nodes[nodes[i].parent].serial_plus1 : 0u;
            nodes[i].serial_plus1 = serial + 1;
            nodes[i].suspended = true;
            return true;
        };
        if (!make_handler(0, head)) return true;
        nodes[0].left = nodes[0].total;
        pos = first_child;
        int next_node = 1;
        for (int cur = 0;;) {
            while (cur >= 0 && nodes[cur].left == 0) cur = nodes[cur].parent;
            if (cur < 0) break;
            --nodes[cur].left;
            const cxxdes_b200_op &ch = prg.ops[pos];
            if (ch.code == CXXDES_B200_OP_CHILD_ALL_OF || ch.code == CXXDES_B200_OP_CHILD_ANY_OF) {
                const int id = next_node++;
                nodes[id].left = nodes[id].total;
                nodes[id].suspended = false;
                if (nodes[cur].suspended && !nodes[id].ready && !make_handler(id, ch)) return true;
                cur = id;
            } else if (nodes[cur].suspended && !(leaf_ready & (1ULL << ((pos - first_child) & 63u)))) {
                const uint32_t tag = ptag(self_id, nodes[cur].serial_plus1);
                const int64_t pr = ch.b == CXXDES_B200_INHERIT ? nodes[cur].prio : (int64_t)ch.b;
                if (ch.code == CXXDES_B200_OP_CHILD_PROC) register_completion(ch.a, tag);
                else if (ch.code == CXXDES_B200_OP_CHILD_DELAY) env.schedule(env.now + ch.imm, pr, tag);
                else if (ch.code == CXXDES_B200_OP_CHILD_UNTIL) env.schedule(ch.imm, pr, tag);
                else if (ch.code == CXXDES_B200_OP_CHILD_EV_WAIT)   // the waiter token carries the handler (event.hpp:136-139)
                    wait_on(e_wait[ch.a], self_id, ch.imm, pr, nodes[cur].serial_plus1);
                else env.status |= CXXDES_B200_REP_BAD_PROGRAM;
            }
            ++pos;
        }
        return true;
    }

    // custom_handler::invoke, any_of.ipp:9-26
    __device__ void invoke_handler(const token &t) {
        const uint32_t hp1 = ptag_serial_plus1(t.tag);
        int slot = -1;
        for (int k = 0; k < HANDLERS; ++k)
            if (handlers[k].serial_plus1 == hp1) slot = k;
        if (slot < 0) return;  // the handler already fired and its storage was reused: nothing left to do
        handler_rt &h = handlers[slot];
        --h.remaining;
        const bool cond = h.is_all ? h.remaining == 0 : h.remaining < h.total;
        if (h.armed && cond) {
            env.schedule(h.out_latency + key_time(t.key), key_prio(t.key), ptag(ptag_target(t.tag), h.out_serial_plus1));
            h.armed = false;
        }
    }

    // Resume process p and run until it suspends or returns (coroutine_data::resume, coroutine_data.ipp:20-29).
    __device__ void resume(uint32_t p) {
        proc_rt &self = procs[p];
        for (int guard = 0; guard < 100000; ++guard) {
            if (self.pc >= prg.n_ops) {
                env.status |= CXXDES_B200_REP_BAD_PROGRAM;
                return;
            }
            const cxxdes_b200_op op = prg.ops[self.pc];
            switch (op.code) {
                case CXXDES_B200_OP_RETURN:
                    do_return(p);
                    return;
                case CXXDES_B200_OP_DELAY:
                    env.schedule(env.now + op.imm, op_priority(op, self), ptag(p));
                    ++self.pc;
                    return;
                case CXXDES_B200_OP_UNTIL:
                    ++self.pc;
                    if (env.now >= op.imm) break;
                    env.schedule(op.imm, op_priority(op, self), ptag(p));
                    return;
                case CXXDES_B200_OP_DELAY_EXP: {
                    addr.stream = prg.procs[p].ordinal;
                    const double v = exponential(addr, (uint64_t)self.draws++, prg.rates[rate_base + op.a], table);
                    env.schedule(env.now + real_to_sim(v, sh.clk), op_priority(op, self), ptag(p));
                    ++self.pc;
                    return;
                }
                case CXXDES_B200_OP_DELAY_REAL:
                    env.schedule(env.now + real_to_sim(bits_as_double((uint64_t)op.imm), sh.clk), op_priority(op, self),
                                 ptag(p));
                    ++self.pc;
                    return;
                case CXXDES_B200_OP_FRESH:
                    fresh(op.a);
                    ++self.pc;
                    break;
                case CXXDES_B200_OP_AWAIT:
                    bind(op.a, self.prio);
                    ++self.pc;
                    if (procs[op.a].complete) break;
                    register_completion(op.a, ptag(p));
                    return;
                case CXXDES_B200_OP_ASYNC:
                    bind(op.a, self.prio);
                    if (!procs[op.a].complete) register_completion(op.a, ptag(PTAG_NONE));
                    ++self.pc;
                    break;
                case CXXDES_B200_OP_ALL_OF:
                case CXXDES_B200_OP_ANY_OF: {
                    const uint32_t first_child = self.pc + 1;
                    uint32_t span = 0;
                    const bool suspended = composition(p, op, first_child, span);
                    self.pc = first_child + span;
                    if (suspended) return;
                    break;
                }
                case CXXDES_B200_OP_Q_PUT: {
                    queue_rt<DEEP> &q = queues[op.a];
                    const uint64_t max_size = prg.queue_max[op.a];
                    if (max_size != 0 && q.count >= max_size) {  // while (!can_put()) co_await event_.wait();
                        wait_on(q_wait[op.a], p, 0, self.prio);
                        return;
                    }
                    if (q.count >= q.capacity()) {
                        env.status |= CXXDES_B200_REP_QUEUE_OVERFLOW;
                        return;
                    }
                    q.item[(q.head + q.count) % q.capacity()] = op.b ? self.reg : env.now;
                    ++q.count;
                    wake(q_wait[op.a]);
                    ++self.pc;
                    break;
                }
                case CXXDES_B200_OP_Q_POP: {
                    queue_rt<DEEP> &q = queues[op.a];
                    if (q.count == 0) {
                        wait_on(q_wait[op.a], p, 0, self.prio);
                        return;
                    }
                    self.reg = q.item[q.head];
                    q.head = (q.head + 1) % q.capacity();
                    if (--q.count == 0) q.head = 0;   // an empty ring restarts at slot 0: short queues stay in one cache line
                    wake(q_wait[op.a]);
                    ++self.pc;
                    break;
                }
                case CXXDES_B200_OP_SEM_DOWN:
                    if (sem_value[op.a] == 0) {
                        wait_on(s_wait[op.a], p, 0, self.prio);
                        return;
                    }
                    --sem_value[op.a];
                    wake(s_wait[op.a]);
                    ++self.pc;
                    break;
                case CXXDES_B200_OP_SEM_UP:
                    if (sem_value[op.a] >= prg.sem_max[op.a]) {
                        wait_on(s_wait[op.a], p, 0, self.prio);
                        return;
                    }
                    ++sem_value[op.a];
                    wake(s_wait[op.a]);
                    ++self.pc;
                    break;
                case CXXDES_B200_OP_MTX_ACQUIRE:
                    if (mtx_owned[op.a]) {
                        wait_on(m_wait[op.a], p, 0, self.prio);
                        return;
                    }
                    mtx_owned[op.a] = true;
                    ++self.pc;
                    break;
                case CXXDES_B200_OP_MTX_RELEASE:
                    mtx_owned[op.a] = false;
                    wake(m_wait[op.a]);
                    ++self.pc;
                    break;
                case CXXDES_B200_OP_EV_WAIT:
                    wait_on(e_wait[op.a], p, op.imm, op_priority(op, self));
                    ++self.pc;
                    return;
                case CXXDES_B200_OP_EV_WAKE:
                    wake(e_wait[op.a]);
                    ++self.pc;
                    break;
                case CXXDES_B200_OP_LOOP:
                    if (op.imm <= 0) {
                        self.pc = (uint32_t)op.b;
                    } else if (self.loop_depth >= 2) {
                        env.status |= CXXDES_B200_REP_BAD_PROGRAM;
                        return;
                    } else {
                        self.loop[self.loop_depth++] = (uint32_t)op.imm;
                        ++self.pc;
                    }
                    break;
                case CXXDES_B200_OP_END_LOOP:
                    if (self.loop_depth == 0) {   // (program_check pairs the loops; never index below the counters)
                        env.status |= CXXDES_B200_REP_BAD_PROGRAM;
                        return;
                    }
                    if (--self.loop[self.loop_depth - 1] > 0) {
                        self.pc = (uint32_t)op.b;
                    } else {
                        --self.loop_depth;
                        ++self.pc;
                    }
                    break;
                case CXXDES_B200_OP_SET_REG_NOW:
                    self.reg = env.now;
                    ++self.pc;
                    break;
                case CXXDES_B200_OP_LAZY_TIMEOUT:   // timeout.ipp:106-163: the deadline is fixed, nothing is scheduled
                    self.reg = env.now + op.imm;
                    ++self.pc;
                    break;
                case CXXDES_B200_OP_UNTIL_REG:
                    ++self.pc;
                    if (env.now >= self.reg) break;
                    env.schedule(self.reg, op_priority(op, self), ptag(p));
                    return;
                case CXXDES_B200_OP_STAT_SECONDS:
                    f64[op.a] += sim_to_seconds(env.now, sh.clk) - sim_to_seconds(self.reg, sh.clk);
                    ++self.pc;
                    break;
                case CXXDES_B200_OP_STAT_TICKS:
                    f64[op.a] += (double)(env.now - self.reg);
                    ++self.pc;
                    break;
                case CXXDES_B200_OP_COUNT:
                    i64[op.a] += op.imm;
                    ++self.pc;
                    break;
                case CXXDES_B200_OP_IF: {
                    const uint32_t idx = op.a & 0xFFu, src = (op.a >> 8) & 0xFu;
                    int64_t v;
                    if (src == CXXDES_B200_SRC_COUNTER) v = i64[idx & (CXXDES_B200_N_I64 - 1)];
                    else if (src == CXXDES_B200_SRC_QUEUE_SIZE) v = (int64_t)queues[idx & (MAX_OBJ - 1)].count;
                    else if (src == CXXDES_B200_SRC_SEM_VALUE) v = (int64_t)sem_value[idx & (MAX_OBJ - 1)];
                    else if (src == CXXDES_B200_SRC_NOW) v = env.now;
                    else if (src == CXXDES_B200_SRC_REGISTER) v = self.reg;
                    else v = mtx_owned[idx & (MAX_OBJ - 1)] ? 1 : 0;
                    self.pc = branch_taken((op.a >> 12) & 0x7u, v, op.imm) ? self.pc + 1 : (uint32_t)op.b;
                    break;
                }
                case CXXDES_B200_OP_JUMP:
                    self.pc = (uint32_t)op.b;
                    break;
                case CXXDES_B200_OP_MARK:
                    ++i64[0];
                    i64[1] = (int64_t)((uint64_t)i64[1] * 1000003ULL + (uint64_t)(env.now * op.imm + (int64_t)op.b));
                    ++self.pc;
                    break;
                default:
                    env.status |= CXXDES_B200_REP_BAD_PROGRAM;
                    return;
            }
            if (env.status) return;
        }
        env.status |= CXXDES_B200_REP_BAD_PROGRAM;  // a process body that never suspends
    }
};

}  // namespace

// heap_cap: tokens a replication may have scheduled at once.  The first pass sizes it from the
// program (2 per process + 8), which leaves shared memory for up to 1024 resident threads per SM --
// the interpreter waits on its per-thread state in local memory most of the time, so resident warps
// are what buys throughput; a replication that overflows the small heap is redone by a second pass
// with PRG_HEAP entries (list != nullptr: only the replications on the retry list).
// The program text (ops, process table, rates, limits) is staged into shared memory when it fits.
// TRACE: the same kernel writing the event trace of its replications (cxxdes_b200_trace).
// DEEP: the last pass.  The reference's containers grow as the model needs (std::priority_queue, std::vector of waiters,
// std::queue); the arrays above hold 80 tokens, 16 waiters per list and 128 queue items.  A timed wait in a loop --
// `co_await (evt.wait() || timeout(t))` -- leaves a waiter behind every time the timeout wins (any_of does not cancel the
// loser, any_of.ipp:41-92; the waiter stays in the event's list until the next wake), so a program whose wakes are rarer
// than sixteen of its timeouts outgrows them.  Such a replication -- its status carries a capacity flag after the passes
// above -- is run again with heap, wait lists and queue rings in GLOBAL memory (a slab per thread, capacities of
// program_deep_plan).  The pass walks the status column; `list_cursor` is its work counter, `list_count` counts what it took
// (list_cursor == nullptr: the event trace of one replication on the same storage).
template <bool TRACE, bool DEEP = false>
__global__ void __launch_bounds__(PRG_THREADS)
program_kernel(device_program prg_in, shard sh, device_columns out, int heap_cap, uint32_t staged_bytes,
               const uint32_t *list, const unsigned long long *list_count, unsigned long long *list_cursor, trace_sink sink,
               program_deep deep) {
    extern __shared__ __align__(16) unsigned char smem[];
    log_entry *table = (log_entry *)smem;
    unsigned char *cursor = smem + sizeof(log_entry) * (1 << CXB_LOG_TABLE_BITS);
    device_program prg = prg_in;
    if (staged_bytes) {
        // one contiguous blob on the device (c_api.cu upload_program): ops first, 16-byte aligned pieces
        const uint4 *src = (const uint4 *)prg_in.ops;
        uint4 *dst = (uint4 *)cursor;
        for (uint32_t k = threadIdx.x; k < staged_bytes / 16; k += blockDim.x) dst[k] = src[k];
        const unsigned char *base = (const unsigned char *)prg_in.ops;
        auto moved = [&](const void *ptr) { return (const void *)(cursor + ((const unsigned char *)ptr - base)); };
        prg.ops = (const cxxdes_b200_op *)moved(prg_in.ops);
        prg.procs = (const cxxdes_b200_process *)moved(prg_in.procs);
        prg.roots = (const uint32_t *)moved(prg_in.roots);
        prg.queue_max = (const uint64_t *)moved(prg_in.queue_max);
        prg.sem_init = (const uint64_t *)moved(prg_in.sem_init);
        prg.sem_max = (const uint64_t *)moved(prg_in.sem_max);
        prg.rates = (const divisor *)moved(prg_in.rates);
        cursor += staged_bytes;
    }
    stage_log_table(table, sh.log_table);   // ends with __syncthreads(): the staged program is visible too
    lane_array<uint64_t> hkey;
    lane_array<uint32_t> htag;
    [[maybe_unused]] unsigned char *slab = nullptr;
    if constexpr (DEEP) {
        slab = deep.base + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * deep.bytes_per_thread;
        heap_cap = deep.heap_cap;
        hkey.base = (uint64_t *)slab;
        htag.base = (uint32_t *)(slab + 8 * (size_t)heap_cap);
        hkey.stride = htag.stride = 1;
        slab += 12 * (size_t)heap_cap;      // (heap_cap is even: the lists that follow stay 8-byte aligned)
    } else {
        hkey = carve<uint64_t>(cursor, heap_cap);
        htag = carve<uint32_t>(cursor, heap_cap);
    }

    uint64_t rep;
    [[maybe_unused]] unsigned long long chunk_next = 0, chunk_end = 0;
    for (;;) {
        if (list) {
            const unsigned long long idx = atomicAdd(list_cursor, 1ULL);
            if (idx >= *list_count) break;
            rep = list[idx];
        } else if (DEEP && list_cursor) {
            if (chunk_next == chunk_end) {      // 16 replications per claim
                chunk_next = atomicAdd(list_cursor, 16ULL);
                chunk_end = chunk_next + 16;
            }
            if (chunk_next >= sh.n_replications) break;
            rep = chunk_next++;
            if (!(out.status[rep] & (CXXDES_B200_REP_HEAP_OVERFLOW | CXXDES_B200_REP_WAITER_OVERFLOW | CXXDES_B200_REP_QUEUE_OVERFLOW)))
                continue;
            atomicAdd((unsigned long long *)list_count, 1ULL);
        } else if (!steal(sh, rep)) {
            break;
        }
        interp<DEEP> it(prg, sh, table);
        it.attach(slab, deep);
        it.env.init(hkey, htag, heap_cap);
        it.addr.seed = sh.seed;
        it.addr.replication = sh.first_replication + rep;
        it.rate_base = prg.sweep_steps ? (uint32_t)((sh.first_replication + rep) % prg.sweep_steps) * prg.n_rates : 0u;
        it.next_serial = 0;
        for (int k = 0; k < CXXDES_B200_N_F64; ++k) it.f64[k] = 0.0;
        for (int k = 0; k < CXXDES_B200_N_I64; ++k) it.i64[k] = 0;
        for (uint32_t p = 0; p < prg.n_procs; ++p) {
            it.fresh(p);
            it.procs[p].draws = 0;
            it.procs[p].reg = 0;
        }
        for (int o = 0; o < interp<DEEP>::MAX_OBJ; ++o) {
            it.q_wait[o].n = it.s_wait[o].n = it.m_wait[o].n = it.e_wait[o].n = 0;
            it.queues[o].head = it.queues[o].count = 0;
            it.sem_value[o] = (uint32_t)o < prg.n_sems ? prg.sem_init[o] : 0;
            it.mtx_owned[o] = false;
        }
        for (int h = 0; h < interp<DEEP>::HANDLERS; ++h) {
            it.handlers[h].serial_plus1 = 0;
            it.handlers[h].armed = false;
        }
        // environment::bind(coroutine) for every root, coroutine.ipp:352-357
        for (uint32_t r = 0; r < prg.n_roots; ++r) {
            const uint32_t p = prg.roots[r];
            it.bind(p, 0);
            if (!it.procs[p].complete) it.register_completion(p, ptag(PTAG_NONE));
        }

        environment &env = it.env;
        while (env.has_event() && (sh.run_until < 0 || env.next_time() <= sh.run_until)) {
            // environment::step, environment.ipp:117-146 (the digest names processes by ordinal)
            token t = env.pop_top();
            const int64_t time = key_time(t.key);
            env.now = time > env.now ? time : env.now;
            ++env.n_events;
            const uint32_t target = ptag_target(t.tag);
            const uint32_t hp1 = ptag_serial_plus1(t.tag);
            const uint32_t kind = hp1 ? TOKEN_HANDLER : (target != PTAG_NONE ? TOKEN_RESUME : TOKEN_NOOP);
            const uint32_t ordinal = target != PTAG_NONE ? prg.procs[target].ordinal : PROC_NONE;
            env.hash = trace_hash_step(env.hash, time, key_prio(t.key), kind, ordinal);
            if constexpr (TRACE) {
                traced_environment rec;
                rec.sink = sink;
                rec.record(env.n_events - 1, time, key_prio(t.key), kind, ordinal);
            }
            if (hp1) it.invoke_handler(t);
            else if (target != PTAG_NONE) it.resume(target);
            if (env.status) break;
        }
        if (sh.run_until >= 0) {
            if (env.has_event()) env.status |= CXXDES_B200_REP_UNFINISHED;
            env.now = env.now > sh.run_until ? env.now : sh.run_until;
        }
        if (!DEEP && !list && heap_cap < PRG_HEAP && (env.status & CXXDES_B200_REP_HEAP_OVERFLOW)) {
            const unsigned long long at = atomicAdd(sh.retry_count, 1ULL);   // redo with the full-size heap
            sh.retry_list[at] = (uint32_t)rep;
            continue;
        }
        rep_out o;
        o.end_time = env.now;
        o.n_events = env.n_events;
        o.trace_hash = env.hash;
        o.status = env.status;
        for (int k = 0; k < CXXDES_B200_N_F64; ++k) o.f64[k] = it.f64[k];
        for (int k = 0; k < CXXDES_B200_N_I64; ++k) o.i64[k] = it.i64[k];
        store_result(out, rep, o);
    }
}

const char *program_check(const cxxdes_b200_program &p) {
    if (p.n_procs < 1 || p.n_procs > PRG_DEEP_MAX_PROCS) return "program: 1..254 processes supported";
    if (p.n_ops < 1 || p.n_ops > 4096) return "program: 1..4096 ops supported";
    if (p.n_roots < 1 || p.n_roots > p.n_procs) return "program: need at least one root process";
    if (p.n_queues > PRG_DEEP_MAX_OBJ || p.n_sems > PRG_DEEP_MAX_OBJ || p.n_mutexes > PRG_DEEP_MAX_OBJ || p.n_events > PRG_DEEP_MAX_OBJ)
        return "program: at most 16 sync objects of each kind";
    if (p.n_rates > 16) return "program: at most 16 rates";
    if (p.sweep_steps > 4096) return "program: at most 4096 sweep points";
    if (!p.ops || !p.procs || !p.roots) return "program: ops/procs/roots must not be NULL";
    // no path may run off the end of the program: the last body ends with co_return, jump targets lie inside
    if (p.ops[p.n_ops - 1].code != CXXDES_B200_OP_RETURN) return "program: the last op must be RETURN";
    for (uint32_t i = 0; i < p.n_procs; ++i) {
        if (p.procs[i].entry >= p.n_ops) return "program: process entry out of range";
        if (p.procs[i].ordinal >= 0xFFFF) return "program: ordinal must be < 0xFFFF";
    }
    for (uint32_t r = 0; r < p.n_roots; ++r)
        if (p.roots[r] >= p.n_procs) return "program: root out of range";
    // a time the 64-bit keys cannot hold (47 bits) can never be scheduled: refused here, so that `now + t` never
    // overflows either (found by tools/fuzz_program_check_on_host.py under UBSan)
    constexpr int64_t TIME_LIMIT = 1LL << 47;
    for (uint32_t i = 0; i < p.n_procs; ++i)
        if (p.procs[i].latency <= -TIME_LIMIT || p.procs[i].latency >= TIME_LIMIT || p.procs[i].return_latency <= -TIME_LIMIT ||
            p.procs[i].return_latency >= TIME_LIMIT)
            return "program: latencies must be below 2^47 ticks";
    for (uint32_t i = 0; i < p.n_ops; ++i) {
        const cxxdes_b200_op &op = p.ops[i];
        switch (op.code) {
            case CXXDES_B200_OP_DELAY: case CXXDES_B200_OP_UNTIL: case CXXDES_B200_OP_CHILD_DELAY: case CXXDES_B200_OP_CHILD_UNTIL:
            case CXXDES_B200_OP_LAZY_TIMEOUT: case CXXDES_B200_OP_ALL_OF: case CXXDES_B200_OP_ANY_OF: case CXXDES_B200_OP_CHILD_ALL_OF:
            case CXXDES_B200_OP_CHILD_ANY_OF: case CXXDES_B200_OP_EV_WAIT: case CXXDES_B200_OP_CHILD_EV_WAIT:
                if (op.imm <= -TIME_LIMIT || op.imm >= TIME_LIMIT) return "program: times and latencies must be below 2^47 ticks";
                break;
            default: break;
        }
        switch (op.code) {
            case CXXDES_B200_OP_FRESH: case CXXDES_B200_OP_AWAIT: case CXXDES_B200_OP_ASYNC:
            case CXXDES_B200_OP_CHILD_PROC:
                if (op.a >= p.n_procs) return "program: process operand out of range";
                break;
            case CXXDES_B200_OP_ALL_OF: case CXXDES_B200_OP_ANY_OF: {
                // the children follow in pre-order; a nested composition has children of its own
                uint32_t pending = op.a, pos = i + 1, nodes = 1, leaves = 0;
                if (op.a < 1 || op.a > 64) return "program: bad composition";
                while (pending > 0) {
                    if (pos >= p.n_ops) return "program: bad composition";
                    const cxxdes_b200_op &ch = p.ops[pos];
                    --pending;
                    if (ch.code == CXXDES_B200_OP_CHILD_ALL_OF || ch.code == CXXDES_B200_OP_CHILD_ANY_OF) {
                        if (ch.a < 1 || ch.a > 32 || ++nodes > 8) return "program: bad nested composition (at most 8 compositions in one co_await)";
                        pending += ch.a;
                    } else if (ch.code == CXXDES_B200_OP_CHILD_PROC || ch.code == CXXDES_B200_OP_CHILD_DELAY ||
                               ch.code == CXXDES_B200_OP_CHILD_UNTIL || ch.code == CXXDES_B200_OP_CHILD_EV_WAIT) {
                        if (++leaves > 64) return "program: at most 64 awaitables in one co_await";
                        if (ch.code == CXXDES_B200_OP_CHILD_PROC && ch.a >= p.n_procs) return "program: process operand out of range";
                        if (ch.code == CXXDES_B200_OP_CHILD_EV_WAIT && ch.a >= p.n_events) return "program: event operand out of range";
                    } else {
                        return "program: composition children must be CHILD_* ops";
                    }
                    ++pos;
                }
                if (pos >= p.n_ops) return "program: bad composition";
                break;
            }
            case CXXDES_B200_OP_Q_PUT: case CXXDES_B200_OP_Q_POP:
                if (op.a >= p.n_queues) return "program: queue operand out of range";
                break;
            case CXXDES_B200_OP_SEM_DOWN: case CXXDES_B200_OP_SEM_UP:
                if (op.a >= p.n_sems) return "program: semaphore operand out of range";
                break;
            case CXXDES_B200_OP_MTX_ACQUIRE: case CXXDES_B200_OP_MTX_RELEASE:
                if (op.a >= p.n_mutexes) return "program: mutex operand out of range";
                break;
            case CXXDES_B200_OP_EV_WAIT: case CXXDES_B200_OP_EV_WAKE:
                if (op.a >= p.n_events) return "program: event operand out of range";
                break;
            case CXXDES_B200_OP_DELAY_EXP:
                if (op.a >= p.n_rates) return "program: rate operand out of range";
                break;
            case CXXDES_B200_OP_LOOP: case CXXDES_B200_OP_END_LOOP:
                if (op.b < 0 || (uint32_t)op.b >= p.n_ops) return "program: loop target out of range";
                if (op.code == CXXDES_B200_OP_LOOP && (op.imm < 0 || op.imm > 0xFFFFFFFFLL))
                    return "program: LOOP count must be in [0, 2^32)";
                break;
            case CXXDES_B200_OP_STAT_SECONDS: case CXXDES_B200_OP_STAT_TICKS:
                if (op.a >= CXXDES_B200_N_F64) return "program: statistic index out of range";
                break;
            case CXXDES_B200_OP_COUNT:
                if (op.a >= CXXDES_B200_N_I64) return "program: counter index out of range";
                break;
            case CXXDES_B200_OP_IF: {
                const uint32_t idx = op.a & 0xFFu, src = (op.a >> 8) & 0xFu, cmp = (op.a >> 12) & 0x7u;
                if (src > CXXDES_B200_SRC_MUTEX || cmp > CXXDES_B200_CMP_GE || (op.a >> 15)) return "program: bad IF operand";
                if ((src == CXXDES_B200_SRC_COUNTER && idx >= CXXDES_B200_N_I64) ||
                    (src == CXXDES_B200_SRC_QUEUE_SIZE && idx >= p.n_queues) || (src == CXXDES_B200_SRC_SEM_VALUE && idx >= p.n_sems) ||
                    (src == CXXDES_B200_SRC_MUTEX && idx >= p.n_mutexes))
                    return "program: IF operand index out of range";
            }   // fall through: the target is checked like a jump's
            case CXXDES_B200_OP_JUMP:
                if (op.b <= (int64_t)i || (uint32_t)op.b >= p.n_ops) return "program: IF / JUMP target must lie ahead, inside the program";
                break;
            default:
                if (op.code > CXXDES_B200_OP_CHILD_EV_WAIT) return "program: unknown op code";
        }
    }
    // Structure of the loops (`for` in a coroutine body): every END_LOOP closes the innermost open LOOP and the two
    // name each other (LOOP.b = END + 1, END.b = LOOP + 1); at most two deep (the interpreters keep two counters per
    // process); no body starts inside another body's loop; an IF / JUMP stays inside the loop body it is in -- jumping
    // out would leave the loop counter on the process's stack, jumping in would run END_LOOP on an empty one.
    {
        uint32_t open[2];
        int depth = 0;
        int32_t *nest = new (std::nothrow) int32_t[p.n_ops];   // innermost enclosing LOOP of each op, -1 = none
        if (!nest) return "program: out of host memory";
        const char *why = nullptr;
        for (uint32_t i = 0; i < p.n_ops && !why; ++i) {
            const cxxdes_b200_op &op = p.ops[i];
            if (op.code == CXXDES_B200_OP_END_LOOP) {
                if (depth == 0) { why = "program: END_LOOP without a LOOP"; break; }
                const uint32_t head = open[depth - 1];
                if ((uint32_t)p.ops[head].b != i + 1 || (uint32_t)op.b != head + 1) { why = "program: LOOP / END_LOOP do not pair up"; break; }
                nest[i] = (int32_t)head;
                --depth;
                continue;
            }
            nest[i] = depth ? (int32_t)open[depth - 1] : -1;
            if (op.code == CXXDES_B200_OP_LOOP) {
                if (depth >= 2) { why = "program: loops nest at most two deep"; break; }
                open[depth++] = i;
            }
        }
        if (!why && depth != 0) why = "program: LOOP without an END_LOOP";
        for (uint32_t i = 0; i < p.n_procs && !why; ++i)
            if (nest[p.procs[i].entry] != -1) why = "program: a process body starts inside a loop";
        for (uint32_t i = 0; i < p.n_ops && !why; ++i) {
            const cxxdes_b200_op &op = p.ops[i];
            if ((op.code == CXXDES_B200_OP_IF || op.code == CXXDES_B200_OP_JUMP) && nest[i] != nest[op.b])
                why = "program: IF / JUMP must stay inside its loop body";
        }
        delete[] nest;
        if (why) return why;
    }
    return nullptr;
}

namespace {
size_t program_smem_bytes(int heap_cap, uint32_t staged_bytes) {
    return sizeof(log_entry) * (1 << CXB_LOG_TABLE_BITS) + 64 + staged_bytes + (size_t)PRG_THREADS * heap_cap * 12;
}
}  // namespace

bool program_needs_last_pass(const device_program &prg) {
    return prg.n_procs > PRG_MAX_PROCS || prg.n_queues > PRG_MAX_OBJ || prg.n_sems > PRG_MAX_OBJ || prg.n_mutexes > PRG_MAX_OBJ ||
           prg.n_events > PRG_MAX_OBJ;
}

// Capacities of the last pass: 2048 tokens, 1024 waiters per list, 4096 items per queue -- a slab per thread holding the
// heap and the lists / rings of the objects the program has; as many threads as `budget` bytes hold, whole blocks, at
// most one block per SM.  Returns the bytes to allocate (0: no such pass).
size_t program_deep_plan(const device_program &prg, int sm_count, size_t budget, program_deep &d) {
    d = program_deep{};
    const size_t H = 2048, W = 1024, Q = 4096;
    const size_t lists = (size_t)prg.n_queues + prg.n_sems + prg.n_mutexes + prg.n_events;
    const size_t per_thread = 12 * H + sizeof(proc_rt_n<PRG_DEEP_AWAITERS>) * (size_t)prg.n_procs + lists * 14 * W + (size_t)prg.n_queues * 8 * Q;
    size_t threads = budget / per_thread;
    if (threads == 0) return 0;
    if (threads > (size_t)sm_count * PRG_THREADS) threads = (size_t)sm_count * PRG_THREADS;
    if (threads > (size_t)PRG_THREADS) threads -= threads % PRG_THREADS;
    d.threads = (int)threads;
    d.heap_cap = (int)H;
    d.waiter_cap = (int)W;
    d.queue_cap = (int)Q;
    d.bytes_per_thread = per_thread;
    return threads * per_thread;
}

cudaError_t launch_program(const device_program &prg, uint32_t blob_bytes, const shard &sh, const device_columns &out,
                           int sm_count, unsigned long long *list_cursor, cudaStream_t stream, int *launches,
                           const program_fast_layout &fast, int64_t *queue_items, const program_fast_state &fast_state,
                           const program_deep &deep) {
    const uint32_t staged = blob_bytes <= 12 * 1024 ? (blob_bytes + 15u) & ~15u : 0u;
    int small = 2 * (int)prg.n_procs + 8;
    if (small > PRG_HEAP) small = PRG_HEAP;
    auto go = [&](int heap_cap, bool retry) -> cudaError_t {
        const size_t smem = program_smem_bytes(heap_cap, staged);
        cudaError_t e = cudaFuncSetAttribute(program_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int per_sm = (int)((227 * 1024) / (smem + 1024));
        if (per_sm < 1) per_sm = 1;
        // throughput grows linearly with resident warps here (measured 2 -> 8 blocks/SM: 3.2 -> 9.5e9 events/s):
        // take what shared memory and the 40 registers per thread allow (12 blocks of 128 threads)
        if (per_sm > 12) per_sm = 12;
        if (const char *env = getenv("CXXDES_B200_PRG_BLOCKS")) { int v = atoi(env); if (v >= 1 && v < per_sm) per_sm = v; }
        program_kernel<false><<<sm_count * per_sm, PRG_THREADS, smem, stream>>>(
            prg, sh, out, heap_cap, staged, retry ? sh.retry_list : nullptr, retry ? sh.retry_count : nullptr,
            retry ? list_cursor : nullptr, trace_sink{}, program_deep{});
        ++*launches;
        return cudaGetLastError();
    };
    cudaError_t e;
    if (program_needs_last_pass(prg)) {
        // more than 4 sync objects of a kind: this kernel's per-thread arrays cannot hold the program; the shared-memory
        // pass (laid out from the program) runs first where it applies, the last pass takes its retry list or the shard
        if (!deep.base) return cudaErrorInvalidConfiguration;      // (refused at create: c_api.cu)
        const size_t smem = program_smem_bytes(0, staged);
        e = cudaFuncSetAttribute(program_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        const int block = deep.threads < PRG_THREADS ? deep.threads : PRG_THREADS;
        if (fast.valid) {
            e = launch_program_fast(prg, fast, sh, out, queue_items, fast_state, sm_count, stream);
            ++*launches;
            if (e != cudaSuccess) return e;
        }
        program_kernel<false, true><<<deep.threads / block, block, smem, stream>>>(
            prg, sh, out, deep.heap_cap, staged, fast.valid ? sh.retry_list : nullptr, fast.valid ? sh.retry_count : nullptr,
            fast.valid ? list_cursor : nullptr, trace_sink{}, deep);
        ++*launches;
        return cudaGetLastError();
    }
    if (fast.valid) {
        // shared-memory interpreter first (program_fast.cu); what its layout cannot hold comes back on the retry list
        e = launch_program_fast(prg, fast, sh, out, queue_items, fast_state, sm_count, stream);
        ++*launches;
        if (e == cudaSuccess) e = go(PRG_HEAP, true);
    } else {
        e = go(small, false);
        if (e == cudaSuccess && small != PRG_HEAP) e = go(PRG_HEAP, true);   // exits at once when nothing overflowed
    }
    if (e != cudaSuccess || !deep.base) return e;
    // what outgrew the fixed capacities: once more with heap, wait lists and queues in global memory (exits at once
    // when nothing did)
    const size_t smem = program_smem_bytes(0, staged);
    e = cudaFuncSetAttribute(program_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int block = deep.threads < PRG_THREADS ? deep.threads : PRG_THREADS;
    program_kernel<false, true><<<deep.threads / block, block, smem, stream>>>(
        prg, sh, out, deep.heap_cap, staged, nullptr, sh.retry_a_count, sh.retry_a_cursor, trace_sink{}, deep);
    ++*launches;
    return cudaGetLastError();
}

// One replication (sh.n_replications == 1) on the full-size interpreter (full heap; the last pass's storage when the
// handle has it, so that any replication the ensemble's passes can run can be traced), with its event trace.
cudaError_t launch_program_trace(const device_program &prg, uint32_t blob_bytes, const shard &sh, const device_columns &out,
                                 const trace_sink &sink, cudaStream_t stream, const program_deep &deep) {
    const uint32_t staged = blob_bytes <= 12 * 1024 ? (blob_bytes + 15u) & ~15u : 0u;
    if (deep.base) {
        const size_t smem = program_smem_bytes(0, staged);
        cudaError_t e = cudaFuncSetAttribute(program_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        program_kernel<true, true><<<1, deep.threads < PRG_THREADS ? deep.threads : PRG_THREADS, smem, stream>>>(
            prg, sh, out, deep.heap_cap, staged, nullptr, nullptr, nullptr, sink, deep);
        return cudaGetLastError();
    }
    const size_t smem = program_smem_bytes(PRG_HEAP, staged);
    cudaError_t e = cudaFuncSetAttribute(program_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    program_kernel<true><<<1, PRG_THREADS, smem, stream>>>(prg, sh, out, PRG_HEAP, staged, nullptr, nullptr, nullptr, sink, program_deep{});
    return cudaGetLastError();
}

}  // namespace cxb
