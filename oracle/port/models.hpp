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

// ---------------------------------------------------------------------------
// ALOHA -- examples/aloha.cpp:25-85.  Engine process index == ordinal: co_main 0, station s 1+s.
// ---------------------------------------------------------------------------
struct aloha_model {
    engine eng;
    cxxdes_b200::clock_config clk;
    size_t num_stations;
    float lambda;       // aloha_config::lambda (float)
    float frame_time;
    uint64_t pps;
    std::vector<exp_stream> interarrival;
    std::vector<uint64_t> i;          // loop index of each station, aloha.cpp:54
    std::vector<char> collided, active;  // the per-iteration `collided` flag and set membership
    size_t n_active = 0;
    uint64_t successful = 0;
    float s_result = 0, g_result = 0;

    aloha_model(const cxxdes_b200_aloha_params &p, const cxxdes_b200_clock &c, uint64_t seed,
                uint64_t rep)
        : clk(fold_clock(c)), num_stations(p.num_stations), frame_time(p.frame_time) {
        if (p.sweep_steps > 1) {
            // main(), aloha.cpp:96,101: float arithmetic with size_t operands converted to float
            const size_t num_steps = p.sweep_steps;
            const size_t step = rep % num_steps;
            const float begin = p.lambda_begin, end = p.lambda_end;
            lambda = begin + (end - begin) * step / (num_steps - 1);
            pps = (uint64_t)(size_t)(lambda * p.packets_scale);
        } else {
            lambda = p.lambda_begin;
            pps = p.packets_per_station;
        }
        eng.add_process(0);
        for (size_t s = 0; s < num_stations; ++s) {
            eng.add_process((uint32_t)(1 + s));
            // exponential_rv{seed, cfg_.lambda / cfg_.num_stations}: float / size_t -> float -> double (:52)
            interarrival.emplace_back(seed, rep, (uint32_t)(1 + s), (double)(lambda / (float)num_stations));
        }
        i.assign(num_stations, 0);
        collided.assign(num_stations, 0);
        active.assign(num_stations, 0);
        eng.bind_root(0);
    }

    double now_seconds() const { return cxxdes_b200::sim_to_seconds(eng.now, clk); }

    void resume(int32_t p) {
        if (p == 0) resume_main();
        else resume_station((size_t)(p - 1));
    }

    // co_main, aloha.cpp:39-48
    void resume_main() {
        process &me = eng.procs[0];
        if (me.pc == 0) {
            // all_of.range copies the handles and binds them in index order (any_of.ipp:41-48,197-209)
            for (size_t s = 0; s < num_stations; ++s) eng.bind((int32_t)(1 + s), me.priority);
            int32_t h = eng.add_handler(true, num_stations, num_stations, token{0, me.priority, 0, NO_HANDLER});
            for (size_t s = 0; s < num_stations; ++s) eng.await_completion((int32_t)(1 + s), 0, h);
            me.pc = 1;
        } else {
            s_result = (float)((double)successful / now_seconds());  // size_t / double -> float (:46)
            g_result = lambda * frame_time;                           // :47
            eng.do_return(0);
        }
    }

    // station, aloha.cpp:51-71
    void resume_station(size_t s) {
        process &me = eng.procs[1 + s];
        int32_t self = (int32_t)(1 + s);
        if (me.pc == 1) {
            // after co_await env.timeout(frame_time): erase, count, draw the waiting time (:62-69)
            active[s] = 0;
            --n_active;
            if (!collided[s]) ++successful;
            eng.timeout(self, cxxdes_b200::real_to_sim(interarrival[s](), clk));
            me.pc = 2;
            return;
        }
        if (me.pc == 2) ++i[s];
        if (i[s] < pps) {
            collided[s] = 0;  // bool collided = false; insert; check_collisions (:55-59,73-78)
            active[s] = 1;
            ++n_active;
            if (n_active > 1)
                for (size_t k = 0; k < num_stations; ++k)
                    if (active[k]) collided[k] = 1;
            eng.timeout(self, cxxdes_b200::real_to_sim_f32(frame_time, clk));  // :60
            me.pc = 1;
        } else {
            eng.do_return(self);
        }
    }

    rep_result result() const {
        rep_result r;
        r.f64[0] = (double)s_result;
        r.f64[1] = (double)g_result;
        r.i64[0] = (int64_t)successful;
        r.i64[1] = (int64_t)pps;
        return r;
    }
};

// ---------------------------------------------------------------------------
// M/M/c with balking: sync::resource (resource.hpp:37-100 over semaphore.hpp:58-78), async
// (async.ipp:21-28) and any_of (any_of.ipp).  Process index == ordinal:
// co_main 0, source 1, customer i 2+2i, acquire_proc i 3+2i.
// ---------------------------------------------------------------------------
struct mmc_model {
    engine eng;
    cxxdes_b200::clock_config clk;
    cxxdes_b200_mmc_params prm;
    uint64_t seed, rep;
    exp_stream arr;
    uint64_t value, max_value;  // semaphore<>{c, c}
    event sem_event;
    std::vector<char> balked_flag, has_handle;
    std::vector<double> t0;
    std::vector<int32_t> handler_of;
    uint64_t i = 0;             // source loop index
    uint64_t served = 0, balked = 0;
    double total_wait = 0;

    mmc_model(const cxxdes_b200_mmc_params &p, const cxxdes_b200_clock &c, uint64_t seed_,
              uint64_t rep_)
        : clk(fold_clock(c)), prm(p), seed(seed_), rep(rep_), arr(seed_, rep_, 1, p.lambda),
          value(p.servers), max_value(p.servers) {
        eng.add_process(0);
        eng.add_process(1);
        for (uint64_t k = 0; k < p.n_customers; ++k) {
            eng.add_process((uint32_t)(2 + 2 * k));
            eng.add_process((uint32_t)(3 + 2 * k));
        }
        balked_flag.assign(p.n_customers, 0);
        has_handle.assign(p.n_customers, 0);
        t0.assign(p.n_customers, 0.0);
        handler_of.assign(p.n_customers, NO_HANDLER);
        eng.bind_root(0);
    }

    double now_seconds() const { return cxxdes_b200::sim_to_seconds(eng.now, clk); }

    void resume(int32_t p) {
        if (p == 0) resume_main();
        else if (p == 1) resume_source();
        else if ((p & 1) == 0) resume_customer((uint64_t)(p - 2) / 2);
        else resume_acquire((uint64_t)(p - 3) / 2);
    }

    void resume_main() {
        process &me = eng.procs[0];
        if (me.pc == 0) {
            eng.bind(1, me.priority);        // co_await source(): start + completion targeting main
            eng.await_completion(1, 0);
            me.pc = 1;
        } else {
            eng.do_return(0);
        }
    }

    void resume_source() {
        process &me = eng.procs[1];
        if (me.pc == 1) ++i;
        if (i < prm.n_customers) {
            int32_t cust = (int32_t)(2 + 2 * i);
            // co_await async(customer(i)): bind + null-target completion, caller continues (async.ipp:21-28)
            eng.bind(cust, me.priority);
            if (!eng.procs[cust].complete) eng.await_completion(cust, NO_TARGET);
            eng.timeout(1, cxxdes_b200::real_to_sim(arr(), clk));
            me.pc = 1;
        } else {
            eng.do_return(1);
        }
    }

    // semaphore::up / down wake(): event.hpp:125-134
    void sem_up() {
        // while (value >= max) wait;  cannot block here: every up() pairs with an earlier down()
        ++value;
        eng.wake(sem_event);
    }

    void resume_customer(uint64_t id) {
        int32_t self = (int32_t)(2 + 2 * id), acq = (int32_t)(3 + 2 * id);
        process &me = eng.procs[self];
        if (me.pc == 0) {
            t0[id] = now_seconds();
            // any_of(acq, env.timeout(patience)): await_bind binds acq (start token) (any_of.ipp:41-48)
            int64_t patience_ticks = cxxdes_b200::real_to_sim(prm.patience, clk);
            eng.bind(acq, me.priority);
            // await_suspend: output token, handler on the completion token of acq and on the timeout (:66-84)
            int32_t h = eng.add_handler(false, 2, 2, token{0, me.priority, self, NO_HANDLER});
            handler_of[id] = h;
            eng.await_completion(acq, self, h);
            eng.timeout(self, patience_ticks, PRIO_INHERIT, h);
            me.pc = 1;
        } else if (me.pc == 1) {
            if (!eng.procs[acq].complete) {
                balked_flag[id] = 1;
                ++balked;
                eng.do_return(self);
                return;
            }
            total_wait += now_seconds() - t0[id];
            exp_stream svc(seed, rep, (uint32_t)(2 + 2 * id), prm.mu);
            eng.timeout(self, cxxdes_b200::real_to_sim(svc(), clk));
            me.pc = 2;
        } else {
            sem_up();  // co_await st[id].h.release()
            ++served;
            eng.do_return(self);
        }
    }

    // acquire_proc: st[id].h = co_await servers.acquire(); if (balked) co_await h.release();
    void resume_acquire(uint64_t id) {
        int32_t self = (int32_t)(3 + 2 * id);
        // semaphore::down, semaphore.hpp:70-78
        if (value == 0) {
            eng.wait(sem_event, self);
            return;
        }
        --value;
        eng.wake(sem_event);
        has_handle[id] = 1;
        if (balked_flag[id]) sem_up();
        eng.do_return(self);
    }

    rep_result result() const {
        rep_result r;
        r.f64[0] = total_wait;
        r.i64[0] = (int64_t)served;
        r.i64[1] = (int64_t)balked;
        return r;
    }
};

// ---------------------------------------------------------------------------
// Memory hierarchy -- examples/basic_arch_sim.cpp:13-110 with k requesters.
// Per requester q four processes are (re)used for each load: mem2.load (ordinal 0x100+q), its
// _Co_with coroutine (0x8100+q), mem1.load (0x200+q), its _Co_with coroutine (0x8200+q).
// ---------------------------------------------------------------------------
struct archsim_model {
    struct level {
        int64_t latency;
        size_t capacity;
        bool has_next;
        bool owned = false;  // sync::mutex::owned_, mutex.hpp:31-109
        event mtx_event;
        struct block { uint64_t addr; int64_t last_access; };
        std::vector<block> blocks;
    };
    engine eng;
    cxxdes_b200_archsim_params prm;
    uint64_t seed, rep;
    level mem2, mem1;
    std::vector<uint64_t> j, addr;     // per requester: loop index, current address
    std::vector<int64_t> t1, stamp2, stamp1;
    double total_latency = 0;
    uint64_t hits = 0, misses = 0;

    int32_t p_req(uint32_t q) const { return (int32_t)(1 + q); }
    int32_t p_load(uint32_t q, int lvl) const { return (int32_t)(1 + prm.n_requesters + 4 * q + (lvl == 2 ? 0 : 2)); }
    int32_t p_with(uint32_t q, int lvl) const { return p_load(q, lvl) + 1; }

    archsim_model(const cxxdes_b200_archsim_params &p, uint64_t seed_, uint64_t rep_)
        : prm(p), seed(seed_), rep(rep_) {
        mem1.latency = p.l1_latency; mem1.capacity = 1024; mem1.has_next = false;
        mem2.latency = p.l2_latency; mem2.capacity = p.l2_capacity; mem2.has_next = true;
        eng.add_process(0);
        for (uint32_t q = 0; q < p.n_requesters; ++q) eng.add_process(1 + q);
        for (uint32_t q = 0; q < p.n_requesters; ++q) {
            eng.add_process(0x100 + q);
            eng.add_process(0x8100 + q);
            eng.add_process(0x200 + q);
            eng.add_process(0x8200 + q);
        }
        j.assign(p.n_requesters, 0);
        addr.assign(p.n_requesters, 0);
        t1.assign(p.n_requesters, 0);
        stamp2.assign(p.n_requesters, 0);
        stamp1.assign(p.n_requesters, 0);
        eng.bind_root(0);
    }

    void fresh(int32_t p) {  // a new coroutine object in the same slot
        process &pr = eng.procs[p];
        uint32_t ord = pr.ordinal;
        pr = process();
        pr.ordinal = ord;
    }

    void resume(int32_t p) {
        uint32_t k = prm.n_requesters;
        if (p == 0) return resume_main();
        if ((uint32_t)p <= k) return resume_requester((uint32_t)p - 1);
        uint32_t rel = (uint32_t)p - 1 - k, q = rel / 4, slot = rel % 4;
        int lvl = slot < 2 ? 2 : 1;
        if ((slot & 1) == 0) resume_load(q, lvl);
        else resume_with(q, lvl);
    }

    void resume_main() {
        process &me = eng.procs[0];
        uint32_t k = prm.n_requesters;
        if (me.pc == 0) {
            for (uint32_t q = 0; q < k; ++q) eng.bind(p_req(q), me.priority);
            int32_t h = eng.add_handler(true, k, k, token{0, me.priority, 0, NO_HANDLER});
            for (uint32_t q = 0; q < k; ++q) eng.await_completion(p_req(q), 0, h);
            me.pc = 1;
        } else {
            eng.do_return(0);
        }
    }

    void resume_requester(uint32_t q) {
        int32_t self = p_req(q);
        process &me = eng.procs[self];
        if (me.pc == 1) {
            total_latency += (double)(eng.now - t1[q]);
            ++j[q];
        }
        if (j[q] < prm.n_loads) {
            cxxdes_b200::stream_addr a{seed, rep, 1 + q};
            addr[q] = cxxdes_b200::stream_draw_bits(a, j[q]) % prm.addr_space;
            t1[q] = eng.now;
            // co_await mem2.load(addr): start token + completion targeting the requester
            int32_t l = p_load(q, 2);
            fresh(l);
            eng.bind(l, me.priority);
            eng.await_completion(l, self);
            me.pc = 1;
        } else {
            eng.do_return(self);
        }
    }

    // memory::load body up to the _Co_with: basic_arch_sim.cpp:21-46
    void resume_load(uint32_t q, int lvl) {
        int32_t self = p_load(q, lvl);
        process &me = eng.procs[self];
        if (me.pc == 0) {
            (lvl == 2 ? stamp2 : stamp1)[q] = eng.now;  // timestamp = env.now()
            // _Co_with(mtx_){...}: co_yield (mtx_ + body) -> coroutine<void>, co_with.ipp:27-35
            int32_t w = p_with(q, lvl);
            fresh(w);
            eng.bind(w, me.priority);
            eng.await_completion(w, self);
            me.pc = 1;
        } else {
            eng.do_return(self);
        }
    }

    static bool contains(const level &m, uint64_t a) {
        for (auto const &b : m.blocks)
            if (b.addr == a) return true;
        return false;
    }

    // operator+(acquirable&, F&&): acquire; body(); release   (co_with.ipp:27-33)
    void resume_with(uint32_t q, int lvl) {
        int32_t self = p_with(q, lvl);
        process &me = eng.procs[self];
        level &m = lvl == 2 ? mem2 : mem1;
        if (me.pc == 0) {
            // mutex::acquire, mutex.hpp:91-99
            if (m.owned) {
                eng.wait(m.mtx_event, self);
                return;
            }
            m.owned = true;
            // body: basic_arch_sim.cpp:25-45
            if (m.has_next && !contains(m, addr[q])) {
                ++misses;
                if (m.blocks.size() == m.capacity) {
                    size_t victim = 0;
                    for (size_t b = 1; b < m.blocks.size(); ++b)
                        if (m.blocks[b].last_access < m.blocks[victim].last_access) victim = b;
                    m.blocks.erase(m.blocks.begin() + (std::ptrdiff_t)victim);
                }
                int32_t l = p_load(q, 1);  // co_await next_->load(addr)
                fresh(l);
                eng.bind(l, me.priority);
                eng.await_completion(l, self);
                me.pc = 1;
                return;
            }
            if (m.has_next) ++hits;
            eng.timeout(self, m.latency);
            me.pc = 2;
        } else if (me.pc == 1) {
            eng.timeout(self, m.latency);  // co_await timeout(latency_)
            me.pc = 2;
        } else {
            // blocks_[addr].last_access = timestamp
            if (m.has_next) {
                int64_t ts = stamp2[q];
                bool found = false;
                for (auto &b : m.blocks)
                    if (b.addr == addr[q]) { b.last_access = ts; found = true; break; }
                if (!found) m.blocks.push_back(level::block{addr[q], ts});
            }
            // handle.release(): owned_ = false; wake  (mutex.hpp:70-79)
            m.owned = false;
            eng.wake(m.mtx_event);
            eng.do_return(self);
        }
    }

    rep_result result() const {
        rep_result r;
        r.f64[0] = total_latency;
        r.i64[0] = (int64_t)hits;
        r.i64[1] = (int64_t)misses;
        return r;
    }
};

rep_result run_model(int model, const void *params, const cxxdes_b200_clock &clock, uint64_t seed,
                     uint64_t rep, int64_t until, std::vector<trace_event> *dump,
                     size_t dump_limit);

}  // namespace port_oracle
