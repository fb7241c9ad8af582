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
#include "kat_models.hpp"
#include "program_runner.hpp"

#include <cxxdes/sync/queue.hpp>

#include <set>

namespace ref_oracle {

using namespace cxxdes::core::time_ops;

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

// ---------------------------------------------------------------------------
// ALOHA: examples/aloha.cpp:25-85 (pure ALOHA, a frame starts immediately).
// Ordinals / stream ids: co_main = 0, station s = 1 + s.  The offered-load sweep of
// main() (aloha.cpp:87-103) is folded into the replication index: step = rep % steps.
// ---------------------------------------------------------------------------
struct aloha_config {
    std::size_t num_stations = 32;
    float lambda = 2.0f;
    std::size_t packets_per_station = 10000;
    float frame_time = 1.0f;
};

struct aloha_result {
    float g;
    float s;
};

inline aloha_config aloha_config_of(cxxdes_b200_aloha_params const &p, uint64_t rep) {
    aloha_config cfg;
    cfg.num_stations = p.num_stations;
    cfg.frame_time = p.frame_time;
    if (p.sweep_steps > 1) {
        // aloha.cpp:96,101 -- float arithmetic, size_t operands converted to float
        const std::size_t num_steps = p.sweep_steps;
        const std::size_t i = rep % num_steps;
        const float begin = p.lambda_begin, end = p.lambda_end;
        const float lambda = begin + (end - begin) * i / (num_steps - 1);
        cfg.lambda = lambda;
        cfg.packets_per_station = std::size_t(lambda * p.packets_scale);
    } else {
        cfg.lambda = p.lambda_begin;
        cfg.packets_per_station = p.packets_per_station;
    }
    return cfg;
}

CXXDES_SIMULATION(aloha_model) {
    aloha_model(aloha_config const &cfg, cxxdes_b200_clock const &c, uint64_t seed, uint64_t rep,
                tracer &tr)
        : cfg_{cfg}, seed_{seed}, rep_{rep}, tr{tr} {
        env.time_unit(to_time(c.unit_t, c.unit_u));
        env.time_precision(to_time(c.precision_t, c.precision_u));
        bind();  // what run() does first (simulation.hpp:51-54); the harness then steps the env
    }

    const auto &result() const { return result_; }
    std::size_t successful() const { return successful_transmissions_; }

    coroutine<> main_body() {
        std::vector<coroutine<void>> ps;
        ps.reserve(cfg_.num_stations);
        for (std::size_t i = 0; i < cfg_.num_stations; ++i)
            ps.emplace_back(tr.named(station((int)i), (uint32_t)(1 + i)));
        co_await all_of.range(ps.begin(), ps.end());

        result_.s = successful_transmissions_ / now_seconds();
        result_.g = cfg_.lambda * cfg_.frame_time;
    }

    coroutine<> co_main() { return tr.named(main_body(), 0); }

private:
    coroutine<void> station(int id) {
        exp_stream interarrival{seed_, rep_, (uint32_t)(1 + id),
                                cfg_.lambda / cfg_.num_stations};  // float / size_t -> float -> double

        for (std::size_t i = 0; i < cfg_.packets_per_station; ++i) {
            bool collided = false;

            active_transmissions_.insert(&collided);

            check_collisions();
            co_await env.timeout(cfg_.frame_time);

            active_transmissions_.erase(&collided);

            if (!collided) {
                successful_transmissions_++;
            }

            auto wait_time = interarrival();
            co_await env.timeout(wait_time);
        }
    }

    void check_collisions() {
        if (active_transmissions_.size() > 1) {
            for (auto collided : active_transmissions_) *collided = true;
        }
    }

    const aloha_config cfg_;
    aloha_result result_{0, 0};
    uint64_t seed_, rep_;
    tracer &tr;

    std::size_t successful_transmissions_ = 0;
    std::set<bool *> active_transmissions_;
};

inline rep_result run_aloha(const void *params, cxxdes_b200_clock const &clock, uint64_t seed,
                            uint64_t rep, int64_t until, tracer &tr) {
    auto const &p = *static_cast<cxxdes_b200_aloha_params const *>(params);
    aloha_config cfg = aloha_config_of(p, rep);
    aloha_model sim{cfg, clock, seed, rep, tr};
    run_traced(sim.env, tr, until);
    rep_result r;
    r.end_time = sim.now();
    r.f64[0] = (double)sim.result().s;
    r.f64[1] = (double)sim.result().g;
    r.i64[0] = (int64_t)sim.successful();
    r.i64[1] = (int64_t)cfg.packets_per_station;
    return r;
}

// ---------------------------------------------------------------------------
// M/M/c with balking customers on sync::resource + async + any_of (the
// examples/resource.cpp pattern; model text in SURVEY Appendix E).
// Ordinals: co_main 0, source 1, customer i = 2 + 2i, acquire_proc i = 3 + 2i.
// Streams: arrivals = 1, service time of customer i = 2 + 2i (draw 0).
// ---------------------------------------------------------------------------
CXXDES_SIMULATION(mmc_model) {
    mmc_model(cxxdes_b200_mmc_params const &p, cxxdes_b200_clock const &c, uint64_t seed,
              uint64_t rep, tracer &tr)
        : servers{p.servers}, n{p.n_customers}, patience{p.patience}, mu{p.mu}, seed_{seed},
          rep_{rep}, arr{seed, rep, 1, p.lambda}, tr{tr} {
        st.resize(n);
        env.time_unit(to_time(c.unit_t, c.unit_u));
        env.time_precision(to_time(c.precision_t, c.precision_u));
        bind();
    }

    cxxdes::sync::resource servers;
    struct cust {
        bool balked = false;
        cxxdes::sync::resource::handle h;
    };
    std::vector<cust> st;
    std::size_t n;
    double patience, mu;
    uint64_t seed_, rep_;
    exp_stream arr;
    tracer &tr;
    std::size_t served = 0, balked = 0;
    double total_wait = 0;

    coroutine<> acquire_proc(std::size_t id) {
        st[id].h = co_await servers.acquire();
        if (st[id].balked) co_await st[id].h.release();
    }

    coroutine<> customer(std::size_t id) {
        double t0 = now_seconds();
        auto acq = tr.named(acquire_proc(id), (uint32_t)(3 + 2 * id));
        co_await any_of(acq, env.timeout(patience));
        if (!acq.complete()) {
            st[id].balked = true;
            ++balked;
            co_return;
        }
        total_wait += now_seconds() - t0;
        exp_stream svc{seed_, rep_, (uint32_t)(2 + 2 * id), mu};
        co_await env.timeout(svc());
        co_await st[id].h.release();
        ++served;
    }

    coroutine<> source() {
        for (std::size_t i = 0; i < n; ++i) {
            co_await async(tr.named(customer(i), (uint32_t)(2 + 2 * i)));
            co_await env.timeout(arr());
        }
    }

    coroutine<> main_body() { co_await tr.named(source(), 1); }

    coroutine<> co_main() { return tr.named(main_body(), 0); }
};

inline rep_result run_mmc(const void *params, cxxdes_b200_clock const &clock, uint64_t seed,
                          uint64_t rep, int64_t until, tracer &tr) {
    auto const &p = *static_cast<cxxdes_b200_mmc_params const *>(params);
    rep_result r;
    {
        mmc_model sim{p, clock, seed, rep, tr};
        run_traced(sim.env, tr, until);
        r.end_time = sim.now();
        r.f64[0] = sim.total_wait;
        r.i64[0] = (int64_t)sim.served;
        r.i64[1] = (int64_t)sim.balked;
    }
    return r;
}

// ---------------------------------------------------------------------------
// Memory hierarchy: examples/basic_arch_sim.cpp:13-110.  memory::load is the example's
// coroutine (start + completion tokens, _Co_with(mtx_) around the body).  The block store is
// user code, not library code: it is written with a fixed slot array and an explicit eviction
// tie rule (least last_access, lowest slot) instead of std::min_element over an unordered_map,
// whose iteration order is unspecified once several requesters produce equal stamps.
// Ordinals: co_main 0, requester q = 1 + q, mem2.load by q = 0x100 + q, mem1.load by q = 0x200 + q;
// the _Co_with helper coroutines are reported as parent | 0x8000 by the tracer.
// Stream: requester q draws address j as stream (1 + q), draw j, bits % addr_space.
// ---------------------------------------------------------------------------
namespace arch {

struct memory {
    memory(time_integral latency, std::size_t capacity, memory *next, uint32_t ordinal_base,
           tracer &tr)
        : latency_{latency}, capacity_{capacity}, next_{next}, ordinal_base_{ordinal_base}, tr{tr} {}

    coroutine<> load(std::size_t addr, uint32_t requester) {
        time_integral timestamp = (co_await this_environment())->now();
        _Co_with(mtx_) {
            if (next_ && !contains(addr)) {
                ++misses;
                if (blocks_.size() == capacity_) evict();
                co_await tr.named(next_->load(addr, requester), next_->ordinal_base_ + requester);
            } else if (next_) {
                ++hits;
            }
            co_await timeout(latency_);
            touch(addr, timestamp);
        };
    }

    std::size_t hits = 0, misses = 0;

private:
    struct block {
        std::size_t addr = 0;
        time_integral last_access = 0;
    };

    bool contains(std::size_t addr) const {
        for (auto const &b : blocks_)
            if (b.addr == addr) return true;
        return false;
    }

    void evict() {
        std::size_t victim = 0;
        for (std::size_t i = 1; i < blocks_.size(); ++i)
            if (blocks_[i].last_access < blocks_[victim].last_access) victim = i;
        blocks_.erase(blocks_.begin() + (std::ptrdiff_t)victim);
    }

    // blocks_[addr].last_access = timestamp: inserts the block when absent
    void touch(std::size_t addr, time_integral timestamp) {
        if (!next_) return;  // the last level's map is write-only in the example
        for (auto &b : blocks_)
            if (b.addr == addr) {
                b.last_access = timestamp;
                return;
            }
        blocks_.push_back(block{addr, timestamp});
    }

    time_integral latency_;
    std::size_t capacity_;
    memory *next_;
    uint32_t ordinal_base_;
    tracer &tr;
    cxxdes::sync::mutex mtx_;
    std::vector<block> blocks_;
};

}  // namespace arch

CXXDES_SIMULATION(archsim_model) {
    archsim_model(cxxdes_b200_archsim_params const &p, uint64_t seed, uint64_t rep, tracer &tr)
        : prm{p}, seed_{seed}, rep_{rep}, tr{tr}, mem1{p.l1_latency, 1024, nullptr, 0x200, tr},
          mem2{p.l2_latency, p.l2_capacity, &mem1, 0x100, tr} {
        bind();
    }

    cxxdes_b200_archsim_params prm;
    uint64_t seed_, rep_;
    tracer &tr;
    arch::memory mem1, mem2;
    double total_latency = 0;

    coroutine<> requester(uint32_t q) {
        cxxdes_b200::stream_addr a{seed_, rep_, 1 + q};
        for (uint64_t j = 0; j < prm.n_loads; ++j) {
            std::size_t addr = (std::size_t)(cxxdes_b200::stream_draw_bits(a, j) % prm.addr_space);
            auto t1 = now();
            co_await tr.named(mem2.load(addr, q), 0x100 + q);
            auto t2 = now();
            total_latency += (double)(t2 - t1);
        }
    }

    coroutine<> main_body() {
        std::vector<coroutine<void>> ps;
        for (uint32_t q = 0; q < prm.n_requesters; ++q) ps.emplace_back(tr.named(requester(q), 1 + q));
        co_await all_of.range(ps.begin(), ps.end());
    }

    coroutine<> co_main() { return tr.named(main_body(), 0); }
};

inline rep_result run_archsim(const void *params, cxxdes_b200_clock const &clock, uint64_t seed,
                              uint64_t rep, int64_t until, tracer &tr) {
    auto const &p = *static_cast<cxxdes_b200_archsim_params const *>(params);
    rep_result r;
    {
        archsim_model sim{p, seed, rep, tr};
        run_traced(sim.env, tr, until);
        r.end_time = sim.now();
        r.f64[0] = sim.total_latency;
        r.i64[0] = (int64_t)sim.mem2.hits;
        r.i64[1] = (int64_t)sim.mem2.misses;
    }
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
        case CXXDES_B200_MODEL_ALOHA: r = run_aloha(params, clock, seed, rep, until, tr); break;
        case CXXDES_B200_MODEL_MMC: r = run_mmc(params, clock, seed, rep, until, tr); break;
        case CXXDES_B200_MODEL_ARCHSIM: r = run_archsim(params, clock, seed, rep, until, tr); break;
        case ORACLE_MODEL_KAT: r = run_kat(params, until, tr, nullptr); break;
        case CXXDES_B200_MODEL_PROGRAM: r = run_program(params, clock, seed, rep, until, tr); break;
        default: throw std::runtime_error("ref_oracle: unknown model");
    }
    r.n_events = tr.n_events;
    r.trace_hash = tr.hash;
    return r;
}

}  // namespace ref_oracle
