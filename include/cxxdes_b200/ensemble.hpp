// cxxdes_b200/ensemble.hpp -- host C++ layer: the reference's vocabulary for a batched run path.
//
// libcxxdes users write
//     CXXDES_SIMULATION(producer_consumer_example) { ... coroutine<> co_main(); };
//     producer_consumer_example sim{lambda, mu};          // env.time_unit(1_s); env.time_precision(1_ms);
//     sim.run();  sim.now();  sim.now_seconds();  sim.avg_latency;
// (include/cxxdes/core/simulation.hpp:31-85, examples/producer_consumer.cpp:61-75).  This header keeps
// those names for an ENSEMBLE of such simulations running on a B200:
//     cxxdes::b200::ensemble<cxxdes::b200::models::producer_consumer> ens{{lambda, mu, 10000}, 1 << 20};
//     ens.env.time_unit(1_s);  ens.env.time_precision(1_ms);
//     ens.run();                       // every replication: bind co_main, while (step());
//     ens.now()[r]; ens.now_seconds()[r]; ens.stat(0)[r] /* avg_latency */; ens.reduce().n_events;
// It is a thin, header-only wrapper over the C ABI in cxxdes_b200.h (plain g++ TU -- nvcc cannot
// parse the reference headers, and does not need to see this one).  Errors follow the reference's
// convention: std::runtime_error (environment.ipp:44-45,61-62).
#pragma once
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "cxxdes_b200.h"

namespace cxxdes {
namespace b200 {

using time_integral = std::intmax_t;  // core/impl/defs.ipp:7
using priority_type = std::intmax_t;  // defs.ipp:4
using real_type = double;             // defs.ipp:10

// cxxdes::time_utils::time<>: {t, unit}  (misc/time.hpp:31-37)
enum class time_units : int { seconds = 0, milliseconds, microseconds, nanoseconds, picoseconds };
struct time {
    time_integral t = 1;
    time_units u = time_units::seconds;
};
constexpr time one_second{1, time_units::seconds};

namespace time_ops {  // the literals of cxxdes::core::time_ops (misc/time.hpp:367-409)
constexpr time operator""_s(unsigned long long x) { return {(time_integral)x, time_units::seconds}; }
constexpr time operator""_ms(unsigned long long x) { return {(time_integral)x, time_units::milliseconds}; }
constexpr time operator""_us(unsigned long long x) { return {(time_integral)x, time_units::microseconds}; }
constexpr time operator""_ns(unsigned long long x) { return {(time_integral)x, time_units::nanoseconds}; }
constexpr time operator""_ps(unsigned long long x) { return {(time_integral)x, time_units::picoseconds}; }
// `3_x`: three model units, mapped to ticks through the environment's time_unit() (misc/time.hpp:128-148,308-310)
struct unitless_time {
    time_integral t = 0;
};
constexpr unitless_time operator""_x(unsigned long long x) { return {(time_integral)x}; }
}  // namespace time_ops

// time::count (misc/time.hpp:75-86): this quantity in ticks of `precision`
constexpr time_integral count(time x, time precision) {
    constexpr time_integral conv[5] = {1, 1000, 1000000, 1000000000, 1000000000000LL};
    if ((int)x.u < (int)precision.u) return conv[(int)precision.u - (int)x.u] * x.t / precision.t;
    return (x.t / precision.t) / conv[(int)x.u - (int)precision.u];
}

// Model descriptors: the constructor arguments of the reference examples.
namespace models {
struct producer_consumer {  // examples/producer_consumer.cpp:10-20
    static constexpr int id = CXXDES_B200_MODEL_MM1;
    using params_type = cxxdes_b200_mm1_params;
    double lambda, mu;
    std::uint64_t n_packets = 10000;
    params_type params() const { return {lambda, mu, n_packets}; }
};
struct aloha {  // aloha_config + the sweep of main(), examples/aloha.cpp:13-18,87-103
    static constexpr int id = CXXDES_B200_MODEL_ALOHA;
    using params_type = cxxdes_b200_aloha_params;
    std::uint32_t num_stations = 32, sweep_steps = 0;
    float lambda_begin = 2.0f, lambda_end = 2.0f, frame_time = 1.0f, packets_scale = 1000.0f;
    std::uint64_t packets_per_station = 10000;
    params_type params() const {
        return {num_stations, sweep_steps, lambda_begin, lambda_end, frame_time, packets_scale, packets_per_station};
    }
};
struct mmc_balking {  // sync::resource + async + any_of (examples/resource.cpp pattern)
    static constexpr int id = CXXDES_B200_MODEL_MMC;
    using params_type = cxxdes_b200_mmc_params;
    double lambda, mu, patience;
    std::uint32_t servers = 8;
    std::uint64_t n_customers = 1000;
    params_type params() const { return {lambda, mu, patience, servers, 0, n_customers}; }
};
struct memory_hierarchy {  // examples/basic_arch_sim.cpp:83-109
    static constexpr int id = CXXDES_B200_MODEL_ARCHSIM;
    using params_type = cxxdes_b200_archsim_params;
    std::int64_t l2_latency = 1, l1_latency = 10;
    std::uint32_t l2_capacity = 16, addr_space = 32, n_requesters = 1;
    std::uint64_t n_loads = 1024;
    params_type params() const { return {l2_latency, l1_latency, l2_capacity, addr_space, n_requesters, 0, n_loads}; }
};
}  // namespace models

// Ensemble estimate of statistic k from the reduced accumulators: mean over the replications and the half-width
// of its 95 % normal-approximation interval (what the examples' main() would compute by looping over runs).
struct estimate {
    double mean, ci95;
};
inline estimate mean_ci(const cxxdes_b200_accumulators &acc, int k, double z = 1.959964) {
    const double n = (double)acc.n_replications;
    if (acc.n_replications == 0) return {0.0 / 0.0, 0.0 / 0.0};
    const double mean = acc.f64_sum[k] / n;
    if (acc.n_replications < 2) return {mean, 0.0 / 0.0};
    double var = (acc.f64_sumsq[k] - n * mean * mean) / (n - 1.0);
    if (var < 0.0) var = 0.0;
    return {mean, z * std::sqrt(var / n)};
}

namespace detail {
inline void check(int rc) {
    if (rc != CXXDES_B200_OK) throw std::runtime_error(std::string("cxxdes_b200: ") + cxxdes_b200_last_error());
}
}  // namespace detail

template <typename Model>
class ensemble {
public:
    // The subset of `environment` that configures the clock (environment.ipp:43-70).
    struct environment_config {
        void time_unit(time x) {
            if (owner_->handle_) throw std::runtime_error("cannot call time_unit(time x) on a used environment!");
            unit_ = x;
        }
        time time_unit() const noexcept { return unit_; }
        void time_precision(time x) {
            if (owner_->handle_) throw std::runtime_error("cannot call time_precision(time x) on a used environment!");
            prec_ = x;
        }
        time time_precision() const noexcept { return prec_; }

    private:
        friend class ensemble;
        ensemble *owner_ = nullptr;
        time unit_ = one_second, prec_ = one_second;  // environment(): unit = precision = one_second
    };

    environment_config env;

    ensemble(Model model, std::uint64_t n_replications, std::uint64_t seed = 0x5EED,
             std::uint64_t first_replication = 0, int device = 0)
        : model_(model), n_(n_replications), seed_(seed), first_(first_replication), device_(device) {
        env.owner_ = this;
    }
    ensemble(const ensemble &) = delete;
    ensemble &operator=(const ensemble &) = delete;
    ~ensemble() {
        if (handle_) cxxdes_b200_destroy(handle_);
    }

    std::uint64_t size() const noexcept { return n_; }

    // simulation::run(): bind co_main once, env.run()  (simulation.hpp:51-54)
    void run() {
        bind();
        detail::check(cxxdes_b200_run(handle_));
        fetched_ = false;
    }
    // simulation::run_for(t) / environment::run_until(t) in ticks (simulation.hpp:58-62, environment.ipp:190-214)
    void run_for(time_integral dt) {
        bind();
        detail::check(cxxdes_b200_run_for(handle_, dt));
        fetched_ = false;
    }
    void run_until(time_integral t) {
        bind();
        detail::check(cxxdes_b200_run_until(handle_, t));
        fetched_ = false;
    }

    // sim.now() of every replication, in ticks (simulation.hpp:33-35)
    const std::vector<std::int64_t> &now() { return fetch().end_time_; }
    // sim.now_seconds() (simulation.hpp:43-45; environment.ipp:29-36; misc/time.hpp:120-125)
    std::vector<real_type> now_seconds() {
        static const double conv[5] = {1, 1e-3, 1e-6, 1e-9, 1e-12};
        const auto &t = now();
        std::vector<real_type> s(t.size());
        for (std::size_t r = 0; r < t.size(); ++r)
            s[r] = (real_type)(t[r] * env.prec_.t) * conv[(int)env.prec_.u];
        return s;
    }
    const std::vector<std::uint64_t> &n_events() { return fetch().n_events_; }
    const std::vector<std::uint64_t> &trace_hash() { return fetch().trace_hash_; }
    const std::vector<std::uint32_t> &status() { return fetch().status_; }
    // model statistics: stat(0) is avg_latency for producer_consumer, S for aloha, ... (cxxdes_b200.h)
    const std::vector<double> &stat(int k) { return fetch().f64_[k]; }
    const std::vector<std::int64_t> &count(int k) { return fetch().i64_[k]; }

    // ensemble accumulators of this shard (summed on the device)
    cxxdes_b200_accumulators reduce() {
        bind();
        cxxdes_b200_accumulators acc;
        detail::check(cxxdes_b200_reduce(handle_, &acc, nullptr));
        return acc;
    }

    // Totals over ALL shards: the one collective of a multi-GPU run (an ncclAllReduce of the 88-byte accumulator
    // struct over `nccl_comm`, an ncclComm_t with one rank per shard).  Every rank must call it.
    cxxdes_b200_accumulators reduce_allgpus(void *nccl_comm) {
        bind();
        cxxdes_b200_accumulators acc;
        detail::check(cxxdes_b200_reduce_allgpus(handle_, nccl_comm, &acc));
        return acc;
    }

    float last_run_ms() {
        float ms = 0;
        detail::check(cxxdes_b200_last_run_ms(handle_, &ms));
        return ms;
    }

    // The tokens environment::step popped for ONE replication, in order (what CXXDES_DEBUG_TOKEN would print,
    // token.ipp:50-60): the replication is run again from tick 0 -- to the end, or to `until` >= 0 -- with a sink on the
    // dispatch loop.  `digest` equal to trace_hash()[replication] means: this is what the ensemble's kernel dispatched.
    struct popped_token {
        time_integral time;
        priority_type priority;
        std::uint32_t kind;       // CXXDES_B200_TOKEN_RESUME / _HANDLER / _NOOP
        std::uint32_t process;    // ordinal, 0xFFFF for a null-target token
    };
    struct event_trace {
        std::uint64_t n_events = 0, digest = 0;
        std::vector<popped_token> tokens;
    };
    event_trace trace(std::uint64_t replication, std::uint64_t max_events = 4096, time_integral until = -1) {
        bind();
        std::vector<std::int64_t> t(max_events), pr(max_events);
        std::vector<std::uint32_t> k(max_events), who(max_events);
        std::uint64_t written = 0;
        event_trace out;
        detail::check(cxxdes_b200_trace(handle_, replication, until, max_events, t.data(), pr.data(), k.data(), who.data(),
                                        &written, &out.n_events, &out.digest));
        for (std::uint64_t i = 0; i < written; ++i) out.tokens.push_back({t[i], pr[i], k[i], who[i]});
        return out;
    }
    // which first-pass kernel instantiation run() launches; how many replications the last launch redid elsewhere
    std::string kernel_name() {
        bind();
        char buf[160];
        detail::check(cxxdes_b200_kernel_name(handle_, buf, sizeof buf));
        return buf;
    }
    // the same answer before anything is created -- no handle, no device (ALOHA, M/M/c, memory hierarchy: the kernel
    // follows from the parameters and the clock; cxxdes_b200_plan_kernel; the others throw)
    std::string planned_kernel() const {
        typename Model::params_type p = model_.params();
        cxxdes_b200_config cfg = config_of(p);
        char buf[160];
        detail::check(cxxdes_b200_plan_kernel(&cfg, buf, sizeof buf));
        return buf;
    }
    struct pass_counts_t {
        std::uint64_t second_pass, helper, fallback, helper_longest_ns;
    };
    pass_counts_t pass_counts() {
        std::uint64_t c[4] = {0, 0, 0, 0};
        detail::check(cxxdes_b200_pass_counts(handle_, c));
        return {c[0], c[1], c[2], c[3]};
    }

private:
    cxxdes_b200_config config_of(const typename Model::params_type &p) const {
        cxxdes_b200_config cfg{};
        cfg.model = Model::id;
        cfg.device = device_;
        cfg.seed = seed_;
        cfg.first_replication = first_;
        cfg.n_replications = n_;
        cfg.clock.unit_t = env.unit_.t;
        cfg.clock.unit_u = (int)env.unit_.u;
        cfg.clock.precision_t = env.prec_.t;
        cfg.clock.precision_u = (int)env.prec_.u;
        cfg.params = &p;
        cfg.params_size = sizeof(p);
        return cfg;
    }
    void bind() {
        if (handle_) return;
        typename Model::params_type p = model_.params();
        cxxdes_b200_config cfg = config_of(p);
        detail::check(cxxdes_b200_create(&cfg, &handle_));
    }

    ensemble &fetch() {
        if (fetched_) return *this;
        bind();
        end_time_.resize(n_);
        n_events_.resize(n_);
        trace_hash_.resize(n_);
        status_.resize(n_);
        cxxdes_b200_columns c{};
        c.end_time = end_time_.data();
        c.n_events = n_events_.data();
        c.trace_hash = trace_hash_.data();
        c.status = status_.data();
        for (int k = 0; k < CXXDES_B200_N_F64; ++k) {
            f64_[k].resize(n_);
            c.f64[k] = f64_[k].data();
        }
        for (int k = 0; k < CXXDES_B200_N_I64; ++k) {
            i64_[k].resize(n_);
            c.i64[k] = i64_[k].data();
        }
        detail::check(cxxdes_b200_fetch(handle_, &c));
        fetched_ = true;
        return *this;
    }

    Model model_;
    std::uint64_t n_, seed_, first_;
    int device_;
    cxxdes_b200_ensemble *handle_ = nullptr;
    bool fetched_ = false;
    std::vector<std::int64_t> end_time_;
    std::vector<std::uint64_t> n_events_, trace_hash_;
    std::vector<std::uint32_t> status_;
    std::vector<double> f64_[CXXDES_B200_N_F64];
    std::vector<std::int64_t> i64_[CXXDES_B200_N_I64];
};

}  // namespace b200
}  // namespace cxxdes
