// oracle/port/oracle_api.cpp -- TEST INFRASTRUCTURE (oracle), not product code.
//
// C entry points of the restated CPU engine (libcxxdes_oracle.so).  Same
// signatures as oracle/ref/ref_oracle.cpp so tests can swap one for the other.
#include <atomic>
#include <queue>
#include <thread>

#include "models.hpp"
#include "program_interp.hpp"

using namespace port_oracle;

namespace port_oracle {

template <typename Model>
static rep_result finish(Model &m, int64_t until, std::vector<trace_event> *dump,
                         size_t dump_limit) {
    m.eng.dump = dump;
    m.eng.dump_limit = dump_limit;
    m.eng.run(m, until);
    rep_result r = m.result();
    r.end_time = m.eng.now;
    r.n_events = m.eng.n_events;
    r.trace_hash = m.eng.hash;
    r.peak_heap = m.eng.peak_heap;
    return r;
}

rep_result run_model(int model, const void *params, const cxxdes_b200_clock &clock, uint64_t seed,
                     uint64_t rep, int64_t until, std::vector<trace_event> *dump,
                     size_t dump_limit) {
    switch (model) {
        case CXXDES_B200_MODEL_MM1: {
            mm1_model m(*static_cast<const cxxdes_b200_mm1_params *>(params), clock, seed, rep);
            return finish(m, until, dump, dump_limit);
        }
        case CXXDES_B200_MODEL_ALOHA: {
            aloha_model m(*static_cast<const cxxdes_b200_aloha_params *>(params), clock, seed, rep);
            return finish(m, until, dump, dump_limit);
        }
        case CXXDES_B200_MODEL_MMC: {
            mmc_model m(*static_cast<const cxxdes_b200_mmc_params *>(params), clock, seed, rep);
            return finish(m, until, dump, dump_limit);
        }
        case CXXDES_B200_MODEL_ARCHSIM: {
            archsim_model m(*static_cast<const cxxdes_b200_archsim_params *>(params), seed, rep);
            return finish(m, until, dump, dump_limit);
        }
        case CXXDES_B200_MODEL_PROGRAM: {
            program_model m(*static_cast<const cxxdes_b200_program *>(params), clock, seed, rep);
            return finish(m, until, dump, dump_limit);
        }
        default:
            throw std::runtime_error("port_oracle: unknown model");
    }
}

}  // namespace port_oracle

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

}  // namespace

extern "C" {

int cxxdes_oracle_run(int model, const void *params, const cxxdes_b200_clock *clock,
                      uint64_t seed, uint64_t first_rep, uint64_t n_reps, int n_threads,
                      int64_t run_until, const cxxdes_b200_columns *out) {
    try {
        parallel_for(n_reps, n_threads, [&](uint64_t i) {
            rep_result r =
                run_model(model, params, *clock, seed, first_rep + i, run_until, nullptr, 0);
            if (out->end_time) out->end_time[i] = r.end_time;
            if (out->n_events) out->n_events[i] = r.n_events;
            if (out->trace_hash) out->trace_hash[i] = r.trace_hash;
            if (out->status) out->status[i] = r.status;
            for (int k = 0; k < CXXDES_B200_N_F64; ++k)
                if (out->f64[k]) out->f64[k][i] = r.f64[k];
            for (int k = 0; k < CXXDES_B200_N_I64; ++k)
                if (out->i64[k]) out->i64[k][i] = r.i64[k];
        });
    } catch (std::exception const &) {
        return 1;
    }
    return 0;
}

uint64_t cxxdes_oracle_trace(int model, const void *params, const cxxdes_b200_clock *clock,
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

int cxxdes_oracle_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }

// ---- pins for the shared arithmetic and the restated heap (used by CPU tests) ----

void cxxdes_oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    cxxdes_b200::philox_block b =
        cxxdes_b200::philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    for (int i = 0; i < 4; ++i) out[i] = b.w[i];
}

void cxxdes_oracle_draw_bits(uint64_t seed, uint64_t rep, uint32_t stream, uint64_t first_draw,
                             uint64_t n, uint64_t *out) {
    cxxdes_b200::stream_addr a{seed, rep, stream};
    for (uint64_t i = 0; i < n; ++i) out[i] = cxxdes_b200::stream_draw_bits(a, first_draw + i);
}

void cxxdes_oracle_log(const double *w, uint64_t n, double *out) {
    for (uint64_t i = 0; i < n; ++i)
        out[i] = cxxdes_b200::log_pos(w[i], cxxdes_b200::LOG_TABLE_HOST);
}

void cxxdes_oracle_unit_complement(const uint64_t *bits, uint64_t n, double *out) {
    for (uint64_t i = 0; i < n; ++i) out[i] = cxxdes_b200::unit_complement_from_bits(bits[i]);
}

void cxxdes_oracle_divide(const double *x, double d, uint64_t n, double *out) {
    cxxdes_b200::divisor v = cxxdes_b200::make_divisor(d);
    for (uint64_t i = 0; i < n; ++i) out[i] = cxxdes_b200::divide(x[i], v);
}

void cxxdes_oracle_exponential_ticks(uint64_t seed, uint64_t rep, uint32_t stream, double rate,
                                     int64_t ticks_per_unit, uint64_t first_draw, uint64_t n,
                                     double *variate, int64_t *ticks) {
    cxxdes_b200::stream_addr a{seed, rep, stream};
    cxxdes_b200::divisor v = cxxdes_b200::make_divisor(rate);
    cxxdes_b200::clock_config c{ticks_per_unit, 1, 1e-3};
    for (uint64_t i = 0; i < n; ++i) {
        double e = cxxdes_b200::exponential(a, first_draw + i, v, cxxdes_b200::LOG_TABLE_HOST);
        if (variate) variate[i] = e;
        if (ticks) ticks[i] = cxxdes_b200::real_to_sim(e, c);
    }
}

uint64_t cxxdes_oracle_trace_hash(const int64_t *time, const int64_t *priority,
                                  const uint32_t *kind, const uint32_t *proc, uint64_t n) {
    uint64_t h = cxxdes_b200::TRACE_HASH_INIT;
    for (uint64_t i = 0; i < n; ++i)
        h = cxxdes_b200::trace_hash_step(h, time[i], priority[i], kind[i], proc[i]);
    return h;
}

// Drive the restated heap and libstdc++'s std::priority_queue with the same
// random schedule of pushes/pops over a small key range (many ties); returns
// the number of pops whose payload differed (must be 0).
uint64_t cxxdes_oracle_heap_check(uint64_t seed, uint64_t n_ops, uint32_t key_range,
                                  uint32_t max_size) {
    struct cmp {
        bool operator()(const token *a, const token *b) const { return token_later(*a, *b); }
    };
    std::priority_queue<token *, std::vector<token *>, cmp> ref;
    ref_heap mine;
    std::vector<token *> owned;
    cxxdes_b200::stream_addr a{seed, 0, 77};
    uint64_t mismatches = 0;
    int32_t serial = 0;
    for (uint64_t op = 0; op < n_ops; ++op) {
        uint64_t bits = cxxdes_b200::stream_draw_bits(a, op);
        bool do_push = mine.size() == 0 || (mine.size() < max_size && (bits & 3) != 0);
        if (do_push) {
            token *t = new token{(int64_t)((bits >> 8) % key_range),
                                 (int64_t)((bits >> 40) % 3) - 1, serial++, NO_HANDLER};
            owned.push_back(t);
            ref.push(t);
            mine.push(*t);
        } else {
            token *r = ref.top();
            ref.pop();
            token m = mine.top();
            mine.pop();
            if (r->target != m.target || r->time != m.time || r->priority != m.priority)
                ++mismatches;
        }
    }
    for (token *t : owned) delete t;
    return mismatches;
}

}  // extern "C"
