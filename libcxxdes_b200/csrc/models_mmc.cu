// models_mmc.cu -- M/M/c with balking customers on the device engine.
//
// The model is written on the reference's primitives: sync::resource
// (resource.hpp:37-100 = semaphore{c, c}, semaphore.hpp:58-78, over one sync::event), async
// (async.ipp:21-28) and any_of (any_of.ipp:3-92).  Everything observable is reproduced: every
// down()/up() wakes ALL waiters (thundering herd, event.hpp:125-134), any_of never cancels the
// loser (its token still pops as a handler event), async leaves a null-target completion that
// pops as a no-op.
//
// Processes / token targets: co_main 0, source 1, customer i = 2 + 2i, acquire_proc i = 3 + 2i.
// A handler token always targets the customer whose any_of owns it, so the handler is found
// from the target.  Per-customer state is one 64-bit word in global memory laid out
// [customer][thread] (lanes walk the customers roughly in step, so accesses coalesce).
#include <cstdlib>

#include "device/engine32.cuh"
#include "device/models_common.cuh"
#include "device/trace.cuh"
#include "kernels.h"

namespace cxb {

namespace {

constexpr int MMC_THREADS = 96;
constexpr int MMC_HEAP = 96;
constexpr int MMC_WAITERS = 96;

// customer word: bits 0-1 pc, 2 balked, 3 acquire_proc complete, 4 any_of armed, 5-6 remaining,
// 16.. tick of t0 = now_seconds() at arrival (a pure function of the tick)
constexpr uint64_t C_PC = 3, C_BALKED = 4, C_ACQ_DONE = 8, C_ARMED = 16, C_REM_SHIFT = 5, C_REM = 3ULL << 5;
constexpr int C_T0_SHIFT = 16;

}  // namespace

// Generic-engine kernel: 64-bit keys, 96-entry heap and waiter list.  Runs the whole shard, or
// (list != nullptr) only the replications the compact kernel below handed back.
// ENV = traced_environment: the same kernel writing the event trace of its replications (cxxdes_b200_trace).
// DEEP: the last pass.  The reference's containers are unbounded (std::priority_queue, the event's std::vector of
// waiters); a replication that outgrew the 96 entries above -- an overloaded system: acquire_procs of customers who
// balked keep waiting for a server -- is run again with its heap and wait list in GLOBAL memory, sized from
// n_customers for the worst case (mmc_deep_plan), so that it gets the reference's result instead of an error flag.
// The pass walks the whole shard and takes the replications whose status carries a capacity flag.
template <typename ENV, bool DEEP = false>
__global__ void __launch_bounds__(MMC_THREADS)
mmc_kernel(cxxdes_b200_mmc_params prm, shard sh, device_columns out, uint64_t *cust_base, const uint32_t *list,
           const unsigned long long *list_count, unsigned long long *list_cursor, trace_sink sink, mmc_deep deep) {
    extern __shared__ __align__(16) unsigned char smem[];
    log_entry *table = (log_entry *)smem;
    unsigned char *cursor = smem + sizeof(log_entry) * (1 << CXB_LOG_TABLE_BITS);
    stage_log_table(table, sh.log_table);
    lane_array<uint64_t> hkey;
    lane_array<uint32_t> htag, wwho;
    lane_array<int32_t> wlat;
    int heap_cap = MMC_HEAP, waiter_cap = MMC_WAITERS;
    if constexpr (DEEP) {
        // [entry][thread of the deep grid]: keys, tags, waiters, wait latencies
        const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (size_t)deep.threads;
        heap_cap = deep.heap_cap;
        waiter_cap = deep.waiter_cap;
        unsigned char *g = deep.base;
        hkey.base = (uint64_t *)g + t;
        g += 8 * nt * (size_t)heap_cap;
        htag.base = (uint32_t *)g + t;
        g += 4 * nt * (size_t)heap_cap;
        wwho.base = (uint32_t *)g + t;
        g += 4 * nt * (size_t)waiter_cap;
        wlat.base = (int32_t *)g + t;
        hkey.stride = htag.stride = wwho.stride = wlat.stride = deep.threads;
    } else {
        hkey = carve<uint64_t>(cursor, MMC_HEAP);
        htag = carve<uint32_t>(cursor, MMC_HEAP);
        wwho = carve<uint32_t>(cursor, MMC_WAITERS);
        wlat = carve<int32_t>(cursor, MMC_WAITERS);
    }

    const uint64_t total_threads = (uint64_t)gridDim.x * blockDim.x;
    uint64_t *cust = cust_base + ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
    const uint64_t n = prm.n_customers;
    const int64_t patience_ticks = real_to_sim(prm.patience, sh.clk);  // env.timeout(patience)
    const divisor d_lambda = make_divisor(prm.lambda), d_mu = make_divisor(prm.mu);
    const uint32_t NOOP_TAG = make_tag(TAG_NO_TARGET);

    uint64_t rep;
    [[maybe_unused]] unsigned long long chunk_next = 0, chunk_end = 0;
    for (;;) {
        if (DEEP && !list_cursor) {
            if (!steal(sh, rep)) break;      // (the event trace of one replication, with the deep pass's storage)
        } else if constexpr (DEEP) {
            // (list_cursor: this pass's own work counter, claimed 16 replications at a time; list_count: where it
            // counts what it took)
            if (chunk_next == chunk_end) {
                chunk_next = atomicAdd(list_cursor, 16ULL);
                chunk_end = chunk_next + 16;
            }
            if (chunk_next >= sh.n_replications) break;
            rep = chunk_next++;
            if (!(out.status[rep] & (CXXDES_B200_REP_HEAP_OVERFLOW | CXXDES_B200_REP_WAITER_OVERFLOW))) continue;
            atomicAdd((unsigned long long *)list_count, 1ULL);
        } else if (list) {
            const unsigned long long idx = atomicAdd(list_cursor, 1ULL);
            if (idx >= *list_count) break;
            rep = list[idx];
        } else if (!steal(sh, rep)) {
            break;
        }
        const uint64_t grep = sh.first_replication + rep;
        stream_addr addr;
        addr.seed = sh.seed;
        addr.replication = grep;

        ENV env;
        env.init(hkey, htag, heap_cap);
        attach_trace(env, sink);
        wait_list waiters;  // the semaphore's event
        waiters.init(wwho, wlat, waiter_cap);
        uint64_t value = prm.servers;  // semaphore{c, c}
        uint64_t i = 0, served = 0, balked = 0;
        uint32_t main_pc = 0, source_pc = 0;
        double total_wait = 0.0;
        for (uint64_t c = 0; c < n; ++c) cust[c * total_threads] = 0;

        env.schedule(0, 0, make_tag(0));  // environment::bind(co_main)

        while (env.has_event() && (sh.run_until < 0 || env.next_time() <= sh.run_until)) {
            token t = env.step_pop();
            const uint32_t target = tag_target(t.tag);
            if (tag_handler_plus1(t.tag)) {
                // any_of handler of customer (target - 2) / 2: any_of.ipp:9-26
                uint64_t *w = &cust[(uint64_t)((target - 2) >> 1) * total_threads];
                uint64_t s = *w;
                uint64_t remaining = ((s & C_REM) >> C_REM_SHIFT) - 1;
                s = (s & ~C_REM) | (remaining << C_REM_SHIFT);
                if ((s & C_ARMED) && remaining < 2) {
                    env.schedule(key_time(t.key), key_prio(t.key), make_tag(target));  // output token
                    s &= ~C_ARMED;
                }
                *w = s;
            } else if (target == 0) {
                if (main_pc == 0) {
                    env.schedule(env.now, 0, make_tag(1));  // co_await source(): start token; completion targets main
                    main_pc = 1;
                } else {
                    env.schedule(env.now, 0, NOOP_TAG);
                }
            } else if (target == 1) {
                // source: for i < n { co_await async(customer(i)); co_await env.timeout(arr()); }
                if (source_pc == 1) ++i;
                if (i < n) {
                    env.schedule(env.now, 0, make_tag((uint32_t)(2 + 2 * i)));  // async: start token (async.ipp:21-28)
                    addr.stream = 1;
                    const double a = exponential(addr, i, d_lambda, table);
                    env.schedule(env.now + real_to_sim(a, sh.clk), 0, make_tag(1));
                    source_pc = 1;
                } else {
                    env.schedule(env.now, 0, make_tag(0));  // source returns -> completion resumes co_main
                }
            } else if (target != TAG_NO_TARGET) {
                const uint64_t id = (uint64_t)(target - 2) >> 1;
                uint64_t *w = &cust[id * total_threads];
                uint64_t s = *w;
                if ((target & 1) == 0) {
                    // ---- customer(id) ----
                    const uint64_t pc = s & C_PC;
                    if (pc == 0) {
                        // t0 = now_seconds(); any_of(acquire_proc, env.timeout(patience))
                        s = ((uint64_t)env.now << C_T0_SHIFT) | 1 | C_ARMED | (2ULL << C_REM_SHIFT);
                        env.schedule(env.now, 0, make_tag(target + 1));                   // await_bind: start(acquire_proc)
                        env.schedule(env.now + patience_ticks, 0, make_tag(target, 1));  // await_suspend: timeout + handler
                    } else if (pc == 1) {
                        if (!(s & C_ACQ_DONE)) {
                            s = (s & ~C_PC) | 3 | C_BALKED;  // balked: co_return
                            ++balked;
                            env.schedule(env.now, 0, NOOP_TAG);  // async's null-target completion
                        } else {
                            const double t0 = sim_to_seconds((int64_t)(s >> C_T0_SHIFT), sh.clk);
                            total_wait += sim_to_seconds(env.now, sh.clk) - t0;
                            addr.stream = target;
                            const double svc = exponential(addr, 0, d_mu, table);
                            env.schedule(env.now + real_to_sim(svc, sh.clk), 0, make_tag(target));
                            s = (s & ~C_PC) | 2;
                        }
                    } else {
                        // release: semaphore::up -> ++value; wake  (semaphore.hpp:58-66)
                        ++value;
                        waiters.wake(env);
                        ++served;
                        s = (s & ~C_PC) | 3;
                        env.schedule(env.now, 0, NOOP_TAG);
                    }
                } else {
                    // ---- acquire_proc(id): st.h = co_await servers.acquire(); if (balked) release ----
                    if (value == 0) {
                        waiters.wait(env, target, 0);  // semaphore::down blocks (semaphore.hpp:70-78)
                    } else {
                        --value;
                        waiters.wake(env);  // down() wakes every waiter even though it consumed a unit
                        if (s & C_BALKED) {
                            ++value;        // late acquisition is released at once
                            waiters.wake(env);
                        }
                        s |= C_ACQ_DONE;
                        env.schedule(env.now, 0, make_tag(target - 1, 1));  // completion token with the any_of handler
                    }
                }
                *w = s;
            }
            if (env.status) break;
        }
        if (sh.run_until >= 0) {
            if (env.has_event()) env.status |= CXXDES_B200_REP_UNFINISHED;
            env.now = env.now > sh.run_until ? env.now : sh.run_until;
        }
        rep_out o;
        o.end_time = env.now;
        o.n_events = env.n_events;
        o.trace_hash = env.hash;
        o.status = env.status;
        o.f64[0] = total_wait;
        o.f64[1] = 0.0;
        o.i64[0] = (int64_t)served;
        o.i64[1] = (int64_t)balked;
        store_result(out, rep, o);
    }
}

// ---------------------------------------------------------------------------------------------
// Compact kernels (the default): the same lowering on device/engine32.cuh, one flat loop with in-loop work stealing,
// 64-entry heap, 64 waiters as 16-bit process ordinals (every waiter is an acquire_proc: priority 0, latency 0).
//   wide tokens    8 bytes {u32 time, u32 target | handler << 16} (heap32x): 640 bytes of shared memory per
//                  replication instead of 1920, 352 threads resident per SM instead of 96;
//   narrow tokens  4 bytes {20 bits of time past a moving epoch, handler << 11 | target} (heap_rel<.., 12>): 384 bytes,
//                  576 threads per SM.  First choice when the ordinals fit 11 bits (at most 1022 customers) and the
//                  delays stay far below 2^20 - 2^18 ticks (mmc_narrow_applies).
// A replication that outgrows the capacities, the 32-bit clock or (narrow) the delay range is handed to mmc_kernel above.
// ---------------------------------------------------------------------------------------------
namespace {
constexpr int MF_HEAP = 64;
constexpr int MF_WAITERS = 64;
#ifndef CXB_PATH_PUSH
#define CXB_PATH_PUSH 1     // pushes as path-parallel sift-ups (engine32.cuh sift_up_path); 0: the loop (A/B)
#endif
#ifndef CXB_MMC_FLAT_POP
#define CXB_MMC_FLAT_POP 0
#endif
#ifndef CXB_MMC_PUSHES_PER_TURN
#define CXB_MMC_PUSHES_PER_TURN 2
#endif
constexpr int MF_PUSHES_PER_TURN = CXB_MMC_PUSHES_PER_TURN;

struct mmc_wide_tokens {
    static constexpr int THREADS = 176, BLOCKS_PER_SM = 2, HEAP = MF_HEAP, WAITERS = MF_WAITERS;
    using heap_t = heap32x<THREADS>;
    using word_t = uint2;
    static constexpr uint32_t HANDLER = 1u << 16, TARGET_MASK = 0xFFFFu, NOOP = TAG_NO_TARGET;
    static constexpr int COLUMN = HEAP;          // entries allocated per replication
    static constexpr bool FLAT_POP = false;
};
#ifndef CXB_MMC_NT          // threads, blocks per SM, heap entries, waiters of the narrow-token kernel (A/B builds)
#define CXB_MMC_NT 288
#define CXB_MMC_NB 2
#define CXB_MMC_NH 64
#define CXB_MMC_NW 64
#endif
constexpr int MF_NARROW[4] = {CXB_MMC_NT, CXB_MMC_NB, CXB_MMC_NH, CXB_MMC_NW};
struct mmc_narrow_tokens {
    static constexpr int THREADS = MF_NARROW[0], BLOCKS_PER_SM = MF_NARROW[1], HEAP = MF_NARROW[2], WAITERS = MF_NARROW[3];
    static_assert(HEAP <= MF_HEAP && WAITERS <= MF_WAITERS, "saved state is laid out for 64 + 64");
    using heap_t = heap_rel<THREADS, 12>;
    using word_t = uint32_t;
    static constexpr uint32_t HANDLER = 1u << 11, TARGET_MASK = 0x7FFu, NOOP = 0x7FFu;
    static constexpr bool FLAT_POP = CXB_MMC_FLAT_POP != 0;      // the pop as straight-line code (heap_rel::pop_after_top_flat)
    static constexpr int COLUMN = FLAT_POP ? heap_t::FLAT_POP_COLUMN(HEAP) : HEAP + heap_t::EXTRA;
};
static_assert(TAG_NO_TARGET == 0xFFFFu, "the wide tokens' no-target ordinal is the generic engine's");

struct wait_list16 {
    lane_array<uint16_t> who;
    int n;
    // wait_awaitable::await_suspend, event.hpp:136-139
    int cap;
    template <typename Env>
    __device__ __forceinline__ void wait(Env &env, uint32_t proc) {
        if (n >= cap) {
            env.status |= CXXDES_B200_REP_WAITER_OVERFLOW;
            return;
        }
        who[n++] = (uint16_t)proc;
    }
    // wake_awaitable::await_ready, event.hpp:125-134: schedule every waiter in list order
    template <typename Env>
    __device__ __forceinline__ void wake(Env &env) {
        for (int i = 0; i < n; ++i) env.schedule(env.now, (uint32_t)who[i]);
        n = 0;
    }
};
}  // namespace

namespace {
constexpr uint32_t MF_PHASE_SAVED = 1, MF_PHASE_DONE = 2, MF_PHASE_REPLAY = 3;   // word 0 of a saved replication
constexpr int MF_SAVED_HEADER = 16;
}  // namespace

// STATEFUL: run_until / run_for in pieces (environment.ipp:190-214).  A replication that stops at the horizon with
// tokens still scheduled is written to global memory, `[word][replication]`: header (clock, counters, semaphore),
// heap, wait list, and the words of the customers that a scheduled token or a waiter still refers to (nobody else
// can be touched again; at most heap + wait-list entries instead of 8 KB for all 1000); the next launch restores it.  DONE / REPLAY as
// in the other kernels.  A separate instantiation, so that the plain run() path carries none of it.
template <bool STATEFUL, typename TK>
__global__ void __launch_bounds__(TK::THREADS, TK::BLOCKS_PER_SM)
mmc_fast_kernel(cxxdes_b200_mmc_params prm, shard sh, device_columns out, uint64_t *cust_base, philox_schedule keys,
                uint32_t until, mmc_saved st) {
    extern __shared__ __align__(16) unsigned char smem[];
    log_entry *table = (log_entry *)smem;
    unsigned char *cursor = smem + sizeof(log_entry) * (1 << CXB_LOG_TABLE_BITS);
    stage_log_table(table, sh.log_table);
    constexpr uint32_t MF_HANDLER = TK::HANDLER, MF_NOOP = TK::NOOP;
    typename TK::word_t *heap = carve<typename TK::word_t>(cursor, TK::COLUMN).base;
    wait_list16 waiters;
    waiters.who = carve<uint16_t>(cursor, TK::WAITERS);
    waiters.cap = TK::WAITERS;
    waiters.n = 0;

    const uint64_t total_threads = (uint64_t)gridDim.x * blockDim.x;
    uint64_t *cust = cust_base + ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
    const uint32_t n = (uint32_t)prm.n_customers;
    const int64_t patience_ticks = real_to_sim(prm.patience, sh.clk);  // env.timeout(patience)
    const divisor d_lambda = make_divisor(prm.lambda), d_mu = make_divisor(prm.mu);

    typename TK::heap_t env;   // (every access to the heap goes through env: explicit ld/st.shared)
    env.init(heap, TK::HEAP);
    uint64_t rep = 0;
    uint32_t ctr2 = 0, ctr3 = 0, value = 0, i = 0, served = 0, balked = 0, main_pc = 0, source_pc = 0;
    double total_wait = 0.0;
    bool active = false, exhausted = false;
    // Tokens an event has scheduled and not yet pushed.  A release or an acquisition wakes EVERY waiter (event.hpp:
    // 125-134): up to ~30 pushes from one event, each a sift to the top of the heap.  Pushed inside the event, such a
    // storm made the other 31 lanes of the warp wait for it; now a lane pushes at most MF_PUSHES_PER_TURN tokens per
    // turn of the loop and simply does not take its next event until its list is empty, while its neighbours go on.
    uint32_t p_time[2] = {0, 0}, p_tag[2] = {0, 0};
    int n_wake = 0, n_total = 0, next_push = 0;

    auto draw = [&](uint32_t stream, uint32_t index, const divisor &rate) {
        const philox_block b = philox4x32_10_scheduled(index >> 1, stream, ctr2, ctr3, keys);
        const uint64_t bits = (index & 1) ? (((uint64_t)b.w[3] << 32) | b.w[2]) : (((uint64_t)b.w[1] << 32) | b.w[0]);
        return real_to_sim(exponential_from_bits(bits, rate, table), sh.clk);
    };

    for (;;) {
        if (!active && !exhausted) {
            if (steal(sh, rep)) {
                const uint32_t phase = STATEFUL && st.resume ? st.words[rep] : 0u;
                if (STATEFUL && phase == MF_PHASE_DONE) {   // now_ = max(now_, t), environment.ipp:194
                    if (sh.run_until >= 0 && out.end_time[rep] < sh.run_until) out.end_time[rep] = sh.run_until;
                } else if (STATEFUL && phase == MF_PHASE_REPLAY) {
                    const unsigned long long at = atomicAdd(sh.retry_count, 1ULL);
                    sh.retry_list[at] = (uint32_t)rep;
                } else {
                    const uint64_t grep = sh.first_replication + rep;
                    ctr2 = (uint32_t)grep;
                    ctr3 = (uint32_t)(grep >> 32);
                    env.init(heap, TK::HEAP);
                    if (STATEFUL && phase == MF_PHASE_SAVED) {
                        const uint32_t *sv = st.words + rep;
                        const size_t m = st.n_replications;
                        env.n = sv[1 * m];
                        env.now = sv[2 * m];
                        env.n_events = sv[3 * m];
                        env.hash = (uint64_t)sv[4 * m] | ((uint64_t)sv[5 * m] << 32);
                        value = sv[6 * m];
                        i = sv[7 * m];
                        served = sv[8 * m];
                        balked = sv[9 * m];
                        main_pc = sv[10 * m] & 0xFFu;
                        source_pc = sv[10 * m] >> 8;
                        total_wait = bits_as_double((uint64_t)sv[11 * m] | ((uint64_t)sv[12 * m] << 32));
                        waiters.n = (int)sv[13 * m];
                        int at = MF_SAVED_HEADER;
                        env.restart_epoch();     // (narrow tokens: times relative to the restored clock; every token is later)
                        for (uint32_t k = 0; k < env.n; ++k) env.set_token(k, sv[(size_t)(at + 2 * k) * m], sv[(size_t)(at + 2 * k + 1) * m]);
                        at += 2 * MF_HEAP;
                        for (int k = 0; k < waiters.n; ++k) waiters.who[k] = (uint16_t)sv[(size_t)(at + k) * m];
                        at += MF_WAITERS;
                        // customer words: only the customers a scheduled token or a waiter still refers to were saved
                        for (uint32_t c = 0; c < n; ++c) cust[(uint64_t)c * total_threads] = 0ull;
                        const uint32_t n_pairs = sv[14 * m];
                        for (uint32_t k = 0; k < n_pairs; ++k) {
                            const uint32_t id = sv[(size_t)(at + 3 * k) * m];
                            cust[(uint64_t)id * total_threads] =
                                (uint64_t)sv[(size_t)(at + 3 * k + 1) * m] | ((uint64_t)sv[(size_t)(at + 3 * k + 2) * m] << 32);
                        }
                    } else {
                        waiters.n = 0;
                        value = prm.servers;  // semaphore{c, c}
                        i = served = balked = 0;
                        main_pc = source_pc = 0;
                        total_wait = 0.0;
                        for (uint32_t c = 0; c < n; ++c) cust[(uint64_t)c * total_threads] = 0;
                        env.schedule(0, 0);  // environment::bind(co_main)
                    }
                    active = true;
                    if (STATEFUL && phase == MF_PHASE_SAVED && env.next_time() > until) {
                        // nothing of this replication falls into the piece: only its clock moves (environment.ipp:194)
                        if (sh.run_until > (int64_t)env.now) {
                            out.end_time[rep] = sh.run_until;
                            st.words[2 * st.n_replications + rep] = (uint32_t)sh.run_until;
                        }
                        active = false;
                    }
                }
            } else {
                exhausted = true;
            }
        }
        // (a claimed replication can leave the lane inactive -- finished in an earlier piece, handed to the
        // retry list, nothing due in this piece -- so the warp only leaves when every lane has seen the end of the work list)
        if (!__any_sync(0xFFFFFFFFu, active)) {
            if (__all_sync(0xFFFFFFFFu, exhausted)) break;
            continue;
        }

        if (active && next_push == n_total) {
            // The token at the root names the event: decode it before the heap moves, and start the load of the
            // customer word it refers to (global memory, L2 latency) so that the sift-down hides it.
            const typename TK::word_t root = env.entry(0);
            token32 t;
            t.time = env.time_of(root);
            t.tag = env.tag_of(root);
            const uint32_t target = t.tag & TK::TARGET_MASK;
            const bool is_customer_token = target >= 2 && target != MF_NOOP;   // customer(id), acquire_proc(id) or their handler
            uint64_t *const w = &cust[(uint64_t)(is_customer_token ? (target - 2) >> 1 : 0u) * total_threads];
            uint64_t s = *w;
            if constexpr (TK::FLAT_POP) env.template pop_after_top_flat<3, 0>();   // (64 entries: three two-level steps)
            else env.pop_after_top();            // environment::step, environment.ipp:117-126
            env.now = t.time > env.now ? t.time : env.now;
            ++env.n_events;
            // An event pushes at most two tokens of its own (wake-ups aside) and draws at most one variate.  The
            // branches below only note which; the variate and the pushes -- the sifts through the heap -- run after
            // them, in one piece of code for all lanes, in the order they were noted (push order decides ties).
            int n_push = 0;
            bool wakes = false;                  // event.wake(): every waiter is scheduled at now, BEFORE the event's own pushes
            bool draws = false;                  // the second push is at now + a variate
            uint32_t d_stream = 0, d_index = 0;
            auto push_at = [&](uint32_t time, uint32_t tag) {   // (constant indices: the two slots stay in registers)
                if (n_push == 0) { p_time[0] = time; p_tag[0] = tag; }
                else { p_time[1] = time; p_tag[1] = tag; }
                ++n_push;
            };
            auto push_after = [&](int64_t delay, uint32_t tag) {
                const uint64_t when = (uint64_t)env.now + (uint64_t)delay;
                if ((uint64_t)delay >= (1ULL << 31) || when >= 0xFFFFFFFFull) env.status |= CXXDES_B200_REP_TIME_RANGE;
                push_at((uint32_t)when, tag);
            };
            // digest of the popped token, once for all kinds of events (three copies in the branches ran one after the other)
            {
                const bool is_handler = (t.tag & MF_HANDLER) != 0, is_noop = !is_handler && target == MF_NOOP;
                env.fold(t.time, is_handler ? (uint32_t)TOKEN_HANDLER : (is_noop ? (uint32_t)TOKEN_NOOP : (uint32_t)TOKEN_RESUME),
                         is_noop ? (uint32_t)PROC_NONE : target);
            }
            if (t.tag & MF_HANDLER) {
                // any_of handler of customer (target - 2) / 2: any_of.ipp:9-26
                const uint64_t remaining = ((s & C_REM) >> C_REM_SHIFT) - 1;
                s = (s & ~C_REM) | (remaining << C_REM_SHIFT);
                if ((s & C_ARMED) && remaining < 2) {
                    push_at(t.time, target);  // output token
                    s &= ~C_ARMED;
                }
                *w = s;
            } else if (target == MF_NOOP) {
                // (nothing to resume)
            } else {
                if (target == 0) {
                    if (main_pc == 0) {
                        push_at(env.now, 1);  // co_await source(): start token; completion targets main
                        main_pc = 1;
                    } else {
                        push_at(env.now, MF_NOOP);
                    }
                } else if (target == 1) {
                    // source: for i < n { co_await async(customer(i)); co_await env.timeout(arr()); }
                    if (source_pc == 1) ++i;
                    if (i < n) {
                        push_at(env.now, 2 + 2 * i);  // async: start token (async.ipp:21-28)
                        draws = true;                 // then env.timeout(arr()): draw i of stream 1
                        d_stream = 1;
                        d_index = i;
                        p_tag[1] = 1;
                        source_pc = 1;
                    } else {
                        push_at(env.now, 0);  // source returns -> completion resumes co_main
                    }
                } else {
                    if ((target & 1) == 0) {
                        // ---- customer(id) ----
                        const uint64_t pc = s & C_PC;
                        if (pc == 0) {
                            // t0 = now_seconds(); any_of(acquire_proc, env.timeout(patience))
                            s = ((uint64_t)env.now << C_T0_SHIFT) | 1 | C_ARMED | (2ULL << C_REM_SHIFT);
                            push_at(env.now, target + 1);                        // await_bind: start(acquire_proc)
                            push_after(patience_ticks, target | MF_HANDLER);     // await_suspend: timeout + handler
                        } else if (pc == 1) {
                            if (!(s & C_ACQ_DONE)) {
                                s = (s & ~C_PC) | 3 | C_BALKED;  // balked: co_return
                                ++balked;
                                push_at(env.now, MF_NOOP);  // async's null-target completion
                            } else {
                                const double t0 = sim_to_seconds((int64_t)(s >> C_T0_SHIFT), sh.clk);
                                total_wait += sim_to_seconds((int64_t)env.now, sh.clk) - t0;
                                draws = true;              // env.timeout(svc()): draw 0 of the customer's own stream
                                d_stream = target;
                                d_index = 0;
                                p_tag[0] = target;
                                s = (s & ~C_PC) | 2;
                            }
                        } else {
                            // release: semaphore::up -> ++value; wake  (semaphore.hpp:58-66)
                            ++value;
                            wakes = true;
                            ++served;
                            s = (s & ~C_PC) | 3;
                            push_at(env.now, MF_NOOP);
                        }
                    } else {
                        // ---- acquire_proc(id): st.h = co_await servers.acquire(); if (balked) release ----
                        if (value == 0) {
                            waiters.wait(env, target);  // semaphore::down blocks (semaphore.hpp:70-78)
                        } else {
                            --value;
                            wakes = true;       // down() wakes every waiter even though it consumed a unit
                            if (s & C_BALKED) ++value;   // late acquisition is released at once: up() wakes again, but
                                                         // the first wake has emptied the list (event.hpp:125-134)
                            s |= C_ACQ_DONE;
                            push_at(env.now, (target - 1) | MF_HANDLER);  // completion token with the any_of handler
                        }
                    }
                    *w = s;
                }
            }
            // the one variate of the event, drawn where the lanes have met again (the source's inter-arrival time
            // and a customer's service time were two copies of Philox + log + divide in two different branches)
            if (draws) {
                const int64_t delay = draw(d_stream, d_index, d_stream == 1 ? d_lambda : d_mu);
                const uint64_t when = (uint64_t)env.now + (uint64_t)delay;
                if ((uint64_t)delay >= (1ULL << 31) || when >= 0xFFFFFFFFull) env.status |= CXXDES_B200_REP_TIME_RANGE;
                // (its tag was noted by the branch: the source's second push, a customer's only one)
                if (d_stream == 1) p_time[1] = (uint32_t)when;
                else p_time[0] = (uint32_t)when;
                ++n_push;
            }
            // what the event scheduled, in program order: the woken waiters in list order (wake_awaitable::await_ready,
            // event.hpp:125-134 -- both events that wake do so before their own push), then the event's own tokens
            n_wake = wakes ? waiters.n : 0;
            n_total = n_wake + n_push;
            next_push = 0;
        }
        if (active) {
            // environment::schedule_token: one copy of the sift for the whole warp, a few tokens per turn
#pragma unroll
            for (int turn = 0; turn < MF_PUSHES_PER_TURN; ++turn) {
                if (next_push < n_total) {
                    const int own = next_push - n_wake;
                    const uint32_t time = own < 0 ? env.now : (own == 0 ? p_time[0] : p_time[1]);
                    const uint32_t tag = own < 0 ? (uint32_t)waiters.who[next_push] : (own == 0 ? p_tag[0] : p_tag[1]);
#if CXB_PATH_PUSH
                    env.template schedule_path<6>(time, tag);     // 64 entries: at most 6 ancestors, loaded in one round trip
#else
                    env.schedule(time, tag);
#endif
                    ++next_push;
                }
            }
        }
        if (active && next_push == n_total) {
            if (n_wake) waiters.n = 0;     // the wake is complete: the list is empty (event.hpp:132)
            n_wake = n_total = next_push = 0;

            if (env.status || !env.has_event() || env.next_time() > until) {
                active = false;
                if (env.status & (CXXDES_B200_REP_TIME_RANGE | CXXDES_B200_REP_HEAP_OVERFLOW | CXXDES_B200_REP_WAITER_OVERFLOW)) {
                    const unsigned long long at = atomicAdd(sh.retry_count, 1ULL);
                    sh.retry_list[at] = (uint32_t)rep;
                    if (STATEFUL) st.words[rep] = MF_PHASE_REPLAY;
                } else {
                    rep_out o;
                    int64_t end = (int64_t)env.now;
                    o.status = env.status;
                    if (sh.run_until >= 0) {
                        if (env.has_event()) o.status |= CXXDES_B200_REP_UNFINISHED;
                        end = end > sh.run_until ? end : sh.run_until;
                    }
                    if (STATEFUL) {
                        uint32_t *sv = st.words + rep;
                        const size_t m = st.n_replications;
                        if (o.status == CXXDES_B200_REP_UNFINISHED) {
                            sv[1 * m] = env.n;
                            sv[2 * m] = (uint32_t)end;      // run_until leaves now_ at the horizon
                            sv[3 * m] = env.n_events;
                            sv[4 * m] = (uint32_t)env.hash;
                            sv[5 * m] = (uint32_t)(env.hash >> 32);
                            sv[6 * m] = value;
                            sv[7 * m] = i;
                            sv[8 * m] = served;
                            sv[9 * m] = balked;
                            sv[10 * m] = main_pc | (source_pc << 8);
                            const uint64_t tw = (uint64_t)__double_as_longlong(total_wait);
                            sv[11 * m] = (uint32_t)tw;
                            sv[12 * m] = (uint32_t)(tw >> 32);
                            sv[13 * m] = (uint32_t)waiters.n;
                            int at = MF_SAVED_HEADER;
                            for (uint32_t k = 0; k < env.n; ++k) {
                                const token32 e = env.token_at(k);
                                sv[(size_t)(at + 2 * k) * m] = e.time;
                                sv[(size_t)(at + 2 * k + 1) * m] = e.tag;
                            }
                            at += 2 * MF_HEAP;
                            for (int k = 0; k < waiters.n; ++k) sv[(size_t)(at + k) * m] = waiters.who[k];
                            at += MF_WAITERS;
                            // Only a customer that a scheduled token or a waiter refers to can be touched again
                            // (targets >= 2 name customer (target - 2) / 2 or its acquire_proc); the customers still
                            // to come start from zero.  At most heap + wait-list entries, against 8 KB for them all.
                            uint32_t n_pairs = 0;
                            auto save_customer = [&](uint32_t target) {
                                if (target < 2 || target == MF_NOOP) return;
                                const uint32_t id = (target - 2) >> 1;
                                const uint64_t wv = cust[(uint64_t)id * total_threads];
                                sv[(size_t)(at + 3 * n_pairs) * m] = id;
                                sv[(size_t)(at + 3 * n_pairs + 1) * m] = (uint32_t)wv;
                                sv[(size_t)(at + 3 * n_pairs + 2) * m] = (uint32_t)(wv >> 32);
                                ++n_pairs;
                            };
                            for (uint32_t k = 0; k < env.n; ++k) save_customer(env.token_at(k).tag & TK::TARGET_MASK);
                            for (int k = 0; k < waiters.n; ++k) save_customer(waiters.who[k]);
                            sv[14 * m] = n_pairs;
                            sv[0] = MF_PHASE_SAVED;
                        } else {
                            sv[0] = MF_PHASE_DONE;
                        }
                    }
                    o.end_time = end;
                    o.n_events = env.n_events;
                    o.trace_hash = env.hash;
                    o.f64[0] = total_wait;
                    o.f64[1] = 0.0;
                    o.i64[0] = (int64_t)served;
                    o.i64[1] = (int64_t)balked;
                    store_result(out, rep, o);
                }
            }
        }
        __syncwarp();
    }
}

size_t mmc_deep_smem_bytes() { return sizeof(log_entry) * (1 << CXB_LOG_TABLE_BITS) + 64; }

size_t mmc_smem_bytes() {
    return sizeof(log_entry) * (1 << CXB_LOG_TABLE_BITS) + 64 +
           (size_t)MMC_THREADS * (MMC_HEAP * (8 + 4) + MMC_WAITERS * (4 + 4));
}

namespace {
template <typename TK>
size_t mmc_fast_smem_bytes() {
    return sizeof(log_entry) * (1 << CXB_LOG_TABLE_BITS) + 64 +
           (size_t)TK::THREADS * (TK::COLUMN * sizeof(typename TK::word_t) + TK::WAITERS * 2);
}
// 32-bit event counter and clock: a variate is at most 36.05 / rate
bool mmc_fast_applies(const cxxdes_b200_mmc_params &p, const shard &sh) {
    const double tpu = (double)sh.clk.ticks_per_unit;
    const double slow = p.lambda < p.mu ? p.lambda : p.mu;
    return 36.05 * tpu / slow < 1073741824.0 && p.patience * tpu < 1073741824.0 && p.n_customers < 100000000ull &&
           p.servers < 0xFFFFFFFFull;
}

template <typename TK>
cudaError_t launch_mmc_fast(const cxxdes_b200_mmc_params &prm, const shard &sh, const device_columns &out,
                            uint64_t *customer_state, int sm_count, cudaStream_t stream, const mmc_saved &st) {
    const size_t fsmem = mmc_fast_smem_bytes<TK>();
    cudaError_t e = cudaFuncSetAttribute(mmc_fast_kernel<false, TK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mmc_fast_kernel<true, TK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem);
    if (e != cudaSuccess) return e;
    const uint32_t until = (sh.run_until < 0 || sh.run_until >= 0xFFFFFFFFll) ? 0xFFFFFFFFu : (uint32_t)sh.run_until;
    const philox_schedule keys = philox_schedule_of((uint32_t)sh.seed, (uint32_t)(sh.seed >> 32));
    const int blocks = sm_count * TK::BLOCKS_PER_SM;
    if (st.words) mmc_fast_kernel<true, TK><<<blocks, TK::THREADS, fsmem, stream>>>(prm, sh, out, customer_state, keys, until, st);
    else mmc_fast_kernel<false, TK><<<blocks, TK::THREADS, fsmem, stream>>>(prm, sh, out, customer_state, keys, until, st);
    return cudaGetLastError();
}
}  // namespace

// Narrow tokens: ordinals up to 3 + 2 (n - 1) must stay below the no-target code 2047; the patience always fits or
// never does; a service or inter-arrival time is an exponential variate, so the kernel is chosen when one passes the
// delay limit with probability below e^-18 per draw (as aloha_compact_applies) and the replication that draws one is
// redone by the 64-bit kernel.  (CXXDES_B200_MMC_TAIL: the exponent, for the tests of that hand-over.)
bool mmc_narrow_applies(const cxxdes_b200_mmc_params &p, const shard &sh) {
    if (!mmc_fast_applies(p, sh) || p.n_customers > 1022) return false;
    const double tpu = (double)sh.clk.ticks_per_unit, limit = (double)mmc_narrow_tokens::heap_t::DELAY_LIMIT;
    const double slow = p.lambda < p.mu ? p.lambda : p.mu;
    double tail = 18.0;
    if (const char *text = getenv("CXXDES_B200_MMC_TAIL")) tail = atof(text) > 0.0 ? atof(text) : tail;
    return tail * tpu / slow < limit && p.patience * tpu < limit;
}

// per-customer words are indexed [customer][global thread]: sized for the largest of the grids
int mmc_blocks(int sm_count) { return sm_count * mmc_narrow_tokens::BLOCKS_PER_SM; }
int mmc_threads() { return mmc_narrow_tokens::THREADS; }
static_assert(mmc_narrow_tokens::THREADS * mmc_narrow_tokens::BLOCKS_PER_SM >= mmc_wide_tokens::THREADS * mmc_wide_tokens::BLOCKS_PER_SM &&
              mmc_narrow_tokens::THREADS * mmc_narrow_tokens::BLOCKS_PER_SM >= MMC_THREADS, "customer words are sized for the narrow-token grid");

bool mmc_keeps_state(const cxxdes_b200_mmc_params &prm, const shard &sh) { return mmc_fast_applies(prm, sh); }
size_t mmc_saved_words(const cxxdes_b200_mmc_params &prm) {
    return MF_SAVED_HEADER + 2 * (size_t)MF_HEAP + MF_WAITERS + 3 * (size_t)(MF_HEAP + MF_WAITERS);
}

cudaError_t launch_mmc(const cxxdes_b200_mmc_params &prm, const shard &sh, const device_columns &out,
                       uint64_t *customer_state, int sm_count, unsigned long long *list_cursor, cudaStream_t stream,
                       int *launches, const mmc_saved &st, bool wide_tokens, const mmc_deep &deep) {
    const size_t smem = mmc_smem_bytes();
    cudaError_t e = cudaFuncSetAttribute(mmc_kernel<environment>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (mmc_fast_applies(prm, sh)) {
        if (!wide_tokens && mmc_narrow_applies(prm, sh)) e = launch_mmc_fast<mmc_narrow_tokens>(prm, sh, out, customer_state, sm_count, stream, st);
        else e = launch_mmc_fast<mmc_wide_tokens>(prm, sh, out, customer_state, sm_count, stream, st);
        ++*launches;
        if (e != cudaSuccess) return e;
        mmc_kernel<environment><<<sm_count, MMC_THREADS, smem, stream>>>(prm, sh, out, customer_state, sh.retry_list, sh.retry_count,
                                                                        list_cursor, trace_sink{}, mmc_deep{});
        ++*launches;
    } else {
        mmc_kernel<environment><<<sm_count, MMC_THREADS, smem, stream>>>(prm, sh, out, customer_state, nullptr, nullptr, nullptr,
                                                                        trace_sink{}, mmc_deep{});
        ++*launches;
    }
    e = cudaGetLastError();
    if (e != cudaSuccess || !deep.base) return e;
    // what outgrew the 96-entry heap / wait list: once more with both in global memory (exits at once when nothing did)
    e = cudaFuncSetAttribute(mmc_kernel<environment, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mmc_deep_smem_bytes());
    if (e != cudaSuccess) return e;
    const int deep_block = deep.threads < MMC_THREADS ? deep.threads : MMC_THREADS;
    mmc_kernel<environment, true><<<deep.threads / deep_block, deep_block, mmc_deep_smem_bytes(), stream>>>(
        prm, sh, out, customer_state, nullptr, sh.retry_a_count, sh.retry_a_cursor, trace_sink{}, deep);
    ++*launches;
    return cudaGetLastError();
}

// Worst case of one replication (n customers): a process is the target of at most one token it waits for, a customer
// owns at most two handler tokens (its any_of's children) and one output token, so 6 n + 16 entries hold any heap the
// model can build; at most one acquire_proc per customer waits.  Threads: as many as `budget` bytes hold, a whole
// number of blocks, at most one 96-thread block per SM.  Returns the bytes to allocate (0: no deep pass).
size_t mmc_deep_plan(const cxxdes_b200_mmc_params &prm, int sm_count, size_t budget, mmc_deep &d) {
    d = mmc_deep{};
    const size_t n = prm.n_customers ? prm.n_customers : 1;
    const size_t heap_cap = 6 * n + 16, waiter_cap = n + 8;
    const size_t per_thread = heap_cap * (8 + 4) + waiter_cap * (4 + 4);
    size_t threads = budget / per_thread;
    if (threads == 0) return 0;
    if (threads > (size_t)sm_count * MMC_THREADS) threads = (size_t)sm_count * MMC_THREADS;
    if (threads > (size_t)MMC_THREADS) threads -= threads % MMC_THREADS;
    d.threads = (int)threads;
    d.heap_cap = (int)heap_cap;
    d.waiter_cap = (int)waiter_cap;
    return threads * per_thread;
}

// One replication (sh.n_replications == 1) on the full-size kernel, with its event trace.  `customer_state` needs
// room for one block of the kernel (the handle's buffer always has).
cudaError_t launch_mmc_trace(const cxxdes_b200_mmc_params &prm, const shard &sh, const device_columns &out,
                             uint64_t *customer_state, const trace_sink &sink, cudaStream_t stream, const mmc_deep &deep) {
    if (deep.base) {   // heap and wait list in global memory: any replication the ensemble's passes can run can be traced
        cudaError_t e = cudaFuncSetAttribute(mmc_kernel<traced_environment, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)mmc_deep_smem_bytes());
        if (e != cudaSuccess) return e;
        mmc_kernel<traced_environment, true><<<1, deep.threads < MMC_THREADS ? deep.threads : MMC_THREADS, mmc_deep_smem_bytes(), stream>>>(
            prm, sh, out, customer_state, nullptr, nullptr, nullptr, sink, deep);
        return cudaGetLastError();
    }
    const size_t smem = mmc_smem_bytes();
    cudaError_t e = cudaFuncSetAttribute(mmc_kernel<traced_environment>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    mmc_kernel<traced_environment><<<1, MMC_THREADS, smem, stream>>>(prm, sh, out, customer_state, nullptr, nullptr, nullptr, sink,
                                                                    mmc_deep{});
    return cudaGetLastError();
}

}  // namespace cxb
