// models_aloha.cu -- pure ALOHA (examples/aloha.cpp:25-85) on the device engine.
//
// One thread per replication.  The event heap (up to N + 2 tokens, real key ties: 0.3-0.5 % of
// pops) lives in shared memory and performs libstdc++'s exact __push_heap / __adjust_heap moves
// (device/engine.cuh), because tie order among equal (time, priority) keys is observable in the
// trace digest.  Station state is bit-packed in registers: the std::set<bool*> of active
// transmissions is a 64-bit mask, the per-iteration `collided` flags another, the resume point
// of each station two more; the loop counters sit in shared memory next to the heap.
//
// Lowering, processes: co_main = 0, station s = 1 + s, all priorities 0:
//   co_main   all_of.range(stations): bind every station in index order (start tokens), one
//             shared all_of handler on their completion tokens (any_of.ipp:41-84,197-209)
//   station   for i < pps: { collided = false; insert; check_collisions; timeout(frame_time);
//                            erase; if (!collided) ++succ; timeout(exponential()) }
#include <cstdlib>

#include "device/engine32.cuh"
#include "device/models_common.cuh"
#include "device/trace.cuh"
#include "kernels.h"

namespace cxb {

namespace {

constexpr int ALOHA_THREADS = 128;

struct aloha_rep_params {
    float lambda;
    uint64_t pps;
};

// main(), aloha.cpp:96,101: float arithmetic, size_t operands converted to float.
__device__ __forceinline__ aloha_rep_params aloha_params_of(const cxxdes_b200_aloha_params &p, uint64_t grep) {
    aloha_rep_params r;
    if (p.sweep_steps > 1) {
        const uint64_t num_steps = p.sweep_steps;
        const uint64_t step = grep % num_steps;
        const float begin = p.lambda_begin, end = p.lambda_end;
        const float num = (end - begin) * (float)step;
        r.lambda = begin + num / (float)(num_steps - 1);
        r.pps = (uint64_t)(r.lambda * p.packets_scale);
    } else {
        r.lambda = p.lambda_begin;
        r.pps = p.packets_per_station;
    }
    return r;
}

}  // namespace

// Generic-engine kernel: 64-bit keys, any parameter range.  Runs the whole shard, or (list != nullptr)
// only the replications the compact kernel below handed back.
// ENV = traced_environment: the same kernel writing the event trace of its replications (cxxdes_b200_trace).
template <typename ENV>
__global__ void __launch_bounds__(ALOHA_THREADS)
aloha_kernel(cxxdes_b200_aloha_params prm, shard sh, device_columns out, const uint32_t *list,
             const unsigned long long *list_count, unsigned long long *list_cursor, trace_sink sink) {
    extern __shared__ __align__(16) unsigned char smem[];
    log_entry *table = (log_entry *)smem;
    unsigned char *cursor = smem + sizeof(log_entry) * (1 << CXB_LOG_TABLE_BITS);
    stage_log_table(table, sh.log_table);
    const int N = (int)prm.num_stations;
    const int heap_cap = N + 2;
    lane_array<uint64_t> hkey = carve<uint64_t>(cursor, heap_cap);
    lane_array<uint32_t> htag = carve<uint32_t>(cursor, heap_cap);
    lane_array<uint32_t> loop_i = carve<uint32_t>(cursor, N);

    const int64_t frame_ticks = real_to_sim_f32(prm.frame_time, sh.clk);  // env.timeout(float), aloha.cpp:60
    const uint32_t HANDLER_TAG = make_tag(0, 1);

    uint64_t rep;
    for (;;) {
        if (list) {
            const unsigned long long idx = atomicAdd(list_cursor, 1ULL);
            if (idx >= *list_count) break;
            rep = list[idx];
        } else if (!steal(sh, rep)) {
            break;
        }
        const uint64_t grep = sh.first_replication + rep;
        const aloha_rep_params rp = aloha_params_of(prm, grep);
        // exponential_rv{seed, cfg_.lambda / cfg_.num_stations}: float / size_t -> float -> double (:52)
        const divisor rate = make_divisor((double)(rp.lambda / (float)prm.num_stations));
        stream_addr addr;
        addr.seed = sh.seed;
        addr.replication = grep;

        ENV env;
        env.init(hkey, htag, heap_cap);
        attach_trace(env, sink);
        uint64_t active = 0, collided = 0, pc_tx = 0, pc_idle = 0;  // pc_tx: awaiting the frame timeout; pc_idle: awaiting the inter-arrival timeout
        uint64_t successful = 0;
        uint32_t main_pc = 0;
        any_all all;
        float s_result = 0.0f, g_result = 0.0f;
        for (int s = 0; s < N; ++s) loop_i[s] = 0;

        // environment::bind(co_main): start token + null-target completion (coroutine.ipp:352-357)
        env.schedule(0, 0, make_tag(0));

        while (env.has_event() && (sh.run_until < 0 || env.next_time() <= sh.run_until)) {
            token t = env.step_pop();
            if (tag_handler_plus1(t.tag)) {
                all.invoke(env, t);
            } else {
                const uint32_t target = tag_target(t.tag);
                if (target == 0) {
                    if (main_pc == 0) {
                        // co_main, aloha.cpp:39-44: bind stations in index order, one all_of handler
                        for (int s = 0; s < N; ++s) env.schedule(env.now, 0, make_tag(1 + s));
                        all.arm(true, (uint32_t)N, (uint32_t)N);
                        main_pc = 1;
                    } else {
                        // :46-47  S = succ / now_seconds() (size_t / double -> float), G = lambda * frame_time
                        s_result = (float)((double)successful / sim_to_seconds(env.now, sh.clk));
                        g_result = rp.lambda * prm.frame_time;
                        env.schedule(env.now, 0, make_tag(TAG_NO_TARGET));  // co_main returns
                    }
                } else if (target != TAG_NO_TARGET) {
                    // station, aloha.cpp:51-71
                    const int s = (int)target - 1;
                    const uint64_t bit = 1ULL << s;
                    if (pc_tx & bit) {
                        // frame finished: erase, count, draw the waiting time (:62-69)
                        pc_tx &= ~bit;
                        active &= ~bit;
                        if (!(collided & bit)) ++successful;
                        addr.stream = target;
                        const double wait = exponential(addr, (uint64_t)loop_i[s], rate, table);
                        env.schedule(env.now + real_to_sim(wait, sh.clk), 0, make_tag(target));
                        pc_idle |= bit;
                    } else {
                        uint32_t i = loop_i[s];
                        if (pc_idle & bit) {  // the for loop's ++i after the inter-arrival timeout
                            pc_idle &= ~bit;
                            loop_i[s] = ++i;
                        }
                        if (i < rp.pps) {
                            // collided = false; insert; check_collisions (:55-59,73-78)
                            collided &= ~bit;
                            active |= bit;
                            if (__popcll(active) > 1) collided |= active;
                            env.schedule(env.now + frame_ticks, 0, make_tag(target));
                            pc_tx |= bit;
                        } else {
                            env.schedule(env.now, 0, HANDLER_TAG);  // co_return -> completion token with the all_of handler
                        }
                    }
                }
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
        o.f64[0] = (double)s_result;
        o.f64[1] = (double)g_result;
        o.i64[0] = (int64_t)successful;
        o.i64[1] = (int64_t)rp.pps;
        store_result(out, rep, o);
    }
}

// ---------------------------------------------------------------------------------------------
// Compact kernel (the default): the same lowering on device/engine32.cuh.
//   * tokens are {u32 time, u32 tag}; the tag's low 7 bits name the target (0 co_main, 1..64
//     station, 65 the all_of handler, 66 a null-target completion) and, because every station
//     has exactly one token scheduled at any time, its loop counter i rides in the upper 25
//     bits -- no per-station arrays at all: 8 bytes x (N + 2) per replication, 384 resident
//     threads per SM instead of 128;
//   * one flat loop, one event per lane per iteration, and a lane whose replication ends takes
//     the next one at once: replications of the load sweep differ 30x in length, a nested
//     "while (events)" loop leaves the warp waiting for its longest lane;
//   * Philox round keys come precomputed from the constant bank (the key is the seed).
// Replications whose clock would pass 2^32 ticks are handed to aloha_kernel above.
// ---------------------------------------------------------------------------------------------
#ifndef CXB_ALOHA_FLAT_POP
#define CXB_ALOHA_FLAT_POP 1   // the pop of the 4-byte-token kernel as straight-line code (heap_rel::pop_after_top_flat: +3 %); 0: the loop (A/B)
#endif
#ifndef CXB_PATH_PUSH
#define CXB_PATH_PUSH 1     // the event's push as a path-parallel sift-up (engine32.cuh sift_up_path); 0: the loop (A/B)
#endif

namespace {
constexpr int AF_THREADS = 128;
constexpr int AF_BLOCKS_PER_SM = 3;
constexpr int AC_BLOCKS_PER_SM = 4;
constexpr uint32_t CODE_MAIN = 0, CODE_HANDLER = 65, CODE_NOOP = 66, CODE_MASK = 127;
constexpr uint32_t AF_PHASE_SAVED = 1, AF_PHASE_DONE = 2, AF_PHASE_REPLAY = 3;   // word 0 of a saved replication
constexpr int COUNTER_SHIFT = 7;
}  // namespace

// STATEFUL: run_until / run_for in pieces (environment.ipp:190-214).  A replication that stops at the horizon with
// tokens still scheduled is written to global memory, `[word][replication]`: 16 header words {phase, heap size,
// clock, event count, digest, the three station masks, counters} and its live heap entries; the next launch
// restores it.  Finished replications keep phase DONE (their end_time follows later horizons); replications the
// 32-bit kernel handed to aloha_kernel get phase REPLAY and are replayed from tick 0 on every later piece.  A
// separate instantiation, so that the plain run() path carries none of it.
template <bool STATEFUL>
__global__ void __launch_bounds__(AF_THREADS, AF_BLOCKS_PER_SM)
aloha_fast_kernel(cxxdes_b200_aloha_params prm, shard sh, device_columns out, philox_schedule keys, uint32_t until,
                  aloha_saved st) {
    extern __shared__ __align__(16) unsigned char smem[];
    log_entry *table = (log_entry *)smem;
    unsigned char *cursor = smem + sizeof(log_entry) * (1 << CXB_LOG_TABLE_BITS);
    stage_log_table(table, sh.log_table);
    const int N = (int)prm.num_stations;
    const int heap_cap = N + 2;
    uint2 *heap = environment32<AF_THREADS>::carve_heap(cursor, heap_cap);

    const int64_t frame_ticks = real_to_sim_f32(prm.frame_time, sh.clk);  // env.timeout(float), aloha.cpp:60

    heap32x<AF_THREADS> env;   // (every access to the heap goes through env: explicit ld/st.shared)
    env.init(heap, heap_cap);
    uint64_t rep = 0, active_set = 0, collided = 0, pc_tx = 0;
    uint32_t successful = 0, main_pc = 0, remaining = 0, ctr2 = 0, ctr3 = 0;
    bool armed = false;
    aloha_rep_params rp{};
    divisor rate{};
    float s_result = 0.0f, g_result = 0.0f;
    bool active = false, exhausted = false;

    // Longest first: the replications of a load sweep differ 30x in length (packets per station = G * 1000,
    // aloha.cpp:101), and a lane that draws a long one near the end of the work list keeps its whole warp -- 31 idle
    // lanes -- running for up to a third of the launch.  So the work list is walked from the heaviest sweep step down:
    // claim k stands for member k % per_step of step (S - 1 - k / per_step); the tail of the launch is then made of
    // the shortest replications.  (Which replication a lane runs never changes a result: the Philox key is the global id.)
    const uint32_t S = prm.sweep_steps > 1 ? prm.sweep_steps : 1u;
    const uint64_t per_step = (sh.n_replications + S - 1) / S;
    const uint64_t claim_limit = per_step * S;
    const bool descending = prm.lambda_end >= prm.lambda_begin;
    const uint32_t first_step = (uint32_t)(sh.first_replication % S);

    for (;;) {
        if (!active && !exhausted) {
            const unsigned long long claim = atomicAdd(sh.work_counter, 1ULL);
            bool claimed = claim < claim_limit;
            if (claimed) {
                const uint32_t rank = (uint32_t)(claim / per_step);
                const uint32_t step = descending ? S - 1 - rank : rank;
                rep = (uint64_t)((step + S - first_step) % S) + (claim % per_step) * S;
                claimed = rep < sh.n_replications;     // (the steps do not all have per_step members)
            } else {
                exhausted = true;
            }
            if (claimed) {
                const uint32_t phase = STATEFUL && st.resume ? st.words[rep] : 0u;
                if (STATEFUL && phase == AF_PHASE_DONE) {   // now_ = max(now_, t), environment.ipp:194
                    if (sh.run_until >= 0 && out.end_time[rep] < sh.run_until) out.end_time[rep] = sh.run_until;
                } else if (STATEFUL && phase == AF_PHASE_REPLAY) {
                    const unsigned long long at = atomicAdd(sh.retry_count, 1ULL);
                    sh.retry_list[at] = (uint32_t)rep;
                } else {
                    const uint64_t grep = sh.first_replication + rep;
                    rp = aloha_params_of(prm, grep);
                    // exponential_rv{seed, cfg_.lambda / cfg_.num_stations}: float / size_t -> float -> double (:52)
                    rate = make_divisor((double)(rp.lambda / (float)prm.num_stations));
                    ctr2 = (uint32_t)grep;
                    ctr3 = (uint32_t)(grep >> 32);
                    env.init(heap, heap_cap);
                    s_result = g_result = 0.0f;
                    if (STATEFUL && phase == AF_PHASE_SAVED) {
                        const uint32_t *sv = st.words + rep;
                        const size_t n = st.n_replications;
                        env.n = sv[1 * n];
                        env.now = sv[2 * n];
                        env.n_events = sv[3 * n];
                        env.hash = (uint64_t)sv[4 * n] | ((uint64_t)sv[5 * n] << 32);
                        active_set = (uint64_t)sv[6 * n] | ((uint64_t)sv[7 * n] << 32);
                        collided = (uint64_t)sv[8 * n] | ((uint64_t)sv[9 * n] << 32);
                        pc_tx = (uint64_t)sv[10 * n] | ((uint64_t)sv[11 * n] << 32);
                        successful = sv[12 * n];
                        main_pc = sv[13 * n] & 0xFFu;
                        armed = (sv[13 * n] >> 8) != 0;
                        remaining = sv[14 * n];
                        for (uint32_t k = 0; k < env.n; ++k)
                            env.set_entry(k, make_uint2(sv[(size_t)(16 + 2 * k) * n], sv[(size_t)(17 + 2 * k) * n]));
                    } else {
                        active_set = collided = pc_tx = 0;
                        successful = 0; main_pc = 0; remaining = 0; armed = false;
                        env.schedule(0, CODE_MAIN);  // environment::bind(co_main)
                    }
                    active = true;
                    if (STATEFUL && phase == AF_PHASE_SAVED && env.next_time() > until) {
                        // nothing of this replication falls into the piece: only its clock moves (environment.ipp:194)
                        if (sh.run_until > (int64_t)env.now) {
                            out.end_time[rep] = sh.run_until;
                            st.words[2 * st.n_replications + rep] = (uint32_t)sh.run_until;
                        }
                        active = false;
                    }
                }
            }
        }
        // (a claimed replication can leave the lane inactive -- finished in an earlier piece, handed to the
        // retry list, nothing due in this piece -- so the warp only leaves when every lane has seen the end of the work list)
        if (!__any_sync(0xFFFFFFFFu, active)) {
            if (__all_sync(0xFFFFFFFFu, exhausted)) break;
            continue;
        }

        if (active) {
            // ---- one event, branch-light: the token at the root decides everything the event does, so it is decoded
            // first; the variate a station draws when its frame ends is computed for EVERY event (half of them use
            // it: a branch would run it for half of the lanes at the same cost and serialise it with the heap moves
            // it does not depend on); the event's effects are selects.  The sifts stay loops: a sift-down has the
            // same trip count in every lane (the heaps hold N or N + 1 tokens), and the predicated fixed-depth
            // sift-up that was tried (engine32.cuh schedule_flat) costs 78 instructions for every lane where the loop
            // costs ~7 per level actually climbed.
            const uint2 root = env.entry(0);
            const uint32_t code = root.y & CODE_MASK;
            const uint32_t i = root.y >> COUNTER_SHIFT;
            const philox_block b = philox4x32_10_scheduled(i >> 1, code, ctr2, ctr3, keys);
            const uint64_t bits = (i & 1) ? (((uint64_t)b.w[3] << 32) | b.w[2]) : (((uint64_t)b.w[1] << 32) | b.w[0]);
            const int64_t wait_ticks = real_to_sim(exponential_from_bits(bits, rate, table), sh.clk);

            env.pop_after_top();                                   // environment::step, environment.ipp:117-126
            token32 t;
            t.time = root.x;
            t.tag = root.y;
            env.now = t.time > env.now ? t.time : env.now;
            ++env.n_events;

            const bool is_station = code - 1u < 64u;
            const bool is_handler = code == CODE_HANDLER, is_main = code == CODE_MAIN;
            const uint64_t bit = is_station ? 1ULL << ((code - 1u) & 63u) : 0ULL;
            // station, aloha.cpp:51-71
            const bool tx_end = (pc_tx & bit) != 0;                          // frame finished: erase, count, draw (:62-69)
            const bool tx_start = is_station && !tx_end && i < rp.pps;       // collided = false; insert; check (:55-59,73-78)
            const bool st_return = is_station && !tx_end && !(i < rp.pps);   // co_return -> completion with the all_of handler
            successful += (tx_end && !(collided & bit)) ? 1u : 0u;
            pc_tx = tx_end ? pc_tx & ~bit : (tx_start ? pc_tx | bit : pc_tx);
            active_set = tx_end ? active_set & ~bit : (tx_start ? active_set | bit : active_set);
            collided = tx_start ? collided & ~bit : collided;
            collided |= (tx_start && __popcll(active_set) > 1) ? active_set : 0ULL;
            // all_of handler (any_of.ipp:9-26): the output token inherits the child token's time (== now)
            remaining -= is_handler ? 1u : 0u;
            const bool fire = is_handler && armed && remaining == 0;
            armed = armed && !fire;
            // digest: kind and process of the popped token
            const uint32_t kind = is_handler ? (uint32_t)TOKEN_HANDLER : ((is_station || is_main) ? (uint32_t)TOKEN_RESUME : (uint32_t)TOKEN_NOOP);
            const uint32_t proc = is_station ? code : ((is_handler || is_main) ? 0u : (uint32_t)PROC_NONE);
            env.fold(t.time, kind, proc);

            bool push = is_station || fire;
            int64_t delay = tx_end ? wait_ticks : (tx_start ? frame_ticks : 0);
            uint32_t push_tag = tx_end ? (code | ((i + 1) << COUNTER_SHIFT)) : (tx_start ? root.y : (st_return ? CODE_HANDLER : CODE_MAIN));
            if (is_main) {   // twice per replication
                if (main_pc == 0) {
                    // co_main, aloha.cpp:39-44: bind stations in index order, one all_of handler
                    for (int s = 0; s < N; ++s) env.schedule(env.now, (uint32_t)(1 + s));
                    remaining = (uint32_t)N;
                    armed = true;
                    main_pc = 1;
                } else {
                    // :46-47  S = succ / now_seconds() (size_t / double -> float), G = lambda * frame_time
                    s_result = (float)((double)successful / sim_to_seconds((int64_t)env.now, sh.clk));
                    g_result = rp.lambda * prm.frame_time;
                    push = true;             // co_main returns: null-target completion
                    push_tag = CODE_NOOP;
                }
            }
            // token at now + delay; the sum must stay inside 32 bits
            const uint64_t when = (uint64_t)env.now + (uint64_t)delay;
            if (push && ((uint64_t)delay >= (1ULL << 31) || when >= 0xFFFFFFFFull)) env.status |= CXXDES_B200_REP_TIME_RANGE;
#if CXB_PATH_PUSH
            if (push) env.schedule_path<6>((uint32_t)when, push_tag);   // every ancestor of the new leaf in one round trip
#else
            if (push) env.schedule((uint32_t)when, push_tag);
#endif

            // environment::run / run_until (environment.ipp:179-196)
            if (env.status || !env.has_event() || env.next_time() > until) {
                active = false;
                if (env.status & CXXDES_B200_REP_TIME_RANGE) {
                    const unsigned long long at = atomicAdd(sh.retry_count, 1ULL);
                    sh.retry_list[at] = (uint32_t)rep;
                    if (STATEFUL) st.words[rep] = AF_PHASE_REPLAY;
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
                        const size_t n = st.n_replications;
                        if (o.status & CXXDES_B200_REP_UNFINISHED) {
                            sv[1 * n] = env.n;
                            sv[2 * n] = (uint32_t)end;      // run_until leaves now_ at the horizon
                            sv[3 * n] = env.n_events;
                            sv[4 * n] = (uint32_t)env.hash;
                            sv[5 * n] = (uint32_t)(env.hash >> 32);
                            sv[6 * n] = (uint32_t)active_set;
                            sv[7 * n] = (uint32_t)(active_set >> 32);
                            sv[8 * n] = (uint32_t)collided;
                            sv[9 * n] = (uint32_t)(collided >> 32);
                            sv[10 * n] = (uint32_t)pc_tx;
                            sv[11 * n] = (uint32_t)(pc_tx >> 32);
                            sv[12 * n] = successful;
                            sv[13 * n] = main_pc | ((armed ? 1u : 0u) << 8);
                            sv[14 * n] = remaining;
                            for (uint32_t k = 0; k < env.n; ++k) {
                                const uint2 e = env.entry(k);
                                sv[(size_t)(16 + 2 * k) * n] = e.x;
                                sv[(size_t)(17 + 2 * k) * n] = e.y;
                            }
                            sv[0] = AF_PHASE_SAVED;
                        } else {
                            sv[0] = AF_PHASE_DONE;
                        }
                    }
                    o.end_time = end;
                    o.n_events = env.n_events;
                    o.trace_hash = env.hash;
                    o.f64[0] = (double)s_result;
                    o.f64[1] = (double)g_result;
                    o.i64[0] = (int64_t)successful;
                    o.i64[1] = (int64_t)rp.pps;
                    store_result(out, rep, o);
                }
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// aloha_compact_kernel: the same event body as aloha_fast_kernel on 4-byte tokens (device/engine32.cuh heap24x: time
// relative to a moving epoch, 8-bit code) with the stations' loop counters in a 16-bit array: 400 instead of 528 bytes of
// shared memory per replication, i.e. FOUR blocks of 128 replications per SM instead of three, and a sift moves one
// word per level instead of two.  First choice when the load sweep keeps packets per station below 2^16 and waiting
// times far below 2^24 - 2^22 ticks (aloha_compact_applies); a replication that draws a longer wait is redone by the
// 64-bit kernel.  Saved state: same header, word 15 = epoch, per heap slot {token word, counter of station k}.
// ---------------------------------------------------------------------------------------------------------------
template <bool STATEFUL>
__global__ void __launch_bounds__(AF_THREADS, AC_BLOCKS_PER_SM)
aloha_compact_kernel(cxxdes_b200_aloha_params prm, shard sh, device_columns out, philox_schedule keys, uint32_t until,
                  aloha_saved st) {
    extern __shared__ __align__(16) unsigned char smem[];
    log_entry *table = (log_entry *)smem;
    unsigned char *cursor = smem + sizeof(log_entry) * (1 << CXB_LOG_TABLE_BITS);
    stage_log_table(table, sh.log_table);
    const int N = (int)prm.num_stations;
    const int heap_cap = N + 2;
    const lane_array<uint32_t> heap_words = carve<uint32_t>(cursor, heap24x<AF_THREADS>::FLAT_POP_COLUMN(heap_cap));   // (+ the sentinel, + what a straight-line pop may read)
    uint32_t *heap = heap_words.base;
    // completed transmissions of every station (the `for` counter of aloha.cpp:53): 16 bits each, [station][thread]
    const uint32_t cnt_base = (uint32_t)__cvta_generic_to_shared(carve<uint16_t>(cursor, N).base);
    auto cnt_load = [&](uint32_t station) {
        uint32_t v;
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(cnt_base + station * (AF_THREADS * 2u)));
        return v;
    };
    auto cnt_store = [&](uint32_t station, uint32_t v) {
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(cnt_base + station * (AF_THREADS * 2u)), "r"(v));
    };

    const int64_t frame_ticks = real_to_sim_f32(prm.frame_time, sh.clk);  // env.timeout(float), aloha.cpp:60
    constexpr uint32_t DELAY_LIMIT = heap24x<AF_THREADS>::DELAY_LIMIT;
    const uint32_t frame32 = (uint64_t)frame_ticks >= (uint64_t)DELAY_LIMIT ? DELAY_LIMIT : (uint32_t)frame_ticks;

    heap24x<AF_THREADS> env;   // 4-byte tokens (time - epoch) << 8 | code; every access goes through env
    env.init(heap, heap_cap);
    uint64_t rep = 0, active_set = 0, collided = 0;   // (a station is past its `insert` exactly while it is in the set:
                                                      // the resume-point mask of the 8-byte kernel IS active_set here)
    uint32_t successful = 0, main_pc = 0, remaining = 0, ctr2 = 0, ctr3 = 0;
    uint32_t n_active = 0;   // |active_set| (the size() of aloha.cpp:73's set), kept beside the mask
    uint32_t root = 0;       // entry 0 of the lane's heap: loaded once, at the end of the event before (or at the claim)
    bool armed = false;
    aloha_rep_params rp{};
    divisor rate{};
    float s_result = 0.0f, g_result = 0.0f;
    bool active = false, exhausted = false;

    // Longest first: the replications of a load sweep differ 30x in length (packets per station = G * 1000,
    // aloha.cpp:101), and a lane that draws a long one near the end of the work list keeps its whole warp -- 31 idle
    // lanes -- running for up to a third of the launch.  So the work list is walked from the heaviest sweep step down:
    // claim k stands for member k % per_step of step (S - 1 - k / per_step); the tail of the launch is then made of
    // the shortest replications.  (Which replication a lane runs never changes a result: the Philox key is the global id.)
    const uint32_t S = prm.sweep_steps > 1 ? prm.sweep_steps : 1u;
    const uint64_t per_step = (sh.n_replications + S - 1) / S;
    const uint64_t claim_limit = per_step * S;
    const bool descending = prm.lambda_end >= prm.lambda_begin;
    const uint32_t first_step = (uint32_t)(sh.first_replication % S);

    for (;;) {
        if (!active && !exhausted) {
            const unsigned long long claim = atomicAdd(sh.work_counter, 1ULL);
            bool claimed = claim < claim_limit;
            if (claimed) {
                const uint32_t rank = (uint32_t)(claim / per_step);
                const uint32_t step = descending ? S - 1 - rank : rank;
                rep = (uint64_t)((step + S - first_step) % S) + (claim % per_step) * S;
                claimed = rep < sh.n_replications;     // (the steps do not all have per_step members)
            } else {
                exhausted = true;
            }
            if (claimed) {
                const uint32_t phase = STATEFUL && st.resume ? st.words[rep] : 0u;
                if (STATEFUL && phase == AF_PHASE_DONE) {   // now_ = max(now_, t), environment.ipp:194
                    if (sh.run_until >= 0 && out.end_time[rep] < sh.run_until) out.end_time[rep] = sh.run_until;
                } else if (STATEFUL && phase == AF_PHASE_REPLAY) {
                    const unsigned long long at = atomicAdd(sh.retry_count, 1ULL);
                    sh.retry_list[at] = (uint32_t)rep;
                } else {
                    const uint64_t grep = sh.first_replication + rep;
                    rp = aloha_params_of(prm, grep);
                    // exponential_rv{seed, cfg_.lambda / cfg_.num_stations}: float / size_t -> float -> double (:52)
                    rate = make_divisor((double)(rp.lambda / (float)prm.num_stations));
                    ctr2 = (uint32_t)grep;
                    ctr3 = (uint32_t)(grep >> 32);
                    env.init(heap, heap_cap);
                    s_result = g_result = 0.0f;
                    if (STATEFUL && phase == AF_PHASE_SAVED) {
                        const uint32_t *sv = st.words + rep;
                        const size_t n = st.n_replications;
                        env.n = sv[1 * n];
                        env.now = sv[2 * n];
                        env.n_events = sv[3 * n];
                        env.hash = (uint64_t)sv[4 * n] | ((uint64_t)sv[5 * n] << 32);
                        active_set = (uint64_t)sv[6 * n] | ((uint64_t)sv[7 * n] << 32);
                        collided = (uint64_t)sv[8 * n] | ((uint64_t)sv[9 * n] << 32);
                        successful = sv[12 * n];
                        main_pc = sv[13 * n] & 0xFFu;
                        armed = (sv[13 * n] >> 8) != 0;
                        remaining = sv[14 * n];
                        env.epoch = sv[15 * n];
                        for (uint32_t k = 0; k < env.n; ++k) env.set_entry(k, sv[(size_t)(16 + 2 * k) * n]);
                        for (int k = 0; k < N; ++k) cnt_store((uint32_t)k, sv[(size_t)(17 + 2 * k) * n]);
                        n_active = (uint32_t)__popcll(active_set);
                    } else {
                        active_set = collided = 0;
                        successful = 0; main_pc = 0; remaining = 0; armed = false; n_active = 0;
                        env.schedule(0, CODE_MAIN);  // environment::bind(co_main)
                    }
                    active = true;
                    root = env.entry(0);
                    if (STATEFUL && phase == AF_PHASE_SAVED && env.time_of(root) > until) {
                        // nothing of this replication falls into the piece: only its clock moves (environment.ipp:194)
                        if (sh.run_until > (int64_t)env.now) {
                            out.end_time[rep] = sh.run_until;
                            st.words[2 * st.n_replications + rep] = (uint32_t)sh.run_until;
                        }
                        active = false;
                    }
                }
            }
        }
        // (a claimed replication can leave the lane inactive -- finished in an earlier piece, handed to the
        // retry list, nothing due in this piece -- so the warp only leaves when every lane has seen the end of the work list)
        if (!__any_sync(0xFFFFFFFFu, active)) {
            if (__all_sync(0xFFFFFFFFu, exhausted)) break;
            continue;
        }

        if (active) {
            // ---- one event, branch-light: the token at the root decides everything the event does, so it is decoded
            // first; the variate a station draws when its frame ends is computed for EVERY event (half of them use
            // it: a branch would run it for half of the lanes at the same cost and serialise it with the heap moves
            // it does not depend on); the event's effects are selects.  No loop is left in the body: the pop is three
            // predicated two-level steps, the push one round trip over the new leaf's ancestors (engine32.cuh).
            const uint32_t code = root & 0xFFu;
            const bool is_station = code - 1u < 64u;
            const uint32_t i = cnt_load(is_station ? code - 1u : 0u);
            const philox_block b = philox4x32_10_scheduled(i >> 1, code, ctr2, ctr3, keys);
            const uint64_t bits = (i & 1) ? (((uint64_t)b.w[3] << 32) | b.w[2]) : (((uint64_t)b.w[1] << 32) | b.w[0]);
            const int64_t wait_ticks = real_to_sim(exponential_from_bits(bits, rate, table), sh.clk);

#if CXB_ALOHA_FLAT_POP
            env.pop_after_top_flat<3, 6>();                        // (straight-line: three two-level steps cover 66 entries)
#else
            env.pop_after_top<6>();                                // environment::step, environment.ipp:117-126 (the displaced
                                                                   // last element re-inserted path-parallel: +1.8 %)
#endif
            token32 t;
            t.time = env.time_of(root);
            t.tag = code;
            env.now = t.time > env.now ? t.time : env.now;
            ++env.n_events;

            const bool is_handler = code == CODE_HANDLER, is_main = code == CODE_MAIN;
            const uint64_t bit = is_station ? 1ULL << ((code - 1u) & 63u) : 0ULL;
            // station, aloha.cpp:51-71
            const bool tx_end = (active_set & bit) != 0;                          // frame finished: erase, count, draw (:62-69)
            const bool tx_start = is_station && !tx_end && i < rp.pps;       // collided = false; insert; check (:55-59,73-78)
            const bool st_return = is_station && !tx_end && !(i < rp.pps);   // co_return -> completion with the all_of handler
            successful += (tx_end && !(collided & bit)) ? 1u : 0u;
            // erase (the bit is set) and insert (it is clear) are both a toggle of the station's bit
            active_set ^= (tx_end || tx_start) ? bit : 0ULL;
            n_active += (tx_start ? 1u : 0u) - (tx_end ? 1u : 0u);
            // collided = false; check_collisions(): with company, every member of the set -- this one included -- collides
            collided = (collided & ~(tx_start ? bit : 0ULL)) | ((tx_start && n_active > 1) ? active_set : 0ULL);
            // all_of handler (any_of.ipp:9-26): the output token inherits the child token's time (== now)
            remaining -= is_handler ? 1u : 0u;
            const bool fire = is_handler && armed && remaining == 0;
            armed = armed && !fire;
            // digest: kind and process of the popped token
            const uint32_t kind = is_handler ? (uint32_t)TOKEN_HANDLER : ((is_station || is_main) ? (uint32_t)TOKEN_RESUME : (uint32_t)TOKEN_NOOP);
            const uint32_t proc = is_station ? code : ((is_handler || is_main) ? 0u : (uint32_t)PROC_NONE);
            env.fold(t.time, kind, proc);

            bool push = is_station || fire;
            // (a wait beyond the relative time of a token is passed on as the limit itself: schedule refuses it)
            const uint32_t wait32 = (uint64_t)wait_ticks >= (uint64_t)DELAY_LIMIT ? DELAY_LIMIT : (uint32_t)wait_ticks;
            const uint32_t delay = tx_end ? wait32 : (tx_start ? frame32 : 0u);
            uint32_t push_tag = (tx_end || tx_start) ? code : (st_return ? CODE_HANDLER : CODE_MAIN);
            if (tx_end) cnt_store(code - 1u, i + 1);     // the `for` loop's ++i (aloha.cpp:53)
            if (is_main) {   // twice per replication
                if (main_pc == 0) {
                    // co_main, aloha.cpp:39-44: bind stations in index order, one all_of handler
                    for (int s = 0; s < N; ++s) {
                        cnt_store((uint32_t)s, 0u);
                        env.schedule(env.now, (uint32_t)(1 + s));
                    }
                    remaining = (uint32_t)N;
                    armed = true;
                    main_pc = 1;
                } else {
                    // :46-47  S = succ / now_seconds() (size_t / double -> float), G = lambda * frame_time
                    s_result = (float)((double)successful / sim_to_seconds((int64_t)env.now, sh.clk));
                    g_result = rp.lambda * prm.frame_time;
                    push = true;             // co_main returns: null-target completion
                    push_tag = CODE_NOOP;
                }
            }
            // token at now + delay: schedule checks that the delay fits the relative time of a token; the sum must fit 32
            // bits (anything else is redone by the 64-bit kernel)
            if (push && env.now >= 0xFF000000u) env.status |= CXXDES_B200_REP_TIME_RANGE;
            const uint32_t when = env.now + delay;
            // (every ancestor of the new leaf in one round trip: indices up to N + 1 = 65 are at most 6 levels deep)
#if CXB_PATH_PUSH
            if (push) env.schedule_path<6>(when, push_tag);
#else
            if (push) env.schedule(when, push_tag);
#endif

            // environment::run / run_until (environment.ipp:179-196); the root loaded here is the next event's
            root = env.entry(0);     // (an empty heap: whatever entry 0 holds, not looked at)
            if (env.status || !env.has_event() || env.time_of(root) > until) {
                active = false;
                if (env.status & CXXDES_B200_REP_TIME_RANGE) {
                    const unsigned long long at = atomicAdd(sh.retry_count, 1ULL);
                    sh.retry_list[at] = (uint32_t)rep;
                    if (STATEFUL) st.words[rep] = AF_PHASE_REPLAY;
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
                        const size_t n = st.n_replications;
                        if (o.status & CXXDES_B200_REP_UNFINISHED) {
                            sv[1 * n] = env.n;
                            sv[2 * n] = (uint32_t)end;      // run_until leaves now_ at the horizon
                            sv[3 * n] = env.n_events;
                            sv[4 * n] = (uint32_t)env.hash;
                            sv[5 * n] = (uint32_t)(env.hash >> 32);
                            sv[6 * n] = (uint32_t)active_set;
                            sv[7 * n] = (uint32_t)(active_set >> 32);
                            sv[8 * n] = (uint32_t)collided;
                            sv[9 * n] = (uint32_t)(collided >> 32);
                            sv[10 * n] = (uint32_t)active_set;
                            sv[11 * n] = (uint32_t)(active_set >> 32);
                            sv[12 * n] = successful;
                            sv[13 * n] = main_pc | ((armed ? 1u : 0u) << 8);
                            sv[14 * n] = remaining;
                            sv[15 * n] = env.epoch;
                            for (uint32_t k = 0; k < env.n; ++k) sv[(size_t)(16 + 2 * k) * n] = env.entry(k);
                            for (int k = 0; k < N; ++k) sv[(size_t)(17 + 2 * k) * n] = cnt_load((uint32_t)k);
                            sv[0] = AF_PHASE_SAVED;
                        } else {
                            sv[0] = AF_PHASE_DONE;
                        }
                    }
                    o.end_time = end;
                    o.n_events = env.n_events;
                    o.trace_hash = env.hash;
                    o.f64[0] = (double)s_result;
                    o.f64[1] = (double)g_result;
                    o.i64[0] = (int64_t)successful;
                    o.i64[1] = (int64_t)rp.pps;
                    store_result(out, rep, o);
                }
            }
        }
        __syncwarp();
    }
}

size_t aloha_smem_bytes(uint32_t num_stations) {
    const size_t n = num_stations;
    return sizeof(log_entry) * (1 << CXB_LOG_TABLE_BITS) + 64 +
           (size_t)ALOHA_THREADS * ((n + 2) * (8 + 4) + n * 4);
}

namespace {
size_t aloha_fast_smem_bytes(uint32_t num_stations) {
    return sizeof(log_entry) * (1 << CXB_LOG_TABLE_BITS) + 64 + (size_t)AF_THREADS * (num_stations + 2) * 8;
}

// The compact kernel counts events and loop iterations in 32 / 25 bits and keeps every delay
// below 2^31 ticks; anything else runs on the generic kernel.
bool aloha_fast_applies(const cxxdes_b200_aloha_params &p, const shard &sh) {
    double lam_lo = p.lambda_begin, lam_hi = p.lambda_begin;
    double pps = (double)p.packets_per_station;
    if (p.sweep_steps > 1) {
        lam_lo = p.lambda_begin < p.lambda_end ? p.lambda_begin : p.lambda_end;
        lam_hi = p.lambda_begin < p.lambda_end ? p.lambda_end : p.lambda_begin;
        pps = lam_hi * (double)p.packets_scale;
    }
    if (!(lam_lo > 0.0) || !(pps < 8388608.0)) return false;
    const double tpu = (double)sh.clk.ticks_per_unit;
    // a variate is at most 36.05 / rate (rate = lambda / N)
    if (!(36.05 * tpu * (double)p.num_stations / lam_lo < 1073741824.0)) return false;
    return (double)p.frame_time * tpu < 1073741824.0;
}
}  // namespace

bool aloha_keeps_state(const cxxdes_b200_aloha_params &prm, const shard &sh) { return aloha_fast_applies(prm, sh); }

namespace {
size_t aloha_compact_smem_bytes(uint32_t num_stations) {
    return sizeof(log_entry) * (1 << CXB_LOG_TABLE_BITS) + 64 +
           (size_t)AF_THREADS * ((size_t)heap24x<AF_THREADS>::FLAT_POP_COLUMN((int)num_stations + 2) * 4 + (size_t)num_stations * 2);
}
}  // namespace

// 4-byte tokens hold times up to 2^24 ticks past a moving epoch and the station counters are 16 bits wide.  A frame
// always fits or never does; a waiting time is an exponential variate, so "fits" is a matter of probability: the
// kernel is chosen when a variate passes heap24x::DELAY_LIMIT with probability below e^-18 = 1.5e-8 (config 3:
// about one replication in a thousand sweeps of 262 144), and the replication that draws one is redone by the 64-bit
// kernel (REP_TIME_RANGE -> retry list), as for any other range the 32-bit kernels leave.
bool aloha_compact_applies(const cxxdes_b200_aloha_params &p, const shard &sh) {
    if (!aloha_fast_applies(p, sh)) return false;
    double lam_lo = p.lambda_begin, pps = (double)p.packets_per_station;
    if (p.sweep_steps > 1) {
        lam_lo = p.lambda_begin < p.lambda_end ? p.lambda_begin : p.lambda_end;
        const double lam_hi = p.lambda_begin < p.lambda_end ? p.lambda_end : p.lambda_begin;
        pps = lam_hi * (double)p.packets_scale;
    }
    if (!(pps < 65535.0)) return false;
    const double tpu = (double)sh.clk.ticks_per_unit, limit = (double)heap24x<AF_THREADS>::DELAY_LIMIT;
    // (CXXDES_B200_ALOHA_TAIL: the exponent, for the tests that want replications to take the hand-over)
    double tail = 18.0;
    if (const char *text = getenv("CXXDES_B200_ALOHA_TAIL")) tail = atof(text) > 0.0 ? atof(text) : tail;
    if (!(tail * tpu * (double)p.num_stations / lam_lo < limit)) return false;
    return (double)p.frame_time * tpu < limit;
}

size_t aloha_saved_words(const cxxdes_b200_aloha_params &prm) { return 16 + 2 * ((size_t)prm.num_stations + 2); }

cudaError_t launch_aloha(const cxxdes_b200_aloha_params &prm, const shard &sh, const device_columns &out,
                         int sm_count, size_t smem_optin, unsigned long long *list_cursor, cudaStream_t stream,
                         int *launches, const aloha_saved &st, bool coop, bool wide_tokens) {
    const size_t smem = aloha_smem_bytes(prm.num_stations);
    if (smem > smem_optin) return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(aloha_kernel<environment>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    if (aloha_fast_applies(prm, sh)) {
        const size_t fsmem = aloha_fast_smem_bytes(prm.num_stations);
        e = cudaFuncSetAttribute(aloha_fast_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(aloha_fast_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem);
        if (e != cudaSuccess) return e;
        const uint32_t until = (sh.run_until < 0 || sh.run_until >= 0xFFFFFFFFll) ? 0xFFFFFFFFu : (uint32_t)sh.run_until;
        const philox_schedule keys = philox_schedule_of((uint32_t)sh.seed, (uint32_t)(sh.seed >> 32));
        if (coop && !st.words) {
            // first pass with the event heap of a replication operated by a group of 8 lanes (models_aloha_coop.cu)
            e = launch_aloha_coop(prm, sh, out, sm_count, stream);
            if (e != cudaSuccess) return e;
        } else if (!wide_tokens && aloha_compact_applies(prm, sh)) {
            const size_t csmem = aloha_compact_smem_bytes(prm.num_stations);
            e = cudaFuncSetAttribute(aloha_compact_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(aloha_compact_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem);
            if (e != cudaSuccess) return e;
            if (st.words) aloha_compact_kernel<true><<<sm_count * AC_BLOCKS_PER_SM, AF_THREADS, csmem, stream>>>(prm, sh, out, keys, until, st);
            else aloha_compact_kernel<false><<<sm_count * AC_BLOCKS_PER_SM, AF_THREADS, csmem, stream>>>(prm, sh, out, keys, until, st);
        } else if (st.words) {
            aloha_fast_kernel<true><<<sm_count * AF_BLOCKS_PER_SM, AF_THREADS, fsmem, stream>>>(prm, sh, out, keys, until, st);
        } else {
            aloha_fast_kernel<false><<<sm_count * AF_BLOCKS_PER_SM, AF_THREADS, fsmem, stream>>>(prm, sh, out, keys, until, st);
        }
        ++*launches;
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        // the replications it handed back (normally none): exits at once when the list is empty
        aloha_kernel<environment><<<sm_count * per_sm, ALOHA_THREADS, smem, stream>>>(prm, sh, out, sh.retry_list, sh.retry_count,
                                                                                      list_cursor, trace_sink{});
        ++*launches;
        return cudaGetLastError();
    }
    aloha_kernel<environment><<<sm_count * per_sm, ALOHA_THREADS, smem, stream>>>(prm, sh, out, nullptr, nullptr, nullptr, trace_sink{});
    ++*launches;
    return cudaGetLastError();
}

// One replication (sh.n_replications == 1) on the full-size kernel, with its event trace.
cudaError_t launch_aloha_trace(const cxxdes_b200_aloha_params &prm, const shard &sh, const device_columns &out,
                               const trace_sink &sink, cudaStream_t stream) {
    const size_t smem = aloha_smem_bytes(prm.num_stations);
    cudaError_t e = cudaFuncSetAttribute(aloha_kernel<traced_environment>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    aloha_kernel<traced_environment><<<1, ALOHA_THREADS, smem, stream>>>(prm, sh, out, nullptr, nullptr, nullptr, sink);
    return cudaGetLastError();
}

}  // namespace cxb
