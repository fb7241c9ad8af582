// models_aloha_coop.cu -- pure ALOHA (examples/aloha.cpp:25-85) with the event heap of each replication operated by a
// GROUP of 8 lanes (device/coop_heap.cuh: sub-warp cooperative push / pop, ballots, libstdc++-exact moves).
//
// Same lowering as aloha_fast_kernel (models_aloha.cu): tokens {u32 time, u32 tag}, the tag's low
// 7 bits name the target (0 co_main, 1..64 station, 65 the all_of handler, 66 a null-target completion) and the station's
// loop counter rides in the upper 25 bits.  A warp holds 4 replications; the scalar part of an event (which process
// resumes, the station masks, the variate) is computed by all 8 lanes of the group alike -- every lane holds the same
// copy of the replication's registers -- and only the heap moves are split between them.
//
// Selected with CXXDES_B200_FLAG_COOP_HEAP (A/B against the thread-per-replication kernel); plain run() / run_until
// from tick 0 only (no saved state between pieces).  Replications that leave the 32-bit range go to the retry list
// and are redone by aloha_kernel, as with aloha_fast_kernel.
#include "device/coop_heap.cuh"
#include "device/engine32.cuh"
#include "device/models_common.cuh"
#include "kernels.h"

namespace cxb {

namespace {

constexpr int AC_THREADS = 256;                       // 32 replications per block
constexpr int AC_BLOCKS_PER_SM = 4;                   // 64 registers per thread
constexpr int AC_GROUPS = AC_THREADS / coop::G;
constexpr int AC_SLOTS = 68;                          // slot 0 unused, 66 entries, one pad: 544 bytes, 16-byte multiple
constexpr int AC_NODE_ROUNDS = 4;                     // nodes 0..31 can have two children while len <= 65
constexpr uint32_t CODE_MAIN = 0, CODE_HANDLER = 65, CODE_NOOP = 66, CODE_MASK = 127;
constexpr int COUNTER_SHIFT = 7;

struct rep_params {
    float lambda;
    uint64_t pps;
};

// main(), aloha.cpp:96,101: float arithmetic, size_t operands converted to float (same as models_aloha.cu).
__device__ __forceinline__ rep_params params_of(const cxxdes_b200_aloha_params &p, uint64_t grep) {
    rep_params r;
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

__global__ void __launch_bounds__(AC_THREADS, AC_BLOCKS_PER_SM)
aloha_coop_kernel(cxxdes_b200_aloha_params prm, shard sh, device_columns out, philox_schedule keys, uint32_t until) {
    extern __shared__ __align__(16) unsigned char smem[];
    log_entry *table = (log_entry *)smem;
    coop::entry *all_slots = (coop::entry *)(smem + sizeof(log_entry) * (1 << CXB_LOG_TABLE_BITS));
    stage_log_table(table, sh.log_table);
    const int N = (int)prm.num_stations;
    const int64_t frame_ticks = real_to_sim_f32(prm.frame_time, sh.clk);  // env.timeout(float), aloha.cpp:60
    const unsigned leader = threadIdx.x & 31u & ~(unsigned)(coop::G - 1);   // lane of the warp that speaks for the group

    coop::group_heap heap;
    heap.attach(all_slots + (size_t)(threadIdx.x / coop::G) * AC_SLOTS, N + 2);

    // the replication's registers: identical in all 8 lanes of the group
    uint32_t now = 0, n_events = 0, status = 0;
    uint64_t hash = TRACE_HASH_INIT;
    uint64_t rep = 0, active_set = 0, collided = 0, pc_tx = 0;
    uint32_t successful = 0, main_pc = 0, remaining = 0, ctr2 = 0, ctr3 = 0;
    bool armed = false;
    rep_params rp{};
    divisor rate{};
    float s_result = 0.0f, g_result = 0.0f;
    bool active = false, exhausted = false;

    for (;;) {
        // ---- claim a replication for the group (its leader asks, everybody hears) ----
        uint32_t first_push = 0;
        const unsigned claiming = __ballot_sync(0xFFFFFFFFu, !active && !exhausted);   // whole groups
        if (!active && !exhausted) {
            unsigned long long r = 0;
            if ((threadIdx.x & (coop::G - 1)) == 0) r = atomicAdd(sh.work_counter, 1ULL);
            r = __shfl_sync(claiming, r, (int)leader);
            if (r < sh.n_replications) {
                rep = r;
                const uint64_t grep = sh.first_replication + rep;
                rp = params_of(prm, grep);
                // exponential_rv{seed, cfg_.lambda / cfg_.num_stations}: float / size_t -> float -> double (:52)
                rate = make_divisor((double)(rp.lambda / (float)prm.num_stations));
                ctr2 = (uint32_t)grep;
                ctr3 = (uint32_t)(grep >> 32);
                now = 0; n_events = 0; status = 0;
                hash = TRACE_HASH_INIT;
                heap.n = 0;
                active_set = collided = pc_tx = 0;
                successful = 0; main_pc = 0; remaining = 0; armed = false;
                s_result = g_result = 0.0f;
                active = true;
                first_push = 1;   // environment::bind(co_main): start token {0, co_main}
            } else {
                exhausted = true;
            }
        }
        __syncwarp();
        if (!__any_sync(0xFFFFFFFFu, active)) break;   // (a claim always makes the group active)

        // ---- environment::step (environment.ipp:117-126): pop, advance the clock, count ----
        int pending = 0;                 // tokens this event schedules
        uint32_t push_time = 0, push_tag = CODE_NOOP;
        const bool popping = active && !first_push;
        const coop::entry t = heap.pop<AC_NODE_ROUNDS>(popping);
        if (first_push) {
            pending = 1;
            push_time = 0;
            push_tag = CODE_MAIN;
        } else if (active) {
            now = t.time > now ? t.time : now;
            ++n_events;
            const uint32_t code = t.tag & CODE_MASK;
            int64_t delay = 0;
            pending = 1;
            if (code - 1u < 64u) {
                // station, aloha.cpp:51-71
                hash = trace_hash_step(hash, (int64_t)t.time, 0, TOKEN_RESUME, code);
                const uint64_t bit = 1ULL << (code - 1);
                const uint32_t i = t.tag >> COUNTER_SHIFT;
                if (pc_tx & bit) {
                    // frame finished: erase, count, draw the waiting time (:62-69)
                    pc_tx &= ~bit;
                    active_set &= ~bit;
                    if (!(collided & bit)) ++successful;
                    const philox_block b = philox4x32_10_scheduled(i >> 1, code, ctr2, ctr3, keys);
                    const uint64_t bits = (i & 1) ? (((uint64_t)b.w[3] << 32) | b.w[2]) : (((uint64_t)b.w[1] << 32) | b.w[0]);
                    delay = real_to_sim(exponential_from_bits(bits, rate, table), sh.clk);
                    push_tag = code | ((i + 1) << COUNTER_SHIFT);
                } else if (i < rp.pps) {
                    // collided = false; insert; check_collisions (:55-59,73-78)
                    collided &= ~bit;
                    active_set |= bit;
                    if (__popcll(active_set) > 1) collided |= active_set;
                    pc_tx |= bit;
                    delay = frame_ticks;
                    push_tag = t.tag;
                } else {
                    push_tag = CODE_HANDLER;  // co_return -> completion token with the all_of handler
                }
            } else if (code == CODE_HANDLER) {
                // all_of handler (any_of.ipp:9-26): the output token inherits the child token's time (== now)
                hash = trace_hash_step(hash, (int64_t)t.time, 0, TOKEN_HANDLER, 0);
                --remaining;
                pending = (armed && remaining == 0) ? 1 : 0;
                if (pending) armed = false;
                push_tag = CODE_MAIN;
            } else if (code == CODE_MAIN) {
                hash = trace_hash_step(hash, (int64_t)t.time, 0, TOKEN_RESUME, 0);
                if (main_pc == 0) {
                    // co_main, aloha.cpp:39-44: bind stations in index order (N pushes below), one all_of handler
                    remaining = (uint32_t)N;
                    armed = true;
                    main_pc = 1;
                    pending = N;
                    push_tag = 1;             // stations 1 .. N, counter 0
                } else {
                    // :46-47  S = succ / now_seconds() (size_t / double -> float), G = lambda * frame_time
                    s_result = (float)((double)successful / sim_to_seconds((int64_t)now, sh.clk));
                    g_result = rp.lambda * prm.frame_time;
                    // co_main returns: null-target completion
                }
            } else {
                hash = trace_hash_step(hash, (int64_t)t.time, 0, TOKEN_NOOP, PROC_NONE);
                pending = 0;
            }
            // token at now + delay; the sum must stay inside 32 bits (else the 64-bit kernel redoes the replication)
            const uint64_t at = (uint64_t)now + (uint64_t)delay;
            if ((uint64_t)delay >= (1ULL << 31) || at >= 0xFFFFFFFFull) status |= CXXDES_B200_REP_TIME_RANGE;
            push_time = (uint32_t)at;
        }

        // ---- environment::schedule_token for what the event scheduled: the whole warp pushes together ----
        while (__any_sync(0xFFFFFFFFu, pending > 0)) {
            const bool mine = pending > 0;
            if (!heap.push(mine, push_time, push_tag)) status |= CXXDES_B200_REP_HEAP_OVERFLOW;
            if (mine) {
                --pending;
                ++push_tag;       // (only co_main's first event schedules more than one token: stations 1, 2, ... N)
            }
        }

        // ---- environment::run / run_until (environment.ipp:179-196) ----
        if (active) {
            const uint32_t next = heap.n > 0 ? heap.slots[coop::slot_of(0)].time : 0xFFFFFFFFu;
            if (status || heap.n == 0 || next > until) {
                active = false;
                if ((threadIdx.x & (coop::G - 1)) == 0) {
                    if (status & CXXDES_B200_REP_TIME_RANGE) {
                        const unsigned long long at = atomicAdd(sh.retry_count, 1ULL);
                        sh.retry_list[at] = (uint32_t)rep;
                    } else {
                        rep_out o;
                        int64_t end = (int64_t)now;
                        o.status = status;
                        if (sh.run_until >= 0) {
                            if (heap.n > 0) o.status |= CXXDES_B200_REP_UNFINISHED;
                            end = end > sh.run_until ? end : sh.run_until;
                        }
                        o.end_time = end;
                        o.n_events = n_events;
                        o.trace_hash = hash;
                        o.f64[0] = (double)s_result;
                        o.f64[1] = (double)g_result;
                        o.i64[0] = (int64_t)successful;
                        o.i64[1] = (int64_t)rp.pps;
                        store_result(out, rep, o);
                    }
                }
            }
        }
        __syncwarp();
    }
}

cudaError_t launch_aloha_coop(const cxxdes_b200_aloha_params &prm, const shard &sh, const device_columns &out,
                              int sm_count, cudaStream_t stream) {
    const size_t smem = sizeof(log_entry) * (1 << CXB_LOG_TABLE_BITS) + (size_t)AC_GROUPS * AC_SLOTS * sizeof(coop::entry);
    cudaError_t e = cudaFuncSetAttribute(aloha_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const uint32_t until = (sh.run_until < 0 || sh.run_until >= 0xFFFFFFFFll) ? 0xFFFFFFFFu : (uint32_t)sh.run_until;
    const philox_schedule keys = philox_schedule_of((uint32_t)sh.seed, (uint32_t)(sh.seed >> 32));
    aloha_coop_kernel<<<sm_count * AC_BLOCKS_PER_SM, AC_THREADS, smem, stream>>>(prm, sh, out, keys, until);
    return cudaGetLastError();
}

}  // namespace cxb
