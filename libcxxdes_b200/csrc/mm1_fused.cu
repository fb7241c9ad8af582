// mm1_fused.cu -- the headline kernel: M/M/1 (examples/producer_consumer.cpp:9-59) with the whole
// replication held on-chip.
//
// One thread owns one replication from its first event to its last.  What the generic engine
// keeps in shared/global memory is specialised here from facts about this model
// (all checked against the reference engine by the parity tests):
//   * the event heap never holds more than 2 tokens and every priority is 0, so the libstdc++
//     heap degenerates to "slot 0 is the minimum, an equal key keeps the earlier push in front"
//     -> two (time, tag) register pairs;
//   * the waiter list of queue<double>::event_ holds at most the consumer -> one flag;
//   * the all_of of co_main is a countdown from 2;
//   * the queue payload now_seconds() is a pure function of the put tick, so the ring stores the
//     tick and converts at pop time (bit-identical, half the bytes).
// Every popped token still counts as one event and is folded into the trace digest, including
// handler events and the trailing null-target completion.
//
// Two instantiations:
//   FAST  = true   32-bit ticks, ring of `ring_cap` u32 per thread in SHARED memory.  A replication
//                  that would exceed the ring or the 32-bit tick range stops, and its id is appended
//                  to the retry list (nothing is written to its result row).
//   FAST  = false  64-bit ticks, ring in GLOBAL memory ([slot][thread], coalesced across lanes).
//                  Runs the retry list (or, for heavily loaded queues, every replication).
#include <cmath>

#include "device/models_common.cuh"
#include "kernels.h"

namespace cxb {

namespace {

enum : uint32_t { M_MAIN = 0, M_PROD = 1, M_CONS = 2, M_HANDLER = 3, M_NOOP = 4 };

// (kind << 56) ^ (proc << 40) of trace_hash_step for each tag; priority is always 0.
__device__ __forceinline__ uint64_t tag_hash_bits(uint32_t m) {
    uint64_t kind = m == M_HANDLER ? TOKEN_HANDLER : (m == M_NOOP ? TOKEN_NOOP : TOKEN_RESUME);
    uint64_t proc = m == M_NOOP ? PROC_NONE : (m == M_HANDLER ? 0u : m);
    return (kind << 56) ^ (proc << 40);
}

template <bool FAST>
struct tick_traits;
template <>
struct tick_traits<true> {
    using type = uint32_t;
    static constexpr uint32_t EMPTY = 0xFFFFFFFFu;
    static constexpr uint32_t LIMIT = 0xFFFF0000u;
};
template <>
struct tick_traits<false> {
    using type = int64_t;
    static constexpr int64_t EMPTY = INT64_MAX;
    static constexpr int64_t LIMIT = INT64_MAX / 2;
};

}  // namespace

struct mm1_rates {
    divisor lambda, mu;
    philox_schedule keys;  // round keys of the seed, read from the constant bank inside the loop
};

template <bool FAST, int THREADS, int MIN_BLOCKS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
mm1_fused_kernel(cxxdes_b200_mm1_params prm, mm1_rates rates, shard sh, device_columns out, uint32_t ring_cap,
                 int64_t *gring_base, const uint32_t *list, const unsigned long long *list_count,
                 unsigned long long *list_cursor) {
    using tick_t = typename tick_traits<FAST>::type;
    constexpr tick_t EMPTY = tick_traits<FAST>::EMPTY;

    extern __shared__ __align__(16) unsigned char smem[];
    log_entry *table = (log_entry *)smem;
    stage_log_table(table, sh.log_table);
    // FAST: ring[slot * THREADS + tid] (u32) in shared memory -> bank = tid % 32, conflict free.
    uint32_t *sring = (uint32_t *)(smem + sizeof(log_entry) * (1 << CXB_LOG_TABLE_BITS)) + threadIdx.x;
    const uint64_t total_threads = (uint64_t)gridDim.x * THREADS;
    int64_t *gring = FAST ? nullptr : gring_base + ((uint64_t)blockIdx.x * THREADS + threadIdx.x);

    // divisors {d, RN(1/d)} are formed on the host (IEEE division, same bits as on the device) and
    // arrive as kernel parameters, so the loop only selects between constant-bank values
    const divisor lambda = rates.lambda;
    const divisor mu = rates.mu;
    const uint32_t n_packets = (uint32_t)prm.n_packets;  // the launcher guarantees 2 <= n < 2^31
    const double ticks_per_unit = (double)sh.clk.ticks_per_unit;
    const int64_t until = sh.run_until;

    // Loop-invariant pieces of the clock conversion (environment::now_seconds).
    const int64_t tick_mult = sh.clk.tick_mult;
    const double sec_per_unit = sh.clk.seconds_per_tick_unit;

    // ---- per-lane replication state, all in registers ----
    bool active = false, exhausted = false;
    uint64_t rep = 0;
    uint32_t ctr2 = 0, ctr3 = 0;  // Philox counter words 2,3 = global replication id
    tick_t t0 = EMPTY, t1 = EMPTY, now = 0, x_tick = 0;
    uint32_t m0 = M_NOOP, m1 = M_NOOP;
    uint32_t i = 0, k = 0, head = 0, count = 0, max_queue = 0;
    uint32_t main_pc = 0, cons_pc = 0, remaining = 2;
    bool waiting = false;  // consumer parked in queue.event_ (event.hpp:136-139)
    bool fast = false, cfirst = false, c_timeout = false;  // steady-state representation, see below
    double total_latency = 0.0;
    uint64_t n_events = 0, hash = TRACE_HASH_INIT;
    uint32_t status = 0;

    auto push = [&](tick_t t, uint32_t m) {
        // std::push_heap on <= 2 entries: the new token goes in front only if the current front
        // is strictly later (stl_heap.h:135-148); an empty front compares as latest.
        if (t1 != EMPTY) status |= CXXDES_B200_REP_HEAP_OVERFLOW;
        bool front = t0 > t;
        t1 = front ? t0 : t;
        m1 = front ? m0 : m;
        t0 = front ? t : t0;
        m0 = front ? m : m0;
    };

    // The warp advances in lockstep, one event per lane per iteration.  Which process a lane
    // resumes differs from lane to lane, so the cheap bookkeeping below diverges; the expensive
    // part (Philox + log + divide for the next timeout) is identical for producer and consumer
    // and is executed once, by all lanes together, after the __syncwarp() join.  A lane whose
    // replication ends steals the next one at once (replication-level work stealing), so lanes
    // only idle when the whole ensemble is exhausted.
    for (;;) {
        if (!active && !exhausted) {
            bool got;
            if (list) {
                unsigned long long idx = atomicAdd(list_cursor, 1ULL);
                got = idx < *list_count;
                if (got) rep = list[idx];
            } else {
                got = steal(sh, rep);
            }
            if (got) {
                const uint64_t grep = sh.first_replication + rep;
                ctr2 = (uint32_t)grep; ctr3 = (uint32_t)(grep >> 32);
                t0 = 0; m0 = M_MAIN;   // environment::bind(co_main): start token {0, 0, main}
                t1 = EMPTY; m1 = M_NOOP;
                now = 0; x_tick = 0;
                i = k = head = count = max_queue = 0;
                main_pc = cons_pc = 0; remaining = 2;
                waiting = false;
                fast = cfirst = c_timeout = false;
                total_latency = 0.0;
                n_events = 0; hash = TRACE_HASH_INIT;
                status = 0;
                active = true;
            } else {
                exhausted = true;
            }
        }
        if (!__any_sync(0xFFFFFFFFu, active)) break;

        bool draw = false, is_prod = false;
        uint32_t draw_index = 0, m = M_NOOP;
        // ---- steady state: exactly one producer token (t0) and at most one consumer token
        // (t1, EMPTY while the consumer is parked in queue.event_).  With <= 2 entries the
        // libstdc++ heap keeps the earlier push in front on equal keys, so the pick below is
        // the pop order: `cfirst` says the consumer's token was pushed before the producer's.
        // (Bitwise &,| on purpose: a short-circuit branch here would split the warp into a
        // producer group and a consumer group that then run the body one after the other.)
        const bool c_goes = (t1 != EMPTY) & ((t1 < t0) | ((t1 == t0) & cfirst));
        const bool leave = c_goes ? ((count != 0) & (k + 1 == n_packets)) : (i == n_packets);
        const bool do_fast = active & fast & !leave;
        bool generic = active & !do_fast;
        if (active & fast & leave) {
            // the next event ends a process (co_return): rebuild the generic two-slot heap
            const tick_t tp = t0, tc = t1;
            waiting = tc == EMPTY;
            t0 = c_goes ? tc : tp;
            m0 = c_goes ? M_CONS : M_PROD;
            t1 = c_goes ? tp : tc;
            m1 = c_goes ? M_PROD : M_CONS;
            cons_pc = c_timeout ? 2u : 1u;
            fast = false;
        }
        __syncwarp();
        if (do_fast) {
            {
                // environment::step (environment.ipp:117-146) for a producer/consumer token
                const tick_t t = c_goes ? t1 : t0;
                now = t;  // tokens are scheduled at now + d, d >= 0: time never runs backwards here
                ++n_events;
                hash = (rotl64(hash, 5) ^ ((uint64_t)t ^ (c_goes ? tag_hash_bits(M_CONS) : tag_hash_bits(M_PROD)))) *
                       TRACE_HASH_MUL;
                is_prod = !c_goes;
                m = c_goes ? M_CONS : M_PROD;
                if (c_goes && c_timeout) {
                    // total_latency += now_seconds() - x   (producer_consumer.cpp:52)
                    double now_s, x_s;
                    if (FAST && tick_mult == 1) {  // uniform branch; (double)(u32) is exact either way
                        now_s = (double)(uint32_t)now * sec_per_unit;
                        x_s = (double)(uint32_t)x_tick * sec_per_unit;
                    } else {
                        now_s = (double)((int64_t)now * tick_mult) * sec_per_unit;
                        x_s = (double)((int64_t)x_tick * tick_mult) * sec_per_unit;
                    }
                    total_latency += (now_s - x_s);
                }
                const bool empty = count == 0;
                const bool pop = c_goes && !empty;   // q.pop() succeeds (queue.hpp:58-65)
                const bool park = c_goes && empty;   // ... or the consumer waits on queue.event_
                const bool put = !c_goes;            // q.put(now_seconds()) (queue.hpp:48-53)
                const bool overflow = put && count >= ring_cap;
                uint32_t slot = put ? head + count : head;
                slot = slot >= ring_cap ? slot - ring_cap : slot;
                if (put && !overflow) {
                    if (FAST) sring[slot * THREADS] = (uint32_t)now;
                    else gring[(uint64_t)slot * total_threads] = (int64_t)now;
                }
                if (pop) {
                    if (FAST) x_tick = sring[slot * THREADS];
                    else x_tick = (tick_t)gring[(uint64_t)slot * total_threads];
                    head = head + 1 >= ring_cap ? 0 : head + 1;
                }
                count += (put ? 1u : 0u) - (pop ? 1u : 0u);
                max_queue = count > max_queue ? count : max_queue;
                if (overflow) status |= CXXDES_B200_REP_QUEUE_OVERFLOW;
                // wake(): a parked consumer gets its token {0 + now} before the producer's timeout
                if (put && t1 == EMPTY) {
                    t1 = now;
                    c_timeout = false;
                }
                if (park) t1 = EMPTY;
                i += put ? 1u : 0u;
                k += pop ? 1u : 0u;
                draw = (put && !overflow) || pop;
                draw_index = is_prod ? i - 1 : k - 1;
            }
        }
        if (generic) {
            // ---- environment::step, environment.ipp:117-146 (start-up and wind-down events) ----
            const tick_t t = t0;
            m = m0;
            t0 = t1; m0 = m1; t1 = EMPTY;           // pop_heap on <= 2 entries
            now = t > now ? t : now;
            ++n_events;
            hash = (rotl64(hash, 5) ^ ((uint64_t)t ^ tag_hash_bits(m))) * TRACE_HASH_MUL;

            if (m == M_PROD || m == M_CONS) {
                // producer and consumer bodies, producer_consumer.cpp:31-54
                is_prod = m == M_PROD;
                const bool put = is_prod && i < n_packets;       // q.put(now_seconds()), queue.hpp:48-53
                const bool overflow = put && count >= ring_cap;
                const bool empty = count == 0;
                const bool pop = !is_prod && !empty;             // q.pop() succeeds, queue.hpp:58-65
                const bool park = !is_prod && empty;             // ... or waits on queue.event_
                if (!is_prod && cons_pc == 2) {
                    double now_s = (double)((int64_t)now * tick_mult) * sec_per_unit;
                    double x_s = (double)((int64_t)x_tick * tick_mult) * sec_per_unit;
                    total_latency += (now_s - x_s);  // :52
                }
                uint32_t slot = put ? head + count : head;
                slot = slot >= ring_cap ? slot - ring_cap : slot;
                if (put && !overflow) {
                    if (FAST) sring[slot * THREADS] = (uint32_t)now;
                    else gring[(uint64_t)slot * total_threads] = (int64_t)now;
                }
                if (pop) {
                    if (FAST) x_tick = sring[slot * THREADS];
                    else x_tick = (tick_t)gring[(uint64_t)slot * total_threads];
                    head = head + 1 >= ring_cap ? 0 : head + 1;
                }
                count += (put ? 1u : 0u) - (pop ? 1u : 0u);
                max_queue = count > max_queue ? count : max_queue;
                if (overflow) status |= CXXDES_B200_REP_QUEUE_OVERFLOW;
                // wake(): the parked consumer's token {0 + now, 0, consumer} goes in before the
                // producer's own timeout (event.hpp:125-134); pop's wake() never finds a waiter
                if (put && waiting) push(now, M_CONS);
                waiting = park || (waiting && !put);
                i += put ? 1u : 0u;
                k += pop ? 1u : 0u;
                const bool last = pop && k == n_packets;          // consumer returns (:45-48)
                const bool ret = (is_prod && !put) || last;       // co_return -> completion token
                if (ret) push(now, M_HANDLER);                    // ... carrying the all_of handler
                draw = (put && !overflow) || (pop && !last);
                draw_index = is_prod ? i - 1 : k - 1;
                cons_pc = park ? 1u : (pop ? 2u : cons_pc);
            } else if (m == M_HANDLER) {
                // all_of handler, any_of.ipp:9-26: output token inherits the child token's time
                if (--remaining == 0) push(t, M_MAIN);
            } else if (m == M_MAIN) {
                if (main_pc == 0) {  // co_await (producer() && consumer()): bind both, any_of.ipp:41-48
                    push(now, M_PROD);
                    push(now, M_CONS);
                    main_pc = 1;
                } else {
                    push(now, M_NOOP);  // co_main returns: null-target completion (coroutine.ipp:352-357)
                }
            }
        }

        __syncwarp();  // join: every lane of the warp takes the draw below together

        if (draw) {
            // env.timeout(exponential()): the only expensive part, shared by both processes
            // key = seed (warp-uniform round keys), counter = (block, stream, replication)
            philox_block b = philox4x32_10_scheduled(draw_index >> 1, is_prod ? 1u : 2u, ctr2, ctr3, rates.keys);
            const bool odd = draw_index & 1;
            uint64_t bits = ((uint64_t)(odd ? b.w[3] : b.w[1]) << 32) | (odd ? b.w[2] : b.w[0]);
            divisor rate;
            rate.d = is_prod ? lambda.d : mu.d;
            rate.inv = is_prod ? lambda.inv : mu.inv;
            double v = exponential_from_bits(bits, rate, table);
            int64_t ticks = (int64_t)(v * ticks_per_unit);  // environment::real_to_sim
            if (FAST && ((uint64_t)ticks >= 0x7FFF0000ULL || now >= tick_traits<FAST>::LIMIT - (tick_t)ticks)) {
                status |= CXXDES_B200_REP_TIME_RANGE;
            } else if (fast) {
                // the drawing process pushes last, so on an equal key the other one's token is in front
                const tick_t due = now + (tick_t)ticks;
                t0 = is_prod ? due : t0;
                t1 = is_prod ? t1 : due;
                cfirst = is_prod;
                c_timeout = is_prod ? c_timeout : true;
            } else {
                push(now + (tick_t)ticks, m);
            }
        }

        if (generic && status == 0 && main_pc == 1 && remaining == 2 && cons_pc != 0 && i >= 1 && i < n_packets &&
            k < n_packets) {
            // both processes are running and only their tokens are scheduled: enter the steady state
            const bool prod_front = m0 == M_PROD && (t1 == EMPTY ? waiting : m1 == M_CONS);
            const bool cons_front = m0 == M_CONS && m1 == M_PROD && t1 != EMPTY;
            if (t0 != EMPTY && (prod_front || cons_front)) {
                const tick_t tp = prod_front ? t0 : t1;
                const tick_t tc = prod_front ? t1 : t0;
                t0 = tp;
                t1 = tc;
                cfirst = cons_front;
                c_timeout = cons_pc == 2;
                fast = true;
            }
        }

        // ---- end of this lane's replication? (environment::run / run_until, environment.ipp:179-196) ----
        const tick_t next = fast && t1 < t0 ? t1 : t0;  // time of the next scheduled token
        if (active && (status != 0 || next == EMPTY || (until >= 0 && (int64_t)next > until))) {
            active = false;
            if (FAST && (status & (CXXDES_B200_REP_QUEUE_OVERFLOW | CXXDES_B200_REP_TIME_RANGE))) {
                // hand the replication to the fallback kernel; its row is written there
                unsigned long long at = atomicAdd(sh.retry_count, 1ULL);
                sh.retry_list[at] = (uint32_t)rep;
            } else {
                int64_t end = (int64_t)now;
                if (until >= 0) {
                    if (next != EMPTY) status |= CXXDES_B200_REP_UNFINISHED;
                    end = end > until ? end : until;  // now_ = max(now_, t), environment.ipp:194
                }
                rep_out o;
                o.end_time = end;
                o.n_events = n_events;
                o.trace_hash = hash;
                o.status = status;
                o.f64[0] = k == n_packets ? total_latency / (double)prm.n_packets : 0.0;  // :46
                o.f64[1] = total_latency;
                o.i64[0] = (int64_t)k;
                o.i64[1] = (int64_t)max_queue;
                store_result(out, rep, o);
            }
        }
    }
}

constexpr int FAST_THREADS = 256;
constexpr int FAST_MIN_BLOCKS = 4;
constexpr int FB_THREADS = 128;

mm1_fused_plan plan_mm1_fused(const cxxdes_b200_mm1_params &prm, uint32_t queue_capacity, int sm_count,
                              size_t smem_optin, int64_t tick_mult, int64_t ticks_per_unit, bool lockstep,
                              uint64_t n_replications) {
    mm1_fused_plan p{};
    p.threads = FAST_THREADS;
    const size_t table_bytes = sizeof(log_entry) * (1 << CXB_LOG_TABLE_BITS);
    // 4 resident blocks of 256 threads per SM: each gets a quarter of the shared memory.
    size_t per_block = (smem_optin < 227 * 1024 ? smem_optin : 227 * 1024) / FAST_MIN_BLOCKS - 1024;
    uint32_t max_ring = (uint32_t)((per_block - table_bytes) / (FAST_THREADS * sizeof(uint32_t)));
    p.smem_ring = queue_capacity ? queue_capacity : max_ring;
    if (p.smem_ring > max_ring) p.smem_ring = max_ring;
    if (p.smem_ring > prm.n_packets) p.smem_ring = (uint32_t)prm.n_packets;
    p.smem_bytes = table_bytes + (size_t)p.smem_ring * FAST_THREADS * sizeof(uint32_t);
    p.blocks = sm_count * FAST_MIN_BLOCKS;
    // Stationary P(queue length > Q) = rho^(Q+1); if more than ~2% of replications are expected to
    // overflow the shared ring at least once, run everything on the global-memory kernel.
    double rho = prm.lambda / prm.mu;
    double p_over = rho >= 1.0 ? 1.0 : pow(rho, (double)p.smem_ring + 1.0) * (double)prm.n_packets;
    p.direct_global = !(p_over < 0.02);
    plan_mm1_ahead(p, prm, queue_capacity, sm_count, tick_mult, ticks_per_unit, n_replications);
    p.fb_threads = FB_THREADS;
    uint64_t want = prm.n_packets;
    if (lockstep ? p.direct_global : p.ah_direct_global) {
        p.fb_blocks = sm_count * 8;
        // memory budget for the rings: 8 GiB
        uint64_t budget = (8ULL << 30) / ((uint64_t)p.fb_blocks * FB_THREADS * sizeof(int64_t));
        if (rho < 1.0) {
            // capacity with overflow probability < 1e-12 per arrival, at least 64
            double need = ceil(log(1e-12 / (double)prm.n_packets) / log(rho)) + 64.0;
            if ((double)want > need) want = (uint64_t)need;
        }
        if (want > budget) want = budget;
    } else {
        p.fb_blocks = sm_count;
        uint64_t budget = (2ULL << 30) / ((uint64_t)p.fb_blocks * FB_THREADS * sizeof(int64_t));
        if (want > budget) want = budget;
    }
    p.fb_ring = (uint32_t)(want < 2 ? 2 : want);
    return p;
}

cudaError_t configure_mm1_fused(const mm1_fused_plan &plan) {
    return cudaFuncSetAttribute(mm1_fused_kernel<true, FAST_THREADS, FAST_MIN_BLOCKS>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem_bytes);
}

// First pass: mm1_ahead.cu (default) or, with `lockstep`, the draw-on-demand kernel above (kept
// for A/B measurements); second pass: the retry list on the global-memory kernel.  Heavily
// loaded queues skip the first pass.
cudaError_t launch_mm1_fused(const cxxdes_b200_mm1_params &prm, const shard &sh, const device_columns &out,
                             const mm1_fused_plan &plan, int64_t *fb_ring, unsigned long long *fb_counter,
                             void *ahead_gring, uint32_t *ahead_state, int resume, cudaStream_t stream, int *launches,
                             bool lockstep, const mm1_overlap &overlap) {
    const size_t table_bytes = sizeof(log_entry) * (1 << CXB_LOG_TABLE_BITS);
    mm1_rates rates;
    rates.lambda = make_divisor(prm.lambda);
    rates.mu = make_divisor(prm.mu);
    rates.keys = philox_schedule_of((uint32_t)sh.seed, (uint32_t)(sh.seed >> 32));
    const bool direct = lockstep ? plan.direct_global : plan.ah_direct_global;
    if (!direct) {
        cudaError_t e;
        if (lockstep) {
            mm1_fused_kernel<true, FAST_THREADS, FAST_MIN_BLOCKS>
                <<<plan.blocks, FAST_THREADS, plan.smem_bytes, stream>>>(prm, rates, sh, out, plan.smem_ring, nullptr,
                                                                        nullptr, nullptr, nullptr);
            e = cudaGetLastError();
            ++*launches;
        } else if (plan.ah_spill) {
            // heavily loaded queue: two-level ring from the start; failures go straight to the fallback
            e = launch_mm1_ahead(prm, sh, out, plan, true, ahead_gring, ahead_state, resume, nullptr, nullptr, nullptr,
                                 sh.retry_list, sh.retry_count, stream);
            ++*launches;
        } else {
            // window-only first pass; what outgrows the window is redone with the two-level ring
            // (a lone replication takes ~4.5 ms there, 13 ms on the fallback kernel) -- by a helper grid that runs
            // NEXT TO the first pass (on the SM the first pass leaves one block short) and polls its list, so that the
            // one-in-a-million replication is done when the first pass ends instead of holding the step (and, at N
            // GPUs, every other rank) for those 4.5 ms.
            // Host order matters: first pass, then the memset that raises the flag behind it, then the helper --
            // if something serialises the launches (a profiler, CUDA_LAUNCH_BLOCKING) the helper finds the flag
            // set and the list complete instead of waiting for a kernel that has not been issued.
            if (overlap.helper_stream) {
                e = cudaMemsetAsync(sh.retry_a_list, 0xFF, sizeof(uint32_t) * sh.n_replications, stream);
                if (e == cudaSuccess) e = cudaMemsetAsync(overlap.done_flag, 0, sizeof(unsigned int), stream);
                if (e == cudaSuccess) e = cudaEventRecord(overlap.fork, stream);
                if (e == cudaSuccess) e = cudaStreamWaitEvent(overlap.helper_stream, overlap.fork, 0);
                if (e != cudaSuccess) return e;
            }
            // (one block short of the full grid when the helper follows: its warps need the registers of that block)
            e = launch_mm1_ahead(prm, sh, out, plan, false, ahead_gring, ahead_state, resume, nullptr, nullptr, nullptr,
                                 sh.retry_a_list, sh.retry_a_count, stream, overlap.helper_stream ? 1 : 0);
            ++*launches;
            if (e != cudaSuccess) return e;
            if (overlap.helper_stream) {
                e = cudaMemsetAsync(overlap.done_flag, 0xFF, sizeof(unsigned int), stream);
                if (e == cudaSuccess)
                    e = launch_mm1_ahead_helper(prm, sh, out, plan, ahead_gring, sh.retry_a_list, (uint32_t)sh.n_replications,
                                                sh.retry_a_cursor, overlap.done_flag, sh.retry_list, sh.retry_count,
                                                overlap.helper_stream);
                ++*launches;
                if (e == cudaSuccess) e = cudaEventRecord(overlap.join, overlap.helper_stream);
                if (e == cudaSuccess) e = cudaStreamWaitEvent(stream, overlap.join, 0);
                if (e != cudaSuccess) return e;
            }
            // what the helper left (a list longer than its few warps could take during the first pass)
            e = launch_mm1_ahead(prm, sh, out, plan, true, ahead_gring, nullptr, 0, sh.retry_a_list, sh.retry_a_count,
                                 sh.retry_a_cursor, sh.retry_list, sh.retry_count, stream);
            ++*launches;
        }
        if (e != cudaSuccess) return e;
        // retry list -> fallback kernel; exits at once when the list is empty
        mm1_fused_kernel<false, FB_THREADS, 1><<<plan.fb_blocks, FB_THREADS, table_bytes, stream>>>(
            prm, rates, sh, out, plan.fb_ring, fb_ring, sh.retry_list, sh.retry_count, fb_counter);
        ++*launches;
        return cudaGetLastError();
    }
    mm1_fused_kernel<false, FB_THREADS, 1><<<plan.fb_blocks, FB_THREADS, table_bytes, stream>>>(
        prm, rates, sh, out, plan.fb_ring, fb_ring, nullptr, nullptr, nullptr);
    ++*launches;
    return cudaGetLastError();
}

}  // namespace cxb
