// device/mm1_lane.cuh -- one replication of the M/M/1 model (examples/producer_consumer.cpp:9-59)
// as a per-lane state machine with the random variates drawn AHEAD of the event loop.
//
// Why ahead: the expensive part of an event is its variate (Philox block, log, divide, tick
// conversion); which process needs one differs from lane to lane and wake-ups need none, so a
// warp that draws "when the event asks for it" wastes lanes.  But the draws do not depend on the
// event order at all: the producer's inter-arrival times are draws 0,1,2,... of stream 1 and the
// service time the consumer draws after popping packet j is draw j of stream 2
// (producer_consumer.cpp:34,50).  So every lane generates packets in lock step -- two per
// iteration, one Philox block per stream yielding both draws of the block -- and writes
//      ring[j] = { A_j = tick at which the producer puts packet j,  s_j = service ticks }
// into its shared-memory ring.  The ring is the sync::queue<double> of the model extended by the
// packets that are generated but not yet put:  entries [k, i) are queued, [i, g) are ahead.
// The event loop (environment::step, environment.ipp:117-146) then only does bookkeeping:
// it takes A_{i+1} as the expiry of the producer's timeout and now + s_k as the consumer's.
//
// The code is __host__ __device__ on purpose: tests/native/mm1_lane_host.cpp compiles the same
// state machine with g++ and checks it against the oracle on the CPU (unit test of the device
// logic; the product has no CPU path -- the library only ever calls it from the kernel).
#pragma once
#include <stdint.h>

#include "cxxdes_b200.h"
#include "cxxdes_b200/shared/philox.hpp"
#include "cxxdes_b200/shared/trace_hash.hpp"
#include "cxxdes_b200/shared/variates.hpp"

namespace cxb {

using namespace cxxdes_b200;

namespace mm1 {

enum : uint32_t { M_MAIN = 0, M_PROD = 1, M_CONS = 2, M_HANDLER = 3, M_NOOP = 4 };

constexpr uint32_t EMPTY = 0xFFFFFFFFu;        // no token in this heap slot
constexpr uint32_t TIME_LIMIT = 1u << 30;      // every event time of a 32-bit replication is below this
constexpr uint32_t DRAW_LIMIT = 1u << 28;      // ... and every single delay below this

// (kind << 56) ^ (proc << 40) of trace_hash_step for each tag; every priority in this model is 0.
CXB_HD uint64_t tag_hash_bits(uint32_t m) {
    uint64_t kind = m == M_HANDLER ? TOKEN_HANDLER : (m == M_NOOP ? TOKEN_NOOP : TOKEN_RESUME);
    uint64_t proc = m == M_NOOP ? PROC_NONE : (m == M_HANDLER ? 0u : m);
    return (kind << 56) ^ (proc << 40);
}
constexpr uint64_t PROD_BITS = (uint64_t)M_PROD << 40;
constexpr uint64_t CONS_BITS = (uint64_t)M_CONS << 40;

CXB_HD uint64_t fold(uint64_t h, uint32_t t, uint64_t bits) {
    return (rotl64(h, 5) ^ ((uint64_t)t ^ bits)) * TRACE_HASH_MUL;
}

// double -> u32 ticks, truncating, saturating (environment::real_to_sim, environment.ipp:79-82,
// for values the range guard accepts; anything >= 2^28 trips the guard on both sides alike).
CXB_HD uint32_t ticks_u32(double x) {
#if defined(__CUDA_ARCH__)
    return __double2uint_rz(x);
#else
    return x >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)x;
#endif
}

// Launch-invariant values (kernel parameters: read straight from the constant bank).
struct constants {
    divisor lambda, mu;
    philox_schedule keys;
    double ticks_per_unit;   // (double)clock.ticks_per_unit
    double sec_per_unit;     // clock.seconds_per_tick_unit
    int64_t tick_mult;       // clock.tick_mult
    uint32_t n_packets;      // 2 <= n < 2^30
    uint32_t ring_cap;       // packets the ring may hold, 2 <= ring_cap <= physical slots
    uint32_t until;          // run_until horizon in ticks, EMPTY when there is none
};

struct ring_entry {
    uint32_t a;  // put tick of the packet
    uint32_t s;  // service ticks the consumer sleeps after popping it
};

// One replication.  `Ring` gives ring_entry load(j) / uint32_t load_a(j) / store(j, e) on slot
// j modulo its physical size.
struct lane {
    uint32_t ctr2, ctr3;        // Philox counter words 2,3 = global replication id
    uint32_t g;                 // packets generated so far (even)
    uint32_t a_gen;             // A_g
    uint32_t budget;            // A_g + sum of s_j, j < g: bounds every event time from above
    uint32_t i, k;              // packets put / popped
    uint32_t t0, t1;            // two-slot heap; in the steady state t0 = producer's token,
    uint32_t m0, m1;            //   t1 = consumer's timeout (EMPTY while it waits on queue.event_)
    uint32_t now, x_tick;
    uint32_t n_events;
    uint32_t max_queue;
    uint32_t status;
    uint32_t main_pc, cons_pc, remaining;
    uint64_t hash;
    double total_latency;
    bool waiting;               // consumer parked in queue.event_ (event.hpp:136-139); generic mode
    bool fast;                  // steady-state representation
    bool cfirst;                // steady state: on equal times the consumer's token pops first

    CXB_HD void init(uint64_t global_replication) {
        ctr2 = (uint32_t)global_replication;
        ctr3 = (uint32_t)(global_replication >> 32);
        g = 0; a_gen = 0; budget = 0;
        i = k = 0;
        t0 = 0; m0 = M_MAIN;     // environment::bind(co_main): start token {0, 0, main}
        t1 = EMPTY; m1 = M_NOOP;
        now = 0; x_tick = 0;
        n_events = 0; max_queue = 0; status = 0;
        main_pc = cons_pc = 0; remaining = 2;
        hash = TRACE_HASH_INIT;
        total_latency = 0.0;
        waiting = false; fast = false; cfirst = false;
    }

    // ---- generation -----------------------------------------------------------------
    CXB_HD bool wants_packets(const constants &c) const { return g < c.n_packets; }
    CXB_HD bool has_room(const constants &c) const { return g - k + 2 <= c.ring_cap; }

    // Packets g and g+1: inter-arrival draws g, g+1 of stream 1 and service draws g, g+1 of
    // stream 2 are the two halves of block g/2 of each stream (philox.hpp addressing).
    template <typename Ring, typename Table>
    CXB_HD void generate_pair(const constants &c, Ring &ring, const Table *table) {
        const philox_block bp = philox4x32_10_scheduled(g >> 1, 1u, ctr2, ctr3, c.keys);
        const philox_block bc = philox4x32_10_scheduled(g >> 1, 2u, ctr2, ctr3, c.keys);
        // env.timeout(exponential()): timeout.ipp:184-187 -> real_to_sim
        const uint32_t ia0 = ticks_u32(exponential_from_bits(((uint64_t)bp.w[1] << 32) | bp.w[0], c.lambda, table) * c.ticks_per_unit);
        const uint32_t ia1 = ticks_u32(exponential_from_bits(((uint64_t)bp.w[3] << 32) | bp.w[2], c.lambda, table) * c.ticks_per_unit);
        const uint32_t s0 = ticks_u32(exponential_from_bits(((uint64_t)bc.w[1] << 32) | bc.w[0], c.mu, table) * c.ticks_per_unit);
        const uint32_t s1 = ticks_u32(exponential_from_bits(((uint64_t)bc.w[3] << 32) | bc.w[2], c.mu, table) * c.ticks_per_unit);
        // range guard: each delay < 2^28 and the bound on the last event time < 2^30, so no
        // 32-bit sum below can wrap; otherwise the replication is redone with 64-bit ticks
        const uint32_t new_budget = budget + ia0 + ia1 + s0 + s1;
        if (((ia0 | ia1 | s0 | s1) >= DRAW_LIMIT) | (new_budget >= TIME_LIMIT)) {
            status |= CXXDES_B200_REP_TIME_RANGE;
            return;
        }
        budget = new_budget;
        const uint32_t a1 = a_gen + ia0;
        ring_entry e0, e1;
        e0.a = a_gen; e0.s = s0;
        e1.a = a1; e1.s = s1;
        ring.store(g, e0);
        ring.store(g + 1, e1);
        a_gen = a1 + ia1;
        g += 2;
    }

    // The producer's event is next, its packet is not generated and the ring has no room: the
    // model's queue is longer than this ring.
    CXB_HD bool ring_exhausted(const constants &c) const {
        if (!wants_packets(c) || has_room(c) || i != g) return false;
        if (fast) return !((t1 < t0) | ((t1 == t0) & cfirst));
        return m0 == M_PROD;
    }

    // ---- time of the next scheduled token -------------------------------------------------
    CXB_HD uint32_t next_time() const { return (fast && t1 < t0) ? t1 : t0; }

    CXB_HD double seconds(const constants &c, uint32_t tick) const {
        // environment::now_seconds, environment.ipp:29-36
        if (c.tick_mult == 1) return (double)tick * c.sec_per_unit;
        return (double)((int64_t)tick * c.tick_mult) * c.sec_per_unit;
    }

    // ---- steady state: one event (two when a put wakes the consumer) ---------------------------
    // Exactly one producer token (t0) and at most one consumer token (t1, always a timeout).
    // With <= 2 entries the libstdc++ heap keeps the earlier push in front on equal keys
    // (stl_heap.h:135-148), which `cfirst` records.  Returns false when nothing could be done.
    template <typename Ring>
    CXB_HD bool fast_slot(const constants &c, Ring &ring) {
        const uint32_t tp = t0, tc = t1;
        const bool c_goes = (tc < tp) | ((tc == tp) & cfirst);   // parked: tc == EMPTY > tp
        const uint32_t t = c_goes ? tc : tp;
        // the producer's event needs A_{i+1}, i.e. packet i generated; run_until stops at `until`
        if (!(fast & (c_goes | (i < g)) & (t <= c.until))) return false;
        // environment::step: pop, advance the clock, count, fold (tokens are scheduled at
        // now + d with d >= 0, so time never runs backwards here)
        now = t;
        ++n_events;
        hash = fold(hash, t, c_goes ? CONS_BITS : PROD_BITS);
        bool attempt = c_goes;  // the consumer runs its pop step
        if (c_goes) {
            // total_latency += now_seconds() - x   (producer_consumer.cpp:52)
            total_latency += seconds(c, t) - seconds(c, x_tick);
        } else {
            // q.put(now_seconds()) (queue.hpp:48-53): packet i, already in the ring with a == now
            ++i;
            const uint32_t count = i - k;
            max_queue = count > max_queue ? count : max_queue;
            const uint32_t next_a = ring.load_a(i);
            t0 = i < g ? next_a : a_gen;  // env.timeout(exponential()) expires at A_{i+1}
            if (tc == EMPTY) {
                // wake(): the parked consumer's token {0 + now} is pushed before the producer's
                // timeout, so it is the very next event (event.hpp:125-134); handled here
                ++n_events;
                hash = fold(hash, t, CONS_BITS);
                attempt = true;
            } else {
                cfirst = true;  // the producer's new token is behind the consumer's on a tie
            }
        }
        if (attempt) {
            if (i != k) {
                // q.pop() succeeds (queue.hpp:58-65); timeout(exponential()) = now + s_k, pushed last
                const ring_entry e = ring.load(k);
                x_tick = e.a;
                t1 = t + e.s;
                cfirst = false;
                ++k;
            } else {
                t1 = EMPTY;  // the consumer waits on queue.event_
            }
        }
        if (i + 1 >= c.n_packets) leave_fast();
        return true;
    }

    // The producer's last put ends the steady state: rebuild the generic two-slot heap.
    CXB_HD void leave_fast() {
        const uint32_t tp = t0, tc = t1;
        const bool c_front = (tc < tp) | ((tc == tp) & cfirst);
        waiting = tc == EMPTY;
        t0 = c_front ? tc : tp;
        m0 = c_front ? M_CONS : M_PROD;
        t1 = c_front ? tp : tc;
        m1 = c_front ? M_PROD : M_CONS;
        cons_pc = waiting ? 1u : 2u;
        fast = false;
    }

    CXB_HD void push(uint32_t t, uint32_t m) {
        // std::push_heap on <= 2 entries: the new token goes in front only if the current front
        // is strictly later (stl_heap.h:135-148); an empty front compares as latest.
        if (t1 != EMPTY) status |= CXXDES_B200_REP_HEAP_OVERFLOW;
        const bool front = t0 > t;
        t1 = front ? t0 : t;
        m1 = front ? m0 : m;
        t0 = front ? t : t0;
        m0 = front ? m : m0;
    }

    // ---- start-up and wind-down: environment::step on the explicit two-slot heap ---------------
    template <typename Ring>
    CXB_HD bool generic_step(const constants &c, Ring &ring) {
        if (fast || t0 == EMPTY || t0 > c.until) return false;
        const uint32_t n = c.n_packets;
        if (m0 == M_PROD && i < n && i >= g) return false;  // A_{i+1} not generated yet
        const uint32_t t = t0;
        const uint32_t m = m0;
        t0 = t1; m0 = m1; t1 = EMPTY; m1 = M_NOOP;   // pop_heap on <= 2 entries
        now = t > now ? t : now;                     // environment.ipp:126
        ++n_events;
        hash = fold(hash, t, tag_hash_bits(m));
        if (m == M_PROD) {
            // producer body, producer_consumer.cpp:31-37
            if (i < n) {
                ++i;                                 // q.put(now_seconds()): ring entry i-1, a == now
                const uint32_t count = i - k;
                max_queue = count > max_queue ? count : max_queue;
                // wake(): the parked consumer's token goes in before the producer's own timeout
                if (waiting) { push(now, M_CONS); waiting = false; }
                const uint32_t next_a = ring.load_a(i);
                push(i < g ? next_a : a_gen, M_PROD);
            } else {
                push(now, M_HANDLER);                // co_return -> completion carrying the all_of handler
            }
        } else if (m == M_CONS) {
            // consumer body, producer_consumer.cpp:39-54
            if (cons_pc == 2) total_latency += seconds(c, now) - seconds(c, x_tick);  // :52
            if (i != k) {
                const ring_entry e = ring.load(k);   // q.pop() succeeds, queue.hpp:58-65
                x_tick = e.a;
                ++k;
                cons_pc = 2;
                if (k == n) push(now, M_HANDLER);    // consumer returns (:45-48)
                else push(now + e.s, M_CONS);
            } else {
                waiting = true;                      // ... or waits on queue.event_
                cons_pc = 1;
            }
        } else if (m == M_HANDLER) {
            // all_of handler, any_of.ipp:9-26: the output token inherits the child token's time
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
        // both processes running, only their tokens scheduled, the consumer parked or asleep in
        // its service timeout: switch to the steady-state representation
        if (status == 0 && main_pc == 1 && remaining == 2 && i >= 1 && i + 1 < n && t0 != EMPTY) {
            const bool prod_front = m0 == M_PROD && (t1 == EMPTY ? waiting : (m1 == M_CONS && cons_pc == 2));
            const bool cons_front = m0 == M_CONS && cons_pc == 2 && t1 != EMPTY && m1 == M_PROD;
            if (prod_front || cons_front) {
                const uint32_t tp = prod_front ? t0 : t1;
                const uint32_t tc = prod_front ? t1 : t0;
                t0 = tp;
                t1 = tc;
                cfirst = cons_front;
                fast = true;
            }
        }
        return true;
    }

    // environment::run / run_until (environment.ipp:179-196): nothing left, or nothing due.
    CXB_HD bool finished(const constants &c) const {
        const uint32_t next = next_time();
        return status != 0 || next == EMPTY || next > c.until;
    }
};

}  // namespace mm1
}  // namespace cxb
