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
// (both stored as heap keys, see `lane`)
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

// PTX text of one steady-state slot (lane::fast_slot_ptx below; the commented C++ twin is
// lane::fast_slot).  Operands: %0 kp, %1 kc, %2 i, %3 k, %4 n_events, %5 max_queue, %6 x_tick,
// %7 hash, %8 total_latency | %9 glim, %10 ring base, %11 sec_per_unit, %12 4*until+3,
// %13 tick_mult, %14 slot mask (= window size), %15 slot stride in bytes, %16/%17 epoch (low/high
// word; wide clocks) | spilling ring: %18 g, %19 global ring column, %20 its slot stride in bytes,
// %21 its slot mask.
//   A    front token; can the lane dispatch (pc: consumer resumes, pp: producer puts)
//   DUE  run_until horizon
//   B    woke?, ring[i+1].ak and ring[k], environment::step (fold) once or, when the put wakes
//        the consumer, twice; now and x as doubles
//   MULT time_precision().t
//   C    total_latency += now_seconds() - x; put: i, max_queue, kp, tie bit; pop or park
#define CXB_MM1_SLOT_A \
    "{\n\t" \
    ".reg .pred pc, pp, pw, pok, pa, ppop, pdue;\n\t" \
    ".reg .b32 key, t, i1, sl, addr, akn, eak, es4, cnt, hl, hh, a0, a1, b0, b1, rl, rh, bh, mh, ts;\n\t" \
    ".reg .b64 m;\n\t" \
    ".reg .f64 dt, dx, lat;\n\t" \
    "setp.lt.u32 pc, %1, %0;\n\t" \
    "min.u32 key, %1, %0;\n\t" \
    "setp.lt.and.u32 pp, %2, %9, !pc;\n\t"
#define CXB_MM1_SLOT_DUE \
    "setp.le.u32 pdue, key, %12;\n\t" \
    "and.pred pc, pc, pdue;\n\t" \
    "and.pred pp, pp, pdue;\n\t"
// Instruction selection knobs (measured on B200): the integer ALU pipe takes a warp
// instruction every 2 cycles; IMAD runs on the FMA-heavy pipe at 2 cycles, IMAD.WIDE / IMAD.HI at 4.
#ifndef CXB_MM1_ROT_MUL
#define CXB_MM1_ROT_MUL 0     /* rotl64(h, 5): 1 = two 32x32->64 multiplies by 32, 0 = two funnel shifts */
#endif
#ifndef CXB_MM1_SHIFT_MUL
#define CXB_MM1_SHIFT_MUL 0   /* x >> 2: 1 = mulhi(x, 2^30), 0 = shift */
#endif
#if CXB_MM1_ROT_MUL
#define CXB_MM1_SLOT_ROT(TLOW, BITS) \
    "mul.wide.u32 m, hl, 32;\n\t" \
    "mov.b64 {a0, a1}, m;\n\t" \
    "mul.wide.u32 m, hh, 32;\n\t" \
    "mov.b64 {b0, b1}, m;\n\t" \
    "lop3.b32 rl, a0, b1, " TLOW ", 0x56;\n\t" \
    "lop3.b32 rh, b0, a1, " BITS ", 0x56;\n\t"
#else
#define CXB_MM1_SLOT_ROT(TLOW, BITS) \
    "shf.l.wrap.b32 a1, hl, hh, 5;\n\t" \
    "shf.l.wrap.b32 a0, hh, hl, 5;\n\t" \
    "xor.b32 rl, a0, " TLOW ";\n\t" \
    "xor.b32 rh, a1, " BITS ";\n\t"
#endif
#if CXB_MM1_SHIFT_MUL
#define CXB_MM1_SLOT_SHR2(PRED, DST, SRC) PRED "mul.hi.u32 " DST ", " SRC ", 0x40000000;\n\t"
#else
#define CXB_MM1_SLOT_SHR2(PRED, DST, SRC) PRED "shr.u32 " DST ", " SRC ", 2;\n\t"
#endif
#define CXB_MM1_SLOT_FOLD(PRED, TLOW, BITS) \
    CXB_MM1_SLOT_ROT(TLOW, BITS) \
    "mul.hi.u32 mh, rl, 0x7F4A7C15;\n\t" \
    "mad.lo.u32 mh, rl, 0x9E3779B9, mh;\n\t" \
    "@" PRED " mad.lo.u32 hh, rh, 0x7F4A7C15, mh;\n\t" \
    "@" PRED " mul.lo.u32 hl, rl, 0x7F4A7C15;\n\t"
/* The time of the event as the digest and now_seconds() see it: the tick itself (NARROW), or
   epoch + tick as a 64-bit value (WIDE: fine time precision, see lane::epoch). */
#define CXB_MM1_SLOT_TABS_NARROW ""
#define CXB_MM1_SLOT_TABS_WIDE \
    ".reg .b32 tl, th, bw, tsl, tsh, xl, xh;\n\t" \
    "add.cc.u32 tl, t, %16;\n\t" \
    "addc.u32 th, 0, %17;\n\t" \
    "xor.b32 bw, th, 0x200;\n\t"
#define CXB_MM1_SLOT_DTDX_NARROW \
    "mov.b64 dt, {ts, 0x43300000};\n\t" \
    "mov.b64 dx, {%6, 0x43300000};\n\t"
#define CXB_MM1_SLOT_DTDX_WIDE \
    "add.cc.u32 tsl, ts, %16;\n\t" \
    "addc.u32 tsh, 0, %17;\n\t" \
    "or.b32 tsh, tsh, 0x43300000;\n\t" \
    "add.cc.u32 xl, %6, %16;\n\t" \
    "addc.u32 xh, 0, %17;\n\t" \
    "or.b32 xh, xh, 0x43300000;\n\t" \
    "mov.b64 dt, {tsl, tsh};\n\t" \
    "mov.b64 dx, {xl, xh};\n\t"
#define CXB_MM1_SLOT_B_ANY(LOAD_K, TABS, TLOW, BH_FIX, BITS2, DTDX) \
    "setp.eq.and.u32 pw, %1, 0xffffffff, pp;\n\t" \
    CXB_MM1_SLOT_SHR2("", "t", "key") \
    TABS \
    "add.u32 i1, %2, 1;\n\t" \
    "and.b32 sl, i1, %14;\n\t" \
    "mad.lo.u32 addr, sl, %15, %10;\n\t" \
    "ld.shared.u32 akn, [addr];\n\t" \
    "and.b32 sl, %3, %14;\n\t" \
    "mad.lo.u32 addr, sl, %15, %10;\n\t" \
    LOAD_K \
    "or.pred pok, pc, pp;\n\t" \
    "mov.b64 {hl, hh}, %7;\n\t" \
    "selp.b32 bh, 0x200, 0x100, pc;\n\t" \
    BH_FIX \
    CXB_MM1_SLOT_FOLD("pok", TLOW, "bh") \
    CXB_MM1_SLOT_FOLD("pw", TLOW, BITS2) \
    "mov.b64 %7, {hl, hh};\n\t" \
    "selp.b32 ts, t, %6, pc;\n\t" \
    DTDX \
    "sub.rn.f64 dt, dt, 0d4330000000000000;\n\t" \
    "sub.rn.f64 dx, dx, 0d4330000000000000;\n\t"
#define CXB_MM1_SLOT_B(LOAD_K) \
    CXB_MM1_SLOT_B_ANY(LOAD_K, CXB_MM1_SLOT_TABS_NARROW, "t", "", "0x200", CXB_MM1_SLOT_DTDX_NARROW)
#define CXB_MM1_SLOT_B_WIDE(LOAD_K) \
    CXB_MM1_SLOT_B_ANY(LOAD_K, CXB_MM1_SLOT_TABS_WIDE, "tl", "xor.b32 bh, bh, th;\n\t", "bw", CXB_MM1_SLOT_DTDX_WIDE)
/* ring[k]: in the shared-memory window, or (spilling ring) in global memory once the queue is
   longer than the window: entry k is in the window iff g - k <= WINDOW */
#define CXB_MM1_SLOT_LOAD_K "ld.shared.v2.u32 {eak, es4}, [addr];\n\t"
#define CXB_MM1_SLOT_LOAD_K_SPILL \
    "{\n\t" \
    ".reg .pred pin;\n\t" \
    ".reg .b32 d, gs;\n\t" \
    ".reg .b64 ga;\n\t" \
    "sub.u32 d, %18, %3;\n\t" \
    "setp.le.u32 pin, d, %14;\n\t" \
    "and.b32 gs, %3, %21;\n\t" \
    "mad.wide.u32 ga, gs, %20, %19;\n\t" \
    "@pin ld.shared.v2.u32 {eak, es4}, [addr];\n\t" \
    "@!pin ld.global.v2.u32 {eak, es4}, [ga];\n\t" \
    "}\n\t"
#define CXB_MM1_SLOT_MULT \
    "mul.rn.f64 dt, dt, %13;\n\t" \
    "mul.rn.f64 dx, dx, %13;\n\t"
#define CXB_MM1_SLOT_C \
    "mul.rn.f64 dt, dt, %11;\n\t" \
    "mul.rn.f64 dx, dx, %11;\n\t" \
    "sub.rn.f64 lat, dt, dx;\n\t" \
    "add.rn.f64 %8, %8, lat;\n\t" \
    "@pp mov.b32 %2, i1;\n\t" \
    "sub.u32 cnt, %2, %3;\n\t" \
    "max.u32 %5, %5, cnt;\n\t" \
    "@pp mov.b32 %0, akn;\n\t" \
    "@pp and.b32 %1, %1, 0xfffffffd;\n\t" \
    "or.pred pa, pc, pw;\n\t" \
    "setp.ne.and.u32 ppop, %2, %3, pa;\n\t" \
    "@pa add.u32 %4, %4, 1;\n\t" \
    "@pa mov.b32 %1, 0xffffffff;\n\t" \
    "@ppop mad.lo.u32 %1, t, 4, es4;\n\t" \
    CXB_MM1_SLOT_SHR2("@ppop ", "%6", "eak") \
    "@ppop add.u32 %3, %3, 1;\n\t" \
    "}"
#define CXB_MM1_SLOT_OPERANDS \
    : "+r"(kp), "+r"(kc), "+r"(i), "+r"(k), "+r"(n_events), "+r"(max_queue), "+r"(x_tick), \
    "+l"(hash), "+d"(total_latency) \
    : "r"(glim), "r"(base), "d"(c.sec_per_unit), "r"(ukey), "d"(mult), "n"(MASK), "n"(STRIDE), "n"(0), "n"(0)
#define CXB_MM1_SLOT_OPERANDS_SPILL CXB_MM1_SLOT_OPERANDS, "r"(g), "l"(gbase), "r"(gstride), "r"(gmask)
/* wide clocks: %16/%17 are the epoch (narrow: two immediates that no text refers to, so that the
   state of a narrow lane does not carry two more live registers into every slot) */
#define CXB_MM1_SLOT_OPERANDS_WIDE \
    : "+r"(kp), "+r"(kc), "+r"(i), "+r"(k), "+r"(n_events), "+r"(max_queue), "+r"(x_tick), \
    "+l"(hash), "+d"(total_latency) \
    : "r"(glim), "r"(base), "d"(c.sec_per_unit), "r"(ukey), "d"(mult), "n"(MASK), "n"(STRIDE), "r"(ep_lo), "r"(ep_hi)
#define CXB_MM1_SLOT_OPERANDS_WIDE_SPILL CXB_MM1_SLOT_OPERANDS_WIDE, "r"(g), "l"(gbase), "r"(gstride), "r"(gmask)

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

CXB_HD uint64_t fold(uint64_t h, uint64_t t, uint64_t bits) {
    return (rotl64(h, 5) ^ (t ^ bits)) * TRACE_HASH_MUL;
}

// Integer <-> double conversions without the XU pipe (I2F/F2I issue at a quarter of the FP64 rate):
// adding 2^52 puts an integer below 2^52 into the low mantissa bits, exactly.
constexpr double TWO52 = 4503599627370496.0;

CXB_HD double u32_to_double(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(0x43300000, (int)x) - TWO52;
#else
    return (double)x;
#endif
}

CXB_HD double u64_to_double(uint64_t x) {   // x < 2^52
#if defined(__CUDA_ARCH__)
    return __hiloint2double((int)(0x43300000u | (uint32_t)(x >> 32)), (int)(uint32_t)x) - TWO52;
#else
    return (double)x;
#endif
}

// double -> ticks, truncating toward zero (environment::real_to_sim, environment.ipp:79-82) for
// 0 <= x < 2^28.  No per-draw check: a variate is at most -log(2^-52) / rate = 36.05 / rate, and
// the launcher only uses this representation when 36.05 * ticks_per_unit / rate < 2^28 for both
// rates (draw_fits_u32 below); everything else runs on the 64-bit kernel.
CXB_HD uint32_t ticks_u32(double x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__double2loint(__dadd_rz(x, TWO52));  // round toward zero = truncate
#else
    return (uint32_t)x;
#endif
}

inline bool draw_fits_u32(double rate, double ticks_per_unit) {
    return 36.05 * ticks_per_unit / rate < (double)DRAW_LIMIT;
}

// Launch-invariant values (kernel parameters: read straight from the constant bank).
struct constants {
    divisor lambda, mu;
    philox_schedule keys;
    double ticks_per_unit;   // (double)clock.ticks_per_unit
    double sec_per_unit;     // clock.seconds_per_tick_unit
    double tick_mult;        // (double)clock.tick_mult, < 2^23 so tick * tick_mult is exact in a double
    uint32_t n_packets;      // 2 <= n < 2^30
    uint32_t ring_cap;       // packets (queued + generated ahead) the shared-memory window may hold, >= 2
    uint32_t spill_cap;      // ... and the ring with its global-memory level (spilling path)
    uint32_t lead_cap;       // spilling path: packets that may be generated ahead of the producer
    uint32_t until;          // run_until horizon in ticks, EMPTY when there is none (or beyond 2^30)
    int64_t until64;         // the same horizon as given, < 0 when there is none (MODE_WIDE)
};

// Compile-time variants of the lane code.
enum : int {
    MODE_UNTIL = 1,      // a run_until horizon is set (else the check is compiled out)
    MODE_TICK_MULT = 2,  // time_precision().t != 1: now_seconds needs the extra multiply
    MODE_WIDE = 4        // the clock passes 2^30 ticks: 32-bit times are relative to a per-lane epoch
};
constexpr uint32_t REBASE_LIMIT = 1u << 29;   // MODE_WIDE: move the epoch when the bound on the live times passes this

// Steady-state tokens are compared as one 32-bit key that carries the tie rule of the two-entry
// libstdc++ heap ("equal (time, priority): the earlier push pops first", stl_heap.h:135-148):
//      producer's token          4 * time + 1
//      consumer's token          4 * time + 2     pushed after the producer's current token
//                                4 * time + 0     pushed before it (the producer re-armed since)
// so "the consumer's token pops next" is kc < kp.  Times stay below 2^30 (range guard).
struct ring_entry {
    uint32_t ak;  // 4 * A_j + 1: the producer's token that puts packet j
    uint32_t s4;  // 4 * s_j + 2: added to 4 * now when the consumer starts serving packet j
};

// One replication.  `Ring` gives ring_entry load(j, g) / uint32_t load_a(j) / store_if(p, j, e) /
// store_a_if(p, j, ak) / spill_if(p, g) / store_level(j, g, e) (rewrite entry j where load(j, g) finds it).  Its first level is a window of WINDOW + 1 slots (j modulo
// that size) holding the entries [g - WINDOW, g) and the arrival key of entry g; a spilling ring
// (SPILL) keeps older queued entries in a second, larger level.
struct lane {
    uint32_t ctr2, ctr3;        // Philox counter words 2,3 = global replication id
    uint32_t g;                 // packets generated so far (even)
    uint32_t agk;               // 4 * A_g + 1
    uint32_t budget;            // A_g + sum of s_j, j < g: bounds every event time from above
    uint32_t i, k;              // packets put / popped
    // steady state: exactly one producer token and at most one consumer token (its service timeout)
    uint32_t kp, kc;            // their keys; kc == EMPTY while the consumer waits on queue.event_;
    uint32_t glim;              //   kp == kc == EMPTY and glim == 0 outside the steady state, else glim == g
    // start-up / wind-down: explicit two-slot heap of (time, tag)
    uint32_t t0, t1, m0, m1;
    uint32_t main_pc, cons_pc, remaining, waiting;
    uint32_t now, x_tick;
    uint32_t n_events;          // outside the steady state: events so far; inside: events - i (see events())
    uint32_t max_queue;
    uint32_t status;
    uint64_t hash;
    double total_latency;
    // MODE_WIDE (fine time precision: the clock passes 2^30 ticks although the live tokens span far
    // less): every 32-bit time / key above is relative to `epoch`, which is moved forward
    // (rebase()) whenever the bound on the live times passes 2^29.  Otherwise epoch stays 0.
    uint64_t epoch;
    uint32_t until_rel;         // run_until horizon relative to epoch, EMPTY when none / out of reach

    CXB_HD void init(uint64_t global_replication, const constants &c) {
        epoch = 0;
        until_rel = c.until;
        ctr2 = (uint32_t)global_replication;
        ctr3 = (uint32_t)(global_replication >> 32);
        g = 0; agk = 1; budget = 0;
        i = k = 0;
        kp = kc = EMPTY; glim = 0;
        t0 = 0; m0 = M_MAIN;     // environment::bind(co_main): start token {0, 0, main}
        t1 = EMPTY; m1 = M_NOOP;
        main_pc = cons_pc = 0; remaining = 2; waiting = 0;
        now = 0; x_tick = 0;
        n_events = 0; max_queue = 0; status = 0;
        hash = TRACE_HASH_INIT;
        total_latency = 0.0;
    }

    // A lane without a replication: every step below is a no-op for it.
    CXB_HD void idle() {
        kp = kc = EMPTY; glim = 0;
        t0 = t1 = EMPTY;
        g = 0xFFFFFFFEu;
    }

    CXB_HD bool is_fast() const { return kp != EMPTY; }

    // environment::step() calls so far.  In the steady state every put is one event, so only the
    // consumer's events are counted there and the puts are taken from i.
    CXB_HD uint32_t events() const { return is_fast() ? n_events + i : n_events; }

    // ---- generation -----------------------------------------------------------------
    CXB_HD bool wants_packets(const constants &c) const { return g < c.n_packets; }
    // (one slot beyond the packets holds the next arrival time: ring_cap < physical slots)
    template <bool SPILL>
    CXB_HD bool has_room(const constants &c) const {
        // (sums, not differences: while a resumed lane regenerates its ring g may be below i, even k - 1)
        if (SPILL) return (g + 2 <= k + c.spill_cap) & (g + 2 <= i + c.lead_cap);
        return g + 2 <= k + c.ring_cap;  // the window bounds the lead as well (i >= k)
    }

    // Packets g and g+1: inter-arrival draws g, g+1 of stream 1 and service draws g, g+1 of
    // stream 2 are the two halves of block g/2 of each stream (philox.hpp addressing).
    // The arithmetic runs for every lane and only the commit is conditional (`enable`), so there
    // is no branch here: the compiler may interleave these four long FP64 chains with the
    // integer bookkeeping of the event slots in the same basic block.
    // ANCHOR_SECOND (only when a saved replication is resumed at an odd k): agk is the key of
    // packet g + 1 instead of packet g.
    template <int MODE, bool ANCHOR_SECOND = false, typename Ring, typename Table>
    CXB_HD void generate_pair(const constants &c, Ring &ring, const Table *table, bool enable) {
        const philox_block bp = philox4x32_10_scheduled(g >> 1, 1u, ctr2, ctr3, c.keys);
        const philox_block bc = philox4x32_10_scheduled(g >> 1, 2u, ctr2, ctr3, c.keys);
        // env.timeout(exponential()): timeout.ipp:184-187 -> real_to_sim
        const uint32_t ia0 = ticks_u32(exponential_from_bits(((uint64_t)bp.w[1] << 32) | bp.w[0], c.lambda, table) * c.ticks_per_unit);
        const uint32_t ia1 = ticks_u32(exponential_from_bits(((uint64_t)bp.w[3] << 32) | bp.w[2], c.lambda, table) * c.ticks_per_unit);
        const uint32_t s0 = ticks_u32(exponential_from_bits(((uint64_t)bc.w[1] << 32) | bc.w[0], c.mu, table) * c.ticks_per_unit);
        const uint32_t s1 = ticks_u32(exponential_from_bits(((uint64_t)bc.w[3] << 32) | bc.w[2], c.mu, table) * c.ticks_per_unit);
        // range guard: each delay < 2^28 (draw_fits_u32) and the bound on the last event time
        // < 2^30, so no 32-bit sum or key below can wrap; otherwise the replication is redone with
        // 64-bit ticks
        uint32_t new_budget = budget + ia0 + ia1 + s0 + s1;
        if constexpr ((MODE & MODE_WIDE) != 0) {
            // relative times: move the epoch up to the oldest live time and retry (rare: once per
            // ~2^29 ticks of simulated time); only a window that itself spans 2^30 ticks is out of range
            if (enable && new_budget >= REBASE_LIMIT) {
                rebase<MODE>(c, ring);
                new_budget = budget + ia0 + ia1 + s0 + s1;
            }
        }
        const bool out_of_range = new_budget >= TIME_LIMIT;
        const bool commit = enable & !out_of_range;
        status |= (enable & out_of_range) ? (uint32_t)CXXDES_B200_REP_TIME_RANGE : 0u;
        const uint32_t a0k = ANCHOR_SECOND ? agk - 4u * ia0 : agk;
        const uint32_t a1k = a0k + 4u * ia0;
        const uint32_t a2k = a1k + 4u * ia1;
        ring_entry e0, e1;
        e0.ak = a0k; e0.s4 = 4u * s0 + 2u;
        e1.ak = a1k; e1.s4 = 4u * s1 + 2u;
        // spilling ring: the two oldest window entries are about to be overwritten; if they are
        // still queued they move to global memory first
        if (Ring::SPILL) ring.spill_if(commit & (g + 2 > k + (uint32_t)Ring::WINDOW), g);
        ring.store_if(commit, g, e0);
        ring.store_if(commit, g + 1, e1);
        ring.store_a_if(commit, g + 2, a2k);   // A_{g+2}: the producer's token after it puts packet g+1
        budget = commit ? new_budget : budget;
        agk = commit ? a2k : agk;
        g += commit ? 2u : 0u;
        glim = is_fast() ? g : 0u;
    }

    // ---- MODE_WIDE: absolute time of a relative tick, and moving the epoch -----------------------
    // the run_until horizon as the lane's relative ticks (EMPTY: none)
    template <int MODE>
    CXB_HD uint32_t horizon(const constants &c) const {
        return (MODE & MODE_WIDE) ? until_rel : c.until;
    }

    template <int MODE>
    CXB_HD uint64_t abs_time(uint32_t tick) const {
        return (MODE & MODE_WIDE) ? epoch + tick : (uint64_t)tick;
    }

    CXB_HD void set_until_rel(const constants &c) {
        if (c.until64 < 0 || (uint64_t)c.until64 < epoch || (uint64_t)c.until64 - epoch >= TIME_LIMIT) until_rel = EMPTY;
        else until_rel = (uint32_t)((uint64_t)c.until64 - epoch);
        // (a horizon already behind the epoch cannot happen: the lane stops at the horizon)
    }

    // Subtract the oldest live time from every relative time and key of the lane (tokens, ring
    // entries, clock, generation frontier) and add it to the epoch; the bound on the live times is
    // recomputed exactly: latest live time + the service times still queued.
    template <int MODE, typename Ring>
    CXB_HD void rebase(const constants &c, Ring &ring) {
        uint32_t lo = agk >> 2, hi = agk >> 2;
        auto see = [&](uint32_t t) { lo = t < lo ? t : lo; hi = t > hi ? t : hi; };
        bool serving;
        if (is_fast()) {
            see(kp >> 2);
            serving = kc != EMPTY;
            if (serving) see(kc >> 2);
        } else {
            if (t0 != EMPTY) see(t0);
            if (t1 != EMPTY) see(t1);
            see(now);
            serving = cons_pc == 2;
        }
        if (serving) see(x_tick);
        uint32_t sum_s = 0;
        for (uint32_t j = k; j < g; ++j) {
            const ring_entry e = ring.load(j, g);
            if (j == k) see(e.ak >> 2);
            sum_s += e.s4 >> 2;
        }
        const uint32_t d = lo, d4 = 4u * lo;
        for (uint32_t j = k; j < g; ++j) {
            ring_entry e = ring.load(j, g);
            e.ak -= d4;
            ring.store_level(j, g, e);
        }
        agk -= d4;
        ring.store_a_if(true, g, agk);
        if (is_fast()) {
            kp -= d4;
            if (kc != EMPTY) kc -= d4;
        } else {
            if (t0 != EMPTY) t0 -= d;
            if (t1 != EMPTY) t1 -= d;
        }
        now = now > d ? now - d : 0u;          // (stale in the steady state, live otherwise)
        x_tick = x_tick > d ? x_tick - d : 0u; // (stale unless a packet is in service)
        epoch += d;
        budget = (hi - d) + sum_s;
        set_until_rel(c);
        // absolute ticks (times tick_mult) must stay exact in a double
        if ((double)(epoch + TIME_LIMIT) * c.tick_mult >= TWO52) status |= CXXDES_B200_REP_TIME_RANGE;
    }

    // ---- run_until / run_for in pieces: state of a replication between two launches ---------------
    // (environment::run_until returns with the tokens still scheduled, environment.ipp:190-196;
    // the next call continues from there.)  Everything but the ring is saved; the ring entries
    // [k, g) are a pure function of (seed, replication, k, A_k), so the resumed lane draws them
    // again -- 8 words fewer than saving a window that may have a second level in global memory.
    static constexpr int STATE_WORDS = 32;   // one 128-byte record per replication
    enum : uint32_t { PHASE_NEW = 0, PHASE_RUNNING = 1, PHASE_DONE = 2, PHASE_REPLAY = 3 };

    template <typename Ring>
    CXB_HD void save(uint32_t *w, Ring &ring, uint32_t phase) const {
        w[0] = phase;
        w[1] = k < g ? ring.load(k, g).ak : agk;   // 4 * A_k + 1
        w[2] = budget; w[3] = i; w[4] = k; w[5] = kp; w[6] = kc;
        w[7] = t0; w[8] = t1; w[9] = m0; w[10] = m1;
        w[11] = main_pc; w[12] = cons_pc; w[13] = remaining; w[14] = waiting;
        w[15] = now; w[16] = x_tick; w[17] = n_events; w[18] = max_queue; w[19] = status;
        w[20] = (uint32_t)hash; w[21] = (uint32_t)(hash >> 32);
        const uint64_t tl = double_as_bits(total_latency);
        w[22] = (uint32_t)tl; w[23] = (uint32_t)(tl >> 32);
        w[24] = (uint32_t)epoch; w[25] = (uint32_t)(epoch >> 32);
    }

    // Counterpart of save(); leaves g <= k with agk anchored at packet k, regenerate() refills.
    CXB_HD void restore(const uint32_t *w, uint64_t global_replication, const constants &c) {
        ctr2 = (uint32_t)global_replication;
        ctr3 = (uint32_t)(global_replication >> 32);
        agk = w[1]; budget = w[2]; i = w[3]; k = w[4]; kp = w[5]; kc = w[6];
        t0 = w[7]; t1 = w[8]; m0 = w[9]; m1 = w[10];
        main_pc = w[11]; cons_pc = w[12]; remaining = w[13]; waiting = w[14];
        now = w[15]; x_tick = w[16]; n_events = w[17]; max_queue = w[18]; status = w[19] & ~(uint32_t)CXXDES_B200_REP_UNFINISHED;
        hash = ((uint64_t)w[21] << 32) | w[20];
        total_latency = bits_as_double(((uint64_t)w[23] << 32) | w[22]);
        epoch = ((uint64_t)w[25] << 32) | w[24];
        set_until_rel(c);
        if (c.until64 < 0 || epoch == 0) until_rel = c.until;   // (no horizon, or not a wide-clock run)
        g = k & ~1u;
        glim = 0;
    }

    // Draw the ring entries of a restored replication again, up to the packet the producer puts
    // next (the invariant of the event loop: g > i while there are packets left).
    template <int MODE, bool SPILL, typename Ring, typename Table>
    CXB_HD void regenerate(const constants &c, Ring &ring, const Table *table) {
        if ((k & 1u) && wants_packets(c)) generate_pair<MODE, true>(c, ring, table, true);   // agk is A_k = A_{g+1}
        while (status == 0 && g <= i && wants_packets(c) && has_room<SPILL>(c)) generate_pair<MODE>(c, ring, table, true);
        glim = is_fast() ? g : 0u;
    }

    // The producer's event is next, its packet is not generated and the ring has no room: the
    // model's queue is longer than this ring.
    CXB_HD bool ring_exhausted(const constants &c) const {
        if (!wants_packets(c) || i != g) return false;   // (called when there is no room)
        if (is_fast()) return !(kc < kp);
        return m0 == M_PROD;
    }

    // ---- time of the next scheduled token -------------------------------------------------
    CXB_HD uint32_t next_time() const { return is_fast() ? (kc < kp ? kc : kp) >> 2 : t0; }

    template <int MODE>
    CXB_HD double seconds(const constants &c, uint32_t tick) const {
        // environment::now_seconds, environment.ipp:29-36: (double)(now * prec.t) * conv[prec.u];
        // tick * prec.t < 2^53, so the product of the two doubles is that integer exactly
        const double ticks = (MODE & MODE_WIDE) ? u64_to_double(epoch + tick) : u32_to_double(tick);
        if (MODE & MODE_TICK_MULT) return (ticks * c.tick_mult) * c.sec_per_unit;
        return ticks * c.sec_per_unit;
    }

    // ---- steady state: one event (two when a put wakes the consumer) ---------------------------
    // Which process a lane resumes differs from lane to lane, so this is one straight line of
    // predicated updates: a branch would make the warp run the producer's and the consumer's side
    // one after the other.  Both ring reads are unconditional (any index maps into the ring).
    // Lanes outside the steady state (or without a replication) fall through untouched.
    template <int MODE, typename Ring>
    CXB_HD void fast_slot(const constants &c, Ring &ring) {
#if defined(__CUDA_ARCH__)
        if constexpr (Ring::HAS_PTX_SLOT) {
            fast_slot_ptx<MODE, Ring::SPILL, Ring::MASK, Ring::STRIDE>(c, ring.base, ring.gbase, ring.gstride, ring.gmask);
            return;
        }
#endif
        const uint32_t key = kc < kp ? kc : kp;   // the front token
        // consumer resumes from its service timeout / producer resumes: q.put(now_seconds()),
        // queue.hpp:48-53, whose packet i must have been generated
        bool cons = kc < kp;
        bool put = !cons & (i < glim);
        if (MODE & MODE_UNTIL) {
            const bool due = key <= 4u * horizon<MODE>(c) + 3u;  // run_until stops at `until` (EMPTY wraps to all ones)
            cons &= due;
            put &= due;
        }
        const bool woke = put & (kc == EMPTY);    // ... and wake() finds the consumer parked
        const uint32_t t = key >> 2;
        const uint32_t i1 = i + 1;
        const uint32_t next_ak = ring.load_a(i1);
        const ring_entry e = ring.load(k, g);
        // environment::step: pop, fold (tokens are scheduled at now + d with d >= 0, so the clock
        // is simply the time of the last token; `now` itself is only needed outside this state)
        if (cons | put) hash = fold(hash, abs_time<MODE>(t), cons ? CONS_BITS : PROD_BITS);
        // The woken consumer's token {0 + now} was pushed before the producer's timeout, so it is
        // the very next event (event.hpp:125-134): it is dispatched in the same slot.
        if (woke) hash = fold(hash, abs_time<MODE>(t), CONS_BITS);
        // total_latency += now_seconds() - x   (producer_consumer.cpp:52); for the other lanes
        // x - x = +0 is added, which leaves the (non-negative) sum as it is
        const uint32_t ts = cons ? t : x_tick;
        total_latency += seconds<MODE>(c, ts) - seconds<MODE>(c, x_tick);
        if (put) {
            // packet i is already in the ring with A_i == now; env.timeout(exponential()) expires
            // at A_{i+1} = ring[i + 1].ak; a consumer asleep in its service keeps the front place
            // on a tie
            i = i1;
            kp = next_ak;
            kc &= ~2u;
        }
        const uint32_t count = i - k;             // the put is in, the pop below is not
        max_queue = count > max_queue ? count : max_queue;
        // consumer: q.pop() (queue.hpp:58-65) succeeds -> timeout(exponential()) = now + s_k, pushed
        // after the producer's token; or the consumer waits on queue.event_
        const bool attempt = cons | woke;
        const bool pop = attempt & (i != k);
        if (attempt) {
            ++n_events;                           // counts consumer events; puts are counted by i
            kc = EMPTY;
        }
        if (pop) {
            kc = 4u * t + e.s4;
            x_tick = e.ak >> 2;
            ++k;
        }
    }

#if defined(__CUDA_ARCH__)
    // The same slot, instruction for instruction, for a ring in shared memory (`base` = shared
    // address of this thread's column, entry j at base + (j & MASK) * STRIDE).  nvcc turns the C++
    // above into ~100 instructions per slot because its short-lived conditions do not fit the seven
    // predicate registers once four slots are interleaved; spelled out it is ~55.
    // Checked against the C++ version by the GPU parity tests (the host unit tests run the C++).
    template <int MODE, bool SPILL, int MASK, int STRIDE>
    __device__ __forceinline__ void fast_slot_ptx(const constants &c, uint32_t base, const void *gbase, uint32_t gstride,
                                                  uint32_t gmask) {
        const uint32_t ukey = (MODE & MODE_UNTIL) ? 4u * horizon<MODE>(c) + 3u : 0xFFFFFFFFu;
        const double mult = (MODE & MODE_TICK_MULT) ? c.tick_mult : 1.0;
        [[maybe_unused]] const uint32_t ep_lo = (uint32_t)epoch, ep_hi = (uint32_t)(epoch >> 32);
        // (the run_until check, the tick_mult multiply, the epoch and the global-memory level are left out when unused)
#define CXB_MM1_SLOT_TEXT(LOAD_K, OPERANDS, OPERANDS_WIDE)                                                         \
        if constexpr (MODE == 0) asm volatile(CXB_MM1_SLOT_A CXB_MM1_SLOT_B(LOAD_K) CXB_MM1_SLOT_C OPERANDS);              \
        else if constexpr (MODE == 1)                                                                              \
            asm volatile(CXB_MM1_SLOT_A CXB_MM1_SLOT_DUE CXB_MM1_SLOT_B(LOAD_K) CXB_MM1_SLOT_C OPERANDS);              \
        else if constexpr (MODE == 2)                                                                              \
            asm volatile(CXB_MM1_SLOT_A CXB_MM1_SLOT_B(LOAD_K) CXB_MM1_SLOT_MULT CXB_MM1_SLOT_C OPERANDS);             \
        else if constexpr (MODE == 3)                                                                              \
            asm volatile(CXB_MM1_SLOT_A CXB_MM1_SLOT_DUE CXB_MM1_SLOT_B(LOAD_K) CXB_MM1_SLOT_MULT CXB_MM1_SLOT_C OPERANDS); \
        else if constexpr (MODE == 4)                                                                              \
            asm volatile(CXB_MM1_SLOT_A CXB_MM1_SLOT_B_WIDE(LOAD_K) CXB_MM1_SLOT_C OPERANDS_WIDE);                    \
        else if constexpr (MODE == 5)                                                                              \
            asm volatile(CXB_MM1_SLOT_A CXB_MM1_SLOT_DUE CXB_MM1_SLOT_B_WIDE(LOAD_K) CXB_MM1_SLOT_C OPERANDS_WIDE);    \
        else if constexpr (MODE == 6)                                                                              \
            asm volatile(CXB_MM1_SLOT_A CXB_MM1_SLOT_B_WIDE(LOAD_K) CXB_MM1_SLOT_MULT CXB_MM1_SLOT_C OPERANDS_WIDE);   \
        else                                                                                                       \
            asm volatile(CXB_MM1_SLOT_A CXB_MM1_SLOT_DUE CXB_MM1_SLOT_B_WIDE(LOAD_K) CXB_MM1_SLOT_MULT CXB_MM1_SLOT_C  \
                             OPERANDS_WIDE);
        if constexpr (SPILL) {
            CXB_MM1_SLOT_TEXT(CXB_MM1_SLOT_LOAD_K_SPILL, CXB_MM1_SLOT_OPERANDS_SPILL, CXB_MM1_SLOT_OPERANDS_WIDE_SPILL)
        } else {
            CXB_MM1_SLOT_TEXT(CXB_MM1_SLOT_LOAD_K, CXB_MM1_SLOT_OPERANDS, CXB_MM1_SLOT_OPERANDS_WIDE)
        }
#undef CXB_MM1_SLOT_TEXT
    }
#endif

    // The steady state ends TAIL puts before the producer's last one (checked once per iteration,
    // so an iteration may run up to TAIL steady-state slots): rebuild the generic two-slot heap.
    static constexpr uint32_t TAIL = 8;
    CXB_HD void leave_fast_near_end(const constants &c) {
        if (is_fast() && i + 1 + TAIL >= c.n_packets) leave_fast();
    }
    CXB_HD void leave_fast() {
        const uint32_t tp = kp >> 2;
        const uint32_t tc = kc == EMPTY ? EMPTY : kc >> 2;
        const bool c_front = kc < kp;
        waiting = kc == EMPTY ? 1u : 0u;
        t0 = c_front ? tc : tp;
        m0 = c_front ? M_CONS : M_PROD;
        t1 = c_front ? tp : tc;
        m1 = c_front ? M_PROD : M_CONS;
        cons_pc = waiting ? 1u : 2u;
        n_events += i;
        kp = kc = EMPTY;
        glim = 0;
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
    template <int MODE, typename Ring>
    CXB_HD bool generic_step(const constants &c, Ring &ring) {
        if (is_fast() || t0 == EMPTY || t0 > horizon<MODE>(c)) return false;
        const uint32_t n = c.n_packets;
        if (m0 == M_PROD && i < n && i >= g) return false;  // A_{i+1} not generated yet
        const uint32_t t = t0;
        const uint32_t m = m0;
        t0 = t1; m0 = m1; t1 = EMPTY; m1 = M_NOOP;   // pop_heap on <= 2 entries
        now = t > now ? t : now;                     // environment.ipp:126
        ++n_events;
        hash = fold(hash, abs_time<MODE>(t), tag_hash_bits(m));
        if (m == M_PROD) {
            // producer body, producer_consumer.cpp:31-37
            if (i < n) {
                ++i;                                 // q.put(now_seconds()): ring entry i-1, A == now
                const uint32_t count = i - k;
                max_queue = count > max_queue ? count : max_queue;
                // wake(): the parked consumer's token goes in before the producer's own timeout
                if (waiting) { push(now, M_CONS); waiting = 0; }
                push(ring.load_a(i) >> 2, M_PROD);   // A_{i+1}
            } else {
                push(now, M_HANDLER);                // co_return -> completion carrying the all_of handler
            }
        } else if (m == M_CONS) {
            // consumer body, producer_consumer.cpp:39-54
            if (cons_pc == 2) total_latency += seconds<MODE>(c, now) - seconds<MODE>(c, x_tick);  // :52
            if (i != k) {
                const ring_entry e = ring.load(k, g);   // q.pop() succeeds, queue.hpp:58-65
                x_tick = e.ak >> 2;
                ++k;
                cons_pc = 2;
                if (k == n) push(now, M_HANDLER);    // consumer returns (:45-48)
                else push(now + (e.s4 >> 2), M_CONS);
            } else {
                waiting = 1;                         // ... or waits on queue.event_
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
        if (status == 0 && main_pc == 1 && remaining == 2 && i >= 1 && i + 1 + TAIL < n && t0 != EMPTY) {
            const bool prod_front = m0 == M_PROD && (t1 == EMPTY ? waiting != 0 : (m1 == M_CONS && cons_pc == 2));
            const bool cons_front = m0 == M_CONS && cons_pc == 2 && t1 != EMPTY && m1 == M_PROD;
            if (prod_front || cons_front) {
                const uint32_t tp = prod_front ? t0 : t1;
                const uint32_t tc = prod_front ? t1 : t0;
                kp = 4u * tp + 1u;
                kc = tc == EMPTY ? EMPTY : 4u * tc + (cons_front ? 0u : 2u);
                glim = g;
                n_events -= i;
            }
        }
        return true;
    }

    // environment::run / run_until (environment.ipp:179-196): nothing left, or nothing due.
    template <int MODE>
    CXB_HD bool finished(const constants &c) const {
        const uint32_t next = next_time();
        // (narrow modes compare with the launch constant, which is EMPTY when there is no horizon:
        // keeping until_rel out of them keeps a register out of the hot loop)
        if (!(MODE & MODE_WIDE)) return status != 0 || next == EMPTY || next > c.until;
        return status != 0 || next == EMPTY || ((MODE & MODE_UNTIL) && next > until_rel);
    }
};

}  // namespace mm1
}  // namespace cxb
