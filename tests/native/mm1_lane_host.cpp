// tests/native/mm1_lane_host.cpp -- TEST INFRASTRUCTURE, never part of the product library.
//
// Compiles the device state machine of the headline kernel (libcxxdes_b200/csrc/device/mm1_lane.cuh,
// a __host__ __device__ header) with g++ and drives one lane at a time through the same
// iteration the kernel runs (generate a pair, one generic step, EVENT_SLOTS steady-state slots),
// so the model logic can be checked against the oracle where there is no GPU.
#include <vector>

#include "device/mm1_lane.cuh"

using namespace cxb;

namespace {
// Same two levels as the device ring: a window of `slots` entries (index modulo slots) and, when
// `spill` is set, a larger second level that receives the entries leaving the window.
template <bool SPILL_>
struct host_ring {
    static constexpr bool HAS_PTX_SLOT = false;
    static constexpr bool SPILL = SPILL_;
    static constexpr int WINDOW = 31;
    std::vector<mm1::ring_entry> slots, global;
    host_ring(uint32_t n_global) : slots(WINDOW + 1), global(n_global ? n_global : 1) {}
    mm1::ring_entry load(uint32_t j, uint32_t g) const {
        if (SPILL && g - j > (uint32_t)WINDOW) return global[j & (global.size() - 1)];
        return slots[j & WINDOW];
    }
    uint32_t load_a(uint32_t j) const { return slots[j & WINDOW].ak; }
    void store_if(bool p, uint32_t j, const mm1::ring_entry &e) { if (p) slots[j & WINDOW] = e; }
    void store_a_if(bool p, uint32_t j, uint32_t ak) { if (p) slots[j & WINDOW].ak = ak; }
    void store_level(uint32_t j, uint32_t g, const mm1::ring_entry &e) {
        if (SPILL && g - j > (uint32_t)WINDOW) global[j & (global.size() - 1)] = e;
        else slots[j & WINDOW] = e;
    }
    void spill_if(bool p, uint32_t g) {
        if (!p) return;
        global[(g - WINDOW) & (global.size() - 1)] = slots[(g + 1) & WINDOW];
        global[(g - WINDOW + 1) & (global.size() - 1)] = slots[(g + 2) & WINDOW];
    }
};

template <int MODE, typename Ring>
void run_mode(const mm1::constants &c, Ring &ring, mm1::lane &L, int event_slots, uint64_t &iters) {
    for (;;) {
        ++iters;
        // same order as one iteration of mm1_ahead_kernel
        if (!L.is_fast()) L.generic_step<MODE>(c, ring);
        L.leave_fast_near_end(c);
        for (int s = 0; s < event_slots; ++s) L.fast_slot<MODE>(c, ring);
        const bool want = L.status == 0 && L.wants_packets(c);
        const bool room = L.template has_room<Ring::SPILL>(c);
        L.generate_pair<MODE>(c, ring, LOG_TABLE_HOST, want && room);
        if (want && !room && L.ring_exhausted(c)) L.status |= CXXDES_B200_REP_QUEUE_OVERFLOW;
        if (L.finished<MODE>(c)) break;
    }
}

// restored = the lane comes from a saved record: draw its ring entries again first
template <typename Ring>
void run_one(const mm1::constants &c, Ring &ring, mm1::lane &L, int event_slots, int mode, bool restored, uint64_t &iters) {
    switch (mode) {
#define CXB_CASE(M)                                                                       \
    case M:                                                                               \
        if (restored) L.regenerate<M, Ring::SPILL>(c, ring, LOG_TABLE_HOST);              \
        run_mode<M>(c, ring, L, event_slots, iters);                                      \
        break;
        CXB_CASE(0) CXB_CASE(1) CXB_CASE(2) CXB_CASE(3) CXB_CASE(4) CXB_CASE(5) CXB_CASE(6)
        default:
            if (restored) L.regenerate<7, Ring::SPILL>(c, ring, LOG_TABLE_HOST);
            run_mode<7>(c, ring, L, event_slots, iters);
            break;
#undef CXB_CASE
    }
}
}  // namespace

extern "C" int mm1_lane_host_run(const cxxdes_b200_mm1_params *prm, int64_t ticks_per_unit, int64_t tick_mult,
                                 double sec_per_unit, uint64_t seed, uint64_t first, uint64_t n_reps,
                                 int64_t until, uint32_t ring_cap, uint32_t global_slots, int event_slots,
                                 int general_mode, int64_t piece, int wide, const cxxdes_b200_columns *out, uint64_t *iterations) {
    mm1::constants c;
    c.lambda = make_divisor(prm->lambda);
    c.mu = make_divisor(prm->mu);
    c.keys = philox_schedule_of((uint32_t)seed, (uint32_t)(seed >> 32));
    c.ticks_per_unit = (double)ticks_per_unit;
    c.sec_per_unit = sec_per_unit;
    c.tick_mult = (double)tick_mult;
    c.n_packets = (uint32_t)prm->n_packets;
    // window only: ring_cap <= 31 packets; spilling ring: ring_cap = global_slots, lead limited to the window
    c.ring_cap = ring_cap;
    c.spill_cap = global_slots ? global_slots : ring_cap;
    c.lead_cap = global_slots ? 30 : ring_cap;
    c.until = (until < 0 || until >= (int64_t)mm1::TIME_LIMIT) ? mm1::EMPTY : (uint32_t)until;
    uint64_t iters = 0;
    // MODE 3 (generic step) = run_until check + tick_mult multiply compiled in: same results as the
    // specialised modes (a multiply by 1.0 and a compare with EMPTY change nothing)
    const int mode = general_mode ? 3 : (c.until != mm1::EMPTY ? 1 : 0) | (tick_mult != 1 ? 2 : 0);
    // `piece` > 0: run_for(piece) again and again up to `until` (or to exhaustion), saving the lane to
    // a 128-byte record at every horizon and restoring + regenerating it, as the kernel does between
    // two launches.  The ring object is new for every piece: nothing may survive in it.
    auto horizon_of = [](int64_t h) { return (h < 0 || h >= (int64_t)mm1::TIME_LIMIT) ? mm1::EMPTY : (uint32_t)h; };
    for (uint64_t r = 0; r < n_reps; ++r) {
        mm1::lane L;
        L.init(first + r, c);
        uint32_t record[mm1::lane::STATE_WORDS] = {0};
        int64_t h = piece > 0 ? piece : until;
        for (;;) {
            if (piece > 0 && until >= 0 && h > until) h = until;
            c.until = horizon_of(h);
            c.until64 = h;
            // general_mode: run_until check + tick_mult multiply compiled in; same results as the specialised
            // modes (a multiply by 1.0 and a compare with EMPTY change nothing)
            const bool has_until = wide ? h >= 0 : c.until != mm1::EMPTY;
            const int mode_h = (general_mode ? 3 : (has_until ? 1 : 0) | (tick_mult != 1 ? 2 : 0)) | (wide ? 4 : 0);
            const bool restored = record[0] == mm1::lane::PHASE_RUNNING;
            if (restored) L.restore(record, first + r, c);
            else if (h == (piece > 0 ? piece : until)) L.init(first + r, c);   // (first piece: until_rel from this horizon)
            bool unfinished;
            if (global_slots) {
                host_ring<true> ring(global_slots);
                run_one(c, ring, L, event_slots, mode_h, restored, iters);
                unfinished = L.status == 0 && L.next_time() != mm1::EMPTY;
                L.save(record, ring, unfinished ? mm1::lane::PHASE_RUNNING : mm1::lane::PHASE_DONE);
            } else {
                host_ring<false> ring(0);
                run_one(c, ring, L, event_slots, mode_h, restored, iters);
                unfinished = L.status == 0 && L.next_time() != mm1::EMPTY;
                L.save(record, ring, unfinished ? mm1::lane::PHASE_RUNNING : mm1::lane::PHASE_DONE);
            }
            if (piece <= 0 || !unfinished || (until >= 0 && h >= until)) break;
            h += piece;
        }
        uint32_t status = L.status;
        int64_t end = (int64_t)(L.epoch + L.now);
        if (until >= 0) {
            if (L.next_time() != mm1::EMPTY) status |= CXXDES_B200_REP_UNFINISHED;
            end = end > until ? end : until;
        }
        out->end_time[r] = end;
        out->n_events[r] = L.events();
        out->trace_hash[r] = L.hash;
        out->status[r] = status;
        out->f64[0][r] = L.k == c.n_packets ? L.total_latency / (double)prm->n_packets : 0.0;
        out->f64[1][r] = L.total_latency;
        out->i64[0][r] = (int64_t)L.k;
        out->i64[1][r] = (int64_t)L.max_queue;
    }
    if (iterations) *iterations = iters;
    return 0;
}
