// mm1_ahead.cu -- headline kernel: M/M/1 (examples/producer_consumer.cpp:9-59), one thread per
// replication, variates generated ahead of the event loop (device/mm1_lane.cuh has the model).
//
// Per loop iteration every lane of the warp, together:
//   1. generates two packets (one Philox block per stream, four log/divide/convert chains with
//      no dependence between them) into its ring in shared memory -- convergent, no lane idles;
//   2. runs SLOTS bookkeeping slots; a slot dispatches the lane's next event (a put that wakes
//      the consumer takes the wake-up token along), so a packet costs two slots whatever the
//      order of events.  A lane that has caught up with its generated packets skips a slot, a lane
//      whose ring is full skips a generation; the ring between the two absorbs the difference.
// Start-up and wind-down events (co_main, all_of handler, completions) go through a generic
// two-slot heap, one per iteration.  A lane whose replication ends takes the next replication
// from the global counter at once (replication-level work stealing).
//
// Replications the 32-bit/shared-memory representation cannot hold (queue longer than the ring,
// ticks beyond 2^30) are appended to the retry list and redone by mm1_fused_kernel<false>.
#include <cmath>
#include <cstdlib>

#include "device/mm1_lane.cuh"
#include "device/models_common.cuh"
#include "kernels.h"

// Generation + event-slot rounds per pass of the outer loop (work stealing, start-up/wind-down and
// end-of-replication checks run once per pass).
// (2 since round 2: 101.90 against 102.38 ms for the headline run -- the 80-odd instructions of a pass of the outer
// loop are mostly branches and uniform-path work that the pipes the block is bound by do not see; 1 measured the same
// within noise in round 1 on an earlier version of the block)
#ifndef CXB_MM1_UNROLL
#define CXB_MM1_UNROLL 2
#endif
// Source order of the event slots and the generation inside the branch-free block (measured: see DESIGN.md)
#ifndef CXB_MM1_ORDER
#define CXB_MM1_ORDER 0
#endif

namespace cxb {

namespace {

// Ring of one replication.  First level: SLOTS 8-byte entries per thread in shared memory,
// THREADS * 8 bytes between slots, so a warp's access touches each bank pair once whatever slot
// each lane is at.  Explicit ld/st.shared on a 32-bit address: no generic-to-shared conversion per
// access.  Second level (SPILL): entries that leave the window while still queued go to a ring in
// global memory, laid out [slot][thread] (lanes generate in lock step, so the spill stores of a
// warp are contiguous); only heavily loaded queues ever touch it.
// All ring accesses are volatile asm, which keeps their relative order; none has a "memory"
// clobber, so the arithmetic around them (and the read-only log table loads) can be scheduled
// freely across them.
template <int SLOTS, int THREADS, bool SPILL_>
struct lane_ring {
    static_assert((SLOTS & (SLOTS - 1)) == 0, "ring slots must be a power of two");
    static constexpr bool HAS_PTX_SLOT = true;
    static constexpr bool SPILL = SPILL_;
    static constexpr int MASK = SLOTS - 1;
    static constexpr int WINDOW = SLOTS - 1;
    static constexpr int STRIDE = THREADS * 8;
    uint32_t base;       // shared-window address of this thread's column
    const void *gbase;   // global ring: this thread's column (entry j at gbase + (j & gmask) * gstride)
    uint32_t gstride;    // bytes between global slots = total threads * 8
    uint32_t gmask;

    __device__ __forceinline__ uint32_t addr(uint32_t j) const { return base + (j & MASK) * STRIDE; }
    __device__ __forceinline__ const char *gaddr(uint32_t j) const {
        return (const char *)gbase + (uint64_t)(j & gmask) * gstride;
    }
    __device__ __forceinline__ mm1::ring_entry load_shared(uint32_t j) const {
        mm1::ring_entry e;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(e.ak), "=r"(e.s4) : "r"(addr(j)));
        return e;
    }
    // entry j of a lane that has generated g packets
    __device__ __forceinline__ mm1::ring_entry load(uint32_t j, uint32_t g) const {
        if (SPILL && g - j > (uint32_t)WINDOW) {
            mm1::ring_entry e;
            asm volatile("ld.global.v2.u32 {%0, %1}, [%2];" : "=r"(e.ak), "=r"(e.s4) : "l"(gaddr(j)));
            return e;
        }
        return load_shared(j);
    }
    // rewrite entry j at the level load(j, g) reads it from (MODE_WIDE rebase)
    __device__ __forceinline__ void store_level(uint32_t j, uint32_t g, const mm1::ring_entry &e) const {
        if (SPILL && g - j > (uint32_t)WINDOW)
            asm volatile("st.global.v2.u32 [%0], {%1, %2};" ::"l"(gaddr(j)), "r"(e.ak), "r"(e.s4));
        else
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr(j)), "r"(e.ak), "r"(e.s4));
    }
    __device__ __forceinline__ uint32_t load_a(uint32_t j) const {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr(j)));
        return v;
    }
    __device__ __forceinline__ void store_if(bool p, uint32_t j, const mm1::ring_entry &e) const {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p st.shared.v2.u32 [%0], {%1, %2};\n\t}"
                     ::"r"(addr(j)), "r"(e.ak), "r"(e.s4), "r"((uint32_t)p));
    }
    __device__ __forceinline__ void store_a_if(bool p, uint32_t j, uint32_t ak) const {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.u32 [%0], %1;\n\t}"
                     ::"r"(addr(j)), "r"(ak), "r"((uint32_t)p));
    }
    // entries g - WINDOW and g - WINDOW + 1 (window slots g + 1 and g + 2) move to the global level
    __device__ __forceinline__ void spill_if(bool p, uint32_t g) const {
        const mm1::ring_entry e1 = load_shared(g + 1), e2 = load_shared(g + 2);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %6, 0;\n\t"
                     "@p st.global.v2.u32 [%0], {%1, %2};\n\t@p st.global.v2.u32 [%3], {%4, %5};\n\t}"
                     ::"l"(gaddr(g - WINDOW)), "r"(e1.ak), "r"(e1.s4), "l"(gaddr(g - WINDOW + 1)), "r"(e2.ak), "r"(e2.s4),
                       "r"((uint32_t)p));
    }
};

}  // namespace

// One round of the spilling path (event slots + generation) as a function of its own, for the
// adaptive kernel.
// SPILL: the ring has its second level in global memory (heavily loaded queues, and the second pass).
// Passes of one run (launch_mm1_fused): this kernel over the whole shard; this kernel with SPILL
// over the replications whose queue outgrew the window (`list`: about one in two million at
// rho = 0.5, but a lone replication on the 64-bit fallback kernel takes 13 ms, on this one 2 ms);
// the fallback kernel over what is left (clock beyond 2^30 ticks, queue beyond both levels).
// A replication this kernel cannot finish goes to `fail_list`.
// POLL (the helper of the first pass, launch_mm1_ahead_helper): the kernel runs WHILE the first pass is still
// appending to `list`.  A lane reserves the next list index and waits for that entry to appear (entries start as
// LIST_UNSET) or for `*done_flag` to become non-zero -- set by a memset that follows the first pass in its stream --
// after which an entry that is still unset will never be written.  So the one-in-two-million replication whose
// queue outgrows the window is redone next to the first pass instead of after it.  (done_flag[2] counts what the
// helper took: done_flag is the low word of counters[5], done_flag + 2 the low word of counters[6].)
template <int THREADS, int MIN_BLOCKS, int RING_SLOTS, int EVENT_SLOTS, int MODE, bool SPILL, bool STATE, bool POLL = false>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
mm1_ahead_kernel(mm1::constants c, cxxdes_b200_mm1_params prm, shard sh, device_columns out, uint2 *gring,
                 uint32_t gring_slots, uint32_t *state, int resume, const uint32_t *list,
                 const unsigned long long *list_count, unsigned long long *list_cursor, uint32_t *fail_list,
                 unsigned long long *fail_count, unsigned int *done_flag = nullptr, uint32_t list_capacity = 0) {
    extern __shared__ __align__(16) unsigned char smem[];
    log_entry *table = (log_entry *)smem;
    stage_log_table(table, sh.log_table);
    lane_ring<RING_SLOTS, THREADS, SPILL> ring;
    ring.gbase = gring + ((size_t)blockIdx.x * THREADS + threadIdx.x);
    ring.gstride = gridDim.x * THREADS * 8u;
    ring.gmask = gring_slots - 1;
    ring.base = (uint32_t)__cvta_generic_to_shared(smem + sizeof(log_entry) * (1 << CXB_LOG_TABLE_BITS)) + threadIdx.x * 8u;

    mm1::lane L;
    L.init(0, c);
    L.idle();
    bool active = false, exhausted = false;
    uint64_t rep = 0;
    [[maybe_unused]] unsigned long long poll_idx = ~0ULL;   // POLL: the list index this lane has reserved
    [[maybe_unused]] unsigned long long poll_t0 = 0;        // POLL: when it took its replication (diagnostics)
    // POLL: one replication per warp.  Only lane 0 reserves list entries and polls; with all 32 lanes polling, the 31
    // idle ones looked at the flag and at their list entry (two dependent L2 round trips) on EVERY turn of the loop
    // their busy neighbour was running its replication in, and made it 2-3 times slower (a replication redone by the
    // helper cost the step 2-3 ms wherever in the run it had failed: 105.5 ms on the ranks that had one).
    if constexpr (POLL) exhausted = (threadIdx.x & 31u) != 0;

    for (;;) {
        if (!active && !exhausted) {
            // `state` (run_until / run_for in pieces): one 128-byte record per replication, written
            // when a replication stops at the horizon and read back by the next launch (`resume`)
            for (int tries = 0; tries < 8 && !active && !exhausted; ++tries) {
                bool got;
                if constexpr (POLL) {
                    const volatile uint32_t *vlist = list;
                    const bool finished = *(volatile unsigned int *)done_flag != 0;   // read BEFORE the entry
                    if (finished) __threadfence();
                    if (poll_idx == ~0ULL) {
                        if (finished) {   // the rest of the list belongs to the pass that follows
                            exhausted = true;
                            break;
                        }
                        poll_idx = atomicAdd(list_cursor, 1ULL);
                    }
                    const uint32_t v = poll_idx < list_capacity ? vlist[poll_idx] : LIST_UNSET;
                    got = v != LIST_UNSET;
                    if (got) {
                        rep = v;
                        poll_idx = ~0ULL;
                        atomicAdd(done_flag + 2, 1u);   // diagnostics (cxxdes_b200_pass_counts): taken by the helper
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(poll_t0));
                    } else if (finished) {
                        exhausted = true;   // the first pass ended without writing this entry
                        break;
                    } else {
                        break;              // not yet: look again on the next turn of the loop
                    }
                } else if (list) {
                    const unsigned long long idx = atomicAdd(list_cursor, 1ULL);
                    got = idx < *list_count;
                    if (got) rep = list[idx];
                } else {
                    got = steal(sh, rep);
                }
                if (!got) {
                    exhausted = true;
                    break;
                }
                const uint64_t grep = sh.first_replication + rep;
                const uint32_t *w = state + rep * mm1::lane::STATE_WORDS;
                // (a second-pass replication starts over: phase REPLAY in its record, or no record at all)
                const uint32_t phase = (STATE && resume && !list) ? w[0] : (uint32_t)mm1::lane::PHASE_NEW;
                if (phase == mm1::lane::PHASE_RUNNING) {
                    L.restore(w, grep, c);
                    L.template regenerate<MODE, SPILL>(c, ring, table);
                    active = true;
                } else if (phase == mm1::lane::PHASE_NEW) {
                    L.init(grep, c);
                    active = true;
                } else if (phase == mm1::lane::PHASE_REPLAY) {
                    // left the 32-bit representation in an earlier piece: the fallback kernel
                    // replays it from tick 0 to the new horizon
                    const unsigned long long at = atomicAdd(fail_count, 1ULL);
                    fail_list[at] = (uint32_t)rep;
                } else if (sh.run_until >= 0 && out.end_time[rep] < sh.run_until) {
                    // PHASE_DONE: finished in an earlier piece, its row is final -- except the clock:
                    // run_until sets now_ = max(now_, t) even when nothing is left (environment.ipp:194)
                    out.end_time[rep] = sh.run_until;
                }
            }
        }
        // (a claimed replication can leave the lane inactive -- finished in an earlier piece, handed to the
        // retry list, nothing due in this piece -- so the warp only leaves when every lane has seen the end of the work list)
        if constexpr (STATE || POLL) {
            if (!__any_sync(0xFFFFFFFFu, active)) {
                if (__all_sync(0xFFFFFFFFu, exhausted)) break;
                if constexpr (POLL) __nanosleep(4000);   // nothing to redo yet: leave the issue slots to the first pass
                continue;
            }
        } else {
            // (without saved state every claimed replication makes its lane active: no lane is idle before the list ends)
            if (!__any_sync(0xFFFFFFFFu, active)) break;
        }

        // ---- 1. start-up / wind-down events (rare) ----
        if (active && !L.is_fast()) L.template generic_step<MODE>(c, ring);
        static_assert(EVENT_SLOTS * CXB_MM1_UNROLL <= (int)mm1::lane::TAIL,
                      "the end-of-steady-state check runs once per iteration");
        L.leave_fast_near_end(c);
        __syncwarp();

        // ---- 2. one branch-free block: steady-state events on the packets generated so far,
        //         and the next two packets for every lane that has room for them ----
        bool want, room;
#pragma unroll
        for (int u = 0; u < CXB_MM1_UNROLL; ++u) {
#if CXB_MM1_ORDER == 1
            want = active && L.status == 0 && L.wants_packets(c);
            room = L.template has_room<SPILL>(c);
            L.template generate_pair<MODE>(c, ring, table, want & room);
#pragma unroll
            for (int s = 0; s < EVENT_SLOTS; ++s) L.template fast_slot<MODE>(c, ring);
#elif CXB_MM1_ORDER == 2
#pragma unroll
            for (int s = 0; s < EVENT_SLOTS / 2; ++s) L.template fast_slot<MODE>(c, ring);
            want = active && L.status == 0 && L.wants_packets(c);
            room = L.template has_room<SPILL>(c);
            L.template generate_pair<MODE>(c, ring, table, want & room);
#pragma unroll
            for (int s = EVENT_SLOTS / 2; s < EVENT_SLOTS; ++s) L.template fast_slot<MODE>(c, ring);
#else
#pragma unroll
            for (int s = 0; s < EVENT_SLOTS; ++s) L.template fast_slot<MODE>(c, ring);  // no-op for idle lanes
            want = active && L.status == 0 && L.wants_packets(c);
            room = L.template has_room<SPILL>(c);
            L.template generate_pair<MODE>(c, ring, table, want & room);
#endif
        }
        // (a ring that is truly exhausted stays so, checking after the last generation is enough)
        if (want && !room && L.ring_exhausted(c)) L.status |= CXXDES_B200_REP_QUEUE_OVERFLOW;

        // ---- 4. end of this lane's replication? ----
        if (active && L.template finished<MODE>(c)) {
            active = false;
            if constexpr (POLL) {   // diagnostics: the longest time the helper spent on one replication, in ns
                unsigned long long t1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                atomicMax((unsigned long long *)(done_flag + 4), t1 - poll_t0);
            }
            const uint32_t status = L.status;
            uint32_t *w = state + rep * mm1::lane::STATE_WORDS;
            if (status & (CXXDES_B200_REP_QUEUE_OVERFLOW | CXXDES_B200_REP_TIME_RANGE)) {
                // hand the replication to the next pass; its row is written there
                unsigned long long at = atomicAdd(fail_count, 1ULL);
                fail_list[at] = (uint32_t)rep;
                if (STATE) w[0] = mm1::lane::PHASE_REPLAY;
            } else {
                rep_out o;
                int64_t end = (MODE & mm1::MODE_WIDE) ? (int64_t)(L.epoch + L.now) : (int64_t)L.now;
                o.status = status;
                if (sh.run_until >= 0) {
                    if (L.next_time() != mm1::EMPTY) o.status |= CXXDES_B200_REP_UNFINISHED;
                    end = end > sh.run_until ? end : sh.run_until;  // now_ = max(now_, t), environment.ipp:194
                }
                o.end_time = end;
                o.n_events = L.events();
                o.trace_hash = L.hash;
                o.f64[0] = L.k == c.n_packets ? L.total_latency / (double)prm.n_packets : 0.0;  // :46
                o.f64[1] = L.total_latency;
                o.i64[0] = (int64_t)L.k;
                o.i64[1] = (int64_t)L.max_queue;
                store_result(out, rep, o);
                if (STATE) L.save(w, ring, L.next_time() != mm1::EMPTY ? mm1::lane::PHASE_RUNNING : mm1::lane::PHASE_DONE);
            }
            L.idle();
        }
    }
}

namespace {
constexpr int AH_EVENT_SLOTS = 4;
// Helper of the first pass: four one-warp blocks (one replication each).  They do NOT fit next to three 256-thread
// blocks of the first pass: an SM's 64 K registers are four banks of 16 K, one per scheduler, six first-pass warps of
// 80 registers leave 1024 in each, and the helper's warp needs 3584 (measured: launched next to a full first pass, the
// helper became resident 2.5 ms before the first pass ended -- when its first blocks retired -- and a replication it redid
// cost the step 2.3 ms wherever in the run it had failed).  So a first pass that has a helper runs ONE BLOCK SHORT of a
// full grid (launch_mm1_fused): the SM with two blocks has 5120 registers free per scheduler, one helper warp each;
// 0.23 % fewer lanes for the first pass against 2.3-4.5 ms whenever some replication outgrows the window (about one
// shard in two at the headline size, some rank nearly always on eight GPUs).
constexpr int AH_HELPER_THREADS = 32, AH_HELPER_MIN_BLOCKS = 16, AH_HELPER_BLOCKS = 4;
constexpr size_t TABLE_BYTES = sizeof(log_entry) * (1 << CXB_LOG_TABLE_BITS);

// Launch geometry: threads per block, resident blocks per SM, window slots per replication.
// (Measured alternatives: 128-thread blocks x 6 per SM: the same time; shallower windows with more
// resident warps, 16 slots x 32 warps/SM: 113.8 ms against 109.2 ms at lambda = 1.)
template <int T, int B, int S>
struct geometry {
    static constexpr int THREADS = T, MIN_BLOCKS = B, RING_SLOTS = S;
    static constexpr size_t SMEM = TABLE_BYTES + (size_t)S * T * sizeof(uint2);
};
using geometry0 = geometry<256, 3, 32>;
// Other geometries were measured for shards that are not a whole number of rounds of these 113 664 lanes (a 1 Mi
// ensemble split over 8 GPUs is 131 072 replications per GPU = 1.15 rounds, 15.0 ms instead of the 12.8 ms of an
// eighth of the full run) and none was kept (profiles/r02_mm1_geometry_ab.jsonl): more lanes with a 16-slot window
// and the two-level ring from the start, <128,7,16> 15.5 ms and <256,4,16> 18.2 ms (the spilling slot text and
// 64-72 registers cost more than the partial round); half as many lanes for two whole rounds, <224,2,32> 16.3 ms
// and <448,1,32> 15.8 ms (one round of 14 resident warps per SM takes 7.8 ms against 11.5 ms for 24: 85 % of the
// throughput, the kernel is not purely issue-bound down there).  A replication is an indivisible 11 ms unit of work.

template <typename G>
void fill_plan(mm1_fused_plan &p, int sm_count) {
    p.ah_threads = G::THREADS;
    p.ah_blocks = sm_count * G::MIN_BLOCKS;
    p.ah_smem_bytes = G::SMEM;
    p.ah_ring_slots = G::RING_SLOTS;
}
}  // namespace

void plan_mm1_ahead(mm1_fused_plan &p, const cxxdes_b200_mm1_params &prm, uint32_t queue_capacity, int sm_count,
                    int64_t tick_mult, int64_t ticks_per_unit, uint64_t n_replications) {
    (void)n_replications;
    p.ah_geometry = 0;
    fill_plan<geometry0>(p, sm_count);
    const uint32_t window = p.ah_ring_slots - 1;  // one slot holds the arrival key after the last packet
    // Stationary P(queue length > Q) = rho^(Q+1) per arrival; the window also holds the packets
    // generated ahead (2 to 4).  When more than ~0.1% of the replications are expected to outgrow
    // the window, the first pass already runs with the spilling ring (ah_spill); otherwise it runs
    // on the window alone and the few replications that outgrow it are redone by a second pass
    // with the spilling ring.  The global-memory level is sized so that overflowing it has
    // probability < 1e-12 per arrival (what still overflows is redone by the fallback kernel).
    const double rho = prm.lambda / prm.mu;
    const double q = (double)window - 3.0;
    const double p_over = rho >= 1.0 ? 1.0 : pow(rho, q) * (double)prm.n_packets;
    p.ah_spill = !queue_capacity && !(p_over < 1e-3);
    uint32_t cap = queue_capacity ? queue_capacity : window;
    if (cap > window) cap = window;
    if (cap < 2) cap = 2;
    p.ah_ring_cap = cap;
    {
        double need = rho < 1.0 ? ceil(log(1e-12 / (double)prm.n_packets) / log(rho)) + 64.0 : (double)prm.n_packets + 64.0;
        if (need > (double)prm.n_packets + 64.0) need = (double)prm.n_packets + 64.0;
        const uint64_t threads = (uint64_t)p.ah_blocks * p.ah_threads;
        uint64_t budget = (8ULL << 30) / (threads * sizeof(uint2));  // 8 GiB of the 180
        uint32_t slots = 64;
        while (slots < need && slots * 2ULL <= budget) slots *= 2;
        p.ah_gring_slots = slots;
        p.ah_spill_cap = slots;
        p.ah_lead_cap = window - 1;
    }
    // tick_mult < 2^23 keeps tick * tick_mult exact in a double (mm1_lane.cuh seconds()),
    // every single delay must fit the 32-bit tick representation,
    // and so must the clock at the end of the run: the lane's bound is A_n + sum of the service
    // times, n (1/lambda + 1/mu) ticks_per_unit on average.  A run that is going to pass 2^30 ticks
    // anyway goes to the 64-bit kernel at once (1 us precision with a 1 s unit: 272 ms instead of
    // 890 ms spent failing late in both fast passes first).
    const double expected_bound = (double)prm.n_packets * (1.0 / prm.lambda + 1.0 / prm.mu) * (double)ticks_per_unit;
    const bool draws_fit = mm1::draw_fits_u32(prm.lambda, (double)ticks_per_unit) && mm1::draw_fits_u32(prm.mu, (double)ticks_per_unit);
    // A clock that passes 2^30 ticks (fine time precision) is no obstacle as long as the LIVE tokens
    // span far less: MODE_WIDE keeps 32-bit times relative to a per-lane epoch.  The live span is a
    // window of ~64 inter-arrival / service times; 16x the mean must stay below 2^29.
    p.ah_wide = !(1.25 * expected_bound < (double)mm1::TIME_LIMIT);
    const double span = 64.0 * 16.0 * (1.0 / prm.lambda + 1.0 / prm.mu) * (double)ticks_per_unit;
    p.ah_direct_global = tick_mult >= (1 << 23) || !draws_fit || (p.ah_wide && !(span < (double)mm1::REBASE_LIMIT)) ||
                         expected_bound * (double)tick_mult >= 1e15;
}

size_t mm1_ahead_state_bytes(uint64_t n_replications) { return (size_t)n_replications * mm1::lane::STATE_WORDS * sizeof(uint32_t); }

size_t mm1_ahead_gring_bytes(const mm1_fused_plan &p) {
    return (size_t)p.ah_gring_slots * p.ah_blocks * p.ah_threads * sizeof(uint2);
}

namespace {
struct pass_args {
    uint2 *gring;
    uint32_t *state;
    int resume;
    const uint32_t *list;
    const unsigned long long *list_count;
    unsigned long long *list_cursor;
    uint32_t *fail_list;
    unsigned long long *fail_count;
    int leave_blocks;   // launch this many blocks short of the full grid (room for the helper, see AH_HELPER_BLOCKS)
};

template <typename G, int MODE, bool SPILL>
cudaError_t configure_one() {
    // (the same shared-memory carve-out preference -- all of it -- for every instantiation: the helper's blocks are to
    // become resident on SMs the first pass has configured for its three 66 KB blocks)
    auto plain = mm1_ahead_kernel<G::THREADS, G::MIN_BLOCKS, G::RING_SLOTS, AH_EVENT_SLOTS, MODE, SPILL, false>;
    auto keeps = mm1_ahead_kernel<G::THREADS, G::MIN_BLOCKS, G::RING_SLOTS, AH_EVENT_SLOTS, MODE, SPILL, true>;
    cudaError_t e = cudaFuncSetAttribute(plain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(keeps, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(plain, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(keeps, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    return e;
}
template <typename G>
cudaError_t configure_geometry() {
    cudaError_t e = cudaSuccess;
#define CXB_CONFIGURE(M)                                              \
    if (e == cudaSuccess) e = configure_one<G, M, false>();           \
    if (e == cudaSuccess) e = configure_one<G, M, true>();
    CXB_CONFIGURE(0) CXB_CONFIGURE(1) CXB_CONFIGURE(2) CXB_CONFIGURE(3)
    CXB_CONFIGURE(4) CXB_CONFIGURE(5) CXB_CONFIGURE(6) CXB_CONFIGURE(7)
#undef CXB_CONFIGURE
    return e;
}
using helper_geometry = geometry<AH_HELPER_THREADS, AH_HELPER_MIN_BLOCKS, 32>;
template <int MODE>
void launch_helper_one(const mm1::constants &c, const cxxdes_b200_mm1_params &prm, const shard &sh, const device_columns &out,
                       const mm1_fused_plan &plan, const pass_args &a, unsigned int *done_flag, uint32_t list_capacity,
                       cudaStream_t stream) {
    using G = helper_geometry;
    static const int blocks = [] {   // CXXDES_B200_MM1_HELPER_BLOCKS: measurements
        const char *v = getenv("CXXDES_B200_MM1_HELPER_BLOCKS");
        const int b = v ? atoi(v) : 0;
        return b >= 1 && b <= 1024 ? b : AH_HELPER_BLOCKS;
    }();
    mm1_ahead_kernel<G::THREADS, G::MIN_BLOCKS, G::RING_SLOTS, AH_EVENT_SLOTS, MODE, true, false, true>
        <<<blocks, G::THREADS, G::SMEM, stream>>>(c, prm, sh, out, a.gring, plan.ah_gring_slots, nullptr, 0, a.list,
                                                            a.list_count, a.list_cursor, a.fail_list, a.fail_count, done_flag,
                                                            list_capacity);
}
template <typename G, int MODE, bool SPILL>
void launch_one(const mm1::constants &c, const cxxdes_b200_mm1_params &prm, const shard &sh, const device_columns &out,
                const mm1_fused_plan &plan, const pass_args &a, cudaStream_t stream) {
    // (the state-keeping variant is a separate instantiation: its extra live values cost the plain
    // run() path 4 % when they are only a run-time switch)
    if (a.state)
        mm1_ahead_kernel<G::THREADS, G::MIN_BLOCKS, G::RING_SLOTS, AH_EVENT_SLOTS, MODE, SPILL, true>
            <<<plan.ah_blocks - a.leave_blocks, G::THREADS, G::SMEM, stream>>>(c, prm, sh, out, a.gring, plan.ah_gring_slots, a.state, a.resume,
                                                              a.list, a.list_count, a.list_cursor, a.fail_list, a.fail_count);
    else
        mm1_ahead_kernel<G::THREADS, G::MIN_BLOCKS, G::RING_SLOTS, AH_EVENT_SLOTS, MODE, SPILL, false>
            <<<plan.ah_blocks - a.leave_blocks, G::THREADS, G::SMEM, stream>>>(c, prm, sh, out, a.gring, plan.ah_gring_slots, nullptr, 0, a.list,
                                                              a.list_count, a.list_cursor, a.fail_list, a.fail_count);
}
template <typename G, bool SPILL>
void launch_geometry(int mode, const mm1::constants &c, const cxxdes_b200_mm1_params &prm, const shard &sh,
                     const device_columns &out, const mm1_fused_plan &plan, const pass_args &a, cudaStream_t stream) {
    switch (mode) {
        case 0: launch_one<G, 0, SPILL>(c, prm, sh, out, plan, a, stream); break;
        case 1: launch_one<G, 1, SPILL>(c, prm, sh, out, plan, a, stream); break;
        case 2: launch_one<G, 2, SPILL>(c, prm, sh, out, plan, a, stream); break;
        case 3: launch_one<G, 3, SPILL>(c, prm, sh, out, plan, a, stream); break;
        case 4: launch_one<G, 4, SPILL>(c, prm, sh, out, plan, a, stream); break;
        case 5: launch_one<G, 5, SPILL>(c, prm, sh, out, plan, a, stream); break;
        case 6: launch_one<G, 6, SPILL>(c, prm, sh, out, plan, a, stream); break;
        default: launch_one<G, 7, SPILL>(c, prm, sh, out, plan, a, stream); break;
    }
}
}  // namespace

namespace {
template <int MODE>
cudaError_t configure_helper() {
    using G = helper_geometry;
    return cudaFuncSetAttribute(mm1_ahead_kernel<G::THREADS, G::MIN_BLOCKS, G::RING_SLOTS, AH_EVENT_SLOTS, MODE, true, false, true>,
                                cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}
}  // namespace

cudaError_t configure_mm1_ahead(const mm1_fused_plan &plan) {
    (void)plan;
    cudaError_t e = configure_geometry<geometry0>();
    if (e == cudaSuccess) e = configure_helper<0>();
    if (e == cudaSuccess) e = configure_helper<1>();
    if (e == cudaSuccess) e = configure_helper<2>();
    if (e == cudaSuccess) e = configure_helper<3>();
    if (e == cudaSuccess) e = configure_helper<4>();
    if (e == cudaSuccess) e = configure_helper<5>();
    if (e == cudaSuccess) e = configure_helper<6>();
    if (e == cudaSuccess) e = configure_helper<7>();
    return e;
}

// One pass of the generate-ahead kernel.  spill: with the two-level ring.  list == nullptr: the whole
// shard (work counter); else the replications on `list`.  Replications it cannot finish are appended
// to fail_list.
namespace {
int constants_and_mode(mm1::constants &c, const cxxdes_b200_mm1_params &prm, const shard &sh, const mm1_fused_plan &plan) {
    c.lambda = make_divisor(prm.lambda);  // IEEE 1/d on the host: same bits as on the device
    c.mu = make_divisor(prm.mu);
    c.keys = philox_schedule_of((uint32_t)sh.seed, (uint32_t)(sh.seed >> 32));
    c.ticks_per_unit = (double)sh.clk.ticks_per_unit;
    c.sec_per_unit = sh.clk.seconds_per_tick_unit;
    c.tick_mult = (double)sh.clk.tick_mult;
    c.n_packets = (uint32_t)prm.n_packets;
    c.ring_cap = plan.ah_ring_cap;
    c.spill_cap = plan.ah_spill_cap;
    c.lead_cap = plan.ah_lead_cap;
    c.until = (sh.run_until < 0 || sh.run_until >= (int64_t)mm1::TIME_LIMIT) ? mm1::EMPTY : (uint32_t)sh.run_until;
    c.until64 = sh.run_until;
    // (with a wide clock a horizon beyond 2^30 ticks is still a horizon: the lanes watch until64 - epoch)
    const bool has_until = plan.ah_wide ? sh.run_until >= 0 : c.until != mm1::EMPTY;
    return (has_until ? mm1::MODE_UNTIL : 0) | (sh.clk.tick_mult != 1 ? mm1::MODE_TICK_MULT : 0) |
           (plan.ah_wide ? mm1::MODE_WIDE : 0);
}
}  // namespace

// The helper of a window-only first pass (see POLL above): AH_HELPER_BLOCKS one-warp blocks with the two-level ring
// that redo the replications on `list` while the first pass is still producing it.  `done_flag` must become non-zero
// (in stream order) after the first pass; `list` must have been filled with LIST_UNSET.
cudaError_t launch_mm1_ahead_helper(const cxxdes_b200_mm1_params &prm, const shard &sh, const device_columns &out,
                                    const mm1_fused_plan &plan, void *gring, const uint32_t *list, uint32_t list_capacity,
                                    unsigned long long *list_cursor, unsigned int *done_flag, uint32_t *fail_list,
                                    unsigned long long *fail_count, cudaStream_t stream) {
    mm1::constants c;
    const int mode = constants_and_mode(c, prm, sh, plan);
    // the helper's two-level ring uses its own (small) thread count as the column count of the global level
    pass_args a{(uint2 *)gring, nullptr, 0, list, nullptr, list_cursor, fail_list, fail_count, 0};
    switch (mode) {
        case 0: launch_helper_one<0>(c, prm, sh, out, plan, a, done_flag, list_capacity, stream); break;
        case 1: launch_helper_one<1>(c, prm, sh, out, plan, a, done_flag, list_capacity, stream); break;
        case 2: launch_helper_one<2>(c, prm, sh, out, plan, a, done_flag, list_capacity, stream); break;
        case 3: launch_helper_one<3>(c, prm, sh, out, plan, a, done_flag, list_capacity, stream); break;
        case 4: launch_helper_one<4>(c, prm, sh, out, plan, a, done_flag, list_capacity, stream); break;
        case 5: launch_helper_one<5>(c, prm, sh, out, plan, a, done_flag, list_capacity, stream); break;
        case 6: launch_helper_one<6>(c, prm, sh, out, plan, a, done_flag, list_capacity, stream); break;
        default: launch_helper_one<7>(c, prm, sh, out, plan, a, done_flag, list_capacity, stream); break;
    }
    return cudaGetLastError();
}

cudaError_t launch_mm1_ahead(const cxxdes_b200_mm1_params &prm, const shard &sh, const device_columns &out,
                             const mm1_fused_plan &plan, bool spill, void *gring, uint32_t *state, int resume,
                             const uint32_t *list, const unsigned long long *list_count, unsigned long long *list_cursor,
                             uint32_t *fail_list, unsigned long long *fail_count, cudaStream_t stream, int leave_blocks) {
    mm1::constants c;
    const int mode = constants_and_mode(c, prm, sh, plan);
    pass_args a{(uint2 *)gring, state, resume, list, list_count, list_cursor, fail_list, fail_count,
                leave_blocks < plan.ah_blocks ? leave_blocks : 0};
    if (spill) launch_geometry<geometry0, true>(mode, c, prm, sh, out, plan, a, stream);
    else launch_geometry<geometry0, false>(mode, c, prm, sh, out, plan, a, stream);
    return cudaGetLastError();
}

}  // namespace cxb
