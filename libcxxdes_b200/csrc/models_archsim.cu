// models_archsim.cu -- memory-hierarchy model (examples/basic_arch_sim.cpp:13-110) on the device engine.
//
// k requester processes issue loads to mem2{latency, capacity} backed by mem1{latency}.  Each
// memory::load is a coroutine (start + completion tokens) whose body runs inside
// _Co_with(mtx_) -- a second coroutine (co_with.ipp:27-35) that acquires the level's sync::mutex
// (mutex.hpp:31-109), runs the body and releases, waking every waiter.  With several requesters
// the mutex contention and wake storms are part of the trace.
//
// Token targets: co_main 0, requester q = 1 + q, mem2.load = 0x100 + q, its _Co_with = 0x8100 + q,
// mem1.load = 0x200 + q, its _Co_with = 0x8200 + q.  All priorities are 0.
// The block store is the model's (user code): an ordered array, eviction = least last_access,
// lowest position on ties, erase keeps order (same rule as the oracle's model).
#include <type_traits>

#include "device/models_common.cuh"
#include "device/trace.cuh"
#include "kernels.h"

namespace cxb {

namespace {

constexpr int ARCH_THREADS = 128;
constexpr int ARCH_MAX_REQ = 16;
constexpr int ARCH_MAX_CAP = 64;
constexpr int ARCH_MAX_BLOCKS_PER_SM = 8;

// mutex::event_ of a level when every waiter has priority 0 and latency 0 (all of them are _Co_with coroutines of this
// model): the requester's index alone, 8 bits.  Same interface as engine.cuh's wait_list.
struct wait_list_plain {
    lane_array<uint8_t> who;     // the requester's index; the ordinal is `base | index` (one list per memory level)
    uint32_t base;
    int n;
    int cap;
    __device__ __forceinline__ void init(lane_array<uint8_t> w, int capacity, uint32_t ordinal_base) {
        who = w;
        base = ordinal_base;
        n = 0;
        cap = capacity;
    }
    template <typename Env>
    __device__ __forceinline__ void wait(Env &env, uint32_t proc, int64_t) {      // event.hpp:136-139
        if (n >= cap) {
            env.status |= CXXDES_B200_REP_WAITER_OVERFLOW;
            return;
        }
        who[n++] = (uint8_t)(proc & 0xFFu);
    }
    template <typename Env>
    __device__ __forceinline__ void wake(Env &env) {                              // event.hpp:125-134, list order
        for (int i = 0; i < n; ++i) env.schedule(env.now, 0, make_tag(base | (uint32_t)who[i]));
        n = 0;
    }
};

// Per-requester state: one entry per requester in shared memory -- or, in the instantiation for the shipped
// sequential driver (one requester), a register.
template <typename T, bool SINGLE>
struct per_requester {
    lane_array<T> many;
    T one;
    __device__ __forceinline__ T &operator[](int q) {
        if constexpr (SINGLE) return one;
        else return many[q];
    }
};

}  // namespace

constexpr uint32_t ARCH_PHASE_SAVED = 1, ARCH_PHASE_DONE = 2;   // word 0 of a saved replication (0 = not started)
constexpr int ARCH_SAVED_HEADER = 24;

// SINGLE: exactly one requester (examples/basic_arch_sim.cpp as shipped, BASELINE config 5).
// STATEFUL: run_until / run_for in pieces (environment.ipp:190-214): a replication that stops at the horizon with
// tokens still scheduled is written to global memory, `[word][replication]` -- header (clock, counters, mutex flags,
// the all_of handler, statistics), the heap, the block store, the per-requester words and both wait lists -- and the
// next launch continues from it; finished replications keep phase DONE.  Separate instantiations, so that the plain
// run() path carries none of it.
// ENV = traced_environment: the same kernel writing the event trace of its replications (cxxdes_b200_trace).
// NARROW: the model's own arrays in half the shared memory -- time stamps in 32 bits, addresses in 16, wait lists as
// bare 16-bit ordinals -- chosen at launch when the run provably ends before tick 2^31 (archsim_narrow_applies: every
// load costs at most both latencies and in the worst case all of them run one after the other) and the address space
// fits (and fewer than 2^16 loads per requester: 16-bit loop counters; resume points in 2 bits each; wait lists as
// requester indices).  With four requesters 244 instead of 452 bytes per replication: seven resident blocks per SM instead of three
// (the kernel waits on dependent shared-memory round trips, three warps per scheduler did not cover them).
template <bool SINGLE, bool STATEFUL, typename ENV = environment, bool NARROW = false>
__global__ void __launch_bounds__(ARCH_THREADS)
archsim_kernel(cxxdes_b200_archsim_params prm, shard sh, device_columns out, archsim_saved st, trace_sink sink) {
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned char *cursor = smem;
    const int K = SINGLE ? 1 : (int)prm.n_requesters;
    const int CAP = (int)prm.l2_capacity;
    const int heap_cap = K + 3;
    using stamp_t = typename std::conditional<NARROW, uint32_t, int64_t>::type;   // a tick of the clock
    using addr_t = typename std::conditional<NARROW, uint16_t, uint32_t>::type;   // a block address
    using who_t = typename std::conditional<NARROW, uint8_t, uint32_t>::type;
    using loop_t = typename std::conditional<NARROW, uint16_t, uint32_t>::type;   // loads done (NARROW: fewer than 2^16)
    using pc_t = typename std::conditional<NARROW, uint16_t, uint32_t>::type;     // five resume points, PCW bits each
    constexpr int PCW = NARROW ? 2 : 4;
    lane_array<uint64_t> hkey = carve<uint64_t>(cursor, heap_cap);
    lane_array<stamp_t> blk_last = carve<stamp_t>(cursor, CAP);
    per_requester<stamp_t, SINGLE> t1, stamp2;
    per_requester<loop_t, SINGLE> loop_j;
    per_requester<pc_t, SINGLE> pcs;  // pcs: field 0 requester, 1 load2, 2 with2, 3 load1, 4 with1 (values 0..2)
    per_requester<addr_t, SINGLE> cur_addr;
    if (!SINGLE) {
        t1.many = carve<stamp_t>(cursor, K);
        stamp2.many = carve<stamp_t>(cursor, K);
    }
    lane_array<uint32_t> htag = carve<uint32_t>(cursor, heap_cap);
    if (!SINGLE) {
        if (!NARROW) {
            loop_j.many = carve<loop_t>(cursor, K);
            pcs.many = carve<pc_t>(cursor, K);
        }
    }
    lane_array<addr_t> blk_addr = carve<addr_t>(cursor, CAP);
    if (!SINGLE) {
        cur_addr.many = carve<addr_t>(cursor, K);
        if (NARROW) {
            loop_j.many = carve<loop_t>(cursor, K);
            pcs.many = carve<pc_t>(cursor, K);
        }
    }
    lane_array<who_t> w2who = carve<who_t>(cursor, K);
    lane_array<who_t> w1who = carve<who_t>(cursor, K);
    lane_array<int32_t> w2lat{}, w1lat{};          // (NARROW: every waiter has latency 0 and priority 0: no such arrays)
    if (!NARROW) {
        w2lat = carve<int32_t>(cursor, K);
        w1lat = carve<int32_t>(cursor, K);
    }

    constexpr uint32_t PCM = (1u << PCW) - 1u;
    auto get_pc = [&](int q, int which) -> uint32_t { return ((uint32_t)pcs[q] >> (PCW * which)) & PCM; };
    auto set_pc = [&](int q, int which, uint32_t v) {
        pcs[q] = (pc_t)(((uint32_t)pcs[q] & ~(PCM << (PCW * which))) | (v << (PCW * which)));
    };
    enum { PC_REQ = 0, PC_LOAD2 = 1, PC_WITH2 = 2, PC_LOAD1 = 3, PC_WITH1 = 4 };

    // One requester and a positive mem2 latency: every load starts later than the one before, so last_access
    // stamps are unique and the "lowest position on ties" rule of the eviction never applies.
    const bool unique_stamps = K == 1 && prm.l2_latency > 0;

    uint64_t rep;
    while (steal(sh, rep)) {
        const uint64_t grep = sh.first_replication + rep;
        stream_addr addr;
        addr.seed = sh.seed;
        addr.replication = grep;

        ENV env;
        env.init(hkey, htag, heap_cap);
        attach_trace(env, sink);
        typename std::conditional<NARROW, wait_list_plain, wait_list>::type wait2, wait1;  // mutex::event_ of mem2 / mem1
        if constexpr (NARROW) {
            wait2.init(w2who, K, 0x8100u);
            wait1.init(w1who, K, 0x8200u);
        } else {
            wait2.init(w2who, w2lat, K);
            wait1.init(w1who, w1lat, K);
        }
        bool owned2 = false, owned1 = false;
        int n_blocks = 0;
        int held_slot = -1;   // block found by the look-up of the load that holds mem2's mutex (-1: a miss)
        uint32_t main_pc = 0;
        any_all all;
        double total_latency = 0.0;
        uint64_t hits = 0, misses = 0;
        const uint32_t phase = STATEFUL && st.resume ? st.words[rep] : 0u;
        if (STATEFUL && phase == ARCH_PHASE_DONE) {   // now_ = max(now_, t), environment.ipp:194
            if (sh.run_until >= 0 && out.end_time[rep] < sh.run_until) out.end_time[rep] = sh.run_until;
            continue;
        }
        if (STATEFUL && phase == ARCH_PHASE_SAVED) {
            const uint32_t *sv = st.words + rep;
            const size_t n = st.n_replications;
            auto u64 = [&](int k) { return (uint64_t)sv[(size_t)k * n] | ((uint64_t)sv[(size_t)(k + 1) * n] << 32); };
            env.n = (int)sv[1 * n];
            env.now = (int64_t)u64(2);
            env.n_events = u64(4);
            env.hash = u64(6);
            const uint32_t flags = sv[8 * n];
            owned2 = flags & 1u; owned1 = (flags >> 1) & 1u; all.is_all = (flags >> 2) & 1u; all.armed = (flags >> 3) & 1u;
            main_pc = flags >> 8;
            n_blocks = (int)sv[9 * n];
            held_slot = (int)sv[10 * n];
            all.total = sv[11 * n];
            all.remaining = sv[12 * n];
            all.out_latency = 0;
            total_latency = bits_as_double(u64(13));
            hits = u64(15);
            misses = u64(17);
            wait2.n = (int)sv[19 * n];
            wait1.n = (int)sv[20 * n];
            int at = ARCH_SAVED_HEADER;
            for (int k = 0; k < env.n; ++k, at += 3) {
                hkey[k] = u64(at);
                htag[k] = sv[(size_t)(at + 2) * n];
            }
            at = ARCH_SAVED_HEADER + 3 * heap_cap;
            for (int b = 0; b < n_blocks; ++b, at += 3) {
                blk_addr[b] = (addr_t)sv[(size_t)at * n];
                blk_last[b] = (stamp_t)u64(at + 1);
            }
            at = ARCH_SAVED_HEADER + 3 * heap_cap + 3 * CAP;
            for (int q = 0; q < K; ++q, at += 11) {
                t1[q] = (stamp_t)u64(at);
                stamp2[q] = (stamp_t)u64(at + 2);
                loop_j[q] = (loop_t)sv[(size_t)(at + 4) * n];
                cur_addr[q] = (addr_t)sv[(size_t)(at + 5) * n];
                pcs[q] = (pc_t)sv[(size_t)(at + 6) * n];
                w2who[q] = (who_t)sv[(size_t)(at + 7) * n];
                w1who[q] = (who_t)sv[(size_t)(at + 9) * n];
                if constexpr (!NARROW) {
                    w2lat[q] = (int32_t)sv[(size_t)(at + 8) * n];
                    w1lat[q] = (int32_t)sv[(size_t)(at + 10) * n];
                }
            }
        } else {
            for (int q = 0; q < K; ++q) {
                pcs[q] = 0;
                loop_j[q] = 0;
            }
            env.schedule(0, 0, make_tag(0));
        }

        while (env.has_event() && (sh.run_until < 0 || env.next_time() <= sh.run_until)) {
            token t = env.step_pop();
            // An event pushes at most one token of its own (co_main's start and the mutex wake-ups aside).  The
            // branches only note it; the push -- range checks and the sift -- runs once after them for all lanes.
            // Order is kept: wherever an event also wakes waiters, the wake-ups come first (mutex.hpp:70-79).
            int64_t p_time = 0;
            uint32_t p_tag = 0;
            bool has_push = false;
            auto note = [&](int64_t time, uint32_t tag) {
                p_time = time;
                p_tag = tag;
                has_push = true;
            };
            const uint32_t target = tag_target(t.tag);
            if (tag_handler_plus1(t.tag)) {
                all.invoke(env, t);
            } else if (target == 0) {
                if (main_pc == 0) {
                    // all_of.range(requesters): bind in order, one handler on the completion tokens
                    for (int q = 0; q < K; ++q) env.schedule(env.now, 0, make_tag(1 + q));
                    all.arm(true, (uint32_t)K, (uint32_t)K);
                    main_pc = 1;
                } else {
                    note(env.now, make_tag(TAG_NO_TARGET));
                }
            } else if (target != TAG_NO_TARGET) {
                if (target < 0x100) {
                    // ---- requester q: for j < n_loads { t1 = now; co_await mem2.load(addr); total += now - t1 } ----
                    const int q = (int)target - 1;
                    uint32_t j = loop_j[q];
                    if (get_pc(q, PC_REQ) == 1) {
                        total_latency += (double)(env.now - (int64_t)t1[q]);
                        loop_j[q] = (loop_t)++j;
                    }
                    if (j < prm.n_loads) {
                        addr.stream = target;
                        cur_addr[q] = (addr_t)(stream_draw_bits(addr, j) % prm.addr_space);
                        t1[q] = (stamp_t)env.now;
                        set_pc(q, PC_LOAD2, 0);
                        note(env.now, make_tag(0x100 + q));  // start(mem2.load); completion targets the requester
                        set_pc(q, PC_REQ, 1);
                    } else {
                        note(env.now, make_tag(0, 1));  // co_return -> completion with the all_of handler
                    }
                } else {
                    const int q = (int)(target & 0xFF);
                    const bool is_with = (target & 0x8000u) != 0;
                    const int lvl = (target & 0x200u) ? 1 : 2;
                    if (!is_with) {
                        // ---- memory::load up to the _Co_with (basic_arch_sim.cpp:21-24) ----
                        const int which = lvl == 2 ? PC_LOAD2 : PC_LOAD1;
                        if (get_pc(q, which) == 0) {
                            if (lvl == 2) stamp2[q] = (stamp_t)env.now;  // timestamp = env.now()
                            set_pc(q, lvl == 2 ? PC_WITH2 : PC_WITH1, 0);
                            note(env.now, make_tag(target | 0x8000u));  // start(_Co_with coroutine)
                            set_pc(q, which, 1);
                        } else {
                            // load returns: completion resumes the awaiter (requester, or mem2's _Co_with)
                            note(env.now, make_tag(lvl == 2 ? 1 + q : 0x8100 + q));
                        }
                    } else {
                        // ---- operator+(mutex&, body): acquire; body(); release  (co_with.ipp:27-33) ----
                        const int which = lvl == 2 ? PC_WITH2 : PC_WITH1;
                        bool &owned = lvl == 2 ? owned2 : owned1;
                        auto &waiters = lvl == 2 ? wait2 : wait1;
                        const int64_t latency = lvl == 2 ? prm.l2_latency : prm.l1_latency;
                        const uint32_t pc = get_pc(q, which);
                        if (pc == 0) {
                            if (owned) {
                                waiters.wait(env, target, 0);  // mutex::acquire blocks (mutex.hpp:91-99)
                            } else {
                                owned = true;
                                bool miss = false;
                                if (lvl == 2) {
                                    const uint32_t a = cur_addr[q];
                                    int at = -1;
                                    for (int b = 0; b < n_blocks; ++b) at = blk_addr[b] == a ? b : at;
                                    miss = at < 0;
                                    held_slot = at;   // this requester holds mem2's mutex until it stamps the block
                                    if (miss) {
                                        ++misses;
                                        if (n_blocks == CAP) {
                                            // evict: least last_access, lowest position on ties
                                            int victim = 0;
                                            stamp_t best = blk_last[0];
                                            for (int b = 1; b < n_blocks; ++b) {
                                                stamp_t v = blk_last[b];
                                                if (v < best) { best = v; victim = b; }
                                            }
                                            if (unique_stamps) {
                                                // stamps never tie, so positions never decide: the last block moves
                                                // into the hole instead of every later block moving down one
                                                blk_addr[victim] = blk_addr[n_blocks - 1];
                                                blk_last[victim] = blk_last[n_blocks - 1];
                                            } else {
                                                for (int b = victim; b + 1 < n_blocks; ++b) {   // erase keeps order
                                                    blk_addr[b] = blk_addr[b + 1];
                                                    blk_last[b] = blk_last[b + 1];
                                                }
                                            }
                                            --n_blocks;
                                        }
                                    } else {
                                        ++hits;
                                    }
                                }
                                if (miss) {
                                    set_pc(q, PC_LOAD1, 0);
                                    note(env.now, make_tag(0x200 + q));  // co_await next_->load(addr)
                                    set_pc(q, which, 1);
                                } else {
                                    note(env.now + latency, make_tag(target));  // co_await timeout(latency_)
                                    set_pc(q, which, 2);
                                }
                            }
                        } else if (pc == 1) {
                            note(env.now + latency, make_tag(target));
                            set_pc(q, which, 2);
                        } else {
                            if (lvl == 2) {
                                // blocks_[addr].last_access = timestamp
                                // (nobody else touched mem2's blocks since the look-up: the mutex was held)
                                const stamp_t ts = stamp2[q];
                                if (held_slot >= 0) {
                                    blk_last[held_slot] = ts;
                                } else {
                                    blk_addr[n_blocks] = cur_addr[q];
                                    blk_last[n_blocks] = ts;
                                    ++n_blocks;
                                }
                            }
                            owned = false;      // handle.release(): owned_ = false; wake (mutex.hpp:70-79)
                            waiters.wake(env);
                            note(env.now, make_tag(target & 0x7FFFu));  // _Co_with returns -> resumes load
                        }
                    }
                }
            }
            if (has_push) env.schedule(p_time, 0, p_tag);
            if (env.status) break;
        }
        if (sh.run_until >= 0) {
            if (env.has_event()) env.status |= CXXDES_B200_REP_UNFINISHED;
            env.now = env.now > sh.run_until ? env.now : sh.run_until;
        }
        if (STATEFUL) {
            uint32_t *sv = st.words + rep;
            const size_t n = st.n_replications;
            if (env.status == CXXDES_B200_REP_UNFINISHED) {
                auto put64 = [&](int k, uint64_t v) {
                    sv[(size_t)k * n] = (uint32_t)v;
                    sv[(size_t)(k + 1) * n] = (uint32_t)(v >> 32);
                };
                sv[1 * n] = (uint32_t)env.n;
                put64(2, (uint64_t)env.now);
                put64(4, env.n_events);
                put64(6, env.hash);
                sv[8 * n] = (owned2 ? 1u : 0u) | (owned1 ? 2u : 0u) | (all.is_all ? 4u : 0u) | (all.armed ? 8u : 0u) | (main_pc << 8);
                sv[9 * n] = (uint32_t)n_blocks;
                sv[10 * n] = (uint32_t)held_slot;
                sv[11 * n] = all.total;
                sv[12 * n] = all.remaining;
                put64(13, (uint64_t)__double_as_longlong(total_latency));
                put64(15, hits);
                put64(17, misses);
                sv[19 * n] = (uint32_t)wait2.n;
                sv[20 * n] = (uint32_t)wait1.n;
                int at = ARCH_SAVED_HEADER;
                for (int k = 0; k < env.n; ++k, at += 3) {
                    put64(at, hkey[k]);
                    sv[(size_t)(at + 2) * n] = htag[k];
                }
                at = ARCH_SAVED_HEADER + 3 * heap_cap;
                for (int b = 0; b < n_blocks; ++b, at += 3) {
                    sv[(size_t)at * n] = blk_addr[b];
                    put64(at + 1, (uint64_t)blk_last[b]);
                }
                at = ARCH_SAVED_HEADER + 3 * heap_cap + 3 * CAP;
                for (int q = 0; q < K; ++q, at += 11) {
                    put64(at, (uint64_t)t1[q]);
                    put64(at + 2, (uint64_t)stamp2[q]);
                    sv[(size_t)(at + 4) * n] = loop_j[q];
                    sv[(size_t)(at + 5) * n] = cur_addr[q];
                    sv[(size_t)(at + 6) * n] = pcs[q];
                    sv[(size_t)(at + 7) * n] = w2who[q];
                    sv[(size_t)(at + 9) * n] = w1who[q];
                    if constexpr (NARROW) {
                        sv[(size_t)(at + 8) * n] = 0u;
                        sv[(size_t)(at + 10) * n] = 0u;
                    } else {
                        sv[(size_t)(at + 8) * n] = (uint32_t)w2lat[q];
                        sv[(size_t)(at + 10) * n] = (uint32_t)w1lat[q];
                    }
                }
                sv[0] = ARCH_PHASE_SAVED;
            } else {
                sv[0] = ARCH_PHASE_DONE;   // finished -- or failed: the status column says so, nothing to continue
            }
        }
        rep_out o;
        o.end_time = env.now;
        o.n_events = env.n_events;
        o.trace_hash = env.hash;
        o.status = env.status;
        o.f64[0] = total_latency;
        o.f64[1] = 0.0;
        o.i64[0] = (int64_t)hits;
        o.i64[1] = (int64_t)misses;
        store_result(out, rep, o);
    }
}

bool archsim_supported(const cxxdes_b200_archsim_params &p) {
    return p.n_requesters >= 1 && p.n_requesters <= ARCH_MAX_REQ && p.l2_capacity >= 1 &&
           p.l2_capacity <= ARCH_MAX_CAP && p.addr_space >= 1 && p.l2_latency >= 0 && p.l1_latency >= 0 &&
           p.n_loads < (1ULL << 32);
}

size_t archsim_smem_bytes(const cxxdes_b200_archsim_params &p) {
    const size_t k = p.n_requesters, cap = p.l2_capacity;
    return 128 + (size_t)ARCH_THREADS * ((k + 3) * 12 + cap * 12 + k * (8 + 8 + 4 + 4 + 4 + 16));
}

namespace {
// The clock of a run stays below the sum of every load's latencies taken one after the other: n_requesters * n_loads
// loads of at most l1 + l2 ticks (nothing else in the model lets time pass).  Below 2^31 the stamps fit 32 bits.
bool archsim_narrow_applies(const cxxdes_b200_archsim_params &p) {
    const double per_load = (double)p.l1_latency + (double)p.l2_latency + 1.0;
    return (double)p.n_requesters * (double)p.n_loads * per_load < 2147483648.0 && p.addr_space <= 65536 && p.n_loads < 65536;
}
size_t archsim_narrow_smem_bytes(const cxxdes_b200_archsim_params &p) {
    const size_t k = p.n_requesters, cap = p.l2_capacity;
    return 128 + (size_t)ARCH_THREADS * ((k + 3) * 12 + cap * (4 + 2) + k * (4 + 4 + 2 + 2 + 2 + 1 + 1));
}
}  // namespace

size_t archsim_saved_words(const cxxdes_b200_archsim_params &p) {
    return ARCH_SAVED_HEADER + 3 * ((size_t)p.n_requesters + 3) + 3 * (size_t)p.l2_capacity + 11 * (size_t)p.n_requesters;
}

cudaError_t launch_archsim(const cxxdes_b200_archsim_params &prm, const shard &sh, const device_columns &out,
                           int sm_count, cudaStream_t stream, const archsim_saved &st, bool wide_layout) {
    const bool narrow = !wide_layout && archsim_narrow_applies(prm);
    const size_t smem = narrow ? archsim_narrow_smem_bytes(prm) : archsim_smem_bytes(prm);
    auto go = [&](auto kernel, int per_sm) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kernel<<<sm_count * per_sm, ARCH_THREADS, smem, stream>>>(prm, sh, out, st, trace_sink{});
        return cudaGetLastError();
    };
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > ARCH_MAX_BLOCKS_PER_SM) per_sm = ARCH_MAX_BLOCKS_PER_SM;   // (72 registers: seven blocks of 128 threads fit an SM)
    if (narrow) {
        if (prm.n_requesters == 1)
            return st.words ? go(archsim_kernel<true, true, environment, true>, per_sm) : go(archsim_kernel<true, false, environment, true>, per_sm);
        return st.words ? go(archsim_kernel<false, true, environment, true>, per_sm) : go(archsim_kernel<false, false, environment, true>, per_sm);
    }
    if (prm.n_requesters == 1) return st.words ? go(archsim_kernel<true, true>, per_sm) : go(archsim_kernel<true, false>, per_sm);
    return st.words ? go(archsim_kernel<false, true>, per_sm) : go(archsim_kernel<false, false>, per_sm);
}

bool archsim_uses_narrow_layout(const cxxdes_b200_archsim_params &prm, bool wide_layout) {
    return !wide_layout && archsim_narrow_applies(prm);
}

// One replication (sh.n_replications == 1) with its event trace (the general k-requester instantiation).
cudaError_t launch_archsim_trace(const cxxdes_b200_archsim_params &prm, const shard &sh, const device_columns &out,
                                 const trace_sink &sink, cudaStream_t stream) {
    const size_t smem = archsim_smem_bytes(prm);
    auto kernel = archsim_kernel<false, false, traced_environment>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kernel<<<1, ARCH_THREADS, smem, stream>>>(prm, sh, out, archsim_saved{}, sink);
    return cudaGetLastError();
}

}  // namespace cxb
