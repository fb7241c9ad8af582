// device/engine32.cuh -- compact variant of the per-replication engine (device/engine.cuh) for
// models whose tokens all carry priority 0 and whose clock stays below 2^32 ticks.
//
// A token is a 32-bit time plus a 32-bit tag (8 bytes instead of 12), so twice as many
// replications fit into an SM's shared memory; with one thread per replication, resident
// warps are what hides the latency of the dependent shared-memory accesses of a heap
// operation.  Same exact libstdc++ moves as engine.cuh (tie order, stl_heap.h:135-148,224-270);
// a replication that would leave the range raises REP_TIME_RANGE and is redone by the 64-bit
// kernel (never a silently different order).
#pragma once
#include "engine.cuh"

namespace cxb {

struct token32 {
    uint32_t time;
    uint32_t tag;
};

// Heap entries are one 8-byte word {time, tag} per slot, slot-major with the thread index as the
// fastest dimension (STRIDE = threads per block, a compile-time constant so that slot addresses
// are shifts and adds): one LDS.64 / STS.64 per entry moved.
template <int STRIDE>
struct environment32 {
    uint32_t now;
    uint32_t n_events;
    uint64_t hash;
    uint32_t status;
    int n;    // heap size
    int cap;  // heap capacity
    uint2 *heap;  // this thread's column: entry i at heap[i * STRIDE]

    __device__ __forceinline__ void init(uint2 *column, int capacity) {
        now = 0;
        n_events = 0;
        hash = TRACE_HASH_INIT;
        status = 0;
        n = 0;
        cap = capacity;
        heap = column;
    }

    // Carve the heap of `capacity` entries per thread out of the block's dynamic shared memory.
    static __device__ __forceinline__ uint2 *carve_heap(unsigned char *&cursor, int capacity) {
        const size_t mis = (size_t)((uintptr_t)cursor & 7);
        unsigned char *p = cursor + ((8 - mis) & 7);
        cursor = p + sizeof(uint2) * (size_t)capacity * STRIDE;
        return (uint2 *)p + threadIdx.x;
    }

    // std::__push_heap, stl_heap.h:135-148; comp(parent, value) = parent strictly later.
    // The loop's loads form a dependent chain (each parent index comes from the previous compare), and with a few
    // resident warps per scheduler that latency is what the heap kernels wait on.  The indices of parent AND
    // grandparent are known up front, so both are loaded in one round trip and two levels are decided per trip;
    // the moves are exactly those of the one-level loop.
    __device__ __forceinline__ void sift_up(int hole, uint32_t key, uint32_t tag) {
        while (hole > 0) {
            const int parent = (hole - 1) >> 1;
            const int grand = parent > 0 ? (parent - 1) >> 1 : 0;
            const uint2 pe = heap[parent * STRIDE];
            const uint2 ge = heap[grand * STRIDE];
            if (!(pe.x > key)) break;
            heap[hole * STRIDE] = pe;
            hole = parent;
            if (parent == 0 || !(ge.x > key)) break;
            heap[hole * STRIDE] = ge;
            hole = grand;
        }
        heap[hole * STRIDE] = make_uint2(key, tag);
    }

    // environment::schedule_token, environment.ipp:94-98 (vector::push_back + push_heap).
    __device__ __forceinline__ void schedule(uint32_t time, uint32_t tag) {
        if (n >= cap) {
            status |= CXXDES_B200_REP_HEAP_OVERFLOW;
            return;
        }
        sift_up(n, time, tag);
        ++n;
    }

    // token at now + delay; the sum must stay inside 32 bits
    __device__ __forceinline__ void schedule_after(int64_t delay, uint32_t tag) {
        const uint64_t t = (uint64_t)now + (uint64_t)delay;
        if ((uint64_t)delay >= (1ULL << 31) || t >= 0xFFFFFFFFull) status |= CXXDES_B200_REP_TIME_RANGE;
        schedule((uint32_t)t, tag);
    }

    // std::pop_heap, stl_heap.h:254-270 (__pop_heap) + :224-249 (__adjust_heap), then pop_back.
    __device__ __forceinline__ token32 pop_top() {
        const uint2 first = heap[0];
        token32 top;
        top.time = first.x;
        top.tag = first.y;
        const int last = n - 1;
        if (last > 0) {
            const uint2 v = heap[last * STRIDE];
            const int len = last;
            int hole = 0;
            int child = 0;
            const int two_children = (len - 1) / 2;   // holes below this index have two children
            while (child < two_children) {
                child = 2 * (child + 1);
                uint2 er = heap[child * STRIDE];
                const uint2 el = heap[(child - 1) * STRIDE];
                if (child < two_children) {
                    // both children have two children themselves: fetch the four grandchildren in the same round
                    // trip and decide two levels (same moves as two turns of the one-level loop).  [Clamping the
                    // indices so that the last levels take this path too measured no better.]
                    const uint2 grr = heap[(2 * child + 2) * STRIDE], grl = heap[(2 * child + 1) * STRIDE];
                    const uint2 glr = heap[(2 * child) * STRIDE], gll = heap[(2 * child - 1) * STRIDE];
                    uint2 gr = grr, gl = grl;
                    if (er.x > el.x) {  // right strictly later than left -> take left; tie -> right
                        --child;
                        er = el;
                        gr = glr;
                        gl = gll;
                    }
                    heap[hole * STRIDE] = er;
                    hole = child;
                    child = 2 * (child + 1);
                    if (gr.x > gl.x) {
                        --child;
                        gr = gl;
                    }
                    heap[hole * STRIDE] = gr;
                    hole = child;
                    continue;
                }
                if (er.x > el.x) {
                    --child;
                    er = el;
                }
                heap[hole * STRIDE] = er;
                hole = child;
            }
            if ((len & 1) == 0 && child == (len - 2) / 2) {
                child = 2 * (child + 1);
                heap[hole * STRIDE] = heap[(child - 1) * STRIDE];
                hole = child - 1;
            }
            sift_up(hole, v.x, v.y);
        }
        n = last;
        return top;
    }

    // ---- the same two operations as straight-line code --------------------------------------------------------------
    // One thread per replication means a warp's 32 heaps are at 32 different places of their sifts; as loops, every
    // lane waits for the deepest one and the dependent chains of the divergent paths run one after the other.  Here a
    // sift is LEVELS predicated steps (LEVELS = depth of the deepest index the heap can have), no branch: all lanes
    // run the same instructions, and the compiler is free to interleave them with whatever else the event needs (the
    // variate of the next timeout does not depend on the heap).  Exactly the moves of pop_top() / schedule().
    template <int LEVELS>
    __device__ __forceinline__ token32 pop_flat() {
        const uint2 first = heap[0];
        token32 top;
        top.time = first.x;
        top.tag = first.y;
        const int len = n - 1;                               // n >= 1
        const uint2 v = heap[len * STRIDE];
        const int two_children = (len - 1) >> 1;             // holes below this index have two children (-1 when len == 0)
        int hole = 0;
#pragma unroll
        for (int k = 0; k < LEVELS; ++k) {
            const bool go = hole < two_children;
            const int right = go ? 2 * hole + 2 : 1;         // (a lane that has arrived reads entries 0 and 1, harmlessly)
            const uint2 er = heap[right * STRIDE], el = heap[(right - 1) * STRIDE];
            const bool left = er.x > el.x;                   // right strictly later than left -> take left; tie -> right
            const uint2 pick = left ? el : er;
            if (go) {
                heap[hole * STRIDE] = pick;
                hole = right - (left ? 1 : 0);
            }
        }
        if ((len & 1) == 0 && len > 0 && hole == (len - 2) / 2) {   // a lone left child at the bottom
            heap[hole * STRIDE] = heap[(2 * hole + 1) * STRIDE];
            hole = 2 * hole + 1;
        }
        // __push_heap of the displaced last element: it came from the bottom level and nearly always stays there
        if (len > 0) {
            while (hole > 0) {
                const int parent = (hole - 1) >> 1;
                const uint2 pe = heap[parent * STRIDE];
                if (!(pe.x > v.x)) break;
                heap[hole * STRIDE] = pe;
                hole = parent;
            }
            heap[hole * STRIDE] = v;
        }
        n = len;
        return top;
    }

    // schedule(time, tag) if `push`, as LEVELS predicated sift-up steps.
    template <int LEVELS>
    __device__ __forceinline__ void schedule_flat(bool push, uint32_t time, uint32_t tag) {
        if (push && n >= cap) {
            status |= CXXDES_B200_REP_HEAP_OVERFLOW;
            push = false;
        }
        int hole = n;
        bool moving = push;
#pragma unroll
        for (int k = 0; k < LEVELS; ++k) {
            const int parent = hole > 0 ? (hole - 1) >> 1 : 0;
            const uint2 pe = heap[parent * STRIDE];
            moving = moving && hole > 0 && pe.x > time;      // comp(parent, value): parent strictly later
            if (moving) {
                heap[hole * STRIDE] = pe;
                hole = parent;
            }
        }
        if (push) {
            heap[hole * STRIDE] = make_uint2(time, tag);
            ++n;
        }
    }

    // First half of environment::step (environment.ipp:117-126): pop, advance the clock, count.
    // The caller folds the token into the digest (it knows the tag layout).
    __device__ __forceinline__ token32 step_pop() {
        const token32 t = pop_top();
        now = t.time > now ? t.time : now;
        ++n_events;
        return t;
    }

    __device__ __forceinline__ void fold(uint32_t time, uint32_t kind, uint32_t proc) {
        hash = trace_hash_step(hash, (int64_t)time, 0, kind, proc);
    }

    __device__ __forceinline__ bool has_event() const { return n > 0; }
    __device__ __forceinline__ uint32_t next_time() const { return heap[0].x; }
};

#if defined(__CUDACC__)   // (explicit shared-memory instructions: device builds only; tests/native compiles the rest with g++)
// ---------------------------------------------------------------------------------------------------------------
// heap32x: the same heap (same layout, same libstdc++ moves) addressed through 32-bit shared-memory addresses with
// explicit ld/st.shared.  Through a generic `uint2 *` the compiler spends 5-7 instructions per entry on 64-bit address
// arithmetic (measured with ncu per source line: the two child loads of a sift-down level cost 9.6 thread instructions);
// here an entry's address is base + index * ENTRY_BYTES in one 32-bit multiply-add and neighbouring entries use
// immediate offsets.  All accesses are volatile asm
// statements (kept in program order among themselves); kernels that use this type touch their heap through it only.
// ---------------------------------------------------------------------------------------------------------------
template <int STRIDE>
struct heap32x {
    static constexpr uint32_t EB = (uint32_t)STRIDE * 8u;   // bytes between consecutive entries of one replication
    uint32_t now;
    uint32_t n_events;
    uint64_t hash;
    uint32_t status;
    uint32_t n;      // heap size
    uint32_t cap;
    uint32_t base;   // shared-memory address of entry 0 of this thread's column

    static constexpr int EXTRA = 0;      // (no sentinel entry in front of the column; see heap_rel)
    __device__ __forceinline__ void init(uint2 *column, int capacity) {
        now = 0;
        n_events = 0;
        hash = TRACE_HASH_INIT;
        status = 0;
        n = 0;
        cap = (uint32_t)capacity;
        base = (uint32_t)__cvta_generic_to_shared(column);
    }
    static __device__ __forceinline__ uint2 lds(uint32_t addr) {
        uint2 v;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
        return v;
    }
    static __device__ __forceinline__ void sts(uint32_t addr, uint2 v) {
        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y));
    }
    __device__ __forceinline__ uint32_t at(uint32_t index) const { return base + index * EB; }
    __device__ __forceinline__ uint2 entry(uint32_t index) const { return lds(at(index)); }
    __device__ __forceinline__ void set_entry(uint32_t index, uint2 v) const { sts(at(index), v); }
    static __device__ __forceinline__ uint32_t time_of(uint2 word) { return word.x; }
    static __device__ __forceinline__ uint32_t tag_of(uint2 word) { return word.y; }
    __device__ __forceinline__ token32 token_at(uint32_t index) const {
        const uint2 w = entry(index);
        token32 t;
        t.time = w.x;
        t.tag = w.y;
        return t;
    }
    __device__ __forceinline__ void set_token(uint32_t index, uint32_t time, uint32_t tag) const { set_entry(index, make_uint2(time, tag)); }
    __device__ __forceinline__ void restart_epoch() {}   // (absolute times: nothing to do; see heap_rel)

    // std::__push_heap from the hole at index `hole` (stl_heap.h:135-148; comp(parent, value) = parent strictly later).
    // Parent and grandparent are loaded in one round trip and two levels decided per trip (as environment32::sift_up:
    // with three or four resident warps per scheduler the kernels wait on these dependent round trips).
    __device__ __forceinline__ void sift_up(uint32_t hole, uint32_t time, uint32_t tag) const {
        while (hole != 0) {
            const uint32_t parent = (hole - 1) >> 1;
            const uint32_t grand = parent != 0 ? (parent - 1) >> 1 : 0u;
            const uint2 pe = lds(at(parent)), ge = lds(at(grand));
            if (!(pe.x > time)) break;
            sts(at(hole), pe);
            hole = parent;
            if (parent == 0 || !(ge.x > time)) break;
            sts(at(hole), ge);
            hole = grand;
        }
        sts(at(hole), make_uint2(time, tag));
    }

    // environment::schedule_token, environment.ipp:94-98 (vector::push_back + push_heap).
    __device__ __forceinline__ void schedule(uint32_t time, uint32_t tag) {
        if (n >= cap) {
            status |= CXXDES_B200_REP_HEAP_OVERFLOW;
            return;
        }
        sift_up(n, time, tag);
        ++n;
    }

    // The same __push_heap with every ancestor of the hole loaded at once.  Which entries a sift-up looks at does not
    // depend on the data: they are the hole's ancestors, (hole - 1) / 2 and so on up to the root.  So all LEVELS of
    // them (LEVELS = depth of the deepest index the heap can have; beyond the root the root again) come in ONE round
    // trip of independent loads instead of one dependent round trip per two levels, and because the times on a path to
    // the root never increase, "strictly later than the new token" holds for a prefix of them: those move down one
    // place on the path, the token lands above the last of them.  No loop, no divergence; a token pushed at `now`
    // (wake-ups, start and completion tokens: most of M/M/c's pushes) climbs nearly to the root, which cost the loop
    // three dependent round trips.  (A root that is loaded again for levels beyond it stores onto itself, before the
    // token's own store.)
    template <int LEVELS>
    __device__ __forceinline__ void sift_up_path(uint32_t hole, uint32_t time, uint32_t tag) const {
        uint32_t addr[LEVELS + 1];
        uint2 e[LEVELS];
        uint32_t a = hole;
        addr[0] = at(hole);
#pragma unroll
        for (int k = 0; k < LEVELS; ++k) {
            a = (a - (a != 0u ? 1u : 0u)) >> 1;
            addr[k + 1] = at(a);
            e[k] = lds(addr[k + 1]);
        }
        uint32_t dest = addr[0];
#pragma unroll
        for (int k = 0; k < LEVELS; ++k) {
            if (e[k].x > time) {
                sts(addr[k], e[k]);
                dest = addr[k + 1];
            }
        }
        sts(dest, make_uint2(time, tag));
    }
    template <int LEVELS>
    __device__ __forceinline__ void schedule_path(uint32_t time, uint32_t tag) {
        if (n >= cap) {
            status |= CXXDES_B200_REP_HEAP_OVERFLOW;
            return;
        }
        sift_up_path<LEVELS>(n, time, tag);
        ++n;
    }

    // std::pop_heap, stl_heap.h:254-270 (__pop_heap) + :224-249 (__adjust_heap), then pop_back -- for a caller that
    // has already loaded entry 0 (it decodes the event from it before the moves).  REINSERT_LEVELS > 0: the displaced last
    // element goes back in with a path-parallel sift-up from the leaf the hole ended at (one round trip of that many
    // ancestor loads) instead of the loop that usually stops after its first look: measured per kernel.
    template <int REINSERT_LEVELS = 0>
    __device__ __forceinline__ void pop_after_top() {
        const uint32_t len = n - 1;
        if (len > 0) {
            const uint2 v = lds(at(len));
            const uint32_t two = (len - 1) >> 1;                 // holes below this index have two children
            uint32_t hole = 0;
            while (hole < two) {
                const uint32_t right = 2 * hole + 2;
                const uint32_t ra = at(right);
                const uint2 er = lds(ra), el = lds(ra - EB);     // (one address, two immediate offsets)
                if (right < two) {
                    // both children have two children themselves: their four children come in the same round trip and
                    // two levels are decided (the moves of two turns of the one-level loop)
                    const uint32_t ga = at(2 * right);           // left child's right child; the others are EB apart
                    const uint2 grr = lds(ga + 2 * EB), grl = lds(ga + EB), glr = lds(ga), gll = lds(ga - EB);
                    const bool left = er.x > el.x;               // right strictly later than left -> take left; tie -> right
                    const uint32_t child = left ? right - 1 : right;
                    sts(at(hole), left ? el : er);
                    const uint2 gr = left ? glr : grr, gl = left ? gll : grl;
                    const bool left2 = gr.x > gl.x;
                    sts(left ? ra - EB : ra, left2 ? gl : gr);
                    hole = 2 * child + (left2 ? 1 : 2);
                    continue;
                }
                const bool left = er.x > el.x;
                sts(at(hole), left ? el : er);
                hole = left ? right - 1 : right;
            }
            if ((len & 1u) == 0 && hole == (len - 2) >> 1) {      // a lone left child at the bottom
                const uint32_t child = 2 * hole + 1;
                sts(at(hole), lds(at(child)));
                hole = child;
            }
            if constexpr (REINSERT_LEVELS > 0) sift_up_path<REINSERT_LEVELS>(hole, v.x, v.y);
            else sift_up(hole, v.x, v.y);
        }
        n = len;
    }

    __device__ __forceinline__ void fold(uint32_t time, uint32_t kind, uint32_t proc) {
        hash = trace_hash_step(hash, (int64_t)time, 0, kind, proc);
    }
    __device__ __forceinline__ bool has_event() const { return n > 0; }
    __device__ __forceinline__ uint32_t next_time() const { return lds(base).x; }
};

#endif  // __CUDACC__

// ---------------------------------------------------------------------------------------------------------------
// heap_rel<STRIDE, CODE_BITS>: 4-byte tokens, (time - epoch) << CODE_BITS | code -- for models whose tokens need only a
// short tag (ALOHA, 8 bits: which process; a station's loop counter lives in a per-station array instead.  M/M/c, 12 bits:
// handler flag + customer / acquire_proc ordinal below 2047) and whose LIVE tokens span less than 2^(32 - CODE_BITS) ticks.
// Half the shared memory of an 8-byte token buys more resident replications per SM (these kernels are bound by the
// dependent chain of each replication times the number of resident chains), and a level of a sift moves one word
// instead of two.  The heap orders by TIME ONLY, exactly as before -- equal times are resolved by libstdc++'s moves,
// never by the code in the low bits: `a later than b` is `a > (b | CODE_MASK)`.
// The epoch moves up to `now` whenever now - epoch reaches REBASE_AT (every live token is at or after now, so all
// relative times stay non-negative); a delay of DELAY_LIMIT ticks or more cannot be represented and raises
// REP_TIME_RANGE (the replication is redone by the 64-bit kernel).
// ---------------------------------------------------------------------------------------------------------------
template <int STRIDE, int CODE_BITS>
struct heap_rel {
    static constexpr uint32_t EB = (uint32_t)STRIDE * 4u;     // bytes between consecutive entries of one replication
    static constexpr int TIME_BITS = 32 - CODE_BITS;
    static constexpr uint32_t CODE_MASK = (1u << CODE_BITS) - 1u;
    static constexpr uint32_t REBASE_AT = 1u << (TIME_BITS - 2);
    static constexpr uint32_t DELAY_LIMIT = (1u << TIME_BITS) - REBASE_AT;
    uint32_t now;        // absolute ticks
    uint32_t epoch;      // absolute tick of relative time 0
    uint32_t n_events;
    uint64_t hash;
    uint32_t status;
    uint32_t n;          // heap size
    uint32_t cap;
    uint32_t base;       // shared-memory address of entry 0 of this thread's column

    // `column` holds EXTRA + capacity words: word 0 is a sentinel in front of entry 0 ("entry -1", relative time 0 and
    // code 0: never later than anything), which the path-parallel sift-up reads for the levels above the root.
    static constexpr int EXTRA = 1;
    __device__ __forceinline__ void init(uint32_t *column, int capacity) {
        now = 0;
        epoch = 0;
        n_events = 0;
        hash = TRACE_HASH_INIT;
        status = 0;
        n = 0;
        cap = (uint32_t)capacity;
        base = shared_address(column) + EB;
        sts(base - EB, 0u);
    }
#if defined(__CUDACC__)
    static __device__ __forceinline__ uint32_t shared_address(uint32_t *column) { return (uint32_t)__cvta_generic_to_shared(column); }
    static __device__ __forceinline__ uint32_t lds(uint32_t addr) {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
        return v;
    }
    static __device__ __forceinline__ void sts(uint32_t addr, uint32_t v) {
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v));
    }
#else   // tests/native: "shared memory" is the array CXB_HOST_SHARED names, addresses are byte offsets into it
    static uint32_t shared_address(uint32_t *column) { return (uint32_t)((column - CXB_HOST_SHARED) * 4); }
    static uint32_t lds(uint32_t addr) { return CXB_HOST_SHARED[addr / 4]; }
    static void sts(uint32_t addr, uint32_t v) { CXB_HOST_SHARED[addr / 4] = v; }
#endif
    static __device__ __forceinline__ bool later(uint32_t a, uint32_t b) { return a > (b | CODE_MASK); }   // a.time > b.time
    __device__ __forceinline__ uint32_t at(uint32_t index) const { return base + index * EB; }
    __device__ __forceinline__ uint32_t entry(uint32_t index) const { return lds(at(index)); }
    __device__ __forceinline__ void set_entry(uint32_t index, uint32_t v) const { sts(at(index), v); }
    __device__ __forceinline__ uint32_t time_of(uint32_t word) const { return epoch + (word >> CODE_BITS); }
    static __device__ __forceinline__ uint32_t tag_of(uint32_t word) { return word & CODE_MASK; }
    // a token as (absolute time, tag), for saved state; the epoch must be at or before every time stored
    __device__ __forceinline__ token32 token_at(uint32_t index) const {
        const uint32_t w = entry(index);
        token32 t;
        t.time = time_of(w);
        t.tag = tag_of(w);
        return t;
    }
    __device__ __forceinline__ void set_token(uint32_t index, uint32_t time, uint32_t tag) const {
        set_entry(index, ((time - epoch) << CODE_BITS) | tag);
    }
    __device__ __forceinline__ void restart_epoch() { epoch = now; }   // after `now` was restored, before set_token

    // std::__push_heap from the hole at index `hole` (stl_heap.h:135-148), parent and grandparent in one round trip
    __device__ __forceinline__ void sift_up(uint32_t hole, uint32_t word) const {
        const uint32_t top = word | CODE_MASK;
        while (hole != 0) {
            const uint32_t parent = (hole - 1) >> 1;
            const uint32_t grand = parent != 0 ? (parent - 1) >> 1 : 0u;
            const uint32_t pe = lds(at(parent)), ge = lds(at(grand));
            if (!(pe > top)) break;
            sts(at(hole), pe);
            hole = parent;
            if (parent == 0 || !(ge > top)) break;
            sts(at(hole), ge);
            hole = grand;
        }
        sts(at(hole), word);
    }

    __device__ __forceinline__ void rebase() {
        const uint32_t delta = now - epoch;
        for (uint32_t k = 0; k < n; ++k) sts(at(k), lds(at(k)) - (delta << CODE_BITS));
        epoch = now;
    }

    // environment::schedule_token (environment.ipp:94-98) of a token at absolute time `when` >= now
    __device__ __forceinline__ void schedule(uint32_t when, uint32_t code) {
        if (n >= cap) {
            status |= CXXDES_B200_REP_HEAP_OVERFLOW;
            return;
        }
        if (now - epoch >= REBASE_AT) rebase();
        if (when - now >= DELAY_LIMIT) {
            status |= CXXDES_B200_REP_TIME_RANGE;
            return;
        }
        sift_up(n, ((when - epoch) << CODE_BITS) | code);
        ++n;
    }

    // the same sift-up with all LEVELS ancestors loaded in one round trip (see heap32x::sift_up_path)
    template <int LEVELS>
    __device__ __forceinline__ void sift_up_path(uint32_t hole, uint32_t word) const {
        const uint32_t top = word | CODE_MASK;
        uint32_t addr[LEVELS + 1], e[LEVELS];
        // ancestor k of node h is ((h + 1) >> k) - 1; beyond the root that is -1, the sentinel in front of the column
        const uint32_t h1 = hole + 1u, before = base - EB;
        addr[0] = at(hole);
#pragma unroll
        for (int k = 0; k < LEVELS; ++k) {
            addr[k + 1] = before + (h1 >> (k + 1)) * EB;
            e[k] = lds(addr[k + 1]);
        }
        uint32_t dest = addr[0];
#pragma unroll
        for (int k = 0; k < LEVELS; ++k) {
            if (e[k] > top) {
                sts(addr[k], e[k]);
                dest = addr[k + 1];
            }
        }
        sts(dest, word);
    }
    template <int LEVELS>
    __device__ __forceinline__ void schedule_path(uint32_t when, uint32_t code) {
        if (n >= cap) {
            status |= CXXDES_B200_REP_HEAP_OVERFLOW;
            return;
        }
        if (now - epoch >= REBASE_AT) rebase();
        if (when - now >= DELAY_LIMIT) {
            status |= CXXDES_B200_REP_TIME_RANGE;
            return;
        }
        sift_up_path<LEVELS>(n, ((when - epoch) << CODE_BITS) | code);
        ++n;
    }

    // std::pop_heap (stl_heap.h:224-270) then pop_back, for a caller that has already loaded entry 0
    // (REINSERT_LEVELS: see heap32x::pop_after_top)
    template <int REINSERT_LEVELS = 0>
    __device__ __forceinline__ void pop_after_top() {
        const uint32_t len = n - 1;
        if (len > 0) {
            const uint32_t v = lds(at(len));
            const uint32_t two = (len - 1) >> 1;                 // holes below this index have two children
            uint32_t hole = 0;
            while (hole < two) {
                const uint32_t right = 2 * hole + 2;
                const uint32_t ra = at(right);
                const uint32_t er = lds(ra), el = lds(ra - EB);
                if (right < two) {
                    // both children have two children themselves: two levels per round trip (see heap32x)
                    const uint32_t ga = at(2 * right);
                    const uint32_t grr = lds(ga + 2 * EB), grl = lds(ga + EB), glr = lds(ga), gll = lds(ga - EB);
                    const bool left = later(er, el);             // right strictly later than left -> take left; tie -> right
                    const uint32_t child = left ? right - 1 : right;
                    sts(at(hole), left ? el : er);
                    const uint32_t gr = left ? glr : grr, gl = left ? gll : grl;
                    const bool left2 = later(gr, gl);
                    sts(left ? ra - EB : ra, left2 ? gl : gr);
                    hole = 2 * child + (left2 ? 1 : 2);
                    continue;
                }
                const bool left = later(er, el);
                sts(at(hole), left ? el : er);
                hole = left ? right - 1 : right;
            }
            if ((len & 1u) == 0 && hole == (len - 2) >> 1) {      // a lone left child at the bottom
                const uint32_t child = 2 * hole + 1;
                sts(at(hole), lds(at(child)));
                hole = child;
            }
            if constexpr (REINSERT_LEVELS > 0) sift_up_path<REINSERT_LEVELS>(hole, v);
            else sift_up(hole, v);
        }
        n = len;
    }

    // The same pop as straight-line code: ROUNDS two-level steps (ROUNDS = half the depth of the deepest index, rounded
    // up), each predicated instead of a loop turn, so that the whole event body is one basic block and the scheduler
    // can run the variate's arithmetic in the shadow of the sift-down's round trips.  Unlike the fixed-depth sift that
    // was rejected earlier (one dependent round trip per LEVEL) this keeps the loop's look-ahead: children and
    // grandchildren in one round trip, two levels decided per trip.  A step that has nothing to do reads entries 1-4,
    // and the look-ahead of the last internal node may read one entry past `cap`: the column must hold
    // max(cap + 1, 5) entries behind the sentinel (FLAT_POP_COLUMN).
    static __host__ __device__ constexpr int FLAT_POP_COLUMN(int capacity) { return (capacity + 1 > 5 ? capacity + 1 : 5) + EXTRA; }
    template <int ROUNDS, int REINSERT_LEVELS>
    __device__ __forceinline__ void pop_after_top_flat() {
        const uint32_t len = n - 1;
        const uint32_t v = lds(at(len));
        const uint32_t two = len ? (len - 1) >> 1 : 0u;          // holes below this index have two children
        uint32_t hole = 0;
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            const bool go1 = hole < two;
            const uint32_t right = go1 ? 2 * hole + 2 : 2u;
            const uint32_t ra = at(right);
            const uint32_t er = lds(ra), el = lds(ra - EB);
            // the chosen child has two children itself iff it is below `two`: both do when right < two, only the left
            // one when right == two
            const uint32_t ga = at((go1 && right <= two) ? 2 * right : 2u);
            const uint32_t grr = lds(ga + 2 * EB), grl = lds(ga + EB), glr = lds(ga), gll = lds(ga - EB);
            const bool left = later(er, el);                     // right strictly later than left -> take left; tie -> right
            const uint32_t child = left ? right - 1 : right;
            if (go1) sts(at(hole), left ? el : er);
            const bool go2 = go1 && child < two;
            const uint32_t gr = left ? glr : grr, gl = left ? gll : grl;
            const bool left2 = later(gr, gl);
            if (go2) sts(left ? ra - EB : ra, left2 ? gl : gr);
            hole = go2 ? 2 * child + (left2 ? 1u : 2u) : (go1 ? child : hole);
        }
        if ((len & 1u) == 0 && len != 0 && hole == (len - 2) >> 1) {   // a lone left child at the bottom
            const uint32_t child = 2 * hole + 1;
            sts(at(hole), lds(at(child)));
            hole = child;
        }
        if (len > 0) {
            if constexpr (REINSERT_LEVELS > 0) sift_up_path<REINSERT_LEVELS>(hole, v);
            else sift_up(hole, v);
        }
        n = len;
    }

    __device__ __forceinline__ void fold(uint32_t time, uint32_t kind, uint32_t proc) {
        hash = trace_hash_step(hash, (int64_t)time, 0, kind, proc);
    }
    __device__ __forceinline__ bool has_event() const { return n > 0; }
    __device__ __forceinline__ uint32_t next_time() const { return time_of(lds(base)); }
};

template <int STRIDE>
using heap24x = heap_rel<STRIDE, 8>;     // ALOHA: 24-bit relative time, 8-bit code

}  // namespace cxb
