// device/coop_heap.cuh -- the event heap of ONE replication operated by a GROUP of 8 lanes (sub-warp cooperative
// push / pop with ballot), still move-for-move libstdc++'s binary heap (bits/stl_heap.h:135-148 __push_heap,
// :224-249 __adjust_heap, :254-270 __pop_heap), i.e. the reference's tie order
// (environment.ipp:94-98,117-126,247-263: std::priority_queue<token*, vector, token_comp>).
//
// Why it can be parallel at all:
//   pop   __adjust_heap moves the hole from the root to a leaf choosing, at every node, between the node's two children
//         by comparing THEM with each other -- never with the value that is being re-inserted.  So "which child would a
//         hole at node p take" is one bit per node that depends only on the current array: the lanes compute those bits
//         for all nodes at once (one 16-byte load of the adjacent children per node), a ballot gathers them, the path is
//         a walk over that bit mask in registers, the values along the path move up one level in parallel (lane k moves
//         level k), and the final __push_heap of the displaced last element compares it with the values on the path --
//         monotone along the path, so one more ballot and a population count give its resting level.
//   push  __push_heap compares the new value with its ancestors only, whose indices follow from the insertion index:
//         lane k loads ancestor k + 1, one ballot, a population count, lanes below the resting level shift down.
// Two shared-memory round trips per operation instead of one per level.
//
// The steps are plain functions of (array, lane, ...) so that tests/native/coop_heap_host.cpp can run them on the CPU with
// a loop over the lanes in place of the warp, against std::priority_queue.  Entries are {time, tag} (engine32.cuh); the
// comparator is strict on time (all priorities equal).  Entry i lives in slot i + 1: the children 2p + 1, 2p + 2 of
// node p are then the aligned 16-byte pair of slots 2p + 2, 2p + 3.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CXB_COOP_HD __host__ __device__ __forceinline__
#else
#define CXB_COOP_HD inline
#endif

namespace cxb {
namespace coop {

constexpr int G = 8;              // lanes per replication
constexpr int MAX_LEVELS = 7;     // heap indices < 2^(MAX_LEVELS + 1) - 1 = 255; a path has at most MAX_LEVELS moves (< G)

struct entry {
    uint32_t time, tag;
};

CXB_COOP_HD int slot_of(int index) { return index + 1; }

// ---- pop, step 1: does a hole at node p take its LEFT child?  (len = size after the pop) ----
// __adjust_heap: secondChild = 2 (hole + 1); if (comp(first + secondChild, first + secondChild - 1)) --secondChild;
// comp(a, b) = a strictly later than b, so the left child is taken only when the right one is strictly later.
// Nodes without two children inside [0, len) give 0 (the walk never asks for them).
CXB_COOP_HD bool takes_left(const entry *slots, int p, int len) {
    if (2 * p + 2 >= len) return false;
    const entry left = slots[slot_of(2 * p + 1)], right = slots[slot_of(2 * p + 2)];
    return right.time > left.time;
}

// ---- pop, step 2: the path of the hole, from the bit mask (bit p = takes_left(p)) ----
// Level k of the path moves the value at index from_k into index to_k (to_0 = 0, to_{k+1} = from_k); `levels` moves in
// all, the hole ends at `leaf`.  A lane only needs its own level, so the walk hands (to, from) of level `my_level` back.
struct path_step {
    int to, from;    // of this lane's level (undefined when my_level >= levels)
    int levels;      // number of moves, <= MAX_LEVELS
    int leaf;        // where the hole ends
};
CXB_COOP_HD path_step walk(uint64_t bits, int len, int my_level) {
    path_step s;
    s.to = s.from = 0;
    int hole = 0, child = 0, level = 0;
    const int two_children = (len - 1) / 2;   // nodes below this index have two children
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < MAX_LEVELS; ++k) {
        if (child < two_children) {
            child = 2 * (child + 1);
            child -= (int)((bits >> hole) & 1u);
            if (level == my_level) { s.to = hole; s.from = child; }
            hole = child;
            ++level;
        }
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {   // a lone left child at the bottom
        child = 2 * (child + 1);
        if (level == my_level) { s.to = hole; s.from = child - 1; }
        hole = child - 1;
        ++level;
    }
    s.levels = level;
    s.leaf = hole;
    return s;
}

// ---- pop, step 3: __push_heap(hole = leaf, top = 0, value) on the path ----
// After the moves the ancestors of the leaf hold x_0 .. x_{levels-1} (x_k = the value that moved into level k); they are
// ordered along the path, so "x_k strictly later than the value" holds exactly for the levels k >= k*.  The value rests
// at level k*: levels below k* keep their moved-up value, levels from k* on keep what they had before the pop.
// later_bits: bit k = (x_k.time > value.time) for k < levels.  Returns k*.
CXB_COOP_HD int resting_level(uint32_t later_bits, int levels) {
    const uint32_t active = (1u << levels) - 1u;
    const uint32_t not_later = ~later_bits & active;
#if defined(__CUDA_ARCH__)
    return __popc(not_later);
#else
    return __builtin_popcount(not_later);
#endif
}

// ---- push: ancestor k of insertion index n (ancestor 0 = n itself), and how many there are above n ----
CXB_COOP_HD int ancestor(int n, int k) { return ((n + 1) >> k) - 1; }
CXB_COOP_HD int depth_of(int n) {   // number of proper ancestors = floor(log2(n + 1))
#if defined(__CUDA_ARCH__)
    return 31 - __clz(n + 1);
#else
    return 31 - __builtin_clz((unsigned)(n + 1));
#endif
}
// later_bits: bit k = (value of ancestor k + 1 strictly later than the new time), k < depth.  Monotone from the bottom:
// the new value passes exactly the first `shift` ancestors.
CXB_COOP_HD int levels_to_shift(uint32_t later_bits, int depth) {
    const uint32_t active = (1u << depth) - 1u;
#if defined(__CUDA_ARCH__)
    return __popc(later_bits & active);
#else
    return __builtin_popcount(later_bits & active);
#endif
}

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------------------------------------------
// The group on the device.  All 32 lanes of the warp must call pop / push together (they contain full-warp ballots);
// a group without work passes enabled = false.  `slots` is the group's array in shared memory (16-byte aligned).
// ---------------------------------------------------------------------------------------------------------------
struct group_heap {
    entry *slots;
    int n;        // heap size (every lane of the group holds the same value)
    int cap;
    int lane;     // 0 .. G-1 within the group
    int shift;    // bit position of the group's lanes in a warp ballot

    __device__ __forceinline__ void attach(entry *group_slots, int capacity) {
        slots = group_slots;
        cap = capacity;
        n = 0;
        lane = (int)(threadIdx.x & (G - 1));
        shift = (int)(threadIdx.x & 31u & ~(unsigned)(G - 1));
    }
    __device__ __forceinline__ uint32_t group_ballot(bool pred) const {
        return (__ballot_sync(0xFFFFFFFFu, pred) >> shift) & ((1u << G) - 1u);
    }

    // NODE_ROUNDS * G nodes can have two children: capacity <= 2 * NODE_ROUNDS * G + 1.
    template <int NODE_ROUNDS>
    __device__ __forceinline__ entry pop(bool enabled) {
        entry top = slots[slot_of(0)];
        const int len = enabled ? n - 1 : 0;
        const entry value = slots[slot_of(len > 0 ? len : 0)];   // the last element, re-inserted from the root
        uint64_t bits = 0;
#pragma unroll
        for (int m = 0; m < NODE_ROUNDS; ++m) {
            const int p = lane + G * m;
            bool left = false;
            if (2 * p + 2 < len) {
                const uint4 pair = *reinterpret_cast<const uint4 *>(slots + slot_of(2 * p + 1));   // {left, right}
                left = pair.z > pair.x;
            }
            bits |= (uint64_t)group_ballot(left) << (G * m);
        }
        const path_step s = walk(bits, len, lane);
        const bool mover = lane < s.levels;
        entry x = value;
        if (mover) x = slots[slot_of(s.from)];
        const int rest = resting_level(group_ballot(mover && x.time > value.time), s.levels);
        __syncwarp();                                   // every load of the group before any store
        if (len > 0) {
            if (lane < rest) slots[slot_of(s.to)] = x;
            if (lane == rest) slots[slot_of(rest < s.levels ? s.to : s.leaf)] = value;
        }
        __syncwarp();
        if (enabled) n = len;
        return top;
    }

    // Returns false (and changes nothing) when the heap is full.
    __device__ __forceinline__ bool push(bool enabled, uint32_t time, uint32_t tag) {
        const bool ok = enabled && n < cap;
        const int at = ok ? n : 0;
        const int depth = depth_of(at);
        const bool has = ok && lane < depth;
        entry y;
        y.time = 0; y.tag = 0;
        if (has) y = slots[slot_of(ancestor(at, lane + 1))];
        const int shift_levels = levels_to_shift(group_ballot(has && y.time > time), depth);
        __syncwarp();
        if (ok) {
            if (lane < shift_levels) slots[slot_of(ancestor(at, lane))] = y;
            if (lane == shift_levels) {
                entry e;
                e.time = time; e.tag = tag;
                slots[slot_of(ancestor(at, shift_levels))] = e;
            }
            n = at + 1;
        }
        __syncwarp();
        return ok || !enabled;
    }
};
#endif  // __CUDACC__

}  // namespace coop
}  // namespace cxb
