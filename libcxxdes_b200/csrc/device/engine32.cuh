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

struct environment32 {
    uint32_t now;
    uint32_t n_events;
    uint64_t hash;
    uint32_t status;
    int n;    // heap size
    int cap;  // heap capacity
    lane_array<uint32_t> hkey;
    lane_array<uint32_t> htag;

    __device__ __forceinline__ void init(lane_array<uint32_t> k, lane_array<uint32_t> t, int capacity) {
        now = 0;
        n_events = 0;
        hash = TRACE_HASH_INIT;
        status = 0;
        n = 0;
        cap = capacity;
        hkey = k;
        htag = t;
    }

    // std::__push_heap, stl_heap.h:135-148; comp(parent, value) = parent strictly later.
    __device__ __forceinline__ void sift_up(int hole, uint32_t key, uint32_t tag) {
        while (hole > 0) {
            const int parent = (hole - 1) >> 1;
            const uint32_t pk = hkey[parent];
            if (!(pk > key)) break;
            hkey[hole] = pk;
            htag[hole] = htag[parent];
            hole = parent;
        }
        hkey[hole] = key;
        htag[hole] = tag;
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
        token32 top;
        top.time = hkey[0];
        top.tag = htag[0];
        const int last = n - 1;
        if (last > 0) {
            const uint32_t vkey = hkey[last];
            const uint32_t vtag = htag[last];
            const int len = last;
            int hole = 0;
            int child = 0;
            while (child < (len - 1) / 2) {
                child = 2 * (child + 1);
                uint32_t kr = hkey[child];
                const uint32_t kl = hkey[child - 1];
                if (kr > kl) {  // right strictly later than left -> take left; tie -> right
                    --child;
                    kr = kl;
                }
                hkey[hole] = kr;
                htag[hole] = htag[child];
                hole = child;
            }
            if ((len & 1) == 0 && child == (len - 2) / 2) {
                child = 2 * (child + 1);
                hkey[hole] = hkey[child - 1];
                htag[hole] = htag[child - 1];
                hole = child - 1;
            }
            sift_up(hole, vkey, vtag);
        }
        n = last;
        return top;
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
    __device__ __forceinline__ uint32_t next_time() const { return hkey[0]; }
};

}  // namespace cxb
