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

}  // namespace cxb
