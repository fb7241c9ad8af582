// device/engine.cuh -- per-replication event-dispatch engine for sm_100a (thread per replication).
//
// One CUDA thread owns one replication's environment (environment.ipp:12-279 in the
// reference): the simulated clock, the (time, priority) token heap, the event counter and
// the trace digest.  The heap lives in shared memory, structure-of-arrays with the thread
// index as the fastest dimension, so every heap access of a warp is bank-conflict free no
// matter which heap level each lane is at.
//
// Tie order: the reference's queue is std::priority_queue<token*, vector, token_comp>
// (environment.ipp:247-263); its comparator has no sequence number, so the order among
// equal (time, priority) keys is whatever libstdc++'s __push_heap/__adjust_heap do.  The
// push/pop below perform exactly those moves (bits/stl_heap.h:135-148, 224-270).
#pragma once
#include <stdint.h>

#include "cxxdes_b200.h"
#include "cxxdes_b200/shared/philox.hpp"
#include "cxxdes_b200/shared/trace_hash.hpp"
#include "cxxdes_b200/shared/variates.hpp"

namespace cxb {

using namespace cxxdes_b200;

// ---- token -------------------------------------------------------------------------
// token.ipp:6-62: {time, priority, coro_data, handler}.  On the device a token is a
// 64-bit key plus a 32-bit tag:
//   key  = time << 16 | (priority + 32768)      lexicographic (time, priority) as one integer
//   tag  = target process (low 16 bits, 0xFFFF = null coro_data)
//        | handler index + 1 (high 16 bits, 0 = null handler)
// Guards: 0 <= time < 2^47 and -32768 <= priority < 32768, else REP_TIME_RANGE is raised
// (never a silently wrong order).
constexpr uint32_t TAG_NO_TARGET = 0xFFFFu;
constexpr int64_t PRIO_INHERIT = INT64_MAX;  // priority_consts::inherit, defs.ipp:35
constexpr uint64_t KEY_EMPTY = ~0ULL;

struct token {
    uint64_t key;
    uint32_t tag;
};

__device__ __forceinline__ uint64_t make_key(int64_t time, int64_t prio, uint32_t &status) {
    if ((uint64_t)time >= (1ULL << 47) || prio < -32768 || prio > 32767)
        status |= CXXDES_B200_REP_TIME_RANGE;
    return ((uint64_t)time << 16) | (uint64_t)((prio + 32768) & 0xFFFF);
}
__device__ __forceinline__ int64_t key_time(uint64_t key) { return (int64_t)(key >> 16); }
__device__ __forceinline__ int64_t key_prio(uint64_t key) { return (int64_t)(key & 0xFFFF) - 32768; }
__device__ __forceinline__ uint32_t make_tag(uint32_t target, uint32_t handler_plus1 = 0) {
    return (target & 0xFFFFu) | (handler_plus1 << 16);
}
__device__ __forceinline__ uint32_t tag_target(uint32_t tag) { return tag & 0xFFFFu; }
__device__ __forceinline__ uint32_t tag_handler_plus1(uint32_t tag) { return tag >> 16; }

// ---- strided shared-memory array: element i of this thread lives at base[i * stride] ----
template <typename T>
struct lane_array {
    T *base;
    int stride;
    __device__ __forceinline__ T &operator[](int i) const { return base[i * stride]; }
};

// Carve `count` elements per thread of type T out of the block's dynamic shared memory.
// The cursor only ever moves by pointer arithmetic (the integer cast below computes the padding,
// not the pointer), so the compiler keeps seeing a pointer into the __shared__ array and emits
// LDS/STS; through an integer round trip it falls back to generic LD/ST, which costs latency
// and sits on the long scoreboard.
template <typename T>
__device__ __forceinline__ lane_array<T> carve(unsigned char *&cursor, int count) {
    const size_t mis = (size_t)((uintptr_t)cursor & (alignof(T) - 1));
    unsigned char *p = cursor + ((alignof(T) - mis) & (alignof(T) - 1));
    lane_array<T> a;
    a.base = (T *)p + threadIdx.x;
    a.stride = blockDim.x;
    cursor = p + sizeof(T) * (size_t)count * blockDim.x;
    return a;
}

// ---- environment ---------------------------------------------------------------------
struct environment {
    int64_t now;
    uint64_t n_events;
    uint64_t hash;
    uint32_t status;
    int n;    // heap size
    int cap;  // heap capacity
    lane_array<uint64_t> hkey;
    lane_array<uint32_t> htag;

    __device__ __forceinline__ void init(lane_array<uint64_t> k, lane_array<uint32_t> t, int capacity) {
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
    __device__ __forceinline__ void sift_up(int hole, uint64_t key, uint32_t tag) {
        while (hole > 0) {
            int parent = (hole - 1) >> 1;
            uint64_t pk = hkey[parent];
            if (!(pk > key)) break;
            hkey[hole] = pk;
            htag[hole] = htag[parent];
            hole = parent;
        }
        hkey[hole] = key;
        htag[hole] = tag;
    }

    // environment::schedule_token, environment.ipp:94-98 (vector::push_back + push_heap).
    __device__ __forceinline__ void schedule(int64_t time, int64_t prio, uint32_t tag) {
        if (n >= cap) {
            status |= CXXDES_B200_REP_HEAP_OVERFLOW;
            return;
        }
        uint64_t key = make_key(time, prio, status);
        sift_up(n, key, tag);
        ++n;
    }

    // std::pop_heap, stl_heap.h:254-270 (__pop_heap) + :224-249 (__adjust_heap), then pop_back.
    __device__ __forceinline__ token pop_top() {
        token top;
        top.key = hkey[0];
        top.tag = htag[0];
        int last = n - 1;
        if (last > 0) {
            uint64_t vkey = hkey[last];
            uint32_t vtag = htag[last];
            int len = last;
            int hole = 0;
            int child = 0;
            while (child < (len - 1) / 2) {
                child = 2 * (child + 1);
                uint64_t kr = hkey[child];
                uint64_t kl = hkey[child - 1];
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

    // First half of environment::step (environment.ipp:117-126): pop, advance the clock,
    // count the event and fold the token into the digest.
    __device__ __forceinline__ token step_pop() {
        token t = pop_top();
        int64_t time = key_time(t.key);
        now = time > now ? time : now;
        ++n_events;
        uint32_t target = tag_target(t.tag);
        uint32_t kind = tag_handler_plus1(t.tag) ? TOKEN_HANDLER
                                                 : (target != TAG_NO_TARGET ? TOKEN_RESUME : TOKEN_NOOP);
        hash = trace_hash_step(hash, time, key_prio(t.key), kind, target);
        return t;
    }

    __device__ __forceinline__ bool has_event() const { return n > 0; }
    __device__ __forceinline__ int64_t next_time() const { return key_time(hkey[0]); }
};

// ---- sync::event (event.hpp:87-139): FIFO list of waiting tokens {latency, priority, proc} ----
// Waiters are stored as (priority, proc) pairs; a wait latency other than 0 is kept per entry.
struct wait_list {
    lane_array<uint32_t> who;   // proc | (priority + 32768) << 16
    lane_array<int32_t> lat;    // latency in ticks (event.wait(latency, priority))
    int n;
    int cap;

    __device__ __forceinline__ void init(lane_array<uint32_t> w, lane_array<int32_t> l, int capacity) {
        who = w;
        lat = l;
        n = 0;
        cap = capacity;
    }

    // wait_awaitable::await_suspend, event.hpp:136-139.
    __device__ __forceinline__ void wait(environment &env, uint32_t proc, int64_t prio, int64_t latency = 0) {
        if (n >= cap || prio < -32768 || prio > 32767 || latency < 0 || latency > INT32_MAX) {
            env.status |= CXXDES_B200_REP_WAITER_OVERFLOW;
            return;
        }
        who[n] = (proc & 0xFFFFu) | ((uint32_t)(prio + 32768) << 16);
        lat[n] = (int32_t)latency;
        ++n;
    }

    // wake_awaitable::await_ready, event.hpp:125-134: schedule every waiter in list order.
    __device__ __forceinline__ void wake(environment &env) {
        for (int i = 0; i < n; ++i) {
            uint32_t w = who[i];
            env.schedule(env.now + lat[i], (int64_t)(w >> 16) - 32768, make_tag(w & 0xFFFFu));
        }
        n = 0;
    }
};

// ---- any_of / all_of handler (any_of.ipp:3-27, conditions :233-247) ----
struct any_all {
    uint32_t total;
    uint32_t remaining;
    bool is_all;
    bool armed;        // completion_tkn != nullptr
    int64_t out_latency;

    __device__ __forceinline__ void arm(bool all, uint32_t total_, uint32_t remaining_, int64_t latency = 0) {
        is_all = all;
        total = total_;
        remaining = remaining_;
        armed = true;
        out_latency = latency;
    }

    // custom_handler::invoke: the output token inherits the satisfying child's
    // time, priority and target (any_of.ipp:19-23).
    __device__ __forceinline__ void invoke(environment &env, const token &tkn) {
        --remaining;
        bool cond = is_all ? (remaining == 0) : (remaining < total);
        if (armed && cond) {
            env.schedule(out_latency + key_time(tkn.key), key_prio(tkn.key), make_tag(tag_target(tkn.tag)));
            armed = false;
        }
    }
};

// ---- random stream + clock shared by all models ----
struct exp_stream {
    stream_addr addr;
    divisor rate;
    uint64_t n;
    __device__ __forceinline__ void init(uint64_t seed, uint64_t rep, uint32_t stream, double lambda) {
        addr.seed = seed;
        addr.replication = rep;
        addr.stream = stream;
        rate = make_divisor(lambda);
        n = 0;
    }
    __device__ __forceinline__ double draw(const log_entry *table) {
        return exponential(addr, n++, rate, table);
    }
};

// Stage the 128-row log table into shared memory (2 KB); all threads of the block call this.
__device__ __forceinline__ void stage_log_table(log_entry *dst, const log_entry *src) {
    for (int i = threadIdx.x; i < (1 << CXB_LOG_TABLE_BITS); i += blockDim.x) dst[i] = src[i];
    __syncthreads();
}

}  // namespace cxb
