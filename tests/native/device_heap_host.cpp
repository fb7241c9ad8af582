// tests/native/device_heap_host.cpp -- TEST INFRASTRUCTURE, never part of the product library.
//
// Compiles the DEVICE event heaps (libcxxdes_b200/csrc/device/engine.cuh and engine32.cuh, the code the CUDA
// kernels run -- including the two-level look-ahead of the compact heap) with g++ and checks them against
// std::priority_queue with the reference's comparator (environment.ipp:247-263) on random push/pop sequences
// with heavy key ties: every popped token must be the SAME token (the tag identifies it), i.e. the tie order
// is libstdc++'s.  The CUDA keywords are defined away; one "thread", stride 1.
#include <cstdint>
#include <cstdio>
#include <queue>
#include <random>
#include <vector>

#define __device__
#define __host__
#define __forceinline__ inline
struct fake_dim3 { unsigned x, y, z; };
static fake_dim3 threadIdx{0, 0, 0}, blockDim{1, 1, 1};
static inline void __syncthreads() {}
struct uint2 { unsigned x, y; };
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
using std::min;
static uint32_t host_shared[4096];          // heap24x's "shared memory" in this build
#define CXB_HOST_SHARED host_shared

#include "device/engine32.cuh"

using namespace cxb;

namespace {

struct ref_token {
    uint64_t time;
    int64_t priority;
    uint32_t tag;
};
struct ref_comp {   // token_comp: a is "less" (pops later) when it is strictly later / lower priority
    bool operator()(const ref_token &a, const ref_token &b) const {
        return a.time > b.time || (a.time == b.time && a.priority > b.priority);
    }
};

int check32(uint32_t seed, int capacity, int n_ops, uint32_t time_span, bool flat = false) {
    std::mt19937 rng(seed);
    std::vector<uint2> storage(capacity);
    environment32<1> env;
    env.init(storage.data(), capacity);
    std::priority_queue<ref_token, std::vector<ref_token>, ref_comp> ref;
    uint32_t next_tag = 1, now = 0;
    for (int k = 0; k < n_ops; ++k) {
        const bool push = ref.empty() || ((int)ref.size() < capacity && rng() % 100 < 55);
        if (push) {
            const uint32_t t = now + rng() % time_span;       // few distinct times: many ties
            // (flat: the straight-line pop / push, LEVELS = 7 covers indices below 255)
            if (flat) env.schedule_flat<7>(true, t, next_tag);
            else env.schedule(t, next_tag);
            ref.push(ref_token{t, 0, next_tag});
            ++next_tag;
        } else {
            if (flat && (rng() & 3) == 0) env.schedule_flat<7>(false, now, 0);   // a push that is predicated off changes nothing
            token32 got;
            if (flat) {
                got = env.pop_flat<7>();
                env.now = got.time > env.now ? got.time : env.now;
            } else {
                got = env.step_pop();
            }
            const ref_token want = ref.top();
            ref.pop();
            if (got.time != want.time || got.tag != want.tag) {
                std::fprintf(stderr, "engine32: seed %u op %d: popped (%u, %u), std::priority_queue (%llu, %u)\n", seed, k,
                             got.time, got.tag, (unsigned long long)want.time, want.tag);
                return 1;
            }
            now = got.time;
        }
        if (env.n != (int)ref.size() || env.status) return 1;
    }
    return 0;
}

// heap_rel (heap24x = 8 code bits: ALOHA; 12 code bits: M/M/c): 4-byte tokens relative to a moving epoch.  Same check, with the clock running far enough for many rebases
// (time steps up to `time_span`, so a run of 3000 operations passes REBASE_AT several times when the span is large)
// and the code byte drawn so that it can never decide an order that libstdc++ would not.
template <int CODE_BITS>
int check_rel(uint32_t seed, int capacity, int n_ops, uint32_t time_span, long *rebases, bool path = false) {
    std::mt19937 rng(seed);
    uint32_t *column = host_shared + 16;
    heap_rel<1, CODE_BITS> env;
    constexpr uint32_t CODES = 1u << CODE_BITS;
    env.init(column, capacity);
    std::priority_queue<ref_token, std::vector<ref_token>, ref_comp> ref;
    std::vector<uint32_t> serial_of_code(CODES, 0);   // the token a code byte stands for (a code is live at most once)
    uint32_t next_serial = 1;
    for (int k = 0; k < n_ops; ++k) {
        const bool push = ref.empty() || ((int)ref.size() < capacity && rng() % 100 < 55);
        if (push) {
            uint32_t code = rng() & (CODES - 1);
            while (serial_of_code[code]) code = (code + 1) & (CODES - 1);
            const uint32_t t = env.now + rng() % time_span;
            const uint32_t epoch_before = env.epoch;
            // (path: every ancestor in one round trip, LEVELS = 8 covers indices below 511)
            if (path) env.template schedule_path<8>(t, code);
            else env.schedule(t, code);
            if (env.epoch != epoch_before) ++*rebases;
            if (env.status) return 1;
            serial_of_code[code] = next_serial;
            ref.push(ref_token{t, 0, next_serial});
            ++next_serial;
        } else {
            const uint32_t root = env.entry(0);
            // (re-insertion of the displaced element path-parallel too; every other run: the straight-line pop, 4 rounds
            // of two levels = indices below 511)
            if (path && (seed & 1)) env.template pop_after_top_flat<4, 8>();
            else if (path) env.template pop_after_top<8>();
            else env.pop_after_top();
            const uint32_t time = env.time_of(root), code = env.tag_of(root);
            env.now = time > env.now ? time : env.now;
            const ref_token want = ref.top();
            ref.pop();
            if (time != want.time || serial_of_code[code] != want.tag) {
                std::fprintf(stderr, "heap_rel: seed %u op %d: popped (%u, code %u), std::priority_queue (%llu, %u)\n", seed, k,
                             time, code, (unsigned long long)want.time, want.tag);
                return 1;
            }
            serial_of_code[code] = 0;
        }
        if (env.n != (uint32_t)ref.size() || env.status) return 1;
    }
    // a delay the 24-bit window cannot hold is refused, never wrapped
    if (env.n >= env.cap) return 0;
    env.schedule(env.now + heap_rel<1, CODE_BITS>::DELAY_LIMIT, 0);
    return (env.status & CXXDES_B200_REP_TIME_RANGE) ? 0 : 1;
}

int check64(uint32_t seed, int capacity, int n_ops, uint32_t time_span) {
    std::mt19937 rng(seed);
    std::vector<uint64_t> keys(capacity);
    std::vector<uint32_t> tags(capacity);
    environment env;
    lane_array<uint64_t> k{keys.data(), 1};
    lane_array<uint32_t> t{tags.data(), 1};
    env.init(k, t, capacity);
    std::priority_queue<ref_token, std::vector<ref_token>, ref_comp> ref;
    uint32_t next_tag = 1;
    int64_t now = 0;
    for (int op = 0; op < n_ops; ++op) {
        const bool push = ref.empty() || ((int)ref.size() < capacity && rng() % 100 < 55);
        if (push) {
            const int64_t time = now + rng() % time_span;
            const int64_t prio = (int64_t)(rng() % 3) - 1;    // priorities -1, 0, 1: ties on (time, priority)
            env.schedule(time, prio, next_tag);
            ref.push(ref_token{(uint64_t)time, prio, next_tag});
            ++next_tag;
        } else {
            const token got = env.pop_top();
            const ref_token want = ref.top();
            ref.pop();
            if ((uint64_t)key_time(got.key) != want.time || key_prio(got.key) != want.priority || got.tag != want.tag) {
                std::fprintf(stderr, "engine: seed %u op %d: wrong token popped\n", seed, op);
                return 1;
            }
            now = key_time(got.key);
        }
        if (env.n != (int)ref.size() || env.status) return 1;
    }
    return 0;
}

}  // namespace

int main() {
    int bad = 0;
    long ops = 0;
    for (uint32_t seed = 1; seed <= 40; ++seed) {
        for (int capacity : {2, 3, 7, 8, 33, 64, 66, 200}) {
            for (uint32_t span : {1u, 2u, 5u, 1000u}) {
                bad += check32(seed * 7919u + capacity, capacity, 3000, span);
                bad += check32(seed * 15485863u + capacity, capacity, 3000, span, true);
                bad += check64(seed * 104729u + capacity, capacity, 3000, span);
                ops += 9000;
            }
        }
    }
    long rebases = 0;
    for (uint32_t seed = 1; seed <= 40; ++seed) {
        for (int capacity : {2, 3, 7, 8, 33, 64, 66, 200}) {
            for (uint32_t span : {1u, 2u, 5u, 1000u, 3000000u}) {
                bad += check_rel<8>(seed * 32452843u + capacity, capacity, 3000, span, &rebases);
                bad += check_rel<8>(seed * 49979687u + capacity, capacity, 3000, span, &rebases, true);
                ops += 6000;
            }
            for (uint32_t span : {1u, 3u, 1000u, 190000u}) {       // 12 code bits: 20 bits of relative time
                bad += check_rel<12>(seed * 86028121u + capacity, capacity, 3000, span, &rebases);
                bad += check_rel<12>(seed * 67867967u + capacity, capacity, 3000, span, &rebases, true);
                ops += 6000;
            }
        }
    }
    if (rebases < 1000) {
        std::fprintf(stderr, "heap_rel: only %ld rebases exercised\n", rebases);
        ++bad;
    }
    std::printf("device heaps vs std::priority_queue: %ld operations, %d mismatching runs\n", ops, bad);
    return bad ? 1 : 0;
}
