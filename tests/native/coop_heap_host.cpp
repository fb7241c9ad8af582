// tests/native/coop_heap_host.cpp -- TEST INFRASTRUCTURE, never part of the product library.
//
// Runs the STEPS of the sub-warp cooperative heap (libcxxdes_b200/csrc/device/coop_heap.cuh: takes_left, walk,
// resting_level, ancestor / depth_of / levels_to_shift -- the functions the CUDA group executes) on the CPU, a loop
// over the 8 lanes standing in for the group and an OR of the lanes' predicates for the ballot, and checks every
// popped token against std::priority_queue with the reference's comparator (environment.ipp:247-263) on random
// push / pop sequences with heavy key ties: the SAME token must pop (the tag identifies it), i.e. libstdc++'s tie
// order survives the parallel formulation.  The array after every operation is compared as well.
#include <cstdint>
#include <cstdio>
#include <queue>
#include <random>
#include <vector>

#include "device/coop_heap.cuh"

using namespace cxb::coop;

namespace {

struct ref_token {
    uint32_t time, tag;
};
struct ref_comp {   // token_comp with equal priorities: a pops later when it is strictly later
    bool operator()(const ref_token &a, const ref_token &b) const { return a.time > b.time; }
};

// A priority_queue whose container can be looked at (the layout is what decides later ties).
struct open_queue : std::priority_queue<ref_token, std::vector<ref_token>, ref_comp> {
    const std::vector<ref_token> &array() const { return c; }
};

struct host_group {
    std::vector<entry> slots;
    int n = 0, cap;
    explicit host_group(int capacity) : slots(capacity + 2), cap(capacity) {}

    entry pop() {
        const entry top = slots[slot_of(0)];
        const int len = n - 1;
        const entry value = slots[slot_of(len > 0 ? len : 0)];
        const int node_rounds = (cap / 2 + G) / G;
        uint64_t bits = 0;
        for (int m = 0; m < node_rounds; ++m)            // one ballot per round
            for (int lane = 0; lane < G; ++lane)
                if (takes_left(slots.data(), lane + G * m, len)) bits |= 1ull << (lane + G * m);
        path_step s[G];
        entry x[G];
        uint32_t later = 0;
        for (int lane = 0; lane < G; ++lane) {            // every lane loads before any lane stores
            s[lane] = walk(bits, len, lane);
            x[lane] = value;
            if (lane < s[lane].levels) x[lane] = slots[slot_of(s[lane].from)];
            if (lane < s[lane].levels && x[lane].time > value.time) later |= 1u << lane;
        }
        const int rest = resting_level(later, s[0].levels);
        if (len > 0)
            for (int lane = 0; lane < G; ++lane) {
                if (lane < rest) slots[slot_of(s[lane].to)] = x[lane];
                if (lane == rest) slots[slot_of(rest < s[lane].levels ? s[lane].to : s[lane].leaf)] = value;
            }
        n = len;
        return top;
    }

    void push(uint32_t time, uint32_t tag) {
        const int at = n, depth = depth_of(at);
        entry y[G];
        uint32_t later = 0;
        for (int lane = 0; lane < G; ++lane) {
            y[lane] = entry{0, 0};
            if (lane < depth) y[lane] = slots[slot_of(ancestor(at, lane + 1))];
            if (lane < depth && y[lane].time > time) later |= 1u << lane;
        }
        const int shift = levels_to_shift(later, depth);
        for (int lane = 0; lane < G; ++lane) {
            if (lane < shift) slots[slot_of(ancestor(at, lane))] = y[lane];
            if (lane == shift) slots[slot_of(ancestor(at, shift))] = entry{time, tag};
        }
        n = at + 1;
    }
};

int check(uint32_t seed, int capacity, int n_ops, uint32_t time_span) {
    std::mt19937 rng(seed);
    host_group g(capacity);
    open_queue ref;
    uint32_t next_tag = 1, now = 0;
    for (int k = 0; k < n_ops; ++k) {
        const bool push = ref.empty() || ((int)ref.size() < capacity && rng() % 100 < 55);
        if (push) {
            const uint32_t t = now + rng() % time_span;       // few distinct times: many ties
            g.push(t, next_tag);
            ref.push(ref_token{t, next_tag});
            ++next_tag;
        } else {
            const entry got = g.pop();
            const ref_token want = ref.top();
            ref.pop();
            if (got.time != want.time || got.tag != want.tag) {
                std::fprintf(stderr, "seed %u capacity %d op %d: popped (%u, %u), std::priority_queue (%u, %u)\n", seed, capacity,
                             k, got.time, got.tag, want.time, want.tag);
                return 1;
            }
            now = got.time;
        }
        if (g.n != (int)ref.size()) return 1;
        for (int i = 0; i < g.n; ++i)
            if (g.slots[slot_of(i)].tag != ref.array()[i].tag) {
                std::fprintf(stderr, "seed %u capacity %d op %d: array differs at index %d\n", seed, capacity, k, i);
                return 1;
            }
    }
    return 0;
}

}  // namespace

int main() {
    int bad = 0;
    long ops = 0;
    for (uint32_t seed = 1; seed <= 40; ++seed)
        for (int capacity : {1, 2, 3, 4, 7, 8, 15, 16, 33, 64, 65, 66, 100, 126})
            for (uint32_t span : {1u, 2u, 5u, 1000u}) {
                bad += check(seed * 7919u + capacity, capacity, 4000, span);
                ops += 4000;
            }
    std::printf("cooperative heap steps vs std::priority_queue: %ld operations, %d mismatching runs\n", ops, bad);
    return bad ? 1 : 0;
}
