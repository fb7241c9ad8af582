// cxxdes_b200/shared/trace_hash.hpp -- order-sensitive digest of the dispatch sequence.
//
// The reference has no trace facility beyond debug prints (token.ipp:50-60).
// The parity contract "same event order as the reference engine" is checked by
// folding every token popped by environment::step (environment.ipp:117-146)
// into a 64-bit digest, using only the token's public fields
// (token.ipp:23-35): time, priority, whether a handler is attached, and which
// process (model-assigned ordinal) it resumes.  The oracle harness computes the
// same digest from env.next_event() before each env.step().
#pragma once
#include <stdint.h>
#include "philox.hpp"

namespace cxxdes_b200 {

enum token_kind : uint32_t {
    TOKEN_RESUME = 0,   // handler == null, coro_data != null  -> coroutine_data::resume
    TOKEN_HANDLER = 1,  // handler != null                     -> any_of / all_of bookkeeping
    TOKEN_NOOP = 2      // handler == null, coro_data == null  -> nothing (still one step)
};

constexpr uint32_t PROC_NONE = 0xFFFFu;  // token without a target coroutine
constexpr uint64_t TRACE_HASH_INIT = 0x243F6A8885A308D3ULL;
constexpr uint64_t TRACE_HASH_MUL = 0x9E3779B97F4A7C15ULL;

CXB_HD uint64_t rotl64(uint64_t x, int r) {
    return (x << r) | (x >> (64 - r));
}

CXB_HD uint64_t trace_hash_step(uint64_t h, int64_t time, int64_t priority, uint32_t kind,
                                uint32_t proc) {
    uint64_t x = (uint64_t)time ^ rotl64((uint64_t)priority, 21) ^ ((uint64_t)kind << 56) ^
                 ((uint64_t)(proc & 0xFFFFu) << 40);
    return (rotl64(h, 5) ^ x) * TRACE_HASH_MUL;
}

}  // namespace cxxdes_b200
