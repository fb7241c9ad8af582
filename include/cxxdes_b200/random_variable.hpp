// cxxdes_b200/random_variable.hpp -- the random source of a model, in the place of the reference's
// examples/random_variable.hpp (std::mt19937_64 seeded from std::random_device: not reproducible).
//
//     exponential_rv lambda{env, /*stream*/ 1, rate};   ...   co_await env.timeout(lambda());
//
// The SAME model source works on both engines:
//   * compiled against the recording <cxxdes/cxxdes.hpp> (cxxdes_b200/record), `lambda()` returns a tag and
//     `env.timeout(tag)` becomes "draw the next variate of the awaiting process's Philox stream" on the device;
//   * compiled against the reference's <cxxdes/cxxdes.hpp>, `lambda()` draws number n = 0, 1, 2, ... of stream
//     `stream` of the replication named by cxxdes_b200::rv_context -- the same numbers, bit for bit
//     (shared/philox.hpp, shared/variates.hpp; host code must be compiled with -ffp-contract=off).
// The device takes the stream id from the ordinal of the process that awaits (creation order: co_main 0, then 1, 2,
// ...), so give every variable the ordinal of the one process that uses it.
#pragma once
#include <cstdint>

#if defined(CXXDES_B200_RECORDING)
#include <cxxdes/cxxdes.hpp>
namespace cxxdes_b200 {
struct exponential_rv {
    std::uint32_t rate_index;
    template <typename Env>
    exponential_rv(Env &env, std::uint32_t /*stream*/, double rate) : rate_index(env.add_rate(rate)) {}
    cxxdes::core::variate_tag operator()() const { return {rate_index}; }
};
}  // namespace cxxdes_b200
#else
#include "cxxdes_b200/shared/philox.hpp"
#include "cxxdes_b200/shared/variates.hpp"
namespace cxxdes_b200 {
struct rv_context {   // which replication the model being constructed is (set by the host before the constructor runs)
    static std::uint64_t &seed() { static thread_local std::uint64_t v = 0; return v; }
    static std::uint64_t &replication() { static thread_local std::uint64_t v = 0; return v; }
};
struct exponential_rv {
    stream_addr addr;
    divisor rate;
    std::uint64_t n = 0;
    template <typename Env>
    exponential_rv(Env &, std::uint32_t stream, double lambda)
        : addr{rv_context::seed(), rv_context::replication(), stream}, rate{make_divisor(lambda)} {}
    double operator()() { return exponential(addr, n++, rate, LOG_TABLE_HOST); }
};
}  // namespace cxxdes_b200
#endif
