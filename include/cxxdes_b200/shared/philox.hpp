// cxxdes_b200/shared/philox.hpp -- counter-based random stream shared by host and device.
//
// The reference seeds std::mt19937_64 from std::random_device
// (examples/random_variable.hpp:17,33,35-37), so its runs are not reproducible.
// The ensemble engine replaces that with Philox4x32-10 (Salmon et al., SC'11)
// addressed by (seed, replication, stream, draw index): any replication can be
// drawn on any GPU, any host thread, in any order, and gives the same numbers.
//
// This header is plain C++ (no libm, no CUDA runtime calls) and is included
// unchanged by the nvcc device code, the host layer and the oracle.
//
// Stream addressing (one 128-bit Philox block yields two 64-bit draws):
//   key     = { lo32(seed), hi32(seed) }                 the same for every replication, so the ten
//                                                         round keys are warp-uniform on the device
//   counter = { lo32(draw >> 1), stream id | hi32(draw >> 1) << 16,
//               lo32(replication), hi32(replication) }    stream id < 2^16, draw < 2^49
//   draw d  = words {0,1} of the block if d is even, words {2,3} if d is odd.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CXB_HD __host__ __device__ __forceinline__
#else
#define CXB_HD inline __attribute__((always_inline))
#endif

namespace cxxdes_b200 {

struct philox_block {
    uint32_t w[4];
};

constexpr uint32_t PHILOX_M0 = 0xD2511F53u;
constexpr uint32_t PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u;
constexpr uint32_t PHILOX_W1 = 0xBB67AE85u;

CXB_HD void philox_mulhilo(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo) {
    uint64_t p = (uint64_t)a * (uint64_t)b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
}

// Philox4x32-10: ten rounds, key bumped by the Weyl constants between rounds.
CXB_HD philox_block philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                  uint32_t k0, uint32_t k1) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        philox_mulhilo(PHILOX_M0, c0, hi0, lo0);
        philox_mulhilo(PHILOX_M1, c2, hi1, lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += PHILOX_W0;
        k1 += PHILOX_W1;
    }
    philox_block b;
    b.w[0] = c0; b.w[1] = c1; b.w[2] = c2; b.w[3] = c3;
    return b;
}

// The ten round keys k + j*W of one 64-bit key, precomputed.  The ensemble uses the seed as the
// Philox key, so on the device the schedule is formed once on the host and read from the constant
// bank; philox4x32_10_scheduled(c, schedule_of(k)) == philox4x32_10(c, k) by construction.
struct philox_schedule {
    uint32_t k0[10];
    uint32_t k1[10];
};

inline philox_schedule philox_schedule_of(uint32_t k0, uint32_t k1) {
    philox_schedule s;
    for (int r = 0; r < 10; ++r) {
        s.k0[r] = k0 + (uint32_t)r * PHILOX_W0;
        s.k1[r] = k1 + (uint32_t)r * PHILOX_W1;
    }
    return s;
}

CXB_HD philox_block philox4x32_10_scheduled(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                            const philox_schedule &s) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        philox_mulhilo(PHILOX_M0, c0, hi0, lo0);
        philox_mulhilo(PHILOX_M1, c2, hi1, lo1);
        uint32_t n0 = hi1 ^ c1 ^ s.k0[r];
        uint32_t n2 = hi0 ^ c3 ^ s.k1[r];
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    philox_block b;
    b.w[0] = c0; b.w[1] = c1; b.w[2] = c2; b.w[3] = c3;
    return b;
}

// Address of one Philox block inside the ensemble's random stream.
struct stream_addr {
    uint64_t seed;
    uint64_t replication;
    uint32_t stream;  // process ordinal inside the model (producer = 1, consumer = 2, ...)
};

CXB_HD philox_block stream_block(const stream_addr &a, uint64_t block_index) {
    return philox4x32_10((uint32_t)block_index,
                         a.stream | ((uint32_t)(block_index >> 32) << 16),
                         (uint32_t)a.replication,
                         (uint32_t)(a.replication >> 32),
                         (uint32_t)a.seed,
                         (uint32_t)(a.seed >> 32));
}

// 64 raw bits of draw number `draw` of the addressed stream.
CXB_HD uint64_t stream_draw_bits(const stream_addr &a, uint64_t draw) {
    philox_block b = stream_block(a, draw >> 1);
    uint32_t lo = (draw & 1) ? b.w[2] : b.w[0];
    uint32_t hi = (draw & 1) ? b.w[3] : b.w[1];
    return ((uint64_t)hi << 32) | lo;
}

}  // namespace cxxdes_b200
