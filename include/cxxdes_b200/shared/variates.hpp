// cxxdes_b200/shared/variates.hpp -- bit-reproducible variates and time conversion.
//
// Everything here is built from IEEE-754 binary64 +, -, *, correctly rounded
// fma and integer operations only, so g++ (-ffp-contract=off) and nvcc
// (-fmad=false) produce identical bits.  No libm: glibc's log and CUDA's log
// differ in the last place now and then, and after `* 1000` and truncation
// (environment::real_to_sim, environment.ipp:79-82) that would flip a tick.
//
// Reference arithmetic restated here:
//   * exponential variate  -log(1 - u) / lambda
//       libstdc++ std::exponential_distribution::operator(), bits/random.h:4900-4905,
//       called from examples/random_variable.hpp:24-26 (producer_consumer.cpp:34,50; aloha.cpp:52,68)
//   * env.timeout(x) -> ticks = (int64)(x * F), F = unit.count(precision)
//       timeout.ipp:184-187, environment.ipp:79-82, misc/time.hpp:75-86,146-148
//   * now_seconds() = (double)(now * prec.t) * conv[prec.u]
//       environment.ipp:29-36, misc/time.hpp:120-125
#pragma once
#include <stdint.h>
#include "philox.hpp"
#include "log_table.inc"

namespace cxxdes_b200 {

struct log_entry {
    uint64_t invc_bits;
    uint64_t logc_bits;
};

// Host copy of the table; device kernels stage the same rows into shared memory.
static const log_entry LOG_TABLE_HOST[1 << CXB_LOG_TABLE_BITS] = {CXB_LOG_TABLE_ROWS};

CXB_HD double bits_as_double(uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double d;
    __builtin_memcpy(&d, &b, 8);
    return d;
#endif
}

CXB_HD uint64_t double_as_bits(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t b;
    __builtin_memcpy(&b, &d, 8);
    return b;
#endif
}

CXB_HD double fma_rn(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}

// w = 1 - u with u = (top 52 bits) * 2^-52 in [0, 1); w is in (0, 1], never 0,
// so log(w) is finite.  Built as 2 - [1,2) to avoid an int->double conversion.
CXB_HD double unit_complement_from_bits(uint64_t bits) {
    double one_plus_u = bits_as_double(0x3FF0000000000000ULL | (bits >> 12));
    return 2.0 - one_plus_u;
}

// Series coefficients of log1p(r) - r = r^2 (c0 + c1 r + ... + c6 r^6), then ln2 split.  On the
// device they live in constant memory so every DFMA takes its coefficient straight from the
// constant bank (no per-use 64-bit immediate materialisation); same bits as the host literals.
#define CXB_LOG_COEFS {-1.0 / 8.0, 1.0 / 7.0, -1.0 / 6.0, 1.0 / 5.0, -1.0 / 4.0, 1.0 / 3.0, -0.5}
#if defined(__CUDACC__)
static __device__ __constant__ double CXB_LOG_COEF_DEV[7] = CXB_LOG_COEFS;
static __device__ __constant__ unsigned long long CXB_LN2_DEV[2] = {CXB_LN2_HI_BITS, CXB_LN2_LO_BITS};
#endif

// Natural logarithm for positive normal doubles (the engine only ever passes
// w in [2^-52, 1]).  Table-driven: w = 2^k * z, z in [0.6875, 1.375),
// r = z*invc - 1 (|r| < 2^-8; < 2^-7 next to 1.0 where c = 1), log(w) = k ln2 + logc +
// log1p(r), log1p by its Taylor series through r^8.  Measured against mpmath: < 1.5 ulp for w <= 0.68,
// absolute error < 2^-54 everywhere (tests/test_shared_math.py).
template <typename Table>
CXB_HD double log_pos(double w, const Table *table) {
    // CXB_LOG_OFF has a zero low word, so the index arithmetic only touches the high word.
    const uint64_t ix = double_as_bits(w);
    const uint32_t hx = (uint32_t)(ix >> 32);
    const uint32_t tmp = hx - (uint32_t)(CXB_LOG_OFF >> 32);
    const int i = (int)((tmp >> (20 - CXB_LOG_TABLE_BITS)) & ((1 << CXB_LOG_TABLE_BITS) - 1));
    const int k = (int)tmp >> 20;
    const uint64_t iz = ((uint64_t)(hx - (tmp & 0xFFF00000u)) << 32) | (uint32_t)ix;
    const double z = bits_as_double(iz);
    const double invc = bits_as_double(table[i].invc_bits);
    const double logc = bits_as_double(table[i].logc_bits);
#if defined(__CUDA_ARCH__)
    const double *c = CXB_LOG_COEF_DEV;
    const double ln2_hi = bits_as_double(CXB_LN2_DEV[0]), ln2_lo = bits_as_double(CXB_LN2_DEV[1]);
#else
    static const double c[7] = CXB_LOG_COEFS;
    const double ln2_hi = bits_as_double(CXB_LN2_HI_BITS), ln2_lo = bits_as_double(CXB_LN2_LO_BITS);
#endif
    const double r = fma_rn(z, invc, -1.0);
#if defined(__CUDA_ARCH__)
    // (double)k without the conversion unit: 2^52 + 2^31 + k sits exactly in the mantissa
    const double kd = __hiloint2double(0x43300000, k ^ (int)0x80000000) - 4503601774854144.0;
#else
    const double kd = (double)k;
#endif
    const double t = fma_rn(kd, ln2_hi, logc);
    const double hi = t + r;
    const double lo = (t - hi) + r;
    const double r2 = r * r;
    // log1p(r) - r = r^2 * (-1/2 + r/3 - r^2/4 + r^3/5 - r^4/6 + r^5/7 - r^6/8)
    double p = fma_rn(r, c[0], c[1]);
    p = fma_rn(r, p, c[2]);
    p = fma_rn(r, p, c[3]);
    p = fma_rn(r, p, c[4]);
    p = fma_rn(r, p, c[5]);
    p = fma_rn(r, p, c[6]);
    const double tail = fma_rn(kd, ln2_lo, r2 * p);
    return hi + (lo + tail);
}

// x / d for a divisor that is reused many times: q0 = x * (1/d), one fma
// residual, one fma correction (Markstein).  With inv = RN(1/d) the result was
// TESTED equal to IEEE `x / d` on 10^7 random pairs (tests/test_shared_math.py);
// it is not proven to be the correctly rounded quotient for every divisor.
// Parity does not depend on that: host oracle and device both compute the
// variate with this function, so they agree bit for bit either way.
struct divisor {
    double d;
    double inv;
};

CXB_HD divisor make_divisor(double d) {
    divisor v;
    v.d = d;
    v.inv = 1.0 / d;
    return v;
}

CXB_HD double divide(double x, const divisor &v) {
    const double q0 = x * v.inv;
    const double rem = fma_rn(-q0, v.d, x);
    return fma_rn(rem, v.inv, q0);
}

// -log(w) / lambda, w = 1 - u.
template <typename Table>
CXB_HD double exponential_from_bits(uint64_t bits, const divisor &lambda, const Table *table) {
    const double w = unit_complement_from_bits(bits);
    const double neg_log = -log_pos(w, table);
    return divide(neg_log, lambda);
}

// One exponential variate = draw number `draw` of the addressed stream.
template <typename Table>
CXB_HD double exponential(const stream_addr &a, uint64_t draw, const divisor &lambda,
                          const Table *table) {
    return exponential_from_bits(stream_draw_bits(a, draw), lambda, table);
}

// Simulation clock configuration folded to two numbers.
//   ticks_per_unit   = time_unit().count(time_precision())        (misc/time.hpp:75-86)
//   seconds_per_tick = conv[precision.u], tick_mult = precision.t (misc/time.hpp:120-125)
struct clock_config {
    int64_t ticks_per_unit;
    int64_t tick_mult;
    double seconds_per_tick_unit;
};

// environment::real_to_sim(double): double * intmax_t -> double, then
// conversion to time_integral truncates toward zero.
CXB_HD int64_t real_to_sim(double x, const clock_config &c) {
    return (int64_t)(x * (double)c.ticks_per_unit);
}

// environment::real_to_sim(float): float * intmax_t -> float (aloha frame_time).
CXB_HD int64_t real_to_sim_f32(float x, const clock_config &c) {
    return (int64_t)(x * (float)c.ticks_per_unit);
}

// environment::now_seconds().
CXB_HD double sim_to_seconds(int64_t now, const clock_config &c) {
    return (double)(now * c.tick_mult) * c.seconds_per_tick_unit;
}

}  // namespace cxxdes_b200
