// Scalar pieces of the NTT programmable bootstrap (modulus switch, monomial rotation, signed
// gadget decomposition), shared by the composed kernels (capi_pbs.cu) and the fused blind-rotation
// kernel (ntt_fast.cuh).  Each function names the reference code it reproduces bit for bit
// (paths under tfhe/src/core_crypto/).
#pragma once
#include <cstdint>

namespace nttb200 {
namespace pbs {

// fft_impl/common.rs:10-23
__host__ __device__ __forceinline__ uint64_t modulus_switch(uint64_t input, unsigned log_modulus) {
    if (log_modulus == 64) return input;
    return (input + (uint64_t(1) << (64 - log_modulus - 1))) >> (64 - log_modulus);
}

// pbs_modulus_switch_non_native, ntt64_pbs.rs:540-550 (divide_round, algorithms/misc.rs:6-18)
__host__ __device__ __forceinline__ uint64_t modulus_switch_non_native(uint64_t input, unsigned log2n,
                                                                       uint64_t p) {
    unsigned __int128 num = (unsigned __int128)input << (log2n + 1);
    unsigned __int128 d = num / p, r = num % p;
    return (uint64_t)d + (r >= (unsigned __int128)(p >> 1) ? 1 : 0);
}

// wrapping_neg_custom_mod (commons/numeric/unsigned.rs:219-225); p == 0 selects the native wrapping_neg
__host__ __device__ __forceinline__ uint64_t neg_mod(uint64_t a, uint64_t p) {
    if (p == 0) return uint64_t(0) - a;
    return a == 0 ? 0 : p - a;
}
// wrapping_sub_custom_mod (unsigned.rs:181-187)
__host__ __device__ __forceinline__ uint64_t sub_mod(uint64_t a, uint64_t b, uint64_t p) {
    return a >= b ? a - b : a - b + p;
}
// wrapping_add_custom_mod (unsigned.rs:174-179): a - neg(b)
__host__ __device__ __forceinline__ uint64_t add_mod(uint64_t a, uint64_t b, uint64_t p) {
    return sub_mod(a, neg_mod(b, p), p);
}

// Coefficient j of poly * X^a (polynomial_wrapping_monic_monomial_mul_assign[_custom_mod],
// algorithms/polynomial_algorithms.rs:462-507): rotate right by a mod N, negate the wrapped
// coefficients, negate everything once more when floor(a / N) is odd.
__host__ __device__ __forceinline__ uint64_t monomial_mul_coeff(const uint64_t* poly, size_t j, unsigned a,
                                                                unsigned log2n, uint64_t p) {
    unsigned n = 1u << log2n, rem = a & (n - 1);
    bool neg = ((a >> log2n) & 1u) ^ (j < rem);
    uint64_t v = poly[(j - rem) & (n - 1)];
    return neg ? neg_mod(v, p) : v;
}
// Coefficient j of poly / X^d (polynomial_wrapping_monic_monomial_div_assign[_custom_mod], :395-442)
__host__ __device__ __forceinline__ uint64_t monomial_div_coeff(const uint64_t* poly, size_t j, unsigned d,
                                                                unsigned log2n, uint64_t p) {
    unsigned n = 1u << log2n, rem = d & (n - 1);
    bool neg = ((d >> log2n) & 1u) ^ (j + rem >= n);
    uint64_t v = poly[(j + rem) & (n - 1)];
    return neg ? neg_mod(v, p) : v;
}

// decompose_one_level, commons/math/decomposition/iter.rs:130-151
__host__ __device__ __forceinline__ uint64_t decompose_one_level(unsigned base_log, uint64_t& state) {
    uint64_t mask = (uint64_t(1) << base_log) - 1;
    uint64_t res = state & mask;
    state = (uint64_t)((int64_t)state >> base_log);
    uint64_t carry = (((res - 1) | state) & res) >> (base_log - 1);
    state += carry;
    return res - (carry << base_log);
}

// SignedDecomposer::init_decomposer_state, decomposer.rs:204-236
__host__ __device__ __forceinline__ uint64_t init_decomposer_state_native(uint64_t input, unsigned base_log,
                                                                          unsigned level) {
    unsigned rep = level * base_log, non_rep = 64 - rep;
    uint64_t res = input >> (non_rep - 1);
    uint64_t rounding_bit = res & 1;
    res += 1;
    res >>= 1;
    res &= ~uint64_t(0) >> (64 - rep);
    uint64_t need_balance = (((res - 1) | (rounding_bit << (rep - 1))) & res) >> (rep - 1);
    return res - (need_balance << rep);
}

__host__ __device__ __forceinline__ unsigned ceil_ilog2(uint64_t x) {  // x >= 2
#ifdef __CUDA_ARCH__
    return 64 - __clzll(x - 1);
#else
    unsigned b = 0;
    for (uint64_t v = x - 1; v; v >>= 1) ++b;
    return b;
#endif
}

// TensorSignedDecompositionLendingIterNonNative::new, iter.rs:640-686, for one coefficient:
// state = closest_representable(|x|) >> (ceil_log2(p) - base_log*level) with |x| the centred
// absolute value (decomposer.rs:487-557, :25-49), neg = "x is in the upper half".
__host__ __device__ __forceinline__ uint64_t init_state_non_native(uint64_t x, unsigned base_log,
                                                                   unsigned level, uint64_t p, bool& neg) {
    uint64_t half_up = p / 2 + (p & 1);
    neg = !(x < half_up);
    uint64_t abs_value = neg ? p - x : x;  // p - x <= floor(p/2) < half_up: the inner sign is positive
    unsigned bits = ceil_ilog2(p), to_native = 64 - bits;
    unsigned shift = 64 - level * base_log - 1;
    uint64_t res = (abs_value << to_native) >> shift;
    res += 1;
    res &= ~uint64_t(1);
    uint64_t closest = (res << shift) >> to_native;
    return closest >> (bits - base_log * level);
}
// next_term, iter.rs:689-737
__host__ __device__ __forceinline__ uint64_t next_term_non_native(unsigned base_log, uint64_t& state, bool neg,
                                                                  uint64_t p) {
    uint64_t t = decompose_one_level(base_log, state);
    if (neg) t = uint64_t(0) - t;
    return (int64_t)t >= 0 ? t : p + t;
}

// level == 1: the whole non-native decomposition of one coefficient in closed form.  The centred
// absolute value is below 2^63 after the shift to the native width, so its rounded top base_log
// bits r = ((|x| << to_native >> (63 - base_log)) + 1) >> 1 are at most B/2: decompose_one_level
// returns r with no carry, and the term is r, or p - r for the upper half (0 stays 0).
// Identical to init_state_non_native + next_term_non_native (checked against them in the tests).
__host__ __device__ __forceinline__ uint64_t single_level_term_non_native(uint64_t x, unsigned base_log,
                                                                          uint64_t p) {
    uint64_t half_up = p / 2 + (p & 1);
    bool neg = !(x < half_up);
    uint64_t abs_value = neg ? p - x : x;
    uint64_t r = (((abs_value << (64 - ceil_ilog2(p))) >> (63 - base_log)) + 1) >> 1;
    return (neg && r) ? p - r : r;
}

// modswitch_from_ntt_prime_to_power_of_two, commons/math/ntt/ntt64.rs:184-196
__host__ __device__ __forceinline__ uint64_t modswitch_prime_to_pow2(uint64_t v, unsigned width, uint64_t p) {
    unsigned __int128 x = ((unsigned __int128)v << width) | (unsigned __int128)(p >> 1);
    uint64_t q = (uint64_t)(x / p);
    return width == 64 ? q : q << (64 - width);
}

// The same modswitch for p = 2^64 - 2^32 + 1 without a 128-bit division: with 2^64 = p + eps
// (eps = 2^32 - 1), x = xh*2^64 + xl = xh*p + y, y = xh*eps + xl = yh*p + z, z = yh*eps + yl < 3p,
// so floor(x / p) = xh + yh + [z >= p] + [z >= 2p].
__host__ __device__ __forceinline__ uint64_t modswitch_solinas_to_pow2(uint64_t v, unsigned width) {
    constexpr uint64_t P = 0xFFFFFFFF00000001ull, EPS = 0xFFFFFFFFull;
    unsigned __int128 x = ((unsigned __int128)v << width) | (unsigned __int128)(P >> 1);
    uint64_t xh = (uint64_t)(x >> 64), xl = (uint64_t)x;
    unsigned __int128 y = (unsigned __int128)xh * EPS + xl;
    uint64_t yh = (uint64_t)(y >> 64), yl = (uint64_t)y;
    unsigned __int128 z = (unsigned __int128)yh * EPS + yl;
    uint64_t q = xh + yh + (z >= (unsigned __int128)P ? 1 : 0) + (z >= ((unsigned __int128)P << 1) ? 1 : 0);
    return width == 64 ? q : q << (64 - width);
}

}  // namespace pbs
}  // namespace nttb200
