// The reference's own tests of the CRT plans (tfhe-ntt/src/native32.rs, native64.rs:1180-1244, native128.rs,
// native_binary*.rs: `negacyclic_polymul` equals the wrapping schoolbook negacyclic convolution), written against
// the C++ host mirror of the API, for all ten plan types; plus `fwd_binary` == `fwd` on binary input, the
// `try_new` -> None cases, fastdiv and the prime helpers.  Self-checking: returns 0 and prints "cpp native ok".
#include <array>
#include <cstdio>
#include <random>
#include <vector>

#include "tfhe_ntt_b200.hpp"

using namespace tfhe_ntt;

template <class V>
static std::vector<V> schoolbook(const std::vector<V>& l, const std::vector<V>& r) {
    const size_t n = l.size();
    std::vector<V> out(n, 0);
    for (size_t i = 0; i < n; ++i)
        for (size_t j = 0; j < n; ++j) {
            V t = (V)(l[i] * r[j]);  // wrapping in the value word
            if (i + j < n)
                out[i + j] = (V)(out[i + j] + t);
            else
                out[i + j - n] = (V)(out[i + j - n] - t);
        }
    return out;
}

template <class V>
static V random_value(std::mt19937_64& g) {
    if constexpr (sizeof(V) == 16)
        return ((V)g() << 64) | g();
    else
        return (V)g();
}

template <class Plan>
static int check_plan(size_t n, bool binary_rhs, uint64_t seed) {
    using V = typename Plan::value_type;
    using R = typename Plan::residue_type;
    constexpr int NP = Plan::num_primes;
    auto plan = Plan::try_new(n);
    if (!plan) return 1;
    if (plan->ntt_size() != n) return 2;
    std::mt19937_64 g(seed);
    std::vector<V> l(n), r(n), prod(n);
    for (auto& v : l) v = random_value<V>(g);
    for (auto& v : r) v = binary_rhs ? (V)(g() & 1) : random_value<V>(g);
    l[0] = ~(V)0;  // the largest value of the word
    plan->negacyclic_polymul(prod.data(), n, l.data(), n, r.data(), n);
    if (prod != schoolbook(l, r)) return 3;
    // batched form: two products in one call
    std::vector<V> l2(2 * n), r2(2 * n), p2(2 * n);
    for (size_t i = 0; i < n; ++i) l2[i] = l[i], r2[i] = r[i], l2[n + i] = r[i], r2[n + i] = binary_rhs ? r[i] : l[i];
    plan->negacyclic_polymul_batch(p2.data(), l2.data(), r2.data(), 2);
    for (size_t i = 0; i < n; ++i)
        if (p2[i] != prod[i]) return 4;
    {
        std::vector<V> a(l2.begin() + n, l2.end()), b(r2.begin() + n, r2.end());
        auto want = schoolbook(a, b);
        for (size_t i = 0; i < n; ++i)
            if (p2[n + i] != want[i]) return 5;
    }
    // length assertion (native64.rs:1049-1051)
    try {
        plan->negacyclic_polymul(prod.data(), n, l.data(), n / 2, r.data(), n);
        return 6;
    } catch (const std::logic_error&) {
    }
    // ntt_0() .. : the per-prime plans; the residues of fwd are their transforms of value mod p_i
    for (int i = 0; i < NP; ++i)
        if (plan->ntt_i(i).ntt_size() != n || plan->ntt_i(i).modulus() < 2) return 8;
    if (plan->ntt_0().modulus() == (NP > 1 ? plan->ntt_i(NP - 1).modulus() : 0)) return 9;
    try {
        plan->ntt_i(NP);
        return 10;
    } catch (const std::out_of_range&) {
    }
    // fwd writes NP residue polynomials below their primes; on the binary plans fwd_binary agrees with fwd
    std::vector<std::vector<R>> res(NP, std::vector<R>(n)), res_b(NP, std::vector<R>(n));
    std::array<R*, NP> ptr, ptr_b;
    for (int i = 0; i < NP; ++i) ptr[i] = res[i].data(), ptr_b[i] = res_b[i].data();
    plan->fwd(r.data(), n, ptr);
    {
        // residue 0 is ntt_0().fwd of the value reduced modulo its prime (native64.rs:970-998)
        auto p0 = plan->ntt_0();
        std::vector<R> direct(n);
        for (size_t i = 0; i < n; ++i) direct[i] = (R)((unsigned __int128)r[i] % (unsigned __int128)p0.modulus());
        p0.fwd(direct);
        if (direct != res[0]) return 11;
    }
    if constexpr (Plan::is_binary) {
        if (binary_rhs) {
            plan->fwd_binary(r.data(), n, ptr_b);
            if (res != res_b) return 7;
        }
    }
    return 0;
}

int main() {
    int rc;
#define RUN(PLAN, N, BIN, TAG)                                        \
    if ((rc = check_plan<PLAN>(N, BIN, TAG)) != 0) {                  \
        std::printf("%s n=%d failed at check %d\n", #PLAN, N, rc);    \
        return TAG;                                                   \
    }
    RUN(native32::Plan32, 32, false, 11)
    RUN(native32::Plan52, 64, false, 12)
    RUN(native64::Plan32, 64, false, 13)
    RUN(native64::Plan52, 32, false, 14)
    RUN(native128::Plan32, 32, false, 15)
    RUN(native_binary32::Plan32, 32, true, 16)
    RUN(native_binary32::Plan52, 64, true, 17)
    RUN(native_binary64::Plan32, 128, true, 18)
    RUN(native_binary64::Plan52, 32, true, 19)
    RUN(native_binary128::Plan32, 64, true, 20)
    // sizes served by the one-CTA-per-product kernels
    RUN(native64::Plan32, 1024, false, 21)
    RUN(native128::Plan32, 1024, false, 22)
    RUN(native_binary64::Plan52, 2048, true, 23)

    // try_new -> None: not a power of two, below the smallest size, beyond the 2-adicity of the primes (native64.rs:932-941)
    if (native64::Plan32::try_new(48) || native64::Plan32::try_new(16) || native64::Plan32::try_new(65536)) return 30;
    if (!native64::Plan32::try_new(32768)) return 31;

    // fastdiv (fastdiv.rs:29-150): exact quotient and remainder by an invariant divisor
    {
        std::mt19937_64 g(5);
        for (int i = 0; i < 1000; ++i) {
            uint32_t d32 = (uint32_t)g() | 1u, n32 = (uint32_t)g();
            uint64_t d64 = g() | 1u, n64 = g();
            fastdiv::Div32 a(d32);
            fastdiv::Div64 b(d64);
            if (fastdiv::Div32::div(n32, a) != n32 / d32 || fastdiv::Div32::rem(n32, a) != n32 % d32) return 40;
            if (fastdiv::Div32::div_u64(n64, a) != n64 / d32 || fastdiv::Div32::rem_u64(n64, a) != n64 % d32) return 41;
            if (fastdiv::Div64::div(n64, b) != n64 / d64 || fastdiv::Div64::rem(n64, b) != n64 % d64) return 42;
            unsigned __int128 w = ((unsigned __int128)g() << 64) | g();
            if (fastdiv::Div64::div_u128(w, b) != w / d64 || fastdiv::Div64::rem_u128(w, b) != (uint64_t)(w % d64)) return 43;
        }
    }
    // prime helpers (prime.rs:188-222)
    if (!prime::is_prime64(prime64::SOLINAS_PRIME) || prime::is_prime64(prime64::SOLINAS_PRIME - 1)) return 50;
    {
        auto q = prime::largest_prime_in_arithmetic_progression64(1ull << 16, 1, 0, 1ull << 32);
        if (!q || !prime::is_prime64(*q) || (*q & 0xFFFF) != 1 || *q >= (1ull << 32)) return 51;
        if (!prime32::Plan::try_new(32768, (uint32_t)*q)) return 52;
    }
    std::printf("cpp native ok\n");
    return 0;
}
