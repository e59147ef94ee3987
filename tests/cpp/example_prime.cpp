// The crate doc example of the reference (tfhe-ntt/src/lib.rs:25-49) and its
// examples/mul_poly_prime.rs, written against the C++ host mirror of the API.
#include <cstdio>
#include <vector>

#include "tfhe_ntt_b200.hpp"

int main() {
    using namespace tfhe_ntt;
    const size_t N = 32;
    const uint32_t p = 1062862849;
    auto plan = prime32::Plan::try_new(N, p);
    if (!plan) return 2;
    std::vector<uint32_t> data(N), t;
    for (size_t i = 0; i < N; ++i) data[i] = (uint32_t)i;
    t = data;
    plan->fwd(t);
    plan->inv(t);
    for (size_t i = 0; i < N; ++i)
        if (t[i] != (uint64_t)data[i] * N % p) return 1;  // roundtrip is x * N

    // negacyclic product through the NTT domain (mul_poly_prime.rs): (1 + x) * x^(N-1) = x^(N-1) - 1
    std::vector<uint32_t> a(N, 0), b(N, 0);
    a[0] = 1, a[1] = 1, b[N - 1] = 1;
    plan->fwd(a);
    plan->fwd(b);
    plan->mul_assign_normalize(a.data(), N, b.data(), N);
    plan->inv(a);
    if (a[N - 1] != 1 || a[0] != p - 1) return 3;
    for (size_t i = 1; i + 1 < N; ++i)
        if (a[i] != 0) return 4;

    // try_new -> None (prime64.rs:1988-1990) and the length assertion (prime64.rs:898)
    if (prime64::Plan::try_new(2048, 1024)) return 5;
    auto p64 = prime64::Plan::try_new(64, prime64::SOLINAS_PRIME);
    std::vector<uint64_t> wrong(32);
    try {
        p64->fwd(wrong);
        return 6;
    } catch (const std::logic_error&) {
    }
    auto nat = native64::Plan32::try_new(64);
    std::vector<uint64_t> l(64, 3), r(64, 0), prod(64);
    r[0] = 5;
    nat->negacyclic_polymul(prod.data(), 64, l.data(), 64, r.data(), 64);
    for (auto v : prod)
        if (v != 15) return 7;
    // custum_radix: the sequence of the fork's test (custum_radix/fwd_1.rs:433-463), here with the inverse table
    // so that it is a round trip: n = 8, p = 17, root = 3^2 = 9 of order 8, tw[k] = 9^k, inv[k] = tw[k]^-1
    {
        const uint32_t q = 17;
        std::vector<uint32_t> tw{1, 9, 13, 15, 16, 8, 4, 2}, inv{1, 2, 4, 8, 16, 15, 13, 9};
        std::vector<uint32_t> v{5, 11, 3, 12, 8, 13, 2, 14}, w = v, u;
        custum_radix::fft_split_radix_recursive(w, tw, q);
        u = v;
        custum_radix::fft_radix4_recursive(u, tw, q);
        if (u != w) return 8;  // one function on a power table
        uint32_t sum = 0;
        for (auto x : v) sum = (sum + x) % q;
        if (w[0] != sum) return 9;  // X[0] = sum of the inputs
        custum_radix::ifft_radix2_recursive(w, inv, q, 15 /* 8^-1 mod 17 */, true);
        if (w != v) return 10;
        // the `_mut` routines count their products (fwd_1.rs:3-37): radix-2 on 8 non-zero inputs = 4 size-2 bases
        // + 4 butterflies on each of the two levels above
        custum_radix::MultStats st;
        u = v;
        custum_radix::fft_radix2_recursive_mut(u, tw, q, st);
        custum_radix::fft_split_radix_recursive(w = v, tw, q);
        if (u != w || st.nonzero_mults + st.skipped_mults != 12) return 12;
        std::vector<uint32_t> bad(6);
        try {
            custum_radix::fft_radix2_recursive(bad, tw, q);
            return 11;
        } catch (const std::logic_error&) {
        }
    }
    std::puts("cpp example ok");
    return 0;
}
