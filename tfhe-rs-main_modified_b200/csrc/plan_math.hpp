// Host-side plan construction for the B200 NTT engine.
//
// Produces the same roots of unity, twiddle values and pointwise constants as the
// reference crate (tfhe-ntt), so that NTT-domain outputs are bit-identical:
//   acceptance rules      tfhe-ntt/src/prime64.rs:769-774, prime32.rs:667-672
//   primality             tfhe-ntt/src/prime.rs:76-126
//   2n-th root of unity   tfhe-ntt/src/roots.rs:17-107, prime64.rs:162-182
//   twiddle placement     tfhe-ntt/src/prime64.rs:184-203, prime32.rs:226-245
//   Barrett thresholds    tfhe-ntt/src/prime64.rs:726-759, :815-817 ; prime32.rs:600-628, :748-749
// The device-side table layout is ours (see ntt_engine.cu); only the values are shared.
#pragma once
#include <cstddef>
#include <cstdint>
#include <optional>
#include <vector>

namespace nttb200 {
namespace pm {

using u32 = uint32_t;
using u64 = uint64_t;
using u128 = unsigned __int128;

constexpr u64 kSolinas = 0xFFFFFFFF00000001ull;  // 2^64 - 2^32 + 1

inline u64 mulmod(u64 a, u64 b, u64 p) { return (u64)((u128)a * b % p); }

// square-and-multiply, LSB first; any exact powmod gives the same value
inline u64 powmod(u64 base, u64 e, u64 p) {
    u64 acc = 1 % p;
    base %= p;
    for (; e; e >>= 1) {
        if (e & 1) acc = mulmod(acc, base, p);
        base = mulmod(base, base, p);
    }
    return acc;
}

inline bool is_prime(u64 n) {
    static const u64 bases[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    if (n < 2) return false;
    for (u64 q : bases)
        if (n % q == 0) return n == q;
    int s = __builtin_ctzll(n - 1);
    u64 d = (n - 1) >> s;
    for (u64 a : bases) {  // deterministic for n < 2^64
        u64 x = powmod(a, d, n);
        if (x == 1 || x == n - 1) continue;
        bool witness = true;
        for (int r = 1; r < s && witness; ++r) {
            x = mulmod(x, x, n);
            if (x == n - 1) witness = false;
        }
        if (witness) return false;
    }
    return true;
}

// Largest prime factor*x + offset in [lo, hi]  (reference: prime.rs:130-186; a public helper
// of the crate, used by its tests and benches to pick NTT primes).
inline std::optional<u64> largest_prime_in_arithmetic_progression64(u64 factor, u64 offset, u64 lo,
                                                                    u64 hi) {
    if (lo > hi || offset > hi) return std::nullopt;
    if (factor == 0) {
        if (lo <= offset && is_prime(offset)) return offset;
        return std::nullopt;
    }
    u64 start = lo > offset ? lo - offset : 0;
    u64 x_lo = start / factor + (start % factor != 0);
    u64 x = (hi - offset) / factor;
    if (x < x_lo) return std::nullopt;
    for (;; --x) {
        u64 v = factor * x + offset;
        if (is_prime(v)) return v;
        if (x == x_lo) break;
    }
    return std::nullopt;
}

// Tonelli-Shanks square root exactly as the reference walks it (roots.rs:31-66): the branch
// of the square root that comes out decides which primitive root the plan uses.
inline std::optional<u64> tonelli_shanks(u64 p, u64 q, u64 s, u64 z, u64 x) {
    u64 m = s, c = powmod(z, q, p), t = powmod(x, q, p), r = powmod(x, (q + 1) / 2, p);
    while (true) {
        if (t == 0) return 0;
        if (t == 1) return r;
        u64 i = 0, tp = t;
        while (i < m) {
            tp = mulmod(tp, tp, p);
            ++i;
            if (tp == 1) break;
        }
        if (i == m) return std::nullopt;
        u64 b = powmod(c, u64(1) << (m - i - 1), p);
        m = i;
        c = mulmod(b, b, p);
        t = mulmod(t, c, p);
        r = mulmod(r, b, p);
    }
}

// roots.rs:68-91: start from -1 and take log2(degree)-1 successive square roots.
inline std::optional<u64> primitive_root(u64 p, u64 degree) {
    if (degree < 2 || (degree & (degree - 1))) return std::nullopt;
    int s = __builtin_ctzll(p - 1);
    u64 q = (p - 1) >> s;
    u64 z = 2;  // smallest quadratic non-residue (roots.rs:17-28)
    while (z < p && powmod(z, (p - 1) / 2, p) != p - 1) ++z;
    if (z >= p) return std::nullopt;
    u64 root = p - 1;
    for (int i = __builtin_ctzll(degree); i > 1; --i) {
        auto r = tonelli_shanks(p, q, (u64)s, z, root);
        if (!r) return std::nullopt;
        root = *r;
    }
    return root;
}

// roots.rs:96-107 (covers the literal table prime64.rs:167-177, see roots.rs:150-172)
inline u64 solinas_root(u64 degree) {
    return powmod(16334397945464290598ull, (u64(1) << 32) / degree, kSolinas);
}

inline size_t bit_reverse(unsigned nbits, size_t i) {
    size_t r = 0;
    for (unsigned b = 0; b < nbits; ++b, i >>= 1) r = (r << 1) | (i & 1);
    return r;
}

struct BarrettInfo {
    u64 big_q = 0;
    u64 p_barrett = 0;
    bool single_step = false;
};
// bits = 32 or 64 (prime32.rs:606-627, prime64.rs:733-758)
inline BarrettInfo barrett_info(u64 p, unsigned bits) {
    BarrettInfo b;
    unsigned big_q = 64 - __builtin_clzll(p);
    unsigned big_l = big_q + bits - 1;
    u128 two_l = (u128)1 << big_l;
    b.big_q = big_q;
    b.p_barrett = (u64)(two_l / p);
    b.single_step = (two_l % p) <= (u128)p - ((u128)1 << (big_q - 1));
    return b;
}

// Natural-order psi powers placed like the reference places them:
//   fwd[bit_rev(k)] = psi^k ; inv[bit_rev((n-k)%n)] = (k==0 ? 1 : p - psi^k)
struct Twiddles {
    std::vector<u64> fwd, inv;
    u64 psi = 0;
};

// Returns nullopt exactly when the reference's try_new returns None.
// min_n = 16 for prime64, 32 for prime32.
inline std::optional<Twiddles> build_twiddles(size_t n, u64 p, size_t min_n) {
    if (n < min_n || (n & (n - 1)) || !is_prime(p)) return std::nullopt;
    auto generic_root = primitive_root(p, 2 * (u64)n);  // existence check always uses this one
    if (!generic_root) return std::nullopt;
    Twiddles t;
    t.psi = (p == kSolinas) ? solinas_root(2 * (u64)n) : *generic_root;
    t.fwd.resize(n);
    t.inv.resize(n);
    unsigned nbits = __builtin_ctzll((u64)n);
    u64 wk = 1;
    for (size_t k = 0; k < n; ++k) {
        t.fwd[bit_reverse(nbits, k)] = wk;
        t.inv[bit_reverse(nbits, (n - k) % n)] = k == 0 ? wk : p - wk;
        wk = mulmod(wk, t.psi, p);
    }
    return t;
}

// ---- CRT primes (tfhe-ntt/src/lib.rs:457-466, :605-610) ----
constexpr u32 kPrimes32[10] = {0x3F5A0001u, 0x3F5D0001u, 0x3F760001u, 0x3F820001u, 0x3FAC0001u,
                               0x3FAF0001u, 0x3FB10001u, 0x3FBB0001u, 0x3FDE0001u, 0x3FFC0001u};
constexpr u64 kPrimes52[6] = {0x3FFFFFE770001ull, 0x3FFFFFEB90001ull, 0x3FFFFFEC80001ull,
                              0x3FFFFFF8B0001ull, 0x3FFFFFFB80001ull, 0x3FFFFFFC70001ull};

inline u64 inv_mod_prime(u64 x, u64 p) { return powmod(x, p - 2, p); }
inline u64 shoup64(u64 w, u64 p) { return (u64)(((u128)w << 64) / p); }
inline u32 shoup32(u32 w, u32 p) { return (u32)(((u64)w << 32) / p); }

}  // namespace pm
}  // namespace nttb200
