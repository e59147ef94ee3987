// NTT-PBS through the C++ host mirror (tfhe_ntt::ntt64_pbs, mirroring tfhe ntt64_pbs.rs):
//  * with an all-zero bootstrap key every CMUX adds nothing, so the blind rotation is
//    lut / X^switch(body) (ntt64_pbs.rs:247-255) and the PBS output is its sample extraction;
//  * on a random key the fused and the composed device paths must agree bit for bit.
#include <cstdio>
#include <random>
#include <vector>

#include "tfhe_ntt_b200.hpp"

int main() {
    using namespace tfhe_ntt;
    using namespace tfhe_ntt::ntt64_pbs;
    const size_t N = 512, k = 1, gs = k + 1, n_lwe = 5;
    const uint32_t base_log = 12, level = 2;
    const uint64_t p = prime64::SOLINAS_PRIME;
    auto plan = prime64::Plan::try_new(N, p);
    if (!plan) return 2;
    std::mt19937_64 rng(7);
    auto rnd = [&] { return rng() % p; };

    std::vector<uint64_t> zero_key(n_lwe * level * gs * gs * N, 0);
    auto key0 = NttLweBootstrapKey::from_container(*plan, zero_key, n_lwe, gs, base_log, level);
    if (key0.input_lwe_dimension() != n_lwe || key0.polynomial_size() != N || key0.glwe_size() != gs) return 3;
    std::vector<uint64_t> lwe(n_lwe + 1), lut(gs * N), acc;
    for (auto& v : lwe) v = rnd();
    for (auto& v : lut) v = rnd();
    acc = lut;
    blind_rotate_ntt64_assign(lwe, acc, key0);
    // switched body = round(body * 2N / p)   (ntt64_pbs.rs:540-550)
    unsigned __int128 num = (unsigned __int128)lwe[n_lwe] << 10;
    size_t d = (size_t)(num / p) + ((num % p) >= (p >> 1) ? 1 : 0);
    for (size_t c = 0; c < gs; ++c)
        for (size_t j = 0; j < N; ++j) {
            size_t e = j + d;
            uint64_t v = lut[c * N + (e % N)];
            if ((e / N) % 2) v = v == 0 ? 0 : p - v;
            if (acc[c * N + j] != v) return 4;
        }
    std::vector<uint64_t> out(k * N + 1);
    programmable_bootstrap_ntt64_lwe_ciphertext(lwe, out, lut, key0);
    if (out[k * N] != acc[k * N]) return 5;  // body = constant coefficient of the rotated body polynomial
    if (out[0] != acc[0]) return 6;
    for (size_t j = 1; j < N; ++j)
        if (out[j] != (acc[N - j] == 0 ? 0 : p - acc[N - j])) return 7;

    std::vector<uint64_t> rkey(zero_key.size());
    for (auto& v : rkey) v = rnd();
    auto key = NttLweBootstrapKey::from_container(*plan, rkey, n_lwe, gs, base_log, level);
    if (key.as_container() != rkey) return 8;
    std::vector<uint64_t> a = lut, b = lut;
    blind_rotate_ntt64_assign(lwe, a, key, Path::Fused);
    blind_rotate_ntt64_assign(lwe, b, key, Path::Composed);
    if (a != b || a == lut) return 9;
    std::vector<uint64_t> msed(n_lwe + 1);
    for (auto& v : msed) v = rng() % (2 * N);
    a = lut, b = lut;
    blind_rotate_ntt64_bnf_assign(msed, a, key, 64, Path::Fused);
    blind_rotate_ntt64_bnf_assign(msed, b, key, 64, Path::Composed);
    if (a != b) return 10;
    try {  // wrong output size: the reference asserts (glwe_sample_extraction.rs:105-109)
        std::vector<uint64_t> bad(k * N);
        programmable_bootstrap_ntt64_lwe_ciphertext(lwe, bad, lut, key);
        return 11;
    } catch (const std::logic_error&) {
    }
    std::puts("cpp pbs example ok");
    return 0;
}
