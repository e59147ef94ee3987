// Per-call latency of the per-polynomial drop-in calls through the C ABI, without any Python layer.
// Build: g++ -O2 -std=c++17 -I include tools/latency_cpp.cpp -o /tmp/latency_cpp -L tfhe-rs-main_modified_b200 -ltfhe_ntt_b200 -Wl,-rpath,$PWD/tfhe-rs-main_modified_b200
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <vector>

#include "tfhe_ntt_b200.h"

template <class F>
double per_call_us(F&& f, int reps = 2000, int warm = 200) {
    for (int i = 0; i < warm; ++i) f();
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < reps; ++i) f();
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double, std::micro>(t1 - t0).count() / reps;
}

int main() {
    const uint64_t p = 0xFFFFFFFF00000001ull;
    for (size_t n : {1024, 2048, 4096}) {
        ntt_b200_plan64* plan = nullptr;
        if (ntt_b200_plan64_try_new(n, p, &plan) != NTT_B200_OK) return 1;
        std::vector<uint64_t> buf(n), acc(n), rhs(n);
        for (size_t i = 0; i < n; ++i) buf[i] = (i * 0x9E3779B97F4A7C15ull) % p, rhs[i] = (i * 12345 + 7) % p;
        double f = per_call_us([&] { ntt_b200_plan64_fwd(plan, buf.data(), n); });
        double v = per_call_us([&] { ntt_b200_plan64_inv(plan, buf.data(), n); });
        double nm = per_call_us([&] { ntt_b200_plan64_normalize(plan, buf.data(), n); });
        double ma = per_call_us([&] { ntt_b200_plan64_mul_accumulate(plan, acc.data(), n, buf.data(), n, rhs.data(), n); });
        printf("C ABI prime64 Solinas n=%zu: fwd %.1f us  inv %.1f us  normalize %.1f us  mul_accumulate %.1f us\n", n, f, v, nm, ma);
        ntt_b200_plan64_free(plan);
    }
    ntt_b200_plan32* p32 = nullptr;
    if (ntt_b200_plan32_try_new(2048, 1073479681u, &p32) != NTT_B200_OK) return 1;
    std::vector<uint32_t> b32(2048);
    for (size_t i = 0; i < 2048; ++i) b32[i] = (uint32_t)((i * 2654435761u) % 1073479681u);
    printf("C ABI prime32 30-bit n=2048: fwd %.1f us  inv %.1f us\n", per_call_us([&] { ntt_b200_plan32_fwd(p32, b32.data(), 2048); }),
           per_call_us([&] { ntt_b200_plan32_inv(p32, b32.data(), 2048); }));
    ntt_b200_plan32_free(p32);
    return 0;
}
