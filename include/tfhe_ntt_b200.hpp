// tfhe_ntt_b200.hpp -- header-only C++17 host layer over the C ABI (tfhe_ntt_b200.h).
//
// The reference is compiled code (Rust) and this image has no Rust toolchain, so this is the
// compiled-language mirror of the reference's public API: same module (namespace), type and
// method names, same argument meaning and error behaviour
//   tfhe_ntt::prime64::Plan   <- tfhe-ntt/src/prime64.rs:245-1223
//   tfhe_ntt::prime32::Plan   <- tfhe-ntt/src/prime32.rs:632-1016
//   tfhe_ntt::native64::Plan32 ... <- tfhe-ntt/src/native{32,64,128}.rs, native_binary*.rs
//   tfhe_ntt::fastdiv::{Div32,Div64} <- tfhe-ntt/src/fastdiv.rs:29-150
//   tfhe_ntt::ntt64_pbs::*         <- tfhe/src/core_crypto/algorithms/lwe_programmable_bootstrapping/ntt64_{,bnf_}pbs.rs
// `try_new` returns std::optional (Rust Option); length assertions throw std::logic_error where
// the reference panics; CUDA failures throw std::runtime_error.
#pragma once
#include <array>
#include <cstdint>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "tfhe_ntt_b200.h"

namespace tfhe_ntt {

inline void check(int status, const char* what) {
    if (status == NTT_B200_OK) return;
    if (status == NTT_B200_ERR_LEN)
        throw std::logic_error(std::string("assertion failed: length mismatch in ") + what);
    if (status == NTT_B200_ERR_UNSUPPORTED)
        throw std::length_error(std::string(what) + ": size beyond a capacity limit of this implementation");
    throw std::runtime_error(std::string(what) + ": " + ntt_b200_last_error());
}

namespace prime {
inline bool is_prime64(uint64_t n) { return ntt_b200_is_prime64(n) != 0; }
inline std::optional<uint64_t> largest_prime_in_arithmetic_progression64(uint64_t factor,
                                                                         uint64_t offset,
                                                                         uint64_t lo, uint64_t hi) {
    uint64_t out = 0;
    if (!ntt_b200_largest_prime_in_arithmetic_progression64(factor, offset, lo, hi, &out))
        return std::nullopt;
    return out;
}
}  // namespace prime

namespace fastdiv {
// Exact division by an invariant divisor (the reference uses Lemire's method; only the exact
// quotient/remainder is observable, fastdiv.rs:159-195).
struct Div32 {
    uint32_t d;
    explicit constexpr Div32(uint32_t divisor) : d(divisor) {}
    static constexpr uint32_t div(uint32_t n, Div32 x) { return n / x.d; }
    static constexpr uint32_t rem(uint32_t n, Div32 x) { return n % x.d; }
    static constexpr uint64_t div_u64(uint64_t n, Div32 x) { return n / x.d; }
    static constexpr uint32_t rem_u64(uint64_t n, Div32 x) { return (uint32_t)(n % x.d); }
    constexpr uint32_t divisor() const { return d; }
};
struct Div64 {
    uint64_t d;
    explicit constexpr Div64(uint64_t divisor) : d(divisor) {}
    static constexpr uint64_t div(uint64_t n, Div64 x) { return n / x.d; }
    static constexpr uint64_t rem(uint64_t n, Div64 x) { return n % x.d; }
    static constexpr unsigned __int128 div_u128(unsigned __int128 n, Div64 x) { return n / x.d; }
    static constexpr uint64_t rem_u128(unsigned __int128 n, Div64 x) { return (uint64_t)(n % x.d); }
    constexpr uint64_t divisor() const { return d; }
};
}  // namespace fastdiv

#define TFHE_NTT_PRIME_PLAN(NS, SFX, ELEM)                                                        \
    namespace NS {                                                                                \
    class Plan {                                                                                  \
        struct Del {                                                                              \
            void operator()(ntt_b200_plan##SFX* p) const { ntt_b200_plan##SFX##_free(p); }        \
        };                                                                                        \
        std::unique_ptr<ntt_b200_plan##SFX, Del> h_;                                              \
        explicit Plan(ntt_b200_plan##SFX* h) : h_(h) {}                                           \
                                                                                                  \
       public:                                                                                    \
        static std::optional<Plan> try_new(size_t polynomial_size, ELEM modulus) {                \
            ntt_b200_plan##SFX* h = nullptr;                                                      \
            int st = ntt_b200_plan##SFX##_try_new(polynomial_size, modulus, &h);                  \
            if (st == NTT_B200_NONE) return std::nullopt;                                         \
            check(st, #NS "::Plan::try_new");                                                     \
            return Plan(h);                                                                       \
        }                                                                                         \
        Plan clone() const {                                                                      \
            ntt_b200_plan##SFX* h = nullptr;                                                      \
            check(ntt_b200_plan##SFX##_clone(h_.get(), &h), "clone");                             \
            return Plan(h);                                                                       \
        }                                                                                         \
        /* an owning handle on a plan borrowed from another object (shares its device tables) */ \
        static Plan clone_of(const ntt_b200_plan##SFX* borrowed) {                                \
            ntt_b200_plan##SFX* h = nullptr;                                                      \
            check(ntt_b200_plan##SFX##_clone(borrowed, &h), "clone");                             \
            return Plan(h);                                                                       \
        }                                                                                         \
        size_t ntt_size() const { return ntt_b200_plan##SFX##_ntt_size(h_.get()); }               \
        ELEM modulus() const { return ntt_b200_plan##SFX##_modulus(h_.get()); }                   \
        bool can_use_fast_reduction_code() const {                                                \
            return ntt_b200_plan##SFX##_can_use_fast_reduction_code(h_.get()) != 0;               \
        }                                                                                         \
        int device() const { return ntt_b200_plan##SFX##_device(h_.get()); }                      \
        const ntt_b200_plan##SFX* raw() const { return h_.get(); }                                \
        void fwd(ELEM* buf, size_t len) const {                                                   \
            check(ntt_b200_plan##SFX##_fwd(h_.get(), buf, len), #NS "::Plan::fwd");               \
        }                                                                                         \
        void inv(ELEM* buf, size_t len) const {                                                   \
            check(ntt_b200_plan##SFX##_inv(h_.get(), buf, len), #NS "::Plan::inv");               \
        }                                                                                         \
        void fwd(std::vector<ELEM>& buf) const { fwd(buf.data(), buf.size()); }                   \
        void inv(std::vector<ELEM>& buf) const { inv(buf.data(), buf.size()); }                   \
        void normalize(ELEM* values, size_t len) const {                                          \
            check(ntt_b200_plan##SFX##_normalize(h_.get(), values, len), "normalize");            \
        }                                                                                         \
        void mul_assign_normalize(ELEM* lhs, size_t lhs_len, const ELEM* rhs,                     \
                                  size_t rhs_len) const {                                         \
            check(ntt_b200_plan##SFX##_mul_assign_normalize(h_.get(), lhs, lhs_len, rhs, rhs_len),\
                  "mul_assign_normalize");                                                        \
        }                                                                                         \
        void mul_accumulate(ELEM* acc, size_t acc_len, const ELEM* lhs, size_t lhs_len,           \
                            const ELEM* rhs, size_t rhs_len) const {                              \
            check(ntt_b200_plan##SFX##_mul_accumulate(h_.get(), acc, acc_len, lhs, lhs_len, rhs,  \
                                                      rhs_len),                                   \
                  "mul_accumulate");                                                              \
        }                                                                                         \
        /* new: batched host / device-resident */                                                 \
        void fwd_batch(ELEM* host, size_t batch) const {                                          \
            check(ntt_b200_plan##SFX##_fwd_batch(h_.get(), host, batch), "fwd_batch");            \
        }                                                                                         \
        void inv_batch(ELEM* host, size_t batch) const {                                          \
            check(ntt_b200_plan##SFX##_inv_batch(h_.get(), host, batch), "inv_batch");            \
        }                                                                                         \
        /* one host batch over several GPUs: plans[g] lives on GPU g (see the C header) */        \
        static void fwd_batch_multi_gpu(const std::vector<const Plan*>& plans, ELEM* host,        \
                                        size_t batch) {                                           \
            std::vector<const ntt_b200_plan##SFX*> raw;                                           \
            for (const Plan* p : plans) raw.push_back(p->h_.get());                               \
            check(ntt_b200_plan##SFX##_fwd_batch_multi_gpu(raw.data(), raw.size(), host, batch),  \
                  "fwd_batch_multi_gpu");                                                         \
        }                                                                                         \
        static void inv_batch_multi_gpu(const std::vector<const Plan*>& plans, ELEM* host,        \
                                        size_t batch) {                                           \
            std::vector<const ntt_b200_plan##SFX*> raw;                                           \
            for (const Plan* p : plans) raw.push_back(p->h_.get());                               \
            check(ntt_b200_plan##SFX##_inv_batch_multi_gpu(raw.data(), raw.size(), host, batch),  \
                  "inv_batch_multi_gpu");                                                         \
        }                                                                                         \
        void fwd_device(ELEM* dev, size_t batch, void* stream = nullptr) const {                  \
            check(ntt_b200_plan##SFX##_fwd_device(h_.get(), dev, batch, stream), "fwd_device");   \
        }                                                                                         \
        void inv_device(ELEM* dev, size_t batch, void* stream = nullptr) const {                  \
            check(ntt_b200_plan##SFX##_inv_device(h_.get(), dev, batch, stream), "inv_device");   \
        }                                                                                         \
        void normalize_device(ELEM* dev, size_t len, void* stream = nullptr) const {              \
            check(ntt_b200_plan##SFX##_normalize_device(h_.get(), dev, len, stream),              \
                  "normalize_device");                                                            \
        }                                                                                         \
        void mul_accumulate_device(ELEM* acc, size_t len, const ELEM* lhs, size_t lhs_len,        \
                                   const ELEM* rhs, size_t rhs_len, void* stream = nullptr) const {\
            check(ntt_b200_plan##SFX##_mul_accumulate_device(h_.get(), acc, len, lhs, lhs_len,    \
                                                             rhs, rhs_len, stream),               \
                  "mul_accumulate_device");                                                       \
        }                                                                                         \
        void ext_product_device(ELEM* out, const ELEM* in, const ELEM* ggsw, size_t rows,         \
                                size_t cols, size_t batch, void* stream = nullptr) const {        \
            check(ntt_b200_plan##SFX##_ext_product_device(h_.get(), out, in, ggsw, rows, cols,    \
                                                          batch, stream),                         \
                  "ext_product_device");                                                          \
        }                                                                                         \
        void fwd_mac_inv_device(ELEM* out, const ELEM* lhs, const ELEM* rhs, size_t rhs_polys,    \
                                const ELEM* acc, size_t acc_polys, size_t batch,                  \
                                void* stream = nullptr) const {                                   \
            check(ntt_b200_plan##SFX##_fwd_mac_inv_device(h_.get(), out, lhs, rhs, rhs_polys, acc,\
                                                          acc_polys, batch, stream),              \
                  "fwd_mac_inv_device");                                                          \
        }                                                                                         \
    };                                                                                            \
    }

TFHE_NTT_PRIME_PLAN(prime64, 64, uint64_t)
TFHE_NTT_PRIME_PLAN(prime32, 32, uint32_t)
#undef TFHE_NTT_PRIME_PLAN

namespace prime64 {
constexpr uint64_t SOLINAS_PRIME = 0xFFFFFFFF00000001ull;  // prime64.rs:8
struct Solinas {
    static constexpr uint64_t P = SOLINAS_PRIME;  // generic_solinas.rs:38-40
};
}  // namespace prime64

// CRT plans: KIND selects the reference type, V the value word, R the residue word, NP the
// number of primes (= number of mod_p* buffers of fwd/inv).
template <int KIND, class V, class R, int NP, bool BINARY>
class NativePlan {
    struct Del {
        void operator()(ntt_b200_native_plan* p) const { ntt_b200_native_free(p); }
    };
    std::unique_ptr<ntt_b200_native_plan, Del> h_;
    explicit NativePlan(ntt_b200_native_plan* h) : h_(h) {}

   public:
    using value_type = V;
    using residue_type = R;
    static constexpr int num_primes = NP;
    static constexpr bool is_binary = BINARY;
    static std::optional<NativePlan> try_new(size_t n) {
        ntt_b200_native_plan* h = nullptr;
        int st = ntt_b200_native_try_new(KIND, n, &h);
        if (st == NTT_B200_NONE) return std::nullopt;
        check(st, "try_new");
        return NativePlan(h);
    }
    size_t ntt_size() const { return ntt_b200_native_ntt_size(h_.get()); }
    // ntt_0() .. ntt_9(): the per-prime plans (native64.rs:950-968, :1095-1103).  The reference returns `&Plan`; here a
    // cheap clone that shares the device tables of the CRT plan.
    using prime_plan = std::conditional_t<sizeof(R) == 4, prime32::Plan, prime64::Plan>;
    prime_plan ntt_i(int i) const {
        const void* b = (i >= 0 && i < NP) ? ntt_b200_native_ntt_i(h_.get(), i) : nullptr;
        if (!b) throw std::out_of_range("ntt_i: this plan has no such prime");
        if constexpr (sizeof(R) == 4)
            return prime32::Plan::clone_of(static_cast<const ntt_b200_plan32*>(b));
        else
            return prime64::Plan::clone_of(static_cast<const ntt_b200_plan64*>(b));
    }
    prime_plan ntt_0() const { return ntt_i(0); }
    prime_plan ntt_1() const { return ntt_i(1); }
    prime_plan ntt_2() const { return ntt_i(2); }
    prime_plan ntt_3() const { return ntt_i(3); }
    prime_plan ntt_4() const { return ntt_i(4); }
    prime_plan ntt_5() const { return ntt_i(5); }
    prime_plan ntt_6() const { return ntt_i(6); }
    prime_plan ntt_7() const { return ntt_i(7); }
    prime_plan ntt_8() const { return ntt_i(8); }
    prime_plan ntt_9() const { return ntt_i(9); }
    // fwd(value, mod_p0, .., mod_p{NP-1})   e.g. native64.rs:970
    void fwd(const V* value, size_t len, const std::array<R*, NP>& mod_p) const {
        void* r[NP];
        for (int i = 0; i < NP; ++i) r[i] = mod_p[i];
        check(ntt_b200_native_fwd(h_.get(), value, len, r, 0), "fwd");
    }
    void fwd_binary(const V* value, size_t len, const std::array<R*, NP>& mod_p) const {
        static_assert(BINARY, "fwd_binary exists on the native_binary plans only");
        void* r[NP];
        for (int i = 0; i < NP; ++i) r[i] = mod_p[i];
        check(ntt_b200_native_fwd(h_.get(), value, len, r, 1), "fwd_binary");
    }
    // inv(value, mod_p0, ..) -- clobbers the residue buffers like the reference (native64.rs:1009)
    void inv(V* value, size_t len, const std::array<R*, NP>& mod_p) const {
        void* r[NP];
        for (int i = 0; i < NP; ++i) r[i] = mod_p[i];
        check(ntt_b200_native_inv(h_.get(), value, len, r), "inv");
    }
    void negacyclic_polymul(V* prod, size_t prod_len, const V* lhs, size_t lhs_len, const V* rhs,
                            size_t rhs_len) const {
        check(ntt_b200_native_negacyclic_polymul(h_.get(), prod, prod_len, lhs, lhs_len, rhs, rhs_len),
              "negacyclic_polymul");
    }
    void negacyclic_polymul_batch(V* prod, const V* lhs, const V* rhs, size_t batch) const {
        check(ntt_b200_native_negacyclic_polymul_batch(h_.get(), prod, lhs, rhs, batch),
              "negacyclic_polymul_batch");
    }
    void negacyclic_polymul_device(V* prod, const V* lhs, const V* rhs, size_t batch,
                                   void* stream = nullptr) const {
        check(ntt_b200_native_negacyclic_polymul_device(h_.get(), prod, lhs, rhs, batch, stream),
              "negacyclic_polymul_device");
    }
};

using u128 = unsigned __int128;
namespace native32 {
using Plan32 = NativePlan<NTT_B200_NATIVE32_PLAN32, uint32_t, uint32_t, 3, false>;
using Plan52 = NativePlan<NTT_B200_NATIVE32_PLAN52, uint32_t, uint64_t, 2, false>;
}  // namespace native32
namespace native64 {
using Plan32 = NativePlan<NTT_B200_NATIVE64_PLAN32, uint64_t, uint32_t, 5, false>;
using Plan52 = NativePlan<NTT_B200_NATIVE64_PLAN52, uint64_t, uint64_t, 3, false>;
}  // namespace native64
namespace native128 {
using Plan32 = NativePlan<NTT_B200_NATIVE128_PLAN32, u128, uint32_t, 10, false>;
}
namespace native_binary32 {
using Plan32 = NativePlan<NTT_B200_NATIVE_BINARY32_PLAN32, uint32_t, uint32_t, 2, true>;
using Plan52 = NativePlan<NTT_B200_NATIVE_BINARY32_PLAN52, uint32_t, uint64_t, 1, true>;
}  // namespace native_binary32
namespace native_binary64 {
using Plan32 = NativePlan<NTT_B200_NATIVE_BINARY64_PLAN32, uint64_t, uint32_t, 3, true>;
using Plan52 = NativePlan<NTT_B200_NATIVE_BINARY64_PLAN52, uint64_t, uint64_t, 2, true>;
}  // namespace native_binary64
namespace native_binary128 {
using Plan32 = NativePlan<NTT_B200_NATIVE_BINARY128_PLAN32, u128, uint32_t, 5, true>;
}

// NTT programmable bootstrap on top of prime64::Plan (the caller of the hot path):
//   tfhe::core_crypto::algorithms::lwe_programmable_bootstrapping::ntt64_pbs / ntt64_bnf_pbs,
//   tfhe::core_crypto::entities::NttLweBootstrapKey, convert_standard_lwe_bootstrap_key_to_ntt64.
// Containers are the reference's flat ones; every call takes `batch` ciphertexts.
namespace ntt64_pbs {
enum class NttLweBootstrapKeyOption { Raw = 0, Normalize = 1 };  // lwe_bootstrap_key_conversion.rs:283-288
enum class Path { Auto = 0, Fused = 1, Composed = 2, Cluster = 3 };

class NttLweBootstrapKey {
    struct Del {
        void operator()(ntt_b200_bsk* p) const { ntt_b200_bsk_free(p); }
    };
    std::unique_ptr<ntt_b200_bsk, Del> h_;
    explicit NttLweBootstrapKey(ntt_b200_bsk* h) : h_(h) {}

   public:
    // NttLweBootstrapKey::from_container, entities/ntt_lwe_bootstrap_key.rs:68-110
    static NttLweBootstrapKey from_container(const prime64::Plan& plan, const std::vector<uint64_t>& container,
                                             size_t input_lwe_dimension, size_t glwe_size,
                                             uint32_t decomposition_base_log, uint32_t decomposition_level_count) {
        if (container.size() != input_lwe_dimension * decomposition_level_count * glwe_size * glwe_size * plan.ntt_size())
            throw std::logic_error("assertion failed: NttLweBootstrapKey container length");
        ntt_b200_bsk* h = nullptr;
        check(ntt_b200_bsk_new(plan.raw(), container.data(), input_lwe_dimension, glwe_size, decomposition_base_log,
                               decomposition_level_count, &h),
              "NttLweBootstrapKey::from_container");
        return NttLweBootstrapKey(h);
    }
    // convert_standard_lwe_bootstrap_key_to_ntt64 (:294-363) straight into device memory
    static NttLweBootstrapKey from_standard(const prime64::Plan& plan, const std::vector<uint64_t>& standard_bsk,
                                            size_t input_lwe_dimension, size_t glwe_size,
                                            uint32_t decomposition_base_log, uint32_t decomposition_level_count,
                                            uint32_t input_modulus_width, NttLweBootstrapKeyOption option) {
        if (standard_bsk.size() !=
            input_lwe_dimension * decomposition_level_count * glwe_size * glwe_size * plan.ntt_size())
            throw std::logic_error("assertion failed: LweBootstrapKey container length");
        ntt_b200_bsk* h = nullptr;
        check(ntt_b200_bsk_convert_new(plan.raw(), standard_bsk.data(), input_lwe_dimension, glwe_size,
                                       decomposition_base_log, decomposition_level_count, input_modulus_width,
                                       (int)option, &h),
              "NttLweBootstrapKey::from_standard");
        return NttLweBootstrapKey(h);
    }
    size_t input_lwe_dimension() const { return ntt_b200_bsk_input_lwe_dimension(h_.get()); }
    size_t glwe_size() const { return ntt_b200_bsk_glwe_size(h_.get()); }
    size_t polynomial_size() const { return ntt_b200_bsk_polynomial_size(h_.get()); }
    uint32_t decomposition_base_log() const { return ntt_b200_bsk_decomposition_base_log(h_.get()); }
    uint32_t decomposition_level_count() const { return ntt_b200_bsk_decomposition_level_count(h_.get()); }
    size_t output_lwe_dimension() const { return (glwe_size() - 1) * polynomial_size(); }
    std::vector<uint64_t> as_container() const {
        std::vector<uint64_t> out(input_lwe_dimension() * decomposition_level_count() * glwe_size() * glwe_size() *
                                  polynomial_size());
        check(ntt_b200_bsk_read(h_.get(), out.data(), out.size()), "NttLweBootstrapKey::as_container");
        return out;
    }
    const ntt_b200_bsk* raw() const { return h_.get(); }
};

// lwe_bootstrap_key_conversion.rs:294-363, host to host
inline void convert_standard_lwe_bootstrap_key_to_ntt64(const prime64::Plan& plan,
                                                        const std::vector<uint64_t>& input_bsk,
                                                        std::vector<uint64_t>& output_bsk,
                                                        NttLweBootstrapKeyOption option,
                                                        uint32_t input_modulus_width = 0) {
    if (input_bsk.size() != output_bsk.size()) throw std::logic_error("assertion failed: mismatched key sizes");
    check(ntt_b200_convert_standard_lwe_bootstrap_key_to_ntt64(plan.raw(), input_bsk.data(), output_bsk.data(),
                                                               input_bsk.size(), input_modulus_width, (int)option),
          "convert_standard_lwe_bootstrap_key_to_ntt64");
}

inline size_t batch_of(size_t len, size_t row, const char* what) {
    if (row == 0 || len % row) throw std::logic_error(std::string("assertion failed: container length in ") + what);
    return len / row;
}
// blind_rotate_ntt64_assign, ntt64_pbs.rs:175-286
inline void blind_rotate_ntt64_assign(const std::vector<uint64_t>& input, std::vector<uint64_t>& lut,
                                      const NttLweBootstrapKey& bsk, Path path = Path::Auto) {
    size_t batch = batch_of(input.size(), bsk.input_lwe_dimension() + 1, "blind_rotate_ntt64_assign");
    if (lut.size() != batch * bsk.glwe_size() * bsk.polynomial_size())
        throw std::logic_error("assertion failed: lut size");
    check(ntt_b200_blind_rotate_ntt64_assign(bsk.raw(), input.data(), lut.data(), batch, (int)path),
          "blind_rotate_ntt64_assign");
}
// blind_rotate_ntt64_bnf_assign, ntt64_bnf_pbs.rs:174-276
inline void blind_rotate_ntt64_bnf_assign(const std::vector<uint64_t>& msed_input, std::vector<uint64_t>& lut,
                                          const NttLweBootstrapKey& bsk, uint32_t ciphertext_modulus_width = 64,
                                          Path path = Path::Auto) {
    size_t batch = batch_of(msed_input.size(), bsk.input_lwe_dimension() + 1, "blind_rotate_ntt64_bnf_assign");
    if (lut.size() != batch * bsk.glwe_size() * bsk.polynomial_size())
        throw std::logic_error("assertion failed: lut size");
    check(ntt_b200_blind_rotate_ntt64_bnf_assign(bsk.raw(), ciphertext_modulus_width, msed_input.data(), lut.data(),
                                                 batch, (int)path),
          "blind_rotate_ntt64_bnf_assign");
}
// programmable_bootstrap_ntt64_lwe_ciphertext, ntt64_pbs.rs:439-538
inline void programmable_bootstrap_ntt64_lwe_ciphertext(const std::vector<uint64_t>& input,
                                                        std::vector<uint64_t>& output,
                                                        const std::vector<uint64_t>& accumulator,
                                                        const NttLweBootstrapKey& bsk, Path path = Path::Auto) {
    size_t batch = batch_of(input.size(), bsk.input_lwe_dimension() + 1, "programmable_bootstrap_ntt64");
    if (output.size() != batch * (bsk.output_lwe_dimension() + 1))
        throw std::logic_error("assertion failed: output size");
    size_t acc_count = batch_of(accumulator.size(), bsk.glwe_size() * bsk.polynomial_size(), "accumulator");
    check(ntt_b200_programmable_bootstrap_ntt64(bsk.raw(), input.data(), output.data(), accumulator.data(), acc_count,
                                                batch, (int)path),
          "programmable_bootstrap_ntt64_lwe_ciphertext");
}
// programmable_bootstrap_ntt64_bnf_lwe_ciphertext, ntt64_bnf_pbs.rs:428-539
inline void programmable_bootstrap_ntt64_bnf_lwe_ciphertext(const std::vector<uint64_t>& input,
                                                            std::vector<uint64_t>& output,
                                                            const std::vector<uint64_t>& accumulator,
                                                            const NttLweBootstrapKey& bsk,
                                                            uint32_t ciphertext_modulus_width = 64,
                                                            Path path = Path::Auto) {
    size_t batch = batch_of(input.size(), bsk.input_lwe_dimension() + 1, "programmable_bootstrap_ntt64_bnf");
    if (output.size() != batch * (bsk.output_lwe_dimension() + 1))
        throw std::logic_error("assertion failed: output size");
    size_t acc_count = batch_of(accumulator.size(), bsk.glwe_size() * bsk.polynomial_size(), "accumulator");
    check(ntt_b200_programmable_bootstrap_ntt64_bnf(bsk.raw(), ciphertext_modulus_width, input.data(), output.data(),
                                                    accumulator.data(), acc_count, batch, (int)path),
          "programmable_bootstrap_ntt64_bnf_lwe_ciphertext");
}
}  // namespace ntt64_pbs

// custum_radix (tfhe-ntt/src/custum_radix/mod.rs:1-22): the fork's recursive cyclic u32 transforms over a
// caller-built table twiddles[k] = root^k mod p, natural order in and out; the `_mut` routines of fwd_1.rs also
// return the fork's MultStats counters.
namespace custum_radix {
inline void fft_radix2_recursive(std::vector<uint32_t>& a, const std::vector<uint32_t>& twiddles, uint32_t p) {  // fwd.rs:170
    check(ntt_b200_custum_radix_fft(NTT_B200_CR_RADIX2, a.data(), a.size(), twiddles.data(), twiddles.size(), p),
          "fft_radix2_recursive");
}
inline void fft_radix4_recursive(std::vector<uint32_t>& a, const std::vector<uint32_t>& twiddles, uint32_t p) {  // fwd.rs:105
    check(ntt_b200_custum_radix_fft(NTT_B200_CR_RADIX4, a.data(), a.size(), twiddles.data(), twiddles.size(), p),
          "fft_radix4_recursive");
}
inline void fft_split_radix_recursive(std::vector<uint32_t>& a, const std::vector<uint32_t>& tw, uint32_t p) {  // fwd.rs:207
    check(ntt_b200_custum_radix_fft(NTT_B200_CR_SPLIT_RADIX, a.data(), a.size(), tw.data(), tw.size(), p),
          "fft_split_radix_recursive");
}
inline void ifft_radix2_recursive(std::vector<uint32_t>& a, const std::vector<uint32_t>& inv_twiddles, uint32_t p,
                                  uint32_t n_inv, bool top) {  // inv.rs:178
    check(ntt_b200_custum_radix_ifft(NTT_B200_CR_RADIX2, a.data(), a.size(), inv_twiddles.data(), inv_twiddles.size(), p,
                                     n_inv, top),
          "ifft_radix2_recursive");
}
inline void ifft_radix4_recursive(std::vector<uint32_t>& a, const std::vector<uint32_t>& inv_twiddles, uint32_t p,
                                  uint32_t n_inv, bool top) {  // inv.rs:106 (halves when log2 n is odd)
    check(ntt_b200_custum_radix_ifft(NTT_B200_CR_RADIX4, a.data(), a.size(), inv_twiddles.data(), inv_twiddles.size(), p,
                                     n_inv, top),
          "ifft_radix4_recursive");
}
inline void ifft_split_radix_recursive(std::vector<uint32_t>& a, const std::vector<uint32_t>& inv_tw, uint32_t p,
                                       uint32_t n_inv, bool top) {  // inv.rs:232
    check(ntt_b200_custum_radix_ifft(NTT_B200_CR_SPLIT_RADIX, a.data(), a.size(), inv_tw.data(), inv_tw.size(), p, n_inv,
                                     top),
          "ifft_split_radix_recursive");
}
// fwd_1.rs:3-7
struct MultStats {
    size_t nonzero_mults = 0;  // nonzero * nonzero
    size_t skipped_mults = 0;  // multiplications with zero
};
namespace detail {
inline void fft_mut(int kind, std::vector<uint32_t>& a, const std::vector<uint32_t>& tw, uint32_t p, MultStats& st,
                    const char* what) {
    uint64_t raw[2] = {st.nonzero_mults, st.skipped_mults};
    check(ntt_b200_custum_radix_fft_mut(kind, a.data(), a.size(), tw.data(), tw.size(), p, raw), what);
    st.nonzero_mults = (size_t)raw[0];
    st.skipped_mults = (size_t)raw[1];
}
}  // namespace detail
// fwd_1.rs:102 / :190 / :232 -- values and counters (one vector per call, n <= 4096)
inline void fft_radix4_recursive_mut(std::vector<uint32_t>& a, const std::vector<uint32_t>& tw, uint32_t p, MultStats& st) {
    detail::fft_mut(NTT_B200_CR_RADIX4, a, tw, p, st, "fft_radix4_recursive_mut");
}
inline void fft_radix2_recursive_mut(std::vector<uint32_t>& a, const std::vector<uint32_t>& tw, uint32_t p, MultStats& st) {
    detail::fft_mut(NTT_B200_CR_RADIX2, a, tw, p, st, "fft_radix2_recursive_mut");
}
inline void fft_split_radix_recursive_mut(std::vector<uint32_t>& a, const std::vector<uint32_t>& tw, uint32_t p,
                                          MultStats& st) {
    detail::fft_mut(NTT_B200_CR_SPLIT_RADIX, a, tw, p, st, "fft_split_radix_recursive_mut");
}
inline void ifft_radix4_recursive_mut(std::vector<uint32_t>& a, const std::vector<uint32_t>& inv_twiddles, uint32_t p,
                                      uint32_t n_inv, bool top, MultStats& st) {  // fwd_1.rs:296
    uint64_t raw[2] = {st.nonzero_mults, st.skipped_mults};
    check(ntt_b200_custum_radix_ifft_radix4_mut(a.data(), a.size(), inv_twiddles.data(), inv_twiddles.size(), p, n_inv,
                                                top, raw),
          "ifft_radix4_recursive_mut");
    st.nonzero_mults = (size_t)raw[0];
    st.skipped_mults = (size_t)raw[1];
}
}  // namespace custum_radix

}  // namespace tfhe_ntt
