// Host-side engine: plan objects that own device tables and launch the kernels.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "ntt_arith.cuh"

namespace nttb200 {

struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define NTT_CUDA_CHECK(expr)                                                                  \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess)                                                               \
            throw ::nttb200::CudaError(std::string(#expr) + ": " + cudaGetErrorString(e__));  \
    } while (0)

// Raw device tables of a 30-bit-prime plan, for kernels that run several plans at once (the fused
// CRT polymul).
struct RawShoup32H {
    const ShoupTw<uint32_t>* fwd;
    const ShoupTw<uint32_t>* inv;
    Shoup<uint32_t, true>::Ctx ctx;
    ShoupTw<uint32_t> n_inv;
};

// Abstract prime plan (prime32::Plan / prime64::Plan); one concrete class per modulus family.
struct PrimePlan {
    size_t n = 0;
    int logn = 0;
    int elem_bytes = 0;  // 4 or 8
    uint64_t p = 0;
    bool can_use_fast_reduction_code = false;
    int device = 0;
    const char* family = "";

    virtual ~PrimePlan() = default;
    // all pointers are device pointers on `device`; asynchronous on `stream`
    virtual void fwd(void* data, size_t batch, cudaStream_t stream) const = 0;
    virtual void inv(void* data, size_t batch, cudaStream_t stream) const = 0;
    virtual void normalize(void* v, size_t total, cudaStream_t stream) const = 0;
    virtual void mul_assign_normalize(void* lhs, const void* rhs, size_t total, size_t rhs_period,
                                      cudaStream_t stream) const = 0;
    virtual void mul_accumulate(void* acc, const void* lhs, const void* rhs, size_t total,
                                size_t lhs_period, size_t rhs_period,
                                cudaStream_t stream) const = 0;
    // out[b] = inv(acc[b % acc_polys] + fwd(lhs[b]) * rhs[b % rhs_polys]); acc may be null
    virtual void fwd_mac_inv(void* out, const void* lhs, const void* rhs, size_t rhs_polys,
                             const void* acc, size_t acc_polys, size_t batch,
                             cudaStream_t stream) const = 0;
    // out[b][c] = inv(sum_r fwd(in[b][r]) * ggsw[r][c]); in: [batch][rows][n], ggsw: [rows][cols][n]
    // (NTT domain, shared by the batch), out: [batch][cols][n]
    virtual void ext_product(void* out, const void* in, const void* ggsw, unsigned rows,
                             unsigned cols, size_t batch, cudaStream_t stream) const = 0;
    // Fused blind rotation of the NTT-PBS (u64 plans, tfhe ntt64_pbs.rs:213-286 / ntt64_bnf_pbs.rs:
    // 208-276): acc_out[b] = blind rotation of lut[b % lut_count] by the switched ciphertext
    // switched[b][n_lwe+1] under the NTT-domain key bsk.  Returns false when this plan / shape has
    // no fused kernel (the caller then composes ext_product with elementwise kernels).
    // bsk_tw: the key in the family's twiddle form (key_to_twiddle_form), or null.
    // latency != 0 asks for the two-CTA cluster kernel (one ciphertext per CTA pair; k = 1, level = 1).
    virtual bool blind_rotate(uint64_t* acc_out, const uint64_t* lut, size_t lut_count,
                              const unsigned* switched, const uint64_t* bsk, const uint64_t* bsk_tw,
                              size_t n_lwe, size_t glwe_size, unsigned base_log, unsigned level,
                              size_t batch, int bnf, unsigned width, int latency,
                              cudaStream_t stream) const {
        (void)latency;
        (void)acc_out, (void)lut, (void)lut_count, (void)switched, (void)bsk, (void)bsk_tw, (void)n_lwe;
        (void)glwe_size, (void)base_log, (void)level, (void)batch, (void)bnf, (void)width, (void)stream;
        return false;
    }
    // out = the key `in` ([matrices][glwe_size][glwe_size][n]) in the layout and form the fused blind
    // rotation reads; false: none for this family / shape
    virtual bool key_to_twiddle_form(uint64_t* out, const uint64_t* in, size_t matrices, size_t glwe_size,
                                     cudaStream_t stream) const {
        (void)out, (void)in, (void)matrices, (void)glwe_size, (void)stream;
        return false;
    }
    virtual std::shared_ptr<PrimePlan> clone() const = 0;
    virtual bool raw_shoup32h(RawShoup32H*) const { return false; }
};

// nullptr <=> the reference's try_new returns None.  Throws CudaError on CUDA failures.
std::shared_ptr<PrimePlan> make_plan64(size_t n, uint64_t p);
std::shared_ptr<PrimePlan> make_plan32(size_t n, uint32_t p);

// RAII device scope
struct DeviceGuard {
    int prev = 0;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        cur = dev;
    }
    ~DeviceGuard() {
        if (prev != cur) cudaSetDevice(prev);
    }
    int cur;
};

}  // namespace nttb200
