// Host-side engine: plan objects that own device tables and launch the kernels.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
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

// The same for a plan of a prime below 2^62 (the 50-bit CRT primes of the Plan52 kinds).
struct RawShoup64H {
    const ShoupTw<uint64_t>* fwd;
    const ShoupTw<uint64_t>* inv;
    Shoup<uint64_t, true>::Ctx ctx;
    ShoupTw<uint64_t> n_inv;
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
    virtual bool raw_shoup64h(RawShoup64H*) const { return false; }
};

// nullptr <=> the reference's try_new returns None.  Throws CudaError on CUDA failures.
std::shared_ptr<PrimePlan> make_plan64(size_t n, uint64_t p);
std::shared_ptr<PrimePlan> make_plan32(size_t n, uint32_t p);

// Profiling ranges (the reference's CUDA backend convention: PUSH_RANGE / POP_RANGE around every host
// entry, backends/tfhe-cuda-backend/cuda/src/utils/helper_profile.cu:15-41).  Compiled in with
// -DNTT_B200_NVTX (python build.py --nvtx); header-only NVTX v3, no extra library.  Off by default.
#ifdef NTT_B200_NVTX
}  // namespace nttb200
#include <nvtx3/nvToolsExt.h>
namespace nttb200 {
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};
#define NTT_NVTX_CAT2(a, b) a##b
#define NTT_NVTX_CAT(a, b) NTT_NVTX_CAT2(a, b)
#define NTT_NVTX(name) ::nttb200::NvtxRange NTT_NVTX_CAT(nvtx_range_, __LINE__)(name)
#else
#define NTT_NVTX(name) ((void)0)
#endif

// Kernels that need more than the default 48 KiB of dynamic shared memory opt in once per device
// (the kernel is a template argument so that the flags are per kernel, not per signature).  The flags
// are shared by concurrent host threads: atomics, and a device index beyond the table simply sets the
// attribute on every call (it is idempotent and cheap).
template <auto Kernel>
void allow_dynamic_smem(size_t bytes) {
    if (bytes <= 48 * 1024) return;
    static std::atomic<bool> done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const bool tracked = dev >= 0 && dev < 64;
    if (tracked && done[dev].load(std::memory_order_acquire)) return;
    NTT_CUDA_CHECK(cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    if (tracked) done[dev].store(true, std::memory_order_release);
}

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
