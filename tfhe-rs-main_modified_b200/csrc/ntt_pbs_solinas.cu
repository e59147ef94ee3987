// Fused blind-rotation kernels (ntt_pbs_fused.cuh) for the Solinas prime -- the modulus of every
// NTT-PBS parameter set in the reference (tfhe/src/core_crypto/algorithms/test/mod.rs:106-130).
// Other primes use the composed path of capi_pbs.cu.
#include <algorithm>

#include "ntt_engine.cuh"
#include "ntt_pbs_fused.cuh"

namespace nttb200 {
namespace {

template <class A, int LOGN, int GS, bool BNF, bool MONT>
bool launch_blind_rotate(uint64_t* acc_out, const uint64_t* lut, size_t lut_count, const unsigned* switched,
                         const uint64_t* bsk, size_t n_lwe, unsigned base_log, unsigned level, size_t batch,
                         unsigned width, const typename A::TW* tw_fwd, const typename A::TW* tw_inv,
                         const typename A::Ctx& c, typename A::TW n_inv, cudaStream_t st) {
    auto kern = ntt_fast_blind_rotate_kernel<A, LOGN, GS, BNF, MONT>;
    size_t smem = PbsShape<LOGN, GS>::bytes(n_lwe);
    if (smem > size_t(227) * 1024) return false;
    NTT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)batch, FastShape<LOGN>::kThreadsPerPoly, smem, st>>>(
        acc_out, lut, lut_count, switched, bsk, (unsigned)n_lwe, base_log, level, width, tw_fwd, tw_inv, c, n_inv);
    NTT_CUDA_CHECK(cudaGetLastError());
    return true;
}

// the fused kernels read the key in twiddle (Montgomery) form
template <class A, int LOGN, int GS, class... Args>
bool by_variant(int bnf, Args... args) {
    return bnf ? launch_blind_rotate<A, LOGN, GS, true, true>(args...)
               : launch_blind_rotate<A, LOGN, GS, false, true>(args...);
}

__global__ void solinas_to_montgomery_kernel(uint64_t* __restrict__ out, const uint64_t* __restrict__ in,
                                             size_t total) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        out[i] = Solinas64::mul_plain(in[i], Solinas64::EPS);  // * 2^64 mod p
}
template <class A, int LOGN, class... Args>
bool by_glwe_size(size_t gs, int bnf, Args... args) {
    switch (gs) {
        case 2: return by_variant<A, LOGN, 2>(bnf, args...);
        case 3: return by_variant<A, LOGN, 3>(bnf, args...);
        case 4: return by_variant<A, LOGN, 4>(bnf, args...);
        default: return false;
    }
}

}  // namespace

template <>
bool fast_blind_rotate<Solinas64>(uint64_t* acc_out, const uint64_t* lut, size_t lut_count,
                                  const unsigned* switched, const uint64_t* bsk_plain, const uint64_t* bsk,
                                  size_t n_lwe, size_t glwe_size, unsigned base_log, unsigned level, size_t batch, int bnf,
                                  unsigned width, int logn, const uint64_t* tw_fwd, const uint64_t* tw_inv,
                                  const Solinas64::Ctx& c, uint64_t n_inv, cudaStream_t st) {
    using A = Solinas64;
    (void)bsk_plain;
    if (!bsk) return false;  // the key has no Montgomery copy
    if (!batch) return true;
    if (batch > 0x7fffffffull) return false;
    switch (logn) {
        case 8: return by_glwe_size<A, 8>(glwe_size, bnf, acc_out, lut, lut_count, switched, bsk, n_lwe, base_log, level, batch, width, tw_fwd, tw_inv, c, n_inv, st);
        case 9: return by_glwe_size<A, 9>(glwe_size, bnf, acc_out, lut, lut_count, switched, bsk, n_lwe, base_log, level, batch, width, tw_fwd, tw_inv, c, n_inv, st);
        case 10: return by_glwe_size<A, 10>(glwe_size, bnf, acc_out, lut, lut_count, switched, bsk, n_lwe, base_log, level, batch, width, tw_fwd, tw_inv, c, n_inv, st);
        case 11: return by_glwe_size<A, 11>(glwe_size, bnf, acc_out, lut, lut_count, switched, bsk, n_lwe, base_log, level, batch, width, tw_fwd, tw_inv, c, n_inv, st);
        case 12: return by_glwe_size<A, 12>(glwe_size, bnf, acc_out, lut, lut_count, switched, bsk, n_lwe, base_log, level, batch, width, tw_fwd, tw_inv, c, n_inv, st);
        default: return false;
    }
}

template <>
bool fast_key_to_twiddle_form<Solinas64>(uint64_t* out, const uint64_t* in, size_t total, const Solinas64::Ctx&,
                                         cudaStream_t st) {
    if (!total) return true;
    unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 16);
    solinas_to_montgomery_kernel<<<blocks, 256, 0, st>>>(out, in, total);
    NTT_CUDA_CHECK(cudaGetLastError());
    return true;
}

}  // namespace nttb200
