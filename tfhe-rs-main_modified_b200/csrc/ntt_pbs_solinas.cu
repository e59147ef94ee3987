// Fused blind-rotation kernels (ntt_pbs_fused.cuh) for the Solinas prime -- the modulus of every
// NTT-PBS parameter set in the reference (tfhe/src/core_crypto/algorithms/test/mod.rs:106-130).
// Other primes use the composed path of capi_pbs.cu.
#include <algorithm>
#include <cstdlib>

#include "ntt_engine.cuh"
#include "ntt_pbs_fused.cuh"

namespace nttb200 {
namespace {

template <class A, int LOGN, int GS, bool BNF, bool SINGLE, bool REGACC = false>
bool launch_blind_rotate(uint64_t* acc_out, const uint64_t* lut, size_t lut_count, const unsigned* switched,
                         const uint64_t* bsk_tw, size_t n_lwe, unsigned base_log, unsigned level, size_t batch,
                         unsigned width, const typename A::TW* tw_fwd, const typename A::TW* tw_inv,
                         const typename A::Ctx& c, typename A::TW n_inv, cudaStream_t st) {
    auto kern = ntt_fast_blind_rotate_kernel<A, LOGN, GS, BNF, SINGLE, REGACC>;
    size_t smem = PbsShape<LOGN, GS>::bytes(n_lwe);
    if (smem > size_t(227) * 1024) return false;
    NTT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ulonglong2* scratch = nullptr;
    if (!SINGLE && !REGACC) NTT_CUDA_CHECK(cudaMallocAsync(&scratch, batch * ((size_t)GS << LOGN) * 8, st));
    kern<<<(unsigned)batch, FastShape<LOGN>::kThreadsPerPoly, smem, st>>>(
        acc_out, lut, lut_count, switched, bsk_tw, scratch, (unsigned)n_lwe, base_log, level, width, tw_fwd, tw_inv,
        c, n_inv);
    cudaError_t e = cudaGetLastError();
    if (scratch) cudaFreeAsync(scratch, st);
    NTT_CUDA_CHECK(e);
    return true;
}

// NTT_B200_PBS_NO_REGACC=1 keeps the multi-level shapes on the L2-scratch accumulators (A/B measurements)
bool reg_acc_enabled() {
    static const bool on = [] {
        const char* e = std::getenv("NTT_B200_PBS_NO_REGACC");
        return !(e && e[0] == '1');
    }();
    return on;
}

template <class A, int LOGN, int GS, class... Args>
bool by_variant(int bnf, unsigned level, Args... args) {
    constexpr bool kCanSingle = PbsShape<LOGN, GS>::kGroups == 1;
    if constexpr (kCanSingle) {
        if (level == 1)
            return bnf ? launch_blind_rotate<A, LOGN, GS, true, true>(args...)
                       : launch_blind_rotate<A, LOGN, GS, false, true>(args...);
        if (reg_acc_enabled())  // level > 1, one transform group per level: accumulators in registers
            return bnf ? launch_blind_rotate<A, LOGN, GS, true, false, true>(args...)
                       : launch_blind_rotate<A, LOGN, GS, false, false, true>(args...);
    }
    return bnf ? launch_blind_rotate<A, LOGN, GS, true, false>(args...)
               : launch_blind_rotate<A, LOGN, GS, false, false>(args...);
}
template <class A, int LOGN, class... Args>
bool by_glwe_size(size_t gs, int bnf, unsigned level, Args... args) {
    switch (gs) {
        case 2: return by_variant<A, LOGN, 2>(bnf, level, args...);
        case 3: return by_variant<A, LOGN, 3>(bnf, level, args...);
        case 4: return by_variant<A, LOGN, 4>(bnf, level, args...);
        default: return false;
    }
}

// Montgomery form (x * 2^64 mod p) in the permuted layout of pbs_key_index
template <int LOGN, int GS>
__global__ void solinas_key_to_twiddle_form_kernel(uint64_t* __restrict__ out, const uint64_t* __restrict__ in,
                                                   size_t matrices) {
    constexpr unsigned N = 1u << LOGN, PPT = PbsShape<LOGN, GS>::kPPT;
    const size_t per = (size_t)GS * GS * N, total = matrices * per;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t m = i / per, r = i % per;
        unsigned row = (unsigned)(r / ((size_t)GS * N)), cc = (unsigned)((r / N) % GS), coef = (unsigned)(r % N);
        unsigned t = coef >> 3, q = (coef >> 1) & 3, e = coef & 1;
        out[m * per + pbs_key_index<LOGN, GS>(row / PPT, q, t, row % PPT, cc, e)] =
            Solinas64::mul_plain(in[i], Solinas64::EPS);
    }
}
template <int LOGN>
bool key_form_by_glwe_size(uint64_t* out, const uint64_t* in, size_t matrices, size_t gs, cudaStream_t st) {
    size_t total = matrices * gs * gs << LOGN;
    unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 16);
    switch (gs) {
        case 2: solinas_key_to_twiddle_form_kernel<LOGN, 2><<<blocks, 256, 0, st>>>(out, in, matrices); break;
        case 3: solinas_key_to_twiddle_form_kernel<LOGN, 3><<<blocks, 256, 0, st>>>(out, in, matrices); break;
        case 4: solinas_key_to_twiddle_form_kernel<LOGN, 4><<<blocks, 256, 0, st>>>(out, in, matrices); break;
        default: return false;
    }
    NTT_CUDA_CHECK(cudaGetLastError());
    return true;
}

}  // namespace

template <>
bool fast_blind_rotate<Solinas64>(uint64_t* acc_out, const uint64_t* lut, size_t lut_count,
                                  const unsigned* switched, const uint64_t* bsk_plain, const uint64_t* bsk,
                                  size_t n_lwe, size_t glwe_size, unsigned base_log, unsigned level, size_t batch,
                                  int bnf, unsigned width, int logn, const uint64_t* tw_fwd, const uint64_t* tw_inv,
                                  const Solinas64::Ctx& c, uint64_t n_inv, cudaStream_t st) {
    using A = Solinas64;
    (void)bsk_plain;
    if (!bsk) return false;  // the key has no twiddle-form copy
    if (!batch) return true;
    if (batch > 0x7fffffffull) return false;
#define NTT_PBS_CASE(L)                                                                                          \
    case L:                                                                                                      \
        return by_glwe_size<A, L>(glwe_size, bnf, level, acc_out, lut, lut_count, switched, bsk, n_lwe, base_log, \
                                  level, batch, width, tw_fwd, tw_inv, c, n_inv, st);
    switch (logn) {
        NTT_PBS_CASE(8)
        NTT_PBS_CASE(9)
        NTT_PBS_CASE(10)
        NTT_PBS_CASE(11)
        NTT_PBS_CASE(12)
        default: return false;
    }
#undef NTT_PBS_CASE
}

namespace {
template <class A, int LOGN, bool BNF>
bool launch_cluster(uint64_t* acc_out, const uint64_t* lut, size_t lut_count, const unsigned* switched,
                    const uint64_t* bsk_tw, size_t n_lwe, unsigned base_log, size_t batch, unsigned width,
                    const typename A::TW* tw_fwd, const typename A::TW* tw_inv, const typename A::Ctx& c,
                    typename A::TW n_inv, cudaStream_t st) {
    auto kern = ntt_fast_blind_rotate_cluster2_kernel<A, LOGN, BNF>;
    size_t smem = PbsClusterShape<LOGN>::bytes(n_lwe);
    if (smem > size_t(227) * 1024) return false;
    NTT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)(2 * batch), FastShape<LOGN>::kThreadsPerPoly, smem, st>>>(
        acc_out, lut, lut_count, switched, bsk_tw, (unsigned)n_lwe, base_log, width, tw_fwd, tw_inv, c, n_inv);
    NTT_CUDA_CHECK(cudaGetLastError());
    return true;
}
}  // namespace

template <>
bool fast_blind_rotate_cluster<Solinas64>(uint64_t* acc_out, const uint64_t* lut, size_t lut_count,
                                          const unsigned* switched, const uint64_t* bsk_tw, size_t n_lwe,
                                          size_t glwe_size, unsigned base_log, unsigned level, size_t batch,
                                          int bnf, unsigned width, int logn, const uint64_t* tw_fwd,
                                          const uint64_t* tw_inv, const Solinas64::Ctx& c, uint64_t n_inv,
                                          cudaStream_t st) {
    using A = Solinas64;
    if (!bsk_tw || glwe_size != 2 || level != 1) return false;
    if (!batch) return true;
    if (batch > 0x3fffffffull) return false;
#define NTT_PBS_CASE(L)                                                                                         \
    case L:                                                                                                     \
        return bnf ? launch_cluster<A, L, true>(acc_out, lut, lut_count, switched, bsk_tw, n_lwe, base_log,     \
                                                batch, width, tw_fwd, tw_inv, c, n_inv, st)                    \
                   : launch_cluster<A, L, false>(acc_out, lut, lut_count, switched, bsk_tw, n_lwe, base_log,    \
                                                 batch, width, tw_fwd, tw_inv, c, n_inv, st);
    switch (logn) {
        NTT_PBS_CASE(9)
        NTT_PBS_CASE(10)
        NTT_PBS_CASE(11)
        NTT_PBS_CASE(12)
        default: return false;
    }
#undef NTT_PBS_CASE
}

template <>
bool fast_key_to_twiddle_form<Solinas64>(uint64_t* out, const uint64_t* in, size_t matrices, size_t glwe_size,
                                         int logn, const Solinas64::Ctx&, cudaStream_t st) {
    if (!matrices) return true;
    switch (logn) {
        case 8: return key_form_by_glwe_size<8>(out, in, matrices, glwe_size, st);
        case 9: return key_form_by_glwe_size<9>(out, in, matrices, glwe_size, st);
        case 10: return key_form_by_glwe_size<10>(out, in, matrices, glwe_size, st);
        case 11: return key_form_by_glwe_size<11>(out, in, matrices, glwe_size, st);
        case 12: return key_form_by_glwe_size<12>(out, in, matrices, glwe_size, st);
        default: return false;
    }
}

}  // namespace nttb200
