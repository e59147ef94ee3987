// Dispatch from a runtime log2(n) to the compile-time fast kernels; included by one .cu per
// modulus family so the families compile in parallel.
#pragma once
#include "ntt_engine.cuh"
#include "ntt_fast.cuh"

namespace nttb200 {

template <int LOGN>
struct FastPolys {  // polynomials per CTA: keep CTAs at >= 256 threads
    static constexpr int value = (1 << LOGN) >= 2048 ? 1 : 2048 / (1 << LOGN);
};

// Polynomials per thread: two when the polynomial is big enough for the shared index arithmetic
// and twiddle loads to matter (the tiles are dynamic shared memory: 72 KiB for two 4096-point
// u64 polynomials, still two CTAs per SM).
#ifndef NTT_FAST_U32_PPT
#define NTT_FAST_U32_PPT 4
#endif
template <class A, int LOGN, bool INV = false>
struct FastPPT {
    // 32-bit Shoup families: four polynomials per thread (16-byte tile slots, TileLayout::kQuad) halve the
    // per-polynomial index arithmetic and twiddle loads once more -- forward at n = 1024 / 2048, inverse at
    // n = 1024 (at 2048 the inverse is as fast with two, at 4096 four lose 20-30 %); short polynomials
    // (n = 256, 512) take two per thread as well: +10..20 % (profiles/r02_u32_kernel_variants.txt)
    static constexpr bool kU32 = sizeof(typename A::T) == 4 && !std::is_same<A, Wide32>::value;
    static constexpr bool kFour = kU32 && (LOGN == 10 || (LOGN == 11 && !INV));
    static constexpr int value = kFour ? NTT_FAST_U32_PPT : 2;
};

// The fused fwd -> pointwise -> inv kernel needs more registers; with two 4096-point u64
// polynomials per thread it drops to one CTA per SM and loses (measured), so it keeps one there.
template <class A, int LOGN>
struct FastPPTFused {
    static constexpr bool fits =
        2 * FastPolys<LOGN>::value * FastShape<LOGN>::kPaddedElems * sizeof(typename A::T) <= 48 * 1024;
#ifndef NTT_FUSED_U32_PPT
#define NTT_FUSED_U32_PPT 2
#endif
    static constexpr bool kU32 = sizeof(typename A::T) == 4 && !std::is_same<A, Wide32>::value;
    static constexpr int value = (kU32 && (LOGN == 10 || LOGN == 11)) ? NTT_FUSED_U32_PPT : (fits ? 2 : 1);
};

template <class A, int LOGN, int PPT>
constexpr size_t fast_smem_bytes() {
    return (size_t)FastPolys<LOGN>::value * PPT * FastShape<LOGN>::kPaddedElems * sizeof(typename A::T);
}
template <auto Kernel>
void fast_allow_smem(size_t bytes) {
    allow_dynamic_smem<Kernel>(bytes);
}

template <class A, int LOGN, int PPT>
void launch_fast_fwd_p(typename A::T* data, size_t rows, unsigned depth, const typename A::TW* tw,
                       const typename A::Ctx& c, cudaStream_t st) {
    constexpr int P = FastPolys<LOGN>::value;
    dim3 block(FastShape<LOGN>::kThreadsPerPoly, P);
    size_t polys = rows >> depth, groups = ((polys + PPT - 1) / PPT) << depth;
    unsigned grid = (unsigned)((groups + P - 1) / P);
    constexpr size_t smem = fast_smem_bytes<A, LOGN, PPT>();
    if (depth == 0) {
        fast_allow_smem<ntt_fast_fwd_kernel<A, LOGN, P, PPT, false>>(smem);
        ntt_fast_fwd_kernel<A, LOGN, P, PPT, false><<<grid, block, smem, st>>>(data, rows, 0u, tw, c);
    } else {
        fast_allow_smem<ntt_fast_fwd_kernel<A, LOGN, P, PPT, true>>(smem);
        ntt_fast_fwd_kernel<A, LOGN, P, PPT, true><<<grid, block, smem, st>>>(data, rows, depth, tw, c);
    }
}
template <class A, int LOGN>
void launch_fast_fwd(typename A::T* data, size_t rows, unsigned depth, const typename A::TW* tw,
                     const typename A::Ctx& c, cudaStream_t st) {
    constexpr int kP = FastPPT<A, LOGN>::value;
    if (kP == 4 && (rows >> depth) >= 4)
        launch_fast_fwd_p<A, LOGN, kP>(data, rows, depth, tw, c, st);
    else if (kP >= 2 && (rows >> depth) > 1)
        launch_fast_fwd_p<A, LOGN, 2>(data, rows, depth, tw, c, st);
    else
        launch_fast_fwd_p<A, LOGN, 1>(data, rows, depth, tw, c, st);
}
template <class A, int LOGN, int PPT>
void launch_fast_inv_p(typename A::T* data, size_t rows, unsigned depth, const typename A::TW* tw,
                       const typename A::Ctx& c, cudaStream_t st) {
    constexpr int P = FastPolys<LOGN>::value;
    dim3 block(FastShape<LOGN>::kThreadsPerPoly, P);
    size_t polys = rows >> depth, groups = ((polys + PPT - 1) / PPT) << depth;
    unsigned grid = (unsigned)((groups + P - 1) / P);
    constexpr size_t smem = fast_smem_bytes<A, LOGN, PPT>();
    if (depth == 0) {
        fast_allow_smem<ntt_fast_inv_kernel<A, LOGN, P, PPT, false>>(smem);
        ntt_fast_inv_kernel<A, LOGN, P, PPT, false><<<grid, block, smem, st>>>(data, rows, 0u, tw, c);
    } else {
        fast_allow_smem<ntt_fast_inv_kernel<A, LOGN, P, PPT, true>>(smem);
        ntt_fast_inv_kernel<A, LOGN, P, PPT, true><<<grid, block, smem, st>>>(data, rows, depth, tw, c);
    }
}
template <class A, int LOGN>
void launch_fast_inv(typename A::T* data, size_t rows, unsigned depth, const typename A::TW* tw,
                     const typename A::Ctx& c, cudaStream_t st) {
    constexpr int kP = FastPPT<A, LOGN, true>::value;
    if (kP == 4 && (rows >> depth) >= 4)
        launch_fast_inv_p<A, LOGN, kP>(data, rows, depth, tw, c, st);
    else if (kP >= 2 && (rows >> depth) > 1)
        launch_fast_inv_p<A, LOGN, 2>(data, rows, depth, tw, c, st);
    else
        launch_fast_inv_p<A, LOGN, 1>(data, rows, depth, tw, c, st);
}
template <class A, int LOGN>
void launch_fast_fmi(typename A::T* out, const typename A::T* lhs, const typename A::T* rhs,
                     size_t rhs_polys, const typename A::T* acc, size_t acc_polys, size_t batch,
                     const typename A::TW* tw_fwd, const typename A::TW* tw_inv,
                     const typename A::Ctx& c, cudaStream_t st) {
    constexpr int P = FastPolys<LOGN>::value;
    constexpr int PPT = FastPPTFused<A, LOGN>::value;
    dim3 block(FastShape<LOGN>::kThreadsPerPoly, P);
    unsigned grid = (unsigned)((batch + P * PPT - 1) / (P * PPT));
    constexpr size_t smem = fast_smem_bytes<A, LOGN, PPT>();
    fast_allow_smem<ntt_fast_fwd_mac_inv_kernel<A, LOGN, P, PPT>>(smem);
    ntt_fast_fwd_mac_inv_kernel<A, LOGN, P, PPT><<<grid, block, smem, st>>>(
        out, lhs, rhs, rhs_polys, acc, acc_polys ? acc_polys : 1, batch, tw_fwd, tw_inv, c);
}

// (A 2^13-point single-CTA kernel -- 1024 threads, 36 KiB tile -- was measured for the 32-bit families and
// is slower stand-alone than the TMA-staged radix-16 pass + 512-point kernels: 31.9 / 36.3 against 36.9 /
// 38.9 M NTT/s, profiles/r02_large_n_u32.txt.  The fused CRT kernels do use fwd_from_regs<.., 13, ..>.)
#define NTT_FAST_SWITCH(CALL)          \
    switch (logn) {                    \
        case 8: CALL(8); break;        \
        case 9: CALL(9); break;        \
        case 10: CALL(10); break;      \
        case 11: CALL(11); break;      \
        case 12: CALL(12); break;      \
        default: return false;         \
    }


template <class A>
bool fast_fwd_impl(typename A::T* data, size_t batch, int logn, unsigned depth,
                   const typename A::TW* tw, const typename A::Ctx& c, cudaStream_t st) {
#define NTT_CALL(L) launch_fast_fwd<A, L>(data, batch, depth, tw, c, st)
    NTT_FAST_SWITCH(NTT_CALL)
#undef NTT_CALL
    NTT_CUDA_CHECK(cudaGetLastError());
    return true;
}
template <class A>
bool fast_inv_impl(typename A::T* data, size_t batch, int logn, unsigned depth,
                   const typename A::TW* tw, const typename A::Ctx& c, cudaStream_t st) {
#define NTT_CALL(L) launch_fast_inv<A, L>(data, batch, depth, tw, c, st)
    NTT_FAST_SWITCH(NTT_CALL)
#undef NTT_CALL
    NTT_CUDA_CHECK(cudaGetLastError());
    return true;
}
template <class A>
bool fast_fmi_impl(typename A::T* out, const typename A::T* lhs, const typename A::T* rhs,
                   size_t rhs_polys, const typename A::T* acc, size_t acc_polys, size_t batch,
                   int logn, const typename A::TW* tw_fwd, const typename A::TW* tw_inv,
                   const typename A::Ctx& c, cudaStream_t st) {
#define NTT_CALL(L) \
    launch_fast_fmi<A, L>(out, lhs, rhs, rhs_polys, acc, acc_polys, batch, tw_fwd, tw_inv, c, st)
    NTT_FAST_SWITCH(NTT_CALL)
#undef NTT_CALL
    NTT_CUDA_CHECK(cudaGetLastError());
    return true;
}

template <class A, int LOGN>
bool launch_fast_ext(typename A::T* out, const typename A::T* in, const typename A::T* ggsw,
                     unsigned rows, unsigned cols, size_t batch, const typename A::TW* tw_fwd,
                     const typename A::TW* tw_inv, const typename A::Ctx& c, cudaStream_t st) {
    dim3 block(FastShape<LOGN>::kThreadsPerPoly);
    unsigned grid = (unsigned)batch;
    if (cols < 1 || cols > 4) return false;
    // the shared GGSW goes to the family's pointwise form once per call (rows * cols polynomials)
    typename A::T* gform = nullptr;
    if (sizeof(typename A::T) == 8) {
        size_t total = ((size_t)rows * cols) << LOGN;
        NTT_CUDA_CHECK(cudaMallocAsync(&gform, total * sizeof(typename A::T), st));
        pw_form_kernel<A><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(gform, ggsw, total, c);
        ggsw = gform;
    }
    struct Release {
        typename A::T* p;
        cudaStream_t st;
        ~Release() {
            if (p) cudaFreeAsync(p, st);
        }
    } release{gform, st};
    switch (cols) {
        case 1: ntt_fast_ext_product_kernel<A, LOGN, 1><<<grid, block, 0, st>>>(out, in, ggsw, rows, tw_fwd, tw_inv, c); break;
        case 2: ntt_fast_ext_product_kernel<A, LOGN, 2><<<grid, block, 0, st>>>(out, in, ggsw, rows, tw_fwd, tw_inv, c); break;
        case 3: ntt_fast_ext_product_kernel<A, LOGN, 3><<<grid, block, 0, st>>>(out, in, ggsw, rows, tw_fwd, tw_inv, c); break;
        case 4: ntt_fast_ext_product_kernel<A, LOGN, 4><<<grid, block, 0, st>>>(out, in, ggsw, rows, tw_fwd, tw_inv, c); break;
        default: return false;
    }
    return true;
}
template <class A>
bool fast_ext_impl(typename A::T* out, const typename A::T* in, const typename A::T* ggsw,
                   unsigned rows, unsigned cols, size_t batch, int logn,
                   const typename A::TW* tw_fwd, const typename A::TW* tw_inv,
                   const typename A::Ctx& c, cudaStream_t st) {
    bool ok = false;
    switch (logn) {  // the PBS polynomial sizes
        case 9: ok = launch_fast_ext<A, 9>(out, in, ggsw, rows, cols, batch, tw_fwd, tw_inv, c, st); break;
        case 10: ok = launch_fast_ext<A, 10>(out, in, ggsw, rows, cols, batch, tw_fwd, tw_inv, c, st); break;
        case 11: ok = launch_fast_ext<A, 11>(out, in, ggsw, rows, cols, batch, tw_fwd, tw_inv, c, st); break;
        case 12: ok = launch_fast_ext<A, 12>(out, in, ggsw, rows, cols, batch, tw_fwd, tw_inv, c, st); break;
        default: return false;
    }
    if (ok) NTT_CUDA_CHECK(cudaGetLastError());
    return ok;
}

template <class A, int LOGSUB, bool INV>
void launch_cluster8(typename A::T* data, size_t polys, const typename A::TW* tw, const typename A::Ctx& c,
                     cudaStream_t st) {
    constexpr size_t smem = (size_t)FastShape<LOGSUB>::kPaddedElems * sizeof(typename A::T);
    constexpr auto kern = INV ? ntt_cluster8_inv_kernel<A, LOGSUB> : ntt_cluster8_fwd_kernel<A, LOGSUB>;
    fast_allow_smem<kern>(smem);
    // grid.x limit is 2^31 - 1 CTAs: far above any batch that fits the device
    kern<<<(unsigned)(polys * 8), FastShape<LOGSUB>::kThreadsPerPoly, smem, st>>>(data, tw, c);
}
template <class A, bool INV>
bool fast_cluster_impl(typename A::T* data, size_t polys, int logn, const typename A::TW* tw,
                       const typename A::Ctx& c, cudaStream_t st) {
    switch (logn) {
        case 13: launch_cluster8<A, 10, INV>(data, polys, tw, c, st); break;
        case 14: launch_cluster8<A, 11, INV>(data, polys, tw, c, st); break;
        default: return false;
    }
    NTT_CUDA_CHECK(cudaGetLastError());
    return true;
}

// explicit specialisations of the entry points declared in ntt_fast.cuh
#define NTT_DEFINE_FAST(A)                                                                         \
    template <>                                                                                    \
    bool fast_fwd<A>(A::T * data, size_t batch, int logn, unsigned depth, const A::TW* tw,         \
                     const A::Ctx& c, cudaStream_t st) {                                           \
        return fast_fwd_impl<A>(data, batch, logn, depth, tw, c, st);                              \
    }                                                                                              \
    template <>                                                                                    \
    bool fast_inv<A>(A::T * data, size_t batch, int logn, unsigned depth, const A::TW* tw,         \
                     const A::Ctx& c, cudaStream_t st) {                                           \
        return fast_inv_impl<A>(data, batch, logn, depth, tw, c, st);                              \
    }                                                                                              \
    template <>                                                                                    \
    bool fast_cluster_fwd<A>(A::T * data, size_t polys, int logn, const A::TW* tw, const A::Ctx& c, \
                             cudaStream_t st) {                                                    \
        return fast_cluster_impl<A, false>(data, polys, logn, tw, c, st);                          \
    }                                                                                              \
    template <>                                                                                    \
    bool fast_cluster_inv<A>(A::T * data, size_t polys, int logn, const A::TW* tw, const A::Ctx& c, \
                             cudaStream_t st) {                                                    \
        return fast_cluster_impl<A, true>(data, polys, logn, tw, c, st);                           \
    }                                                                                              \
    template <>                                                                                    \
    bool fast_ext_product<A>(A::T * out, const A::T* in, const A::T* ggsw, unsigned rows,          \
                             unsigned cols, size_t batch, int logn, const A::TW* tw_fwd,           \
                             const A::TW* tw_inv, const A::Ctx& c, cudaStream_t st) {              \
        return fast_ext_impl<A>(out, in, ggsw, rows, cols, batch, logn, tw_fwd, tw_inv, c, st);    \
    }                                                                                              \
    template <>                                                                                    \
    bool fast_fwd_mac_inv<A>(A::T * out, const A::T* lhs, const A::T* rhs, size_t rhs_polys,       \
                             const A::T* acc, size_t acc_polys, size_t batch, int logn,            \
                             const A::TW* tw_fwd, const A::TW* tw_inv, const A::Ctx& c,            \
                             cudaStream_t st) {                                                    \
        return fast_fmi_impl<A>(out, lhs, rhs, rhs_polys, acc, acc_polys, batch, logn, tw_fwd,     \
                                tw_inv, c, st);                                                    \
    }

}  // namespace nttb200
