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

// Polynomials per thread: two when both tiles fit the 48 KiB static shared-memory window and the
// polynomial is big enough for the shared index arithmetic to matter.
template <class A, int LOGN>
struct FastPPT {
    static constexpr bool fits =
        2 * FastPolys<LOGN>::value * FastShape<LOGN>::kPaddedElems * sizeof(typename A::T) <= 48 * 1024;
    static constexpr int value = (LOGN >= 10 && fits) ? 2 : 1;
};

template <class A, int LOGN, int PPT>
void launch_fast_fwd_p(typename A::T* data, size_t rows, unsigned depth, const typename A::TW* tw,
                       const typename A::Ctx& c, cudaStream_t st) {
    constexpr int P = FastPolys<LOGN>::value;
    dim3 block(FastShape<LOGN>::kThreadsPerPoly, P);
    unsigned grid = (unsigned)((rows + P * PPT - 1) / (P * PPT));
    ntt_fast_fwd_kernel<A, LOGN, P, PPT><<<grid, block, 0, st>>>(data, rows, depth, tw, c);
}
template <class A, int LOGN>
void launch_fast_fwd(typename A::T* data, size_t rows, unsigned depth, const typename A::TW* tw,
                     const typename A::Ctx& c, cudaStream_t st) {
    if (FastPPT<A, LOGN>::value == 2 && depth == 0 && rows > 1)
        launch_fast_fwd_p<A, LOGN, FastPPT<A, LOGN>::value>(data, rows, depth, tw, c, st);
    else
        launch_fast_fwd_p<A, LOGN, 1>(data, rows, depth, tw, c, st);
}
template <class A, int LOGN, int PPT>
void launch_fast_inv_p(typename A::T* data, size_t rows, unsigned depth, const typename A::TW* tw,
                       const typename A::Ctx& c, cudaStream_t st) {
    constexpr int P = FastPolys<LOGN>::value;
    dim3 block(FastShape<LOGN>::kThreadsPerPoly, P);
    unsigned grid = (unsigned)((rows + P * PPT - 1) / (P * PPT));
    ntt_fast_inv_kernel<A, LOGN, P, PPT><<<grid, block, 0, st>>>(data, rows, depth, tw, c);
}
template <class A, int LOGN>
void launch_fast_inv(typename A::T* data, size_t rows, unsigned depth, const typename A::TW* tw,
                     const typename A::Ctx& c, cudaStream_t st) {
    if (FastPPT<A, LOGN>::value == 2 && depth == 0 && rows > 1)
        launch_fast_inv_p<A, LOGN, FastPPT<A, LOGN>::value>(data, rows, depth, tw, c, st);
    else
        launch_fast_inv_p<A, LOGN, 1>(data, rows, depth, tw, c, st);
}
template <class A, int LOGN>
void launch_fast_fmi(typename A::T* out, const typename A::T* lhs, const typename A::T* rhs,
                     size_t rhs_polys, const typename A::T* acc, size_t acc_polys, size_t batch,
                     const typename A::TW* tw_fwd, const typename A::TW* tw_inv,
                     const typename A::Ctx& c, cudaStream_t st) {
    constexpr int P = FastPolys<LOGN>::value;
    constexpr int PPT = FastPPT<A, LOGN>::value;
    dim3 block(FastShape<LOGN>::kThreadsPerPoly, P);
    unsigned grid = (unsigned)((batch + P * PPT - 1) / (P * PPT));
    ntt_fast_fwd_mac_inv_kernel<A, LOGN, P, PPT><<<grid, block, 0, st>>>(
        out, lhs, rhs, rhs_polys, acc, acc_polys ? acc_polys : 1, batch, tw_fwd, tw_inv, c);
}

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
