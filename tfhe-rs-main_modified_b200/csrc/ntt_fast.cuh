// Fast single-CTA NTT kernels with compile-time size (2^8 .. 2^12 coefficients).
//
// One thread owns 8 coefficients; a polynomial is N/8 threads.  The log2(N) radix-2 stages of
// the reference (generic_solinas.rs:449-514 / shoup.rs:544-615) are executed as
//   [radix-8 register passes over strides N/8, N/64, ...]  +  [a last pass over 8 consecutive
//   coefficients holding the remaining 1..3 stages]
// Forward: the first pass reads global memory directly (coalesced 64/32-bit loads, stride N/8),
// passes exchange data through shared memory, the last pass writes 8 consecutive coefficients
// per thread with 128-bit stores.  Inverse is the exact mirror.
//
// Shared-memory layout: 16 bytes of padding after every 128 bytes of a polynomial; with it
// every pass's 64-bit accesses and the last pass's 128-bit accesses are bank-conflict free
// (DESIGN.md "shared-memory layout").
#pragma once
#include "ntt_kernels.cuh"

namespace nttb200 {

template <int LOGN>
struct FastShape {
    static constexpr int kLastStages = (LOGN % 3 == 0) ? 3 : (LOGN % 3);
    static constexpr int kRadix8Passes = (LOGN - kLastStages) / 3;
    static constexpr int kThreadsPerPoly = (1 << LOGN) / 8;
    static constexpr int kPaddedElems = (1 << LOGN) + ((1 << LOGN) >> 3);  // 16 B per 128 B
};

// 16 bytes of padding after every 128 bytes: u64 -> a + 2*(a>>4), u32 -> a + 4*(a>>5)
template <class T>
NTT_DEVINL constexpr unsigned pad_index(unsigned a) {
    return sizeof(T) == 8 ? a + 2u * (a >> 4) : a + 4u * (a >> 5);
}

template <class T>
NTT_DEVINL T ldg_tw(const T* p) {
    return __ldg(p);
}
template <class T>
NTT_DEVINL ShoupTw<T> ldg_tw(const ShoupTw<T>* p);
template <>
NTT_DEVINL ShoupTw<uint32_t> ldg_tw(const ShoupTw<uint32_t>* p) {
    uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    return ShoupTw<uint32_t>{v.x, v.y};
}
template <>
NTT_DEVINL ShoupTw<uint64_t> ldg_tw(const ShoupTw<uint64_t>* p) {
    ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(p));
    return ShoupTw<uint64_t>{v.x, v.y};
}

// radix-2^R butterflies on x[OFF .. OFF + 2^R) with twiddles fetched through the read-only path
template <class A, int R, int OFF, bool INV>
NTT_DEVINL void tuple_ro(typename A::T (&x)[8], const typename A::TW* __restrict__ tw, unsigned w0,
                         const typename A::Ctx& c) {
    if (!INV) {
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const int d = 1 << (R - 1 - q);
#pragma unroll
            for (int h = 0; h < (1 << q); ++h) {
                typename A::TW w = ldg_tw(tw + ((w0 << q) + h));
#pragma unroll
                for (int k = 0; k < d; ++k)
                    A::fwd_bf(c, x[OFF + h * 2 * d + k], x[OFF + h * 2 * d + k + d], w);
            }
        }
    } else {
#pragma unroll
        for (int q = R - 1; q >= 0; --q) {
            const int d = 1 << (R - 1 - q);
#pragma unroll
            for (int h = 0; h < (1 << q); ++h) {
                typename A::TW w = ldg_tw(tw + ((w0 << q) + h));
#pragma unroll
                for (int k = 0; k < d; ++k)
                    A::inv_bf(c, x[OFF + h * 2 * d + k], x[OFF + h * 2 * d + k + d], w);
            }
        }
    }
}

// The "last" pass: 8 consecutive coefficients 8u .. 8u+7 of the polynomial, stages
// [LOGN - S, LOGN).  S = 3: one radix-8 tuple; S = 2: two radix-4 tuples; S = 1: four radix-2.
template <class A, int LOGN, bool INV>
NTT_DEVINL void last_pass(typename A::T (&x)[8], const typename A::TW* __restrict__ tw, unsigned u,
                          const typename A::Ctx& c) {
    constexpr int S = FastShape<LOGN>::kLastStages;
    constexpr unsigned m = 1u << (LOGN - S);  // groups in the first fused stage
    if (S == 3) {
        tuple_ro<A, 3, 0, INV>(x, tw, m + u, c);
    } else if (S == 2) {
        tuple_ro<A, 2, 0, INV>(x, tw, m + 2 * u, c);
        tuple_ro<A, 2, 4, INV>(x, tw, m + 2 * u + 1, c);
    } else {
        tuple_ro<A, 1, 0, INV>(x, tw, m + 4 * u, c);
        tuple_ro<A, 1, 2, INV>(x, tw, m + 4 * u + 1, c);
        tuple_ro<A, 1, 4, INV>(x, tw, m + 4 * u + 2, c);
        tuple_ro<A, 1, 6, INV>(x, tw, m + 4 * u + 3, c);
    }
}

// 8 consecutive elements <-> global memory with 128-bit accesses
template <class T>
NTT_DEVINL void load8_consecutive(const T* __restrict__ g, T (&x)[8]) {
    if (sizeof(T) == 8) {
        const ulonglong2* v = reinterpret_cast<const ulonglong2*>(g);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ulonglong2 t = v[k];
            x[2 * k] = (T)t.x;
            x[2 * k + 1] = (T)t.y;
        }
    } else {
        const uint4* v = reinterpret_cast<const uint4*>(g);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            uint4 t = v[k];
            x[4 * k] = (T)t.x;
            x[4 * k + 1] = (T)t.y;
            x[4 * k + 2] = (T)t.z;
            x[4 * k + 3] = (T)t.w;
        }
    }
}
template <class T>
NTT_DEVINL void store8_consecutive(T* __restrict__ g, const T (&x)[8]) {
    if (sizeof(T) == 8) {
        ulonglong2* v = reinterpret_cast<ulonglong2*>(g);
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = make_ulonglong2(x[2 * k], x[2 * k + 1]);
    } else {
        uint4* v = reinterpret_cast<uint4*>(g);
#pragma unroll
        for (int k = 0; k < 2; ++k) v[k] = make_uint4(x[4 * k], x[4 * k + 1], x[4 * k + 2], x[4 * k + 3]);
    }
}
// 8 consecutive elements <-> padded shared memory (never straddles a padding gap: 8 | 16)
template <class T>
NTT_DEVINL void lds8(const T* s, unsigned u, T (&x)[8]) {
    load8_consecutive(s + pad_index<T>(8 * u), x);
}
template <class T>
NTT_DEVINL void sts8(T* s, unsigned u, const T (&x)[8]) {
    store8_consecutive(s + pad_index<T>(8 * u), x);
}

// Core transform of one polynomial by N/8 cooperating threads.  `t` = thread index within the
// polynomial, `s` = this polynomial's padded shared-memory tile.  Within a pass every thread
// loads and stores the same 8 positions, so one CTA barrier per pass is enough.
template <class A, int LOGN>
NTT_DEVINL void fwd_from_regs(typename A::T (&x)[8], typename A::T* s, unsigned t,
                              const typename A::TW* __restrict__ tw, const typename A::Ctx& c) {
    // on entry x holds elements t + k*(N/8): exactly the first radix-8 tuple (stages 0..2)
    using S = FastShape<LOGN>;
#pragma unroll
    for (int pass = 0; pass < S::kRadix8Passes; ++pass) {
        const int stage = 3 * pass;
        const int log_t2 = LOGN - stage - 3;
        unsigned i = t >> log_t2, j = t & ((1u << log_t2) - 1u);
        unsigned base = (i << (log_t2 + 3)) + j;
        if (pass > 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] = s[pad_index<typename A::T>(base + ((unsigned)k << log_t2))];
        }
        tuple_ro<A, 3, 0, false>(x, tw, (1u << stage) + i, c);
#pragma unroll
        for (int k = 0; k < 8; ++k) s[pad_index<typename A::T>(base + ((unsigned)k << log_t2))] = x[k];
        __syncthreads();
    }
    lds8(s, t, x);
    last_pass<A, LOGN, false>(x, tw, t, c);
}

// inverse: x holds 8 consecutive elements 8t..8t+7 on entry, elements t + k*(N/8) on exit
template <class A, int LOGN>
NTT_DEVINL void inv_to_regs(typename A::T (&x)[8], typename A::T* s, unsigned t,
                            const typename A::TW* __restrict__ tw, const typename A::Ctx& c) {
    using S = FastShape<LOGN>;
    last_pass<A, LOGN, true>(x, tw, t, c);
    sts8(s, t, x);
#pragma unroll
    for (int pass = S::kRadix8Passes - 1; pass >= 0; --pass) {
        const int stage = 3 * pass;
        const int log_t2 = LOGN - stage - 3;
        unsigned i = t >> log_t2, j = t & ((1u << log_t2) - 1u);
        unsigned base = (i << (log_t2 + 3)) + j;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = s[pad_index<typename A::T>(base + ((unsigned)k << log_t2))];
        tuple_ro<A, 3, 0, true>(x, tw, (1u << stage) + i, c);
        if (pass > 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) s[pad_index<typename A::T>(base + ((unsigned)k << log_t2))] = x[k];
        }
    }
}

// ---- kernels ------------------------------------------------------------------------------
// blockDim = (N/8, POLYS): POLYS polynomials per CTA (more than one only for small N).
template <class A, int LOGN, int POLYS>
__global__ void __launch_bounds__(FastShape<LOGN>::kThreadsPerPoly* POLYS)
    ntt_fast_fwd_kernel(typename A::T* __restrict__ data, size_t batch,
                        const typename A::TW* __restrict__ tw, typename A::Ctx c) {
    using T = typename A::T;
    using S = FastShape<LOGN>;
    __shared__ __align__(16) T smem[POLYS * S::kPaddedElems];
    const unsigned t = threadIdx.x;
    size_t poly = (size_t)blockIdx.x * POLYS + threadIdx.y;
    // whole-CTA barriers below: out-of-range polynomials run on a clamped index and skip the store
    const bool live = poly < batch;
    if (!live) poly = batch - 1;
    T* g = data + (poly << LOGN);
    T* s = smem + threadIdx.y * S::kPaddedElems;
    T x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = g[t + k * S::kThreadsPerPoly];
    fwd_from_regs<A, LOGN>(x, s, t, tw, c);
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = A::fwd_fin(c, x[k]);
    if (live) store8_consecutive(g + 8 * t, x);
}

template <class A, int LOGN, int POLYS>
__global__ void __launch_bounds__(FastShape<LOGN>::kThreadsPerPoly* POLYS)
    ntt_fast_inv_kernel(typename A::T* __restrict__ data, size_t batch,
                        const typename A::TW* __restrict__ tw, typename A::Ctx c) {
    using T = typename A::T;
    using S = FastShape<LOGN>;
    __shared__ __align__(16) T smem[POLYS * S::kPaddedElems];
    const unsigned t = threadIdx.x;
    size_t poly = (size_t)blockIdx.x * POLYS + threadIdx.y;
    const bool live = poly < batch;
    if (!live) poly = batch - 1;
    T* g = data + (poly << LOGN);
    T* s = smem + threadIdx.y * S::kPaddedElems;
    T x[8];
    load8_consecutive(g + 8 * t, x);
    inv_to_regs<A, LOGN>(x, s, t, tw, c);
    if (live) {
#pragma unroll
        for (int k = 0; k < 8; ++k) g[t + k * S::kThreadsPerPoly] = A::inv_fin(c, x[k]);
    }
}

// Fused fwd -> pointwise multiply(-accumulate) -> inv, one pass over HBM (BASELINE C2 / the
// PBS external product shape):  out = inv(acc + fwd(lhs) * rhs).
template <class A, int LOGN, int POLYS>
__global__ void __launch_bounds__(FastShape<LOGN>::kThreadsPerPoly* POLYS)
    ntt_fast_fwd_mac_inv_kernel(typename A::T* __restrict__ out,
                                const typename A::T* __restrict__ lhs,
                                const typename A::T* __restrict__ rhs, size_t rhs_polys,
                                const typename A::T* __restrict__ acc, size_t acc_polys,
                                size_t batch, const typename A::TW* __restrict__ tw_fwd,
                                const typename A::TW* __restrict__ tw_inv, typename A::Ctx c) {
    using T = typename A::T;
    using S = FastShape<LOGN>;
    __shared__ __align__(16) T smem[POLYS * S::kPaddedElems];
    const unsigned t = threadIdx.x;
    size_t poly = (size_t)blockIdx.x * POLYS + threadIdx.y;
    const bool live = poly < batch;
    if (!live) poly = batch - 1;
    T* s = smem + threadIdx.y * S::kPaddedElems;
    T x[8];
    const T* gl = lhs + (poly << LOGN);
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = gl[t + k * S::kThreadsPerPoly];
    fwd_from_regs<A, LOGN>(x, s, t, tw_fwd, c);
    // pointwise step on the 8 consecutive NTT-domain coefficients this thread holds
    {
        T r[8];
        load8_consecutive(rhs + ((poly % rhs_polys) << LOGN) + 8 * t, r);
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = A::mul_full(c, A::fwd_fin(c, x[k]), r[k]);
        if (acc) {
            load8_consecutive(acc + ((poly % acc_polys) << LOGN) + 8 * t, r);
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] = A::add_full(c, x[k], r[k]);
        }
    }
    // no barrier needed: the inverse first writes the 8 positions this thread just read
    inv_to_regs<A, LOGN>(x, s, t, tw_inv, c);
    if (live) {
        T* go = out + (poly << LOGN);
#pragma unroll
        for (int k = 0; k < 8; ++k) go[t + k * S::kThreadsPerPoly] = A::inv_fin(c, x[k]);
    }
}

// Host-side dispatch (defined in ntt_fast_*.cu, one translation unit per modulus family).
// Returns false when (A, logn) has no fast kernel; the caller then uses the generic path.
template <class A>
bool fast_fwd(typename A::T* data, size_t batch, int logn, const typename A::TW* tw,
              const typename A::Ctx& c, cudaStream_t st);
template <class A>
bool fast_inv(typename A::T* data, size_t batch, int logn, const typename A::TW* tw,
              const typename A::Ctx& c, cudaStream_t st);
template <class A>
bool fast_fwd_mac_inv(typename A::T* out, const typename A::T* lhs, const typename A::T* rhs,
                      size_t rhs_polys, const typename A::T* acc, size_t acc_polys, size_t batch,
                      int logn, const typename A::TW* tw_fwd, const typename A::TW* tw_inv,
                      const typename A::Ctx& c, cudaStream_t st);

}  // namespace nttb200
