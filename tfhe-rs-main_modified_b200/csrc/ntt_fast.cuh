// Fast single-CTA NTT kernels with compile-time size (2^8 .. 2^12 coefficients).
//
// One thread owns 8 coefficients of each of PPT (1 or 2) polynomials; a polynomial is N/8 threads.
// Two polynomials per thread share every twiddle load and all index arithmetic.  The log2(N) radix-2 stages of
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
#include <cooperative_groups.h>

#include <type_traits>

#include "ntt_kernels.cuh"

#ifndef NTT_EXT_THREADS_PER_SM
#define NTT_EXT_THREADS_PER_SM 768
#endif
#ifndef NTT_FUSED_U32_THREADS_PER_SM
#define NTT_FUSED_U32_THREADS_PER_SM 1280
#endif
#ifndef NTT_FAST_U32_THREADS_PER_SM
#define NTT_FAST_U32_THREADS_PER_SM 1536
#endif

namespace nttb200 {
namespace cg = cooperative_groups;

// A "row" handled by one thread group may be one of the 2^depth contiguous sub-blocks of a longer
// polynomial whose first `depth` stages ran in global memory (the reference's depth-first
// recursion, generic_solinas.rs:1338-1386): its stage with 2^s groups then reads
// twid[(2^s << depth) + half * 2^s + i].  depth = 0 is the plain case.
struct SubPoly {
    unsigned depth, half;
    NTT_DEVINL unsigned base(int stage) const { return ((1u << stage) << depth) + (half << stage); }
};

template <int LOGN>
struct FastShape {
    static constexpr int kLastStages = (LOGN % 3 == 0) ? 3 : (LOGN % 3);
    static constexpr int kRadix8Passes = (LOGN - kLastStages) / 3;
    static constexpr int kThreadsPerPoly = (1 << LOGN) / 8;
    static constexpr int kPaddedElems = (1 << LOGN) + ((1 << LOGN) >> 3);  // 16 B per 128 B
};

// Resident CTAs per SM the register allocator is asked to allow.  64-bit families: 1024 threads per
// SM, i.e. 64 registers per thread (they fit with at most a few spilled words); four 256-thread CTAs
// per SM hide the load phase of one CTA behind the butterflies of the others.  32-bit families need
// 40-48 registers, so they are given 1536 threads per SM (six 256-thread CTAs; measured +8 % at
// n = 2048); the exact path for p >= 2^31 spills there and keeps 1024.
template <class A, int THREADS, int PPT = 2>
struct FastMinBlocks {
    // (u32, four polynomials per thread: 72 registers, three 256-thread CTAs per SM measured best,
    // profiles/r02_u32_kernel_variants.txt)
    static constexpr bool kU32 = sizeof(typename A::T) == 4 && !std::is_same<A, Wide32>::value;
    static constexpr int kThreadsPerSM = kU32 ? (PPT <= 2 ? NTT_FAST_U32_THREADS_PER_SM : 768) : 1024;
    static constexpr int value = kThreadsPerSM / THREADS > 0 ? kThreadsPerSM / THREADS : 1;
};

// The fused fwd -> pointwise -> inv kernel: 32-bit Shoup families are held to
// NTT_FUSED_U32_THREADS_PER_SM threads per SM, everything else to 1024 (64 registers: the 64-bit
// families then spill a few words but run four 256-thread CTAs per SM instead of two).
template <class A, int THREADS, int PPT = 2>
struct FusedMinBlocks {
    static constexpr int kThreadsPerSM =
        (sizeof(typename A::T) == 4 && !std::is_same<A, Wide32>::value) ? (PPT <= 2 ? NTT_FUSED_U32_THREADS_PER_SM : 768) : 1024;
    static constexpr int value = kThreadsPerSM / THREADS > 0 ? kThreadsPerSM / THREADS : 1;
};

// Whether the 8-byte twiddle records of a stage are fetched two per 128-bit load (ldg_tw_run).
// Measured: a gain everywhere except the forward 4096-point kernels (u32 -9 %, u64 -1.5 %).
template <int LOGN, bool INV>
struct FastPairLoads {
    static constexpr bool value = INV || LOGN <= 11;
};

// 16 bytes of padding after every 128 bytes: u64 -> a + 2*(a>>4), u32 -> a + 4*(a>>5), 16-byte slots -> a + (a>>3)
template <class T>
NTT_DEVINL constexpr unsigned pad_index(unsigned a) {
    return sizeof(T) == 16 ? a + (a >> 3) : sizeof(T) == 8 ? a + 2u * (a >> 4) : a + 4u * (a >> 5);
}

// Tuple addressing.  A radix-8 tuple's base is i*(8*t2) + j with j < t2; when 8*t2 is a multiple of
// the padding period (16 u64 / 32 u32 elements) the padding of base + k*t2 splits into pad(base)
// plus a compile-time constant, so the eight accesses of a pass use immediate offsets instead of
// per-element index arithmetic (`const_off` in the pass loops below).

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

// CNT consecutive table entries starting at an index that is a multiple of CNT.  8-byte records
// (u64 Montgomery twiddles, u32 Shoup pairs) are fetched two at a time with one 128-bit load: the
// stages of a pass whose groups differ from lane to lane then touch half as many L1 wavefronts.
template <class TW, int CNT, bool PAIR>
NTT_DEVINL void ldg_tw_run(const TW* __restrict__ p, TW (&w)[CNT]) {
    if constexpr (PAIR && sizeof(TW) == 8 && CNT >= 2) {
        static_assert(sizeof(ulonglong2) == 2 * sizeof(TW), "pair load");
        const ulonglong2* v = reinterpret_cast<const ulonglong2*>(p);
#pragma unroll
        for (int k = 0; k < CNT / 2; ++k) {
            ulonglong2 t = __ldg(v + k);
            if constexpr (sizeof(w[0]) == sizeof(unsigned long long)) {
                __builtin_memcpy(&w[2 * k], &t.x, 8);
                __builtin_memcpy(&w[2 * k + 1], &t.y, 8);
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < CNT; ++k) w[k] = ldg_tw(p + k);
    }
}

// radix-2^R butterflies on x[.][OFF .. OFF + 2^R) with twiddles fetched through the read-only
// path.  PPT polynomials per thread share every twiddle load and all index arithmetic.
// ENTRY_CANON (inverse only): the tuple's inputs are known canonical.  Within the tuple the
// Gentleman-Sande stage of distance d leaves sums at positions with bit d clear and canonical
// products at positions with bit d set, so which inputs of the next stage are canonical is
// static.
// one radix-2 stage (Q-th of R, 2^Q twiddles) of a tuple, twiddles already in registers
template <class A, int R, int OFF, bool INV, bool ENTRY_CANON, int PPT, int Q>
NTT_DEVINL void tuple_stage_regs(typename A::T (&x)[PPT][8], const typename A::TW* wq, const typename A::Ctx& c) {
    constexpr int d = 1 << (R - 1 - Q);
#pragma unroll
    for (int h = 0; h < (1 << Q); ++h) {
        const typename A::TW w = wq[h];
#pragma unroll
        for (int k = 0; k < d; ++k) {
            const int ja = OFF + h * 2 * d + k, jb = ja + d;  // positions inside the tuple
            if (!INV) {
#pragma unroll
                for (int pp = 0; pp < PPT; ++pp) A::fwd_bf(c, x[pp][ja], x[pp][jb], w);
            } else {
                const bool b_canon = (d == 1) ? ENTRY_CANON : (((jb - OFF) & (d >> 1)) != 0);
#pragma unroll
                for (int pp = 0; pp < PPT; ++pp) A::inv_bf(c, x[pp][ja], x[pp][jb], w, b_canon);
            }
        }
    }
}
// ... fetched through the read-only path
template <class A, int R, int OFF, bool INV, bool ENTRY_CANON, int PPT, int Q, bool PAIR>
NTT_DEVINL void tuple_stage(typename A::T (&x)[PPT][8], const typename A::TW* __restrict__ tw,
                            unsigned w0, const typename A::Ctx& c) {
    typename A::TW wq[1 << Q];
    ldg_tw_run<typename A::TW, (1 << Q), PAIR>(tw + (w0 << Q), wq);
    tuple_stage_regs<A, R, OFF, INV, ENTRY_CANON, PPT, Q>(x, wq, c);
}
template <class A, int R, int OFF, bool INV, bool ENTRY_CANON, int PPT, bool PAIR = true>
NTT_DEVINL void tuple_ro(typename A::T (&x)[PPT][8], const typename A::TW* __restrict__ tw,
                         unsigned w0, const typename A::Ctx& c) {
    if (!INV) {
        tuple_stage<A, R, OFF, INV, ENTRY_CANON, PPT, 0, PAIR>(x, tw, w0, c);
        if constexpr (R >= 2) tuple_stage<A, R, OFF, INV, ENTRY_CANON, PPT, 1, PAIR>(x, tw, w0, c);
        if constexpr (R >= 3) tuple_stage<A, R, OFF, INV, ENTRY_CANON, PPT, 2, PAIR>(x, tw, w0, c);
    } else {
        if constexpr (R >= 3) tuple_stage<A, R, OFF, INV, ENTRY_CANON, PPT, 2, PAIR>(x, tw, w0, c);
        if constexpr (R >= 2) tuple_stage<A, R, OFF, INV, ENTRY_CANON, PPT, 1, PAIR>(x, tw, w0, c);
        tuple_stage<A, R, OFF, INV, ENTRY_CANON, PPT, 0, PAIR>(x, tw, w0, c);
    }
}

// The "last" pass: 8 consecutive coefficients 8u .. 8u+7 of the polynomial, stages
// [LOGN - S, LOGN).  S = 3: one radix-8 tuple; S = 2: two radix-4 tuples; S = 1: four radix-2.
// `sub` locates this polynomial inside a longer one (see SubPoly).
template <class A, int LOGN, bool INV, bool ENTRY_CANON, int PPT>
NTT_DEVINL void last_pass(typename A::T (&x)[PPT][8], const typename A::TW* __restrict__ tw,
                          unsigned u, const typename A::Ctx& c, const SubPoly& sub) {
    constexpr int S = FastShape<LOGN>::kLastStages;
    constexpr bool PAIR = FastPairLoads<LOGN, INV>::value;
    const unsigned m = sub.base(LOGN - S);  // table index of group 0 of the first fused stage
    using TW = typename A::TW;
    if constexpr (S == 3) {
        tuple_ro<A, 3, 0, INV, ENTRY_CANON, PPT, PAIR>(x, tw, m + u, c);
    } else if constexpr (S == 2) {
        // Two radix-4 tuples whose twiddles are neighbours in the table (records m+2u, m+2u+1 of the first
        // stage, 2(m+2u) .. +3 of the second): fetched as one 16-byte and one 32-byte run per thread instead of
        // per tuple, which halves the sectors / wavefronts a warp touches for them.
        TW wa[2], wb[4];
        ldg_tw_run<TW, 2, true>(tw + (m + 2 * u), wa);
        ldg_tw_run<TW, 4, true>(tw + 2 * (m + 2 * u), wb);
        if (!INV) {
            tuple_stage_regs<A, 2, 0, INV, ENTRY_CANON, PPT, 0>(x, wa, c);
            tuple_stage_regs<A, 2, 4, INV, ENTRY_CANON, PPT, 0>(x, wa + 1, c);
            tuple_stage_regs<A, 2, 0, INV, ENTRY_CANON, PPT, 1>(x, wb, c);
            tuple_stage_regs<A, 2, 4, INV, ENTRY_CANON, PPT, 1>(x, wb + 2, c);
        } else {
            tuple_stage_regs<A, 2, 0, INV, ENTRY_CANON, PPT, 1>(x, wb, c);
            tuple_stage_regs<A, 2, 4, INV, ENTRY_CANON, PPT, 1>(x, wb + 2, c);
            tuple_stage_regs<A, 2, 0, INV, ENTRY_CANON, PPT, 0>(x, wa, c);
            tuple_stage_regs<A, 2, 4, INV, ENTRY_CANON, PPT, 0>(x, wa + 1, c);
        }
    } else {
        // four radix-2 butterflies on consecutive pairs: records m+4u .. m+4u+3 in one 32-byte run
        TW w[4];
        ldg_tw_run<TW, 4, true>(tw + (m + 4 * u), w);
        tuple_stage_regs<A, 1, 0, INV, ENTRY_CANON, PPT, 0>(x, w, c);
        tuple_stage_regs<A, 1, 2, INV, ENTRY_CANON, PPT, 0>(x, w + 1, c);
        tuple_stage_regs<A, 1, 4, INV, ENTRY_CANON, PPT, 0>(x, w + 2, c);
        tuple_stage_regs<A, 1, 6, INV, ENTRY_CANON, PPT, 0>(x, w + 3, c);
    }
}

// 8 consecutive elements <-> memory with 128-bit accesses
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

// Core transform of PPT polynomials by N/8 cooperating threads.  `t` = thread index within the
// polynomial, `s` = the exchange tile (PPT * kPaddedElems elements; its inner layout is private to
// these two functions).  Within a pass every thread loads and stores the same 8 positions, so one CTA
// barrier per pass is enough.
//
// Tile layout.  64-bit elements, or one polynomial per thread: PPT consecutive padded tiles.  Two
// 32-bit polynomials per thread: ONE tile of 64-bit slots, slot e = (poly0[e], poly1[e]), padded like a
// u64 polynomial -- every pass moves both polynomials with one LDS.64 / STS.64 per position instead of
// two 32-bit accesses (half the shared-memory instructions of the u32 kernels, same bank behaviour as
// the u64 kernels), and the last pass reads 8 consecutive slots with four LDS.128.
// Four 32-bit polynomials per thread: one tile of 16-byte slots (poly0[e] .. poly3[e]), one LDS.128 /
// STS.128 per position; with one slot of padding per eight slots the radix-8 passes (slot strides n/8,
// n/64, 4) and the last pass (eight consecutive slots per thread, one 128-bit access each) are free of
// bank conflicts per quarter warp.
template <class T, int PPT>
struct TileLayout {
    static constexpr bool kPaired = sizeof(T) == 4 && PPT == 2;
    static constexpr bool kQuad = sizeof(T) == 4 && PPT == 4;
    // Two 64-bit polynomials per thread as 16-byte slots (poly0[e], poly1[e]): implemented, measured at the
    // SASS level and NOT used -- a 128-bit access needs four consecutive registers, and the moves that line
    // them up (IMAD.MOV 116 -> 198 in the Solinas n = 2048 forward kernel) outweigh the 40 shared-memory
    // instructions saved (2548 against 2507 instructions).
    static constexpr bool kPair64 = false && sizeof(T) == 8 && PPT == 2;
    using Slot = typename std::conditional<kQuad || kPair64, uint4, typename std::conditional<kPaired, uint64_t, T>::type>::type;
};
template <class T, int PPT, int PADDED>
NTT_DEVINL void tile_load(const T* s, unsigned off, T (&x)[PPT][8], int k) {
    if constexpr (TileLayout<T, PPT>::kPair64) {
        ulonglong2 v = reinterpret_cast<const ulonglong2*>(s)[off];
        x[0][k] = v.x, x[1][k] = v.y;
    } else if constexpr (TileLayout<T, PPT>::kQuad) {
        uint4 v = reinterpret_cast<const uint4*>(s)[off];
        x[0][k] = v.x, x[1][k] = v.y, x[2][k] = v.z, x[3][k] = v.w;
    } else if constexpr (TileLayout<T, PPT>::kPaired) {
        uint2 v = reinterpret_cast<const uint2*>(s)[off];
        x[0][k] = v.x;
        x[1][k] = v.y;
    } else {
#pragma unroll
        for (int pp = 0; pp < PPT; ++pp) x[pp][k] = s[pp * PADDED + off];
    }
}
template <class T, int PPT, int PADDED>
NTT_DEVINL void tile_store(T* s, unsigned off, const T (&x)[PPT][8], int k) {
    if constexpr (TileLayout<T, PPT>::kPair64) {
        reinterpret_cast<ulonglong2*>(s)[off] = make_ulonglong2(x[0][k], x[1][k]);
    } else if constexpr (TileLayout<T, PPT>::kQuad) {
        reinterpret_cast<uint4*>(s)[off] = make_uint4(x[0][k], x[1][k], x[2][k], x[3][k]);
    } else if constexpr (TileLayout<T, PPT>::kPaired) {
        reinterpret_cast<uint2*>(s)[off] = make_uint2(x[0][k], x[1][k]);
    } else {
#pragma unroll
        for (int pp = 0; pp < PPT; ++pp) s[pp * PADDED + off] = x[pp][k];
    }
}
// the 8 consecutive elements 8t .. 8t+7 of every polynomial <-> the tile
template <class T, int PPT, int PADDED>
NTT_DEVINL void tile_load8(const T* s, unsigned t, T (&x)[PPT][8]) {
    if constexpr (TileLayout<T, PPT>::kPair64) {
        const ulonglong2* v = reinterpret_cast<const ulonglong2*>(s) + pad_index<uint4>(8 * t);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            ulonglong2 q = v[k];
            x[0][k] = q.x, x[1][k] = q.y;
        }
    } else if constexpr (TileLayout<T, PPT>::kQuad) {
        const uint4* v = reinterpret_cast<const uint4*>(s) + pad_index<uint4>(8 * t);  // 8 slots, no padding inside
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint4 q = v[k];
            x[0][k] = q.x, x[1][k] = q.y, x[2][k] = q.z, x[3][k] = q.w;
        }
    } else if constexpr (TileLayout<T, PPT>::kPaired) {
        const uint4* v = reinterpret_cast<const uint4*>(reinterpret_cast<const uint2*>(s) + pad_index<uint64_t>(8 * t));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint4 q = v[k];
            x[0][2 * k] = q.x, x[1][2 * k] = q.y, x[0][2 * k + 1] = q.z, x[1][2 * k + 1] = q.w;
        }
    } else {
#pragma unroll
        for (int pp = 0; pp < PPT; ++pp) load8_consecutive(s + pp * PADDED + pad_index<T>(8 * t), x[pp]);
    }
}
template <class T, int PPT, int PADDED>
NTT_DEVINL void tile_store8(T* s, unsigned t, const T (&x)[PPT][8]) {
    if constexpr (TileLayout<T, PPT>::kPair64) {
        ulonglong2* v = reinterpret_cast<ulonglong2*>(s) + pad_index<uint4>(8 * t);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = make_ulonglong2(x[0][k], x[1][k]);
    } else if constexpr (TileLayout<T, PPT>::kQuad) {
        uint4* v = reinterpret_cast<uint4*>(s) + pad_index<uint4>(8 * t);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = make_uint4(x[0][k], x[1][k], x[2][k], x[3][k]);
    } else if constexpr (TileLayout<T, PPT>::kPaired) {
        uint4* v = reinterpret_cast<uint4*>(reinterpret_cast<uint2*>(s) + pad_index<uint64_t>(8 * t));
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = make_uint4(x[0][2 * k], x[1][2 * k], x[0][2 * k + 1], x[1][2 * k + 1]);
    } else {
#pragma unroll
        for (int pp = 0; pp < PPT; ++pp) store8_consecutive(s + pp * PADDED + pad_index<T>(8 * t), x[pp]);
    }
}

template <class A, int LOGN, int PPT>
NTT_DEVINL void fwd_from_regs(typename A::T (&x)[PPT][8], typename A::T* s, unsigned t,
                              const typename A::TW* __restrict__ tw, const typename A::Ctx& c,
                              const SubPoly& sub) {
    // on entry x holds elements t + k*(N/8): exactly the first radix-8 tuple (stages 0..2)
    using S = FastShape<LOGN>;
    using T = typename A::T;
    using Slot = typename TileLayout<T, PPT>::Slot;
#pragma unroll
    for (int pass = 0; pass < S::kRadix8Passes; ++pass) {
        const int stage = 3 * pass;
        const int log_t2 = LOGN - stage - 3;
        unsigned i = t >> log_t2, j = t & ((1u << log_t2) - 1u);
        unsigned base = (i << (log_t2 + 3)) + j;
        // (pass and k are unrolled, so these fold to constants)
        const bool const_off = ((8u << log_t2) % (128u / sizeof(Slot))) == 0;
        const unsigned pbase = pad_index<Slot>(base);
        if (pass > 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                unsigned off = const_off ? pbase + pad_index<Slot>((unsigned)k << log_t2)
                                         : pad_index<Slot>(base + ((unsigned)k << log_t2));
                tile_load<T, PPT, S::kPaddedElems>(s, off, x, k);
            }
        }
        tuple_ro<A, 3, 0, false, false, PPT, FastPairLoads<LOGN, false>::value>(x, tw, sub.base(stage) + i, c);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            unsigned off = const_off ? pbase + pad_index<Slot>((unsigned)k << log_t2)
                                     : pad_index<Slot>(base + ((unsigned)k << log_t2));
            tile_store<T, PPT, S::kPaddedElems>(s, off, x, k);
        }
        __syncthreads();
    }
    tile_load8<T, PPT, S::kPaddedElems>(s, t, x);
    last_pass<A, LOGN, false, false, PPT>(x, tw, t, c, sub);
}

// inverse: x holds 8 consecutive elements 8t..8t+7 on entry (canonical: fresh input or the
// output of a pointwise product), elements t + k*(N/8) on exit
template <class A, int LOGN, int PPT>
NTT_DEVINL void inv_to_regs(typename A::T (&x)[PPT][8], typename A::T* s, unsigned t,
                            const typename A::TW* __restrict__ tw, const typename A::Ctx& c,
                            const SubPoly& sub) {
    using S = FastShape<LOGN>;
    using T = typename A::T;
    using Slot = typename TileLayout<T, PPT>::Slot;
    last_pass<A, LOGN, true, true, PPT>(x, tw, t, c, sub);
    tile_store8<T, PPT, S::kPaddedElems>(s, t, x);
#pragma unroll
    for (int pass = S::kRadix8Passes - 1; pass >= 0; --pass) {
        const int stage = 3 * pass;
        const int log_t2 = LOGN - stage - 3;
        unsigned i = t >> log_t2, j = t & ((1u << log_t2) - 1u);
        unsigned base = (i << (log_t2 + 3)) + j;
        // (pass and k are unrolled, so these fold to constants)
        const bool const_off = ((8u << log_t2) % (128u / sizeof(Slot))) == 0;
        const unsigned pbase = pad_index<Slot>(base);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            unsigned off = const_off ? pbase + pad_index<Slot>((unsigned)k << log_t2)
                                     : pad_index<Slot>(base + ((unsigned)k << log_t2));
            tile_load<T, PPT, S::kPaddedElems>(s, off, x, k);
        }
        tuple_ro<A, 3, 0, true, false, PPT, FastPairLoads<LOGN, true>::value>(x, tw, sub.base(stage) + i, c);
        if (pass > 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                unsigned off = const_off ? pbase + pad_index<Slot>((unsigned)k << log_t2)
                                         : pad_index<Slot>(base + ((unsigned)k << log_t2));
                tile_store<T, PPT, S::kPaddedElems>(s, off, x, k);
            }
        }
    }
}

// ---- kernels ------------------------------------------------------------------------------
// blockDim = (N/8, POLYS); each thread group works on PPT rows at once, so a CTA covers
// POLYS*PPT rows.  `rows` counts rows of length N; with depth > 0 row r is sub-block
// (r mod 2^depth) of polynomial r >> depth; a thread group then pairs the same sub-block of PPT
// consecutive polynomials, which share their twiddles.  Out-of-range rows are clamped and not
// stored (whole-CTA barriers inside).
// SUB = false: the rows are whole polynomials (depth == 0, the common case): the twiddle bases and the
// row mapping fold to constants (7-11 % fewer instructions in the u32 kernels, 2-3 % in the u64 ones).
template <class A, int LOGN, int POLYS, int PPT, bool SUB = true>
__global__ void __launch_bounds__(FastShape<LOGN>::kThreadsPerPoly* POLYS,
                                  FastMinBlocks<A, FastShape<LOGN>::kThreadsPerPoly * POLYS, PPT>::value)
    ntt_fast_fwd_kernel(typename A::T* __restrict__ data, size_t rows, unsigned depth_arg,
                        const typename A::TW* __restrict__ tw, typename A::Ctx c) {
    using T = typename A::T;
    using S = FastShape<LOGN>;
    extern __shared__ __align__(16) unsigned char fast_smem_raw[];  // POLYS * PPT padded tiles
    T* smem = reinterpret_cast<T*>(fast_smem_raw);
    const unsigned depth = SUB ? depth_arg : 0u;
    const unsigned t = threadIdx.x;
    // thread group `grp` works on sub-block `half` of PPT consecutive polynomials (same twiddles)
    const size_t grp = (size_t)blockIdx.x * POLYS + threadIdx.y;
    const unsigned half = (unsigned)grp & ((1u << depth) - 1u);
    const SubPoly sub{depth, half};
    T* s = smem + threadIdx.y * PPT * S::kPaddedElems;
    T x[PPT][8];
    T* g[PPT];
    bool live[PPT];
#pragma unroll
    for (int pp = 0; pp < PPT; ++pp) {
        size_t row = ((((grp >> depth) * PPT + pp)) << depth) + half;
        live[pp] = row < rows;
        g[pp] = data + ((live[pp] ? row : rows - 1) << LOGN);
#pragma unroll
        for (int k = 0; k < 8; ++k) x[pp][k] = g[pp][t + k * S::kThreadsPerPoly];
    }
    fwd_from_regs<A, LOGN, PPT>(x, s, t, tw, c, sub);
#pragma unroll
    for (int pp = 0; pp < PPT; ++pp) {
#pragma unroll
        for (int k = 0; k < 8; ++k) x[pp][k] = A::fwd_fin(c, x[pp][k]);
        if (live[pp]) store8_consecutive(g[pp] + 8 * t, x[pp]);
    }
}

template <class A, int LOGN, int POLYS, int PPT, bool SUB = true>
__global__ void __launch_bounds__(FastShape<LOGN>::kThreadsPerPoly* POLYS,
                                  FastMinBlocks<A, FastShape<LOGN>::kThreadsPerPoly * POLYS, PPT>::value)
    ntt_fast_inv_kernel(typename A::T* __restrict__ data, size_t rows, unsigned depth_arg,
                        const typename A::TW* __restrict__ tw, typename A::Ctx c) {
    using T = typename A::T;
    using S = FastShape<LOGN>;
    extern __shared__ __align__(16) unsigned char fast_smem_raw[];  // POLYS * PPT padded tiles
    T* smem = reinterpret_cast<T*>(fast_smem_raw);
    const unsigned depth = SUB ? depth_arg : 0u;
    const unsigned t = threadIdx.x;
    // thread group `grp` works on sub-block `half` of PPT consecutive polynomials (same twiddles)
    const size_t grp = (size_t)blockIdx.x * POLYS + threadIdx.y;
    const unsigned half = (unsigned)grp & ((1u << depth) - 1u);
    const SubPoly sub{depth, half};
    T* s = smem + threadIdx.y * PPT * S::kPaddedElems;
    T x[PPT][8];
    T* g[PPT];
    bool live[PPT];
#pragma unroll
    for (int pp = 0; pp < PPT; ++pp) {
        size_t row = ((((grp >> depth) * PPT + pp)) << depth) + half;
        live[pp] = row < rows;
        g[pp] = data + ((live[pp] ? row : rows - 1) << LOGN);
        load8_consecutive(g[pp] + 8 * t, x[pp]);
    }
    inv_to_regs<A, LOGN, PPT>(x, s, t, tw, c, sub);
    // after the last radix-8 pass positions 0..3 hold sums, 4..7 canonical-range products; with
    // depth > 0 the lazy values feed the global passes, which canonicalise at the end
#pragma unroll
    for (int pp = 0; pp < PPT; ++pp) {
        if (!live[pp]) continue;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            T v = x[pp][k];
            if (depth == 0) v = (k < 4) ? A::inv_fin(c, v) : A::inv_fin_prod(c, v);
            g[pp][t + k * S::kThreadsPerPoly] = v;
        }
    }
}

// Index of polynomial `poly` into an operand that holds `count` polynomials reused cyclically by a batch of
// `batch`.  The common cases cost nothing: an operand per polynomial (count == batch: the index itself) and
// one operand for all (count == 1); only a truly periodic operand pays a modulo, 32-bit when it can be --
// a 64-bit `%` is a ~100-instruction subroutine call, and the fused kernel used to run four of them per thread.
NTT_DEVINL size_t shared_index(size_t poly, size_t count, size_t batch) {
    if (count >= batch) return poly;
    if (count == 1) return 0;
    if (batch <= 0xFFFFFFFFull) return (size_t)((unsigned)poly % (unsigned)count);
    return poly % count;
}

// Fused fwd -> pointwise multiply(-accumulate) -> inv, one pass over HBM (BASELINE C2 / the
// PBS external product shape):  out = inv(acc + fwd(lhs) * rhs).
template <class A, int LOGN, int POLYS, int PPT>
__global__ void __launch_bounds__(FastShape<LOGN>::kThreadsPerPoly* POLYS,
                                  FusedMinBlocks<A, FastShape<LOGN>::kThreadsPerPoly * POLYS, PPT>::value)
    ntt_fast_fwd_mac_inv_kernel(typename A::T* __restrict__ out,
                                const typename A::T* __restrict__ lhs,
                                const typename A::T* __restrict__ rhs, size_t rhs_polys,
                                const typename A::T* __restrict__ acc, size_t acc_polys,
                                size_t batch, const typename A::TW* __restrict__ tw_fwd,
                                const typename A::TW* __restrict__ tw_inv, typename A::Ctx c) {
    using T = typename A::T;
    using S = FastShape<LOGN>;
    extern __shared__ __align__(16) unsigned char fast_smem_raw[];  // POLYS * PPT padded tiles
    T* smem = reinterpret_cast<T*>(fast_smem_raw);
    const unsigned t = threadIdx.x;
    const size_t poly0 = ((size_t)blockIdx.x * POLYS + threadIdx.y) * PPT;
    const SubPoly sub{0u, 0u};
    T* s = smem + threadIdx.y * PPT * S::kPaddedElems;
    T x[PPT][8];
    size_t poly[PPT];
    bool live[PPT];
#pragma unroll
    for (int pp = 0; pp < PPT; ++pp) {
        live[pp] = poly0 + pp < batch;
        poly[pp] = live[pp] ? poly0 + pp : batch - 1;
        const T* gl = lhs + (poly[pp] << LOGN);
#pragma unroll
        for (int k = 0; k < 8; ++k) x[pp][k] = gl[t + k * S::kThreadsPerPoly];
    }
    fwd_from_regs<A, LOGN, PPT>(x, s, t, tw_fwd, c, sub);
    // pointwise step on the 8 consecutive NTT-domain coefficients this thread holds
#pragma unroll
    for (int pp = 0; pp < PPT; ++pp) {
        T r[8];
        load8_consecutive(rhs + (shared_index(poly[pp], rhs_polys, batch) << LOGN) + 8 * t, r);
#pragma unroll
        for (int k = 0; k < 8; ++k) x[pp][k] = A::mul_full(c, A::fwd_fin(c, x[pp][k]), r[k]);
        if (acc) {
            load8_consecutive(acc + (shared_index(poly[pp], acc_polys, batch) << LOGN) + 8 * t, r);
#pragma unroll
            for (int k = 0; k < 8; ++k) x[pp][k] = A::add_full(c, x[pp][k], r[k]);
        }
    }
    // no barrier needed: the inverse first writes the 8 positions this thread just read
    inv_to_regs<A, LOGN, PPT>(x, s, t, tw_inv, c, sub);
#pragma unroll
    for (int pp = 0; pp < PPT; ++pp) {
        if (!live[pp]) continue;
        T* go = out + (poly[pp] << LOGN);
#pragma unroll
        for (int k = 0; k < 8; ++k)
            go[t + k * S::kThreadsPerPoly] = (k < 4) ? A::inv_fin(c, x[pp][k]) : A::inv_fin_prod(c, x[pp][k]);
    }
}

// reused pointwise operand -> the family's pointwise form (Montgomery form for the 64-bit families)
template <class A>
__global__ void pw_form_kernel(typename A::T* __restrict__ out, const typename A::T* __restrict__ in,
                               size_t total, typename A::Ctx c) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        out[i] = A::pw_form(c, in[i]);
}

// External-product core (the NTT-PBS shape, tfhe ntt64_pbs.rs:598-661): for every batch item b
//   out[b][c] = inv( sum_r fwd(in[b][r]) (*) ggsw[r][c] ),   r < rows, c < COLS
// `ggsw` (rows x COLS NTT-domain polynomials) is shared by the whole batch and stays in L2.
// One CTA per batch item: rows forward transforms, rows*COLS pointwise multiply-accumulates
// into register accumulators, COLS inverse transforms; HBM traffic is rows*n*w in, COLS*n*w out.
// Equivalent to the reference's sequence Plan::fwd / Plan::mul_accumulate / Plan::inv.
template <class A, int LOGN, int COLS>
__global__ void __launch_bounds__(FastShape<LOGN>::kThreadsPerPoly,
                                  (NTT_EXT_THREADS_PER_SM / FastShape<LOGN>::kThreadsPerPoly > 0
                                       ? NTT_EXT_THREADS_PER_SM / FastShape<LOGN>::kThreadsPerPoly : 1))
    ntt_fast_ext_product_kernel(typename A::T* __restrict__ out, const typename A::T* __restrict__ in,
                                const typename A::T* __restrict__ ggsw, unsigned rows,
                                const typename A::TW* __restrict__ tw_fwd,
                                const typename A::TW* __restrict__ tw_inv, typename A::Ctx c) {
    using T = typename A::T;
    using S = FastShape<LOGN>;
    __shared__ __align__(16) T smem[S::kPaddedElems];
    const unsigned t = threadIdx.x;
    const size_t b = blockIdx.x;
    const SubPoly sub{0u, 0u};
    T acc[COLS][8];
#pragma unroll
    for (int cc = 0; cc < COLS; ++cc)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[cc][k] = 0;
    for (unsigned r = 0; r < rows; ++r) {
        T x[1][8];
        const T* gi = in + ((b * rows + r) << LOGN);
#pragma unroll
        for (int k = 0; k < 8; ++k) x[0][k] = gi[t + k * S::kThreadsPerPoly];
        fwd_from_regs<A, LOGN, 1>(x, smem, t, tw_fwd, c, sub);
        if (!A::kPwLazyIn) {
#pragma unroll
            for (int k = 0; k < 8; ++k) x[0][k] = A::fwd_fin(c, x[0][k]);
        }
#pragma unroll
        for (int cc = 0; cc < COLS; ++cc) {
            T g[8];  // `ggsw` is in the family's pointwise form (A::pw_form, applied by the launcher)
            load8_consecutive(ggsw + (((size_t)r * COLS + cc) << LOGN) + 8 * t, g);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[cc][k] = A::acc_add(c, acc[cc][k], A::mul_pw(c, x[0][k], g[k]));
        }
        __syncthreads();  // the tile is reused by the next transform
    }
#pragma unroll
    for (int cc = 0; cc < COLS; ++cc) {
        T x[1][8];
#pragma unroll
        for (int k = 0; k < 8; ++k) x[0][k] = A::acc_fin(c, acc[cc][k]);
        inv_to_regs<A, LOGN, 1>(x, smem, t, tw_inv, c, sub);
        T* go = out + ((b * COLS + cc) << LOGN);
#pragma unroll
        for (int k = 0; k < 8; ++k)
            go[t + k * S::kThreadsPerPoly] = (k < 4) ? A::inv_fin(c, x[0][k]) : A::inv_fin_prod(c, x[0][k]);
        __syncthreads();
    }
}

// ---- polynomials of 2^13 / 2^14 coefficients: one thread-block cluster per polynomial -------------
// n = 8 * 2^LOGSUB.  A cluster of eight CTAs (2^LOGSUB / 8 threads each) holds the polynomial in
// its eight shared-memory tiles, so it crosses HBM once in each direction (the two-launch path
// ntt_global_pass_kernel + ntt_fast_*_kernel crosses it twice).
//   forward: every thread of the cluster loads the radix-8 tuple {j + k * n/8} from global memory
//            (stages 0..2 of the reference's sweep, twid[1], twid[2..3], twid[4..7]) and writes
//            element k into the tile of CTA k through distributed shared memory; after one cluster
//            barrier CTA r runs the remaining stages on sub-block r exactly like the single-CTA
//            kernel (SubPoly{3, r}).
//   inverse: the mirror image; the last radix-8 tuple gathers its inputs from the eight tiles.
template <class A, int LOGSUB>
__global__ void __cluster_dims__(8, 1, 1)
    __launch_bounds__(FastShape<LOGSUB>::kThreadsPerPoly, FastMinBlocks<A, FastShape<LOGSUB>::kThreadsPerPoly>::value)
        ntt_cluster8_fwd_kernel(typename A::T* __restrict__ data, const typename A::TW* __restrict__ tw,
                                typename A::Ctx c) {
    using T = typename A::T;
    using S = FastShape<LOGSUB>;
    constexpr unsigned TPP = S::kThreadsPerPoly, SUB = 1u << LOGSUB;
    extern __shared__ __align__(16) unsigned char fast_smem_raw[];
    T* tile = reinterpret_cast<T*>(fast_smem_raw);
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned r = cluster.block_rank(), t = threadIdx.x;
    T* g = data + ((size_t)(blockIdx.x >> 3) << (LOGSUB + 3));
    T x[1][8];
    const unsigned j = r * TPP + t;
#pragma unroll
    for (int k = 0; k < 8; ++k) x[0][k] = g[j + k * SUB];
    cluster.sync();  // every CTA of the cluster is resident before its shared memory is written
    tuple_ro<A, 3, 0, false, false, 1, false>(x, tw, 1u, c);
#pragma unroll
    for (int k = 0; k < 8; ++k) cluster.map_shared_rank(tile, k)[pad_index<T>(j)] = x[0][k];
    cluster.sync();
    const SubPoly sub{3u, r};
#pragma unroll
    for (int k = 0; k < 8; ++k) x[0][k] = tile[pad_index<T>(t + k * TPP)];
    fwd_from_regs<A, LOGSUB, 1>(x, tile, t, tw, c, sub);
#pragma unroll
    for (int k = 0; k < 8; ++k) x[0][k] = A::fwd_fin(c, x[0][k]);
    store8_consecutive(g + r * SUB + 8 * t, x[0]);
}

template <class A, int LOGSUB>
__global__ void __cluster_dims__(8, 1, 1)
    __launch_bounds__(FastShape<LOGSUB>::kThreadsPerPoly, FastMinBlocks<A, FastShape<LOGSUB>::kThreadsPerPoly>::value)
        ntt_cluster8_inv_kernel(typename A::T* __restrict__ data, const typename A::TW* __restrict__ tw,
                                typename A::Ctx c) {
    using T = typename A::T;
    using S = FastShape<LOGSUB>;
    constexpr unsigned TPP = S::kThreadsPerPoly, SUB = 1u << LOGSUB;
    extern __shared__ __align__(16) unsigned char fast_smem_raw[];
    T* tile = reinterpret_cast<T*>(fast_smem_raw);
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned r = cluster.block_rank(), t = threadIdx.x;
    T* g = data + ((size_t)(blockIdx.x >> 3) << (LOGSUB + 3));
    T x[1][8];
    load8_consecutive(g + r * SUB + 8 * t, x[0]);
    const SubPoly sub{3u, r};
    inv_to_regs<A, LOGSUB, 1>(x, tile, t, tw, c, sub);
    // x holds the lazy values of elements t + k*TPP of sub-block r (the slots this thread owns)
#pragma unroll
    for (int k = 0; k < 8; ++k) tile[pad_index<T>(t + k * TPP)] = x[0][k];
    cluster.sync();
    const unsigned j = r * TPP + t;
#pragma unroll
    for (int k = 0; k < 8; ++k) x[0][k] = cluster.map_shared_rank(tile, k)[pad_index<T>(j)];
    tuple_ro<A, 3, 0, true, false, 1, false>(x, tw, 1u, c);
#pragma unroll
    for (int k = 0; k < 8; ++k) g[j + k * SUB] = (k < 4) ? A::inv_fin(c, x[0][k]) : A::inv_fin_prod(c, x[0][k]);
    cluster.sync();  // no CTA leaves while its tile may still be read
}

// Host-side dispatch (defined in ntt_fast_*.cu, one translation unit per modulus family).
// Returns false when (A, logn) has no fast kernel; the caller then uses the generic path.
// `rows` rows of 2^logn coefficients; depth as in SubPoly.
template <class A>
bool fast_fwd(typename A::T* data, size_t rows, int logn, unsigned depth, const typename A::TW* tw,
              const typename A::Ctx& c, cudaStream_t st);
template <class A>
bool fast_inv(typename A::T* data, size_t rows, int logn, unsigned depth, const typename A::TW* tw,
              const typename A::Ctx& c, cudaStream_t st);
// one cluster of eight CTAs per polynomial, logn = 13 or 14 (false: no such kernel)
template <class A>
bool fast_cluster_fwd(typename A::T* data, size_t polys, int logn, const typename A::TW* tw,
                      const typename A::Ctx& c, cudaStream_t st);
template <class A>
bool fast_cluster_inv(typename A::T* data, size_t polys, int logn, const typename A::TW* tw,
                      const typename A::Ctx& c, cudaStream_t st);
template <class A>
bool fast_ext_product(typename A::T* out, const typename A::T* in, const typename A::T* ggsw,
                      unsigned rows, unsigned cols, size_t batch, int logn,
                      const typename A::TW* tw_fwd, const typename A::TW* tw_inv,
                      const typename A::Ctx& c, cudaStream_t st);
template <class A>
bool fast_fwd_mac_inv(typename A::T* out, const typename A::T* lhs, const typename A::T* rhs,
                      size_t rhs_polys, const typename A::T* acc, size_t acc_polys, size_t batch,
                      int logn, const typename A::TW* tw_fwd, const typename A::TW* tw_inv,
                      const typename A::Ctx& c, cudaStream_t st);

}  // namespace nttb200
