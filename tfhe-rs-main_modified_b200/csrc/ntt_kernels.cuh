// Generic (any n, any modulus class) batched negacyclic NTT kernels for sm_100a.
//
// Transform structure is the reference's (tfhe-ntt/src/prime64/generic_solinas.rs:449-514,
// shoup.rs:544-615, :1306-1377): forward = Cooley-Tukey, natural order in, bit-reversed
// order out, stage with m groups reads twid[m + i]; inverse = Gentleman-Sande with inv_twid.
// Here 1..3 consecutive stages are fused into one register-resident radix-2/4/8 pass and a
// polynomial ("row") lives in shared memory between passes.  Rows longer than one CTA's
// shared memory are first cut down by strided global-memory passes (the same butterflies the
// reference's depth-first recursion performs above RECURSION_THRESHOLD,
// generic_solinas.rs:1338-1386: twiddle index (m << depth) + half * m + i).
#pragma once
#include "ntt_arith.cuh"

namespace nttb200 {

// One radix-2^R register pass, forward: stages q = 0..R-1; w0 is the table index of the
// first stage's twiddle for this tuple.
template <class A, int R>
NTT_DEVINL void fwd_tuple(typename A::T (&x)[1 << R], const typename A::TW* __restrict__ tw,
                          size_t w0, const typename A::Ctx& c) {
#pragma unroll
    for (int q = 0; q < R; ++q) {
        const int d = 1 << (R - 1 - q);
#pragma unroll
        for (int h = 0; h < (1 << q); ++h) {
            typename A::TW w = tw[(w0 << q) + h];
#pragma unroll
            for (int k = 0; k < d; ++k) A::fwd_bf(c, x[h * 2 * d + k], x[h * 2 * d + k + d], w);
        }
    }
}
// inverse: stages q = R-1..0 (smallest distance first)
template <class A, int R>
NTT_DEVINL void inv_tuple(typename A::T (&x)[1 << R], const typename A::TW* __restrict__ tw,
                          size_t w0, const typename A::Ctx& c) {
#pragma unroll
    for (int q = R - 1; q >= 0; --q) {
        const int d = 1 << (R - 1 - q);
#pragma unroll
        for (int h = 0; h < (1 << q); ++h) {
            typename A::TW w = tw[(w0 << q) + h];
#pragma unroll
            for (int k = 0; k < d; ++k) A::inv_bf(c, x[h * 2 * d + k], x[h * 2 * d + k + d], w);
        }
    }
}

// One pass over `rows` rows of length 2^log_row held contiguously at `s` (shared or global),
// fusing stages [stage, stage+R) of each row.
template <class A, int R, bool INV>
NTT_DEVINL void pass_over_rows(typename A::T* s, unsigned rows, unsigned row0, int log_row,
                               int depth, int stage, const typename A::TW* __restrict__ tw,
                               const typename A::Ctx& c, bool finalize, unsigned tid,
                               unsigned nthreads) {
    using T = typename A::T;
    const int log_t2 = log_row - stage - R;           // distance of the last fused stage
    const unsigned tuples_per_row = 1u << (log_row - R);
    const unsigned total = rows << (log_row - R);
    const unsigned half_mask = (1u << depth) - 1u;
    for (unsigned o = tid; o < total; o += nthreads) {
        unsigned r = o >> (log_row - R);
        unsigned ot = o & (tuples_per_row - 1);
        unsigned i = ot >> log_t2;
        unsigned j = ot & ((1u << log_t2) - 1u);
        T* base = s + ((size_t)r << log_row) + ((size_t)i << (log_t2 + R)) + j;
        size_t half = (row0 + r) & half_mask;
        size_t w0 = (((size_t)1 << stage) << depth) + (half << stage) + i;
        T x[1 << R];
#pragma unroll
        for (int k = 0; k < (1 << R); ++k) x[k] = base[(size_t)k << log_t2];
        if (INV)
            inv_tuple<A, R>(x, tw, w0, c);
        else
            fwd_tuple<A, R>(x, tw, w0, c);
        if (finalize) {
#pragma unroll
            for (int k = 0; k < (1 << R); ++k) x[k] = INV ? A::inv_fin(c, x[k]) : A::fwd_fin(c, x[k]);
        }
#pragma unroll
        for (int k = 0; k < (1 << R); ++k) base[(size_t)k << log_t2] = x[k];
    }
}

// Rows kernel: each CTA stages `rows_per_cta` rows in shared memory and runs all log_row
// stages of each.  depth > 0 means the rows are the 2^depth sub-blocks of longer polynomials
// whose first `depth` stages run in ntt_global_pass_kernel.
template <class A, bool INV>
__global__ void __launch_bounds__(512) ntt_rows_kernel(typename A::T* __restrict__ data, size_t num_rows, int log_row,
                                int depth, const typename A::TW* __restrict__ tw,
                                typename A::Ctx c, unsigned rows_per_cta, int finalize) {
    using T = typename A::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* s = reinterpret_cast<T*>(smem_raw);
    const unsigned tid = threadIdx.x, nth = blockDim.x;
    size_t row0 = (size_t)blockIdx.x * rows_per_cta;
    if (row0 >= num_rows) return;
    unsigned rows = (unsigned)((num_rows - row0 < rows_per_cta) ? (num_rows - row0) : rows_per_cta);
    const size_t elems = (size_t)rows << log_row;
    T* g = data + (row0 << log_row);

    // global -> shared, 128-bit when the base is aligned (rows are multiples of 128 B)
    constexpr unsigned V = 16 / sizeof(T);
    if ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) {
        const uint4* gv = reinterpret_cast<const uint4*>(g);
        uint4* sv = reinterpret_cast<uint4*>(s);
        for (size_t v = tid; v < elems / V; v += nth) sv[v] = gv[v];
    } else {
        for (size_t e = tid; e < elems; e += nth) s[e] = g[e];
    }
    __syncthreads();

    const int r0 = log_row % 3;
    const unsigned row0u = (unsigned)(row0 & 0xFFFFFFFFu);
    if (!INV) {
        int stage = 0;
        if (r0 == 1)
            pass_over_rows<A, 1, false>(s, rows, row0u, log_row, depth, 0, tw, c,
                                        finalize && log_row == 1, tid, nth);
        else if (r0 == 2)
            pass_over_rows<A, 2, false>(s, rows, row0u, log_row, depth, 0, tw, c,
                                        finalize && log_row == 2, tid, nth);
        if (r0) __syncthreads();
        for (stage = r0; stage < log_row; stage += 3) {
            pass_over_rows<A, 3, false>(s, rows, row0u, log_row, depth, stage, tw, c,
                                        finalize && stage + 3 == log_row, tid, nth);
            __syncthreads();
        }
    } else {
        for (int stage = log_row - 3; stage >= r0; stage -= 3) {
            pass_over_rows<A, 3, true>(s, rows, row0u, log_row, depth, stage, tw, c,
                                       finalize && stage == 0, tid, nth);
            __syncthreads();
        }
        if (r0 == 1)
            pass_over_rows<A, 1, true>(s, rows, row0u, log_row, depth, 0, tw, c, finalize, tid, nth);
        else if (r0 == 2)
            pass_over_rows<A, 2, true>(s, rows, row0u, log_row, depth, 0, tw, c, finalize, tid, nth);
        if (r0) __syncthreads();
    }

    if ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) {
        uint4* gv = reinterpret_cast<uint4*>(g);
        const uint4* sv = reinterpret_cast<const uint4*>(s);
        for (size_t v = tid; v < elems / V; v += nth) gv[v] = sv[v];
    } else {
        for (size_t e = tid; e < elems; e += nth) g[e] = s[e];
    }
}

// stage Q (2^Q twiddles, distance 2^(R-1-Q)) of a radix-2^R tuple held in registers
template <class A, int R, int Q, bool INV>
NTT_DEVINL void global_stage(typename A::T (&x)[1 << R], const typename A::TW* __restrict__ tw, unsigned w0,
                             const typename A::Ctx& c) {
    constexpr int d = 1 << (R - 1 - Q);
#pragma unroll
    for (int h = 0; h < (1 << Q); ++h) {
        const typename A::TW w = tw[(w0 << Q) + h];
#pragma unroll
        for (int k = 0; k < d; ++k) {
            constexpr int dd = d;
            const int ja = h * 2 * dd + k, jb = ja + dd;
            if (!INV)
                A::fwd_bf(c, x[ja], x[jb], w);
            else  // products of the previous stage sit where bit d/2 of the position is set
                A::inv_bf(c, x[ja], x[jb], w, dd > 1 && (jb & (dd >> 1)) != 0);
        }
    }
}

// Strided pass in global memory over whole polynomials of length 2^logn: stages
// [stage, stage+R).  Used for the top `depth` stages when a polynomial does not fit one CTA.
// grid = (tuples per polynomial / 256, polys [folded into y and z]); a CTA covers 256 consecutive tuples of
// one polynomial, which share their group index i (and so their twiddles) whenever the distance
// 2^log_t2 of the last fused stage is at least 256 -- always the case for the passes the plans
// launch (log_t2 >= 12).  The inverse knows statically which tuple positions hold products of the
// previous stage (bit d/2 of the position), so only sums are re-canonicalised.
template <class A, int R, bool INV>
__global__ void __launch_bounds__(256, 3) ntt_global_pass_kernel(typename A::T* __restrict__ data, size_t num_polys,
                                       int logn, int stage,
                                       const typename A::TW* __restrict__ tw, typename A::Ctx c,
                                       int finalize) {
    using T = typename A::T;
    const int log_t2 = logn - stage - R;
    const unsigned ot = blockIdx.x * 256u + threadIdx.x;  // tuple index inside the polynomial
    const unsigned i = log_t2 >= 8 ? (blockIdx.x * 256u) >> log_t2 : ot >> log_t2;
    const unsigned j = ot & ((1u << log_t2) - 1u);
    const size_t off = ((size_t)i << (log_t2 + R)) + j;
    const unsigned w0 = (1u << stage) + i;
    const size_t poly = (size_t)blockIdx.z * gridDim.y + blockIdx.y;
    if (poly >= num_polys) return;
    {
        T* base = data + (poly << logn) + off;
        T x[1 << R];
#pragma unroll
        for (int k = 0; k < (1 << R); ++k) x[k] = base[(size_t)k << log_t2];
        if (!INV) {
            global_stage<A, R, 0, false>(x, tw, w0, c);
            if constexpr (R >= 2) global_stage<A, R, 1, false>(x, tw, w0, c);
            if constexpr (R >= 3) global_stage<A, R, 2, false>(x, tw, w0, c);
            if constexpr (R >= 4) global_stage<A, R, 3, false>(x, tw, w0, c);
        } else {
            if constexpr (R >= 4) global_stage<A, R, 3, true>(x, tw, w0, c);
            if constexpr (R >= 3) global_stage<A, R, 2, true>(x, tw, w0, c);
            if constexpr (R >= 2) global_stage<A, R, 1, true>(x, tw, w0, c);
            global_stage<A, R, 0, true>(x, tw, w0, c);
        }
        if (finalize) {
            if (!INV) {
#pragma unroll
                for (int k = 0; k < (1 << R); ++k) x[k] = A::fwd_fin(c, x[k]);
            } else {  // the last stage leaves sums in the lower half, canonical products in the upper
#pragma unroll
                for (int k = 0; k < (1 << (R - 1)); ++k) x[k] = A::inv_fin(c, x[k]);
#pragma unroll
                for (int k = (1 << (R - 1)); k < (1 << R); ++k) x[k] = A::inv_fin_prod(c, x[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < (1 << R); ++k) base[(size_t)k << log_t2] = x[k];
    }
}

// ---- TMA-staged strided pass -----------------------------------------------------------------
// The same butterflies as ntt_global_pass_kernel, with the strided rows moved by the bulk-copy
// engine (cp.async.bulk, 1-D TMA) instead of by the threads: a CTA walks over tiles of 256 tuples
// (2^R rows of 256 contiguous coefficients), double-buffered in shared memory.  One thread arms an
// mbarrier with the tile's byte count and issues the 2^R row loads of the NEXT tile, everybody
// waits for the current tile, transforms it in place in shared memory, and the same thread sends
// the rows back with bulk stores -- so the loads of tile t+1 and the stores of tile t-1 overlap
// the arithmetic of tile t, which the register-staged kernel (two or three resident CTAs of
// 80-100 registers) cannot do.
namespace tma {
NTT_DEVINL unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
NTT_DEVINL void mbar_init(void* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
NTT_DEVINL void mbar_expect_tx(void* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
NTT_DEVINL void mbar_wait(void* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
NTT_DEVINL void bulk_load(void* dst_smem, const void* src_gmem, unsigned bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
NTT_DEVINL void bulk_store(void* dst_gmem, const void* src_smem, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
}
NTT_DEVINL void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
NTT_DEVINL void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
NTT_DEVINL void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
NTT_DEVINL void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
}  // namespace tma

template <class A, int R>
constexpr size_t global_pass_tma_smem() {
    return 2 * (size_t)(1 << R) * 256 * sizeof(typename A::T) + 16;
}

template <class A, int R, bool INV>
__global__ void __launch_bounds__(256) ntt_global_pass_tma_kernel(typename A::T* __restrict__ data, size_t total_tiles,
                                                                  unsigned log_tiles_per_poly, int logn, int stage,
                                                                  const typename A::TW* __restrict__ tw, typename A::Ctx c,
                                                                  int finalize) {
    using T = typename A::T;
    constexpr unsigned ROWS = 1u << R, ROW_BYTES = 256u * sizeof(T);
    extern __shared__ __align__(128) unsigned char tma_smem_raw[];
    T* buf = reinterpret_cast<T*>(tma_smem_raw);                                                // [2][ROWS][256]
    unsigned long long* full = reinterpret_cast<unsigned long long*>(buf + 2 * ROWS * 256);  // [2]
    const unsigned t = threadIdx.x;
    const int log_t2 = logn - stage - R;
    // global address of row 0 of tile g (row k is k << log_t2 elements further)
    auto tile_base = [&](size_t g) {
        const size_t poly = g >> log_tiles_per_poly;
        const unsigned ot0 = ((unsigned)g & ((1u << log_tiles_per_poly) - 1u)) * 256u;
        const unsigned i = ot0 >> log_t2, j0 = ot0 & ((1u << log_t2) - 1u);
        return data + (poly << logn) + ((size_t)i << (log_t2 + R)) + j0;
    };
    if (t == 0) {
        tma::mbar_init(&full[0], 1);
        tma::mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    size_t g = blockIdx.x;
    if (g >= total_tiles) return;
    if (t == 0) {
        tma::mbar_expect_tx(&full[0], ROWS * ROW_BYTES);
        const T* src = tile_base(g);
#pragma unroll
        for (unsigned k = 0; k < ROWS; ++k) tma::bulk_load(buf + k * 256, src + ((size_t)k << log_t2), ROW_BYTES, &full[0]);
    }
    for (unsigned it = 0; g < total_tiles; ++it, g += gridDim.x) {
        const unsigned cur = it & 1u;
        T* tile = buf + cur * ROWS * 256;
        const size_t gn = g + gridDim.x;
        if (t == 0 && gn < total_tiles) {
            tma::bulk_wait_read0();  // the stores of the previous tile have finished reading the other buffer
            tma::mbar_expect_tx(&full[cur ^ 1u], ROWS * ROW_BYTES);
            const T* src = tile_base(gn);
            T* dst = buf + (cur ^ 1u) * ROWS * 256;
#pragma unroll
            for (unsigned k = 0; k < ROWS; ++k)
                tma::bulk_load(dst + k * 256, src + ((size_t)k << log_t2), ROW_BYTES, &full[cur ^ 1u]);
        }
        tma::mbar_wait(&full[cur], (it >> 1) & 1u);
        const unsigned ot0 = ((unsigned)g & ((1u << log_tiles_per_poly) - 1u)) * 256u;
        const unsigned w0 = (1u << stage) + (ot0 >> log_t2);  // log_t2 >= 8: one group per tile
        T x[ROWS];
#pragma unroll
        for (unsigned k = 0; k < ROWS; ++k) x[k] = tile[k * 256 + t];
        if (!INV) {
            global_stage<A, R, 0, false>(x, tw, w0, c);
            if constexpr (R >= 2) global_stage<A, R, 1, false>(x, tw, w0, c);
            if constexpr (R >= 3) global_stage<A, R, 2, false>(x, tw, w0, c);
            if constexpr (R >= 4) global_stage<A, R, 3, false>(x, tw, w0, c);
        } else {
            if constexpr (R >= 4) global_stage<A, R, 3, true>(x, tw, w0, c);
            if constexpr (R >= 3) global_stage<A, R, 2, true>(x, tw, w0, c);
            if constexpr (R >= 2) global_stage<A, R, 1, true>(x, tw, w0, c);
            global_stage<A, R, 0, true>(x, tw, w0, c);
        }
        if (finalize) {
            if (!INV) {
#pragma unroll
                for (unsigned k = 0; k < ROWS; ++k) x[k] = A::fwd_fin(c, x[k]);
            } else {
#pragma unroll
                for (unsigned k = 0; k < ROWS / 2; ++k) x[k] = A::inv_fin(c, x[k]);
#pragma unroll
                for (unsigned k = ROWS / 2; k < ROWS; ++k) x[k] = A::inv_fin_prod(c, x[k]);
            }
        }
#pragma unroll
        for (unsigned k = 0; k < ROWS; ++k) tile[k * 256 + t] = x[k];
        tma::fence_async_smem();  // generic-proxy writes -> visible to the bulk-copy engine
        __syncthreads();
        if (t == 0) {
            T* dst = tile_base(g);
#pragma unroll
            for (unsigned k = 0; k < ROWS; ++k) tma::bulk_store(dst + ((size_t)k << log_t2), tile + k * 256, ROW_BYTES);
            tma::bulk_commit();
        }
    }
    if (t == 0) tma::bulk_wait0();
}

// ---- pointwise kernels (reference: prime64.rs:1050-1222, prime32.rs:900-1015) ----
// rhs may be shared by the whole batch: rhs index = i % rhs_period (rhs_period == total: none).
// The *_vec kernels handle 16 bytes (V = 2 u64 / 4 u32 coefficients) per thread and iteration with
// 128-bit accesses; the launcher uses them when every pointer is 16-byte aligned and every length
// and period is a multiple of V (measured on u32: normalize 58 % -> of the HBM copy rate with scalars).
template <class T>
struct Vec16 {
    static constexpr int V = 16 / sizeof(T);
    T v[V];
    NTT_DEVINL static Vec16 load(const T* p) {
        Vec16 r;
        *reinterpret_cast<uint4*>(r.v) = *reinterpret_cast<const uint4*>(p);
        return r;
    }
    NTT_DEVINL void store(T* p) const { *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(v); }
};
template <class A>
__global__ void mul_accumulate_vec_kernel(typename A::T* __restrict__ acc, const typename A::T* __restrict__ lhs,
                                          const typename A::T* __restrict__ rhs, size_t total, size_t lhs_period,
                                          size_t rhs_period, typename A::Ctx c) {
    using VT = Vec16<typename A::T>;
    constexpr int V = VT::V;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * V; i < total;
         i += (size_t)gridDim.x * blockDim.x * V) {
        VT a = VT::load(acc + i), l = VT::load(lhs + (lhs_period == total ? i : i % lhs_period)),
           r = VT::load(rhs + (rhs_period == total ? i : i % rhs_period));
#pragma unroll
        for (int k = 0; k < V; ++k) a.v[k] = A::add_full(c, a.v[k], A::mul_full(c, l.v[k], r.v[k]));
        a.store(acc + i);
    }
}
template <class A>
__global__ void mul_assign_normalize_vec_kernel(typename A::T* __restrict__ lhs, const typename A::T* __restrict__ rhs,
                                                size_t total, size_t rhs_period, typename A::Ctx c,
                                                typename A::TW n_inv) {
    using VT = Vec16<typename A::T>;
    constexpr int V = VT::V;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * V; i < total;
         i += (size_t)gridDim.x * blockDim.x * V) {
        VT l = VT::load(lhs + i), r = VT::load(rhs + (rhs_period == total ? i : i % rhs_period));
#pragma unroll
        for (int k = 0; k < V; ++k) l.v[k] = A::mul_const(c, A::mul_full(c, l.v[k], r.v[k]), n_inv);
        l.store(lhs + i);
    }
}
template <class A>
__global__ void normalize_vec_kernel(typename A::T* __restrict__ v, size_t total, typename A::Ctx c,
                                     typename A::TW n_inv) {
    using VT = Vec16<typename A::T>;
    constexpr int V = VT::V;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * V; i < total;
         i += (size_t)gridDim.x * blockDim.x * V) {
        VT x = VT::load(v + i);
#pragma unroll
        for (int k = 0; k < V; ++k) x.v[k] = A::mul_const(c, x.v[k], n_inv);
        x.store(v + i);
    }
}
template <class A>
__global__ void mul_accumulate_kernel(typename A::T* __restrict__ acc,
                                      const typename A::T* __restrict__ lhs,
                                      const typename A::T* __restrict__ rhs, size_t total,
                                      size_t lhs_period, size_t rhs_period, typename A::Ctx c) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        typename A::T l = lhs[lhs_period == total ? i : i % lhs_period];
        typename A::T r = rhs[rhs_period == total ? i : i % rhs_period];
        acc[i] = A::add_full(c, acc[i], A::mul_full(c, l, r));
    }
}
template <class A>
__global__ void mul_assign_normalize_kernel(typename A::T* __restrict__ lhs,
                                            const typename A::T* __restrict__ rhs, size_t total,
                                            size_t rhs_period, typename A::Ctx c,
                                            typename A::TW n_inv) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        typename A::T r = rhs[rhs_period == total ? i : i % rhs_period];
        lhs[i] = A::mul_const(c, A::mul_full(c, lhs[i], r), n_inv);
    }
}
template <class A>
__global__ void normalize_kernel(typename A::T* __restrict__ v, size_t total, typename A::Ctx c,
                                 typename A::TW n_inv) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x)
        v[i] = A::mul_const(c, v[i], n_inv);
}

}  // namespace nttb200
