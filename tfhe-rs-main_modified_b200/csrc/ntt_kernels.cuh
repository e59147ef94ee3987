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

// Strided pass in global memory over whole polynomials of length 2^logn: stages
// [stage, stage+R).  Used for the top `depth` stages when a polynomial does not fit one CTA.
template <class A, int R, bool INV>
__global__ void __launch_bounds__(256) ntt_global_pass_kernel(typename A::T* __restrict__ data, size_t num_polys,
                                       int logn, int stage,
                                       const typename A::TW* __restrict__ tw, typename A::Ctx c,
                                       int finalize) {
    using T = typename A::T;
    const int log_t2 = logn - stage - R;
    const size_t total = num_polys << (logn - R);
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < total;
         o += (size_t)gridDim.x * blockDim.x) {
        size_t poly = o >> (logn - R);
        size_t ot = o & (((size_t)1 << (logn - R)) - 1);
        size_t i = ot >> log_t2, j = ot & (((size_t)1 << log_t2) - 1);
        T* base = data + (poly << logn) + (i << (log_t2 + R)) + j;
        size_t w0 = ((size_t)1 << stage) + i;
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

// ---- pointwise kernels (reference: prime64.rs:1050-1222, prime32.rs:900-1015) ----
// rhs may be shared by the whole batch: rhs index = i % rhs_period (rhs_period == total: none).
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
