// custum_radix -- the fork's cyclic u32 transforms (reference: tfhe-ntt/src/custum_radix/{fwd.rs,inv.rs,fwd_1.rs}).
//
// The reference has three recursive formulations (radix-2, radix-4, split-radix) of ONE function: for a
// table tw[k] = root^k (root of order n modulo p) each forward routine returns X[k] = sum_j a[j] root^(jk),
// natural order in and out, canonical in [0, p); the inverse routines return the same sum over the inverse
// table times a constant that depends on the routine (which of its bases scale).  Because every value is
// canonical the formulation cannot be observed in the result, so the GPU runs one schedule for all of
// them: the iterative form of the reference's radix-2 recursion (identical to it for ANY table of values
// below p, including the multiplication-free size-2 base), a batch of vectors per launch.  Vectors of 8 to
// 2^15 elements take `cr_fast_kernel` (radix-8 register passes, the bit reversal folded into the first
// pass, the caller's table turned into stage-compact Shoup pairs once per call); longer ones are bit-reversed
// in place, run the same kernel block by block and finish with one pass over global memory per further
// stage; n = 1, 2, 4 are done by one thread per vector.
//
// What the recursions add on top of the sum (restated in `final_factor`):
//   ifft_radix2 / ifft_split_radix (inv.rs:178-303)  n_inv when `top`, except n <= 2 (the bases return early)
//   ifft_radix4                    (inv.rs:106-176)  the size-2 base halves: 1/2 when log2 n is odd; n_inv
//                                                    when `top` and n > 2
//   ifft_radix4_recursive_mut      (fwd_1.rs:296-379) n_inv when `top`, every n
// The MultStats counters of fwd_1.rs count zero operands inside each particular recursion; `cr_stats_kernel`
// reproduces them from the levels of the radix-2 schedule (see the section before the host code).
#include <algorithm>
#include <cstring>
#include <vector>

#include "capi_common.cuh"

using namespace nttb200;

namespace {

constexpr unsigned kLogBlockMax = 15;  // 2^15 u32 (+ padding) = 146 KiB of shared memory per CTA

struct CrMod {
    uint32_t p;
    uint64_t mu;  // floor(2^64 / p)
};

// fwd.rs:1-19, literally
NTT_DEVINL uint32_t cr_add(uint32_t a, uint32_t b, uint32_t p) {
    uint64_t s = (uint64_t)a + b;
    return s >= p ? (uint32_t)(s - p) : (uint32_t)s;
}
NTT_DEVINL uint32_t cr_sub(uint32_t a, uint32_t b, uint32_t p) {
    return a >= b ? a - b : (uint32_t)((uint64_t)a + p - b);
}
// (a * b) % p without a division: q = floor(x * mu / 2^64) is floor(x / p) or one less
NTT_DEVINL uint32_t cr_mul(uint32_t a, uint32_t b, const CrMod& m) {
    uint64_t x = (uint64_t)a * b;
    uint64_t r = x - __umul64hi(x, m.mu) * m.p;
    if (r >= m.p) r -= m.p;
    if (r >= m.p) r -= m.p;
    return (uint32_t)r;
}

// n = 1, 2, 4: one thread per vector, in registers.  n = 4 is the size-2 bases on (a0, a2) and (a1, a3) followed by
// the combine with tw[0] and tw[1] (fwd.rs:170-205); the size-2 base does not multiply (fwd.rs:173-178).
__global__ void cr_tiny_kernel(uint32_t* __restrict__ data, const uint32_t* __restrict__ tw, unsigned logn, size_t batch,
                               CrMod m, uint32_t factor) {
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < batch; v += (size_t)gridDim.x * blockDim.x) {
        uint32_t* a = data + (v << logn);
        uint32_t x[4] = {a[0], 0u, 0u, 0u};
        for (unsigned i = 1; i < (1u << logn); ++i) x[i] = a[i];
        if (logn == 1) {
            const uint32_t e = x[0], o = x[1];
            x[0] = cr_add(e, o, m.p);
            x[1] = cr_sub(e, o, m.p);
        } else if (logn == 2) {
            const uint32_t e0 = cr_add(x[0], x[2], m.p), e1 = cr_sub(x[0], x[2], m.p);
            const uint32_t o0 = cr_add(x[1], x[3], m.p), o1 = cr_sub(x[1], x[3], m.p);
            const uint32_t t0 = cr_mul(o0, __ldg(tw), m), t1 = cr_mul(o1, __ldg(tw + 1), m);
            x[0] = cr_add(e0, t0, m.p);
            x[1] = cr_add(e1, t1, m.p);
            x[2] = cr_sub(e0, t0, m.p);
            x[3] = cr_sub(e1, t1, m.p);
        }
        for (unsigned i = 0; i < (1u << logn); ++i) a[i] = factor != 1 ? cr_mul(x[i], factor, m) : x[i];
    }
}

// ---- the main path: 8 <= n <= 2^15, one vector per n/8 threads, radix-8 register passes ----------------
//
// MODE 0 (p < 2^31): the caller's table is first turned into Shoup pairs {w, floor(w 2^32 / p)} in a
// stream-ordered scratch array (cr_stage_table_kernel: one division per entry, once per call), after which a
// product is IMAD.HI + two IMAD + one min in one-word arithmetic: r = o w - floor(o w' / 2^32) p lies in
// [0, 2p) and 2p < 2^32.  MODE 2 (any p): plain entries and the 64-bit Barrett of cr_mul.  Every result
// is canonical, so the two agree wherever both apply.
struct CrFast {
    uint32_t p, factor_shoup;
    const uint2* stw;  // stage-compact table (cr_stage_table_kernel)
    CrMod wide;        // MODE 2
};
template <int MODE>
struct CrTw {
    uint32_t w, ws;
};
template <int MODE>
NTT_DEVINL uint32_t crf_add(uint32_t a, uint32_t b, const CrFast& m) {
    if (MODE == 2) return cr_add(a, b, m.p);
    const uint32_t s = a + b;
    return min(s, s - m.p);
}
template <int MODE>
NTT_DEVINL uint32_t crf_sub(uint32_t a, uint32_t b, const CrFast& m) {
    if (MODE == 2) return a >= b ? a - b : a + m.p - b;  // wraps to the exact value also when a + p >= 2^32
    const uint32_t d = a - b;
    return min(d, d + m.p);
}
template <int MODE>
NTT_DEVINL uint32_t crf_mul(uint32_t o, CrTw<MODE> w, const CrFast& m) {
    if (MODE == 2) return cr_mul(o, w.w, m.wide);
    const uint32_t r = o * w.w - __umulhi(o, w.ws) * m.p;
    return min(r, r - m.p);
}
// Stage-compact copy of the caller's table: the stage of length 2h reads tw[k * n / (2h)], k < h (the
// subsampled tables of fwd.rs:188-192), a stride of n / (2h) entries between neighbouring lanes; the copy holds
// that stage's entries side by side at [h, 2h), so a warp's twiddle load is one or two wavefronts instead of
// up to 32.  With `shoup` each entry carries floor(w 2^32 / p).
__global__ void cr_stage_table_kernel(const uint32_t* __restrict__ tw, unsigned logn, uint32_t p, int shoup,
                                      uint2* __restrict__ out) {
    const unsigned n = 1u << logn;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (i == 0) {
            out[0] = make_uint2(0u, 0u);
            continue;
        }
        const unsigned lh = 31 - __clz(i), k = i - (1u << lh);  // h = 2^lh, stage length 2^(lh+1)
        const uint32_t w = tw[(size_t)k << (logn - lh - 1)];
        out[i] = make_uint2(w, shoup ? (uint32_t)(((uint64_t)w << 32) / p) : 0u);
    }
}

// one word of padding after every 8 and one more after every 64: the tuple accesses of stride 1, 8, 64 ... and
// the bit-reversed stores of the first pass spread over the banks
NTT_DEVINL unsigned cr_pad(unsigned a) { return a + (a >> 3) + (a >> 6); }

// Stages ll .. ll+R-1 of the reference's radix-2 recursion on the 2^R elements x[j] at positions
// base + j * 2^(ll-1), base = group * 2^(ll-1+R) + k: in stage ll+q element j (bit q clear) pairs with
// j + 2^q and sits at offset k + (j mod 2^q) * 2^(ll-1) of its group, which indexes the subsampled table
// (fwd.rs:188-201).  The size-2 stage does not multiply (fwd.rs:173-178).
template <int MODE, int R, int LL>
NTT_DEVINL void cr_tuple(uint32_t (&x)[8], unsigned k, const CrFast& m) {
#pragma unroll
    for (int q = 0; q < R; ++q) {
        const uint2* __restrict__ stage = m.stw + (1u << (LL + q - 1)) + k;
#pragma unroll
        for (int j = 0; j < (1 << R); ++j) {
            if (j & (1 << q)) continue;
            const uint32_t e = x[j], o = x[j + (1 << q)];
            uint32_t t = o;
            if (LL + q != 1) {
                const uint2 w = __ldg(stage + ((j & ((1 << q) - 1)) << (LL - 1)));  // immediate offset
                t = crf_mul<MODE>(o, CrTw<MODE>{w.x, w.y}, m);
            }
            x[j] = crf_add<MODE>(e, t, m);
            x[j + (1 << q)] = crf_sub<MODE>(e, t, m);
        }
    }
}

// Every pass works at half = 2^(LL-1) in {8, 64, 512, 4096}: the padding of base + j * half is then the
// padding of base plus j times a constant, so a tuple's accesses use immediate offsets.
template <int LL>
struct CrStride {
    static constexpr unsigned kHalf = 1u << (LL - 1);
    static constexpr unsigned kPadded = kHalf + (kHalf >> 3) + (kHalf >> 6);
};

// stages LL .. LOGN: three per pass through shared memory, the last 1-3 out to global memory (coalesced)
template <int MODE, int LOGN, int LL>
NTT_DEVINL void cr_passes(uint32_t* __restrict__ g, uint32_t* s, unsigned vectors, const CrFast& m, uint32_t factor) {
    constexpr int REM = LOGN - LL + 1;
    constexpr unsigned PN = (1u << LOGN) + (1u << LOGN >> 3) + (1u << LOGN >> 6);
    constexpr unsigned STEP = CrStride<LL>::kPadded, HALF = CrStride<LL>::kHalf;
    if constexpr (REM > 3) {
        constexpr unsigned LT = LOGN - 3;
        for (unsigned t = threadIdx.x; t < (vectors << LT); t += blockDim.x) {
            const unsigned v = t >> LT, tl = t & ((1u << LT) - 1);
            const unsigned k = tl & (HALF - 1), base = ((tl >> (LL - 1)) << (LL + 2)) + k;
            uint32_t* sv = s + v * PN + cr_pad(base);
            uint32_t x[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = sv[j * STEP];
            cr_tuple<MODE, 3, LL>(x, k, m);
#pragma unroll
            for (int j = 0; j < 8; ++j) sv[j * STEP] = x[j];
        }
        __syncthreads();
        cr_passes<MODE, LOGN, LL + 3>(g, s, vectors, m, factor);
    } else {
        constexpr unsigned LT = LOGN - REM;  // the group is the whole vector: k = tuple index, HALF = 2^LT
        for (unsigned t = threadIdx.x; t < (vectors << LT); t += blockDim.x) {
            const unsigned v = t >> LT, k = t & (HALF - 1);
            const uint32_t* sv = s + v * PN + cr_pad(k);
            uint32_t x[8];
#pragma unroll
            for (int j = 0; j < (1 << REM); ++j) x[j] = sv[j * STEP];
            cr_tuple<MODE, REM, LL>(x, k, m);
            uint32_t* out = g + ((size_t)v << LOGN) + k;
#pragma unroll
            for (int j = 0; j < (1 << REM); ++j)
                out[j * HALF] = factor != 1 ? crf_mul<MODE>(x[j], CrTw<MODE>{factor, m.factor_shoup}, m) : x[j];
        }
    }
}

// The first pass reads the vector in place of the bit reversal: thread tl takes a[tl + brev3(j) * n/8], which
// are the elements 8 * brev(tl) + j of the reversed order, runs stages 1-3 and parks them there.
// PRE: the "vector" is a 2^LOGN block of a longer vector that cr_bitrev_kernel already permuted, so the tuple
// at position tl is read where it lies.
template <int MODE, int LOGN, bool PRE>
__global__ void __launch_bounds__(LOGN >= 13 ? 1024 : (LOGN >= 11 ? (1 << (LOGN - 3)) : 256))
cr_fast_kernel(uint32_t* __restrict__ data, size_t total_vectors, unsigned per_cta, CrFast m, uint32_t factor) {
    extern __shared__ uint32_t s[];
    constexpr unsigned LT = LOGN - 3, PN = (1u << LOGN) + (1u << LOGN >> 3) + (1u << LOGN >> 6);
    const size_t first = (size_t)blockIdx.x * per_cta;
    const unsigned vectors = (unsigned)min((size_t)per_cta, total_vectors - first);
    uint32_t* g = data + (first << LOGN);
    for (unsigned t = threadIdx.x; t < (vectors << LT); t += blockDim.x) {
        const unsigned v = t >> LT, tl = t & ((1u << LT) - 1);
        const uint32_t* src = g + ((size_t)v << LOGN) + (PRE ? 8 * tl : tl);
        uint32_t x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = PRE ? src[j] : src[(size_t)(((j & 1) << 2) | (j & 2) | (j >> 2)) << LT];
        cr_tuple<MODE, 3, 1>(x, 0, m);
        if constexpr (LT == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                g[(size_t)v * 8 + j] = factor != 1 ? crf_mul<MODE>(x[j], CrTw<MODE>{factor, m.factor_shoup}, m) : x[j];
        } else {
            uint32_t* sv = s + v * PN + cr_pad(8 * (PRE ? tl : (__brev(tl) >> (32 - LT))));
#pragma unroll
            for (int j = 0; j < 8; ++j) sv[j] = x[j];
        }
    }
    if constexpr (LT > 0) {
        __syncthreads();
        cr_passes<MODE, LOGN, 4>(g, s, vectors, m, factor);
    }
}

template <int MODE, int LOGN, bool PRE = false>
void launch_fast_n(uint32_t* dev, size_t batch, const CrFast& m, uint32_t factor, cudaStream_t st) {
    constexpr size_t n = size_t(1) << LOGN, tuples = n / 8;
    // short vectors share a CTA (256 tuples per CTA); long ones get up to 1024 threads
    const unsigned per_cta = (unsigned)std::max<size_t>(1, std::min<size_t>(256 / tuples, batch));
    const unsigned threads = (unsigned)std::min<size_t>(1024, std::max<size_t>(32, per_cta * tuples));
    const size_t smem = (size_t)per_cta * (n + n / 8 + n / 64) * sizeof(uint32_t);
    // (smem is the same for every call of this instantiation that needs the opt-in: per_cta is 1 there)
    allow_dynamic_smem<cr_fast_kernel<MODE, LOGN, PRE>>(smem);
    const size_t ctas = (batch + per_cta - 1) / per_cta;
    cr_fast_kernel<MODE, LOGN, PRE><<<(unsigned)ctas, threads, smem, st>>>(dev, batch, per_cta, m, factor);
}

template <int MODE>
void launch_fast(uint32_t* dev, unsigned logn, size_t batch, const CrFast& m, uint32_t factor, cudaStream_t st) {
    switch (logn) {
#define CR_CASE(L) case L: launch_fast_n<MODE, L>(dev, batch, m, factor, st); break;
        CR_CASE(3) CR_CASE(4) CR_CASE(5) CR_CASE(6) CR_CASE(7) CR_CASE(8) CR_CASE(9) CR_CASE(10) CR_CASE(11)
        CR_CASE(12) CR_CASE(13) CR_CASE(14) CR_CASE(15)
#undef CR_CASE
        default: throw CudaError("custum_radix: no single-CTA kernel for this length");
    }
}

// in-place bit reversal of each vector (n > 2^kLogBlockMax)
__global__ void cr_bitrev_kernel(uint32_t* __restrict__ data, unsigned logn, size_t total) {
    const size_t n = (size_t)1 << logn;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
        size_t poly = g >> logn;
        unsigned i = (unsigned)(g & (n - 1)), r = __brev(i) >> (32 - logn);
        if (i < r) {
            uint32_t* v = data + (poly << logn);
            uint32_t t = v[i];
            v[i] = v[r];
            v[r] = t;
        }
    }
}

// one stage of length 2^ll in global memory (ll > kLogBlockMax) over the stage-compact table; the last one
// applies `factor`
template <int MODE>
__global__ void cr_global_stage_kernel(uint32_t* __restrict__ data, unsigned ll, size_t total_bf, CrFast m,
                                       uint32_t factor) {
    const unsigned half = 1u << (ll - 1);
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total_bf; t += (size_t)gridDim.x * blockDim.x) {
        const unsigned k = (unsigned)t & (half - 1);
        const size_t i0 = ((t >> (ll - 1)) << ll) + k, i1 = i0 + half;
        const uint2 w = __ldg(m.stw + half + k);
        const uint32_t e = data[i0], x = crf_mul<MODE>(data[i1], CrTw<MODE>{w.x, w.y}, m);
        uint32_t y0 = crf_add<MODE>(e, x, m), y1 = crf_sub<MODE>(e, x, m);
        if (factor != 1) {
            y0 = crf_mul<MODE>(y0, CrTw<MODE>{factor, m.factor_shoup}, m);
            y1 = crf_mul<MODE>(y1, CrTw<MODE>{factor, m.factor_shoup}, m);
        }
        data[i0] = y0;
        data[i1] = y1;
    }
}

// vectors longer than 2^15: bit reversal in place, the stages up to 2^15 block by block in the single-CTA
// kernel, one pass over global memory for every further stage
template <int MODE>
void launch_long(uint32_t* dev, unsigned logn, size_t batch, const CrFast& m, uint32_t factor, cudaStream_t st) {
    const size_t total = batch << logn;
    cr_bitrev_kernel<<<(unsigned)std::min<size_t>((total + 255) / 256, size_t(148) * 32), 256, 0, st>>>(dev, logn, total);
    launch_fast_n<MODE, kLogBlockMax, true>(dev, batch << (logn - kLogBlockMax), m, 1u, st);
    for (unsigned ll = kLogBlockMax + 1; ll <= logn; ++ll)
        cr_global_stage_kernel<MODE><<<(unsigned)std::min<size_t>((total / 2 + 255) / 256, size_t(148) * 32), 256, 0, st>>>(
            dev, ll, total / 2, m, ll == logn ? factor : 1u);
}

// ---- fwd_1.rs: the `_mut` routines with their MultStats counters ------------------------------------------
//
// Every counted product of the three forward recursions (and of ifft_radix4_recursive_mut) has operands that
// are sub-transforms of the input: S_m^(r)[k], the size-m transform of a[r], a[r + n/m], ... -- exactly what the
// radix-2 schedule holds after its stage of length m (block brev(r)).  One CTA per vector keeps ALL levels in
// shared memory (level l at lv + l * n; n <= 4096), then counts per routine:
//   radix-2 (fwd_1.rs:190-230)      every butterfly above the size-2 base: (tw, odd); the base counts a[1] != 0
//   split-radix (:232-294)          nodes reached by tokens {0, 10, 11} of the residue; per k < m/4:
//                                   (w^k, A1), (w^3k, A2), (J, w^k A1 - w^3k A2), A1 / A2 = the 1 / 3 mod 4 children
//   radix-4 (:102-188)              nodes at even depth; size-4 base: (tw[1], x1 - x3); above it per k < m/4:
//                                   (w^k, A1), (w^2k, A2), (w^3k, A3), (w^(m/4), w^k A1 - w^3k A3)
//   inverse radix-4 (:296-379)      nodes at even depth, m >= 4; per i < m/4 and output e = i + s m/4:
//                                   (w^e, A1), (w^2e, A2), (w^3e, A3)
// A product counts as `nonzero` when both operands are non-zero, else as `skipped` (fwd_1.rs:28-37).
constexpr unsigned kStatsLogMax = 12;
#define NTT_B200_CR_STATS_IFFT_RADIX4 3

struct CrCount {
    unsigned long long nz = 0, sk = 0;
    NTT_DEVINL void add(uint32_t x, uint32_t y) {
        if (x != 0 && y != 0)
            ++nz;
        else
            ++sk;
    }
};

__global__ void __launch_bounds__(256)
cr_stats_kernel(uint32_t* __restrict__ data, const uint32_t* __restrict__ tw, unsigned logn, int kind, CrMod m,
                uint32_t factor, unsigned long long* __restrict__ stats) {
    extern __shared__ uint32_t lv[];
    __shared__ unsigned long long tot[2];
    const unsigned n = 1u << logn;
    uint32_t* a = data + ((size_t)blockIdx.x << logn);
    if (threadIdx.x < 2) tot[threadIdx.x] = 0;
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) lv[logn ? (__brev(i) >> (32 - logn)) : 0u] = a[i];
    __syncthreads();
    for (unsigned l = 1; l <= logn; ++l) {  // level l from level l - 1 (fwd.rs:170-205)
        const uint32_t *src = lv + (size_t)(l - 1) * n;
        uint32_t* dst = lv + (size_t)l * n;
        const unsigned half = 1u << (l - 1);
        for (unsigned t = threadIdx.x; t < n / 2; t += blockDim.x) {
            const unsigned k = t & (half - 1), i0 = ((t >> (l - 1)) << l) + k, i1 = i0 + half;
            const uint32_t e = src[i0], o = src[i1];
            const uint32_t x = l == 1 ? o : cr_mul(o, __ldg(tw + ((size_t)k << (logn - l))), m);
            dst[i0] = cr_add(e, x, m.p);
            dst[i1] = cr_sub(e, x, m.p);
        }
        __syncthreads();
    }
    CrCount c;
    auto T = [&](unsigned l, unsigned i) { return __ldg(tw + ((size_t)(i & ((1u << l) - 1)) << (logn - l))); };  // T_m[i % m]
    if (kind == NTT_B200_CR_RADIX2) {
        for (unsigned t = threadIdx.x; t < n / 2; t += blockDim.x) c.nz += lv[2 * t + 1] != 0;
        for (unsigned l = 2; l <= logn; ++l) {
            const uint32_t* src = lv + (size_t)(l - 1) * n;
            const unsigned half = 1u << (l - 1);
            for (unsigned t = threadIdx.x; t < n / 2; t += blockDim.x) {
                const unsigned k = t & (half - 1);
                c.add(T(l, k), src[((t >> (l - 1)) << l) + k + half]);
            }
        }
    } else {
        for (unsigned l = 2; l <= logn; ++l) {
            const unsigned L = logn - l, q = 1u << (l - 2);
            if (kind != NTT_B200_CR_SPLIT_RADIX && (L & 1)) continue;  // radix-4 nodes sit at even depth
            const uint32_t* sub = lv + (size_t)(l - 2) * n;
            for (unsigned t = threadIdx.x; t < n / 4; t += blockDim.x) {  // (node b, k)
                const unsigned b = t >> (l - 2), k = t & (q - 1);
                if (kind == NTT_B200_CR_SPLIT_RADIX) {
                    const unsigned r = L ? (__brev(b) >> (32 - L)) : 0u;
                    unsigned pos = 0;
                    while (pos < L) pos += ((r >> pos) & 1u) ? 2 : 1;
                    if (pos != L) continue;  // not a node of the split-radix recursion
                }
                const uint32_t* node = sub + ((size_t)b << l);
                const uint32_t a1 = node[2 * q + k], a2 = node[q + k], a3 = node[3 * q + k];  // classes 1, 2, 3 mod 4
                if (kind == NTT_B200_CR_SPLIT_RADIX) {
                    const uint32_t wk = T(l, k), w3 = T(l, 3 * k);
                    c.add(wk, a1);
                    c.add(w3, a3);
                    c.add(T(l, q), cr_sub(cr_mul(a1, wk, m), cr_mul(a3, w3, m), m.p));
                } else if (kind == NTT_B200_CR_RADIX4) {
                    if (l == 2) {  // the size-4 base works on the inputs: x1 - x3
                        c.add(T(2, 1), cr_sub(lv[4 * b + 2], lv[4 * b + 3], m.p));
                    } else {
                        const uint32_t w1 = T(l, k), w3 = T(l, 3 * k);
                        c.add(w1, a1);
                        c.add(T(l, 2 * k), a2);
                        c.add(w3, a3);
                        c.add(T(l, q), cr_sub(cr_mul(w1, a1, m), cr_mul(w3, a3, m), m.p));
                    }
                } else {
                    for (unsigned s4 = 0; s4 < 4; ++s4) {
                        const unsigned e = k + s4 * q;
                        c.add(T(l, e), a1);
                        c.add(T(l, 2 * e), a2);
                        c.add(T(l, 3 * e), a3);
                    }
                }
            }
        }
    }
    atomicAdd(&tot[0], c.nz);
    atomicAdd(&tot[1], c.sk);
    const uint32_t* res = lv + (size_t)logn * n;
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) a[i] = factor != 1 ? cr_mul(res[i], factor, m) : res[i];
    __syncthreads();
    if (threadIdx.x < 2) stats[2 * blockIdx.x + threadIdx.x] = tot[threadIdx.x];
}

// host restatement of fwd.rs:22-39
uint32_t mulmod_h(uint32_t a, uint32_t b, uint32_t p) { return (uint32_t)(((uint64_t)a * b) % p); }
uint32_t powmod_h(uint32_t base, uint32_t exp, uint32_t p) {
    uint32_t res = 1;
    base %= p;
    while (exp > 0) {
        if (exp & 1) res = mulmod_h(res, base, p);
        base = mulmod_h(base, base, p);
        exp >>= 1;
    }
    return res;
}

bool shape_ok(size_t n, size_t tw_len, uint32_t p, int kind, bool inverse, int* status) {
    const int max_kind = inverse ? NTT_B200_CR_RADIX4_MUT : NTT_B200_CR_SPLIT_RADIX;
    if (kind < 0 || kind > max_kind || p < 2) {
        *status = NTT_B200_ERR_ARG;
        return false;
    }
    // n = 0 recurses forever in the reference, other non-powers of two index out of bounds, and so does a
    // table shorter than the vector
    if (n == 0 || (n & (n - 1)) || n > (size_t(1) << 30) || (n > 2 && tw_len < n)) {
        *status = NTT_B200_ERR_LEN;
        return false;
    }
    return true;
}

uint32_t final_factor(int kind, size_t n, unsigned logn, uint32_t p, uint32_t n_inv, bool inverse, bool top) {
    if (!inverse) return 1;
    uint32_t f = 1;
    if (kind == NTT_B200_CR_RADIX4_MUT) return top ? n_inv % p : 1;
    if (kind == NTT_B200_CR_RADIX4 && (logn & 1)) f = powmod_h(2, p - 2, p);  // inv.rs:112-114
    if (top && n > 2) f = mulmod_h(f, n_inv, p);
    return f;
}

void enqueue(uint32_t* dev, size_t n, size_t batch, const uint32_t* tw_dev, uint32_t p, uint32_t factor,
             cudaStream_t st) {
    if (!batch || (n == 1 && factor == 1)) return;
    unsigned logn = 0;
    while ((size_t(1) << logn) < n) ++logn;
    const CrMod m{p, ~0ull / p + ((~0ull % p) + 1 == p ? 1 : 0)};  // floor(2^64 / p), p >= 2
    if (logn >= 3) {
        struct Scratch {  // stream-ordered, released on every exit path
            cudaStream_t st;
            uint2* ptr = nullptr;
            ~Scratch() {
                if (ptr) cudaFreeAsync(ptr, st);
            }
        } table{st};
        int device = 0;
        NTT_CUDA_CHECK(cudaGetDevice(&device));
        keep_pool_cached(device);  // the scratch table comes from and returns to the cached pool
        NTT_CUDA_CHECK(cudaMallocAsync(&table.ptr, n * sizeof(uint2), st));
        const bool shoup = p < (1u << 31);
        cr_stage_table_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, 148), 256, 0, st>>>(tw_dev, logn, p, shoup ? 1 : 0,
                                                                                             table.ptr);
        const CrFast f{p, shoup ? (uint32_t)(((uint64_t)factor << 32) / p) : 0u, table.ptr, m};
        if (logn <= kLogBlockMax) {
            if (shoup)
                launch_fast<0>(dev, logn, batch, f, factor, st);
            else
                launch_fast<2>(dev, logn, batch, f, factor, st);
        } else {
            if (shoup)
                launch_long<0>(dev, logn, batch, f, factor, st);
            else
                launch_long<2>(dev, logn, batch, f, factor, st);
        }
        NTT_CUDA_CHECK(cudaGetLastError());
        return;
    }
    // n = 1, 2, 4
    cr_tiny_kernel<<<(unsigned)std::min<size_t>((batch + 255) / 256, size_t(148) * 8), 256, 0, st>>>(dev, tw_dev, logn, batch,
                                                                                                 m, factor);
    NTT_CUDA_CHECK(cudaGetLastError());
}

int run_device(int kind, uint32_t* dev, size_t n, size_t batch, const uint32_t* tw_dev, size_t tw_len, uint32_t p,
               uint32_t n_inv, bool inverse, bool top, void* stream) {
    int status = NTT_B200_OK;
    if (!shape_ok(n, tw_len, p, kind, inverse, &status)) return status;
    if (batch && (!dev || (n > 2 && !tw_dev))) return NTT_B200_ERR_ARG;
    return guarded([&] {
        unsigned logn = 0;
        while ((size_t(1) << logn) < n) ++logn;
        enqueue(dev, n, batch, tw_dev, p, final_factor(kind, n, logn, p, n_inv, inverse, top),
                static_cast<cudaStream_t>(stream));
        return NTT_B200_OK;
    });
}

// host vectors: table and data go up, the kernels run, data comes back (32 MiB chunks on one stream)
int run_host(int kind, uint32_t* host, size_t n, size_t batch, const uint32_t* tw, size_t tw_len, uint32_t p,
             uint32_t n_inv, bool inverse, bool top) {
    int status = NTT_B200_OK;
    if (!shape_ok(n, tw_len, p, kind, inverse, &status)) return status;
    if (batch && (!host || (n > 2 && !tw))) return NTT_B200_ERR_ARG;
    if (!batch) return NTT_B200_OK;
    return guarded([&] {
        int device = 0;
        NTT_CUDA_CHECK(cudaGetDevice(&device));
        keep_pool_cached(device);
        cudaStream_t st = cached_stream(device);
        unsigned logn = 0;
        while ((size_t(1) << logn) < n) ++logn;
        const uint32_t factor = final_factor(kind, n, logn, p, n_inv, inverse, top);
        const size_t poly_bytes = n * sizeof(uint32_t);
        const size_t chunk = std::min(batch, std::max<size_t>(1, (size_t(32) << 20) / poly_bytes));
        struct Scratch {  // stream-ordered, released on every exit path
            cudaStream_t st;
            uint32_t* ptr = nullptr;
            ~Scratch() {
                if (ptr) cudaFreeAsync(ptr, st);
            }
        } s_tw{st}, s_d{st};
        const size_t tw_bytes = (n > 2 ? n : 1) * sizeof(uint32_t);
        NTT_CUDA_CHECK(cudaMallocAsync(&s_tw.ptr, tw_bytes, st));
        NTT_CUDA_CHECK(cudaMallocAsync(&s_d.ptr, chunk * poly_bytes, st));
        uint32_t *d_tw = s_tw.ptr, *d = s_d.ptr;
        if (n > 2) NTT_CUDA_CHECK(cudaMemcpyAsync(d_tw, tw, tw_bytes, cudaMemcpyHostToDevice, st));
        for (size_t b0 = 0; b0 < batch; b0 += chunk) {
            const size_t nb = std::min(chunk, batch - b0);
            uint32_t* h = host + b0 * n;
            NTT_CUDA_CHECK(cudaMemcpyAsync(d, h, nb * poly_bytes, cudaMemcpyHostToDevice, st));
            enqueue(d, n, nb, d_tw, p, factor, st);
            NTT_CUDA_CHECK(cudaMemcpyAsync(h, d, nb * poly_bytes, cudaMemcpyDeviceToHost, st));
        }
        NTT_CUDA_CHECK(cudaStreamSynchronize(st));
        return NTT_B200_OK;
    });
}

// the `_mut` routines: values and counters of `batch` vectors, one CTA each; the counters of vector v are ADDED onto
// stats[2v] (nonzero) and stats[2v + 1] (skipped)
int run_stats(int kind, uint32_t* host, size_t n, size_t batch, const uint32_t* tw, size_t tw_len, uint32_t p,
              uint32_t factor, uint64_t* stats) {
    if (kind < 0 || kind > NTT_B200_CR_STATS_IFFT_RADIX4 || p < 2 || (batch && !stats)) return NTT_B200_ERR_ARG;
    if (n == 0 || (n & (n - 1)) || (n > 2 && tw_len < n)) return NTT_B200_ERR_LEN;
    if (n > (size_t(1) << kStatsLogMax)) return NTT_B200_ERR_UNSUPPORTED;  // capacity of the counting kernel, not a caller error
    if (batch && (!host || (n > 2 && !tw))) return NTT_B200_ERR_ARG;
    if (!batch) return NTT_B200_OK;
    return guarded([&] {
        int device = 0;
        NTT_CUDA_CHECK(cudaGetDevice(&device));
        keep_pool_cached(device);
        cudaStream_t st = cached_stream(device);
        unsigned logn = 0;
        while ((size_t(1) << logn) < n) ++logn;
        struct Scratch {
            cudaStream_t st;
            void* ptr = nullptr;
            ~Scratch() {
                if (ptr) cudaFreeAsync(ptr, st);
            }
        } buf{st};
        const size_t smem = (size_t)(logn + 1) * n * sizeof(uint32_t);
        if (smem > 48 * 1024)
            allow_dynamic_smem<cr_stats_kernel>((kStatsLogMax + 1) * (sizeof(uint32_t) << kStatsLogMax));
        const CrMod m{p, ~0ull / p + ((~0ull % p) + 1 == p ? 1 : 0)};
        const size_t chunk = std::min(batch, std::max<size_t>(1, (size_t(32) << 20) / (n * sizeof(uint32_t))));
        // counters (64-bit, first), vectors, table
        NTT_CUDA_CHECK(cudaMallocAsync(&buf.ptr, chunk * 2 * sizeof(uint64_t) + (chunk + 1) * n * sizeof(uint32_t), st));
        unsigned long long* d_stats = static_cast<unsigned long long*>(buf.ptr);
        uint32_t* d = reinterpret_cast<uint32_t*>(d_stats + 2 * chunk);
        uint32_t* d_tw = d + chunk * n;
        if (n > 2) NTT_CUDA_CHECK(cudaMemcpyAsync(d_tw, tw, n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        std::vector<unsigned long long> got(2 * chunk);
        for (size_t b0 = 0; b0 < batch; b0 += chunk) {
            const size_t nb = std::min(chunk, batch - b0);
            uint32_t* h = host + b0 * n;
            NTT_CUDA_CHECK(cudaMemcpyAsync(d, h, nb * n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
            cr_stats_kernel<<<(unsigned)nb, 256, smem, st>>>(d, d_tw, logn, kind, m, factor, d_stats);
            NTT_CUDA_CHECK(cudaGetLastError());
            NTT_CUDA_CHECK(cudaMemcpyAsync(h, d, nb * n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            NTT_CUDA_CHECK(cudaMemcpyAsync(got.data(), d_stats, nb * 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
            NTT_CUDA_CHECK(cudaStreamSynchronize(st));
            for (size_t i = 0; i < 2 * nb; ++i) stats[2 * b0 + i] += got[i];
        }
        return NTT_B200_OK;
    });
}

}  // namespace

extern "C" {

int ntt_b200_custum_radix_fft(int kind, uint32_t* a, size_t n, const uint32_t* twiddles, size_t tw_len, uint32_t p) {
    return run_host(kind, a, n, 1, twiddles, tw_len, p, 1, false, false);
}
int ntt_b200_custum_radix_ifft(int kind, uint32_t* a, size_t n, const uint32_t* inv_twiddles, size_t tw_len,
                               uint32_t p, uint32_t n_inv, int top) {
    return run_host(kind, a, n, 1, inv_twiddles, tw_len, p, n_inv, true, top != 0);
}
int ntt_b200_custum_radix_fft_batch(int kind, uint32_t* host, size_t n, size_t batch, const uint32_t* twiddles,
                                    size_t tw_len, uint32_t p) {
    return run_host(kind, host, n, batch, twiddles, tw_len, p, 1, false, false);
}
int ntt_b200_custum_radix_ifft_batch(int kind, uint32_t* host, size_t n, size_t batch, const uint32_t* inv_twiddles,
                                     size_t tw_len, uint32_t p, uint32_t n_inv, int top) {
    return run_host(kind, host, n, batch, inv_twiddles, tw_len, p, n_inv, true, top != 0);
}
int ntt_b200_custum_radix_fft_device(int kind, uint32_t* dev, size_t n, size_t batch, const uint32_t* twiddles_dev,
                                     size_t tw_len, uint32_t p, void* stream) {
    return run_device(kind, dev, n, batch, twiddles_dev, tw_len, p, 1, false, false, stream);
}
int ntt_b200_custum_radix_ifft_device(int kind, uint32_t* dev, size_t n, size_t batch,
                                      const uint32_t* inv_twiddles_dev, size_t tw_len, uint32_t p, uint32_t n_inv,
                                      int top, void* stream) {
    return run_device(kind, dev, n, batch, inv_twiddles_dev, tw_len, p, n_inv, true, top != 0, stream);
}

int ntt_b200_custum_radix_fft_mut(int kind, uint32_t* a, size_t n, const uint32_t* twiddles, size_t tw_len, uint32_t p,
                                  uint64_t* stats) {
    if (kind == NTT_B200_CR_RADIX4_MUT) return NTT_B200_ERR_ARG;
    return run_stats(kind, a, n, 1, twiddles, tw_len, p, 1u, stats);
}
int ntt_b200_custum_radix_fft_mut_batch(int kind, uint32_t* host, size_t n, size_t batch, const uint32_t* twiddles,
                                        size_t tw_len, uint32_t p, uint64_t* stats) {
    if (kind == NTT_B200_CR_RADIX4_MUT) return NTT_B200_ERR_ARG;
    return run_stats(kind, host, n, batch, twiddles, tw_len, p, 1u, stats);
}
int ntt_b200_custum_radix_ifft_radix4_mut(uint32_t* a, size_t n, const uint32_t* inv_twiddles, size_t tw_len, uint32_t p,
                                          uint32_t n_inv, int top, uint64_t* stats) {
    if (p < 2) return NTT_B200_ERR_ARG;
    return run_stats(NTT_B200_CR_STATS_IFFT_RADIX4, a, n, 1, inv_twiddles, tw_len, p, top ? n_inv % p : 1u, stats);
}

}  // extern "C"
