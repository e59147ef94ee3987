// Register-resident throughput of Solinas (p = 2^64 - 2^32 + 1) radix-8 tuple variants on B200 (sm_100a).
//
// VERDICT r1 item 1 asks for the butterfly's integer work to be cut by (a) shift twiddles in the
// stages whose twiddles are powers of two, (b) radix-4 tuples with a 2^48 rotation, (c) removing the
// IMAD.MOVs of the 64x64 product, (d) 96-bit lazy accumulators -- or for the measurements that show
// each of them losing.  Every variant below works on the unit the shipped kernels work on: eight
// 64-bit values per polynomial and thread, two polynomials per thread (PPT = 2), twelve butterflies
// per polynomial (three radix-2 stages), twiddles in registers, 256 threads per CTA, four CTAs per SM
// (64 registers), no memory traffic in the timed loop.  The harness checks every variant against a
// host big-integer evaluation of the same linear map before it times it.
//
//   V0  shipped: Solinas64::fwd_bf (Montgomery-form twiddle, mult-free REDC, add_lazy / sub_lazy)
//   V1  V0 with the wrap of the sum folded by one IMAD.WIDE (c * eps + s) instead of three ALU ops
//   V2  (a) stages 0..2 of the reference's sweep (twiddles 2^48 | 2^24 2^72 | 2^12 2^60 2^36 2^84,
//       tfhe-ntt/src/prime64.rs:162-179) as shift + fold: no 64x64 product at all
//   V3  (b) two stages as radix-4 tuples: three Montgomery products + one rotation by 2^48 per four points
//   V4  (c) V0 with the 128-bit product written as four mad.wide.u32 whose partial sums cannot overflow
//       (no carry flag, no IMAD.MOV of a high half into an accumulator pair)
//   V6  (a) in its most favourable case only: twelve butterflies whose twiddle is 2^24 (e < 32: no limb rotation)
//   V5  (d) V0 with 96-bit (three-limb) accumulators inside the tuple, folded where a value is multiplied
//       and at the tuple's exit
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o solinas_bf_variants.bin solinas_bf_variants.cu
// SASS histograms: cuobjdump -sass solinas_bf_variants.bin (tools/sass_hist.py summarises them).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../tfhe-rs-main_modified_b200/csrc/ntt_arith.cuh"

using nttb200::Solinas64;
typedef unsigned __int128 u128;
static constexpr uint64_t P = 0xFFFFFFFF00000001ull;
static constexpr uint64_t EPS = 0xFFFFFFFFull;

#define DEVINL __device__ __forceinline__

// ---------------------------------------------------------------------------------------------
// V1: a + b with the wrap folded by IMAD.WIDE (b <= 2^64 - 2^32, so a single fold suffices)
__constant__ uint32_t c_eps;  // 2^32 - 1 at run time: a literal would be strength-reduced to IMAD.HI + moves
DEVINL uint32_t eps_reg() { return c_eps; }
DEVINL uint64_t add_lazy_w(uint64_t a, uint64_t b) {
    uint64_t s;
    uint32_t c;
    asm("{ .reg .u32 a0,a1,b0,b1;\n\t"
        "mov.b64 {a0,a1}, %2; mov.b64 {b0,b1}, %3;\n\t"
        "add.cc.u32 a0,a0,b0; addc.cc.u32 a1,a1,b1; addc.u32 %1,0,0;\n\t"
        "mov.b64 %0, {a0,a1}; }"
        : "=l"(s), "=r"(c)
        : "l"(a), "l"(b));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(s) : "r"(c), "r"(eps_reg()));
    return s;
}

// V4: the 128-bit product as four mad.wide.u32 without carries: every partial sum stays below 2^64
//   m  = b0*w1 + hi(b0*w0)            <= (2^32-1)^2 + 2^32 - 1
//   m2 = b1*w0 + lo(m)                <= same
//   h  = b1*w1 + hi(m) + hi(m2)       <  2^64
// followed by the shipped eight-instruction REDC.
DEVINL uint64_t mulm_chain(uint64_t b, uint64_t wm) {
    uint64_t r;
    asm("{ .reg .u32 b0,b1,w0,w1,x0,x1,h0,h1,m1,tt,m,z;\n\t"
        ".reg .u64 p00,mm,m2,hh,t0;\n\t"
        "mov.b64 {b0,b1}, %1; mov.b64 {w0,w1}, %2;\n\t"
        "mul.wide.u32 p00,b0,w0; mov.b64 {x0,z}, p00; cvt.u64.u32 t0,z;\n\t"
        "mad.wide.u32 mm,b0,w1,t0; mov.b64 {m1,z}, mm; cvt.u64.u32 t0,m1; cvt.u64.u32 hh,z;\n\t"
        "mad.wide.u32 m2,b1,w0,t0; mov.b64 {x1,z}, m2; cvt.u64.u32 t0,z; add.u64 hh,hh,t0;\n\t"
        "mad.wide.u32 hh,b1,w1,hh; mov.b64 {h0,h1}, hh;\n\t"
        "add.cc.u32 m1,x0,x1; madc.lo.u32 tt,m1,1,0;\n\t"
        "addc.cc.u32 h0,h0,x1; madc.lo.u32 h1,h1,1,0;\n\t"
        "sub.cc.u32 h1,h1,tt; subc.u32 m,0,0;\n\t"
        "sub.cc.u32 h0,h0,m; subc.u32 h1,h1,0;\n\t"
        "mov.b64 %0,{h0,h1}; }"
        : "=l"(r)
        : "l"(b), "l"(wm));
    return r;
}

// ---------------------------------------------------------------------------------------------
// V2: butterfly with the twiddle 2^E, 0 < E < 96, E = 32 q + r.
// y = b << r = (y2 : y1 : y0), y2 < 2^r.  With phi = 2^32: phi^2 = phi - 1, phi^3 = -1, so
//   q = 0: t = (y1 : y0) + y2 * eps
//   q = 1: t = (y0 + y1) * phi - (y1 + y2)
//   q = 2: t = (y0 - y2) * phi - (y0 + y1)
// t is brought to one 64-bit representative <= p (exact, every wrap folded), then the shipped lazy
// add / sub finish the butterfly.  The q = 0 case is also written with the three-limb sums a +- y
// folded once each (shift_bf_q0_direct), its cheapest form.
template <int R>
DEVINL void shl96(uint64_t b, uint32_t& y0, uint32_t& y1, uint32_t& y2) {
    uint32_t b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32);
    if constexpr (R == 0) {
        y0 = b0, y1 = b1, y2 = 0;
    } else {
        y0 = b0 << R;
        y1 = __funnelshift_l(b0, b1, R);
        y2 = b1 >> (32 - R);
    }
}
// value = lo64 + top * eps - sub32, 0 <= top, sub32 < 2^32 -> canonical representative
DEVINL uint64_t fold_pos_neg(uint64_t lo64, uint32_t top, uint32_t sub32) {
    // all arithmetic exact in 128 bits, then the reference's reduction of a 96-bit value
    u128 v = (u128)lo64 + (u128)top * EPS + (u128)P - sub32;  // < 2^64 + 2^64 + 2^64
    uint64_t lo = (uint64_t)v;
    uint32_t hi = (uint32_t)(v >> 64);  // 0..2
    uint64_t m = (uint64_t)hi * EPS, s = lo + m;
    if (s < m) s += EPS;
    return Solinas64::canon(s);
}
template <int E>
DEVINL uint64_t shl_mod(uint64_t b) {
    constexpr int Q = E / 32, R = E % 32;
    uint32_t y0, y1, y2;
    shl96<R>(b, y0, y1, y2);
    if constexpr (Q == 0) {
        return fold_pos_neg(((uint64_t)y1 << 32) | y0, y2, 0);
    } else if constexpr (Q == 1) {
        // (y0 + y1) * phi - (y1 + y2): phi * (y0 + y1) = ((y0 + y1) mod 2^32 : 0) + carry * eps
        uint64_t h = (uint64_t)y0 + y1;
        uint64_t l = (uint64_t)y1 + y2;  // < 2^33
        u128 v = ((u128)(uint32_t)h << 32) + (u128)(uint32_t)(h >> 32) * EPS + 2 * (u128)P - l;
        uint64_t lo = (uint64_t)v;
        uint32_t hi = (uint32_t)(v >> 64);
        uint64_t m = (uint64_t)hi * EPS, s = lo + m;
        if (s < m) s += EPS;
        return Solinas64::canon(s);
    } else {
        // (y0 - y2) * phi - (y0 + y1)
        uint64_t l = (uint64_t)y0 + y1;
        u128 v = ((u128)y0 << 32) + 2 * (u128)P - ((u128)y2 << 32) - l;  // y2 < 2^31: positive
        uint64_t lo = (uint64_t)v;
        uint32_t hi = (uint32_t)(v >> 64);
        uint64_t m = (uint64_t)hi * EPS, s = lo + m;
        if (s < m) s += EPS;
        return Solinas64::canon(s);
    }
}
// q = 0 in its cheapest form: S = a + y and D = a - y as three-limb integers, one fold each.
//   S = (s2 : s1 : s0), s2 <= 2^R:  S == (s1 + s2 : s0) - s2 + (k - j) * eps, k = carry of s1 + s2, j = borrow
//   D = (-(n2) : d1 : d0), n2 = y2 + borrow:  D == (d1 - n2 : d0) + n2 - (j - k) * eps
template <int R>
DEVINL void shift_bf_q0_direct(uint64_t& a, uint64_t& b) {
    uint32_t y0, y1, y2;
    shl96<R>(b, y0, y1, y2);
    uint64_t op, om;
    asm("{ .reg .u32 a0,a1,s0,s1,s2,k,w;\n\t"
        "mov.b64 {a0,a1}, %1;\n\t"
        "add.cc.u32 s0,a0,%2; addc.cc.u32 s1,a1,%3; addc.u32 s2,%4,0;\n\t"
        "add.cc.u32 s1,s1,s2; addc.u32 k,0,0;\n\t"
        "sub.cc.u32 s0,s0,s2; subc.cc.u32 s1,s1,0; subc.u32 w,k,0;\n\t"
        "mov.b64 %0,{s0,s1}; mad.wide.u32 %0,w,%5,%0; }"
        : "=l"(op)
        : "l"(a), "r"(y0), "r"(y1), "r"(y2), "r"(eps_reg()));
    asm("{ .reg .u32 a0,a1,d0,d1,n2,j,w;\n\t"
        "mov.b64 {a0,a1}, %1;\n\t"
        "sub.cc.u32 d0,a0,%2; subc.cc.u32 d1,a1,%3; subc.u32 n2,0,%4;\n\t"  // n2 = -(y2 + borrow)
        "neg.s32 n2,n2;\n\t"
        "sub.cc.u32 d1,d1,n2; subc.u32 j,0,0;\n\t"                           // j = 0 or 0xFFFFFFFF
        "add.cc.u32 d0,d0,n2; addc.cc.u32 d1,d1,0; addc.u32 w,j,0;\n\t"      // w = 0xFFFFFFFF iff net -eps
        "sub.cc.u32 d0,d0,w; subc.u32 d1,d1,0;\n\t"
        "mov.b64 %0,{d0,d1}; }"
        : "=l"(om)
        : "l"(a), "r"(y0), "r"(y1), "r"(y2));
    a = op;
    b = om;
}
template <int E>
DEVINL void shift_bf(uint64_t& a, uint64_t& b) {
    if constexpr (E < 32) {
        shift_bf_q0_direct<E>(a, b);
    } else {
        uint64_t t = shl_mod<E>(b), z0 = a;
        a = Solinas64::add_lazy(z0, t);
        b = Solinas64::sub_lazy(z0, t);
    }
}

// ---------------------------------------------------------------------------------------------
// V5: three-limb values (signed top limb).  a' = a + t, a - t cost three instructions each; a value is
// folded to 64 bits (fold3) before it is multiplied and when it leaves the tuple.
struct L3 {
    uint32_t x0, x1;
    int32_t x2;
};
DEVINL L3 l3_from(uint64_t a) { return L3{(uint32_t)a, (uint32_t)(a >> 32), 0}; }
DEVINL L3 l3_add(L3 a, uint64_t t) {
    L3 r;
    asm("add.cc.u32 %0,%3,%6; addc.cc.u32 %1,%4,%7; addc.u32 %2,%5,0;"
        : "=r"(r.x0), "=r"(r.x1), "=r"(r.x2)
        : "r"(a.x0), "r"(a.x1), "r"(a.x2), "r"((uint32_t)t), "r"((uint32_t)(t >> 32)));
    return r;
}
DEVINL L3 l3_sub(L3 a, uint64_t t) {
    L3 r;
    asm("sub.cc.u32 %0,%3,%6; subc.cc.u32 %1,%4,%7; subc.u32 %2,%5,0;"
        : "=r"(r.x0), "=r"(r.x1), "=r"(r.x2)
        : "r"(a.x0), "r"(a.x1), "r"(a.x2), "r"((uint32_t)t), "r"((uint32_t)(t >> 32)));
    return r;
}
// (x2 : x1 : x0) with -4 <= x2 <= 4  ->  one 64-bit representative: lo64 + x2 * eps, wraps folded
DEVINL uint64_t fold3(L3 v) {
    uint64_t lo = ((uint64_t)v.x1 << 32) | v.x0;
    // lo + (x2 + 4) * eps - 4 * eps, kept non-negative by adding p:  4 * eps < p
    uint64_t up = (uint64_t)(uint32_t)(v.x2 + 4) * EPS;
    uint64_t s = lo + up;
    if (s < up) s += EPS;  // cannot wrap twice: after a wrap s < 8 * eps
    uint64_t d = s - 4 * EPS;
    if (s < 4 * EPS) d -= EPS;  // s - 4 eps + 2^64 - eps >= 0 because the borrow means s < 4 eps only when ... see note
    return d;
}

// ---------------------------------------------------------------------------------------------
// the tuple variants; w[0] = stage-0 twiddle, w[1..2] stage 1, w[3..6] stage 2 (Montgomery form)
template <int V>
DEVINL void bf(uint64_t& a, uint64_t& b, uint64_t w) {
    if (V == 0) {
        Solinas64::fwd_bf(Solinas64::Ctx{P}, a, b, w);
    } else if (V == 1) {
        uint64_t t = Solinas64::mulm(b, w), z0 = a;
        a = add_lazy_w(z0, t);
        b = Solinas64::sub_lazy(z0, t);
    } else if (V == 4) {
        uint64_t t = mulm_chain(b, w), z0 = a;
        a = Solinas64::add_lazy(z0, t);
        b = Solinas64::sub_lazy(z0, t);
    }
}
template <int V>
DEVINL void tuple8(uint64_t (&x)[8], const uint64_t (&w)[7]) {
    if (V == 0 || V == 1 || V == 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) bf<V>(x[k], x[k + 4], w[0]);
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int k = 0; k < 2; ++k) bf<V>(x[4 * h + k], x[4 * h + k + 2], w[1 + h]);
#pragma unroll
        for (int h = 0; h < 4; ++h) bf<V>(x[2 * h], x[2 * h + 1], w[3 + h]);
    } else if (V == 2) {
#pragma unroll
        for (int k = 0; k < 4; ++k) shift_bf<48>(x[k], x[k + 4]);
#pragma unroll
        for (int k = 0; k < 2; ++k) shift_bf<24>(x[k], x[k + 2]);
#pragma unroll
        for (int k = 0; k < 2; ++k) shift_bf<72>(x[4 + k], x[4 + k + 2]);
        shift_bf<12>(x[0], x[1]);
        shift_bf<60>(x[2], x[3]);
        shift_bf<36>(x[4], x[5]);
        shift_bf<84>(x[6], x[7]);
    } else if (V == 6) {  // twelve butterflies with the twiddle 2^24 (q = 0, hand-written direct form)
#pragma unroll
        for (int k = 0; k < 4; ++k) shift_bf<24>(x[k], x[k + 4]);
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int k = 0; k < 2; ++k) shift_bf<24>(x[4 * h + k], x[4 * h + k + 2]);
#pragma unroll
        for (int h = 0; h < 4; ++h) shift_bf<24>(x[2 * h], x[2 * h + 1]);
    } else if (V == 3) {
        // stages 0 and 1 as two radix-4 tuples {k, k+2, k+4, k+6}: y = DFT4(x0, w1 x2', ...) with the
        // twiddles of the reference's stage pair folded: t1 = w[1] x[k+2], t2 = w[0] x[k+4], t3 = w[0] w[1] x[k+6]
        // (w[2] = 2^48 w[1] for a stage pair of the negacyclic sweep), then one rotation by 2^48.
        // w[5] carries w[0]*w[1] in Montgomery form for this variant.
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            uint64_t t1 = Solinas64::mulm(x[k + 2], w[1]);
            uint64_t t2 = Solinas64::mulm(x[k + 4], w[0]);
            uint64_t t3 = Solinas64::mulm(x[k + 6], w[5]);
            uint64_t A = Solinas64::add_lazy(x[k], t2), B = Solinas64::sub_lazy(x[k], t2);
            uint64_t C = Solinas64::canon(Solinas64::add_lazy(t1, t3));
            uint64_t D = shl_mod<48>(Solinas64::sub_lazy(t1, t3));
            x[k] = Solinas64::add_lazy(A, C);
            x[k + 2] = Solinas64::sub_lazy(A, C);
            x[k + 4] = Solinas64::add_lazy(B, D);
            x[k + 6] = Solinas64::sub_lazy(B, D);
        }
        // third stage as in V0 so that all variants do twelve butterflies
#pragma unroll
        for (int h = 0; h < 4; ++h) bf<0>(x[2 * h], x[2 * h + 1], w[3 + (h & 1)]);
    } else if (V == 5) {
        // three-limb accumulators: fold only where a value is multiplied or leaves
        L3 y[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint64_t t = Solinas64::mulm(x[k + 4], w[0]);
            L3 a = l3_from(x[k]);
            y[k] = l3_add(a, t);
            y[k + 4] = l3_sub(a, t);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                uint64_t t = Solinas64::mulm(fold3(y[4 * h + k + 2]), w[1 + h]);
                L3 a = y[4 * h + k];
                y[4 * h + k] = l3_add(a, t);
                y[4 * h + k + 2] = l3_sub(a, t);
            }
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            uint64_t t = Solinas64::mulm(fold3(y[2 * h + 1]), w[3 + h]);
            L3 a = y[2 * h];
            x[2 * h] = fold3(l3_add(a, t));
            x[2 * h + 1] = fold3(l3_sub(a, t));
        }
    }
}

template <int V>
__global__ void __launch_bounds__(256, 4) bench_kernel(uint64_t* out, const uint64_t* in, const uint64_t* tw, int iters) {
    uint64_t x[2][8], w[7];
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int pp = 0; pp < 2; ++pp)
#pragma unroll
        for (int k = 0; k < 8; ++k) x[pp][k] = in[(tid * 2 + pp) * 8 + k];
#pragma unroll
    for (int k = 0; k < 7; ++k) w[k] = tw[k];
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int pp = 0; pp < 2; ++pp) tuple8<V>(x[pp], w);
    }
#pragma unroll
    for (int pp = 0; pp < 2; ++pp)
#pragma unroll
        for (int k = 0; k < 8; ++k) out[(tid * 2 + pp) * 8 + k] = Solinas64::canon(x[pp][k]);
}

// occupancy / ILP probe of the shipped tuple: PPT polynomials per thread, MINB resident CTAs per SM asked for
template <int PPT, int MINB>
__global__ void __launch_bounds__(256, MINB) occ_kernel(uint64_t* out, const uint64_t* in, const uint64_t* tw, int iters) {
    uint64_t x[PPT][8], w[7];
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int pp = 0; pp < PPT; ++pp)
#pragma unroll
        for (int k = 0; k < 8; ++k) x[pp][k] = in[(tid * PPT + pp) * 8 + k];
#pragma unroll
    for (int k = 0; k < 7; ++k) w[k] = tw[k];
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int pp = 0; pp < PPT; ++pp) tuple8<0>(x[pp], w);
    }
#pragma unroll
    for (int pp = 0; pp < PPT; ++pp)
#pragma unroll
        for (int k = 0; k < 8; ++k) out[(tid * PPT + pp) * 8 + k] = Solinas64::canon(x[pp][k]);
}
template <int PPT, int MINB>
void occ_probe(uint64_t* d_out, const uint64_t* d_in, const uint64_t* d_tw, int iters, int sms, int clk_khz) {
    const int blocks = sms * MINB * 4;
    for (int wu = 0; wu < 2; ++wu) occ_kernel<PPT, MINB><<<blocks, 256>>>(d_out, d_in, d_tw, iters);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) occ_kernel<PPT, MINB><<<blocks, 256>>>(d_out, d_in, d_tw, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 5;
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, occ_kernel<PPT, MINB>, 256, 0);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, occ_kernel<PPT, MINB>);
    double per_s = (double)blocks * 256 * PPT * 12 * iters / (ms * 1e-3);
    printf("occupancy probe: PPT %d, asked %d CTAs/SM, got %d (%d regs, %zu B local): %6.2f G butterflies/s, %5.1f clk per warp-butterfly per SMSP\n",
           PPT, MINB, occ, fa.numRegs, (size_t)fa.localSizeBytes, per_s / 1e9, (clk_khz * 1e3 * sms * 4) / (per_s / 32));
}

// ---------------------------------------------------------------------------------------------
// host reference
static uint64_t hmul(uint64_t a, uint64_t b) { return (uint64_t)((u128)a * b % P); }
static uint64_t hadd(uint64_t a, uint64_t b) { return (uint64_t)(((u128)a + b) % P); }
static uint64_t hsub(uint64_t a, uint64_t b) { return (uint64_t)(((u128)a + P - b % P) % P); }
static uint64_t hpow(uint64_t a, uint64_t e) {
    uint64_t r = 1;
    for (; e; e >>= 1, a = hmul(a, a))
        if (e & 1) r = hmul(r, a);
    return r;
}
static void host_tuple(uint64_t (&x)[8], const uint64_t (&w)[7]) {  // plain twiddles
    auto b = [&](uint64_t& a, uint64_t& c, uint64_t t) {
        uint64_t m = hmul(c % P, t), z = a % P;
        a = hadd(z, m);
        c = hsub(z, m);
    };
    for (int k = 0; k < 4; ++k) b(x[k], x[k + 4], w[0]);
    for (int h = 0; h < 2; ++h)
        for (int k = 0; k < 2; ++k) b(x[4 * h + k], x[4 * h + k + 2], w[1 + h]);
    for (int h = 0; h < 4; ++h) b(x[2 * h], x[2 * h + 1], w[3 + h]);
}

struct Variant {
    int id;
    const char* name;
    void (*launch)(uint64_t*, const uint64_t*, const uint64_t*, int, int);
};
template <int V>
void launch(uint64_t* out, const uint64_t* in, const uint64_t* tw, int iters, int blocks) {
    bench_kernel<V><<<blocks, 256>>>(out, in, tw, iters);
}

int main(int argc, char** argv) {
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev);
    const int blocks = sms * 4 * 4, threads = blocks * 256;
    const int iters = argc > 1 ? atoi(argv[1]) : 200;
    std::vector<uint64_t> h_in((size_t)threads * 16), h_out(h_in.size());
    uint64_t s = 88172645463325252ull;
    for (auto& v : h_in) {
        s ^= s << 13, s ^= s >> 7, s ^= s << 17;
        v = s % P;
    }
    // edge values in the first threads
    h_in[0] = P - 1, h_in[1] = 0, h_in[2] = 1, h_in[3] = P - 1, h_in[4] = P - 1, h_in[5] = EPS, h_in[6] = P - EPS, h_in[7] = 1ull << 63;
    const uint64_t R64 = EPS;  // 2^64 mod p
    // twiddles of stages 0..2 of the reference sweep (powers of two) and a random set
    uint64_t tw_pow2[7] = {hpow(2, 48), hpow(2, 24), hpow(2, 72), hpow(2, 12), hpow(2, 60), hpow(2, 36), hpow(2, 84)};
    uint64_t tw_24[7];
    for (auto& v : tw_24) v = hpow(2, 24);
    uint64_t tw_rand[7];
    for (auto& v : tw_rand) {
        s ^= s << 13, s ^= s >> 7, s ^= s << 17;
        v = s % P;
    }
    // V3 computes a stage pair (w0; w1, 2^48 w1) + third stage (w3, w4, w3, w4): make the random set of that shape
    uint64_t tw_r4[7] = {tw_rand[0], tw_rand[1], hmul(tw_rand[1], hpow(2, 48)), tw_rand[3], tw_rand[4], tw_rand[3], tw_rand[4]};

    uint64_t *d_in, *d_out, *d_tw;
    cudaMalloc(&d_in, h_in.size() * 8);
    cudaMalloc(&d_out, h_in.size() * 8);
    cudaMalloc(&d_tw, 7 * 8);
    cudaMemcpy(d_in, h_in.data(), h_in.size() * 8, cudaMemcpyHostToDevice);
    const uint32_t eps32 = 0xFFFFFFFFu;
    cudaMemcpyToSymbol(c_eps, &eps32, 4);

    Variant variants[] = {{0, "V0 shipped (mulm + add_lazy + sub_lazy)", launch<0>},
                          {1, "V1 sum wrap folded by IMAD.WIDE", launch<1>},
                          {2, "V2 (a) stages 0-2 as shift + fold", launch<2>},
                          {3, "V3 (b) radix-4 stage pair + 2^48 rotation", launch<3>},
                          {4, "V4 (c) carry-free mad.wide product", launch<4>},
                          {5, "V5 (d) 96-bit accumulators in the tuple", launch<5>},
                          {6, "V6 (a) best case: 12 x twiddle 2^24 (e < 32)", launch<6>}};
    printf("device %d: %d SMs, nominal SM clock %d MHz; %d CTAs x 256 threads, 2 polynomials x 8 values per thread, %d tuple iterations\n",
           dev, sms, clk_khz / 1000, blocks, iters);
    for (auto& v : variants) {
        const uint64_t* twp = v.id == 2 ? tw_pow2 : (v.id == 3 ? tw_r4 : (v.id == 6 ? tw_24 : tw_rand));
        uint64_t dev_tw[7];
        for (int k = 0; k < 7; ++k) dev_tw[k] = hmul(twp[k], R64);  // Montgomery form
        if (v.id == 3) dev_tw[5] = hmul(hmul(twp[0], twp[1]), R64);
        cudaMemcpy(d_tw, dev_tw, sizeof dev_tw, cudaMemcpyHostToDevice);
        // correctness: one iteration against the host
        v.launch(d_out, d_in, d_tw, 1, blocks);
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) {
            printf("%s: CUDA error %s\n", v.name, cudaGetErrorString(err));
            return 1;
        }
        cudaMemcpy(h_out.data(), d_out, h_out.size() * 8, cudaMemcpyDeviceToHost);
        size_t bad = 0;
        for (size_t t = 0; t < 4096; ++t) {
            uint64_t x[8];
            for (int k = 0; k < 8; ++k) x[k] = h_in[t * 8 + k];
            uint64_t wp[7];
            for (int k = 0; k < 7; ++k) wp[k] = twp[k];
            host_tuple(x, wp);
            for (int k = 0; k < 8; ++k) bad += (x[k] != h_out[t * 8 + k]);
        }
        // timing
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        for (int wu = 0; wu < 3; ++wu) v.launch(d_out, d_in, d_tw, iters, blocks);
        cudaEventRecord(e0);
        const int reps = 5;
        for (int r = 0; r < reps; ++r) v.launch(d_out, d_in, d_tw, iters, blocks);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        ms /= reps;
        double bfs = (double)threads * 2 * 12 * iters;  // butterflies per launch
        double per_s = bfs / (ms * 1e-3);
        double clk = clk_khz * 1e3;
        double warp_bf_clk = (clk * sms * 4) / (per_s / 32);  // SMSP clocks per warp-butterfly
        printf("%-46s %s  %8.3f ms  %7.2f G butterflies/s  %5.1f clk per warp-butterfly per SMSP  -> N=2048 bound %6.1f M NTT/s\n",
               v.name, bad ? "MISMATCH" : "ok      ", ms, per_s / 1e9, warp_bf_clk, per_s / 11264 / 1e6);
    }
    uint64_t dev_tw[7];
    for (int k = 0; k < 7; ++k) dev_tw[k] = hmul(tw_rand[k], R64);
    cudaMemcpy(d_tw, dev_tw, sizeof dev_tw, cudaMemcpyHostToDevice);
    occ_probe<1, 4>(d_out, d_in, d_tw, iters, sms, clk_khz);
    occ_probe<1, 6>(d_out, d_in, d_tw, iters, sms, clk_khz);
    occ_probe<1, 8>(d_out, d_in, d_tw, iters, sms, clk_khz);
    occ_probe<2, 2>(d_out, d_in, d_tw, iters, sms, clk_khz);
    occ_probe<2, 3>(d_out, d_in, d_tw, iters, sms, clk_khz);
    occ_probe<2, 4>(d_out, d_in, d_tw, iters, sms, clk_khz);
    return 0;
}
