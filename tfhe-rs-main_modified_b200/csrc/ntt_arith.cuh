// Device-side modular arithmetic classes for the B200 NTT engine (sm_100a).
//
// Each class implements one modulus family that the reference dispatches on
// (tfhe-ntt/src/prime64.rs:897-968, prime32.rs:797-843):
//
//   Shoup<T, HARVEY>  p < 2^(W-1)  lazy Shoup/Harvey butterflies; HARVEY (values in [0,4p))
//                     needs p < 2^(W-2)      (reference: prime64/less_than_6{2,3}bit.rs,
//                                              prime32/less_than_3{0,1}bit.rs)
//   Solinas64         p = 2^64-2^32+1        (reference: prime64/generic_solinas.rs:77-129)
//   Mont64            any other p >= 2^63    (reference: generic_solinas.rs:42-75, exact `%`)
//   Wide32            p >= 2^31              (reference: prime32/generic.rs:9-31, exact `%`)
//
// Every class ends a transform with canonical values in [0,p), so results are bit-identical
// to the reference for canonical inputs no matter which exact algorithm runs in between
// ("canonical-form invariants").
//
// Interface of an arithmetic class A:
//   A::T            element type (uint32_t / uint64_t)
//   A::TW           twiddle record as stored in device tables
//   A::Ctx          per-plan constants passed by value to kernels
//   fwd_bf / inv_bf lazy Cooley-Tukey / Gentleman-Sande butterflies
//   fwd_fin/inv_fin canonicalise after the last stage
//   mul_full        exact (a*b) mod p of two canonical values, canonical result (pointwise ops)
//   add_full        exact (a+b) mod p
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace nttb200 {

#define NTT_DEVINL __device__ __forceinline__

template <class T>
NTT_DEVINL T umin_(T a, T b) {
    return a < b ? a : b;
}
NTT_DEVINL uint32_t mulhi_(uint32_t a, uint32_t b) { return __umulhi(a, b); }
NTT_DEVINL uint64_t mulhi_(uint64_t a, uint64_t b) { return __umul64hi(a, b); }

// ------------------------------------------------------------------------------------
// Shoup / Harvey lazy arithmetic, W = 32 or 64
// ------------------------------------------------------------------------------------
template <class T_>
struct ShoupTw {
    T_ w, ws;  // w and floor(w * 2^W / p)
};

template <class T_, bool HARVEY>
struct Shoup {
    using T = T_;
    using TW = ShoupTw<T_>;
    struct Ctx {
        T p, two_p;
        // pointwise constants: Montgomery (W=64) or 64-bit Barrett (W=32)
        T pinv;     // p^-1 mod 2^W
        T r2;       // 2^(2W) mod p
        uint64_t barrett64;  // floor(2^64 / p), W=32 only
        // W=32, p < 2^30: one-word Barrett for products of canonical operands (barrett32_narrow)
        uint32_t bar_mu;     // floor(2^(k+31) / p), k = bit length of p
        uint32_t bar_shift;  // k - 1
    };
    static constexpr bool kHarvey = HARVEY;

    NTT_DEVINL static T csub(T a, T m) { return umin_<T>(a, a - m); }

    // t = b*w mod p in [0,2p) for any b < 2^W
    NTT_DEVINL static T mul_lazy(const Ctx& c, T b, TW w) {
        T q = mulhi_(b, w.ws);
        return b * w.w - q * c.p;
    }
    NTT_DEVINL static void fwd_bf(const Ctx& c, T& a, T& b, TW w) {
        if (HARVEY) {  // [0,4p) -> [0,4p)
            T z0 = csub(a, c.two_p);
            T t = mul_lazy(c, b, w);
            a = z0 + t;
            b = z0 - t + c.two_p;
        } else {  // [0,2p) -> [0,2p)
            T z0 = csub(a, c.p);
            T t = csub(mul_lazy(c, b, w), c.p);
            a = z0 + t;
            b = z0 - t + c.p;
        }
    }
    NTT_DEVINL static void inv_bf(const Ctx& c, T& a, T& b, TW w, bool /*b_canonical*/ = false) {
        if (HARVEY) {  // [0,2p) -> [0,2p)
            T y0 = csub(a + b, c.two_p);
            T t = a - b + c.two_p;
            a = y0;
            b = mul_lazy(c, t, w);
        } else {  // [0,p) -> [0,p)
            T y0 = csub(a + b, c.p);
            T t = a - b + c.p;
            a = y0;
            b = csub(mul_lazy(c, t, w), c.p);
        }
    }
    NTT_DEVINL static T fwd_fin(const Ctx& c, T a) {
        if (HARVEY) a = csub(a, c.two_p);
        return csub(a, c.p);
    }
    NTT_DEVINL static T inv_fin(const Ctx& c, T a) { return HARVEY ? csub(a, c.p) : a; }
    NTT_DEVINL static T inv_fin_prod(const Ctx& c, T a) { return inv_fin(c, a); }
    // multiply by a plan constant given as a Shoup pair, canonical result (used by normalize)
    NTT_DEVINL static T mul_const(const Ctx& c, T a, TW w) { return csub(mul_lazy(c, a, w), c.p); }

    NTT_DEVINL static T add_full(const Ctx& c, T a, T b) { return csub(a + b, c.p); }  // p < 2^(W-1)
    // running sum of canonical products: acc stays canonical
    NTT_DEVINL static T acc_add(const Ctx& c, T acc, T prod) { return add_full(c, acc, prod); }
    NTT_DEVINL static T acc_fin(const Ctx&, T acc) { return acc; }
    NTT_DEVINL static T mul_full(const Ctx& c, T a, T b);
    // Pointwise product against an operand that is reused many times (a GGSW shared by a batch):
    // pw_form() is applied to it once, mul_pw(x, pw_form(g)) = x*g mod p canonical.  W = 64: the
    // operand goes to Montgomery form so that one REDC suffices and x may be any lazy value.
    static constexpr bool kPwLazyIn = sizeof(T) == 8;
    NTT_DEVINL static T pw_form(const Ctx& c, T g);
    NTT_DEVINL static T mul_pw(const Ctx& c, T x, T gf);
};

// Montgomery REDC for odd p < 2^64: returns (hi:lo) * 2^-64 mod p, canonical, for hi < p.
NTT_DEVINL uint64_t redc64(uint64_t lo, uint64_t hi, uint64_t p, uint64_t pinv) {
    uint64_t m = lo * pinv;
    uint64_t t = __umul64hi(m, p);
    uint64_t r = hi - t;
    return hi < t ? r + p : r;
}

template <>
NTT_DEVINL uint64_t Shoup<uint64_t, true>::mul_full(const Ctx& c, uint64_t a, uint64_t b) {
    uint64_t x = redc64(a * b, __umul64hi(a, b), c.p, c.pinv);
    return redc64(x * c.r2, __umul64hi(x, c.r2), c.p, c.pinv);
}
template <>
NTT_DEVINL uint64_t Shoup<uint64_t, false>::mul_full(const Ctx& c, uint64_t a, uint64_t b) {
    uint64_t x = redc64(a * b, __umul64hi(a, b), c.p, c.pinv);
    return redc64(x * c.r2, __umul64hi(x, c.r2), c.p, c.pinv);
}
template <>
NTT_DEVINL uint64_t Shoup<uint64_t, true>::pw_form(const Ctx& c, uint64_t g) {
    return redc64(g * c.r2, __umul64hi(g, c.r2), c.p, c.pinv);
}
template <>
NTT_DEVINL uint64_t Shoup<uint64_t, false>::pw_form(const Ctx& c, uint64_t g) {
    return redc64(g * c.r2, __umul64hi(g, c.r2), c.p, c.pinv);
}
template <>
NTT_DEVINL uint64_t Shoup<uint64_t, true>::mul_pw(const Ctx& c, uint64_t x, uint64_t gf) {
    return redc64(x * gf, __umul64hi(x, gf), c.p, c.pinv);
}
template <>
NTT_DEVINL uint64_t Shoup<uint64_t, false>::mul_pw(const Ctx& c, uint64_t x, uint64_t gf) {
    return redc64(x * gf, __umul64hi(x, gf), c.p, c.pinv);
}
// exact a*b mod p for any p < 2^32 via a 64-bit Barrett quotient
NTT_DEVINL uint32_t barrett32(uint64_t d, uint32_t p, uint64_t b64) {
    uint64_t q = __umul64hi(d, b64);  // floor(d/p) - {0,1}
    uint64_t r = d - q * p;           // < 2p
    return (uint32_t)(r >= p ? r - p : r);
}
// exact d mod p for d < p^2, p < 2^30 (k = bit length of p): with s = floor(d / 2^(k-1)) < 2^32 and
// mu = floor(2^(k+31) / p) < 2^32 the estimate q = floor(s * mu / 2^32) satisfies
// floor(d/p) - 2 <= q <= floor(d/p)   (d / 2^(k+31) <= 1/2 and 2^(k-1) / p <= 1),
// so d - q*p lies in [0, 3p) < 2^32 and one 32-bit multiply-subtract plus two conditional
// subtractions finish.  One IMAD.HI + one IMAD against the four IMAD.WIDE of the 64-bit Barrett.
NTT_DEVINL uint32_t barrett32_narrow(uint64_t d, uint32_t p, uint32_t two_p, uint32_t mu, uint32_t shift) {
    uint32_t s = __funnelshift_r((uint32_t)d, (uint32_t)(d >> 32), shift);
    uint32_t q = __umulhi(s, mu);
    uint32_t r = (uint32_t)d - q * p;
    r = umin_<uint32_t>(r, r - two_p);
    return umin_<uint32_t>(r, r - p);
}
template <>
NTT_DEVINL uint32_t Shoup<uint32_t, true>::mul_full(const Ctx& c, uint32_t a, uint32_t b) {
    return barrett32_narrow((uint64_t)a * b, c.p, c.two_p, c.bar_mu, c.bar_shift);
}
template <>
NTT_DEVINL uint32_t Shoup<uint32_t, false>::mul_full(const Ctx& c, uint32_t a, uint32_t b) {
    return barrett32((uint64_t)a * b, c.p, c.barrett64);
}
template <>
NTT_DEVINL uint32_t Shoup<uint32_t, true>::pw_form(const Ctx&, uint32_t g) { return g; }
template <>
NTT_DEVINL uint32_t Shoup<uint32_t, false>::pw_form(const Ctx&, uint32_t g) { return g; }
template <>
NTT_DEVINL uint32_t Shoup<uint32_t, true>::mul_pw(const Ctx& c, uint32_t x, uint32_t gf) { return mul_full(c, x, gf); }
template <>
NTT_DEVINL uint32_t Shoup<uint32_t, false>::mul_pw(const Ctx& c, uint32_t x, uint32_t gf) { return mul_full(c, x, gf); }

// ------------------------------------------------------------------------------------
// Wide32: p >= 2^31, exact canonical arithmetic (the reference's prime32/generic.rs path).
// Round 2: twiddles in Montgomery form (w * 2^32 mod p), one 32-bit REDC per product -- IMAD.WIDE,
// IMAD, IMAD.HI and a conditional add instead of a Shoup quotient with a 64-bit remainder -- and
// add / sub through the complement (p - b) so that nothing needs a 33rd bit: 13 instead of 19
// instructions per butterfly.
// ------------------------------------------------------------------------------------
struct Wide32 {
    using T = uint32_t;
    using TW = uint32_t;  // w * 2^32 mod p
    struct Ctx {
        uint32_t p, two_p, pinv, r2;  // pinv = p^-1 mod 2^32, r2 = 2^64 mod p
        uint64_t barrett64;
    };
    static constexpr bool kHarvey = false;
    // a * wm * 2^-32 mod p, canonical, for any 32-bit a and wm < p
    NTT_DEVINL static T mul_exact(const Ctx& c, T a, TW wm) {
        uint64_t t = (uint64_t)a * wm;
        uint32_t lo = (uint32_t)t, hi = (uint32_t)(t >> 32);  // hi <= wm - 1 < p
        uint32_t u = __umulhi(lo * c.pinv, c.p);              // < p
        uint32_t r = hi - u;
        return hi < u ? r + c.p : r;
    }
    NTT_DEVINL static T add_full(const Ctx& c, T a, T b) {  // a, b < p
        uint32_t nb = c.p - b;                               // in (0, p]
        return a >= nb ? a - nb : a + b;                     // a + b < p when a < p - b: no 33rd bit
    }
    NTT_DEVINL static T sub_full(const Ctx& c, T a, T b) {
        uint32_t d = a - b;
        return a >= b ? d : d + c.p;
    }
    NTT_DEVINL static void fwd_bf(const Ctx& c, T& a, T& b, TW w) {
        T t = mul_exact(c, b, w), z0 = a;
        a = add_full(c, z0, t);
        b = sub_full(c, z0, t);
    }
    NTT_DEVINL static void inv_bf(const Ctx& c, T& a, T& b, TW w, bool /*b_canonical*/ = false) {
        T s = add_full(c, a, b), d = sub_full(c, a, b);
        a = s;
        b = mul_exact(c, d, w);
    }
    NTT_DEVINL static T fwd_fin(const Ctx&, T a) { return a; }
    NTT_DEVINL static T inv_fin(const Ctx&, T a) { return a; }
    NTT_DEVINL static T inv_fin_prod(const Ctx&, T a) { return a; }
    NTT_DEVINL static T mul_const(const Ctx& c, T a, TW wm) { return mul_exact(c, a, wm); }
    NTT_DEVINL static T mul_full(const Ctx& c, T a, T b) {
        return barrett32((uint64_t)a * b, c.p, c.barrett64);
    }
    NTT_DEVINL static T acc_add(const Ctx& c, T acc, T prod) { return add_full(c, acc, prod); }
    NTT_DEVINL static T acc_fin(const Ctx&, T acc) { return acc; }
    static constexpr bool kPwLazyIn = false;
    NTT_DEVINL static T pw_form(const Ctx&, T g) { return g; }
    NTT_DEVINL static T mul_pw(const Ctx& c, T x, T gf) { return mul_full(c, x, gf); }
};

// ------------------------------------------------------------------------------------
// Solinas64: p = 2^64 - 2^32 + 1.  Values travel as arbitrary 64-bit representatives
// (2^64 == 2^32-1 =: eps mod p) and are canonicalised once at the end.
// ------------------------------------------------------------------------------------
struct Solinas64 {
    using T = uint64_t;
    using TW = uint64_t;  // twiddle in Montgomery form: w * 2^64 mod p
    struct Ctx {
        uint64_t p;  // compile-time constant in the butterflies; kept for a uniform interface
    };
    static constexpr bool kHarvey = false;
    static constexpr uint64_t P = 0xFFFFFFFF00000001ull;
    static constexpr uint64_t EPS = 0xFFFFFFFFull;

    // Montgomery product b * wm * 2^-64 mod p, canonical, for any 64-bit b and wm < p.
    // With p = 2^64 - 2^32 + 1 the REDC needs no multiplication (p^-1 = 2^32 + 1 mod 2^64):
    //   T = b*wm = (h1 : h0 : x1 : x0);  m1 = (x0 + x1) mod 2^32, c = its carry
    //   S = m1*(2^32-1) + x0 - c = ((m1 + c) << 32) - (x1 + c)          (0 <= S < 2^64)
    //   r = (h1:h0) - S = ((h1 + k - (m1 + c)) : r0)  with (k : r0) = h0 + x1 + c
    //   r in (-p, p); negative (borrow in the high word) -> + p, i.e. - (2^32-1) wrapping.
    // h1 + k cannot overflow because the high half of b*wm is at most p - 2.  Eight carry-chain
    // instructions; add.cc/addc and sub.cc/subc chains are never mixed.
    NTT_DEVINL static uint64_t mulm(uint64_t b, uint64_t wm) {
        unsigned __int128 t = (unsigned __int128)b * wm;  // 4 IMAD.WIDE.U32 + 3
        uint64_t lo = (uint64_t)t, hi = (uint64_t)(t >> 64), r;
        asm("{ .reg .u32 x0,x1,h0,h1,m1,tt,m;\n\t"
            "mov.b64 {x0,x1}, %1; mov.b64 {h0,h1}, %2;\n\t"
            "add.cc.u32 m1,x0,x1; madc.lo.u32 tt,m1,1,0;\n\t"
            "addc.cc.u32 h0,h0,x1; madc.lo.u32 h1,h1,1,0;\n\t"
            "sub.cc.u32 h1,h1,tt; subc.u32 m,0,0;\n\t"
            "sub.cc.u32 h0,h0,m; subc.u32 h1,h1,0;\n\t"
            "mov.b64 %0,{h0,h1}; }"
            : "=l"(r)
            : "l"(lo), "l"(hi));
        return r;
    }
    // a arbitrary 64-bit representative, b <= p: a + b with a wrap folded as +eps
    // (s + 2^64 == s + eps; cannot wrap twice because b <= p)
    NTT_DEVINL static uint64_t add_lazy(uint64_t a, uint64_t b) {
        uint64_t r;
        asm("{ .reg .u32 a0,a1,b0,b1,c;\n\t"
            "mov.b64 {a0,a1}, %1; mov.b64 {b0,b1}, %2;\n\t"
            "add.cc.u32 a0,a0,b0; addc.cc.u32 a1,a1,b1; madc.lo.u32 c,0,0,0;\n\t"
            "add.u32 a1,a1,c; sub.cc.u32 a0,a0,c; subc.u32 a1,a1,0;\n\t"
            "mov.b64 %0, {a0,a1}; }"
            : "=l"(r)
            : "l"(a), "l"(b));
        return r;
    }
    // a arbitrary, b <= p: a - b with a borrow folded as -eps (cannot underflow twice)
    NTT_DEVINL static uint64_t sub_lazy(uint64_t a, uint64_t b) {
        uint64_t r;
        asm("{ .reg .u32 a0,a1,b0,b1,m;\n\t"
            "mov.b64 {a0,a1}, %1; mov.b64 {b0,b1}, %2;\n\t"
            "sub.cc.u32 a0,a0,b0; subc.cc.u32 a1,a1,b1; subc.u32 m,0,0;\n\t"
            "sub.cc.u32 a0,a0,m; subc.u32 a1,a1,0;\n\t"
            "mov.b64 %0, {a0,a1}; }"
            : "=l"(r)
            : "l"(a), "l"(b));
        return r;
    }
    // canonical representative of an arbitrary 64-bit value: a >= p <=> a + eps wraps
    NTT_DEVINL static uint64_t canon(uint64_t a) {
        // a >= p  <=>  high word all ones and low word nonzero; then a - p = (0 : lo - 1)
        uint64_t r;
        asm("{ .reg .pred q0,q1; .reg .u32 lo,hi;\n\t"
            "mov.b64 {lo,hi}, %1;\n\t"
            "setp.eq.u32 q0,hi,0xFFFFFFFF; setp.ne.and.u32 q1,lo,0,q0;\n\t"
            "@q1 add.u32 lo,lo,0xFFFFFFFF; @q1 mov.u32 hi,0;\n\t"
            "mov.b64 %0, {lo,hi}; }"
            : "=l"(r)
            : "l"(a));
        return r;
    }
    // plain (non-Montgomery) product of two arbitrary 64-bit values, canonical:
    // the reference's fold (generic_solinas.rs:102-128) with 2^64 == eps, 2^96 == -1
    NTT_DEVINL static uint64_t mul_plain(uint64_t a, uint64_t b) {
        unsigned __int128 t = (unsigned __int128)a * b;
        uint64_t lo = (uint64_t)t, hi = (uint64_t)(t >> 64);
        uint32_t hh = (uint32_t)(hi >> 32), mid = (uint32_t)hi;
        uint64_t t0 = lo - hh;
        if (lo < (uint64_t)hh) t0 -= EPS;
        uint64_t m = ((uint64_t)mid << 32) - mid;
        uint64_t t1 = t0 + m;
        if (t1 < m) t1 += EPS;
        return canon(t1);
    }
    NTT_DEVINL static void fwd_bf(const Ctx&, T& a, T& b, TW w) {
        T t = mulm(b, w), z0 = a;
        a = add_lazy(z0, t);
        b = sub_lazy(z0, t);
    }
    // Gentleman-Sande: both inputs may be arbitrary representatives, so bring b into [0,p) first
    // unless the caller knows it already is (b is the product output of the previous stage).
    NTT_DEVINL static void inv_bf(const Ctx&, T& a, T& b, TW w, bool b_canonical = false) {
        T bc = b_canonical ? b : canon(b), z0 = a;
        a = add_lazy(z0, bc);
        b = mulm(sub_lazy(z0, bc), w);
    }
    NTT_DEVINL static T fwd_fin(const Ctx&, T a) { return canon(a); }
    NTT_DEVINL static T inv_fin(const Ctx&, T a) { return canon(a); }
    NTT_DEVINL static T inv_fin_prod(const Ctx&, T a) { return a; }  // mulm output is canonical
    NTT_DEVINL static T mul_const(const Ctx&, T a, TW wm) { return mulm(a, wm); }
    NTT_DEVINL static T add_full(const Ctx&, T a, T b) { return canon(add_lazy(a, b)); }
    NTT_DEVINL static T mul_full(const Ctx&, T a, T b) { return mul_plain(a, b); }
    // running sum: arbitrary representative + canonical product, canonicalised once at the end
    NTT_DEVINL static T acc_add(const Ctx&, T acc, T prod) { return add_lazy(acc, prod); }
    NTT_DEVINL static T acc_fin(const Ctx&, T acc) { return canon(acc); }
    // Montgomery form of a reused pointwise operand: g * 2^64 = g * eps mod p; then one REDC per product
    static constexpr bool kPwLazyIn = true;
    NTT_DEVINL static T pw_form(const Ctx&, T g) { return mul_plain(g, EPS); }
    NTT_DEVINL static T mul_pw(const Ctx&, T x, T gf) { return mulm(x, gf); }
};

// ------------------------------------------------------------------------------------
// Mont64: any odd prime p >= 2^63 other than Solinas.  Twiddles are stored in Montgomery
// form (w * 2^64 mod p) so that one REDC gives the exact canonical product.
// Round 2: values travel as arbitrary 64-bit representatives, like Solinas64, with the wrap of a
// sum folded as + (2^64 - p) and the wrap of a difference as + p (both modulo 2^64): for b < p a
// single fold can never wrap again (a + b - 2^64 <= p - 2, and a - b + 2^64 >= 2^64 - p + 1), so a
// butterfly needs no comparison at all -- 5 + 5 carry-chain instructions instead of 10 + 6 with
// compare / select.  The REDC accepts any 64-bit multiplicand (its high product word stays below p).
// ------------------------------------------------------------------------------------
struct Mont64 {
    using T = uint64_t;
    using TW = uint64_t;  // w * 2^64 mod p
    struct Ctx {
        uint64_t p, pinv, r2;
    };
    static constexpr bool kHarvey = false;
    NTT_DEVINL static T mul_exact(const Ctx& c, T a, TW wm) {
        return redc64(a * wm, __umul64hi(a, wm), c.p, c.pinv);
    }
    // a arbitrary, b < p
    NTT_DEVINL static T add_lazy(const Ctx& c, T a, T b) {
        const uint64_t cc = 0 - c.p;  // 2^64 - p
        uint64_t r;
        asm("{ .reg .u32 a0,a1,b0,b1,c0,c1,k;\n\t"
            "mov.b64 {a0,a1}, %1; mov.b64 {b0,b1}, %2; mov.b64 {c0,c1}, %3;\n\t"
            "add.cc.u32 a0,a0,b0; addc.cc.u32 a1,a1,b1; addc.u32 k,0,0;\n\t"
            "mad.lo.cc.u32 a0,k,c0,a0; madc.lo.u32 a1,k,c1,a1;\n\t"
            "mov.b64 %0, {a0,a1}; }"
            : "=l"(r)
            : "l"(a), "l"(b), "l"(cc));
        return r;
    }
    NTT_DEVINL static T sub_lazy(const Ctx& c, T a, T b) {
        uint64_t r;
        asm("{ .reg .u32 a0,a1,b0,b1,p0,p1,k;\n\t"
            "mov.b64 {a0,a1}, %1; mov.b64 {b0,b1}, %2; mov.b64 {p0,p1}, %3;\n\t"
            "sub.cc.u32 a0,a0,b0; subc.cc.u32 a1,a1,b1; subc.u32 k,0,0;\n\t"  // k = 0 or 0xFFFFFFFF
            "and.b32 p0,p0,k; and.b32 p1,p1,k;\n\t"
            "add.cc.u32 a0,a0,p0; addc.u32 a1,a1,p1;\n\t"
            "mov.b64 %0, {a0,a1}; }"
            : "=l"(r)
            : "l"(a), "l"(b), "l"(c.p));
        return r;
    }
    NTT_DEVINL static T canon(const Ctx& c, T a) { return a >= c.p ? a - c.p : a; }  // a < 2^64 < 2p
    NTT_DEVINL static T add_full(const Ctx& c, T a, T b) { return canon(c, add_lazy(c, a, b)); }
    NTT_DEVINL static void fwd_bf(const Ctx& c, T& a, T& b, TW w) {
        T t = mul_exact(c, b, w), z0 = a;
        a = add_lazy(c, z0, t);
        b = sub_lazy(c, z0, t);
    }
    // Gentleman-Sande: b is brought into [0, p) unless the caller knows it is (a product of the previous stage)
    NTT_DEVINL static void inv_bf(const Ctx& c, T& a, T& b, TW w, bool b_canonical = false) {
        T bc = b_canonical ? b : canon(c, b), z0 = a;
        a = add_lazy(c, z0, bc);
        b = mul_exact(c, sub_lazy(c, z0, bc), w);
    }
    NTT_DEVINL static T fwd_fin(const Ctx& c, T a) { return canon(c, a); }
    NTT_DEVINL static T inv_fin(const Ctx& c, T a) { return canon(c, a); }
    NTT_DEVINL static T inv_fin_prod(const Ctx&, T a) { return a; }  // REDC output is canonical
    NTT_DEVINL static T mul_const(const Ctx& c, T a, TW wm) { return mul_exact(c, a, wm); }
    NTT_DEVINL static T mul_full(const Ctx& c, T a, T b) {
        uint64_t x = redc64(a * b, __umul64hi(a, b), c.p, c.pinv);
        return redc64(x * c.r2, __umul64hi(x, c.r2), c.p, c.pinv);
    }
    NTT_DEVINL static T acc_add(const Ctx& c, T acc, T prod) { return add_lazy(c, acc, prod); }
    NTT_DEVINL static T acc_fin(const Ctx& c, T acc) { return canon(c, acc); }
    static constexpr bool kPwLazyIn = false;  // the pointwise step canonicalises the transform output first
    NTT_DEVINL static T pw_form(const Ctx& c, T g) { return redc64(g * c.r2, __umul64hi(g, c.r2), c.p, c.pinv); }
    NTT_DEVINL static T mul_pw(const Ctx& c, T x, T gf) { return mul_exact(c, x, gf); }
};

}  // namespace nttb200
