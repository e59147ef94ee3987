/*
 * tfhe_ntt_custum_radix_oracle.c -- TEST INFRASTRUCTURE ONLY (see tfhe_ntt_oracle.h).
 *
 * Literal C restatement of the fork's `tfhe_ntt::custum_radix` module
 * (/root/reference/tfhe-ntt/src/custum_radix/{fwd.rs,inv.rs,fwd_1.rs}): recursive radix-2, radix-4
 * and split-radix CYCLIC transforms of u32 vectors (natural order in and out) over a caller-supplied
 * table tw[k] = root^k, including the quirks of the recursions (which bases scale, which do not) and
 * the multiplication counters of the `_mut` variants.  The recursion allocates per level exactly as
 * the reference does; nothing here is meant to be fast.
 *
 * Parity pinning: the reference holds no known-answer vector for this module (its only test,
 * fwd_1.rs:433-463, prints).  tests/test_custum_radix.py pins this file against an arbitrary-
 * precision Python restatement of the definition (X[k] = sum_j a[j] root^(jk)) on that test's
 * input (n = 64, p = 65537) and on random vectors.
 */
#include <stdlib.h>
#include <string.h>

#include "tfhe_ntt_oracle.h"

/* fwd.rs:1-19 */
static inline uint32_t add_mod(uint32_t a, uint32_t b, uint32_t p) {
    uint64_t s = (uint64_t)a + b;
    return s >= p ? (uint32_t)(s - p) : (uint32_t)s;
}
static inline uint32_t sub_mod(uint32_t a, uint32_t b, uint32_t p) {
    return a >= b ? a - b : (uint32_t)((uint64_t)a + p - b);
}
static inline uint32_t mul_mod(uint32_t a, uint32_t b, uint32_t p) {
    return (uint32_t)(((uint64_t)a * b) % p);
}
/* fwd_1.rs:28-37 */
static inline uint32_t mul_mod_counted(uint32_t a, uint32_t b, uint32_t p, tfo_mult_stats *st) {
    if (a != 0 && b != 0)
        st->nonzero_mults += 1;
    else
        st->skipped_mults += 1;
    return mul_mod(a, b, p);
}

/* fwd.rs:22-34 */
uint32_t tfo_cr_pow_mod(uint32_t base, uint32_t exp, uint32_t p) {
    uint32_t res = 1;
    base %= p;
    while (exp > 0) {
        if (exp & 1) res = mul_mod(res, base, p);
        base = mul_mod(base, base, p);
        exp >>= 1;
    }
    return res;
}
/* fwd.rs:37-39 */
uint32_t tfo_cr_mod_inverse(uint32_t a, uint32_t p) { return tfo_cr_pow_mod(a, p - 2, p); }

/* fwd.rs:42-68: smallest generator of (Z/p)^*; 0 where the reference panics */
uint32_t tfo_cr_compute_primitive_root(uint32_t p) {
    uint32_t m = p - 1, factors[32];
    int nf = 0;
    for (uint32_t i = 2; (uint64_t)i * i <= m; ++i)
        if (m % i == 0) {
            factors[nf++] = i;
            while (m % i == 0) m /= i;
        }
    if (m > 1) factors[nf++] = m;
    for (uint32_t g = 2; g < p; ++g) {
        int ok = 1;
        for (int k = 0; k < nf && ok; ++k)
            if (tfo_cr_pow_mod(g, (p - 1) / factors[k], p) == 1) ok = 0;
        if (ok) return g;
    }
    return 0;
}

/* fwd.rs:72-93 make_twiddles: tw[k] = root^k with root = g^((p-1)/n); 0 on the reference's asserts */
int tfo_cr_make_twiddles(size_t n, uint32_t p, uint32_t *tw) {
    if (n == 0 || (n & (n - 1)) || (p - 1) % n != 0) return 0;
    uint32_t g = tfo_cr_compute_primitive_root(p);
    if (!g) return 0;
    uint32_t root = tfo_cr_pow_mod(g, (uint32_t)((p - 1) / n), p), cur = 1;
    for (size_t k = 0; k < n; ++k) {
        tw[k] = cur;
        cur = mul_mod(cur, root, p);
    }
    return 1;
}
/* fwd.rs:96-103 make_inv_twiddles */
void tfo_cr_make_inv_twiddles(const uint32_t *tw, size_t n, uint32_t p, uint32_t *inv) {
    for (size_t k = 0; k < n; ++k) inv[k] = tfo_cr_mod_inverse(tw[k], p);
}

static uint32_t *subsample(const uint32_t *tw, size_t n, size_t step, size_t count) {
    uint32_t *t = (uint32_t *)calloc(count ? count : 1, sizeof(uint32_t));
    for (size_t k = 0; k < count; ++k) t[k] = tw[(step * k) % n];
    return t;
}
static uint32_t *gather(const uint32_t *a, size_t stride, size_t off, size_t count) {
    uint32_t *t = (uint32_t *)calloc(count ? count : 1, sizeof(uint32_t));
    for (size_t i = 0; i < count; ++i) t[i] = a[stride * i + off];
    return t;
}
static void scale(uint32_t *a, size_t n, uint32_t f, uint32_t p) {
    for (size_t i = 0; i < n; ++i) a[i] = mul_mod(a[i], f, p);
}
static void base2(uint32_t *a, uint32_t p) {
    uint32_t t = a[0];
    a[0] = add_mod(a[0], a[1], p);
    a[1] = sub_mod(t, a[1], p);
}

/* the 4-output combine shared by fwd.rs:139-167 and inv.rs:141-169 (tw indexes modulo n) */
static void radix4_combine(uint32_t *a, size_t n, const uint32_t *a0, const uint32_t *a1, const uint32_t *a2,
                           const uint32_t *a3, const uint32_t *tw, uint32_t p, tfo_mult_stats *st) {
    size_t q = n / 4;
    for (size_t i = 0; i < q; ++i)
        for (size_t r = 0; r < 4; ++r) {
            size_t e = i + r * q;
            uint32_t t1, t2, t3;
            if (st) {
                t1 = mul_mod_counted(tw[e % n], a1[i], p, st);
                t2 = mul_mod_counted(tw[(2 * e) % n], a2[i], p, st);
                t3 = mul_mod_counted(tw[(3 * e) % n], a3[i], p, st);
            } else {
                t1 = mul_mod(tw[e % n], a1[i], p);
                t2 = mul_mod(tw[(2 * e) % n], a2[i], p);
                t3 = mul_mod(tw[(3 * e) % n], a3[i], p);
            }
            a[e] = add_mod(add_mod(a0[i], t1, p), add_mod(t2, t3, p), p);
        }
}

/* fwd.rs:105-168 */
void tfo_cr_fft_radix4_recursive(uint32_t *a, size_t n, const uint32_t *tw, uint32_t p) {
    if (n == 1) return;
    if (n == 2) {
        base2(a, p);
        return;
    }
    size_t q = n / 4;
    uint32_t *a0 = gather(a, 4, 0, q), *a1 = gather(a, 4, 1, q), *a2 = gather(a, 4, 2, q), *a3 = gather(a, 4, 3, q);
    uint32_t *tw4 = subsample(tw, n, 4, q);
    tfo_cr_fft_radix4_recursive(a0, q, tw4, p);
    tfo_cr_fft_radix4_recursive(a1, q, tw4, p);
    tfo_cr_fft_radix4_recursive(a2, q, tw4, p);
    tfo_cr_fft_radix4_recursive(a3, q, tw4, p);
    radix4_combine(a, n, a0, a1, a2, a3, tw, p, NULL);
    free(a0), free(a1), free(a2), free(a3), free(tw4);
}

/* fwd.rs:170-205 (st == NULL) and fwd_1.rs:190-230 (counted) */
static void fft_radix2(uint32_t *a, size_t n, const uint32_t *tw, uint32_t p, tfo_mult_stats *st) {
    if (n == 1) return;
    if (n == 2) {
        if (st) st->nonzero_mults += (a[1] != 0); /* fwd_1.rs:195-196 counts the implicit product */
        base2(a, p);
        return;
    }
    size_t h = n / 2;
    uint32_t *even = gather(a, 2, 0, h), *odd = gather(a, 2, 1, h), *tw2 = subsample(tw, n, 2, h);
    fft_radix2(even, h, tw2, p, st);
    fft_radix2(odd, h, tw2, p, st);
    for (size_t k = 0; k < h; ++k) {
        uint32_t t = st ? mul_mod_counted(odd[k], tw[k], p, st) : mul_mod(odd[k], tw[k], p);
        a[k] = add_mod(even[k], t, p);
        a[k + h] = sub_mod(even[k], t, p);
    }
    free(even), free(odd), free(tw2);
}
void tfo_cr_fft_radix2_recursive(uint32_t *a, size_t n, const uint32_t *tw, uint32_t p) { fft_radix2(a, n, tw, p, NULL); }

/* fwd.rs:207-272 (st == NULL), inv.rs:232-302 (the same body on the inverse table) and
 * fwd_1.rs:232-294 (counted) */
static void split_radix(uint32_t *a, size_t n, const uint32_t *tw, uint32_t p, tfo_mult_stats *st) {
    if (n == 1) return;
    if (n == 2) {
        base2(a, p);
        return;
    }
    size_t n2 = n / 2, n4 = n / 4;
    uint32_t *a0 = gather(a, 2, 0, n2), *a1 = gather(a, 4, 1, n4), *a2 = gather(a, 4, 3, n4);
    uint32_t *tw2 = subsample(tw, n, 2, n2), *tw4 = subsample(tw, n, 4, n4);
    split_radix(a0, n2, tw2, p, st);
    split_radix(a1, n4, tw4, p, st);
    split_radix(a2, n4, tw4, p, st);
    uint32_t j = tw[n4 % n];
    for (size_t k = 0; k < n4; ++k) {
        uint32_t wk = tw[k % n], w3k = tw[(3 * k) % n], t1, t2, jd;
        if (st) {
            t1 = mul_mod_counted(a1[k], wk, p, st);
            t2 = mul_mod_counted(a2[k], w3k, p, st);
        } else {
            t1 = mul_mod(a1[k], wk, p);
            t2 = mul_mod(a2[k], w3k, p);
        }
        uint32_t sum = add_mod(t1, t2, p), diff = sub_mod(t1, t2, p);
        jd = st ? mul_mod_counted(diff, j, p, st) : mul_mod(diff, j, p);
        uint32_t u0 = a0[k], u1 = a0[k + n4];
        a[k] = add_mod(u0, sum, p);
        a[k + n4] = add_mod(u1, jd, p);
        a[k + n2] = sub_mod(u0, sum, p);
        a[k + n2 + n4] = sub_mod(u1, jd, p);
    }
    free(a0), free(a1), free(a2), free(tw2), free(tw4);
}
void tfo_cr_fft_split_radix_recursive(uint32_t *a, size_t n, const uint32_t *tw, uint32_t p) {
    split_radix(a, n, tw, p, NULL);
}

/* inv.rs:106-176: the n == 2 base halves, the children's table is inv_tw^4 at full length */
void tfo_cr_ifft_radix4_recursive(uint32_t *a, size_t n, const uint32_t *inv_tw, uint32_t p, uint32_t n_inv, int top) {
    if (n == 1) return;
    if (n == 2) {
        uint32_t tmp = a[0], inv2 = tfo_cr_mod_inverse(2, p);
        a[0] = mul_mod(add_mod(a[0], a[1], p), inv2, p);
        a[1] = mul_mod(sub_mod(tmp, a[1], p), inv2, p);
        return;
    }
    size_t q = n / 4;
    uint32_t *a0 = gather(a, 4, 0, q), *a1 = gather(a, 4, 1, q), *a2 = gather(a, 4, 2, q), *a3 = gather(a, 4, 3, q);
    uint32_t *tw4 = (uint32_t *)calloc(n, sizeof(uint32_t));
    for (size_t i = 0; i < n; ++i) tw4[i] = tfo_cr_pow_mod(inv_tw[i], 4, p);
    tfo_cr_ifft_radix4_recursive(a0, q, tw4, p, n_inv, 0);
    tfo_cr_ifft_radix4_recursive(a1, q, tw4, p, n_inv, 0);
    tfo_cr_ifft_radix4_recursive(a2, q, tw4, p, n_inv, 0);
    tfo_cr_ifft_radix4_recursive(a3, q, tw4, p, n_inv, 0);
    radix4_combine(a, n, a0, a1, a2, a3, inv_tw, p, NULL);
    if (top) scale(a, n, n_inv, p);
    free(a0), free(a1), free(a2), free(a3), free(tw4);
}

/* inv.rs:178-230 and fwd_1.rs:381-428: the bases return before the scaling */
void tfo_cr_ifft_radix2_recursive(uint32_t *a, size_t n, const uint32_t *inv_tw, uint32_t p, uint32_t n_inv, int top) {
    if (n <= 2) {
        fft_radix2(a, n, inv_tw, p, NULL);
        return;
    }
    fft_radix2(a, n, inv_tw, p, NULL);
    if (top) scale(a, n, n_inv, p);
}

/* inv.rs:232-303 */
void tfo_cr_ifft_split_radix_recursive(uint32_t *a, size_t n, const uint32_t *inv_tw, uint32_t p, uint32_t n_inv,
                                       int top) {
    split_radix(a, n, inv_tw, p, NULL);
    if (n > 2 && top) scale(a, n, n_inv, p);
}

/* fwd_1.rs:102-188: n == 4 is a base of its own, the combine multiplies by tw[n/4] */
void tfo_cr_fft_radix4_recursive_mut(uint32_t *a, size_t n, const uint32_t *tw, uint32_t p, tfo_mult_stats *st) {
    if (n == 1) return;
    if (n == 2) {
        base2(a, p);
        return;
    }
    if (n == 4) {
        uint32_t t0 = add_mod(a[0], a[2], p), t1 = sub_mod(a[0], a[2], p);
        uint32_t t2 = add_mod(a[1], a[3], p), t3 = sub_mod(a[1], a[3], p);
        uint32_t r = mul_mod_counted(tw[1], t3, p, st);
        a[0] = add_mod(t0, t2, p);
        a[1] = add_mod(t1, r, p);
        a[2] = sub_mod(t0, t2, p);
        a[3] = sub_mod(t1, r, p);
        return;
    }
    size_t q = n / 4;
    uint32_t *a0 = gather(a, 4, 0, q), *a1 = gather(a, 4, 1, q), *a2 = gather(a, 4, 2, q), *a3 = gather(a, 4, 3, q);
    uint32_t *tw4 = subsample(tw, n, 4, q);
    tfo_cr_fft_radix4_recursive_mut(a0, q, tw4, p, st);
    tfo_cr_fft_radix4_recursive_mut(a1, q, tw4, p, st);
    tfo_cr_fft_radix4_recursive_mut(a2, q, tw4, p, st);
    tfo_cr_fft_radix4_recursive_mut(a3, q, tw4, p, st);
    for (size_t k = 0; k < q; ++k) {
        uint32_t t1 = mul_mod_counted(tw[k], a1[k], p, st);
        uint32_t t2 = mul_mod_counted(tw[(2 * k) % n], a2[k], p, st);
        uint32_t t3 = mul_mod_counted(tw[(3 * k) % n], a3[k], p, st);
        uint32_t b0 = add_mod(a0[k], t2, p), b1 = sub_mod(a0[k], t2, p);
        uint32_t b2 = add_mod(t1, t3, p), b3 = sub_mod(t1, t3, p);
        uint32_t b3r = mul_mod_counted(tw[q], b3, p, st);
        a[k] = add_mod(b0, b2, p);
        a[k + q] = add_mod(b1, b3r, p);
        a[k + 2 * q] = sub_mod(b0, b2, p);
        a[k + 3 * q] = sub_mod(b1, b3r, p);
    }
    free(a0), free(a1), free(a2), free(a3), free(tw4);
}
void tfo_cr_fft_radix2_recursive_mut(uint32_t *a, size_t n, const uint32_t *tw, uint32_t p, tfo_mult_stats *st) {
    fft_radix2(a, n, tw, p, st);
}
void tfo_cr_fft_split_radix_recursive_mut(uint32_t *a, size_t n, const uint32_t *tw, uint32_t p, tfo_mult_stats *st) {
    split_radix(a, n, tw, p, st);
}

/* fwd_1.rs:296-379: subsampled child tables, no halving, the bases scale when top */
void tfo_cr_ifft_radix4_recursive_mut(uint32_t *a, size_t n, const uint32_t *inv_tw, uint32_t p, uint32_t n_inv,
                                      int top, tfo_mult_stats *st) {
    if (n == 1) {
        if (top) a[0] = mul_mod(a[0], n_inv, p);
        return;
    }
    if (n == 2) {
        base2(a, p);
        if (top) scale(a, 2, n_inv, p);
        return;
    }
    size_t q = n / 4;
    uint32_t *a0 = gather(a, 4, 0, q), *a1 = gather(a, 4, 1, q), *a2 = gather(a, 4, 2, q), *a3 = gather(a, 4, 3, q);
    uint32_t *tw4 = subsample(inv_tw, n, 4, q);
    tfo_cr_ifft_radix4_recursive_mut(a0, q, tw4, p, n_inv, 0, st);
    tfo_cr_ifft_radix4_recursive_mut(a1, q, tw4, p, n_inv, 0, st);
    tfo_cr_ifft_radix4_recursive_mut(a2, q, tw4, p, n_inv, 0, st);
    tfo_cr_ifft_radix4_recursive_mut(a3, q, tw4, p, n_inv, 0, st);
    radix4_combine(a, n, a0, a1, a2, a3, inv_tw, p, st);
    if (top) scale(a, n, n_inv, p);
    free(a0), free(a1), free(a2), free(a3), free(tw4);
}
