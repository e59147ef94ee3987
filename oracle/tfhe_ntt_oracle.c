/*
 * tfhe_ntt_oracle.c -- TEST INFRASTRUCTURE ONLY (see tfhe_ntt_oracle.h).
 *
 * Plain-C restatement of the *scalar* code paths of the reference crate
 * /root/reference/tfhe-ntt/src.  Nothing in the product
 * (tfhe-rs-main_modified_b200/) links or calls this file.
 *
 * Where the reference divides through its Lemire Div32/Div64 helpers
 * (fastdiv.rs:43-150) this file uses the C `%` / `/` operators on
 * unsigned __int128: the helpers compute exact quotients/remainders
 * (fastdiv.rs tests :159-195), so the values are identical.
 */
#include "tfhe_ntt_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef tfo_u128 u128;

#define SOLINAS_P 0xFFFFFFFF00000001ull /* prime64.rs:8 */

/* ------------------------------------------------------------------ */
/* prime.rs                                                            */
/* ------------------------------------------------------------------ */

/* prime.rs:8-10 */
uint64_t tfo_mul_mod64(uint64_t p, uint64_t a, uint64_t b) { return (uint64_t)(((u128)a * b) % p); }
/* prime.rs:4-6 */
uint32_t tfo_mul_mod32(uint32_t p, uint32_t a, uint32_t b) {
    return (uint32_t)(((uint64_t)a * b) % p);
}

/* prime.rs:31-48 */
uint64_t tfo_exp_mod64(uint64_t p, uint64_t base, uint64_t pow) {
    if (pow == 0) return 1;
    uint64_t y = 1, x = base;
    while (pow > 1) {
        if (pow % 2 == 1) y = tfo_mul_mod64(p, x, y);
        x = tfo_mul_mod64(p, x, x);
        pow /= 2;
    }
    return tfo_mul_mod64(p, x, y);
}

/* prime.rs:12-29 */
uint32_t tfo_exp_mod32(uint32_t p, uint32_t base, uint32_t pow) {
    if (pow == 0) return 1;
    uint32_t y = 1, x = base;
    while (pow > 1) {
        if (pow % 2 == 1) y = tfo_mul_mod32(p, x, y);
        x = tfo_mul_mod32(p, x, x);
        pow /= 2;
    }
    return tfo_mul_mod32(p, x, y);
}

/* prime.rs:50-66 */
static int miller_rabin_iter(uint64_t n, uint64_t s, uint64_t d, uint64_t a) {
    uint64_t x = tfo_exp_mod64(n, a, d);
    uint64_t n_minus_1 = n - 1;
    if (x == 1 || x == n_minus_1) return 1;
    uint64_t count = 0;
    while (count < s - 1) {
        x = tfo_mul_mod64(n, x, x);
        if (x == n_minus_1) return 1;
        count += 1;
    }
    return 0;
}

/* prime.rs:76-126 */
int tfo_is_prime64(uint64_t n) {
    static const uint64_t small[12] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    if (n < 2) return 0;
    for (int i = 0; i < 12; i++)
        if (n % small[i] == 0) return n == small[i];
    uint64_t s = 0, d = n - 1;
    while (d % 2 == 0) {
        s += 1;
        d /= 2;
    }
    for (int i = 0; i < 12; i++)
        if (!miller_rabin_iter(n, s, d, small[i])) return 0;
    return 1;
}

/* prime.rs:130-186 */
int tfo_largest_prime_in_arithmetic_progression64(uint64_t factor, uint64_t offset, uint64_t lo,
                                                  uint64_t hi, uint64_t *out) {
    if (lo > hi) return 0;
    uint64_t a = factor, b = offset;
    if (b > hi) return 0;
    if (a == 0) {
        if (lo <= b && b <= hi && tfo_is_prime64(b)) {
            *out = b;
            return 1;
        }
        return 0;
    }
    uint64_t mx = lo > b ? lo : b;
    uint64_t x_lo = (mx - b) / a;
    if ((mx - b) % a != 0) x_lo += 1;
    uint64_t x_hi = (hi - b) / a;
    uint64_t x = x_hi;
    for (;;) {
        uint64_t val = a * x + b;
        if (tfo_is_prime64(val)) {
            *out = val;
            return 1;
        }
        if (x == x_lo) break;
        x -= 1;
    }
    return 0;
}

/* ------------------------------------------------------------------ */
/* roots.rs                                                            */
/* ------------------------------------------------------------------ */

/* roots.rs:6-15 */
static void get_q_s64(uint64_t p, uint64_t *q, uint64_t *s) {
    uint64_t qq = p - 1, ss = 0;
    while (qq % 2 == 0) {
        qq /= 2;
        ss += 1;
    }
    *q = qq;
    *s = ss;
}

/* roots.rs:17-28 */
static int get_z64(uint64_t p, uint64_t *z) {
    uint64_t n = 2;
    while (n < p) {
        if (tfo_exp_mod64(p, n, (p - 1) / 2) == p - 1) {
            *z = n;
            return 1;
        }
        n += 1;
    }
    return 0;
}

/* roots.rs:31-66 (Tonelli-Shanks; which of the two square roots comes out fixes psi) */
static int sqrt_mod_ex64(uint64_t p, uint64_t q, uint64_t s, uint64_t z, uint64_t n, uint64_t *out) {
    uint64_t m = s;
    uint64_t c = tfo_exp_mod64(p, z, q);
    uint64_t t = tfo_exp_mod64(p, n, q);
    uint64_t r = tfo_exp_mod64(p, n, q / 2 + (q % 2)); /* q.div_ceil(2) */
    for (;;) {
        if (t == 0) {
            *out = 0;
            return 1;
        }
        if (t == 1) {
            *out = r;
            return 1;
        }
        uint64_t i = 0, t_pow = t;
        while (i < m) {
            t_pow = tfo_mul_mod64(p, t_pow, t_pow);
            i += 1;
            if (t_pow == 1) break;
        }
        if (i == m) return 0;
        uint64_t b = tfo_exp_mod64(p, c, (uint64_t)1 << (m - i - 1));
        m = i;
        c = tfo_mul_mod64(p, b, b);
        t = tfo_mul_mod64(p, t, c);
        r = tfo_mul_mod64(p, r, b);
    }
}

/* roots.rs:68-91 */
int tfo_find_primitive_root64(uint64_t p, uint64_t degree, uint64_t *out) {
    if (degree < 2 || (degree & (degree - 1)) != 0) return 0; /* reference asserts */
    uint32_t n = (uint32_t)__builtin_ctzll(degree);
    uint64_t root = p - 1, q, s, z;
    get_q_s64(p, &q, &s);
    if (!get_z64(p, &z)) return 0;
    for (uint32_t i = 0; i + 1 < n; i++) {
        if (!sqrt_mod_ex64(p, q, s, z, root, &root)) return 0;
    }
    *out = root;
    return 1;
}

/* roots.rs:96-107 */
int tfo_find_root_solinas_64(uint64_t n, uint64_t *out) {
    if (n == 0 || n > ((uint64_t)1 << 32)) return 0;
    const uint64_t OMG_2_32 = 16334397945464290598ull;
    uint64_t pow = ((uint64_t)1 << 32) / n;
    *out = tfo_exp_mod64(SOLINAS_P, OMG_2_32, pow);
    return 1;
}

/* lib.rs:122-125 */
size_t tfo_bit_rev(uint32_t nbits, size_t i) {
    size_t r = 0;
    for (uint32_t b = 0; b < nbits; b++) r |= ((i >> b) & 1) << (nbits - 1 - b);
    return r;
}

static uint32_t ilog2_u64(uint64_t x) { return 63u - (uint32_t)__builtin_clzll(x); }

/* ------------------------------------------------------------------ */
/* prime64.rs -- plan                                                  */
/* ------------------------------------------------------------------ */

/* generic_solinas.rs:42-75 (u64 modulus) and :77-100 (Solinas): identical add/sub */
static inline uint64_t add64(uint64_t p, uint64_t a, uint64_t b) {
    uint64_t neg_b = p - b;
    return a >= neg_b ? a - neg_b : a + b;
}
static inline uint64_t sub64(uint64_t p, uint64_t a, uint64_t b) {
    uint64_t neg_b = p - b;
    return a >= b ? a - b : a + neg_b;
}
/* generic_solinas.rs:102-128 */
static inline uint64_t solinas_mul(uint64_t a, uint64_t b) {
    const uint64_t p = SOLINAS_P;
    u128 wide = (u128)a * b;
    uint64_t lo = (uint64_t)wide;
    uint64_t hi = (uint64_t)(wide >> 64);
    uint64_t mid = hi & 0x00000000FFFFFFFFull;
    hi = (hi & 0xFFFFFFFF00000000ull) >> 32;
    uint64_t low2 = lo - hi;
    low2 += (hi > lo) ? p : 0; /* `if hi > lo { low2 += p }`, as a select */
    uint64_t product = mid << 32;
    product -= mid;
    uint64_t result = low2 + product;
    result -= ((result < product) | (result >= p)) ? p : 0;
    return result;
}
/* generic_solinas.rs:72-74 */
static inline uint64_t generic_mul64(uint64_t p, uint64_t a, uint64_t b) {
    return (uint64_t)(((u128)a * b) % p);
}
static inline uint64_t min64(uint64_t a, uint64_t b) { return a < b ? a : b; }
static inline uint32_t min32(uint32_t a, uint32_t b) { return a < b ? a : b; }

/* prime64.rs:159-241 (both initialisers; `bits` is 64 because use_ifma is never set here) */
static void init_twiddles64(uint64_t p, size_t n, uint64_t *twid, uint64_t *twid_shoup,
                            uint64_t *inv_twid, uint64_t *inv_twid_shoup) {
    uint64_t w;
    if (twid_shoup == NULL && p == SOLINAS_P) {
        /* prime64.rs:162-179: table for n in 32..32768, formula otherwise.  The table
         * is reproduced by the formula (roots.rs:150-172, checked in tests). */
        static const struct {
            size_t n;
            uint64_t w;
        } table[] = {{32, 8ull},
                     {64, 2198989700608ull},
                     {128, 14041890976876060974ull},
                     {256, 14430643036723656017ull},
                     {512, 4440654710286119610ull},
                     {1024, 8816101479115663336ull},
                     {2048, 10974926054405199669ull},
                     {4096, 1206500561358145487ull},
                     {8192, 10930245224889659871ull},
                     {16384, 3333600369887534767ull},
                     {32768, 15893793146607301539ull}};
        int found = 0;
        w = 0;
        for (size_t i = 0; i < sizeof(table) / sizeof(table[0]); i++)
            if (table[i].n == n) {
                w = table[i].w;
                found = 1;
            }
        if (!found) tfo_find_root_solinas_64(2 * (uint64_t)n, &w);
    } else {
        tfo_find_primitive_root64(p, 2 * (uint64_t)n, &w);
    }
    uint32_t nbits = (uint32_t)__builtin_ctzll((uint64_t)n);
    uint64_t wk = 1;
    for (size_t k = 0; k < n; k++) {
        size_t fwd_idx = tfo_bit_rev(nbits, k);
        twid[fwd_idx] = wk;
        if (twid_shoup) twid_shoup[fwd_idx] = (uint64_t)((((u128)wk) << 64) / p);
        size_t inv_idx = tfo_bit_rev(nbits, (n - k) % n);
        uint64_t x = (k == 0) ? wk : p - wk;
        inv_twid[inv_idx] = x;
        if (inv_twid_shoup) inv_twid_shoup[inv_idx] = (uint64_t)((((u128)x) << 64) / p);
        wk = (uint64_t)(((u128)wk * w) % p);
    }
}

/* prime64.rs:764-862 */
tfo_plan64 *tfo_plan64_try_new(size_t n, uint64_t p) {
    uint64_t dummy;
    if (n < 16 || (n & (n - 1)) != 0 || !tfo_is_prime64(p) ||
        !tfo_find_primitive_root64(p, 2 * (uint64_t)n, &dummy))
        return NULL;
    tfo_plan64 *pl = (tfo_plan64 *)calloc(1, sizeof(*pl));
    pl->n = n;
    pl->p = p;
    pl->use_ifma = 0; /* prime64.rs:776-806: the IFMA branch needs an x86 IFMA CPU path; not restated */
    /* BarrettInit64::new(modulus, 64), prime64.rs:733-758 */
    uint32_t big_q = ilog2_u64(p) + 1;
    uint32_t big_l = big_q + 64 - 1;
    u128 two_to_the_l = (u128)1 << big_l;
    pl->p_barrett = (uint64_t)(two_to_the_l / p);
    u128 beta = two_to_the_l % p;
    u128 single_reduction_threshold = (u128)p - ((u128)1 << (big_q - 1));
    int single_step = beta <= single_reduction_threshold;
    pl->big_q = big_q;
    /* prime64.rs:815-817 */
    pl->can_use_fast_reduction_code =
        (p < 6148914691236517206ull) || (single_step && p < ((uint64_t)1 << 63));
    pl->twid = (uint64_t *)calloc(n, sizeof(uint64_t));
    pl->inv_twid = (uint64_t *)calloc(n, sizeof(uint64_t));
    if (p < ((uint64_t)1 << 63)) {
        pl->twid_shoup = (uint64_t *)calloc(n, sizeof(uint64_t));
        pl->inv_twid_shoup = (uint64_t *)calloc(n, sizeof(uint64_t));
    }
    init_twiddles64(p, n, pl->twid, pl->twid_shoup, pl->inv_twid, pl->inv_twid_shoup);
    /* prime64.rs:844-845 */
    pl->n_inv_mod_p = tfo_exp_mod64(p, (uint64_t)n, p - 2);
    pl->n_inv_mod_p_shoup = (uint64_t)((((u128)pl->n_inv_mod_p) << 64) / p);
    return pl;
}

void tfo_plan64_free(tfo_plan64 *pl) {
    if (!pl) return;
    free(pl->twid);
    free(pl->twid_shoup);
    free(pl->inv_twid);
    free(pl->inv_twid_shoup);
    free(pl);
}

/* generic_solinas.rs:449-481 at recursion depth 0.  The depth-first recursion
 * (generic_solinas.rs:1338-1386) performs the same butterflies with the same
 * twiddles in another order, so one breadth-first sweep gives the same values.
 * The Solinas and the generic-modulus instantiations are separate loops, as the
 * reference monomorphises them (PrimeModulus for Solinas / for u64). */
static void fwd64_solinas(uint64_t *data, size_t n, const uint64_t *twid) {
    const uint64_t p = SOLINAS_P;
    size_t t = n / 2, m = 1;
    while (m < n) {
        for (size_t i = 0; i < m; i++) {
            uint64_t w1 = twid[m + i];
            uint64_t *z0 = data + 2 * i * t, *z1 = z0 + t;
            for (size_t j = 0; j < t; j++) {
                uint64_t z1w = solinas_mul(z1[j], w1);
                uint64_t a = z0[j];
                z0[j] = add64(p, a, z1w);
                z1[j] = sub64(p, a, z1w);
            }
        }
        t /= 2;
        m *= 2;
    }
}
static void fwd64_generic(uint64_t *data, size_t n, uint64_t p, const uint64_t *twid) {
    size_t t = n / 2, m = 1;
    while (m < n) {
        for (size_t i = 0; i < m; i++) {
            uint64_t w1 = twid[m + i];
            uint64_t *z0 = data + 2 * i * t, *z1 = z0 + t;
            for (size_t j = 0; j < t; j++) {
                uint64_t z1w = generic_mul64(p, z1[j], w1);
                uint64_t a = z0[j];
                z0[j] = add64(p, a, z1w);
                z1[j] = sub64(p, a, z1w);
            }
        }
        t /= 2;
        m *= 2;
    }
}
static void fwd64_exact(uint64_t *data, size_t n, uint64_t p, const uint64_t *twid, int solinas) {
    if (solinas)
        fwd64_solinas(data, n, twid);
    else
        fwd64_generic(data, n, p, twid);
}
/* generic_solinas.rs:483-514 */
static void inv64_solinas(uint64_t *data, size_t n, const uint64_t *inv_twid) {
    const uint64_t p = SOLINAS_P;
    size_t t = 1, m = n;
    while (m > 1) {
        m /= 2;
        for (size_t i = 0; i < m; i++) {
            uint64_t w1 = inv_twid[m + i];
            uint64_t *z0 = data + 2 * i * t, *z1 = z0 + t;
            for (size_t j = 0; j < t; j++) {
                uint64_t a = z0[j], b = z1[j];
                z0[j] = add64(p, a, b);
                z1[j] = solinas_mul(sub64(p, a, b), w1);
            }
        }
        t *= 2;
    }
}
static void inv64_generic(uint64_t *data, size_t n, uint64_t p, const uint64_t *inv_twid) {
    size_t t = 1, m = n;
    while (m > 1) {
        m /= 2;
        for (size_t i = 0; i < m; i++) {
            uint64_t w1 = inv_twid[m + i];
            uint64_t *z0 = data + 2 * i * t, *z1 = z0 + t;
            for (size_t j = 0; j < t; j++) {
                uint64_t a = z0[j], b = z1[j];
                z0[j] = add64(p, a, b);
                z1[j] = generic_mul64(p, sub64(p, a, b), w1);
            }
        }
        t *= 2;
    }
}
static void inv64_exact(uint64_t *data, size_t n, uint64_t p, const uint64_t *inv_twid,
                        int solinas) {
    if (solinas)
        inv64_solinas(data, n, inv_twid);
    else
        inv64_generic(data, n, p, inv_twid);
}

/* shoup.rs:544-615 driver with the less_than_62bit.rs:117-154 (lazy in [0,4p)) or
 * less_than_63bit.rs:117-153 (lazy in [0,2p)) butterflies. */
static void fwd64_shoup(uint64_t *data, size_t n, uint64_t p, const uint64_t *twid,
                        const uint64_t *twid_shoup, int bits62) {
    size_t t = n, m = 1;
    uint64_t neg_p = (uint64_t)0 - p, two_p = 2 * p;
    while (m < n) {
        t /= 2;
        for (size_t i = 0; i < m; i++) {
            uint64_t w = twid[m + i], ws = twid_shoup[m + i];
            uint64_t *d0 = data + 2 * i * t, *d1 = d0 + t;
            for (size_t j = 0; j < t; j++) {
                uint64_t z0 = d0[j], z1 = d1[j];
                uint64_t q = (uint64_t)(((u128)z1 * ws) >> 64);
                uint64_t tt = z1 * w + q * neg_p;
                if (bits62) {
                    z0 = min64(z0, z0 - two_p);
                    if (t == 1) {
                        z0 = min64(z0, z0 - p);
                        tt = min64(tt, tt - p);
                        uint64_t r0 = z0 + tt, r1 = z0 - tt + p;
                        d0[j] = min64(r0, r0 - p);
                        d1[j] = min64(r1, r1 - p);
                    } else {
                        d0[j] = z0 + tt;
                        d1[j] = z0 - tt + two_p;
                    }
                } else {
                    z0 = min64(z0, z0 - p);
                    tt = min64(tt, tt - p);
                    uint64_t r0 = z0 + tt, r1 = z0 - tt + p;
                    if (t == 1) {
                        d0[j] = min64(r0, r0 - p);
                        d1[j] = min64(r1, r1 - p);
                    } else {
                        d0[j] = r0;
                        d1[j] = r1;
                    }
                }
            }
        }
        m *= 2;
    }
}
/* shoup.rs:1306-1377 with less_than_62bit.rs:271-310 / less_than_63bit.rs:214-234 */
static void inv64_shoup(uint64_t *data, size_t n, uint64_t p, const uint64_t *inv_twid,
                        const uint64_t *inv_twid_shoup, int bits62) {
    size_t t = 1, m = n;
    uint64_t neg_p = (uint64_t)0 - p, two_p = 2 * p;
    while (m > 1) {
        m /= 2;
        for (size_t i = 0; i < m; i++) {
            uint64_t w = inv_twid[m + i], ws = inv_twid_shoup[m + i];
            uint64_t *d0 = data + 2 * i * t, *d1 = d0 + t;
            for (size_t j = 0; j < t; j++) {
                uint64_t z0 = d0[j], z1 = d1[j];
                uint64_t y0 = z0 + z1;
                if (bits62) {
                    y0 = min64(y0, y0 - two_p);
                    uint64_t tt = z0 - z1 + two_p;
                    uint64_t q = (uint64_t)(((u128)tt * ws) >> 64);
                    uint64_t y1 = tt * w + q * neg_p;
                    if (m == 1) {
                        y0 = min64(y0, y0 - p);
                        y1 = min64(y1, y1 - p);
                    }
                    d0[j] = y0;
                    d1[j] = y1;
                } else {
                    y0 = min64(y0, y0 - p);
                    uint64_t tt = z0 - z1 + p;
                    uint64_t q = (uint64_t)(((u128)tt * ws) >> 64);
                    uint64_t y1 = tt * w + q * neg_p;
                    d0[j] = y0;
                    d1[j] = min64(y1, y1 - p);
                }
            }
        }
        t *= 2;
    }
}

void tfo_plan64_fwd_generic(const tfo_plan64 *pl, uint64_t *buf) {
    fwd64_exact(buf, pl->n, pl->p, pl->twid, 0);
}
void tfo_plan64_inv_generic(const tfo_plan64 *pl, uint64_t *buf) {
    inv64_exact(buf, pl->n, pl->p, pl->inv_twid, 0);
}

/* prime64.rs:897-968 (scalar arms of the dispatch) */
void tfo_plan64_fwd(const tfo_plan64 *pl, uint64_t *buf) {
    uint64_t p = pl->p;
    if (p < ((uint64_t)1 << 62))
        fwd64_shoup(buf, pl->n, p, pl->twid, pl->twid_shoup, 1);
    else if (p < ((uint64_t)1 << 63))
        fwd64_shoup(buf, pl->n, p, pl->twid, pl->twid_shoup, 0);
    else
        fwd64_exact(buf, pl->n, p, pl->twid, p == SOLINAS_P);
}
/* prime64.rs:975-1046 */
void tfo_plan64_inv(const tfo_plan64 *pl, uint64_t *buf) {
    uint64_t p = pl->p;
    if (p < ((uint64_t)1 << 62))
        inv64_shoup(buf, pl->n, p, pl->inv_twid, pl->inv_twid_shoup, 1);
    else if (p < ((uint64_t)1 << 63))
        inv64_shoup(buf, pl->n, p, pl->inv_twid, pl->inv_twid_shoup, 0);
    else
        inv64_exact(buf, pl->n, p, pl->inv_twid, p == SOLINAS_P);
}

/* prime64.rs:1137-1179 ; scalar :715-724 */
void tfo_plan64_normalize(const tfo_plan64 *pl, uint64_t *values, size_t len) {
    uint64_t p = pl->p, ninv = pl->n_inv_mod_p, ninv_s = pl->n_inv_mod_p_shoup;
    if (pl->can_use_fast_reduction_code) {
        for (size_t i = 0; i < len; i++) {
            uint64_t val = values[i];
            uint64_t q = (uint64_t)(((u128)val * ninv_s) >> 64);
            uint64_t t = val * ninv - q * p;
            values[i] = min64(t, t - p);
        }
    } else if (p == SOLINAS_P) {
        for (size_t i = 0; i < len; i++) values[i] = solinas_mul(values[i], ninv);
    } else {
        for (size_t i = 0; i < len; i++) values[i] = generic_mul64(p, values[i], ninv);
    }
}

/* prime64.rs:1050-1133 ; scalar :559-584 */
void tfo_plan64_mul_assign_normalize(const tfo_plan64 *pl, uint64_t *lhs, const uint64_t *rhs,
                                     size_t len) {
    uint64_t p = pl->p, ninv = pl->n_inv_mod_p, ninv_s = pl->n_inv_mod_p_shoup;
    if (pl->can_use_fast_reduction_code) {
        uint64_t big_q_m1 = pl->big_q - 1;
        for (size_t i = 0; i < len; i++) {
            u128 d = (u128)lhs[i] * rhs[i];
            uint64_t c1 = (uint64_t)(d >> big_q_m1);
            uint64_t c3 = (uint64_t)(((u128)c1 * pl->p_barrett) >> 64);
            uint64_t prod = (uint64_t)d - p * c3;
            uint64_t q = (uint64_t)(((u128)prod * ninv_s) >> 64);
            uint64_t t = prod * ninv - q * p;
            lhs[i] = min64(t, t - p);
        }
    } else if (p == SOLINAS_P) {
        for (size_t i = 0; i < len; i++) lhs[i] = solinas_mul(solinas_mul(lhs[i], rhs[i]), ninv);
    } else {
        for (size_t i = 0; i < len; i++)
            lhs[i] = generic_mul64(p, generic_mul64(p, lhs[i], rhs[i]), ninv);
    }
}

/* prime64.rs:1182-1222 ; scalar :586-609 */
void tfo_plan64_mul_accumulate(const tfo_plan64 *pl, uint64_t *acc, const uint64_t *lhs,
                               const uint64_t *rhs, size_t len) {
    uint64_t p = pl->p;
    if (pl->can_use_fast_reduction_code) {
        uint64_t big_q_m1 = pl->big_q - 1;
        for (size_t i = 0; i < len; i++) {
            u128 d = (u128)lhs[i] * rhs[i];
            uint64_t c1 = (uint64_t)(d >> big_q_m1);
            uint64_t c3 = (uint64_t)(((u128)c1 * pl->p_barrett) >> 64);
            uint64_t prod = (uint64_t)d - p * c3;
            prod = min64(prod, prod - p);
            uint64_t a = prod + acc[i];
            acc[i] = min64(a, a - p);
        }
    } else if (p == SOLINAS_P) {
        for (size_t i = 0; i < len; i++) acc[i] = add64(p, acc[i], solinas_mul(lhs[i], rhs[i]));
    } else {
        for (size_t i = 0; i < len; i++)
            acc[i] = add64(p, acc[i], generic_mul64(p, lhs[i], rhs[i]));
    }
}

/* ------------------------------------------------------------------ */
/* prime32.rs -- plan                                                  */
/* ------------------------------------------------------------------ */

/* prime32/generic.rs:9-31 */
static inline uint32_t add32(uint32_t p, uint32_t a, uint32_t b) {
    uint32_t neg_b = p - b;
    return a >= neg_b ? a - neg_b : a + b;
}
static inline uint32_t sub32(uint32_t p, uint32_t a, uint32_t b) {
    uint32_t neg_b = p - b;
    return a >= b ? a - b : a + neg_b;
}
static inline uint32_t generic_mul32(uint32_t p, uint32_t a, uint32_t b) {
    return (uint32_t)(((uint64_t)a * b) % p);
}

/* prime32.rs:223-282 */
static void init_twiddles32(uint32_t p, size_t n, uint32_t *twid, uint32_t *twid_shoup,
                            uint32_t *inv_twid, uint32_t *inv_twid_shoup) {
    uint64_t w64 = 0;
    tfo_find_primitive_root64((uint64_t)p, 2 * (uint64_t)n, &w64);
    uint32_t w = (uint32_t)w64;
    uint32_t nbits = (uint32_t)__builtin_ctzll((uint64_t)n);
    uint32_t wk = 1;
    for (size_t k = 0; k < n; k++) {
        size_t fwd_idx = tfo_bit_rev(nbits, k);
        twid[fwd_idx] = wk;
        if (twid_shoup) twid_shoup[fwd_idx] = (uint32_t)((((uint64_t)wk) << 32) / p);
        size_t inv_idx = tfo_bit_rev(nbits, (n - k) % n);
        uint32_t x = (k == 0) ? wk : p - wk;
        inv_twid[inv_idx] = x;
        if (inv_twid_shoup) inv_twid_shoup[inv_idx] = (uint32_t)((((uint64_t)x) << 32) / p);
        wk = (uint32_t)(((uint64_t)wk * w) % p);
    }
}

/* prime32.rs:662-765 */
tfo_plan32 *tfo_plan32_try_new(size_t n, uint32_t p) {
    uint64_t dummy;
    if (n < 32 || (n & (n - 1)) != 0 || !tfo_is_prime64((uint64_t)p) ||
        !tfo_find_primitive_root64((uint64_t)p, 2 * (uint64_t)n, &dummy))
        return NULL;
    tfo_plan32 *pl = (tfo_plan32 *)calloc(1, sizeof(*pl));
    pl->n = n;
    pl->p = p;
    pl->twid = (uint32_t *)calloc(n, sizeof(uint32_t));
    pl->inv_twid = (uint32_t *)calloc(n, sizeof(uint32_t));
    if (p < ((uint32_t)1 << 31)) {
        pl->twid_shoup = (uint32_t *)calloc(n, sizeof(uint32_t));
        pl->inv_twid_shoup = (uint32_t *)calloc(n, sizeof(uint32_t));
    }
    init_twiddles32(p, n, pl->twid, pl->twid_shoup, pl->inv_twid, pl->inv_twid_shoup);
    pl->n_inv_mod_p = tfo_exp_mod32(p, (uint32_t)n, p - 2);
    pl->n_inv_mod_p_shoup = (uint32_t)((((uint64_t)pl->n_inv_mod_p) << 32) / p);
    /* BarrettInit32::new, prime32.rs:606-627 */
    uint32_t big_q = ilog2_u64((uint64_t)p) + 1;
    uint32_t big_l = big_q + 31;
    uint64_t two_to_the_l = (uint64_t)1 << big_l;
    pl->p_barrett = (uint32_t)(two_to_the_l / p);
    uint64_t beta = two_to_the_l % p;
    uint64_t thr = (uint64_t)p - ((uint64_t)1 << (big_q - 1));
    int single_step = beta <= thr;
    pl->big_q = big_q;
    /* prime32.rs:748-749 */
    pl->can_use_fast_reduction_code =
        (p < 1431655766u) || (single_step && p <= ((uint32_t)1 << 31));
    return pl;
}

void tfo_plan32_free(tfo_plan32 *pl) {
    if (!pl) return;
    free(pl->twid);
    free(pl->twid_shoup);
    free(pl->inv_twid);
    free(pl->inv_twid_shoup);
    free(pl);
}

/* prime32/generic.rs:228-260 / :312-343 at depth 0 */
static void fwd32_exact(uint32_t *data, size_t n, uint32_t p, const uint32_t *twid) {
    size_t t = n / 2, m = 1;
    while (m < n) {
        for (size_t i = 0; i < m; i++) {
            uint32_t w1 = twid[m + i];
            uint32_t *z0 = data + 2 * i * t, *z1 = z0 + t;
            for (size_t j = 0; j < t; j++) {
                uint32_t z1w = generic_mul32(p, z1[j], w1);
                uint32_t a = z0[j];
                z0[j] = add32(p, a, z1w);
                z1[j] = sub32(p, a, z1w);
            }
        }
        t /= 2;
        m *= 2;
    }
}
static void inv32_exact(uint32_t *data, size_t n, uint32_t p, const uint32_t *inv_twid) {
    size_t t = 1, m = n;
    while (m > 1) {
        m /= 2;
        for (size_t i = 0; i < m; i++) {
            uint32_t w1 = inv_twid[m + i];
            uint32_t *z0 = data + 2 * i * t, *z1 = z0 + t;
            for (size_t j = 0; j < t; j++) {
                uint32_t a = z0[j], b = z1[j];
                z0[j] = add32(p, a, b);
                z1[j] = generic_mul32(p, sub32(p, a, b), w1);
            }
        }
        t *= 2;
    }
}

/* prime32/shoup.rs:582-635 with less_than_30bit.rs:115-153 / less_than_31bit.rs:117-157 */
static void fwd32_shoup(uint32_t *data, size_t n, uint32_t p, const uint32_t *twid,
                        const uint32_t *twid_shoup, int bits30) {
    size_t t = n, m = 1;
    uint32_t neg_p = (uint32_t)0 - p, two_p = 2 * p;
    while (m < n) {
        t /= 2;
        for (size_t i = 0; i < m; i++) {
            uint32_t w = twid[m + i], ws = twid_shoup[m + i];
            uint32_t *d0 = data + 2 * i * t, *d1 = d0 + t;
            for (size_t j = 0; j < t; j++) {
                uint32_t z0 = d0[j], z1 = d1[j];
                uint32_t q = (uint32_t)(((uint64_t)z1 * ws) >> 32);
                uint32_t tt = z1 * w + q * neg_p;
                if (bits30) {
                    z0 = min32(z0, z0 - two_p);
                    if (t == 1) {
                        z0 = min32(z0, z0 - p);
                        tt = min32(tt, tt - p);
                        uint32_t r0 = z0 + tt, r1 = z0 - tt + p;
                        d0[j] = min32(r0, r0 - p);
                        d1[j] = min32(r1, r1 - p);
                    } else {
                        d0[j] = z0 + tt;
                        d1[j] = z0 - tt + two_p;
                    }
                } else {
                    z0 = min32(z0, z0 - p);
                    tt = min32(tt, tt - p);
                    uint32_t r0 = z0 + tt, r1 = z0 - tt + p;
                    if (t == 1) {
                        d0[j] = min32(r0, r0 - p);
                        d1[j] = min32(r1, r1 - p);
                    } else {
                        d0[j] = r0;
                        d1[j] = r1;
                    }
                }
            }
        }
        m *= 2;
    }
}
/* prime32/shoup.rs:1355-1408 with less_than_30bit.rs:265-303 / less_than_31bit.rs:214-234 */
static void inv32_shoup(uint32_t *data, size_t n, uint32_t p, const uint32_t *inv_twid,
                        const uint32_t *inv_twid_shoup, int bits30) {
    size_t t = 1, m = n;
    uint32_t neg_p = (uint32_t)0 - p, two_p = 2 * p;
    while (m > 1) {
        m /= 2;
        for (size_t i = 0; i < m; i++) {
            uint32_t w = inv_twid[m + i], ws = inv_twid_shoup[m + i];
            uint32_t *d0 = data + 2 * i * t, *d1 = d0 + t;
            for (size_t j = 0; j < t; j++) {
                uint32_t z0 = d0[j], z1 = d1[j];
                uint32_t y0 = z0 + z1;
                if (bits30) {
                    y0 = min32(y0, y0 - two_p);
                    uint32_t tt = z0 - z1 + two_p;
                    uint32_t q = (uint32_t)(((uint64_t)tt * ws) >> 32);
                    uint32_t y1 = tt * w + q * neg_p;
                    if (m == 1) {
                        y0 = min32(y0, y0 - p);
                        y1 = min32(y1, y1 - p);
                    }
                    d0[j] = y0;
                    d1[j] = y1;
                } else {
                    y0 = min32(y0, y0 - p);
                    uint32_t tt = z0 - z1 + p;
                    uint32_t q = (uint32_t)(((uint64_t)tt * ws) >> 32);
                    uint32_t y1 = tt * w + q * neg_p;
                    d0[j] = y0;
                    d1[j] = min32(y1, y1 - p);
                }
            }
        }
        t *= 2;
    }
}

void tfo_plan32_fwd_generic(const tfo_plan32 *pl, uint32_t *buf) {
    fwd32_exact(buf, pl->n, pl->p, pl->twid);
}
void tfo_plan32_inv_generic(const tfo_plan32 *pl, uint32_t *buf) {
    inv32_exact(buf, pl->n, pl->p, pl->inv_twid);
}

/* prime32.rs:797-843 */
void tfo_plan32_fwd(const tfo_plan32 *pl, uint32_t *buf) {
    uint32_t p = pl->p;
    if (p < ((uint32_t)1 << 30))
        fwd32_shoup(buf, pl->n, p, pl->twid, pl->twid_shoup, 1);
    else if (p < ((uint32_t)1 << 31))
        fwd32_shoup(buf, pl->n, p, pl->twid, pl->twid_shoup, 0);
    else
        fwd32_exact(buf, pl->n, p, pl->twid);
}
/* prime32.rs:850-896 */
void tfo_plan32_inv(const tfo_plan32 *pl, uint32_t *buf) {
    uint32_t p = pl->p;
    if (p < ((uint32_t)1 << 30))
        inv32_shoup(buf, pl->n, p, pl->inv_twid, pl->inv_twid_shoup, 1);
    else if (p < ((uint32_t)1 << 31))
        inv32_shoup(buf, pl->n, p, pl->inv_twid, pl->inv_twid_shoup, 0);
    else
        inv32_exact(buf, pl->n, p, pl->inv_twid);
}

/* prime32.rs:956-990 ; scalar :477-486 */
void tfo_plan32_normalize(const tfo_plan32 *pl, uint32_t *values, size_t len) {
    uint32_t p = pl->p, ninv = pl->n_inv_mod_p, ninv_s = pl->n_inv_mod_p_shoup;
    if (pl->can_use_fast_reduction_code) {
        for (size_t i = 0; i < len; i++) {
            uint32_t val = values[i];
            uint32_t q = (uint32_t)(((uint64_t)val * ninv_s) >> 32);
            uint32_t t = val * ninv - q * p;
            values[i] = min32(t, t - p);
        }
    } else {
        for (size_t i = 0; i < len; i++) values[i] = generic_mul32(p, values[i], ninv);
    }
}

/* prime32.rs:900-952 ; scalar :383-408 */
void tfo_plan32_mul_assign_normalize(const tfo_plan32 *pl, uint32_t *lhs, const uint32_t *rhs,
                                     size_t len) {
    uint32_t p = pl->p, ninv = pl->n_inv_mod_p, ninv_s = pl->n_inv_mod_p_shoup;
    if (pl->can_use_fast_reduction_code) {
        uint32_t big_q_m1 = pl->big_q - 1;
        for (size_t i = 0; i < len; i++) {
            uint64_t d = (uint64_t)lhs[i] * rhs[i];
            uint32_t c1 = (uint32_t)(d >> big_q_m1);
            uint32_t c3 = (uint32_t)(((uint64_t)c1 * pl->p_barrett) >> 32);
            uint32_t prod = (uint32_t)d - p * c3;
            uint32_t q = (uint32_t)(((uint64_t)prod * ninv_s) >> 32);
            uint32_t t = prod * ninv - q * p;
            lhs[i] = min32(t, t - p);
        }
    } else {
        for (size_t i = 0; i < len; i++)
            lhs[i] = generic_mul32(p, generic_mul32(p, lhs[i], rhs[i]), ninv);
    }
}

/* prime32.rs:993-1015 ; scalar :575-598 */
void tfo_plan32_mul_accumulate(const tfo_plan32 *pl, uint32_t *acc, const uint32_t *lhs,
                               const uint32_t *rhs, size_t len) {
    uint32_t p = pl->p;
    if (pl->can_use_fast_reduction_code) {
        uint32_t big_q_m1 = pl->big_q - 1;
        for (size_t i = 0; i < len; i++) {
            uint64_t d = (uint64_t)lhs[i] * rhs[i];
            uint32_t c1 = (uint32_t)(d >> big_q_m1);
            uint32_t c3 = (uint32_t)(((uint64_t)c1 * pl->p_barrett) >> 32);
            uint32_t prod = (uint32_t)d - p * c3;
            prod = min32(prod, prod - p);
            uint32_t a = prod + acc[i];
            acc[i] = min32(a, a - p);
        }
    } else {
        for (size_t i = 0; i < len; i++)
            acc[i] = add32(p, acc[i], generic_mul32(p, lhs[i], rhs[i]));
    }
}

/* ------------------------------------------------------------------ */
/* schoolbook convolutions (the reference tests' own oracle)           */
/* ------------------------------------------------------------------ */

/* prime64.rs:1264-1276 */
void tfo_negacyclic_convolution_mod64(size_t n, uint64_t p, const uint64_t *lhs,
                                      const uint64_t *rhs, uint64_t *out) {
    u128 *full = (u128 *)calloc(2 * n, sizeof(u128));
    for (size_t i = 0; i < n; i++)
        for (size_t j = 0; j < n; j++)
            full[i + j] = (full[i + j] + (((u128)lhs[i] * rhs[j]) % p)) % p;
    for (size_t i = 0; i < n; i++) out[i] = sub64(p, (uint64_t)full[i], (uint64_t)full[i + n]);
    free(full);
}
void tfo_negacyclic_convolution_mod32(size_t n, uint32_t p, const uint32_t *lhs,
                                      const uint32_t *rhs, uint32_t *out) {
    uint64_t *full = (uint64_t *)calloc(2 * n, sizeof(uint64_t));
    for (size_t i = 0; i < n; i++)
        for (size_t j = 0; j < n; j++)
            full[i + j] = (full[i + j] + (((uint64_t)lhs[i] * rhs[j]) % p)) % p;
    for (size_t i = 0; i < n; i++) out[i] = sub32(p, (uint32_t)full[i], (uint32_t)full[i + n]);
    free(full);
}
void tfo_negacyclic_convolution_wrapping_u32(size_t n, const uint32_t *lhs, const uint32_t *rhs,
                                             uint32_t *out) {
    uint32_t *full = (uint32_t *)calloc(2 * n, sizeof(uint32_t));
    for (size_t i = 0; i < n; i++)
        for (size_t j = 0; j < n; j++) full[i + j] += lhs[i] * rhs[j];
    for (size_t i = 0; i < n; i++) out[i] = full[i] - full[i + n];
    free(full);
}
void tfo_negacyclic_convolution_wrapping_u64(size_t n, const uint64_t *lhs, const uint64_t *rhs,
                                             uint64_t *out) {
    uint64_t *full = (uint64_t *)calloc(2 * n, sizeof(uint64_t));
    for (size_t i = 0; i < n; i++)
        for (size_t j = 0; j < n; j++) full[i + j] += lhs[i] * rhs[j];
    for (size_t i = 0; i < n; i++) out[i] = full[i] - full[i + n];
    free(full);
}
void tfo_negacyclic_convolution_wrapping_u128(size_t n, const u128 *lhs, const u128 *rhs,
                                              u128 *out) {
    u128 *full = (u128 *)calloc(2 * n, sizeof(u128));
    for (size_t i = 0; i < n; i++)
        for (size_t j = 0; j < n; j++) full[i + j] += lhs[i] * rhs[j];
    for (size_t i = 0; i < n; i++) out[i] = full[i] - full[i + n];
    free(full);
}

/* ------------------------------------------------------------------ */
/* CRT constants, lib.rs:451-656 -- recomputed from the literal primes */
/* ------------------------------------------------------------------ */

static const uint32_t P32[10] = {
    /* lib.rs:457-466 (binary literals) */
    0x3F5A0001u, 0x3F5D0001u, 0x3F760001u, 0x3F820001u, 0x3FAC0001u,
    0x3FAF0001u, 0x3FB10001u, 0x3FBB0001u, 0x3FDE0001u, 0x3FFC0001u};
/* lib.rs:605-610 (binary literals) */
static const uint64_t primes52_tab[6] = {0x3FFFFFE770001ull, 0x3FFFFFEB90001ull, 0x3FFFFFEC80001ull,
                                         0x3FFFFFF8B0001ull, 0x3FFFFFFB80001ull, 0x3FFFFFFC70001ull};

uint32_t tfo_primes32(int i) { return P32[i]; }
uint64_t tfo_primes52(int i) { return primes52_tab[i]; }

/* lib.rs:490-497 */
static uint32_t c_inv_mod32(uint32_t modulus, uint32_t x) { return tfo_exp_mod32(modulus, x, modulus - 2); }
/* lib.rs:499-515 */
static uint64_t c_shoup64(uint64_t modulus, uint64_t w) { return (uint64_t)((((u128)w) << 64) / modulus); }
static uint64_t c_inv_mod52(uint64_t modulus, uint64_t x) { return tfo_exp_mod64(modulus, x, modulus - 2); }

/* native32.rs:21-24 */
static inline uint32_t mul_mod32(uint32_t p, uint32_t a, uint32_t b) {
    return (uint32_t)(((uint64_t)a * b) % p);
}
/* native64.rs:36-40 */
static inline uint64_t mul_mod64s(uint64_t p_neg, uint64_t a, uint64_t b, uint64_t b_shoup) {
    uint64_t q = (uint64_t)(((u128)a * b_shoup) >> 64);
    uint64_t r = a * b + p_neg * q;
    return min64(r, r + p_neg);
}
/* native32.rs:95-106 (mul_mod52_avx512) for operands < 2^52: the IFMA Shoup product with one
 * conditional subtract equals the exact a*b mod p (a < 2^52 keeps the Shoup remainder in [0,2p)). */
static inline uint64_t mul_mod52(uint64_t p, uint64_t a, uint64_t b) {
    return (uint64_t)(((u128)a * b) % p);
}

typedef struct {
    uint32_t P0_INV_MOD_P1, P01_INV_MOD_P2, P012_INV_MOD_P3, P0123_INV_MOD_P4;
    uint32_t P1_INV_MOD_P2, P3_INV_MOD_P4;
    uint64_t P12, P34, P0_INV_MOD_P12, P0_INV_MOD_P12_SHOUP, P0_MOD_P34_SHOUP, P012_INV_MOD_P34,
        P012_INV_MOD_P34_SHOUP;
    uint32_t P2_INV_MOD_P3, P4_INV_MOD_P5, P6_INV_MOD_P7, P8_INV_MOD_P9;
    uint64_t P01, P23, P45, P67, P89;
    uint64_t P01_MOD_P45_SHOUP, P01_MOD_P67_SHOUP, P01_MOD_P89_SHOUP, P23_MOD_P67_SHOUP,
        P23_MOD_P89_SHOUP, P45_MOD_P89_SHOUP;
    uint64_t P01_INV_MOD_P23, P01_INV_MOD_P23_SHOUP, P0123_INV_MOD_P45, P0123_INV_MOD_P45_SHOUP,
        P012345_INV_MOD_P67, P012345_INV_MOD_P67_SHOUP, P01234567_INV_MOD_P89,
        P01234567_INV_MOD_P89_SHOUP;
    u128 P0123, P012345, P01234567, P0123456789;
    /* primes52 */
    uint64_t Q0_INV_MOD_Q1, Q01_INV_MOD_Q2;
} crt_consts;

static crt_consts K;
static pthread_once_t K_once = PTHREAD_ONCE_INIT;

static void crt_init(void) {
    const uint32_t *P = P32;
    /* lib.rs:517-525 */
    K.P0_INV_MOD_P1 = c_inv_mod32(P[1], P[0]);
    K.P01_INV_MOD_P2 = c_inv_mod32(P[2], mul_mod32(P[2], P[0], P[1]));
    K.P012_INV_MOD_P3 = c_inv_mod32(P[3], mul_mod32(P[3], mul_mod32(P[3], P[0], P[1]), P[2]));
    K.P0123_INV_MOD_P4 = c_inv_mod32(
        P[4], mul_mod32(P[4], mul_mod32(P[4], mul_mod32(P[4], P[0], P[1]), P[2]), P[3]));
    /* lib.rs:534-548 */
    K.P1_INV_MOD_P2 = c_inv_mod32(P[2], P[1]);
    K.P3_INV_MOD_P4 = c_inv_mod32(P[4], P[3]);
    K.P12 = (uint64_t)P[1] * P[2];
    K.P34 = (uint64_t)P[3] * P[4];
    K.P0_INV_MOD_P12 =
        tfo_exp_mod64(K.P12, P[0], ((uint64_t)P[1] - 1) * ((uint64_t)P[2] - 1) - 1);
    K.P0_INV_MOD_P12_SHOUP = c_shoup64(K.P12, K.P0_INV_MOD_P12);
    K.P0_MOD_P34_SHOUP = c_shoup64(K.P34, P[0]);
    K.P012_INV_MOD_P34 = tfo_exp_mod64(K.P34, tfo_mul_mod64(K.P34, P[0], K.P12),
                                       ((uint64_t)P[3] - 1) * ((uint64_t)P[4] - 1) - 1);
    K.P012_INV_MOD_P34_SHOUP = c_shoup64(K.P34, K.P012_INV_MOD_P34);
    /* lib.rs:550-557 */
    K.P2_INV_MOD_P3 = c_inv_mod32(P[3], P[2]);
    K.P4_INV_MOD_P5 = c_inv_mod32(P[5], P[4]);
    K.P6_INV_MOD_P7 = c_inv_mod32(P[7], P[6]);
    K.P8_INV_MOD_P9 = c_inv_mod32(P[9], P[8]);
    /* lib.rs:559-563 */
    K.P01 = (uint64_t)P[0] * P[1];
    K.P23 = (uint64_t)P[2] * P[3];
    K.P45 = (uint64_t)P[4] * P[5];
    K.P67 = (uint64_t)P[6] * P[7];
    K.P89 = (uint64_t)P[8] * P[9];
    /* lib.rs:565-572 */
    K.P01_MOD_P45_SHOUP = c_shoup64(K.P45, K.P01);
    K.P01_MOD_P67_SHOUP = c_shoup64(K.P67, K.P01);
    K.P01_MOD_P89_SHOUP = c_shoup64(K.P89, K.P01);
    K.P23_MOD_P67_SHOUP = c_shoup64(K.P67, K.P23);
    K.P23_MOD_P89_SHOUP = c_shoup64(K.P89, K.P23);
    K.P45_MOD_P89_SHOUP = c_shoup64(K.P89, K.P45);
    /* lib.rs:574-593 */
    K.P01_INV_MOD_P23 =
        tfo_exp_mod64(K.P23, K.P01, ((uint64_t)P[2] - 1) * ((uint64_t)P[3] - 1) - 1);
    K.P01_INV_MOD_P23_SHOUP = c_shoup64(K.P23, K.P01_INV_MOD_P23);
    K.P0123_INV_MOD_P45 = tfo_exp_mod64(K.P45, tfo_mul_mod64(K.P45, K.P01, K.P23),
                                        ((uint64_t)P[4] - 1) * ((uint64_t)P[5] - 1) - 1);
    K.P0123_INV_MOD_P45_SHOUP = c_shoup64(K.P45, K.P0123_INV_MOD_P45);
    K.P012345_INV_MOD_P67 = tfo_exp_mod64(
        K.P67, tfo_mul_mod64(K.P67, tfo_mul_mod64(K.P67, K.P01, K.P23), K.P45),
        ((uint64_t)P[6] - 1) * ((uint64_t)P[7] - 1) - 1);
    K.P012345_INV_MOD_P67_SHOUP = c_shoup64(K.P67, K.P012345_INV_MOD_P67);
    K.P01234567_INV_MOD_P89 = tfo_exp_mod64(
        K.P89,
        tfo_mul_mod64(K.P89, tfo_mul_mod64(K.P89, tfo_mul_mod64(K.P89, K.P01, K.P23), K.P45),
                      K.P67),
        ((uint64_t)P[8] - 1) * ((uint64_t)P[9] - 1) - 1);
    K.P01234567_INV_MOD_P89_SHOUP = c_shoup64(K.P89, K.P01234567_INV_MOD_P89);
    /* lib.rs:595-598 */
    K.P0123 = (u128)K.P01 * (u128)K.P23;
    K.P012345 = K.P0123 * (u128)K.P45;
    K.P01234567 = K.P012345 * (u128)K.P67;
    K.P0123456789 = K.P01234567 * (u128)K.P89;
    /* lib.rs:639-643 (primes52; the 52-bit Shoup companions are an IFMA detail and unobservable) */
    const uint64_t *Q = primes52_tab;
    K.Q0_INV_MOD_Q1 = c_inv_mod52(Q[1], Q[0]);
    K.Q01_INV_MOD_Q2 = c_inv_mod52(Q[2], tfo_mul_mod64(Q[2], Q[0], Q[1]));
}
static const crt_consts *crt(void) {
    pthread_once(&K_once, crt_init);
    return &K;
}

/* native32.rs:27-55 */
uint32_t tfo_reconstruct_32bit_012(uint32_t mod_p0, uint32_t mod_p1, uint32_t mod_p2) {
    const crt_consts *k = crt();
    const uint32_t P0 = P32[0], P1 = P32[1], P2 = P32[2];
    uint32_t v0 = mod_p0;
    uint32_t v1 = mul_mod32(P1, k->P0_INV_MOD_P1, 2 * P1 + mod_p1 - v0);
    uint32_t v2 = mul_mod32(P2, k->P01_INV_MOD_P2, 2 * P2 + mod_p2 - (v0 + mul_mod32(P2, P0, v1)));
    int sign = v2 > (P2 / 2);
    const uint32_t _0 = P0, _01 = _0 * P1, _012 = _01 * P2;
    uint32_t pos = v0 + v1 * _0 + v2 * _01;
    uint32_t neg = pos - _012;
    return sign ? neg : pos;
}

/* native32.rs:222-252 */
uint32_t tfo_reconstruct_52bit_01_u32(uint64_t mod_p0, uint64_t mod_p1) {
    const crt_consts *k = crt();
    const uint64_t P0 = primes52_tab[0], P1 = primes52_tab[1];
    uint64_t v0 = mod_p0;
    uint64_t v1 = mul_mod52(P1, 2 * P1 + mod_p1 - v0, k->Q0_INV_MOD_Q1);
    int sign = v1 > (P1 / 2);
    uint64_t pos = v0 + v1 * P0;
    uint64_t neg = pos - P0 * P1;
    return (uint32_t)(sign ? neg : pos);
}

/* native64.rs:90-140 */
uint64_t tfo_reconstruct_32bit_01234_v2(uint32_t mod_p0, uint32_t mod_p1, uint32_t mod_p2,
                                        uint32_t mod_p3, uint32_t mod_p4) {
    const crt_consts *k = crt();
    const uint32_t P0 = P32[0], P1 = P32[1], P2 = P32[2], P3 = P32[3], P4 = P32[4];
    uint64_t mod_p12, mod_p34;
    {
        uint32_t v1 = mod_p1;
        uint32_t v2 = mul_mod32(P2, k->P1_INV_MOD_P2, 2 * P2 + mod_p2 - v1);
        mod_p12 = (uint64_t)v1 + ((uint64_t)v2 * P1);
    }
    {
        uint32_t v3 = mod_p3;
        uint32_t v4 = mul_mod32(P4, k->P3_INV_MOD_P4, 2 * P4 + mod_p4 - v3);
        mod_p34 = (uint64_t)v3 + ((uint64_t)v4 * P3);
    }
    uint64_t v0 = mod_p0;
    uint64_t v12 = mul_mod64s(0 - k->P12, 2 * k->P12 + mod_p12 - v0, k->P0_INV_MOD_P12,
                              k->P0_INV_MOD_P12_SHOUP);
    uint64_t v34 = mul_mod64s(
        0 - k->P34,
        2 * k->P34 + mod_p34 - (v0 + mul_mod64s(0 - k->P34, v12, P0, k->P0_MOD_P34_SHOUP)),
        k->P012_INV_MOD_P34, k->P012_INV_MOD_P34_SHOUP);
    int sign = v34 > (k->P34 / 2);
    const uint64_t _0 = P0, _012 = _0 * k->P12, _01234 = _012 * k->P34;
    uint64_t pos = v0 + v12 * _0 + v34 * _012;
    uint64_t neg = pos - _01234;
    return sign ? neg : pos;
}

/* native64.rs:44-87 (dead code in the reference; kept as a cross-check of the v2 formula) */
uint64_t tfo_reconstruct_32bit_01234(uint32_t mod_p0, uint32_t mod_p1, uint32_t mod_p2,
                                     uint32_t mod_p3, uint32_t mod_p4) {
    const crt_consts *k = crt();
    const uint32_t P0 = P32[0], P1 = P32[1], P2 = P32[2], P3 = P32[3], P4 = P32[4];
    uint32_t v0 = mod_p0;
    uint32_t v1 = mul_mod32(P1, k->P0_INV_MOD_P1, 2 * P1 + mod_p1 - v0);
    uint32_t v2 = mul_mod32(P2, k->P01_INV_MOD_P2, 2 * P2 + mod_p2 - (v0 + mul_mod32(P2, P0, v1)));
    uint32_t v3 = mul_mod32(P3, k->P012_INV_MOD_P3,
                            2 * P3 + mod_p3 - (v0 + mul_mod32(P3, P0, v1 + mul_mod32(P3, P1, v2))));
    uint32_t v4 = mul_mod32(
        P4, k->P0123_INV_MOD_P4,
        2 * P4 + mod_p4 -
            (v0 + mul_mod32(P4, P0, v1 + mul_mod32(P4, P1, v2 + mul_mod32(P4, P2, v3)))));
    int sign = v4 > (P4 / 2);
    const uint64_t _0 = P0, _01 = _0 * P1, _012 = _01 * P2, _0123 = _012 * P3, _01234 = _0123 * P4;
    uint64_t pos = (uint64_t)v0 + (uint64_t)v1 * _0 + (uint64_t)v2 * _01 + (uint64_t)v3 * _012 +
                   (uint64_t)v4 * _0123;
    uint64_t neg = pos - _01234;
    return sign ? neg : pos;
}

/* native64.rs:769-828 */
uint64_t tfo_reconstruct_52bit_012(uint64_t mod_p0, uint64_t mod_p1, uint64_t mod_p2) {
    const crt_consts *k = crt();
    const uint64_t P0 = primes52_tab[0], P1 = primes52_tab[1], P2 = primes52_tab[2];
    uint64_t v0 = mod_p0;
    uint64_t v1 = mul_mod52(P1, 2 * P1 + mod_p1 - v0, k->Q0_INV_MOD_Q1);
    uint64_t v2 =
        mul_mod52(P2, 2 * P2 + mod_p2 - (v0 + mul_mod52(P2, v1, P0)), k->Q01_INV_MOD_Q2);
    int sign = v2 > (P2 / 2);
    uint64_t pos = v0 + v1 * P0 + v2 * (P0 * P1);
    uint64_t neg = pos - P0 * P1 * P2;
    return sign ? neg : pos;
}

/* native128.rs:20-118 */
u128 tfo_reconstruct_32bit_0123456789_v2(const uint32_t r[10]) {
    const crt_consts *k = crt();
    const uint32_t *P = P32;
    uint64_t mod_pair[5];
    const uint32_t pair_inv[5] = {k->P0_INV_MOD_P1, k->P2_INV_MOD_P3, k->P4_INV_MOD_P5,
                                  k->P6_INV_MOD_P7, k->P8_INV_MOD_P9};
    for (int i = 0; i < 5; i++) {
        uint32_t va = r[2 * i];
        uint32_t vb = mul_mod32(P[2 * i + 1], pair_inv[i], 2 * P[2 * i + 1] + r[2 * i + 1] - va);
        mod_pair[i] = (uint64_t)va + ((uint64_t)vb * P[2 * i]);
    }
    uint64_t mod_p23 = mod_pair[1], mod_p45 = mod_pair[2], mod_p67 = mod_pair[3],
             mod_p89 = mod_pair[4];
    uint64_t v01 = mod_pair[0];
    uint64_t v23 = mul_mod64s(0 - k->P23, 2 * k->P23 + mod_p23 - v01, k->P01_INV_MOD_P23,
                              k->P01_INV_MOD_P23_SHOUP);
    uint64_t v45 = mul_mod64s(
        0 - k->P45,
        2 * k->P45 + mod_p45 - (v01 + mul_mod64s(0 - k->P45, v23, k->P01, k->P01_MOD_P45_SHOUP)),
        k->P0123_INV_MOD_P45, k->P0123_INV_MOD_P45_SHOUP);
    uint64_t v67 = mul_mod64s(
        0 - k->P67,
        2 * k->P67 + mod_p67 -
            (v01 + mul_mod64s(0 - k->P67,
                              v23 + mul_mod64s(0 - k->P67, v45, k->P23, k->P23_MOD_P67_SHOUP),
                              k->P01, k->P01_MOD_P67_SHOUP)),
        k->P012345_INV_MOD_P67, k->P012345_INV_MOD_P67_SHOUP);
    uint64_t v89 = mul_mod64s(
        0 - k->P89,
        2 * k->P89 + mod_p89 -
            (v01 +
             mul_mod64s(0 - k->P89,
                        v23 + mul_mod64s(0 - k->P89,
                                         v45 + mul_mod64s(0 - k->P89, v67, k->P45,
                                                          k->P45_MOD_P89_SHOUP),
                                         k->P23, k->P23_MOD_P89_SHOUP),
                        k->P01, k->P01_MOD_P89_SHOUP)),
        k->P01234567_INV_MOD_P89, k->P01234567_INV_MOD_P89_SHOUP);
    int sign = v89 > (k->P89 / 2);
    u128 pos = (u128)v01 + (u128)v23 * (u128)k->P01 + (u128)v45 * k->P0123 +
               (u128)v67 * k->P012345 + (u128)v89 * k->P01234567;
    u128 neg = pos - k->P0123456789;
    return sign ? neg : pos;
}

/* native_binary32.rs:21-40 */
uint32_t tfo_reconstruct_32bit_01(uint32_t mod_p0, uint32_t mod_p1) {
    const crt_consts *k = crt();
    const uint32_t P0 = P32[0], P1 = P32[1];
    uint32_t v0 = mod_p0;
    uint32_t v1 = mul_mod32(P1, k->P0_INV_MOD_P1, 2 * P1 + mod_p1 - v0);
    int sign = v1 > (P1 / 2);
    const uint32_t _0 = P0, _01 = _0 * P1;
    uint32_t pos = v0 + v1 * _0;
    uint32_t neg = pos - _01;
    return sign ? neg : pos;
}

/* native_binary32.rs:110-123 */
uint32_t tfo_reconstruct_52bit_0_u32(uint64_t mod_p0) {
    const uint64_t P0 = primes52_tab[0];
    uint64_t v0 = mod_p0;
    int sign = v0 > (P0 / 2);
    uint64_t pos = v0, neg = pos - P0;
    return (uint32_t)(sign ? neg : pos);
}

/* native_binary64.rs:32-60 */
uint64_t tfo_reconstruct_32bit_012_u64(uint32_t mod_p0, uint32_t mod_p1, uint32_t mod_p2) {
    const crt_consts *k = crt();
    const uint32_t P0 = P32[0], P1 = P32[1], P2 = P32[2];
    uint32_t v0 = mod_p0;
    uint32_t v1 = mul_mod32(P1, k->P0_INV_MOD_P1, 2 * P1 + mod_p1 - v0);
    uint32_t v2 = mul_mod32(P2, k->P01_INV_MOD_P2, 2 * P2 + mod_p2 - (v0 + mul_mod32(P2, P0, v1)));
    int sign = v2 > (P2 / 2);
    const uint64_t _0 = P0, _01 = _0 * P1, _012 = _01 * P2;
    uint64_t pos = (uint64_t)v0 + (uint64_t)v1 * _0 + (uint64_t)v2 * _01;
    uint64_t neg = pos - _012;
    return sign ? neg : pos;
}

/* native_binary64.rs:229-260 */
uint64_t tfo_reconstruct_52bit_01_u64(uint64_t mod_p0, uint64_t mod_p1) {
    const crt_consts *k = crt();
    const uint64_t P0 = primes52_tab[0], P1 = primes52_tab[1];
    uint64_t v0 = mod_p0;
    uint64_t v1 = mul_mod52(P1, 2 * P1 + mod_p1 - v0, k->Q0_INV_MOD_Q1);
    int sign = v1 > (P1 / 2);
    uint64_t pos = v0 + v1 * P0;
    uint64_t neg = pos - P0 * P1;
    return sign ? neg : pos;
}

/* native_binary128.rs:13-63 */
u128 tfo_reconstruct_32bit_01234_v2_u128(uint32_t mod_p0, uint32_t mod_p1, uint32_t mod_p2,
                                         uint32_t mod_p3, uint32_t mod_p4) {
    const crt_consts *k = crt();
    const uint32_t P0 = P32[0], P1 = P32[1], P2 = P32[2], P3 = P32[3], P4 = P32[4];
    uint64_t mod_p12, mod_p34;
    {
        uint32_t v1 = mod_p1;
        uint32_t v2 = mul_mod32(P2, k->P1_INV_MOD_P2, 2 * P2 + mod_p2 - v1);
        mod_p12 = (uint64_t)v1 + ((uint64_t)v2 * P1);
    }
    {
        uint32_t v3 = mod_p3;
        uint32_t v4 = mul_mod32(P4, k->P3_INV_MOD_P4, 2 * P4 + mod_p4 - v3);
        mod_p34 = (uint64_t)v3 + ((uint64_t)v4 * P3);
    }
    uint64_t v0 = mod_p0;
    uint64_t v12 = mul_mod64s(0 - k->P12, 2 * k->P12 + mod_p12 - v0, k->P0_INV_MOD_P12,
                              k->P0_INV_MOD_P12_SHOUP);
    uint64_t v34 = mul_mod64s(
        0 - k->P34,
        2 * k->P34 + mod_p34 - (v0 + mul_mod64s(0 - k->P34, v12, P0, k->P0_MOD_P34_SHOUP)),
        k->P012_INV_MOD_P34, k->P012_INV_MOD_P34_SHOUP);
    int sign = v34 > (k->P34 / 2);
    const u128 _0 = P0, _012 = _0 * (u128)k->P12, _01234 = _012 * (u128)k->P34;
    u128 pos = (u128)v0 + (u128)v12 * _0 + (u128)v34 * _012;
    u128 neg = pos - _01234;
    return sign ? neg : pos;
}

/* ------------------------------------------------------------------ */
/* CRT plans                                                           */
/* ------------------------------------------------------------------ */

static const struct {
    int num_primes, residue_bytes, value_bytes;
} KIND[10] = {
    {3, 4, 4},   /* native32::Plan32 */
    {2, 8, 4},   /* native32::Plan52 */
    {5, 4, 8},   /* native64::Plan32 */
    {3, 8, 8},   /* native64::Plan52 */
    {10, 4, 16}, /* native128::Plan32 */
    {2, 4, 4},   /* native_binary32::Plan32 */
    {1, 8, 4},   /* native_binary32::Plan52 */
    {3, 4, 8},   /* native_binary64::Plan32 */
    {2, 8, 8},   /* native_binary64::Plan52 */
    {5, 4, 16},  /* native_binary128::Plan32 */
};

int tfo_native_num_primes(int kind) { return KIND[kind].num_primes; }
int tfo_native_residue_bytes(int kind) { return KIND[kind].residue_bytes; }
int tfo_native_value_bytes(int kind) { return KIND[kind].value_bytes; }

/* native32.rs:337-345, :440-445; native64.rs:932-941, :1077-1086; native128.rs:123-137;
 * native_binary32.rs:189-192, :262-266; native_binary64.rs:344-351; native_binary128.rs:68-77.
 * Plan52::try_new additionally needs an AVX512-IFMA CPU in the reference; the oracle (like the
 * GPU engine) always provides the 52-bit plans. */
tfo_native_plan *tfo_native_try_new(int kind, size_t n) {
    if (kind < 0 || kind > 9) return NULL;
    tfo_native_plan *pl = (tfo_native_plan *)calloc(1, sizeof(*pl));
    pl->kind = kind;
    pl->n = n;
    pl->num_primes = KIND[kind].num_primes;
    pl->residue_bytes = KIND[kind].residue_bytes;
    pl->value_bytes = KIND[kind].value_bytes;
    for (int i = 0; i < pl->num_primes; i++) {
        int ok;
        if (pl->residue_bytes == 4) {
            pl->p32[i] = tfo_plan32_try_new(n, P32[i]);
            ok = pl->p32[i] != NULL;
        } else {
            pl->p64[i] = tfo_plan64_try_new(n, primes52_tab[i]);
            ok = pl->p64[i] != NULL;
        }
        if (!ok) {
            tfo_native_free(pl);
            return NULL;
        }
    }
    return pl;
}

void tfo_native_free(tfo_native_plan *pl) {
    if (!pl) return;
    for (int i = 0; i < 10; i++) tfo_plan32_free(pl->p32[i]);
    for (int i = 0; i < 3; i++) tfo_plan64_free(pl->p64[i]);
    free(pl);
}

static u128 load_value(const void *value, int value_bytes, size_t i) {
    switch (value_bytes) {
    case 4: return ((const uint32_t *)value)[i];
    case 8: return ((const uint64_t *)value)[i];
    default: {
        u128 v;
        memcpy(&v, (const char *)value + 16 * i, 16);
        return v;
    }
    }
}
static void store_value(void *value, int value_bytes, size_t i, u128 v) {
    switch (value_bytes) {
    case 4: ((uint32_t *)value)[i] = (uint32_t)v; break;
    case 8: ((uint64_t *)value)[i] = (uint64_t)v; break;
    default: memcpy((char *)value + 16 * i, &v, 16); break;
    }
}

/* fwd: native32.rs:365-377, :452-464; native64.rs:970-998, :1107-1126; native128.rs:195-247;
 * fwd_binary: native_binary32.rs:210-217, :283-285; native_binary64.rs:371-388, :479-492;
 * native_binary128.rs fwd_binary.  native32::Plan52 / native_binary32::Plan52 copy `value as u64`
 * without reduction (u32 < P0). */
void tfo_native_fwd(const tfo_native_plan *pl, const void *value, void *const *residues,
                    int binary) {
    size_t n = pl->n;
    int no_reduce =
        binary || pl->kind == TFO_NATIVE32_PLAN52 || pl->kind == TFO_NATIVE_BINARY32_PLAN52;
    for (int k = 0; k < pl->num_primes; k++) {
        if (pl->residue_bytes == 4) {
            uint32_t *r = (uint32_t *)residues[k];
            uint32_t p = pl->p32[k]->p;
            for (size_t i = 0; i < n; i++) {
                u128 v = load_value(value, pl->value_bytes, i);
                r[i] = no_reduce ? (uint32_t)v : (uint32_t)(v % p);
            }
            tfo_plan32_fwd(pl->p32[k], r);
        } else {
            uint64_t *r = (uint64_t *)residues[k];
            uint64_t p = pl->p64[k]->p;
            for (size_t i = 0; i < n; i++) {
                u128 v = load_value(value, pl->value_bytes, i);
                r[i] = no_reduce ? (uint64_t)v : (uint64_t)(v % p);
            }
            tfo_plan64_fwd(pl->p64[k], r);
        }
    }
}

/* inv: native32.rs:379-408, :466-472; native64.rs:1000-1037, :1128-1140; native128.rs:249-292;
 * native_binary32.rs:219-240, :287-292; native_binary64.rs inv; native_binary128.rs inv */
void tfo_native_inv(const tfo_native_plan *pl, void *value, void *const *residues) {
    size_t n = pl->n;
    for (int k = 0; k < pl->num_primes; k++) {
        if (pl->residue_bytes == 4)
            tfo_plan32_inv(pl->p32[k], (uint32_t *)residues[k]);
        else
            tfo_plan64_inv(pl->p64[k], (uint64_t *)residues[k]);
    }
    for (size_t i = 0; i < n; i++) {
        uint32_t r32[10];
        uint64_t r64[3];
        for (int k = 0; k < pl->num_primes; k++) {
            if (pl->residue_bytes == 4)
                r32[k] = ((const uint32_t *)residues[k])[i];
            else
                r64[k] = ((const uint64_t *)residues[k])[i];
        }
        u128 v = 0;
        switch (pl->kind) {
        case TFO_NATIVE32_PLAN32: v = tfo_reconstruct_32bit_012(r32[0], r32[1], r32[2]); break;
        case TFO_NATIVE32_PLAN52: v = tfo_reconstruct_52bit_01_u32(r64[0], r64[1]); break;
        case TFO_NATIVE64_PLAN32:
            v = tfo_reconstruct_32bit_01234_v2(r32[0], r32[1], r32[2], r32[3], r32[4]);
            break;
        case TFO_NATIVE64_PLAN52: v = tfo_reconstruct_52bit_012(r64[0], r64[1], r64[2]); break;
        case TFO_NATIVE128_PLAN32: v = tfo_reconstruct_32bit_0123456789_v2(r32); break;
        case TFO_NATIVE_BINARY32_PLAN32: v = tfo_reconstruct_32bit_01(r32[0], r32[1]); break;
        case TFO_NATIVE_BINARY32_PLAN52: v = tfo_reconstruct_52bit_0_u32(r64[0]); break;
        case TFO_NATIVE_BINARY64_PLAN32:
            v = tfo_reconstruct_32bit_012_u64(r32[0], r32[1], r32[2]);
            break;
        case TFO_NATIVE_BINARY64_PLAN52: v = tfo_reconstruct_52bit_01_u64(r64[0], r64[1]); break;
        case TFO_NATIVE_BINARY128_PLAN32:
            v = tfo_reconstruct_32bit_01234_v2_u128(r32[0], r32[1], r32[2], r32[3], r32[4]);
            break;
        }
        store_value(value, pl->value_bytes, i, v);
    }
}

/* negacyclic_polymul: native32.rs:412-432, :476-497; native64.rs:1041-1068, :1144-1163;
 * native128.rs:296-349; native_binary32.rs:244-261, :296-309; native_binary64.rs:497-515 */
void tfo_native_negacyclic_polymul(const tfo_native_plan *pl, void *prod, const void *lhs,
                                   const void *rhs) {
    size_t n = pl->n;
    int np = pl->num_primes, rb = pl->residue_bytes;
    void *l[10] = {0}, *r[10] = {0};
    for (int k = 0; k < np; k++) {
        l[k] = calloc(n, (size_t)rb);
        r[k] = calloc(n, (size_t)rb);
    }
    int is_binary = pl->kind >= TFO_NATIVE_BINARY32_PLAN32;
    tfo_native_fwd(pl, lhs, l, 0);
    tfo_native_fwd(pl, rhs, r, is_binary);
    for (int k = 0; k < np; k++) {
        if (rb == 4)
            tfo_plan32_mul_assign_normalize(pl->p32[k], (uint32_t *)l[k], (const uint32_t *)r[k], n);
        else
            tfo_plan64_mul_assign_normalize(pl->p64[k], (uint64_t *)l[k], (const uint64_t *)r[k], n);
    }
    tfo_native_inv(pl, prod, l);
    for (int k = 0; k < np; k++) {
        free(l[k]);
        free(r[k]);
    }
}

/* ------------------------------------------------------------------ */
/* batch helpers for the CPU baseline                                  */
/* ------------------------------------------------------------------ */

typedef struct {
    const void *plan;
    void *buf;
    size_t begin, end;
    int op; /* 0 fwd64, 1 inv64, 2 fwd32, 3 inv32 */
} batch_job;

static void *batch_worker(void *arg) {
    batch_job *j = (batch_job *)arg;
    for (size_t b = j->begin; b < j->end; b++) {
        switch (j->op) {
        case 0: {
            const tfo_plan64 *pl = (const tfo_plan64 *)j->plan;
            tfo_plan64_fwd(pl, (uint64_t *)j->buf + b * pl->n);
        } break;
        case 1: {
            const tfo_plan64 *pl = (const tfo_plan64 *)j->plan;
            tfo_plan64_inv(pl, (uint64_t *)j->buf + b * pl->n);
        } break;
        case 2: {
            const tfo_plan32 *pl = (const tfo_plan32 *)j->plan;
            tfo_plan32_fwd(pl, (uint32_t *)j->buf + b * pl->n);
        } break;
        default: {
            const tfo_plan32 *pl = (const tfo_plan32 *)j->plan;
            tfo_plan32_inv(pl, (uint32_t *)j->buf + b * pl->n);
        } break;
        }
    }
    return NULL;
}

static void run_batch(const void *plan, void *buf, size_t batch, int threads, int op) {
    if (threads < 1) threads = 1;
    if ((size_t)threads > batch) threads = batch ? (int)batch : 1;
    pthread_t *tid = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    batch_job *jobs = (batch_job *)calloc((size_t)threads, sizeof(batch_job));
    size_t chunk = (batch + (size_t)threads - 1) / (size_t)threads;
    for (int t = 0; t < threads; t++) {
        size_t b = (size_t)t * chunk, e = b + chunk;
        if (b > batch) b = batch;
        if (e > batch) e = batch;
        jobs[t] = (batch_job){plan, buf, b, e, op};
        if (t == threads - 1)
            batch_worker(&jobs[t]);
        else
            pthread_create(&tid[t], NULL, batch_worker, &jobs[t]);
    }
    for (int t = 0; t + 1 < threads; t++) pthread_join(tid[t], NULL);
    free(tid);
    free(jobs);
}

void tfo_plan64_fwd_batch(const tfo_plan64 *pl, uint64_t *buf, size_t batch, int threads) {
    run_batch(pl, buf, batch, threads, 0);
}
void tfo_plan64_inv_batch(const tfo_plan64 *pl, uint64_t *buf, size_t batch, int threads) {
    run_batch(pl, buf, batch, threads, 1);
}
void tfo_plan32_fwd_batch(const tfo_plan32 *pl, uint32_t *buf, size_t batch, int threads) {
    run_batch(pl, buf, batch, threads, 2);
}
void tfo_plan32_inv_batch(const tfo_plan32 *pl, uint32_t *buf, size_t batch, int threads) {
    run_batch(pl, buf, batch, threads, 3);
}

/* ------------------------------------------------------------------ */
/* product.rs -- plan over a product of distinct primes                */
/* ------------------------------------------------------------------ */

/* product.rs:22-64 (extended Euclid exactly as the reference iterates it) */
static uint64_t modular_inv_u64(uint64_t modulus, uint64_t n) {
    uint64_t old_r = n % modulus, r = modulus, old_s = 1, s = 0;
    while (r != 0) {
        uint64_t q = old_r / r;
        uint64_t nr = old_r - q * r;
        old_r = r;
        r = nr;
        uint64_t qs = (uint64_t)(((u128)q * s) % modulus);
        uint64_t ns = old_s >= qs ? old_s - qs : old_s - qs + modulus; /* sub_mod, product.rs:67-73 */
        old_s = s;
        s = ns;
    }
    return old_s;
}
/* product.rs:85-93 */
static inline uint64_t add_mod_u64(uint64_t modulus, uint64_t a, uint64_t b) {
    uint64_t sum = a + b;
    int overflow = sum < a;
    return (sum >= modulus || overflow) ? sum - modulus : sum;
}
static inline uint64_t sub_mod_u64(uint64_t modulus, uint64_t a, uint64_t b) {
    return a >= b ? a - b : a - b + modulus;
}

static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}

/* product.rs:153-246 */
tfo_product_plan *tfo_product_try_new(size_t n, uint64_t modulus, const uint64_t *factors,
                                      size_t nfactors) {
    if (n % 2 != 0 || nfactors > 64) return NULL;
    uint64_t sorted[64];
    memcpy(sorted, factors, nfactors * sizeof(uint64_t));
    qsort(sorted, nfactors, sizeof(uint64_t), cmp_u64);
    uint64_t prev = 0;
    for (size_t i = 0; i < nfactors; i++) { /* zeros / duplicates */
        if (sorted[i] == prev) return NULL;
        prev = sorted[i];
    }
    size_t start = 0;
    while (start < nfactors && sorted[start] == 1) start++;
    const uint64_t *primes = sorted + start;
    size_t len = nfactors - start;
    u128 prod = 1;
    for (size_t i = 0; i < len; i++) { /* checked_mul */
        prod *= primes[i];
        if (prod >> 64) return NULL;
    }
    if ((uint64_t)prod != modulus) return NULL;
    tfo_product_plan *pl = (tfo_product_plan *)calloc(1, sizeof(*pl));
    pl->n = n;
    pl->modulus = modulus;
    for (size_t i = 0; i < len; i++) {
        if (primes[i] < ((uint64_t)1 << 32)) {
            pl->p32[pl->n32] = tfo_plan32_try_new(n, (uint32_t)primes[i]);
            if (!pl->p32[pl->n32]) {
                tfo_product_free(pl);
                return NULL;
            }
            pl->n32++;
        }
    }
    for (size_t i = 0; i < len; i++) {
        if (primes[i] >= ((uint64_t)1 << 32)) {
            pl->p64[pl->n64] = tfo_plan64_try_new(n, primes[i]);
            if (!pl->p64[pl->n64]) {
                tfo_product_free(pl);
                return NULL;
            }
            pl->n64++;
        }
    }
    for (size_t i = 0; i < len; i++) pl->primes[i] = primes[i];
    /* modular_inverses[offset(j) + i] = p_i^-1 mod p_j for i < j (product.rs:203-225) */
    size_t offset = 0;
    for (size_t j = 0; j < len; j++) {
        for (size_t i = 0; i < j; i++) pl->inverses[offset + i] = modular_inv_u64(primes[j], primes[i]);
        offset += j;
    }
    return pl;
}

void tfo_product_free(tfo_product_plan *pl) {
    if (!pl) return;
    for (int i = 0; i < 16; i++) tfo_plan32_free(pl->p32[i]);
    for (int i = 0; i < 16; i++) tfo_plan64_free(pl->p64[i]);
    free(pl);
}

/* product.rs:261-270 */
size_t tfo_product_ntt_domain_len(const tfo_product_plan *pl) {
    return (pl->n / 2) * (size_t)pl->n32 + pl->n * (size_t)pl->n64;
}

/* product.rs:273-357.  FwdMode::Bounded (the 2 x u32 arm, :304-325) is restated too: for inputs
 * that honour its precondition it produces the same residues as Generic. */
void tfo_product_fwd(const tfo_product_plan *pl, uint64_t *ntt, const uint64_t *standard,
                     int bounded, uint64_t bound) {
    size_t n = pl->n;
    uint32_t *ntt_32 = (uint32_t *)ntt;
    uint64_t *ntt_64 = ntt + (n / 2) * (size_t)pl->n32;
    if (pl->n32 == 0 && pl->n64 == 1) {
        memcpy(ntt_64, standard, n * sizeof(uint64_t));
        tfo_plan64_fwd(pl->p64[0], ntt_64);
        return;
    }
    if (pl->n32 == 1 && pl->n64 == 0) {
        for (size_t i = 0; i < n; i++) ntt_32[i] = (uint32_t)standard[i];
        tfo_plan32_fwd(pl->p32[0], ntt_32);
        return;
    }
    if (pl->n32 == 2 && pl->n64 == 0) {
        uint32_t *ntt0 = ntt_32, *ntt1 = ntt_32 + n;
        uint32_t p0 = pl->p32[0]->p, p1 = pl->p32[1]->p;
        uint64_t p = pl->modulus;
        uint32_t p_u32 = (uint32_t)p;
        if (bounded && bound < p0 && bound < p1) {
            for (size_t i = 0; i < n; i++) {
                int positive = standard[i] < p / 2;
                uint32_t st = (uint32_t)standard[i];
                uint32_t complement = p_u32 - st;
                ntt0[i] = positive ? st : p0 - complement;
                ntt1[i] = positive ? st : p1 - complement;
            }
        } else {
            for (size_t i = 0; i < n; i++) {
                ntt0[i] = (uint32_t)(standard[i] % p0);
                ntt1[i] = (uint32_t)(standard[i] % p1);
            }
        }
        tfo_plan32_fwd(pl->p32[0], ntt0);
        tfo_plan32_fwd(pl->p32[1], ntt1);
        return;
    }
    for (int j = 0; j < pl->n32; j++) {
        uint32_t *r = ntt_32 + (size_t)j * n;
        for (size_t i = 0; i < n; i++) r[i] = (uint32_t)(standard[i] % pl->p32[j]->p);
        tfo_plan32_fwd(pl->p32[j], r);
    }
    for (int j = 0; j < pl->n64; j++) {
        uint64_t *r = ntt_64 + (size_t)j * n;
        for (size_t i = 0; i < n; i++) r[i] = standard[i] % pl->p64[j]->p;
        tfo_plan64_fwd(pl->p64[j], r);
    }
}

/* product.rs:360-880: per-prime inverse transforms, then Knuth 4.3.2 mixed-radix recombination.
 * The special-cased arms (u64x1 :386-398, u32x1 :399-415, u32x2 :420-790) compute the same
 * values as the general loop (:792-879) for canonical residues; the general loop is restated and
 * the single-prime arms, which skip the recombination, explicitly. */
void tfo_product_inv(const tfo_product_plan *pl, uint64_t *standard, uint64_t *ntt, int accumulate) {
    size_t n = pl->n;
    uint32_t *ntt_32 = (uint32_t *)ntt;
    uint64_t *ntt_64 = ntt + (n / 2) * (size_t)pl->n32;
    for (int j = 0; j < pl->n32; j++) tfo_plan32_inv(pl->p32[j], ntt_32 + (size_t)j * n);
    for (int j = 0; j < pl->n64; j++) tfo_plan64_inv(pl->p64[j], ntt_64 + (size_t)j * n);
    uint64_t p = pl->modulus;
    if (pl->n32 == 0 && pl->n64 == 0) {
        if (!accumulate) memset(standard, 0, n * sizeof(uint64_t));
        return;
    }
    if (pl->n32 == 1 && pl->n64 == 0 && accumulate) { /* add_mod_u32 on the truncated word, :406-413 */
        uint32_t q = pl->p32[0]->p;
        for (size_t i = 0; i < n; i++) {
            uint32_t a = (uint32_t)standard[i], b = ntt_32[i];
            uint32_t sum = a + b;
            int overflow = sum < a;
            standard[i] = (sum >= q || overflow) ? (uint32_t)(sum - q) : sum;
        }
        return;
    }
    int c32 = pl->n32, c64 = pl->n64;
    for (size_t idx = 0; idx < n; idx++) {
        uint64_t v[32];
        size_t offset = 0;
        for (int j = 0; j < c32 + c64; j++) {
            uint64_t pj = pl->primes[j];
            uint64_t x = j < c32 ? ntt_32[(size_t)j * n + idx] : ntt_64[(size_t)(j - c32) * n + idx];
            for (int i = 0; i < j; i++) {
                uint64_t diff = sub_mod_u64(pj, x, v[i]);
                x = (uint64_t)(((u128)diff * pl->inverses[offset + (size_t)i]) % pj);
            }
            offset += (size_t)j;
            v[j] = x;
        }
        uint64_t acc = 0;
        for (int j = c32 + c64 - 1; j >= 0; j--) acc = acc * pl->primes[j] + v[j];
        standard[idx] = accumulate ? add_mod_u64(p, standard[idx], acc) : acc;
    }
}

/* product.rs:885-967 */
void tfo_product_mul_assign_normalize(const tfo_product_plan *pl, uint64_t *lhs, const uint64_t *rhs) {
    size_t n = pl->n, off64 = (n / 2) * (size_t)pl->n32;
    for (int j = 0; j < pl->n32; j++)
        tfo_plan32_mul_assign_normalize(pl->p32[j], (uint32_t *)lhs + (size_t)j * n,
                                        (const uint32_t *)rhs + (size_t)j * n, n);
    for (int j = 0; j < pl->n64; j++)
        tfo_plan64_mul_assign_normalize(pl->p64[j], lhs + off64 + (size_t)j * n, rhs + off64 + (size_t)j * n, n);
}
void tfo_product_normalize(const tfo_product_plan *pl, uint64_t *values) {
    size_t n = pl->n, off64 = (n / 2) * (size_t)pl->n32;
    for (int j = 0; j < pl->n32; j++) tfo_plan32_normalize(pl->p32[j], (uint32_t *)values + (size_t)j * n, n);
    for (int j = 0; j < pl->n64; j++) tfo_plan64_normalize(pl->p64[j], values + off64 + (size_t)j * n, n);
}
void tfo_product_mul_accumulate(const tfo_product_plan *pl, uint64_t *acc, const uint64_t *lhs,
                                const uint64_t *rhs) {
    size_t n = pl->n, off64 = (n / 2) * (size_t)pl->n32;
    for (int j = 0; j < pl->n32; j++)
        tfo_plan32_mul_accumulate(pl->p32[j], (uint32_t *)acc + (size_t)j * n,
                                  (const uint32_t *)lhs + (size_t)j * n, (const uint32_t *)rhs + (size_t)j * n, n);
    for (int j = 0; j < pl->n64; j++)
        tfo_plan64_mul_accumulate(pl->p64[j], acc + off64 + (size_t)j * n, lhs + off64 + (size_t)j * n,
                                  rhs + off64 + (size_t)j * n, n);
}

/* ------------------------------------------------------------------ */
/* tfhe::core_crypto::commons::math::ntt::ntt64::Ntt64View             */
/* (tfhe/src/core_crypto/commons/math/ntt/ntt64.rs:89-266): the thin   */
/* wrapper through which the NTT-PBS calls prime64::Plan               */
/* ------------------------------------------------------------------ */

/* mode 0: forward (:89-95); 1: forward_normalized (:97-108); 2: forward_from_decomp (:218-238);
 * 3: forward_from_power_of_two_modulus(width) (:201-214 with the modswitch :165-177) */
/* CPU-baseline switch (see the header): vectorised transforms inside the Ntt64View / PBS restatement */
static int g_simd_transforms = 0;
void tfo_use_simd_transforms(int on) { g_simd_transforms = on; }
static void view_fwd(const tfo_plan64 *pl, uint64_t *buf) {
    if (g_simd_transforms && tfo_plan64_fwd_simd1(pl, buf)) return;
    tfo_plan64_fwd(pl, buf);
}
static void view_inv(const tfo_plan64 *pl, uint64_t *buf) {
    if (g_simd_transforms && tfo_plan64_inv_simd1(pl, buf)) return;
    tfo_plan64_inv(pl, buf);
}

void tfo_ntt64_forward(const tfo_plan64 *pl, uint64_t *ntt, const uint64_t *standard, int mode,
                       uint32_t width) {
    size_t n = pl->n;
    uint64_t p = pl->p;
    memcpy(ntt, standard, n * sizeof(uint64_t));
    if (mode == 2) {
        for (size_t i = 0; i < n; i++)
            if ((int64_t)ntt[i] < 0) ntt[i] = ntt[i] + p;
    } else if (mode == 3) {
        for (size_t i = 0; i < n; i++) {
            u128 v = (u128)ntt[i] >> (64 - width);
            ntt[i] = (uint64_t)(((v * (u128)p) + ((u128)1 << (width - 1))) >> width);
        }
    }
    view_fwd(pl, ntt);
    if (mode == 1) tfo_plan64_normalize(pl, ntt, n);
}

/* mode 0: add_backward (:110-131, wrapping_add_custom_mod); 1: add_backward_on_power_of_two_modulus
 * (:242-266 with the modswitch :184-196).  ntt is transformed (and, in mode 1, modswitched) in
 * place exactly as the reference leaves it. */
void tfo_ntt64_add_backward(const tfo_plan64 *pl, uint64_t *standard, uint64_t *ntt, int mode,
                            uint32_t width) {
    size_t n = pl->n;
    uint64_t p = pl->p;
    view_inv(pl, ntt);
    if (mode == 0) {
        for (size_t i = 0; i < n; i++) {
            /* wrapping_add_custom_mod = a - neg(b) (mod p), tfhe .../numeric/unsigned.rs:174-190, :219-225 */
            uint64_t a = standard[i], b = ntt[i];
            uint64_t nb = b == 0 ? 0 : p - b;
            standard[i] = a >= nb ? a - nb : a - nb + p;
        }
    } else {
        for (size_t i = 0; i < n; i++) {
            u128 x = (((u128)ntt[i]) << width) | ((u128)p >> 1);
            ntt[i] = (uint64_t)(x / p) << (64 - width);
            standard[i] = standard[i] + ntt[i];
        }
    }
}
