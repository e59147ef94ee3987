/*
 * tfhe_ntt_simd.c -- TEST / BASELINE INFRASTRUCTURE ONLY (see tfhe_ntt_oracle.h).
 *
 * AVX-512 port of the reference's vectorised Solinas path, used only by bench.py's CPU-baseline
 * legs so that the reported CPU number is not a scalar strawman:
 *   widening 64x64->128 multiply from 32-bit products   tfhe-ntt/src/lib.rs:175-207
 *   Solinas add / sub / mul on u64x8                    prime64/generic_solinas.rs:324-446
 *   breadth-first drivers (splat twiddle while t >= 8, then three in-register stages)
 *                                                       prime64/generic_solinas.rs:801-1000
 * Results are identical to the scalar oracle (tests/test_oracle_simd.py).
 * Compiled into libtfhe_ntt_oracle_native.so with -march=native; without AVX-512F the entry
 * points return 0 and the caller falls back to the scalar oracle.
 */
#include <pthread.h>
#include <stdlib.h>

#include "tfhe_ntt_oracle.h"

#define SOLINAS_P 0xFFFFFFFF00000001ull

#if defined(__AVX512F__) && defined(__AVX512DQ__)
#include <immintrin.h>

typedef __m512i v8;

/* lib.rs:175-207 */
static inline void widening_mul(v8 x, v8 y, v8 *lo, v8 *hi) {
    const v8 lo_mask = _mm512_set1_epi64(0x00000000FFFFFFFFll);
    v8 x_hi = _mm512_shuffle_epi32(x, 0xB1), y_hi = _mm512_shuffle_epi32(y, 0xB1);
    v8 z_lo_lo = _mm512_mul_epu32(x, y), z_lo_hi = _mm512_mul_epu32(x, y_hi);
    v8 z_hi_lo = _mm512_mul_epu32(x_hi, y), z_hi_hi = _mm512_mul_epu32(x_hi, y_hi);
    v8 sum_tmp = _mm512_add_epi64(z_lo_hi, _mm512_srli_epi64(z_lo_lo, 32));
    v8 sum_lo = _mm512_and_si512(sum_tmp, lo_mask), sum_mid = _mm512_srli_epi64(sum_tmp, 32);
    v8 sum_mid2 = _mm512_add_epi64(z_hi_lo, sum_lo);
    v8 sum_hi = _mm512_add_epi64(z_hi_hi, sum_mid);
    *hi = _mm512_add_epi64(sum_hi, _mm512_srli_epi64(sum_mid2, 32));
    *lo = _mm512_add_epi64(_mm512_slli_epi64(_mm512_add_epi64(z_lo_hi, z_hi_lo), 32), z_lo_lo);
}
/* generic_solinas.rs:415-446 */
static inline v8 sol_mul(v8 a, v8 b) {
    const v8 p = _mm512_set1_epi64((long long)SOLINAS_P);
    v8 lo, hi;
    widening_mul(a, b, &lo, &hi);
    v8 mid = _mm512_and_si512(hi, _mm512_set1_epi64(0x00000000FFFFFFFFll));
    hi = _mm512_srli_epi64(hi, 32);
    v8 low2 = _mm512_sub_epi64(lo, hi);
    low2 = _mm512_mask_add_epi64(low2, _mm512_cmpgt_epu64_mask(hi, lo), low2, p);
    v8 product = _mm512_sub_epi64(_mm512_slli_epi64(mid, 32), mid);
    v8 result = _mm512_add_epi64(low2, product);
    __mmask8 keep = (__mmask8)(~_mm512_cmpgt_epu64_mask(product, result) & _mm512_cmpgt_epu64_mask(p, result));
    return _mm512_mask_sub_epi64(result, (__mmask8)~keep, result, p);
}
/* generic_solinas.rs:324-360 (u64 modulus add/sub on u64x8) */
static inline v8 sol_add(v8 a, v8 b) {
    const v8 p = _mm512_set1_epi64((long long)SOLINAS_P);
    v8 neg_b = _mm512_sub_epi64(p, b);
    __mmask8 ge = _mm512_cmpge_epu64_mask(a, neg_b);
    return _mm512_mask_sub_epi64(_mm512_add_epi64(a, b), ge, a, neg_b);
}
static inline v8 sol_sub(v8 a, v8 b) {
    const v8 p = _mm512_set1_epi64((long long)SOLINAS_P);
    v8 neg_b = _mm512_sub_epi64(p, b);
    __mmask8 ge = _mm512_cmpge_epu64_mask(a, b);
    return _mm512_mask_sub_epi64(_mm512_add_epi64(a, neg_b), ge, a, b);
}

/* index vectors that gather the z0 / z1 halves of two consecutive vectors for t = 4, 2, 1 */
static const long long IDX_Z0[3][8] = {{0, 1, 2, 3, 8, 9, 10, 11}, {0, 1, 4, 5, 8, 9, 12, 13}, {0, 2, 4, 6, 8, 10, 12, 14}};
static const long long IDX_Z1[3][8] = {{4, 5, 6, 7, 12, 13, 14, 15}, {2, 3, 6, 7, 10, 11, 14, 15}, {1, 3, 5, 7, 9, 11, 13, 15}};
static const long long IDX_LO[3][8] = {{0, 1, 2, 3, 8, 9, 10, 11}, {0, 1, 8, 9, 2, 3, 10, 11}, {0, 8, 1, 9, 2, 10, 3, 11}};
static const long long IDX_HI[3][8] = {{4, 5, 6, 7, 12, 13, 14, 15}, {4, 5, 12, 13, 6, 7, 14, 15}, {4, 12, 5, 13, 6, 14, 7, 15}};
static const long long IDX_TW[3][8] = {{0, 0, 0, 0, 1, 1, 1, 1}, {0, 0, 1, 1, 2, 2, 3, 3}, {0, 1, 2, 3, 4, 5, 6, 7}};

static void fwd_solinas_avx512(uint64_t *data, size_t n, const uint64_t *twid) {
    size_t t = n / 2, m = 1;
    while (t >= 8) {
        for (size_t i = 0; i < m; i++) {
            v8 w = _mm512_set1_epi64((long long)twid[m + i]);
            uint64_t *z0 = data + 2 * i * t, *z1 = z0 + t;
            for (size_t j = 0; j < t; j += 8) {
                v8 a = _mm512_loadu_si512(z0 + j), b = _mm512_loadu_si512(z1 + j);
                v8 bw = sol_mul(b, w);
                _mm512_storeu_si512(z0 + j, sol_add(a, bw));
                _mm512_storeu_si512(z1 + j, sol_sub(a, bw));
            }
        }
        t /= 2;
        m *= 2;
    }
    for (int s = 0; s < 3; s++, t /= 2, m *= 2) { /* t = 4, 2, 1 on 16 coefficients at a time */
        v8 iz0 = _mm512_loadu_si512(IDX_Z0[s]), iz1 = _mm512_loadu_si512(IDX_Z1[s]);
        v8 ilo = _mm512_loadu_si512(IDX_LO[s]), ihi = _mm512_loadu_si512(IDX_HI[s]);
        v8 itw = _mm512_loadu_si512(IDX_TW[s]);
        size_t tw_per_16 = 8 / t; /* 2, 4, 8 */
        for (size_t e = 0; e < n; e += 16) {
            v8 A = _mm512_loadu_si512(data + e), B = _mm512_loadu_si512(data + e + 8);
            v8 a = _mm512_permutex2var_epi64(A, iz0, B), b = _mm512_permutex2var_epi64(A, iz1, B);
            v8 wraw = _mm512_maskz_loadu_epi64((__mmask8)((1u << tw_per_16) - 1), twid + m + e / (2 * t));
            v8 w = _mm512_permutexvar_epi64(itw, wraw);
            v8 bw = sol_mul(b, w);
            v8 x = sol_add(a, bw), y = sol_sub(a, bw);
            _mm512_storeu_si512(data + e, _mm512_permutex2var_epi64(x, ilo, y));
            _mm512_storeu_si512(data + e + 8, _mm512_permutex2var_epi64(x, ihi, y));
        }
    }
}

static void inv_solinas_avx512(uint64_t *data, size_t n, const uint64_t *inv_twid) {
    size_t t = 1, m = n;
    for (int s = 2; s >= 0; s--, t *= 2) { /* t = 1, 2, 4 */
        m /= 2;
        v8 iz0 = _mm512_loadu_si512(IDX_Z0[s]), iz1 = _mm512_loadu_si512(IDX_Z1[s]);
        v8 ilo = _mm512_loadu_si512(IDX_LO[s]), ihi = _mm512_loadu_si512(IDX_HI[s]);
        v8 itw = _mm512_loadu_si512(IDX_TW[s]);
        size_t tw_per_16 = 8 / t;
        for (size_t e = 0; e < n; e += 16) {
            v8 A = _mm512_loadu_si512(data + e), B = _mm512_loadu_si512(data + e + 8);
            v8 a = _mm512_permutex2var_epi64(A, iz0, B), b = _mm512_permutex2var_epi64(A, iz1, B);
            v8 wraw = _mm512_maskz_loadu_epi64((__mmask8)((1u << tw_per_16) - 1), inv_twid + m + e / (2 * t));
            v8 w = _mm512_permutexvar_epi64(itw, wraw);
            v8 x = sol_add(a, b), y = sol_mul(sol_sub(a, b), w);
            _mm512_storeu_si512(data + e, _mm512_permutex2var_epi64(x, ilo, y));
            _mm512_storeu_si512(data + e + 8, _mm512_permutex2var_epi64(x, ihi, y));
        }
    }
    while (m > 1) {
        m /= 2;
        for (size_t i = 0; i < m; i++) {
            v8 w = _mm512_set1_epi64((long long)inv_twid[m + i]);
            uint64_t *z0 = data + 2 * i * t, *z1 = z0 + t;
            for (size_t j = 0; j < t; j += 8) {
                v8 a = _mm512_loadu_si512(z0 + j), b = _mm512_loadu_si512(z1 + j);
                _mm512_storeu_si512(z0 + j, sol_add(a, b));
                _mm512_storeu_si512(z1 + j, sol_mul(sol_sub(a, b), w));
            }
        }
        t *= 2;
    }
}

typedef struct {
    const tfo_plan64 *pl;
    uint64_t *buf;
    size_t begin, end;
    int inverse;
} simd_job;

static void *simd_worker(void *arg) {
    simd_job *j = (simd_job *)arg;
    size_t n = j->pl->n;
    for (size_t b = j->begin; b < j->end; b++) {
        if (j->inverse)
            inv_solinas_avx512(j->buf + b * n, n, j->pl->inv_twid);
        else
            fwd_solinas_avx512(j->buf + b * n, n, j->pl->twid);
    }
    return NULL;
}

static int run_simd(const tfo_plan64 *pl, uint64_t *buf, size_t batch, int threads, int inverse) {
    if (pl->p != SOLINAS_P || pl->n < 16 || !__builtin_cpu_supports("avx512f") || !__builtin_cpu_supports("avx512dq"))
        return 0;
    if (threads < 1) threads = 1;
    if ((size_t)threads > batch) threads = batch ? (int)batch : 1;
    pthread_t *tid = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    simd_job *jobs = (simd_job *)calloc((size_t)threads, sizeof(simd_job));
    size_t chunk = (batch + (size_t)threads - 1) / (size_t)threads;
    for (int t = 0; t < threads; t++) {
        size_t b = (size_t)t * chunk, e = b + chunk;
        if (b > batch) b = batch;
        if (e > batch) e = batch;
        jobs[t] = (simd_job){pl, buf, b, e, inverse};
        if (t == threads - 1)
            simd_worker(&jobs[t]);
        else
            pthread_create(&tid[t], NULL, simd_worker, &jobs[t]);
    }
    for (int t = 0; t + 1 < threads; t++) pthread_join(tid[t], NULL);
    free(tid);
    free(jobs);
    return 1;
}
static int simd_ok(const tfo_plan64 *pl) {
    return pl->p == SOLINAS_P && pl->n >= 16 && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512dq");
}
int tfo_plan64_fwd_simd1(const tfo_plan64 *pl, uint64_t *buf) {
    if (!simd_ok(pl)) return 0;
    fwd_solinas_avx512(buf, pl->n, pl->twid);
    return 1;
}
int tfo_plan64_inv_simd1(const tfo_plan64 *pl, uint64_t *buf) {
    if (!simd_ok(pl)) return 0;
    inv_solinas_avx512(buf, pl->n, pl->inv_twid);
    return 1;
}
int tfo_plan64_fwd_batch_simd(const tfo_plan64 *pl, uint64_t *buf, size_t batch, int threads) {
    return run_simd(pl, buf, batch, threads, 0);
}
int tfo_plan64_inv_batch_simd(const tfo_plan64 *pl, uint64_t *buf, size_t batch, int threads) {
    return run_simd(pl, buf, batch, threads, 1);
}
#else
int tfo_plan64_fwd_simd1(const tfo_plan64 *pl, uint64_t *buf) {
    (void)pl; (void)buf;
    return 0;
}
int tfo_plan64_inv_simd1(const tfo_plan64 *pl, uint64_t *buf) {
    (void)pl; (void)buf;
    return 0;
}
int tfo_plan64_fwd_batch_simd(const tfo_plan64 *pl, uint64_t *buf, size_t batch, int threads) {
    (void)pl; (void)buf; (void)batch; (void)threads;
    return 0;
}
int tfo_plan64_inv_batch_simd(const tfo_plan64 *pl, uint64_t *buf, size_t batch, int threads) {
    (void)pl; (void)buf; (void)batch; (void)threads;
    return 0;
}
#endif
