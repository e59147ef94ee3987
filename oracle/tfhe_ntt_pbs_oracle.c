/*
 * tfhe_ntt_pbs_oracle.c -- TEST INFRASTRUCTURE ONLY (see tfhe_ntt_oracle.h).
 *
 * CPU restatement of the NTT programmable bootstrap that calls the hot path: SURVEY.md section 8f
 * row 1.  Reference files (all under /root/reference/tfhe/src/core_crypto):
 *   algorithms/lwe_programmable_bootstrapping/ntt64_pbs.rs      (classic: ciphertext modulus = NTT prime)
 *   algorithms/lwe_programmable_bootstrapping/ntt64_bnf_pbs.rs  (bnf: power-of-two ciphertext modulus)
 *   commons/math/decomposition/{decomposer,iter}.rs, fft_impl/fft64/math/decomposition.rs
 *   algorithms/polynomial_algorithms.rs, algorithms/glwe_sample_extraction.rs,
 *   algorithms/lwe_bootstrap_key_conversion.rs, fft_impl/common.rs, algorithms/misc.rs
 *
 * Layouts (the reference's flat containers):
 *   lwe   [n_lwe + 1]                      mask ..., body                 (entities/lwe_ciphertext.rs)
 *   glwe  [(k+1) * N]                      k mask polynomials, body       (entities/glwe_ciphertext.rs)
 *   bsk   [n_lwe][l][k+1][k+1][N]          NTT domain; the first level slice is level l
 *                                          (entities/ntt_ggsw_ciphertext.rs:176-190)
 */
#include "tfhe_ntt_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef tfo_u128 u128;

static unsigned ceil_ilog2_u64(uint64_t x) { /* x >= 2 */
    unsigned b = 0;
    uint64_t v = x - 1;
    while (v) {
        b++;
        v >>= 1;
    }
    return b;
}

/* algorithms/misc.rs:6-18 */
static u128 divide_round_u128(u128 num, u128 den) {
    u128 d = num / den, r = num % den;
    return d + (r >= (den >> 1) ? 1 : 0);
}

/* ntt64_pbs.rs:540-550 */
uint64_t tfo_pbs_modulus_switch_non_native(uint64_t input, uint32_t log2_poly_size, uint64_t modulus) {
    return (uint64_t)divide_round_u128((u128)input << (log2_poly_size + 1), (u128)modulus);
}

/* fft_impl/common.rs:10-23 */
uint64_t tfo_modulus_switch(uint64_t input, uint32_t log_modulus) {
    if (log_modulus == 64) return input;
    uint64_t t = input + ((uint64_t)1 << (64 - log_modulus - 1));
    return t >> (64 - log_modulus);
}

/* commons/numeric/unsigned.rs:219-225 and the native wrapping_neg; modulus 0 selects native */
static uint64_t neg_mod(uint64_t a, uint64_t modulus) {
    if (modulus == 0) return (uint64_t)0 - a;
    return a == 0 ? 0 : modulus - a;
}
/* commons/numeric/unsigned.rs:181-187 */
static uint64_t sub_mod(uint64_t a, uint64_t b, uint64_t modulus) {
    if (modulus == 0) return a - b;
    return a >= b ? a - b : a - b + modulus;
}

static void rotate_left(uint64_t *v, size_t n, size_t r) {
    if (r == 0) return;
    uint64_t *t = (uint64_t *)malloc(n * sizeof(uint64_t));
    for (size_t i = 0; i < n; i++) t[i] = v[(i + r) % n];
    memcpy(v, t, n * sizeof(uint64_t));
    free(t);
}

/* polynomial_algorithms.rs:485-507 (custom modulus) and :462-483 (native, modulus == 0) */
void tfo_monomial_mul_assign(uint64_t *poly, size_t n, size_t degree, uint64_t modulus) {
    size_t full = degree / n;
    if (full % 2 != 0)
        for (size_t i = 0; i < n; i++) poly[i] = neg_mod(poly[i], modulus);
    size_t rem = degree % n;
    rotate_left(poly, n, (n - rem) % n); /* rotate_right(rem) */
    for (size_t i = 0; i < rem; i++) poly[i] = neg_mod(poly[i], modulus);
}

/* polynomial_algorithms.rs:419-442 (custom modulus) and :395-417 (native) */
void tfo_monomial_div_assign(uint64_t *poly, size_t n, size_t degree, uint64_t modulus) {
    size_t full = degree / n;
    if (full % 2 != 0)
        for (size_t i = 0; i < n; i++) poly[i] = neg_mod(poly[i], modulus);
    size_t rem = degree % n;
    rotate_left(poly, n, rem);
    for (size_t i = 0; i < rem; i++) poly[n - 1 - i] = neg_mod(poly[n - 1 - i], modulus);
}

/* decomposer.rs:25-49 */
static uint64_t native_closest_representable(uint64_t input, unsigned level, unsigned base_log) {
    unsigned non_rep = 64 - level * base_log;
    unsigned shift = non_rep - 1;
    uint64_t res = input >> shift;
    res += 1;
    res &= ~(uint64_t)1;
    return res << shift;
}

/* SignedDecomposerNonNative::closest_representable, decomposer.rs:487-557 */
uint64_t tfo_closest_representable_non_native(uint64_t input, uint32_t base_log, uint32_t level,
                                              uint64_t modulus) {
    uint64_t half_up = modulus / 2 + (modulus & 1); /* div_ceil(2) */
    int negative = !(input < half_up);
    uint64_t abs_value = negative ? modulus - input : input;
    unsigned shift_to_native = 64 - ceil_ilog2_u64(modulus);
    uint64_t abs_closest =
        native_closest_representable(abs_value << shift_to_native, level, base_log) >> shift_to_native;
    return negative ? neg_mod(abs_closest, modulus) : abs_closest;
}

/* SignedDecomposer::init_decomposer_state, decomposer.rs:204-236 (with the bit trick :52-60) */
uint64_t tfo_init_decomposer_state_native(uint64_t input, uint32_t base_log, uint32_t level) {
    unsigned rep = level * base_log;
    unsigned non_rep = 64 - rep;
    uint64_t res = input >> (non_rep - 1);
    uint64_t rounding_bit = res & 1;
    res += 1;
    res >>= 1;
    uint64_t mod_mask = ~(uint64_t)0 >> (64 - rep);
    res &= mod_mask;
    uint64_t shifted_random = rounding_bit << (rep - 1);
    uint64_t need_balance = (((res - 1) | shifted_random) & res) >> (rep - 1);
    return res - (need_balance << rep);
}

/* iter.rs:130-151 */
static uint64_t decompose_one_level(unsigned base_log, uint64_t *state, uint64_t mod_b_mask) {
    uint64_t res = *state & mod_b_mask;
    *state = (uint64_t)((int64_t)*state >> base_log);
    uint64_t carry = (((res - 1) | *state) & res) >> (base_log - 1);
    *state += carry;
    return res - (carry << base_log);
}

/* TensorSignedDecompositionLendingIterNonNative::new, iter.rs:640-686 */
void tfo_decomp_non_native_init(const uint64_t *input, size_t len, uint32_t base_log, uint32_t level,
                                uint64_t modulus, uint64_t *states, uint8_t *signs) {
    unsigned shift = ceil_ilog2_u64(modulus) - base_log * level;
    uint64_t half_up = modulus / 2 + (modulus & 1);
    for (size_t i = 0; i < len; i++) {
        if (input[i] < half_up) {
            states[i] = tfo_closest_representable_non_native(input[i], base_log, level, modulus) >> shift;
            signs[i] = 0;
        } else {
            states[i] =
                tfo_closest_representable_non_native(modulus - input[i], base_log, level, modulus) >> shift;
            signs[i] = 1;
        }
    }
}

/* next_term, iter.rs:689-737: one level (the iterator runs from level l down to 1) */
void tfo_decomp_non_native_next(uint64_t *states, const uint8_t *signs, size_t len, uint32_t base_log,
                                uint64_t modulus, uint64_t *term) {
    uint64_t mask = ((uint64_t)1 << base_log) - 1;
    for (size_t i = 0; i < len; i++) {
        uint64_t t = decompose_one_level(base_log, &states[i], mask);
        if (signs[i]) t = (uint64_t)0 - t;
        term[i] = (int64_t)t >= 0 ? t : modulus + t;
    }
}

/* fft_impl/fft64/math/decomposition.rs:41-75 over init_decomposer_state (ntt64_bnf_pbs.rs:591-599) */
void tfo_decomp_native_init(const uint64_t *input, size_t len, uint32_t base_log, uint32_t level,
                            uint64_t *states) {
    for (size_t i = 0; i < len; i++) states[i] = tfo_init_decomposer_state_native(input[i], base_log, level);
}
void tfo_decomp_native_next(uint64_t *states, size_t len, uint32_t base_log, uint64_t *term) {
    uint64_t mask = ((uint64_t)1 << base_log) - 1;
    for (size_t i = 0; i < len; i++) term[i] = decompose_one_level(base_log, &states[i], mask);
}

/* add_external_product_ntt64_assign, ntt64_pbs.rs:553-663 (bnf == 0), and
 * add_external_product_ntt64_bnf_assign, ntt64_bnf_pbs.rs:541-681 (bnf != 0, width = log2 of the
 * power-of-two ciphertext modulus).  ggsw is [l][k+1][k+1][N]. */
void tfo_add_external_product_ntt64_assign(const tfo_plan64 *pl, uint64_t *out, const uint64_t *ggsw,
                                           const uint64_t *glwe, size_t glwe_size, uint32_t base_log,
                                           uint32_t level, int bnf, uint32_t width) {
    size_t n = pl->n, len = glwe_size * n;
    uint64_t p = pl->p;
    uint64_t *acc = (uint64_t *)calloc(len, sizeof(uint64_t)); /* update_with_fmadd: fill(0) first */
    uint64_t *states = (uint64_t *)malloc(len * sizeof(uint64_t));
    uint8_t *signs = (uint8_t *)malloc(len);
    uint64_t *term = (uint64_t *)malloc(len * sizeof(uint64_t));
    uint64_t *ntt_poly = (uint64_t *)malloc(n * sizeof(uint64_t));
    if (bnf)
        tfo_decomp_native_init(glwe, len, base_log, level, states);
    else
        tfo_decomp_non_native_init(glwe, len, base_log, level, p, states, signs);
    for (uint32_t lv = 0; lv < level; lv++) { /* ggsw.into_levels(): slice lv is level l - lv */
        if (bnf)
            tfo_decomp_native_next(states, len, base_log, term);
        else
            tfo_decomp_non_native_next(states, signs, len, base_log, p, term);
        const uint64_t *matrix = ggsw + (size_t)lv * glwe_size * glwe_size * n;
        for (size_t row = 0; row < glwe_size; row++) {
            /* ntt.forward (ntt64.rs:89-95) or forward_from_decomp (:218-238) */
            tfo_ntt64_forward(pl, ntt_poly, term + row * n, bnf ? 2 : 0, 0);
            const uint64_t *ggsw_row = matrix + row * glwe_size * n;
            for (size_t col = 0; col < glwe_size; col++) /* update_with_fmadd_ntt64, :683-702 */
                tfo_plan64_mul_accumulate(pl, acc + col * n, ggsw_row + col * n, ntt_poly, n);
        }
    }
    for (size_t col = 0; col < glwe_size; col++) {
        if (bnf) { /* ntt64_bnf_pbs.rs:669-673 */
            tfo_plan64_normalize(pl, acc + col * n, n);
            tfo_ntt64_add_backward(pl, out + col * n, acc + col * n, 1, width);
        } else { /* ntt64_pbs.rs:652-661 */
            tfo_ntt64_add_backward(pl, out + col * n, acc + col * n, 0, 0);
        }
    }
    free(acc);
    free(states);
    free(signs);
    free(term);
    free(ntt_poly);
}

/* cmux_ntt64_assign, ntt64_pbs.rs:669-680 / cmux_ntt64_bnf_assign, ntt64_bnf_pbs.rs:683-705 */
void tfo_cmux_ntt64_assign(const tfo_plan64 *pl, uint64_t *ct0, uint64_t *ct1, const uint64_t *ggsw,
                           size_t glwe_size, uint32_t base_log, uint32_t level, int bnf, uint32_t width) {
    size_t len = glwe_size * pl->n;
    for (size_t i = 0; i < len; i++) ct1[i] = sub_mod(ct1[i], ct0[i], bnf ? 0 : pl->p);
    tfo_add_external_product_ntt64_assign(pl, ct0, ggsw, ct1, glwe_size, base_log, level, bnf, width);
}

static unsigned log2_exact(size_t n) {
    unsigned l = 0;
    while (((size_t)1 << l) < n) l++;
    return l;
}

/* blind_rotate_ntt64_assign_mem_optimized, ntt64_pbs.rs:213-286 */
void tfo_blind_rotate_ntt64_assign(const tfo_plan64 *pl, const uint64_t *bsk, size_t n_lwe,
                                   size_t glwe_size, uint32_t base_log, uint32_t level,
                                   const uint64_t *lwe, uint64_t *lut) {
    size_t n = pl->n, len = glwe_size * n;
    uint64_t p = pl->p;
    unsigned logn = log2_exact(n);
    size_t ggsw_len = (size_t)level * glwe_size * glwe_size * n;
    uint64_t body = lwe[n_lwe];
    size_t deg = (size_t)tfo_pbs_modulus_switch_non_native(body, logn, p);
    for (size_t c = 0; c < glwe_size; c++) tfo_monomial_div_assign(lut + c * n, n, deg, p);
    uint64_t *ct1 = (uint64_t *)malloc(len * sizeof(uint64_t));
    for (size_t i = 0; i < n_lwe; i++) {
        if (lwe[i] == 0) continue;
        memcpy(ct1, lut, len * sizeof(uint64_t));
        size_t a = (size_t)tfo_pbs_modulus_switch_non_native(lwe[i], logn, p);
        for (size_t c = 0; c < glwe_size; c++) tfo_monomial_mul_assign(ct1 + c * n, n, a, p);
        tfo_cmux_ntt64_assign(pl, lut, ct1, bsk + i * ggsw_len, glwe_size, base_log, level, 0, 0);
    }
    free(ct1);
}

/* blind_rotate_ntt64_bnf_assign_mem_optimized, ntt64_bnf_pbs.rs:208-276; msed = the modulus
 * switched ciphertext ([n_lwe] mask values then the body, each in [0, 2N)) */
void tfo_blind_rotate_ntt64_bnf_assign(const tfo_plan64 *pl, const uint64_t *bsk, size_t n_lwe,
                                       size_t glwe_size, uint32_t base_log, uint32_t level,
                                       uint32_t width, const uint64_t *msed, uint64_t *lut) {
    size_t n = pl->n, len = glwe_size * n;
    size_t ggsw_len = (size_t)level * glwe_size * glwe_size * n;
    uint64_t *ct1 = (uint64_t *)malloc(len * sizeof(uint64_t));
    for (size_t i = 0; i < n_lwe; i++) {
        if (msed[i] == 0) continue;
        memcpy(ct1, lut, len * sizeof(uint64_t));
        for (size_t c = 0; c < glwe_size; c++) tfo_monomial_mul_assign(ct1 + c * n, n, (size_t)msed[i], 0);
        tfo_cmux_ntt64_assign(pl, lut, ct1, bsk + i * ggsw_len, glwe_size, base_log, level, 1, width);
    }
    free(ct1);
    for (size_t c = 0; c < glwe_size; c++) tfo_monomial_div_assign(lut + c * n, n, (size_t)msed[n_lwe], 0);
}

/* extract_lwe_sample_from_glwe_ciphertext, glwe_sample_extraction.rs:89-164; modulus 0 = native */
void tfo_extract_lwe_sample(const uint64_t *glwe, size_t glwe_size, size_t n, size_t nth,
                            uint64_t modulus, uint64_t *lwe_out) {
    size_t k = glwe_size - 1;
    lwe_out[k * n] = glwe[k * n + nth];
    memcpy(lwe_out, glwe, k * n * sizeof(uint64_t));
    size_t opposite_count = n - nth - 1;
    for (size_t c = 0; c < k; c++) {
        uint64_t *poly = lwe_out + c * n;
        for (size_t i = 0; i < n / 2; i++) {
            uint64_t t = poly[i];
            poly[i] = poly[n - 1 - i];
            poly[n - 1 - i] = t;
        }
        for (size_t i = 0; i < opposite_count; i++) poly[i] = neg_mod(poly[i], modulus);
        rotate_left(poly, n, opposite_count);
    }
}

/* programmable_bootstrap_ntt64_lwe_ciphertext_mem_optimized, ntt64_pbs.rs:482-538 */
void tfo_programmable_bootstrap_ntt64(const tfo_plan64 *pl, const uint64_t *bsk, size_t n_lwe,
                                      size_t glwe_size, uint32_t base_log, uint32_t level,
                                      const uint64_t *lwe_in, uint64_t *lwe_out,
                                      const uint64_t *accumulator) {
    size_t len = glwe_size * pl->n;
    uint64_t *local = (uint64_t *)malloc(len * sizeof(uint64_t));
    memcpy(local, accumulator, len * sizeof(uint64_t));
    tfo_blind_rotate_ntt64_assign(pl, bsk, n_lwe, glwe_size, base_log, level, lwe_in, local);
    tfo_extract_lwe_sample(local, glwe_size, pl->n, 0, pl->p, lwe_out);
    free(local);
}

/* A batch of the above over host threads, static contiguous chunks (how the reference's callers
 * parallelise independent ciphertexts with rayon, e.g. lwe_bootstrap_key_conversion.rs:419-447):
 * bench.py's CPU arm for the PBS. */
typedef struct {
    const tfo_plan64 *pl;
    const uint64_t *bsk, *lwe_in, *accumulator;
    uint64_t *lwe_out;
    size_t n_lwe, glwe_size, acc_count, begin, end;
    uint32_t base_log, level;
} pbs_job;
static void *pbs_worker(void *arg) {
    pbs_job *j = (pbs_job *)arg;
    size_t per = j->glwe_size * j->pl->n, out = (j->glwe_size - 1) * j->pl->n + 1;
    for (size_t b = j->begin; b < j->end; b++)
        tfo_programmable_bootstrap_ntt64(j->pl, j->bsk, j->n_lwe, j->glwe_size, j->base_log, j->level,
                                         j->lwe_in + b * (j->n_lwe + 1), j->lwe_out + b * out,
                                         j->accumulator + (j->acc_count == 1 ? 0 : b) * per);
    return NULL;
}
void tfo_programmable_bootstrap_ntt64_batch(const tfo_plan64 *pl, const uint64_t *bsk, size_t n_lwe,
                                            size_t glwe_size, uint32_t base_log, uint32_t level,
                                            const uint64_t *lwe_in, uint64_t *lwe_out,
                                            const uint64_t *accumulator, size_t acc_count, size_t batch,
                                            int threads) {
    if (threads < 1) threads = 1;
    if ((size_t)threads > batch) threads = batch ? (int)batch : 1;
    pthread_t *tid = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    pbs_job *jobs = (pbs_job *)calloc((size_t)threads, sizeof(pbs_job));
    size_t chunk = (batch + (size_t)threads - 1) / (size_t)threads;
    for (int t = 0; t < threads; t++) {
        size_t b = (size_t)t * chunk, e = b + chunk;
        if (b > batch) b = batch;
        if (e > batch) e = batch;
        jobs[t] = (pbs_job){pl, bsk, lwe_in, accumulator, lwe_out, n_lwe, glwe_size, acc_count, b, e, base_log, level};
        if (t == threads - 1)
            pbs_worker(&jobs[t]);
        else
            pthread_create(&tid[t], NULL, pbs_worker, &jobs[t]);
    }
    for (int t = 0; t + 1 < threads; t++) pthread_join(tid[t], NULL);
    free(tid);
    free(jobs);
}

/* programmable_bootstrap_ntt64_bnf_lwe_ciphertext_mem_optimized, ntt64_bnf_pbs.rs:469-539 */
void tfo_programmable_bootstrap_ntt64_bnf(const tfo_plan64 *pl, const uint64_t *bsk, size_t n_lwe,
                                          size_t glwe_size, uint32_t base_log, uint32_t level,
                                          uint32_t width, const uint64_t *lwe_in, uint64_t *lwe_out,
                                          const uint64_t *accumulator) {
    size_t len = glwe_size * pl->n;
    unsigned log_modulus = log2_exact(pl->n) + 1; /* to_blind_rotation_input_modulus_log, parameters.rs:162 */
    uint64_t *local = (uint64_t *)malloc(len * sizeof(uint64_t));
    uint64_t *msed = (uint64_t *)malloc((n_lwe + 1) * sizeof(uint64_t));
    memcpy(local, accumulator, len * sizeof(uint64_t));
    /* lwe_ciphertext_modulus_switch (modulus_switch.rs:14-24): lazy, no body correction */
    for (size_t i = 0; i <= n_lwe; i++) msed[i] = tfo_modulus_switch(lwe_in[i], log_modulus);
    tfo_blind_rotate_ntt64_bnf_assign(pl, bsk, n_lwe, glwe_size, base_log, level, width, msed, local);
    tfo_extract_lwe_sample(local, glwe_size, pl->n, 0, 0, lwe_out);
    free(local);
    free(msed);
}

/* convert_standard_lwe_bootstrap_key_to_ntt64, lwe_bootstrap_key_conversion.rs:294-363:
 * input_width = 0 when the input key already lives modulo the NTT prime (ntt.forward), else the
 * log2 of its power-of-two modulus (forward_from_power_of_two_modulus); normalize = the
 * NttLweBootstrapKeyOption::Normalize option. */
void tfo_convert_standard_lwe_bootstrap_key_to_ntt64(const tfo_plan64 *pl, const uint64_t *input,
                                                     uint64_t *output, size_t poly_count,
                                                     uint32_t input_width, int normalize) {
    size_t n = pl->n;
    for (size_t i = 0; i < poly_count; i++) {
        tfo_ntt64_forward(pl, output + i * n, input + i * n, input_width ? 3 : 0, input_width);
        if (normalize) tfo_plan64_normalize(pl, output + i * n, n);
    }
}
