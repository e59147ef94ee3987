/*
 * tfhe_ntt_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the scalar code paths of the reference crate
 * `tfhe-ntt` (/root/reference/tfhe-ntt/src).  It exists to check the CUDA
 * engine; nothing under tfhe-rs-main_modified_b200/ may link, import or call
 * it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it.
 *
 * Parity pinning: the reference is Rust and cannot be built in this image
 * (no rustc/cargo), so there is no oracle/_ref.  The oracle is pinned against
 * every known-answer vector the reference's own tests hold for this path
 * (roots.rs:150-172, prime.rs:188-222, prime32.rs:1338-1397,
 * prime64.rs:1557-1569, :1988-1990, lib.rs:25-49) and against the schoolbook
 * negacyclic convolution the reference tests use as their own oracle
 * (prime64.rs:1264-1276) -- see tests/test_oracle_*.py.
 *
 * Every function cites the reference file:line it follows.
 */
#ifndef TFHE_NTT_ORACLE_H
#define TFHE_NTT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned __int128 tfo_u128;

/* ---- number theory (prime.rs, roots.rs, lib.rs) ---- */
uint64_t tfo_mul_mod64(uint64_t p, uint64_t a, uint64_t b);
uint32_t tfo_mul_mod32(uint32_t p, uint32_t a, uint32_t b);
uint64_t tfo_exp_mod64(uint64_t p, uint64_t base, uint64_t pow);
uint32_t tfo_exp_mod32(uint32_t p, uint32_t base, uint32_t pow);
int tfo_is_prime64(uint64_t n);
int tfo_largest_prime_in_arithmetic_progression64(uint64_t factor, uint64_t offset, uint64_t lo,
                                                  uint64_t hi, uint64_t *out);
int tfo_find_primitive_root64(uint64_t p, uint64_t degree, uint64_t *out);
int tfo_find_root_solinas_64(uint64_t n, uint64_t *out);
size_t tfo_bit_rev(uint32_t nbits, size_t i);

/* ---- prime64::Plan (prime64.rs:245-261) ---- */
typedef struct tfo_plan64 {
    size_t n;
    uint64_t p;
    uint64_t *twid, *twid_shoup, *inv_twid, *inv_twid_shoup; /* *_shoup NULL when p >= 2^63 */
    int use_ifma;                                            /* always 0: no IFMA path restated */
    int can_use_fast_reduction_code;
    uint64_t p_barrett, big_q;
    uint64_t n_inv_mod_p, n_inv_mod_p_shoup;
} tfo_plan64;

tfo_plan64 *tfo_plan64_try_new(size_t n, uint64_t p); /* NULL <=> None */
void tfo_plan64_free(tfo_plan64 *);
void tfo_plan64_fwd(const tfo_plan64 *, uint64_t *buf);
void tfo_plan64_inv(const tfo_plan64 *, uint64_t *buf);
void tfo_plan64_normalize(const tfo_plan64 *, uint64_t *values, size_t len);
void tfo_plan64_mul_assign_normalize(const tfo_plan64 *, uint64_t *lhs, const uint64_t *rhs,
                                     size_t len);
void tfo_plan64_mul_accumulate(const tfo_plan64 *, uint64_t *acc, const uint64_t *lhs,
                               const uint64_t *rhs, size_t len);
/* always the exact generic path (generic_solinas.rs:42-75 + :449-514), any p */
void tfo_plan64_fwd_generic(const tfo_plan64 *, uint64_t *buf);
void tfo_plan64_inv_generic(const tfo_plan64 *, uint64_t *buf);

/* ---- prime32::Plan (prime32.rs:632-648) ---- */
typedef struct tfo_plan32 {
    size_t n;
    uint32_t p;
    uint32_t *twid, *twid_shoup, *inv_twid, *inv_twid_shoup; /* *_shoup NULL when p >= 2^31 */
    int can_use_fast_reduction_code;
    uint32_t p_barrett, big_q;
    uint32_t n_inv_mod_p, n_inv_mod_p_shoup;
} tfo_plan32;

tfo_plan32 *tfo_plan32_try_new(size_t n, uint32_t p);
void tfo_plan32_free(tfo_plan32 *);
void tfo_plan32_fwd(const tfo_plan32 *, uint32_t *buf);
void tfo_plan32_inv(const tfo_plan32 *, uint32_t *buf);
void tfo_plan32_normalize(const tfo_plan32 *, uint32_t *values, size_t len);
void tfo_plan32_mul_assign_normalize(const tfo_plan32 *, uint32_t *lhs, const uint32_t *rhs,
                                     size_t len);
void tfo_plan32_mul_accumulate(const tfo_plan32 *, uint32_t *acc, const uint32_t *lhs,
                               const uint32_t *rhs, size_t len);
void tfo_plan32_fwd_generic(const tfo_plan32 *, uint32_t *buf);
void tfo_plan32_inv_generic(const tfo_plan32 *, uint32_t *buf);

/* ---- schoolbook negacyclic convolutions (the reference tests' own oracle) ---- */
void tfo_negacyclic_convolution_mod64(size_t n, uint64_t p, const uint64_t *lhs,
                                      const uint64_t *rhs, uint64_t *out); /* prime64.rs:1264-1276 */
void tfo_negacyclic_convolution_mod32(size_t n, uint32_t p, const uint32_t *lhs,
                                      const uint32_t *rhs, uint32_t *out); /* prime32.rs tests */
void tfo_negacyclic_convolution_wrapping_u32(size_t n, const uint32_t *lhs, const uint32_t *rhs,
                                             uint32_t *out);
void tfo_negacyclic_convolution_wrapping_u64(size_t n, const uint64_t *lhs, const uint64_t *rhs,
                                             uint64_t *out); /* native64.rs tests */
void tfo_negacyclic_convolution_wrapping_u128(size_t n, const tfo_u128 *lhs, const tfo_u128 *rhs,
                                              tfo_u128 *out); /* native128.rs:359-372 */

/* ---- CRT constants (lib.rs:451-656), recomputed, exposed for tests ---- */
uint32_t tfo_primes32(int i); /* P0..P9 */
uint64_t tfo_primes52(int i); /* P0..P5 */

/* ---- CRT plans ----
 * kind selects the reference plan type; residue buffers are an array of
 * `tfo_native_num_primes(kind)` pointers (u32 for *_PLAN32, u64 for *_PLAN52).
 * value buffers are u32 / u64 / u128 according to tfo_native_value_bytes(kind). */
enum tfo_native_kind {
    TFO_NATIVE32_PLAN32 = 0,        /* native32.rs:8-12    3 x prime32 */
    TFO_NATIVE32_PLAN52 = 1,        /* native32.rs:18      2 x prime64 (primes52) */
    TFO_NATIVE64_PLAN32 = 2,        /* native64.rs:16-22   5 x prime32 */
    TFO_NATIVE64_PLAN52 = 3,        /* native64.rs:28-33   3 x prime64 */
    TFO_NATIVE128_PLAN32 = 4,       /* native128.rs:6-17   10 x prime32 */
    TFO_NATIVE_BINARY32_PLAN32 = 5, /* native_binary32.rs:11   2 x prime32 */
    TFO_NATIVE_BINARY32_PLAN52 = 6, /* native_binary32.rs:18   1 x prime64 */
    TFO_NATIVE_BINARY64_PLAN32 = 7, /* native_binary64.rs:17-21 3 x prime32 */
    TFO_NATIVE_BINARY64_PLAN52 = 8, /* native_binary64.rs:28   2 x prime64 */
    TFO_NATIVE_BINARY128_PLAN32 = 9 /* native_binary128.rs:4-10 5 x prime32 */
};

typedef struct tfo_native_plan {
    int kind;
    size_t n;
    int num_primes;
    int residue_bytes; /* 4 or 8 */
    int value_bytes;   /* 4, 8 or 16 */
    tfo_plan32 *p32[10];
    tfo_plan64 *p64[3];
} tfo_native_plan;

int tfo_native_num_primes(int kind);
int tfo_native_residue_bytes(int kind);
int tfo_native_value_bytes(int kind);
tfo_native_plan *tfo_native_try_new(int kind, size_t n);
void tfo_native_free(tfo_native_plan *);
/* fwd: value -> residues (split + per-prime fwd).  binary != 0 selects fwd_binary. */
void tfo_native_fwd(const tfo_native_plan *, const void *value, void *const *residues, int binary);
/* inv: per-prime inv (clobbers residues) + CRT recombination into value */
void tfo_native_inv(const tfo_native_plan *, void *value, void *const *residues);
/* negacyclic_polymul; for the native_binary* kinds rhs is the binary operand */
void tfo_native_negacyclic_polymul(const tfo_native_plan *, void *prod, const void *lhs,
                                   const void *rhs);

/* scalar CRT recombinations, exposed for direct tests */
uint32_t tfo_reconstruct_32bit_012(uint32_t r0, uint32_t r1, uint32_t r2);            /* native32.rs:27-55 */
uint32_t tfo_reconstruct_52bit_01_u32(uint64_t r0, uint64_t r1);                      /* native32.rs:222-252 */
uint64_t tfo_reconstruct_32bit_01234_v2(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3,
                                        uint32_t r4);                                 /* native64.rs:90-140 */
uint64_t tfo_reconstruct_32bit_01234(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3,
                                     uint32_t r4);                                    /* native64.rs:44-87 */
uint64_t tfo_reconstruct_52bit_012(uint64_t r0, uint64_t r1, uint64_t r2);            /* native64.rs:769-828 */
tfo_u128 tfo_reconstruct_32bit_0123456789_v2(const uint32_t r[10]);                   /* native128.rs:20-118 */
uint32_t tfo_reconstruct_32bit_01(uint32_t r0, uint32_t r1);                          /* native_binary32.rs:21-40 */
uint32_t tfo_reconstruct_52bit_0_u32(uint64_t r0);                                    /* native_binary32.rs:110-123 */
uint64_t tfo_reconstruct_32bit_012_u64(uint32_t r0, uint32_t r1, uint32_t r2);        /* native_binary64.rs:32-60 */
uint64_t tfo_reconstruct_52bit_01_u64(uint64_t r0, uint64_t r1);                      /* native_binary64.rs:229-260 */
tfo_u128 tfo_reconstruct_32bit_01234_v2_u128(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3,
                                             uint32_t r4);                            /* native_binary128.rs:13-63 */

/* ---- batch helpers for the CPU baseline (bench.py only) ----
 * Transform `batch` contiguous polynomials with `threads` worker threads using
 * static contiguous chunks (the decomposition the reference's caller uses with
 * rayon par_chunks, tfhe/src/core_crypto/algorithms/lwe_bootstrap_key_conversion.rs:419-447). */
void tfo_plan64_fwd_batch(const tfo_plan64 *, uint64_t *buf, size_t batch, int threads);
void tfo_plan64_inv_batch(const tfo_plan64 *, uint64_t *buf, size_t batch, int threads);
void tfo_plan32_fwd_batch(const tfo_plan32 *, uint32_t *buf, size_t batch, int threads);
void tfo_plan32_inv_batch(const tfo_plan32 *, uint32_t *buf, size_t batch, int threads);

/* ---- product::Plan (product.rs:139-967): NTT over a product of distinct primes ----
 * NTT-domain layout of one polynomial (product.rs:261-283): n32 arrays of n u32 (packed two per
 * u64 word) followed by n64 arrays of n u64; ntt_domain_len() = n/2 * n32 + n * n64 words. */
typedef struct tfo_product_plan {
    size_t n;
    uint64_t modulus;
    int n32, n64;
    tfo_plan32 *p32[16];
    tfo_plan64 *p64[16];
    uint64_t primes[32];     /* sorted: the u32 primes, then the u64 primes */
    uint64_t inverses[512];  /* modular_inverses, product.rs:203-225 */
} tfo_product_plan;
tfo_product_plan *tfo_product_try_new(size_t n, uint64_t modulus, const uint64_t *factors,
                                      size_t nfactors); /* NULL <=> None */
void tfo_product_free(tfo_product_plan *);
size_t tfo_product_ntt_domain_len(const tfo_product_plan *);
/* bounded != 0 selects FwdMode::Bounded(bound) */
void tfo_product_fwd(const tfo_product_plan *, uint64_t *ntt, const uint64_t *standard, int bounded,
                     uint64_t bound);
/* accumulate != 0 selects InvMode::Accumulate; ntt is transformed in place like the reference */
void tfo_product_inv(const tfo_product_plan *, uint64_t *standard, uint64_t *ntt, int accumulate);
void tfo_product_mul_assign_normalize(const tfo_product_plan *, uint64_t *lhs, const uint64_t *rhs);
void tfo_product_normalize(const tfo_product_plan *, uint64_t *values);
void tfo_product_mul_accumulate(const tfo_product_plan *, uint64_t *acc, const uint64_t *lhs,
                                const uint64_t *rhs);

/* ---- tfhe Ntt64View (tfhe/src/core_crypto/commons/math/ntt/ntt64.rs:89-266) ----
 * forward modes: 0 forward, 1 forward_normalized, 2 forward_from_decomp,
 * 3 forward_from_power_of_two_modulus(width); add_backward modes: 0 add_backward,
 * 1 add_backward_on_power_of_two_modulus(width) */
void tfo_ntt64_forward(const tfo_plan64 *, uint64_t *ntt, const uint64_t *standard, int mode,
                       uint32_t width);
void tfo_ntt64_add_backward(const tfo_plan64 *, uint64_t *standard, uint64_t *ntt, int mode,
                            uint32_t width);

/* ---- NTT programmable bootstrap (tfhe_ntt_pbs_oracle.c): tfhe ntt64_pbs.rs / ntt64_bnf_pbs.rs ----
 * lwe [n_lwe+1]; glwe / lut / accumulator [(k+1)*N]; bsk [n_lwe][l][k+1][k+1][N] (NTT domain).
 * "modulus == 0" selects the native 2^64 wrapping arithmetic. */
uint64_t tfo_pbs_modulus_switch_non_native(uint64_t input, uint32_t log2_poly_size, uint64_t modulus);
uint64_t tfo_modulus_switch(uint64_t input, uint32_t log_modulus);
void tfo_monomial_mul_assign(uint64_t *poly, size_t n, size_t degree, uint64_t modulus);
void tfo_monomial_div_assign(uint64_t *poly, size_t n, size_t degree, uint64_t modulus);
uint64_t tfo_closest_representable_non_native(uint64_t input, uint32_t base_log, uint32_t level,
                                              uint64_t modulus);
uint64_t tfo_init_decomposer_state_native(uint64_t input, uint32_t base_log, uint32_t level);
void tfo_decomp_non_native_init(const uint64_t *input, size_t len, uint32_t base_log, uint32_t level,
                                uint64_t modulus, uint64_t *states, uint8_t *signs);
void tfo_decomp_non_native_next(uint64_t *states, const uint8_t *signs, size_t len, uint32_t base_log,
                                uint64_t modulus, uint64_t *term);
void tfo_decomp_native_init(const uint64_t *input, size_t len, uint32_t base_log, uint32_t level,
                            uint64_t *states);
void tfo_decomp_native_next(uint64_t *states, size_t len, uint32_t base_log, uint64_t *term);
void tfo_add_external_product_ntt64_assign(const tfo_plan64 *, uint64_t *out, const uint64_t *ggsw,
                                           const uint64_t *glwe, size_t glwe_size, uint32_t base_log,
                                           uint32_t level, int bnf, uint32_t width);
void tfo_cmux_ntt64_assign(const tfo_plan64 *, uint64_t *ct0, uint64_t *ct1, const uint64_t *ggsw,
                           size_t glwe_size, uint32_t base_log, uint32_t level, int bnf, uint32_t width);
void tfo_blind_rotate_ntt64_assign(const tfo_plan64 *, const uint64_t *bsk, size_t n_lwe,
                                   size_t glwe_size, uint32_t base_log, uint32_t level,
                                   const uint64_t *lwe, uint64_t *lut);
void tfo_blind_rotate_ntt64_bnf_assign(const tfo_plan64 *, const uint64_t *bsk, size_t n_lwe,
                                       size_t glwe_size, uint32_t base_log, uint32_t level,
                                       uint32_t width, const uint64_t *msed, uint64_t *lut);
void tfo_extract_lwe_sample(const uint64_t *glwe, size_t glwe_size, size_t n, size_t nth,
                            uint64_t modulus, uint64_t *lwe_out);
void tfo_programmable_bootstrap_ntt64(const tfo_plan64 *, const uint64_t *bsk, size_t n_lwe,
                                      size_t glwe_size, uint32_t base_log, uint32_t level,
                                      const uint64_t *lwe_in, uint64_t *lwe_out,
                                      const uint64_t *accumulator);
void tfo_programmable_bootstrap_ntt64_bnf(const tfo_plan64 *, const uint64_t *bsk, size_t n_lwe,
                                          size_t glwe_size, uint32_t base_log, uint32_t level,
                                          uint32_t width, const uint64_t *lwe_in, uint64_t *lwe_out,
                                          const uint64_t *accumulator);
void tfo_convert_standard_lwe_bootstrap_key_to_ntt64(const tfo_plan64 *, const uint64_t *input,
                                                     uint64_t *output, size_t poly_count,
                                                     uint32_t input_width, int normalize);

/* AVX-512 port of the reference's vectorised Solinas path (tfhe_ntt_simd.c; bench.py only).
 * Return 1 when the SIMD path ran, 0 when the CPU / build has no AVX-512F+DQ or p is not the
 * Solinas prime (the caller then uses the scalar batch helpers). */
int tfo_plan64_fwd_batch_simd(const tfo_plan64 *, uint64_t *buf, size_t batch, int threads);
/* one polynomial on the calling thread; 0 = no vector path on this host / for this plan (nothing done) */
int tfo_plan64_fwd_simd1(const tfo_plan64 *, uint64_t *buf);
int tfo_plan64_inv_simd1(const tfo_plan64 *, uint64_t *buf);
/* CPU BASELINE ONLY (bench.py): when on, tfo_ntt64_forward / tfo_ntt64_add_backward -- the transforms
 * inside the PBS restatement -- run the vectorised Solinas port where the host has AVX-512 (same bits
 * as the scalar path, tests/test_oracle_simd.py).  Off by default: the parity tests use the scalar path. */
void tfo_use_simd_transforms(int on);
/* batch of programmable bootstraps (classic) over `threads` host threads in static contiguous chunks,
 * the decomposition the reference's callers use with rayon; accumulator [acc_count][(k+1)N], acc_count 1 or batch */
void tfo_programmable_bootstrap_ntt64_batch(const tfo_plan64 *, const uint64_t *bsk, size_t n_lwe,
                                            size_t glwe_size, uint32_t base_log, uint32_t level,
                                            const uint64_t *lwe_in, uint64_t *lwe_out,
                                            const uint64_t *accumulator, size_t acc_count, size_t batch,
                                            int threads);
int tfo_plan64_inv_batch_simd(const tfo_plan64 *, uint64_t *buf, size_t batch, int threads);

/* custum_radix (tfhe_ntt_custum_radix_oracle.c): the fork's recursive cyclic u32 transforms,
 * tfhe-ntt/src/custum_radix/{fwd.rs,inv.rs,fwd_1.rs}.  tw[k] = root^k, natural order in and out. */
typedef struct {
    size_t nonzero_mults, skipped_mults; /* fwd_1.rs:3-7 MultStats */
} tfo_mult_stats;
uint32_t tfo_cr_pow_mod(uint32_t base, uint32_t exp, uint32_t p);
uint32_t tfo_cr_mod_inverse(uint32_t a, uint32_t p);
uint32_t tfo_cr_compute_primitive_root(uint32_t p);
int tfo_cr_make_twiddles(size_t n, uint32_t p, uint32_t *tw);
void tfo_cr_make_inv_twiddles(const uint32_t *tw, size_t n, uint32_t p, uint32_t *inv);
void tfo_cr_fft_radix4_recursive(uint32_t *a, size_t n, const uint32_t *tw, uint32_t p);
void tfo_cr_fft_radix2_recursive(uint32_t *a, size_t n, const uint32_t *tw, uint32_t p);
void tfo_cr_fft_split_radix_recursive(uint32_t *a, size_t n, const uint32_t *tw, uint32_t p);
void tfo_cr_ifft_radix4_recursive(uint32_t *a, size_t n, const uint32_t *inv_tw, uint32_t p, uint32_t n_inv, int top);
void tfo_cr_ifft_radix2_recursive(uint32_t *a, size_t n, const uint32_t *inv_tw, uint32_t p, uint32_t n_inv, int top);
void tfo_cr_ifft_split_radix_recursive(uint32_t *a, size_t n, const uint32_t *inv_tw, uint32_t p, uint32_t n_inv,
                                       int top);
void tfo_cr_fft_radix4_recursive_mut(uint32_t *a, size_t n, const uint32_t *tw, uint32_t p, tfo_mult_stats *st);
void tfo_cr_fft_radix2_recursive_mut(uint32_t *a, size_t n, const uint32_t *tw, uint32_t p, tfo_mult_stats *st);
void tfo_cr_fft_split_radix_recursive_mut(uint32_t *a, size_t n, const uint32_t *tw, uint32_t p, tfo_mult_stats *st);
void tfo_cr_ifft_radix4_recursive_mut(uint32_t *a, size_t n, const uint32_t *inv_tw, uint32_t p, uint32_t n_inv,
                                      int top, tfo_mult_stats *st);

#ifdef __cplusplus
}
#endif
#endif
