/*
 * tfhe_ntt_b200.h -- C ABI of the B200-native batched NTT engine (libtfhe_ntt_b200.so).
 *
 * The reference (`tfhe-ntt`, /root/reference/tfhe-ntt/src) is a plain Rust crate with no FFI
 * layer of its own; its public Rust API is the drop-in boundary (lib.rs:83-116).  Each entry
 * point below is what a Rust host crate binds to stand in for one reference method; the
 * citation names the method it replaces.  INTEGRATION.md shows the Rust `extern "C"` block and
 * the safe wrappers (`prime64::Plan::fwd` ...) a maintainer adds on the reference side.
 *
 * Conventions
 *   - Plain pointers and sizes only.  Return value: NTT_B200_OK (0) on success,
 *     NTT_B200_NONE where the reference returns `None`, NTT_B200_ERR_LEN where the reference
 *     panics on a length assertion, NTT_B200_ERR_CUDA on a CUDA failure (text from
 *     ntt_b200_last_error()).
 *   - `*_fwd`, `*_inv`, ... with HOST pointers are the per-polynomial drop-ins.
 *     `*_batch` take HOST pointers to `batch` contiguous polynomials (new; the reference is
 *     per-polynomial, prime64.rs:897-898).  `*_device` take DEVICE pointers on the plan's GPU
 *     and a CUDA stream (`void *`, a cudaStream_t; NULL = default stream) and do not synchronise.
 *   - A plan lives on the CUDA device that was current when it was created; tables are
 *     immutable afterwards, all entry points are re-entrant on a shared plan
 *     (reference: plans are Send + Sync, shared through Arc, tfhe ntt64.rs:26-31).
 *   - Outputs are bit-identical to the reference for inputs in [0, p) (prime plans) / any
 *     value (native plans).  Inputs >= p are outside the contract, as in the reference.
 */
#ifndef TFHE_NTT_B200_H
#define TFHE_NTT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NTT_B200_OK 0
#define NTT_B200_NONE 1      /* constructor: the reference returns None */
#define NTT_B200_ERR_LEN 2   /* the reference's assert_eq!(buf.len(), ntt_size()) would panic */
#define NTT_B200_ERR_CUDA 3  /* CUDA runtime failure; see ntt_b200_last_error() */
#define NTT_B200_ERR_ARG 4   /* null pointer / unknown plan kind */
#define NTT_B200_ERR_UNSUPPORTED 5 /* a valid call beyond a capacity limit of this implementation (custum_radix *_mut: n > 4096) */

const char *ntt_b200_last_error(void); /* thread-local text of the last NTT_B200_ERR_CUDA */
int ntt_b200_device_count(void);
int ntt_b200_set_device(int device);   /* device used by subsequently created plans */

/* ------------------------------------------------------------------------------------------
 * prime64::Plan   (tfhe-ntt/src/prime64.rs:245-261)
 * ------------------------------------------------------------------------------------------ */
typedef struct ntt_b200_plan64 ntt_b200_plan64;

/* prime64::Plan::try_new(polynomial_size, modulus) -> Option<Plan>   prime64.rs:764-862
 * NTT_B200_NONE iff n < 16, n not a power of two, p composite, or 2n does not divide p-1. */
int ntt_b200_plan64_try_new(size_t n, uint64_t p, ntt_b200_plan64 **out);
/* impl Clone for Plan                                                prime64.rs:244 */
int ntt_b200_plan64_clone(const ntt_b200_plan64 *plan, ntt_b200_plan64 **out);
/* Drop */
void ntt_b200_plan64_free(ntt_b200_plan64 *plan);
/* Plan::ntt_size / modulus / use_ifma / can_use_fast_reduction_code  prime64.rs:870-890
 * use_ifma is always 0 (there is no CPU IFMA path); can_use_fast_reduction_code reproduces
 * prime64.rs:815-817 with use_ifma = false. */
size_t ntt_b200_plan64_ntt_size(const ntt_b200_plan64 *plan);
uint64_t ntt_b200_plan64_modulus(const ntt_b200_plan64 *plan);
int ntt_b200_plan64_use_ifma(const ntt_b200_plan64 *plan);
int ntt_b200_plan64_can_use_fast_reduction_code(const ntt_b200_plan64 *plan);
int ntt_b200_plan64_device(const ntt_b200_plan64 *plan);

/* Plan::fwd(&self, buf: &mut [u64])   prime64.rs:897-968 : natural -> bit-reversed, in [0,p) */
int ntt_b200_plan64_fwd(const ntt_b200_plan64 *plan, uint64_t *buf, size_t len);
/* Plan::inv(&self, buf: &mut [u64])   prime64.rs:975-1046: bit-reversed -> natural, times n */
int ntt_b200_plan64_inv(const ntt_b200_plan64 *plan, uint64_t *buf, size_t len);
/* Plan::normalize(&self, values)      prime64.rs:1137-1179 */
int ntt_b200_plan64_normalize(const ntt_b200_plan64 *plan, uint64_t *values, size_t len);
/* Plan::mul_assign_normalize(lhs, rhs) prime64.rs:1050-1133; like the reference's izip!
 * (lib.rs:658-688) the shorter of the two slices bounds the work */
int ntt_b200_plan64_mul_assign_normalize(const ntt_b200_plan64 *plan, uint64_t *lhs,
                                         size_t lhs_len, const uint64_t *rhs, size_t rhs_len);
/* Plan::mul_accumulate(acc, lhs, rhs) prime64.rs:1182-1222.  Always the canonical (acc + lhs*rhs) mod p; the
 * reference's one-step Barrett code (prime64.rs:586-609) returns that value + p for about one product in 10^5 on
 * moduli below 2^64/3 that need two correction steps -- the only known divergence. */
int ntt_b200_plan64_mul_accumulate(const ntt_b200_plan64 *plan, uint64_t *acc, size_t acc_len,
                                   const uint64_t *lhs, size_t lhs_len, const uint64_t *rhs,
                                   size_t rhs_len);

/* batched, host memory: `batch` polynomials of ntt_size() coefficients each, contiguous */
int ntt_b200_plan64_fwd_batch(const ntt_b200_plan64 *plan, uint64_t *host, size_t batch);
int ntt_b200_plan64_inv_batch(const ntt_b200_plan64 *plan, uint64_t *host, size_t batch);
/* NEW (no reference counterpart: tfhe-ntt is single-device CPU code): the same host batch split over
 * several GPUs of one box.  plans[g] are plans for the same (n, p), each created on its own GPU
 * (ntt_b200_set_device(g) before try_new); GPU g gets a contiguous slice, batch / n_plans polynomials
 * with the remainder one by one to the first GPUs -- the split rule of the reference's CUDA backend,
 * backends/tfhe-cuda-backend/cuda/src/utils/helper_multi_gpu.cu:57-88 -- one host thread per GPU, no
 * exchange between GPUs.  ERR_ARG if the plans disagree on n or p. */
int ntt_b200_plan64_fwd_batch_multi_gpu(const ntt_b200_plan64 *const *plans, size_t n_plans,
                                        uint64_t *host, size_t batch);
int ntt_b200_plan64_inv_batch_multi_gpu(const ntt_b200_plan64 *const *plans, size_t n_plans,
                                        uint64_t *host, size_t batch);

/* device-resident: pointers on the plan's GPU, asynchronous on `stream` */
int ntt_b200_plan64_fwd_device(const ntt_b200_plan64 *plan, uint64_t *dev, size_t batch,
                               void *stream);
int ntt_b200_plan64_inv_device(const ntt_b200_plan64 *plan, uint64_t *dev, size_t batch,
                               void *stream);
int ntt_b200_plan64_normalize_device(const ntt_b200_plan64 *plan, uint64_t *dev, size_t len,
                                     void *stream);
/* rhs (and lhs for mul_accumulate) may be shorter than the destination: it is then reused
 * cyclically (len must be a multiple of its length) -- e.g. one GGSW row shared by the batch. */
int ntt_b200_plan64_mul_assign_normalize_device(const ntt_b200_plan64 *plan, uint64_t *lhs,
                                                size_t len, const uint64_t *rhs, size_t rhs_len,
                                                void *stream);
int ntt_b200_plan64_mul_accumulate_device(const ntt_b200_plan64 *plan, uint64_t *acc, size_t len,
                                          const uint64_t *lhs, size_t lhs_len,
                                          const uint64_t *rhs, size_t rhs_len, void *stream);

/* ------------------------------------------------------------------------------------------
 * prime32::Plan   (tfhe-ntt/src/prime32.rs:632-648); same shape with u32 and n >= 32
 * ------------------------------------------------------------------------------------------ */
typedef struct ntt_b200_plan32 ntt_b200_plan32;

int ntt_b200_plan32_try_new(size_t n, uint32_t p, ntt_b200_plan32 **out); /* prime32.rs:662-765 */
int ntt_b200_plan32_clone(const ntt_b200_plan32 *plan, ntt_b200_plan32 **out);
void ntt_b200_plan32_free(ntt_b200_plan32 *plan);
size_t ntt_b200_plan32_ntt_size(const ntt_b200_plan32 *plan);                 /* prime32.rs:773 */
uint32_t ntt_b200_plan32_modulus(const ntt_b200_plan32 *plan);                /* prime32.rs:779 */
int ntt_b200_plan32_can_use_fast_reduction_code(const ntt_b200_plan32 *plan); /* prime32.rs:785 */
int ntt_b200_plan32_device(const ntt_b200_plan32 *plan);

int ntt_b200_plan32_fwd(const ntt_b200_plan32 *plan, uint32_t *buf, size_t len); /* :797-843 */
int ntt_b200_plan32_inv(const ntt_b200_plan32 *plan, uint32_t *buf, size_t len); /* :850-896 */
int ntt_b200_plan32_normalize(const ntt_b200_plan32 *plan, uint32_t *values, size_t len); /* :956 */
int ntt_b200_plan32_mul_assign_normalize(const ntt_b200_plan32 *plan, uint32_t *lhs,
                                         size_t lhs_len, const uint32_t *rhs,
                                         size_t rhs_len); /* :900-952 */
int ntt_b200_plan32_mul_accumulate(const ntt_b200_plan32 *plan, uint32_t *acc, size_t acc_len,
                                   const uint32_t *lhs, size_t lhs_len, const uint32_t *rhs,
                                   size_t rhs_len); /* :993-1015 */
int ntt_b200_plan32_fwd_batch(const ntt_b200_plan32 *plan, uint32_t *host, size_t batch);
int ntt_b200_plan32_inv_batch(const ntt_b200_plan32 *plan, uint32_t *host, size_t batch);
int ntt_b200_plan32_fwd_batch_multi_gpu(const ntt_b200_plan32 *const *plans, size_t n_plans,
                                        uint32_t *host, size_t batch);
int ntt_b200_plan32_inv_batch_multi_gpu(const ntt_b200_plan32 *const *plans, size_t n_plans,
                                        uint32_t *host, size_t batch);
int ntt_b200_plan32_fwd_device(const ntt_b200_plan32 *plan, uint32_t *dev, size_t batch,
                               void *stream);
int ntt_b200_plan32_inv_device(const ntt_b200_plan32 *plan, uint32_t *dev, size_t batch,
                               void *stream);
int ntt_b200_plan32_normalize_device(const ntt_b200_plan32 *plan, uint32_t *dev, size_t len,
                                     void *stream);
int ntt_b200_plan32_mul_assign_normalize_device(const ntt_b200_plan32 *plan, uint32_t *lhs,
                                                size_t len, const uint32_t *rhs, size_t rhs_len,
                                                void *stream);
int ntt_b200_plan32_mul_accumulate_device(const ntt_b200_plan32 *plan, uint32_t *acc, size_t len,
                                          const uint32_t *lhs, size_t lhs_len,
                                          const uint32_t *rhs, size_t rhs_len, void *stream);

/* Fused fwd -> mul_accumulate -> inv on device (BASELINE config C2):
 *   out[b] = inv(acc[b] + fwd(lhs[b]) (*) rhs[b])   with rhs, acc already in the NTT domain.
 * rhs_polys / acc_polys < batch reuse those operands cyclically (0 acc_polys = no accumulator).
 * Equivalent to Plan::fwd + Plan::mul_accumulate + Plan::inv called in sequence. */
int ntt_b200_plan32_fwd_mac_inv_device(const ntt_b200_plan32 *plan, uint32_t *out,
                                       const uint32_t *lhs, const uint32_t *rhs, size_t rhs_polys,
                                       const uint32_t *acc, size_t acc_polys, size_t batch,
                                       void *stream);
int ntt_b200_plan64_fwd_mac_inv_device(const ntt_b200_plan64 *plan, uint64_t *out,
                                       const uint64_t *lhs, const uint64_t *rhs, size_t rhs_polys,
                                       const uint64_t *acc, size_t acc_polys, size_t batch,
                                       void *stream);

/* External-product core of the NTT-PBS (tfhe ntt64_pbs.rs:598-661), new:
 *   out[b][c] = inv( sum_{r<rows} fwd(in[b][r]) (*) ggsw[r][c] )      c < cols
 * in: batch x rows polynomials (e.g. the (k+1)*l decomposition digits of one GLWE per batch item),
 * ggsw: rows x cols NTT-domain polynomials shared by the whole batch, out: batch x cols polynomials.
 * Same result as Plan::fwd on every input, Plan::mul_accumulate for every (r, c), Plan::inv on
 * every accumulator.  rows >= 1, 1 <= cols <= 4 use one fused kernel for n = 512..4096. */
int ntt_b200_plan64_ext_product_device(const ntt_b200_plan64 *plan, uint64_t *out,
                                       const uint64_t *in, const uint64_t *ggsw, size_t rows,
                                       size_t cols, size_t batch, void *stream);
int ntt_b200_plan32_ext_product_device(const ntt_b200_plan32 *plan, uint32_t *out,
                                       const uint32_t *in, const uint32_t *ggsw, size_t rows,
                                       size_t cols, size_t batch, void *stream);

/* Host-memory form of the fused call: lhs/out hold `batch` polynomials in host memory (pinned
 * memory lets the copies overlap the kernels); rhs / acc hold rhs_polys / acc_polys polynomials
 * (== batch, or fewer and reused cyclically; acc may be NULL).  H2D, kernel and D2H of successive
 * chunks (32 MiB) are pipelined on four streams.  out may alias lhs. */
int ntt_b200_plan32_fwd_mac_inv_batch(const ntt_b200_plan32 *plan, uint32_t *out,
                                      const uint32_t *lhs, const uint32_t *rhs, size_t rhs_polys,
                                      const uint32_t *acc, size_t acc_polys, size_t batch);
int ntt_b200_plan64_fwd_mac_inv_batch(const ntt_b200_plan64 *plan, uint64_t *out,
                                      const uint64_t *lhs, const uint64_t *rhs, size_t rhs_polys,
                                      const uint64_t *acc, size_t acc_polys, size_t batch);

/* ------------------------------------------------------------------------------------------
 * CRT ("native") plans
 *   native32::{Plan32,Plan52}          native32.rs:8-18,  :334-498
 *   native64::{Plan32,Plan52}          native64.rs:16-33, :929-1164
 *   native128::Plan32                  native128.rs:6-17, :120-349
 *   native_binary32::{Plan32,Plan52}   native_binary32.rs:11-18, :186-310
 *   native_binary64::{Plan32,Plan52}   native_binary64.rs:17-28, :341-515
 *   native_binary128::Plan32           native_binary128.rs:4-10, :65-200
 * One handle type; `kind` selects the reference type.  value elements are u32 / u64 / u128
 * (16 bytes, little endian) and residues are u32 (Plan32) or u64 (Plan52):
 * see ntt_b200_native_value_bytes / _residue_bytes / _num_primes.
 * Plan52::try_new in the reference also requires an AVX512-IFMA CPU (native64.rs:1079);
 * the GPU engine always provides the 52-bit plans.
 * ------------------------------------------------------------------------------------------ */
enum ntt_b200_native_kind {
    NTT_B200_NATIVE32_PLAN32 = 0,
    NTT_B200_NATIVE32_PLAN52 = 1,
    NTT_B200_NATIVE64_PLAN32 = 2,
    NTT_B200_NATIVE64_PLAN52 = 3,
    NTT_B200_NATIVE128_PLAN32 = 4,
    NTT_B200_NATIVE_BINARY32_PLAN32 = 5,
    NTT_B200_NATIVE_BINARY32_PLAN52 = 6,
    NTT_B200_NATIVE_BINARY64_PLAN32 = 7,
    NTT_B200_NATIVE_BINARY64_PLAN52 = 8,
    NTT_B200_NATIVE_BINARY128_PLAN32 = 9
};
typedef struct ntt_b200_native_plan ntt_b200_native_plan;

int ntt_b200_native_num_primes(int kind);
int ntt_b200_native_residue_bytes(int kind);
int ntt_b200_native_value_bytes(int kind);

/* PlanXX::try_new(n) -> Option<Self>   e.g. native64.rs:932-941 */
int ntt_b200_native_try_new(int kind, size_t n, ntt_b200_native_plan **out);
void ntt_b200_native_free(ntt_b200_native_plan *plan);
size_t ntt_b200_native_ntt_size(const ntt_b200_native_plan *plan); /* e.g. native64.rs:945 */
int ntt_b200_native_kind_of(const ntt_b200_native_plan *plan);
/* PlanXX::ntt_0() .. ntt_9(): the per-prime plan (borrowed; owned by the native plan).
 * Returns ntt_b200_plan32* for Plan32 kinds, ntt_b200_plan64* for Plan52 kinds. */
const void *ntt_b200_native_ntt_i(const ntt_b200_native_plan *plan, int i); /* native64.rs:950-968 */

/* PlanXX::fwd(value, mod_p0, ..)        e.g. native64.rs:970-998; binary != 0 selects
 * fwd_binary (native_binary64.rs:371-388).  residues[i] points at `len` elements. */
int ntt_b200_native_fwd(const ntt_b200_native_plan *plan, const void *value, size_t len,
                        void *const *residues, int binary);
/* PlanXX::inv(value, mod_p0, ..)        e.g. native64.rs:1000-1037.  The residue buffers are
 * transformed in place (clobbered), as in the reference. */
int ntt_b200_native_inv(const ntt_b200_native_plan *plan, void *value, size_t len,
                        void *const *residues);
/* PlanXX::negacyclic_polymul(prod, lhs, rhs)  e.g. native64.rs:1041-1068; the three slices must
 * have equal length n (NTT_B200_ERR_LEN otherwise, where the reference asserts). */
int ntt_b200_native_negacyclic_polymul(const ntt_b200_native_plan *plan, void *prod,
                                       size_t prod_len, const void *lhs, size_t lhs_len,
                                       const void *rhs, size_t rhs_len);
/* batched host / device-resident forms: `batch` contiguous polynomials per operand */
int ntt_b200_native_negacyclic_polymul_batch(const ntt_b200_native_plan *plan, void *prod,
                                             const void *lhs, const void *rhs, size_t batch);
int ntt_b200_native_negacyclic_polymul_device(const ntt_b200_native_plan *plan, void *prod,
                                              const void *lhs, const void *rhs, size_t batch,
                                              void *stream);
int ntt_b200_native_fwd_device(const ntt_b200_native_plan *plan, const void *value,
                               void *const *residues, size_t batch, int binary, void *stream);
int ntt_b200_native_inv_device(const ntt_b200_native_plan *plan, void *value,
                               void *const *residues, size_t batch, void *stream);

/* ------------------------------------------------------------------------------------------
 * tfhe's Ntt64View (tfhe/src/core_crypto/commons/math/ntt/ntt64.rs:89-266): the wrapper through
 * which the NTT-PBS calls prime64::Plan.  `len` = batch * ntt_size() coefficients.
 *   forward mode 0 forward (:89-95)            1 forward_normalized (:97-108)
 *                2 forward_from_decomp (:218-238)
 *                3 forward_from_power_of_two_modulus(width) (:201-214)
 *   add_backward mode 0 add_backward (:110-131: standard += inv(ntt) modulo p)
 *                     1 add_backward_on_power_of_two_modulus(width) (:242-266)
 * `ntt` is left as the reference leaves it (inverse-transformed, and modswitched in mode 1).
 * ------------------------------------------------------------------------------------------ */
int ntt_b200_ntt64_forward(const ntt_b200_plan64 *plan, uint64_t *ntt, const uint64_t *standard,
                           size_t len, int mode, uint32_t width);
int ntt_b200_ntt64_add_backward(const ntt_b200_plan64 *plan, uint64_t *standard, uint64_t *ntt,
                                size_t len, int mode, uint32_t width);
int ntt_b200_ntt64_forward_device(const ntt_b200_plan64 *plan, uint64_t *ntt,
                                  const uint64_t *standard, size_t batch, int mode, uint32_t width,
                                  void *stream);
int ntt_b200_ntt64_add_backward_device(const ntt_b200_plan64 *plan, uint64_t *standard,
                                       uint64_t *ntt, size_t batch, int mode, uint32_t width,
                                       void *stream);

/* ------------------------------------------------------------------------------------------
 * NTT programmable bootstrap, the caller of the hot path:
 *   tfhe/src/core_crypto/algorithms/lwe_programmable_bootstrapping/ntt64_pbs.rs       "classic":
 *       the ciphertext modulus is the NTT prime (SignedDecomposerNonNative, custom-mod rotations)
 *   tfhe/src/core_crypto/algorithms/lwe_programmable_bootstrapping/ntt64_bnf_pbs.rs   "bnf":
 *       ciphertexts live modulo 2^width (MSB aligned), modswitched around each NTT
 * Containers are the reference's flat ones: lwe [n_lwe+1] (mask, body), glwe / lut / accumulator
 * [(k+1)*N], NTT bootstrap key [n_lwe][level][k+1][k+1][N] with the first level slice = level l
 * (entities/ntt_ggsw_ciphertext.rs:176-190).  New: every call takes `batch` ciphertexts.
 * `path`: 0 = automatic (two-CTA cluster kernel when k = 1, level = 1 and N = 512..4096; else
 * the one-CTA fused kernel when the shape has one; else composed);
 * 1 = one-CTA fused only, 3 = cluster only (ERR_CUDA when there is none); 2 = composed only.
 * All paths give identical bits.
 * ------------------------------------------------------------------------------------------ */
typedef struct ntt_b200_bsk ntt_b200_bsk;
/* NttLweBootstrapKey::from_container   entities/ntt_lwe_bootstrap_key.rs:68-110; copies the
 * NTT-domain container to the plan's device.  base_log * level must be < 64. */
int ntt_b200_bsk_new(const ntt_b200_plan64 *plan, const uint64_t *ntt_bsk, size_t n_lwe,
                     size_t glwe_size, uint32_t base_log, uint32_t level, ntt_b200_bsk **out);
/* convert_standard_lwe_bootstrap_key_to_ntt64   lwe_bootstrap_key_conversion.rs:294-363, into a
 * device-resident key.  input_width = 0: the standard key is modulo the NTT prime (ntt.forward);
 * else log2 of its power-of-two modulus (forward_from_power_of_two_modulus).  normalize != 0 =
 * NttLweBootstrapKeyOption::Normalize. */
int ntt_b200_bsk_convert_new(const ntt_b200_plan64 *plan, const uint64_t *standard_bsk, size_t n_lwe,
                             size_t glwe_size, uint32_t base_log, uint32_t level,
                             uint32_t input_width, int normalize, ntt_b200_bsk **out);
/* the same conversion host to host (`len` coefficients, a multiple of ntt_size()) */
int ntt_b200_convert_standard_lwe_bootstrap_key_to_ntt64(const ntt_b200_plan64 *plan,
                                                         const uint64_t *input, uint64_t *output,
                                                         size_t len, uint32_t input_width,
                                                         int normalize);
void ntt_b200_bsk_free(ntt_b200_bsk *key);
size_t ntt_b200_bsk_input_lwe_dimension(const ntt_b200_bsk *key);         /* ntt_lwe_bootstrap_key.rs:113 */
size_t ntt_b200_bsk_glwe_size(const ntt_b200_bsk *key);                   /* :123 */
size_t ntt_b200_bsk_polynomial_size(const ntt_b200_bsk *key);             /* :118 */
uint32_t ntt_b200_bsk_decomposition_base_log(const ntt_b200_bsk *key);    /* :128 */
uint32_t ntt_b200_bsk_decomposition_level_count(const ntt_b200_bsk *key); /* :133 */
const uint64_t *ntt_b200_bsk_device_data(const ntt_b200_bsk *key);        /* device pointer */
/* copies the NTT-domain key back (len = n_lwe * level * (k+1)^2 * N) */
int ntt_b200_bsk_read(const ntt_b200_bsk *key, uint64_t *out, size_t len);
/* blind_rotate_ntt64_assign(input, lut, bsk)   ntt64_pbs.rs:175-286
 * lwe [batch][n_lwe+1]; lut [batch][(k+1)N] is rotated in place */
int ntt_b200_blind_rotate_ntt64_assign(const ntt_b200_bsk *key, const uint64_t *lwe, uint64_t *lut,
                                       size_t batch, int path);
/* blind_rotate_ntt64_bnf_assign(msed_input, lut, bsk)   ntt64_bnf_pbs.rs:174-276
 * msed [batch][n_lwe+1]: the modulus-switched mask and body (each in [0, 2N)) */
int ntt_b200_blind_rotate_ntt64_bnf_assign(const ntt_b200_bsk *key, uint32_t width,
                                           const uint64_t *msed, uint64_t *lut, size_t batch,
                                           int path);
/* programmable_bootstrap_ntt64_lwe_ciphertext(input, output, accumulator, bsk)  ntt64_pbs.rs:439-538
 * lwe_in [batch][n_lwe+1], lwe_out [batch][k*N+1], accumulator [acc_count][(k+1)N] with
 * acc_count = 1 (one LUT for the whole batch) or batch (ERR_LEN otherwise) */
int ntt_b200_programmable_bootstrap_ntt64(const ntt_b200_bsk *key, const uint64_t *lwe_in,
                                          uint64_t *lwe_out, const uint64_t *accumulator,
                                          size_t acc_count, size_t batch, int path);
/* programmable_bootstrap_ntt64_bnf_lwe_ciphertext   ntt64_bnf_pbs.rs:428-539 */
int ntt_b200_programmable_bootstrap_ntt64_bnf(const ntt_b200_bsk *key, uint32_t width,
                                              const uint64_t *lwe_in, uint64_t *lwe_out,
                                              const uint64_t *accumulator, size_t acc_count,
                                              size_t batch, int path);
/* device-resident forms (pointers on the key's device, asynchronous on `stream`).
 * lwe_is_switched: bnf only, `lwe` already holds modulus-switched values.
 * acc_out [batch][(k+1)N] must not alias lut. */
int ntt_b200_blind_rotate_ntt64_device(const ntt_b200_bsk *key, int bnf, uint32_t width,
                                       const uint64_t *lwe, int lwe_is_switched, const uint64_t *lut,
                                       size_t lut_count, uint64_t *acc_out, size_t batch, int path,
                                       void *stream);
/* extract_lwe_sample_from_glwe_ciphertext(glwe, lwe, MonomialDegree(0))
 * glwe_sample_extraction.rs:89-164; glwe [batch][(k+1)N] -> lwe_out [batch][k*N+1] */
int ntt_b200_extract_lwe_sample_device(const ntt_b200_bsk *key, int bnf, const uint64_t *glwe,
                                       uint64_t *lwe_out, size_t batch, void *stream);

/* ------------------------------------------------------------------------------------------
 * product::Plan   (tfhe-ntt/src/product.rs:139-967): negacyclic NTT modulo a product of distinct
 * primes (each < 2^32 prime runs as a prime32 plan, the others as prime64 plans).
 * NTT-domain layout of ONE polynomial = the reference's (product.rs:261-283): the u32 residue
 * arrays (n/2 words each) followed by the u64 residue arrays; ntt_domain_len() words in total.
 * The *_device forms take `batch` polynomials with the residue arrays prime-major (prime j owns
 * batch*n contiguous residues); batch = 1 is the reference layout.
 * ------------------------------------------------------------------------------------------ */
typedef struct ntt_b200_product_plan ntt_b200_product_plan;
/* Plan::try_new(polynomial_size, modulus, factors) -> Option<Plan>     product.rs:153-246
 * NONE: odd size, a zero / duplicate factor, product(factors) != modulus (or overflow), or any
 * per-prime try_new is None. */
int ntt_b200_product_try_new(size_t n, uint64_t modulus, const uint64_t *factors, size_t nfactors,
                             ntt_b200_product_plan **out);
void ntt_b200_product_free(ntt_b200_product_plan *plan);
size_t ntt_b200_product_ntt_size(const ntt_b200_product_plan *plan);       /* product.rs:251 */
uint64_t ntt_b200_product_modulus(const ntt_b200_product_plan *plan);      /* product.rs:257 */
size_t ntt_b200_product_ntt_domain_len(const ntt_b200_product_plan *plan); /* product.rs:268 */
/* Plan::fwd(ntt, standard, mode)   product.rs:273-357.  FwdMode::Bounded(b) produces the residues
 * of FwdMode::Generic for every input that honours the bound, so there is one exact entry point. */
int ntt_b200_product_fwd(const ntt_b200_product_plan *plan, uint64_t *ntt, size_t ntt_len,
                         const uint64_t *standard, size_t standard_len);
/* Plan::inv(standard, ntt, mode)   product.rs:360-880; accumulate != 0 = InvMode::Accumulate
 * (standard += value, modulo the product).  ntt is transformed in place, as in the reference. */
int ntt_b200_product_inv(const ntt_b200_product_plan *plan, uint64_t *standard, size_t standard_len,
                         uint64_t *ntt, size_t ntt_len, int accumulate);
int ntt_b200_product_normalize(const ntt_b200_product_plan *plan, uint64_t *values,
                               size_t len); /* product.rs:917-932 */
int ntt_b200_product_mul_assign_normalize(const ntt_b200_product_plan *plan, uint64_t *lhs,
                                          size_t lhs_len, const uint64_t *rhs,
                                          size_t rhs_len); /* product.rs:885-913 */
int ntt_b200_product_mul_accumulate(const ntt_b200_product_plan *plan, uint64_t *acc, size_t acc_len,
                                    const uint64_t *lhs, size_t lhs_len, const uint64_t *rhs,
                                    size_t rhs_len); /* product.rs:935-967 */
int ntt_b200_product_fwd_device(const ntt_b200_product_plan *plan, uint64_t *ntt,
                                const uint64_t *standard, size_t batch, void *stream);
int ntt_b200_product_inv_device(const ntt_b200_product_plan *plan, uint64_t *standard, uint64_t *ntt,
                                size_t batch, int accumulate, void *stream);

/* ------------------------------------------------------------------------------------------
 * plan-build helpers that are public in the reference crate
 * ------------------------------------------------------------------------------------------ */
/* prime::is_prime64                                    prime.rs:76-126 */
int ntt_b200_is_prime64(uint64_t n);
/* prime::largest_prime_in_arithmetic_progression64     prime.rs:130-186 ; 0 = None */
int ntt_b200_largest_prime_in_arithmetic_progression64(uint64_t factor, uint64_t offset,
                                                       uint64_t lo, uint64_t hi, uint64_t *out);

/* ------------------------------------------------------------------------------------------
 * custum_radix   (tfhe-ntt/src/custum_radix/mod.rs:1-22, exported at lib.rs:116)
 * The fork's recursive CYCLIC transforms of u32 vectors: natural order in and out, over a caller-built
 * table twiddles[k] = root^k mod p (root of order n).  `kind` names the reference routine; on such a
 * table the three forward routines are one function, the inverse routines differ by the constant
 * their bases apply (see csrc/capi_custum_radix.cu).  n must be a power of two (1 allowed) and the
 * table at least n long (the reference indexes it modulo n): NTT_B200_ERR_LEN otherwise, where the
 * reference overflows its stack or panics on an index.
 * PRECONDITION (not checked): the table is the power table the module's make_twiddles builds
 * (fwd.rs:72-103) -- twiddles[0] = 1, twiddles[k] = twiddles[k-1] * twiddles[1] mod p, every entry
 * below p -- and the inputs are below p.  For any other table the reference's radix-4 and split-radix
 * recursions (which read tw[(i + q n/4) % n], tw[2k], tw[3k] and J = tw[n/4] directly) and the GPU's
 * single schedule are different functions of the table and the results diverge.
 * The _mut entry points keep every level of one vector in shared memory: n <= 4096; a longer power-of-two vector
 * answers NTT_B200_ERR_UNSUPPORTED (a capacity limit of this implementation, not a length mismatch of the caller).
 * ------------------------------------------------------------------------------------------ */
#define NTT_B200_CR_RADIX2 0      /* fft_radix2_recursive fwd.rs:170-205 (= fwd_1.rs:190-230), ifft inv.rs:178-230 */
#define NTT_B200_CR_RADIX4 1      /* fft_radix4_recursive fwd.rs:105-168 (= fwd_1.rs:102-188), ifft inv.rs:106-176 */
#define NTT_B200_CR_SPLIT_RADIX 2 /* fft_split_radix_recursive fwd.rs:207-272 (= fwd_1.rs:232-294), ifft inv.rs:232-303 */
#define NTT_B200_CR_RADIX4_MUT 3  /* inverse only: ifft_radix4_recursive_mut fwd_1.rs:296-379 */

/* fft_*_recursive(a: &mut [u32], twiddles: &[u32], p: u32), host memory, in place */
int ntt_b200_custum_radix_fft(int kind, uint32_t *a, size_t n, const uint32_t *twiddles,
                              size_t tw_len, uint32_t p);
/* ifft_*_recursive(a, inv_twiddles, p, n_inv, top) */
int ntt_b200_custum_radix_ifft(int kind, uint32_t *a, size_t n, const uint32_t *inv_twiddles,
                               size_t tw_len, uint32_t p, uint32_t n_inv, int top);
/* fft_{radix2,radix4,split_radix}_recursive_mut(a, twiddles, p, stats: &mut MultStats)   fwd_1.rs:102-294
 * and ifft_radix4_recursive_mut(a, inv_twiddles, p, n_inv, top, stats)                  fwd_1.rs:296-379:
 * the values of the routines above plus the fork's multiplication counters, ADDED onto stats[0]
 * (MultStats::nonzero_mults) and stats[1] (skipped_mults), fwd_1.rs:3-7, :28-37.  One vector per call,
 * n <= 4096 (all levels of the transform are kept in shared memory); NTT_B200_ERR_UNSUPPORTED above that. */
int ntt_b200_custum_radix_fft_mut(int kind, uint32_t *a, size_t n, const uint32_t *twiddles,
                                  size_t tw_len, uint32_t p, uint64_t *stats);
/* NEW: the same for `batch` contiguous vectors in one launch (one CTA per vector): the counters of vector v
 * are added onto stats[2v] and stats[2v + 1] -- the shape of the fork's dataset builder (examples/model/Dataset.rs),
 * which runs the three routines over many inputs. */
int ntt_b200_custum_radix_fft_mut_batch(int kind, uint32_t *host, size_t n, size_t batch,
                                        const uint32_t *twiddles, size_t tw_len, uint32_t p,
                                        uint64_t *stats);
int ntt_b200_custum_radix_ifft_radix4_mut(uint32_t *a, size_t n, const uint32_t *inv_twiddles,
                                          size_t tw_len, uint32_t p, uint32_t n_inv, int top,
                                          uint64_t *stats);
/* NEW: `batch` contiguous vectors in host memory */
int ntt_b200_custum_radix_fft_batch(int kind, uint32_t *host, size_t n, size_t batch,
                                    const uint32_t *twiddles, size_t tw_len, uint32_t p);
int ntt_b200_custum_radix_ifft_batch(int kind, uint32_t *host, size_t n, size_t batch,
                                     const uint32_t *inv_twiddles, size_t tw_len, uint32_t p,
                                     uint32_t n_inv, int top);
/* NEW: device-resident vectors and table on the current device, asynchronous on `stream` */
int ntt_b200_custum_radix_fft_device(int kind, uint32_t *dev, size_t n, size_t batch,
                                     const uint32_t *twiddles_dev, size_t tw_len, uint32_t p,
                                     void *stream);
int ntt_b200_custum_radix_ifft_device(int kind, uint32_t *dev, size_t n, size_t batch,
                                      const uint32_t *inv_twiddles_dev, size_t tw_len, uint32_t p,
                                      uint32_t n_inv, int top, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* TFHE_NTT_B200_H */
