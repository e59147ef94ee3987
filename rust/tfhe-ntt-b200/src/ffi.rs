//! Hand-written `extern "C"` block for include/tfhe_ntt_b200.h (no bindgen: no libclang needed).
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

pub const OK: c_int = 0;
pub const NONE: c_int = 1;
pub const ERR_LEN: c_int = 2;
pub const ERR_CUDA: c_int = 3;
pub const ERR_ARG: c_int = 4;
pub const ERR_UNSUPPORTED: c_int = 5;

#[repr(C)] pub struct ntt_b200_plan64 { _p: [u8; 0] }
#[repr(C)] pub struct ntt_b200_plan32 { _p: [u8; 0] }
#[repr(C)] pub struct ntt_b200_native_plan { _p: [u8; 0] }
#[repr(C)] pub struct ntt_b200_product_plan { _p: [u8; 0] }
#[repr(C)] pub struct ntt_b200_bsk { _p: [u8; 0] }

extern "C" {
    pub fn ntt_b200_last_error() -> *const c_char;
    pub fn ntt_b200_set_device(device: c_int) -> c_int;
    pub fn ntt_b200_is_prime64(n: u64) -> c_int;
    pub fn ntt_b200_largest_prime_in_arithmetic_progression64(factor: u64, offset: u64, lo: u64, hi: u64, out: *mut u64) -> c_int;

    pub fn ntt_b200_plan64_try_new(n: usize, p: u64, out: *mut *mut ntt_b200_plan64) -> c_int;
    pub fn ntt_b200_plan64_clone(plan: *const ntt_b200_plan64, out: *mut *mut ntt_b200_plan64) -> c_int;
    pub fn ntt_b200_plan64_free(plan: *mut ntt_b200_plan64);
    pub fn ntt_b200_plan64_ntt_size(plan: *const ntt_b200_plan64) -> usize;
    pub fn ntt_b200_plan64_modulus(plan: *const ntt_b200_plan64) -> u64;
    pub fn ntt_b200_plan64_use_ifma(plan: *const ntt_b200_plan64) -> c_int;
    pub fn ntt_b200_plan64_can_use_fast_reduction_code(plan: *const ntt_b200_plan64) -> c_int;
    pub fn ntt_b200_plan64_fwd(plan: *const ntt_b200_plan64, buf: *mut u64, len: usize) -> c_int;
    pub fn ntt_b200_plan64_inv(plan: *const ntt_b200_plan64, buf: *mut u64, len: usize) -> c_int;
    pub fn ntt_b200_plan64_normalize(plan: *const ntt_b200_plan64, values: *mut u64, len: usize) -> c_int;
    pub fn ntt_b200_plan64_mul_assign_normalize(plan: *const ntt_b200_plan64, lhs: *mut u64, lhs_len: usize, rhs: *const u64, rhs_len: usize) -> c_int;
    pub fn ntt_b200_plan64_mul_accumulate(plan: *const ntt_b200_plan64, acc: *mut u64, acc_len: usize, lhs: *const u64, lhs_len: usize, rhs: *const u64, rhs_len: usize) -> c_int;
    pub fn ntt_b200_plan64_fwd_batch(plan: *const ntt_b200_plan64, host: *mut u64, batch: usize) -> c_int;
    pub fn ntt_b200_plan64_inv_batch(plan: *const ntt_b200_plan64, host: *mut u64, batch: usize) -> c_int;
    pub fn ntt_b200_plan64_fwd_batch_multi_gpu(plans: *const *const ntt_b200_plan64, n_plans: usize, host: *mut u64, batch: usize) -> c_int;
    pub fn ntt_b200_plan64_inv_batch_multi_gpu(plans: *const *const ntt_b200_plan64, n_plans: usize, host: *mut u64, batch: usize) -> c_int;
    pub fn ntt_b200_plan64_fwd_device(plan: *const ntt_b200_plan64, dev: *mut u64, batch: usize, stream: *mut c_void) -> c_int;
    pub fn ntt_b200_plan64_inv_device(plan: *const ntt_b200_plan64, dev: *mut u64, batch: usize, stream: *mut c_void) -> c_int;

    pub fn ntt_b200_plan64_normalize_device(plan: *const ntt_b200_plan64, dev: *mut u64, len: usize, stream: *mut c_void) -> c_int;
    pub fn ntt_b200_plan64_mul_assign_normalize_device(plan: *const ntt_b200_plan64, lhs: *mut u64, len: usize, rhs: *const u64, rhs_len: usize, stream: *mut c_void) -> c_int;
    pub fn ntt_b200_plan64_mul_accumulate_device(plan: *const ntt_b200_plan64, acc: *mut u64, len: usize, lhs: *const u64, lhs_len: usize, rhs: *const u64, rhs_len: usize, stream: *mut c_void) -> c_int;
    pub fn ntt_b200_plan64_fwd_mac_inv_device(plan: *const ntt_b200_plan64, out: *mut u64, lhs: *const u64, rhs: *const u64, rhs_polys: usize, acc: *const u64, acc_polys: usize, batch: usize, stream: *mut c_void) -> c_int;
    pub fn ntt_b200_plan64_fwd_mac_inv_batch(plan: *const ntt_b200_plan64, out: *mut u64, lhs: *const u64, rhs: *const u64, rhs_polys: usize, acc: *const u64, acc_polys: usize, batch: usize) -> c_int;
    pub fn ntt_b200_plan64_ext_product_device(plan: *const ntt_b200_plan64, out: *mut u64, input: *const u64, ggsw: *const u64, rows: usize, cols: usize, batch: usize, stream: *mut c_void) -> c_int;

    pub fn ntt_b200_plan32_try_new(n: usize, p: u32, out: *mut *mut ntt_b200_plan32) -> c_int;
    pub fn ntt_b200_plan32_clone(plan: *const ntt_b200_plan32, out: *mut *mut ntt_b200_plan32) -> c_int;
    pub fn ntt_b200_plan32_free(plan: *mut ntt_b200_plan32);
    pub fn ntt_b200_plan32_ntt_size(plan: *const ntt_b200_plan32) -> usize;
    pub fn ntt_b200_plan32_modulus(plan: *const ntt_b200_plan32) -> u32;
    pub fn ntt_b200_plan32_can_use_fast_reduction_code(plan: *const ntt_b200_plan32) -> c_int;
    pub fn ntt_b200_plan32_fwd(plan: *const ntt_b200_plan32, buf: *mut u32, len: usize) -> c_int;
    pub fn ntt_b200_plan32_inv(plan: *const ntt_b200_plan32, buf: *mut u32, len: usize) -> c_int;
    pub fn ntt_b200_plan32_normalize(plan: *const ntt_b200_plan32, values: *mut u32, len: usize) -> c_int;
    pub fn ntt_b200_plan32_mul_assign_normalize(plan: *const ntt_b200_plan32, lhs: *mut u32, lhs_len: usize, rhs: *const u32, rhs_len: usize) -> c_int;
    pub fn ntt_b200_plan32_mul_accumulate(plan: *const ntt_b200_plan32, acc: *mut u32, acc_len: usize, lhs: *const u32, lhs_len: usize, rhs: *const u32, rhs_len: usize) -> c_int;
    pub fn ntt_b200_plan32_fwd_batch(plan: *const ntt_b200_plan32, host: *mut u32, batch: usize) -> c_int;
    pub fn ntt_b200_plan32_inv_batch(plan: *const ntt_b200_plan32, host: *mut u32, batch: usize) -> c_int;
    pub fn ntt_b200_plan32_fwd_batch_multi_gpu(plans: *const *const ntt_b200_plan32, n_plans: usize, host: *mut u32, batch: usize) -> c_int;
    pub fn ntt_b200_plan32_inv_batch_multi_gpu(plans: *const *const ntt_b200_plan32, n_plans: usize, host: *mut u32, batch: usize) -> c_int;

    pub fn ntt_b200_plan32_normalize_device(plan: *const ntt_b200_plan32, dev: *mut u32, len: usize, stream: *mut c_void) -> c_int;
    pub fn ntt_b200_plan32_mul_assign_normalize_device(plan: *const ntt_b200_plan32, lhs: *mut u32, len: usize, rhs: *const u32, rhs_len: usize, stream: *mut c_void) -> c_int;
    pub fn ntt_b200_plan32_mul_accumulate_device(plan: *const ntt_b200_plan32, acc: *mut u32, len: usize, lhs: *const u32, lhs_len: usize, rhs: *const u32, rhs_len: usize, stream: *mut c_void) -> c_int;
    pub fn ntt_b200_plan32_fwd_mac_inv_device(plan: *const ntt_b200_plan32, out: *mut u32, lhs: *const u32, rhs: *const u32, rhs_polys: usize, acc: *const u32, acc_polys: usize, batch: usize, stream: *mut c_void) -> c_int;
    pub fn ntt_b200_plan32_fwd_device(plan: *const ntt_b200_plan32, dev: *mut u32, batch: usize, stream: *mut c_void) -> c_int;
    pub fn ntt_b200_plan32_inv_device(plan: *const ntt_b200_plan32, dev: *mut u32, batch: usize, stream: *mut c_void) -> c_int;

    pub fn ntt_b200_product_try_new(n: usize, modulus: u64, factors: *const u64, nfactors: usize, out: *mut *mut ntt_b200_product_plan) -> c_int;
    pub fn ntt_b200_product_free(plan: *mut ntt_b200_product_plan);
    pub fn ntt_b200_product_ntt_size(plan: *const ntt_b200_product_plan) -> usize;
    pub fn ntt_b200_product_modulus(plan: *const ntt_b200_product_plan) -> u64;
    pub fn ntt_b200_product_ntt_domain_len(plan: *const ntt_b200_product_plan) -> usize;
    pub fn ntt_b200_product_fwd(plan: *const ntt_b200_product_plan, ntt: *mut u64, ntt_len: usize, standard: *const u64, standard_len: usize) -> c_int;
    pub fn ntt_b200_product_inv(plan: *const ntt_b200_product_plan, standard: *mut u64, standard_len: usize, ntt: *mut u64, ntt_len: usize, accumulate: c_int) -> c_int;
    pub fn ntt_b200_product_normalize(plan: *const ntt_b200_product_plan, values: *mut u64, len: usize) -> c_int;
    pub fn ntt_b200_product_mul_assign_normalize(plan: *const ntt_b200_product_plan, lhs: *mut u64, lhs_len: usize, rhs: *const u64, rhs_len: usize) -> c_int;
    pub fn ntt_b200_product_mul_accumulate(plan: *const ntt_b200_product_plan, acc: *mut u64, acc_len: usize, lhs: *const u64, lhs_len: usize, rhs: *const u64, rhs_len: usize) -> c_int;

    pub fn ntt_b200_bsk_new(plan: *const ntt_b200_plan64, ntt_bsk: *const u64, n_lwe: usize, glwe_size: usize, base_log: u32, level: u32, out: *mut *mut ntt_b200_bsk) -> c_int;
    pub fn ntt_b200_bsk_convert_new(plan: *const ntt_b200_plan64, standard_bsk: *const u64, n_lwe: usize, glwe_size: usize, base_log: u32, level: u32, input_width: u32, normalize: c_int, out: *mut *mut ntt_b200_bsk) -> c_int;
    pub fn ntt_b200_convert_standard_lwe_bootstrap_key_to_ntt64(plan: *const ntt_b200_plan64, input: *const u64, output: *mut u64, len: usize, input_width: u32, normalize: c_int) -> c_int;
    pub fn ntt_b200_bsk_free(key: *mut ntt_b200_bsk);
    pub fn ntt_b200_bsk_input_lwe_dimension(key: *const ntt_b200_bsk) -> usize;
    pub fn ntt_b200_bsk_glwe_size(key: *const ntt_b200_bsk) -> usize;
    pub fn ntt_b200_bsk_polynomial_size(key: *const ntt_b200_bsk) -> usize;
    pub fn ntt_b200_bsk_decomposition_base_log(key: *const ntt_b200_bsk) -> u32;
    pub fn ntt_b200_bsk_decomposition_level_count(key: *const ntt_b200_bsk) -> u32;
    pub fn ntt_b200_bsk_read(key: *const ntt_b200_bsk, out: *mut u64, len: usize) -> c_int;
    pub fn ntt_b200_blind_rotate_ntt64_assign(key: *const ntt_b200_bsk, lwe: *const u64, lut: *mut u64, batch: usize, path: c_int) -> c_int;
    pub fn ntt_b200_blind_rotate_ntt64_bnf_assign(key: *const ntt_b200_bsk, width: u32, msed: *const u64, lut: *mut u64, batch: usize, path: c_int) -> c_int;
    pub fn ntt_b200_programmable_bootstrap_ntt64(key: *const ntt_b200_bsk, lwe_in: *const u64, lwe_out: *mut u64, accumulator: *const u64, acc_count: usize, batch: usize, path: c_int) -> c_int;
    pub fn ntt_b200_programmable_bootstrap_ntt64_bnf(key: *const ntt_b200_bsk, width: u32, lwe_in: *const u64, lwe_out: *mut u64, accumulator: *const u64, acc_count: usize, batch: usize, path: c_int) -> c_int;
    pub fn ntt_b200_blind_rotate_ntt64_device(key: *const ntt_b200_bsk, bnf: c_int, width: u32, lwe: *const u64, lwe_is_switched: c_int, lut: *const u64, lut_count: usize, acc_out: *mut u64, batch: usize, path: c_int, stream: *mut c_void) -> c_int;
    pub fn ntt_b200_extract_lwe_sample_device(key: *const ntt_b200_bsk, bnf: c_int, glwe: *const u64, lwe_out: *mut u64, batch: usize, stream: *mut c_void) -> c_int;

    pub fn ntt_b200_native_try_new(kind: c_int, n: usize, out: *mut *mut ntt_b200_native_plan) -> c_int;
    pub fn ntt_b200_native_free(plan: *mut ntt_b200_native_plan);
    pub fn ntt_b200_native_ntt_size(plan: *const ntt_b200_native_plan) -> usize;
    pub fn ntt_b200_native_ntt_i(plan: *const ntt_b200_native_plan, i: c_int) -> *const c_void;
    pub fn ntt_b200_native_fwd(plan: *const ntt_b200_native_plan, value: *const c_void, len: usize, residues: *const *mut c_void, binary: c_int) -> c_int;
    pub fn ntt_b200_native_inv(plan: *const ntt_b200_native_plan, value: *mut c_void, len: usize, residues: *const *mut c_void) -> c_int;
    pub fn ntt_b200_native_negacyclic_polymul(plan: *const ntt_b200_native_plan, prod: *mut c_void, prod_len: usize, lhs: *const c_void, lhs_len: usize, rhs: *const c_void, rhs_len: usize) -> c_int;
    pub fn ntt_b200_native_negacyclic_polymul_batch(plan: *const ntt_b200_native_plan, prod: *mut c_void, lhs: *const c_void, rhs: *const c_void, batch: usize) -> c_int;
    // custum_radix (tfhe-ntt/src/custum_radix/mod.rs:1-22)
    pub fn ntt_b200_custum_radix_fft(kind: c_int, a: *mut u32, n: usize, twiddles: *const u32, tw_len: usize, p: u32) -> c_int;
    pub fn ntt_b200_custum_radix_ifft(kind: c_int, a: *mut u32, n: usize, inv_twiddles: *const u32, tw_len: usize, p: u32, n_inv: u32, top: c_int) -> c_int;
    pub fn ntt_b200_custum_radix_fft_mut(kind: c_int, a: *mut u32, n: usize, twiddles: *const u32, tw_len: usize, p: u32, stats: *mut u64) -> c_int;
    pub fn ntt_b200_custum_radix_fft_mut_batch(kind: c_int, host: *mut u32, n: usize, batch: usize, twiddles: *const u32, tw_len: usize, p: u32, stats: *mut u64) -> c_int;
    pub fn ntt_b200_custum_radix_ifft_radix4_mut(a: *mut u32, n: usize, inv_twiddles: *const u32, tw_len: usize, p: u32, n_inv: u32, top: c_int, stats: *mut u64) -> c_int;
    pub fn ntt_b200_custum_radix_fft_batch(kind: c_int, host: *mut u32, n: usize, batch: usize, twiddles: *const u32, tw_len: usize, p: u32) -> c_int;
    pub fn ntt_b200_custum_radix_ifft_batch(kind: c_int, host: *mut u32, n: usize, batch: usize, inv_twiddles: *const u32, tw_len: usize, p: u32, n_inv: u32, top: c_int) -> c_int;
    pub fn ntt_b200_custum_radix_fft_device(kind: c_int, dev: *mut u32, n: usize, batch: usize, twiddles_dev: *const u32, tw_len: usize, p: u32, stream: *mut c_void) -> c_int;
    pub fn ntt_b200_custum_radix_ifft_device(kind: c_int, dev: *mut u32, n: usize, batch: usize, inv_twiddles_dev: *const u32, tw_len: usize, p: u32, n_inv: u32, top: c_int, stream: *mut c_void) -> c_int;
}

/// Status -> the reference's behaviour: length errors panic like `assert_eq!` (prime64.rs:898),
/// CUDA failures print and abort like the reference's own CUDA backend
/// (backends/tfhe-cuda-backend/cuda/include/device.h:11-19).
#[track_caller]
pub fn check(status: c_int, what: &str) {
    match status {
        OK => {}
        ERR_LEN => panic!("assertion failed: length mismatch in {what}"),
        ERR_UNSUPPORTED => panic!("{what}: size beyond a capacity limit of tfhe-ntt-b200 (custum_radix *_mut: n <= 4096)"),
        _ => {
            let msg = unsafe { core::ffi::CStr::from_ptr(ntt_b200_last_error()) };
            eprintln!("tfhe-ntt-b200: {what}: {}", msg.to_string_lossy());
            std::process::abort();
        }
    }
}
