//! NTT programmable bootstrap on the B200 engine -- what `tfhe` binds instead of
//! `tfhe/src/core_crypto/algorithms/lwe_programmable_bootstrapping/ntt64_pbs.rs` (classic) and
//! `ntt64_bnf_pbs.rs` (bnf), with the key entity of `entities/ntt_lwe_bootstrap_key.rs` and the
//! conversion of `algorithms/lwe_bootstrap_key_conversion.rs:294-447`.
//! Containers are the reference's flat `&[u64]`; every call takes `len / row` ciphertexts.
use crate::ffi::{self, check};
use crate::prime64::Plan;

/// lwe_bootstrap_key_conversion.rs:283-288
#[derive(Copy, Clone, Debug, PartialEq, Eq)]
pub enum NttLweBootstrapKeyOption { Raw, Normalize }

/// Device path: fused persistent kernel when the shape has one, else composed kernels.
#[derive(Copy, Clone, Debug, PartialEq, Eq)]
pub enum Path { Auto = 0, Fused = 1, Composed = 2, Cluster = 3 }

/// Device-resident `NttLweBootstrapKey` (entities/ntt_lwe_bootstrap_key.rs:26-33).
pub struct NttLweBootstrapKey { raw: *mut ffi::ntt_b200_bsk }
unsafe impl Send for NttLweBootstrapKey {}
unsafe impl Sync for NttLweBootstrapKey {}
impl Drop for NttLweBootstrapKey { fn drop(&mut self) { unsafe { ffi::ntt_b200_bsk_free(self.raw) } } }

impl NttLweBootstrapKey {
    /// `from_container` (:68-110): `container` is already in the NTT domain.
    #[track_caller]
    pub fn from_container(plan: &Plan, container: &[u64], input_lwe_dimension: usize, glwe_size: usize,
                          decomp_base_log: u32, decomp_level_count: u32) -> Self {
        assert_eq!(container.len(),
                   input_lwe_dimension * decomp_level_count as usize * glwe_size * glwe_size * plan.ntt_size());
        let mut raw = core::ptr::null_mut();
        check(unsafe { ffi::ntt_b200_bsk_new(plan.as_raw(), container.as_ptr(), input_lwe_dimension, glwe_size,
                                             decomp_base_log, decomp_level_count, &mut raw) }, "NttLweBootstrapKey::from_container");
        Self { raw }
    }
    /// `convert_standard_lwe_bootstrap_key_to_ntt64` (:294-363) straight into device memory.
    /// `input_modulus_width` = 0 when the standard key is modulo the NTT prime.
    #[track_caller]
    pub fn from_standard(plan: &Plan, standard_bsk: &[u64], input_lwe_dimension: usize, glwe_size: usize,
                         decomp_base_log: u32, decomp_level_count: u32, input_modulus_width: u32,
                         option: NttLweBootstrapKeyOption) -> Self {
        assert_eq!(standard_bsk.len(),
                   input_lwe_dimension * decomp_level_count as usize * glwe_size * glwe_size * plan.ntt_size());
        let mut raw = core::ptr::null_mut();
        check(unsafe { ffi::ntt_b200_bsk_convert_new(plan.as_raw(), standard_bsk.as_ptr(), input_lwe_dimension, glwe_size,
                                                     decomp_base_log, decomp_level_count, input_modulus_width,
                                                     (option == NttLweBootstrapKeyOption::Normalize) as i32, &mut raw) },
              "NttLweBootstrapKey::from_standard");
        Self { raw }
    }
    pub fn input_lwe_dimension(&self) -> usize { unsafe { ffi::ntt_b200_bsk_input_lwe_dimension(self.raw) } }
    pub fn glwe_size(&self) -> usize { unsafe { ffi::ntt_b200_bsk_glwe_size(self.raw) } }
    pub fn polynomial_size(&self) -> usize { unsafe { ffi::ntt_b200_bsk_polynomial_size(self.raw) } }
    pub fn decomposition_base_log(&self) -> u32 { unsafe { ffi::ntt_b200_bsk_decomposition_base_log(self.raw) } }
    pub fn decomposition_level_count(&self) -> u32 { unsafe { ffi::ntt_b200_bsk_decomposition_level_count(self.raw) } }
    pub fn output_lwe_dimension(&self) -> usize { (self.glwe_size() - 1) * self.polynomial_size() }
    pub fn as_container(&self) -> Vec<u64> {
        let len = self.input_lwe_dimension() * self.decomposition_level_count() as usize
            * self.glwe_size() * self.glwe_size() * self.polynomial_size();
        let mut out = vec![0u64; len];
        check(unsafe { ffi::ntt_b200_bsk_read(self.raw, out.as_mut_ptr(), len) }, "NttLweBootstrapKey::as_container");
        out
    }
}

/// lwe_bootstrap_key_conversion.rs:294-363, host to host.
#[track_caller]
pub fn convert_standard_lwe_bootstrap_key_to_ntt64(plan: &Plan, input_bsk: &[u64], output_bsk: &mut [u64],
                                                   option: NttLweBootstrapKeyOption, input_modulus_width: u32) {
    assert_eq!(input_bsk.len(), output_bsk.len());
    check(unsafe { ffi::ntt_b200_convert_standard_lwe_bootstrap_key_to_ntt64(
        plan.as_raw(), input_bsk.as_ptr(), output_bsk.as_mut_ptr(), input_bsk.len(), input_modulus_width,
        (option == NttLweBootstrapKeyOption::Normalize) as i32) }, "convert_standard_lwe_bootstrap_key_to_ntt64")
}

#[track_caller]
fn batch_of(len: usize, row: usize) -> usize { assert_eq!(len % row, 0); len / row }

/// ntt64_pbs.rs:175-286: `lut` (one GLWE per input ciphertext) is rotated in place.
#[track_caller]
pub fn blind_rotate_ntt64_assign(input: &[u64], lut: &mut [u64], bsk: &NttLweBootstrapKey) {
    let batch = batch_of(input.len(), bsk.input_lwe_dimension() + 1);
    assert_eq!(lut.len(), batch * bsk.glwe_size() * bsk.polynomial_size());
    check(unsafe { ffi::ntt_b200_blind_rotate_ntt64_assign(bsk.raw, input.as_ptr(), lut.as_mut_ptr(), batch, Path::Auto as i32) },
          "blind_rotate_ntt64_assign")
}
/// ntt64_bnf_pbs.rs:174-276: `msed_input` = modulus-switched mask and body.
#[track_caller]
pub fn blind_rotate_ntt64_bnf_assign(msed_input: &[u64], lut: &mut [u64], bsk: &NttLweBootstrapKey,
                                     ciphertext_modulus_width: u32) {
    let batch = batch_of(msed_input.len(), bsk.input_lwe_dimension() + 1);
    assert_eq!(lut.len(), batch * bsk.glwe_size() * bsk.polynomial_size());
    check(unsafe { ffi::ntt_b200_blind_rotate_ntt64_bnf_assign(bsk.raw, ciphertext_modulus_width, msed_input.as_ptr(),
                                                               lut.as_mut_ptr(), batch, Path::Auto as i32) },
          "blind_rotate_ntt64_bnf_assign")
}
/// ntt64_pbs.rs:439-538; `accumulator`: one GLWE for the whole batch, or one per input.
#[track_caller]
pub fn programmable_bootstrap_ntt64_lwe_ciphertext(input: &[u64], output: &mut [u64], accumulator: &[u64],
                                                   bsk: &NttLweBootstrapKey) {
    let batch = batch_of(input.len(), bsk.input_lwe_dimension() + 1);
    assert_eq!(output.len(), batch * (bsk.output_lwe_dimension() + 1));
    let acc_count = batch_of(accumulator.len(), bsk.glwe_size() * bsk.polynomial_size());
    check(unsafe { ffi::ntt_b200_programmable_bootstrap_ntt64(bsk.raw, input.as_ptr(), output.as_mut_ptr(),
                                                              accumulator.as_ptr(), acc_count, batch, Path::Auto as i32) },
          "programmable_bootstrap_ntt64_lwe_ciphertext")
}
/// ntt64_bnf_pbs.rs:428-539
#[track_caller]
pub fn programmable_bootstrap_ntt64_bnf_lwe_ciphertext(input: &[u64], output: &mut [u64], accumulator: &[u64],
                                                       bsk: &NttLweBootstrapKey, ciphertext_modulus_width: u32) {
    let batch = batch_of(input.len(), bsk.input_lwe_dimension() + 1);
    assert_eq!(output.len(), batch * (bsk.output_lwe_dimension() + 1));
    let acc_count = batch_of(accumulator.len(), bsk.glwe_size() * bsk.polynomial_size());
    check(unsafe { ffi::ntt_b200_programmable_bootstrap_ntt64_bnf(bsk.raw, ciphertext_modulus_width, input.as_ptr(),
                                                                  output.as_mut_ptr(), accumulator.as_ptr(), acc_count,
                                                                  batch, Path::Auto as i32) },
          "programmable_bootstrap_ntt64_bnf_lwe_ciphertext")
}
