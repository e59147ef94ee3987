//! `product::Plan` (reference: tfhe-ntt/src/product.rs:139-967).
use crate::ffi::{self, check};

#[derive(Copy, Clone, Debug, PartialEq, Eq)]
pub enum FwdMode { Generic, Bounded(u64) }
#[derive(Copy, Clone, Debug, PartialEq, Eq)]
pub enum InvMode { Replace, Accumulate }

/// Negacyclic NTT plan for 64bit product of distinct primes.
pub struct Plan { raw: *mut ffi::ntt_b200_product_plan }
unsafe impl Send for Plan {}
unsafe impl Sync for Plan {}

impl Plan {
    /// product.rs:153
    pub fn try_new(polynomial_size: usize, modulus: u64, factors: impl AsRef<[u64]>) -> Option<Self> {
        let f = factors.as_ref();
        let mut raw = core::ptr::null_mut();
        match unsafe { ffi::ntt_b200_product_try_new(polynomial_size, modulus, f.as_ptr(), f.len(), &mut raw) } {
            ffi::OK => Some(Self { raw }), ffi::NONE => None, e => { check(e, "product::Plan::try_new"); None }
        }
    }
    #[inline] pub fn ntt_size(&self) -> usize { unsafe { ffi::ntt_b200_product_ntt_size(self.raw) } }
    #[inline] pub fn modulus(&self) -> u64 { unsafe { ffi::ntt_b200_product_modulus(self.raw) } }
    pub fn ntt_domain_len(&self) -> usize { unsafe { ffi::ntt_b200_product_ntt_domain_len(self.raw) } }
    /// product.rs:273 — `Bounded` yields the residues of `Generic` whenever its bound holds
    #[track_caller]
    pub fn fwd(&self, ntt: &mut [u64], standard: &[u64], _mode: FwdMode) {
        check(unsafe { ffi::ntt_b200_product_fwd(self.raw, ntt.as_mut_ptr(), ntt.len(), standard.as_ptr(), standard.len()) }, "product::Plan::fwd")
    }
    /// product.rs:360
    #[track_caller]
    pub fn inv(&self, standard: &mut [u64], ntt: &mut [u64], mode: InvMode) {
        check(unsafe { ffi::ntt_b200_product_inv(self.raw, standard.as_mut_ptr(), standard.len(), ntt.as_mut_ptr(), ntt.len(), (mode == InvMode::Accumulate) as i32) }, "product::Plan::inv")
    }
    #[track_caller]
    pub fn mul_assign_normalize(&self, lhs: &mut [u64], rhs: &[u64]) {
        check(unsafe { ffi::ntt_b200_product_mul_assign_normalize(self.raw, lhs.as_mut_ptr(), lhs.len(), rhs.as_ptr(), rhs.len()) }, "mul_assign_normalize")
    }
    #[track_caller]
    pub fn normalize(&self, values: &mut [u64]) {
        check(unsafe { ffi::ntt_b200_product_normalize(self.raw, values.as_mut_ptr(), values.len()) }, "normalize")
    }
    #[track_caller]
    pub fn mul_accumulate(&self, acc: &mut [u64], lhs: &[u64], rhs: &[u64]) {
        check(unsafe { ffi::ntt_b200_product_mul_accumulate(self.raw, acc.as_mut_ptr(), acc.len(), lhs.as_ptr(), lhs.len(), rhs.as_ptr(), rhs.len()) }, "mul_accumulate")
    }
}
impl Drop for Plan { fn drop(&mut self) { unsafe { ffi::ntt_b200_product_free(self.raw) } } }
