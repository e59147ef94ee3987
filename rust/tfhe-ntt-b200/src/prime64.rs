//! `prime64::Plan` (reference: tfhe-ntt/src/prime64.rs:245-1223).
use crate::ffi::{self, check};
use core::ptr;

pub const SOLINAS_PRIME: u64 = ((1u128 << 64) - (1u128 << 32) + 1) as u64;
#[derive(Copy, Clone, Debug)]
pub struct Solinas;
impl Solinas { pub const P: u64 = SOLINAS_PRIME; }

/// Negacyclic NTT plan for 64bit primes.
pub struct Plan { raw: *mut ffi::ntt_b200_plan64 }
// tables are immutable after construction and every entry point is re-entrant
unsafe impl Send for Plan {}
unsafe impl Sync for Plan {}

impl Plan {
    /// prime64.rs:764
    pub fn try_new(polynomial_size: usize, modulus: u64) -> Option<Self> {
        let mut raw = ptr::null_mut();
        match unsafe { ffi::ntt_b200_plan64_try_new(polynomial_size, modulus, &mut raw) } {
            ffi::OK => Some(Self { raw }),
            ffi::NONE => None,
            e => { check(e, "prime64::Plan::try_new"); None }
        }
    }
    /// the C handle, for the sibling modules that take a plan (ntt64_pbs)
    #[inline] pub(crate) fn as_raw(&self) -> *const ffi::ntt_b200_plan64 { self.raw }
    #[inline] pub fn ntt_size(&self) -> usize { unsafe { ffi::ntt_b200_plan64_ntt_size(self.raw) } }
    #[inline] pub fn modulus(&self) -> u64 { unsafe { ffi::ntt_b200_plan64_modulus(self.raw) } }
    #[inline] pub fn use_ifma(&self) -> bool { unsafe { ffi::ntt_b200_plan64_use_ifma(self.raw) != 0 } }
    #[inline] pub fn can_use_fast_reduction_code(&self) -> bool {
        unsafe { ffi::ntt_b200_plan64_can_use_fast_reduction_code(self.raw) != 0 }
    }
    /// prime64.rs:897 — standard order in, bit-reversed order out
    pub fn fwd(&self, buf: &mut [u64]) {
        check(unsafe { ffi::ntt_b200_plan64_fwd(self.raw, buf.as_mut_ptr(), buf.len()) }, "prime64::Plan::fwd")
    }
    /// prime64.rs:975
    pub fn inv(&self, buf: &mut [u64]) {
        check(unsafe { ffi::ntt_b200_plan64_inv(self.raw, buf.as_mut_ptr(), buf.len()) }, "prime64::Plan::inv")
    }
    /// prime64.rs:1050
    pub fn mul_assign_normalize(&self, lhs: &mut [u64], rhs: &[u64]) {
        check(unsafe { ffi::ntt_b200_plan64_mul_assign_normalize(self.raw, lhs.as_mut_ptr(), lhs.len(), rhs.as_ptr(), rhs.len()) }, "mul_assign_normalize")
    }
    /// prime64.rs:1137
    pub fn normalize(&self, values: &mut [u64]) {
        check(unsafe { ffi::ntt_b200_plan64_normalize(self.raw, values.as_mut_ptr(), values.len()) }, "normalize")
    }
    /// prime64.rs:1182
    pub fn mul_accumulate(&self, acc: &mut [u64], lhs: &[u64], rhs: &[u64]) {
        check(unsafe { ffi::ntt_b200_plan64_mul_accumulate(self.raw, acc.as_mut_ptr(), acc.len(), lhs.as_ptr(), lhs.len(), rhs.as_ptr(), rhs.len()) }, "mul_accumulate")
    }
    /// New: `polys.len() / ntt_size()` transforms in one call (host memory).
    pub fn fwd_batch(&self, polys: &mut [u64]) {
        assert_eq!(polys.len() % self.ntt_size(), 0);
        check(unsafe { ffi::ntt_b200_plan64_fwd_batch(self.raw, polys.as_mut_ptr(), polys.len() / self.ntt_size()) }, "fwd_batch")
    }
    /// One host batch over several GPUs: `plans[g]` was created on GPU g (`set_device(g)` before
    /// `try_new`); contiguous slices, no exchange between GPUs.
    pub fn fwd_batch_multi_gpu(plans: &[&Plan], polys: &mut [u64]) {
        let raw: Vec<*const ffi::ntt_b200_plan64> = plans.iter().map(|p| p.raw as *const _).collect();
        let n = plans[0].ntt_size();
        assert_eq!(polys.len() % n, 0);
        check(unsafe { ffi::ntt_b200_plan64_fwd_batch_multi_gpu(raw.as_ptr(), raw.len(), polys.as_mut_ptr(), polys.len() / n) }, "fwd_batch_multi_gpu")
    }
    pub fn inv_batch_multi_gpu(plans: &[&Plan], polys: &mut [u64]) {
        let raw: Vec<*const ffi::ntt_b200_plan64> = plans.iter().map(|p| p.raw as *const _).collect();
        let n = plans[0].ntt_size();
        assert_eq!(polys.len() % n, 0);
        check(unsafe { ffi::ntt_b200_plan64_inv_batch_multi_gpu(raw.as_ptr(), raw.len(), polys.as_mut_ptr(), polys.len() / n) }, "inv_batch_multi_gpu")
    }
    pub fn inv_batch(&self, polys: &mut [u64]) {
        assert_eq!(polys.len() % self.ntt_size(), 0);
        check(unsafe { ffi::ntt_b200_plan64_inv_batch(self.raw, polys.as_mut_ptr(), polys.len() / self.ntt_size()) }, "inv_batch")
    }
    /// New: device-resident, asynchronous on `stream` (a `cudaStream_t`).
    /// # Safety
    /// `dev` must point at `batch * ntt_size()` u64 on the plan's GPU.
    pub unsafe fn fwd_device(&self, dev: *mut u64, batch: usize, stream: *mut core::ffi::c_void) {
        check(ffi::ntt_b200_plan64_fwd_device(self.raw, dev, batch, stream), "fwd_device")
    }
    /// # Safety
    /// as [`Plan::fwd_device`]
    pub unsafe fn inv_device(&self, dev: *mut u64, batch: usize, stream: *mut core::ffi::c_void) {
        check(ffi::ntt_b200_plan64_inv_device(self.raw, dev, batch, stream), "inv_device")
    }
    /// New: `out = inv(acc + fwd(lhs) * rhs)` on host slices in one pipelined call; `rhs` / `acc` may
    /// hold fewer polynomials than `lhs` (reused cyclically).
    pub fn fwd_mac_inv_batch(&self, out: &mut [u64], lhs: &[u64], rhs: &[u64], acc: Option<&[u64]>) {
        let n = self.ntt_size();
        assert_eq!(out.len(), lhs.len());
        assert_eq!(lhs.len() % n, 0);
        let (ap, al) = acc.map_or((core::ptr::null(), 0), |a| (a.as_ptr(), a.len() / n));
        check(unsafe { ffi::ntt_b200_plan64_fwd_mac_inv_batch(self.raw, out.as_mut_ptr(), lhs.as_ptr(), rhs.as_ptr(), rhs.len() / n, ap, al, lhs.len() / n) }, "fwd_mac_inv_batch")
    }
    /// New: the NTT core of the PBS external product on device memory:
    /// `out[b][c] = inv(sum_r fwd(input[b][r]) * ggsw[r][c])`.
    /// # Safety
    /// device pointers on the plan's GPU with `batch*rows*n`, `rows*cols*n`, `batch*cols*n` elements.
    pub unsafe fn ext_product_device(&self, out: *mut u64, input: *const u64, ggsw: *const u64, rows: usize, cols: usize, batch: usize, stream: *mut core::ffi::c_void) {
        check(ffi::ntt_b200_plan64_ext_product_device(self.raw, out, input, ggsw, rows, cols, batch, stream), "ext_product_device")
    }
    /// New: `Plan::normalize` on device memory (prime64.rs normalize), asynchronous on `stream`.
    /// # Safety
    /// `dev` must point at `len` elements on the plan's GPU.
    pub unsafe fn normalize_device(&self, dev: *mut u64, len: usize, stream: *mut core::ffi::c_void) {
        check(ffi::ntt_b200_plan64_normalize_device(self.raw, dev, len, stream), "normalize_device")
    }
    /// New: `Plan::mul_assign_normalize` on device memory; `rhs` may be shorter than `lhs` (its length
    /// divides `len`: one operand shared by a batch).
    /// # Safety
    /// device pointers on the plan's GPU with `len` / `rhs_len` elements.
    pub unsafe fn mul_assign_normalize_device(&self, lhs: *mut u64, len: usize, rhs: *const u64, rhs_len: usize, stream: *mut core::ffi::c_void) {
        check(ffi::ntt_b200_plan64_mul_assign_normalize_device(self.raw, lhs, len, rhs, rhs_len, stream), "mul_assign_normalize_device")
    }
    /// New: `Plan::mul_accumulate` on device memory; `lhs` / `rhs` may be batch-shared (their lengths divide `len`).
    /// # Safety
    /// device pointers on the plan's GPU with `len` / `lhs_len` / `rhs_len` elements.
    pub unsafe fn mul_accumulate_device(&self, acc: *mut u64, len: usize, lhs: *const u64, lhs_len: usize, rhs: *const u64, rhs_len: usize, stream: *mut core::ffi::c_void) {
        check(ffi::ntt_b200_plan64_mul_accumulate_device(self.raw, acc, len, lhs, lhs_len, rhs, rhs_len, stream), "mul_accumulate_device")
    }
    /// New: `out = inv(acc + fwd(lhs) * rhs)` on device memory in one kernel (`acc` may be null).
    /// # Safety
    /// device pointers on the plan's GPU; `rhs_polys` / `acc_polys` divide `batch`.
    pub unsafe fn fwd_mac_inv_device(&self, out: *mut u64, lhs: *const u64, rhs: *const u64, rhs_polys: usize, acc: *const u64, acc_polys: usize, batch: usize, stream: *mut core::ffi::c_void) {
        check(ffi::ntt_b200_plan64_fwd_mac_inv_device(self.raw, out, lhs, rhs, rhs_polys, acc, acc_polys, batch, stream), "fwd_mac_inv_device")
    }
    pub(crate) unsafe fn borrowed(raw: *const ffi::ntt_b200_plan64) -> core::mem::ManuallyDrop<Self> {
        core::mem::ManuallyDrop::new(Self { raw: raw as *mut _ })
    }
}
impl Clone for Plan {
    fn clone(&self) -> Self {
        let mut raw = ptr::null_mut();
        check(unsafe { ffi::ntt_b200_plan64_clone(self.raw, &mut raw) }, "clone");
        Self { raw }
    }
}
impl Drop for Plan { fn drop(&mut self) { unsafe { ffi::ntt_b200_plan64_free(self.raw) } } }
impl core::fmt::Debug for Plan {
    // prime64.rs:263-270
    fn fmt(&self, f: &mut core::fmt::Formatter<'_>) -> core::fmt::Result {
        f.debug_struct("Plan").field("ntt_size", &self.ntt_size()).field("modulus", &self.modulus()).finish()
    }
}
