//! `prime32::Plan` (reference: tfhe-ntt/src/prime32.rs:632-1016).
use crate::ffi::{self, check};
use core::ptr;


/// Negacyclic NTT plan for 32bit primes.
pub struct Plan { raw: *mut ffi::ntt_b200_plan32 }
// tables are immutable after construction and every entry point is re-entrant
unsafe impl Send for Plan {}
unsafe impl Sync for Plan {}

impl Plan {
    /// prime32.rs:662
    pub fn try_new(polynomial_size: usize, modulus: u32) -> Option<Self> {
        let mut raw = ptr::null_mut();
        match unsafe { ffi::ntt_b200_plan32_try_new(polynomial_size, modulus, &mut raw) } {
            ffi::OK => Some(Self { raw }),
            ffi::NONE => None,
            e => { check(e, "prime32::Plan::try_new"); None }
        }
    }
    #[inline] pub fn ntt_size(&self) -> usize { unsafe { ffi::ntt_b200_plan32_ntt_size(self.raw) } }
    #[inline] pub fn modulus(&self) -> u32 { unsafe { ffi::ntt_b200_plan32_modulus(self.raw) } }
    #[inline] pub fn can_use_fast_reduction_code(&self) -> bool {
        unsafe { ffi::ntt_b200_plan32_can_use_fast_reduction_code(self.raw) != 0 }
    }
    /// prime32.rs:797 — standard order in, bit-reversed order out
    pub fn fwd(&self, buf: &mut [u32]) {
        check(unsafe { ffi::ntt_b200_plan32_fwd(self.raw, buf.as_mut_ptr(), buf.len()) }, "prime32::Plan::fwd")
    }
    /// prime32.rs:850
    pub fn inv(&self, buf: &mut [u32]) {
        check(unsafe { ffi::ntt_b200_plan32_inv(self.raw, buf.as_mut_ptr(), buf.len()) }, "prime32::Plan::inv")
    }
    /// prime32.rs:900
    pub fn mul_assign_normalize(&self, lhs: &mut [u32], rhs: &[u32]) {
        check(unsafe { ffi::ntt_b200_plan32_mul_assign_normalize(self.raw, lhs.as_mut_ptr(), lhs.len(), rhs.as_ptr(), rhs.len()) }, "mul_assign_normalize")
    }
    /// prime32.rs:956
    pub fn normalize(&self, values: &mut [u32]) {
        check(unsafe { ffi::ntt_b200_plan32_normalize(self.raw, values.as_mut_ptr(), values.len()) }, "normalize")
    }
    /// prime32.rs:993
    pub fn mul_accumulate(&self, acc: &mut [u32], lhs: &[u32], rhs: &[u32]) {
        check(unsafe { ffi::ntt_b200_plan32_mul_accumulate(self.raw, acc.as_mut_ptr(), acc.len(), lhs.as_ptr(), lhs.len(), rhs.as_ptr(), rhs.len()) }, "mul_accumulate")
    }
    /// New: `polys.len() / ntt_size()` transforms in one call (host memory).
    pub fn fwd_batch(&self, polys: &mut [u32]) {
        assert_eq!(polys.len() % self.ntt_size(), 0);
        check(unsafe { ffi::ntt_b200_plan32_fwd_batch(self.raw, polys.as_mut_ptr(), polys.len() / self.ntt_size()) }, "fwd_batch")
    }
    /// One host batch over several GPUs: `plans[g]` was created on GPU g (`set_device(g)` before
    /// `try_new`); contiguous slices, no exchange between GPUs.
    pub fn fwd_batch_multi_gpu(plans: &[&Plan], polys: &mut [u32]) {
        let raw: Vec<*const ffi::ntt_b200_plan32> = plans.iter().map(|p| p.raw as *const _).collect();
        let n = plans[0].ntt_size();
        assert_eq!(polys.len() % n, 0);
        check(unsafe { ffi::ntt_b200_plan32_fwd_batch_multi_gpu(raw.as_ptr(), raw.len(), polys.as_mut_ptr(), polys.len() / n) }, "fwd_batch_multi_gpu")
    }
    pub fn inv_batch_multi_gpu(plans: &[&Plan], polys: &mut [u32]) {
        let raw: Vec<*const ffi::ntt_b200_plan32> = plans.iter().map(|p| p.raw as *const _).collect();
        let n = plans[0].ntt_size();
        assert_eq!(polys.len() % n, 0);
        check(unsafe { ffi::ntt_b200_plan32_inv_batch_multi_gpu(raw.as_ptr(), raw.len(), polys.as_mut_ptr(), polys.len() / n) }, "inv_batch_multi_gpu")
    }
    pub fn inv_batch(&self, polys: &mut [u32]) {
        assert_eq!(polys.len() % self.ntt_size(), 0);
        check(unsafe { ffi::ntt_b200_plan32_inv_batch(self.raw, polys.as_mut_ptr(), polys.len() / self.ntt_size()) }, "inv_batch")
    }
    /// New: device-resident, asynchronous on `stream` (a `cudaStream_t`).
    /// # Safety
    /// `dev` must point at `batch * ntt_size()` u32 on the plan's GPU.
    pub unsafe fn fwd_device(&self, dev: *mut u32, batch: usize, stream: *mut core::ffi::c_void) {
        check(ffi::ntt_b200_plan32_fwd_device(self.raw, dev, batch, stream), "fwd_device")
    }
    /// # Safety
    /// as [`Plan::fwd_device`]
    pub unsafe fn inv_device(&self, dev: *mut u32, batch: usize, stream: *mut core::ffi::c_void) {
        check(ffi::ntt_b200_plan32_inv_device(self.raw, dev, batch, stream), "inv_device")
    }
    /// New: `Plan::normalize` on device memory (prime32.rs normalize), asynchronous on `stream`.
    /// # Safety
    /// `dev` must point at `len` elements on the plan's GPU.
    pub unsafe fn normalize_device(&self, dev: *mut u32, len: usize, stream: *mut core::ffi::c_void) {
        check(ffi::ntt_b200_plan32_normalize_device(self.raw, dev, len, stream), "normalize_device")
    }
    /// New: `Plan::mul_assign_normalize` on device memory; `rhs` may be shorter than `lhs` (its length
    /// divides `len`: one operand shared by a batch).
    /// # Safety
    /// device pointers on the plan's GPU with `len` / `rhs_len` elements.
    pub unsafe fn mul_assign_normalize_device(&self, lhs: *mut u32, len: usize, rhs: *const u32, rhs_len: usize, stream: *mut core::ffi::c_void) {
        check(ffi::ntt_b200_plan32_mul_assign_normalize_device(self.raw, lhs, len, rhs, rhs_len, stream), "mul_assign_normalize_device")
    }
    /// New: `Plan::mul_accumulate` on device memory; `lhs` / `rhs` may be batch-shared (their lengths divide `len`).
    /// # Safety
    /// device pointers on the plan's GPU with `len` / `lhs_len` / `rhs_len` elements.
    pub unsafe fn mul_accumulate_device(&self, acc: *mut u32, len: usize, lhs: *const u32, lhs_len: usize, rhs: *const u32, rhs_len: usize, stream: *mut core::ffi::c_void) {
        check(ffi::ntt_b200_plan32_mul_accumulate_device(self.raw, acc, len, lhs, lhs_len, rhs, rhs_len, stream), "mul_accumulate_device")
    }
    /// New: `out = inv(acc + fwd(lhs) * rhs)` on device memory in one kernel (`acc` may be null).
    /// # Safety
    /// device pointers on the plan's GPU; `rhs_polys` / `acc_polys` divide `batch`.
    pub unsafe fn fwd_mac_inv_device(&self, out: *mut u32, lhs: *const u32, rhs: *const u32, rhs_polys: usize, acc: *const u32, acc_polys: usize, batch: usize, stream: *mut core::ffi::c_void) {
        check(ffi::ntt_b200_plan32_fwd_mac_inv_device(self.raw, out, lhs, rhs, rhs_polys, acc, acc_polys, batch, stream), "fwd_mac_inv_device")
    }
    pub(crate) unsafe fn borrowed(raw: *const ffi::ntt_b200_plan32) -> core::mem::ManuallyDrop<Self> {
        core::mem::ManuallyDrop::new(Self { raw: raw as *mut _ })
    }
}
impl Clone for Plan {
    fn clone(&self) -> Self {
        let mut raw = ptr::null_mut();
        check(unsafe { ffi::ntt_b200_plan32_clone(self.raw, &mut raw) }, "clone");
        Self { raw }
    }
}
impl Drop for Plan { fn drop(&mut self) { unsafe { ffi::ntt_b200_plan32_free(self.raw) } } }
impl core::fmt::Debug for Plan {
    // prime32.rs:650-657
    fn fmt(&self, f: &mut core::fmt::Formatter<'_>) -> core::fmt::Result {
        f.debug_struct("Plan").field("ntt_size", &self.ntt_size()).field("modulus", &self.modulus()).finish()
    }
}
