//! CRT plans (reference: native32.rs, native64.rs, native128.rs, native_binary{32,64,128}.rs).
//! One macro instantiates every reference type with its own value / residue types and arity.
use crate::ffi::{self, check};
use core::ffi::c_void;

macro_rules! native_plan {
    ($modname:ident, $ty:ident, $kind:expr, $V:ty, $R:ty, $prime:ident, [$(($res:ident, $acc:ident, $idx:expr)),+], $binary:expr) => {
        pub struct $ty {
            raw: *mut ffi::ntt_b200_native_plan,
            // the per-prime plans, borrowed from the C object (it owns them): what ntt_0() .. return
            subs: Vec<core::mem::ManuallyDrop<crate::$prime::Plan>>,
        }
        unsafe impl Send for $ty {}
        unsafe impl Sync for $ty {}
        impl $ty {
            /// e.g. native64.rs:932 — `None` when any per-prime plan is `None`
            pub fn try_new(n: usize) -> Option<Self> {
                let mut raw = core::ptr::null_mut();
                match unsafe { ffi::ntt_b200_native_try_new($kind, n, &mut raw) } {
                    ffi::OK => {
                        let count = [$($idx),+].len();
                        let subs = (0..count).map(|i| unsafe {
                            crate::$prime::Plan::borrowed(ffi::ntt_b200_native_ntt_i(raw, i as i32) as *const _)
                        }).collect();
                        Some(Self { raw, subs })
                    }
                    ffi::NONE => None,
                    e => { check(e, "try_new"); None }
                }
            }
            #[inline] pub fn ntt_size(&self) -> usize { unsafe { ffi::ntt_b200_native_ntt_size(self.raw) } }
            $(
            /// the reference's accessor of the same name (e.g. native64.rs:950-968, :1095-1103)
            #[inline] pub fn $acc(&self) -> &crate::$prime::Plan { &self.subs[$idx] }
            )+
            /// the same by index
            #[inline] pub fn ntt_i(&self, i: usize) -> &crate::$prime::Plan { &self.subs[i] }
            /// e.g. native64.rs:970
            pub fn fwd(&self, value: &[$V], $($res: &mut [$R]),+) {
                let r = [$($res.as_mut_ptr() as *mut c_void),+];
                check(unsafe { ffi::ntt_b200_native_fwd(self.raw, value.as_ptr() as *const c_void, value.len(), r.as_ptr(), 0) }, "fwd")
            }
            /// native_binary64.rs:371 (present on the binary plans only in the reference)
            pub fn fwd_binary(&self, value: &[$V], $($res: &mut [$R]),+) {
                assert!($binary, "fwd_binary exists on the native_binary plans only");
                let r = [$($res.as_mut_ptr() as *mut c_void),+];
                check(unsafe { ffi::ntt_b200_native_fwd(self.raw, value.as_ptr() as *const c_void, value.len(), r.as_ptr(), 1) }, "fwd_binary")
            }
            /// e.g. native64.rs:1000 — transforms (clobbers) the residue buffers like the reference
            pub fn inv(&self, value: &mut [$V], $($res: &mut [$R]),+) {
                let r = [$($res.as_mut_ptr() as *mut c_void),+];
                check(unsafe { ffi::ntt_b200_native_inv(self.raw, value.as_mut_ptr() as *mut c_void, value.len(), r.as_ptr()) }, "inv")
            }
            /// e.g. native64.rs:1041
            pub fn negacyclic_polymul(&self, prod: &mut [$V], lhs: &[$V], rhs: &[$V]) {
                check(unsafe { ffi::ntt_b200_native_negacyclic_polymul(self.raw, prod.as_mut_ptr() as *mut c_void, prod.len(),
                    lhs.as_ptr() as *const c_void, lhs.len(), rhs.as_ptr() as *const c_void, rhs.len()) }, "negacyclic_polymul")
            }
            /// New: `prod.len() / ntt_size()` products in one call.
            pub fn negacyclic_polymul_batch(&self, prod: &mut [$V], lhs: &[$V], rhs: &[$V]) {
                assert_eq!(prod.len(), lhs.len()); assert_eq!(prod.len(), rhs.len());
                assert_eq!(prod.len() % self.ntt_size(), 0);
                check(unsafe { ffi::ntt_b200_native_negacyclic_polymul_batch(self.raw, prod.as_mut_ptr() as *mut c_void,
                    lhs.as_ptr() as *const c_void, rhs.as_ptr() as *const c_void, prod.len() / self.ntt_size()) }, "negacyclic_polymul_batch")
            }
        }
        impl Drop for $ty { fn drop(&mut self) { unsafe { ffi::ntt_b200_native_free(self.raw) } } }
    };
}

pub mod native32 {
    use super::*;
    native_plan!(native32, Plan32, 0, u32, u32, prime32, [(mod_p0, ntt_0, 0), (mod_p1, ntt_1, 1), (mod_p2, ntt_2, 2)], false);
    native_plan!(native32, Plan52, 1, u32, u64, prime64, [(mod_p0, ntt_0, 0), (mod_p1, ntt_1, 1)], false);
}
pub mod native64 {
    use super::*;
    native_plan!(native64, Plan32, 2, u64, u32, prime32, [(mod_p0, ntt_0, 0), (mod_p1, ntt_1, 1), (mod_p2, ntt_2, 2), (mod_p3, ntt_3, 3), (mod_p4, ntt_4, 4)], false);
    native_plan!(native64, Plan52, 3, u64, u64, prime64, [(mod_p0, ntt_0, 0), (mod_p1, ntt_1, 1), (mod_p2, ntt_2, 2)], false);
}
pub mod native128 {
    use super::*;
    native_plan!(native128, Plan32, 4, u128, u32, prime32,
        [(mod_p0, ntt_0, 0), (mod_p1, ntt_1, 1), (mod_p2, ntt_2, 2), (mod_p3, ntt_3, 3), (mod_p4, ntt_4, 4), (mod_p5, ntt_5, 5), (mod_p6, ntt_6, 6), (mod_p7, ntt_7, 7), (mod_p8, ntt_8, 8), (mod_p9, ntt_9, 9)], false);
}
pub mod native_binary32 {
    use super::*;
    native_plan!(native_binary32, Plan32, 5, u32, u32, prime32, [(mod_p0, ntt_0, 0), (mod_p1, ntt_1, 1)], true);
    native_plan!(native_binary32, Plan52, 6, u32, u64, prime64, [(mod_p0, ntt_0, 0)], true);
}
pub mod native_binary64 {
    use super::*;
    native_plan!(native_binary64, Plan32, 7, u64, u32, prime32, [(mod_p0, ntt_0, 0), (mod_p1, ntt_1, 1), (mod_p2, ntt_2, 2)], true);
    native_plan!(native_binary64, Plan52, 8, u64, u64, prime64, [(mod_p0, ntt_0, 0), (mod_p1, ntt_1, 1)], true);
}
pub mod native_binary128 {
    use super::*;
    native_plan!(native_binary128, Plan32, 9, u128, u32, prime32, [(mod_p0, ntt_0, 0), (mod_p1, ntt_1, 1), (mod_p2, ntt_2, 2), (mod_p3, ntt_3, 3), (mod_p4, ntt_4, 4)], true);
}
