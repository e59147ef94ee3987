//! `custum_radix` (reference: tfhe-ntt/src/custum_radix/mod.rs:1-22): the fork's recursive cyclic
//! transforms of `u32` vectors over a caller-built table `twiddles[k] = root^k mod p`, natural order
//! in and out.  Same names and argument order; on such a table the three forward routines are one
//! function and one CUDA schedule serves them (csrc/capi_custum_radix.cu); the `_mut` routines of
//! fwd_1.rs also return the fork's `MultStats` counters, reproduced by a dedicated kernel.
use crate::ffi::{self, check};
use core::ffi::c_int;

const RADIX2: c_int = 0;
const RADIX4: c_int = 1;
const SPLIT_RADIX: c_int = 2;
#[allow(dead_code)]
const RADIX4_MUT: c_int = 3;

#[track_caller]
fn fft(kind: c_int, a: &mut [u32], twiddles: &[u32], p: u32) {
    check(unsafe { ffi::ntt_b200_custum_radix_fft(kind, a.as_mut_ptr(), a.len(), twiddles.as_ptr(), twiddles.len(), p) }, "custum_radix fft")
}
#[track_caller]
fn ifft(kind: c_int, a: &mut [u32], inv_twiddles: &[u32], p: u32, n_inv: u32, top: bool) {
    check(unsafe { ffi::ntt_b200_custum_radix_ifft(kind, a.as_mut_ptr(), a.len(), inv_twiddles.as_ptr(), inv_twiddles.len(), p, n_inv, top as c_int) }, "custum_radix ifft")
}

/// fwd.rs:105
pub fn fft_radix4_recursive(a: &mut [u32], twiddles: &[u32], p: u32) { fft(RADIX4, a, twiddles, p) }
/// fwd.rs:170
pub fn fft_radix2_recursive(a: &mut [u32], twiddles: &[u32], p: u32) { fft(RADIX2, a, twiddles, p) }
/// fwd.rs:207
pub fn fft_split_radix_recursive(a: &mut [u32], tw: &[u32], p: u32) { fft(SPLIT_RADIX, a, tw, p) }
/// inv.rs:106 (its size-2 base halves: the result carries 1/2 when log2 n is odd)
pub fn ifft_radix4_recursive(a: &mut [u32], inv_twiddles: &[u32], p: u32, n_inv: u32, top: bool) { ifft(RADIX4, a, inv_twiddles, p, n_inv, top) }
/// inv.rs:178
pub fn ifft_radix2_recursive(a: &mut [u32], inv_twiddles: &[u32], p: u32, n_inv: u32, top: bool) { ifft(RADIX2, a, inv_twiddles, p, n_inv, top) }
/// inv.rs:232
pub fn ifft_split_radix_recursive(a: &mut [u32], inv_tw: &[u32], p: u32, n_inv: u32, top: bool) { ifft(SPLIT_RADIX, a, inv_tw, p, n_inv, top) }

pub mod fwd_1 {
    //! fwd_1.rs: the `_mut` routines with the fork's multiplication counters (one vector per call, n <= 4096:
    //! the counting kernel keeps every level of the transform in shared memory).
    use super::*;

    /// fwd_1.rs:3-7
    #[derive(Debug, Clone, Default)]
    pub struct MultStats {
        pub nonzero_mults: usize, // nonzero * nonzero
        pub skipped_mults: usize, // multiplications with zero
    }

    #[track_caller]
    fn fft_mut(kind: c_int, a: &mut [u32], twiddles: &[u32], p: u32, stats: &mut MultStats) {
        let mut raw = [stats.nonzero_mults as u64, stats.skipped_mults as u64];
        check(unsafe { ffi::ntt_b200_custum_radix_fft_mut(kind, a.as_mut_ptr(), a.len(), twiddles.as_ptr(), twiddles.len(), p, raw.as_mut_ptr()) }, "custum_radix fft_mut");
        stats.nonzero_mults = raw[0] as usize;
        stats.skipped_mults = raw[1] as usize;
    }
    /// fwd_1.rs:102
    pub fn fft_radix4_recursive_mut(a: &mut [u32], twiddles: &[u32], p: u32, stats: &mut MultStats) { fft_mut(RADIX4, a, twiddles, p, stats) }
    /// fwd_1.rs:190
    pub fn fft_radix2_recursive_mut(a: &mut [u32], twiddles: &[u32], p: u32, stats: &mut MultStats) { fft_mut(RADIX2, a, twiddles, p, stats) }
    /// fwd_1.rs:232
    pub fn fft_split_radix_recursive_mut(a: &mut [u32], tw: &[u32], p: u32, stats: &mut MultStats) { fft_mut(SPLIT_RADIX, a, tw, p, stats) }
    /// fwd_1.rs:296: the bases scale when `top`, nothing is halved
    #[track_caller]
    pub fn ifft_radix4_recursive_mut(a: &mut [u32], inv_twiddles: &[u32], p: u32, n_inv: u32, top: bool, stats: &mut MultStats) {
        let mut raw = [stats.nonzero_mults as u64, stats.skipped_mults as u64];
        check(unsafe { ffi::ntt_b200_custum_radix_ifft_radix4_mut(a.as_mut_ptr(), a.len(), inv_twiddles.as_ptr(), inv_twiddles.len(), p, n_inv, top as c_int, raw.as_mut_ptr()) }, "ifft_radix4_recursive_mut");
        stats.nonzero_mults = raw[0] as usize;
        stats.skipped_mults = raw[1] as usize;
    }
    /// fwd_1.rs:381 (takes no counters in the reference either)
    pub fn ifft_radix2_recursive_mut(a: &mut [u32], inv_twiddles: &[u32], p: u32, n_inv: u32, top: bool) { ifft(RADIX2, a, inv_twiddles, p, n_inv, top) }
}
pub use fwd_1::{fft_radix2_recursive_mut, fft_radix4_recursive_mut, fft_split_radix_recursive_mut};  // mod.rs:17-21

/// New: `batch` contiguous vectors of `n` elements in host memory.
#[track_caller]
pub fn fft_batch(a: &mut [u32], n: usize, twiddles: &[u32], p: u32) {
    assert!(n > 0 && a.len() % n == 0);
    check(unsafe { ffi::ntt_b200_custum_radix_fft_batch(RADIX2, a.as_mut_ptr(), n, a.len() / n, twiddles.as_ptr(), twiddles.len(), p) }, "custum_radix fft_batch")
}
#[track_caller]
pub fn ifft_batch(a: &mut [u32], n: usize, inv_twiddles: &[u32], p: u32, n_inv: u32, top: bool) {
    assert!(n > 0 && a.len() % n == 0);
    check(unsafe { ffi::ntt_b200_custum_radix_ifft_batch(RADIX2, a.as_mut_ptr(), n, a.len() / n, inv_twiddles.as_ptr(), inv_twiddles.len(), p, n_inv, top as c_int) }, "custum_radix ifft_batch")
}
