//! reference: tfhe-ntt/src/prime.rs:76-186
pub fn is_prime64(n: u64) -> bool { unsafe { crate::ffi::ntt_b200_is_prime64(n) != 0 } }
pub fn largest_prime_in_arithmetic_progression64(factor: u64, offset: u64, lo: u64, hi: u64) -> Option<u64> {
    let mut out = 0u64;
    let ok = unsafe { crate::ffi::ntt_b200_largest_prime_in_arithmetic_progression64(factor, offset, lo, hi, &mut out) };
    (ok != 0).then_some(out)
}
