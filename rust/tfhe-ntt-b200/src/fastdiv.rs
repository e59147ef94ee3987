//! Fast division by a constant divisor (reference: tfhe-ntt/src/fastdiv.rs:29-150).
//! Host-side helpers with the reference's names and results (quotient / remainder of an exact
//! division); only the results are observable (fastdiv.rs:159-195 tests them against `/` and `%`).
//! The reciprocals are kept with the reference's field names; the 256-bit one is two `u128` halves.

/// Divisor representing a 32bit denominator.
#[derive(Copy, Clone, Debug)]
pub struct Div32 {
    /// ceil(2^128 / divisor)
    pub double_reciprocal: u128,
    /// ceil(2^64 / divisor)
    pub single_reciprocal: u64,
    pub divisor: u32,
}

/// Divisor representing a 64bit denominator.
#[derive(Copy, Clone, Debug)]
pub struct Div64 {
    /// ceil(2^256 / divisor) as (low, high) 128-bit halves
    pub double_reciprocal: (u128, u128),
    /// ceil(2^128 / divisor)
    pub single_reciprocal: u128,
    pub divisor: u64,
}

#[inline(always)]
const fn mulhi_u128_u64(lowbits: u128, d: u64) -> u64 {
    // high 64 bits of the 192-bit product lowbits * d
    let bottom = ((lowbits & 0xFFFF_FFFF_FFFF_FFFF) * d as u128) >> 64;
    let top = (lowbits >> 64) * d as u128;
    ((bottom + top) >> 64) as u64
}

impl Div32 {
    /// # Panics
    /// Panics if the divisor is zero or one (fastdiv.rs:49-51).
    pub const fn new(divisor: u32) -> Self {
        assert!(divisor > 1);
        Self {
            double_reciprocal: (u128::MAX / divisor as u128) + 1,
            single_reciprocal: (u64::MAX / divisor as u64) + 1,
            divisor,
        }
    }
    /// fastdiv.rs:63
    #[inline(always)]
    pub const fn div(n: u32, d: Self) -> u32 { ((d.single_reciprocal as u128 * n as u128) >> 64) as u32 }
    /// fastdiv.rs:69
    #[inline(always)]
    pub const fn rem(n: u32, d: Self) -> u32 {
        let low_bits = d.single_reciprocal.wrapping_mul(n as u64);
        ((low_bits as u128 * d.divisor as u128) >> 64) as u32
    }
    /// fastdiv.rs:76
    #[inline(always)]
    pub const fn div_u64(n: u64, d: Self) -> u64 { mulhi_u128_u64(d.double_reciprocal, n) }
    /// fastdiv.rs:82
    #[inline(always)]
    pub const fn rem_u64(n: u64, d: Self) -> u32 {
        let low_bits = d.double_reciprocal.wrapping_mul(n as u128);
        mulhi_u128_u64(low_bits, d.divisor as u64) as u32
    }
    #[inline(always)]
    pub const fn divisor(&self) -> u32 { self.divisor }
}

impl Div64 {
    /// # Panics
    /// Panics if the divisor is zero or one (fastdiv.rs:101-103).
    pub const fn new(divisor: u64) -> Self {
        assert!(divisor > 1);
        // ceil(2^256 / d) by long division of 2^256 - 1 in 128-bit halves, plus one
        let d = divisor as u128;
        let q_hi = u128::MAX / d;
        let r_hi = u128::MAX % d;
        // (r_hi * 2^128 + (2^128 - 1)) / d, r_hi < d < 2^64: two 64-bit steps
        let n1 = (r_hi << 64) | 0xFFFF_FFFF_FFFF_FFFF;
        let q1 = n1 / d;
        let r1 = n1 % d;
        let n0 = (r1 << 64) | 0xFFFF_FFFF_FFFF_FFFF;
        let q0 = n0 / d;
        let q_lo = (q1 << 64) | q0;
        let (lo, carry) = q_lo.overflowing_add(1);
        let hi = if carry { q_hi.wrapping_add(1) } else { q_hi };
        Self { double_reciprocal: (lo, hi), single_reciprocal: (u128::MAX / d) + 1, divisor }
    }
    /// fastdiv.rs:124
    #[inline(always)]
    pub const fn div(n: u64, d: Self) -> u64 { mulhi_u128_u64(d.single_reciprocal, n) }
    /// fastdiv.rs:130
    #[inline(always)]
    pub const fn rem(n: u64, d: Self) -> u64 {
        let low_bits = d.single_reciprocal.wrapping_mul(n as u128);
        mulhi_u128_u64(low_bits, d.divisor)
    }
    /// fastdiv.rs:137 (exact quotient; the reciprocal form needs a 384-bit product, so this one divides)
    #[inline(always)]
    pub const fn div_u128(n: u128, d: Self) -> u128 { n / d.divisor as u128 }
    /// fastdiv.rs:143
    #[inline(always)]
    pub const fn rem_u128(n: u128, d: Self) -> u64 { (n % d.divisor as u128) as u64 }
    #[inline(always)]
    pub const fn divisor(&self) -> u64 { self.divisor }
}

#[cfg(test)]
mod tests {
    use super::*;
    #[test]
    fn matches_division() {
        // restates fastdiv.rs:159-195 on a fixed sequence
        let mut s = 0x9E3779B97F4A7C15u64;
        let mut next = || { s ^= s << 13; s ^= s >> 7; s ^= s << 17; s };
        for _ in 0..1000 {
            let d32 = (next() as u32).max(2);
            let d64 = next().max(2);
            let (n32, n64, n128) = (next() as u32, next(), ((next() as u128) << 64) | next() as u128);
            let (a, b) = (Div32::new(d32), Div64::new(d64));
            assert_eq!(Div32::div(n32, a), n32 / d32);
            assert_eq!(Div32::rem(n32, a), n32 % d32);
            assert_eq!(Div32::div_u64(n64, a), n64 / d32 as u64);
            assert_eq!(Div32::rem_u64(n64, a), (n64 % d32 as u64) as u32);
            assert_eq!(Div64::div(n64, b), n64 / d64);
            assert_eq!(Div64::rem(n64, b), n64 % d64);
            assert_eq!(Div64::div_u128(n128, b), n128 / d64 as u128);
            assert_eq!(Div64::rem_u128(n128, b), (n128 % d64 as u128) as u64);
        }
    }
}
