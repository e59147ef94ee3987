//! `tfhe-ntt`-compatible API over the B200 CUDA engine (reference: tfhe-ntt/src/lib.rs:83-116).
//! Module tree, type names and method signatures are the reference's; batched and
//! device-resident entry points are added alongside.  There is no CPU fallback.
pub mod ffi;
pub mod fastdiv;
pub mod prime;
mod primes;
pub use primes::{primes32, primes52};
pub mod prime32;
pub mod prime64;
pub mod product;
pub mod ntt64_pbs;
pub mod custum_radix;
mod native;
pub use native::{native128, native32, native64, native_binary128, native_binary32, native_binary64};
