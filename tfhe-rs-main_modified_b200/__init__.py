"""tfhe_ntt_b200 -- host-side mirror of the `tfhe-ntt` API over the B200 C ABI.

The reference is a Rust crate (/root/reference/tfhe-ntt/src/lib.rs:83-116); this image has no
Rust toolchain, so the tested host layer is this ctypes binding of libtfhe_ntt_b200.so (the Rust
crate source that binds the same symbols lives in rust/ and is described in INTEGRATION.md).
Module, type and method names follow the reference:

    prime32.Plan / prime64.Plan    try_new, ntt_size, modulus, fwd, inv, normalize,
                                   mul_assign_normalize, mul_accumulate,
                                   can_use_fast_reduction_code (+ use_ifma on prime64)
    native32 / native64 / native128 / native_binary32 / native_binary64 / native_binary128
                                   Plan32 / Plan52: try_new, ntt_size, ntt_0().., fwd, fwd_binary,
                                   inv, negacyclic_polymul
    product.Plan (FwdMode, InvMode)   try_new, ntt_size, modulus, ntt_domain_len, fwd, inv,
                                   normalize, mul_assign_normalize, mul_accumulate
    prime.is_prime64, prime.largest_prime_in_arithmetic_progression64
    ntt64.Ntt64View                tfhe's wrapper over prime64.Plan (forward*, add_backward*)
    ntt64_pbs                      NttLweBootstrapKey, convert_standard_lwe_bootstrap_key_to_ntt64,
                                   blind_rotate_ntt64[_bnf]_assign,
                                   programmable_bootstrap_ntt64[_bnf]_lwe_ciphertext
    custum_radix                   the fork's recursive cyclic u32 transforms: fft_/ifft_{radix2,radix4,split_radix}_recursive,
                                   fft_*_recursive_mut / ifft_radix{2,4}_recursive_mut with MultStats (+ fft_batch / ifft_batch / *_device)

Host calls take numpy arrays and work in place exactly like the reference's `&mut [T]` slices.
New, alongside: `*_batch` (host arrays holding many polynomials) and `*_device` (device pointers
or anything with .data_ptr(), e.g. torch tensors, plus a CUDA stream handle).

There is no CPU fallback: importing works without a GPU (so the symbol table can be checked),
but every compute call needs the CUDA library and a device and raises otherwise.

The directory name has hyphens, so import it through the repo-root shim: `import tfhe_ntt_b200`.
"""
from ._binding import (  # noqa: F401
    NttB200Error, lib, library_path, last_error, device_count, set_device,
)
from . import prime32, prime64, prime  # noqa: F401
from . import native32, native64, native128  # noqa: F401
from . import native_binary32, native_binary64, native_binary128  # noqa: F401
from . import product  # noqa: F401
from . import ntt64  # noqa: F401
from . import ntt64_pbs  # noqa: F401
from . import custum_radix  # noqa: F401
from . import sharding  # noqa: F401
