"""prime::{is_prime64, largest_prime_in_arithmetic_progression64} (reference: prime.rs:76-186)."""
import ctypes as C

from . import _binding as B


def is_prime64(n):
    return bool(B.lib().ntt_b200_is_prime64(n))


def largest_prime_in_arithmetic_progression64(factor, offset, lo, hi):
    out = C.c_uint64()
    ok = B.lib().ntt_b200_largest_prime_in_arithmetic_progression64(factor, offset, lo, hi, C.byref(out))
    return out.value if ok else None
