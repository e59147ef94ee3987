"""product::Plan (reference: tfhe-ntt/src/product.rs:139-967)."""
import ctypes as C

import numpy as np

from . import _binding as B


class FwdMode:
    """product.rs:10-14.  Bounded(b) promises |coefficients| < b; the residues are those of Generic."""
    Generic = ("Generic",)

    @staticmethod
    def Bounded(bound):
        return ("Bounded", bound)


class InvMode:
    """product.rs:16-20"""
    Replace = 0
    Accumulate = 1


class Plan:
    def __init__(self, handle):
        self._h = handle
        self._L = B.lib()

    @classmethod
    def try_new(cls, polynomial_size, modulus, factors):
        f = (C.c_uint64 * len(factors))(*factors)
        out = C.c_void_p()
        st = B.lib().ntt_b200_product_try_new(polynomial_size, modulus, f, len(factors), C.byref(out))
        if st == B.NONE:
            return None
        B.check(st, "try_new")
        return cls(out.value)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.ntt_b200_product_free(self._h)
        self._h = None

    def ntt_size(self):
        return self._L.ntt_b200_product_ntt_size(self._h)

    def modulus(self):
        return self._L.ntt_b200_product_modulus(self._h)

    def ntt_domain_len(self):
        return self._L.ntt_b200_product_ntt_domain_len(self._h)

    def fwd(self, ntt, standard, mode=FwdMode.Generic):
        B.check(self._L.ntt_b200_product_fwd(self._h, B.host_ptr(ntt, np.uint64, True), ntt.size,
                                             B.host_ptr(standard, np.uint64), standard.size), "in fwd")

    def inv(self, standard, ntt, mode=InvMode.Replace):
        B.check(self._L.ntt_b200_product_inv(self._h, B.host_ptr(standard, np.uint64, True), standard.size,
                                             B.host_ptr(ntt, np.uint64, True), ntt.size, int(mode)), "in inv")

    def normalize(self, values):
        B.check(self._L.ntt_b200_product_normalize(self._h, B.host_ptr(values, np.uint64, True), values.size),
                "in normalize")

    def mul_assign_normalize(self, lhs, rhs):
        B.check(self._L.ntt_b200_product_mul_assign_normalize(self._h, B.host_ptr(lhs, np.uint64, True), lhs.size,
                                                              B.host_ptr(rhs, np.uint64), rhs.size),
                "in mul_assign_normalize")

    def mul_accumulate(self, acc, lhs, rhs):
        B.check(self._L.ntt_b200_product_mul_accumulate(self._h, B.host_ptr(acc, np.uint64, True), acc.size,
                                                        B.host_ptr(lhs, np.uint64), lhs.size,
                                                        B.host_ptr(rhs, np.uint64), rhs.size), "in mul_accumulate")

    def fwd_device(self, ntt, standard, batch, stream=None):
        B.check(self._L.ntt_b200_product_fwd_device(self._h, B.dev_ptr(ntt), B.dev_ptr(standard), batch,
                                                    B.stream_ptr(stream)))

    def inv_device(self, standard, ntt, batch, mode=InvMode.Replace, stream=None):
        B.check(self._L.ntt_b200_product_inv_device(self._h, B.dev_ptr(standard), B.dev_ptr(ntt), batch, int(mode),
                                                    B.stream_ptr(stream)))
