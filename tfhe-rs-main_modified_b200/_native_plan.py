"""Shared implementation of the CRT plans (reference: native32.rs, native64.rs, native128.rs,
native_binary32.rs, native_binary64.rs, native_binary128.rs)."""
import ctypes as C

import numpy as np

from . import _binding as B
from . import prime32, prime64

U128 = np.dtype([("lo", np.uint64), ("hi", np.uint64)])  # little-endian u128


class NativePlanBase:
    _kind = None
    _binary = False

    def __init__(self, handle):
        self._h = handle
        self._L = B.lib()
        k = self._kind
        self.num_primes = self._L.ntt_b200_native_num_primes(k)
        self.residue_bytes = self._L.ntt_b200_native_residue_bytes(k)
        self.value_bytes = self._L.ntt_b200_native_value_bytes(k)
        self.residue_dtype = np.uint32 if self.residue_bytes == 4 else np.uint64
        self.value_dtype = {4: np.dtype(np.uint32), 8: np.dtype(np.uint64), 16: U128}[self.value_bytes]

    @classmethod
    def try_new(cls, n):
        out = C.c_void_p()
        st = B.lib().ntt_b200_native_try_new(cls._kind, n, C.byref(out))
        if st == B.NONE:
            return None
        B.check(st, "try_new")
        return cls(out.value)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.ntt_b200_native_free(self._h)
        self._h = None

    def ntt_size(self):
        return self._L.ntt_b200_native_ntt_size(self._h)

    def ntt_i(self, i):
        """ntt_0() .. ntt_9(): the i-th per-prime plan (borrowed)."""
        h = self._L.ntt_b200_native_ntt_i(self._h, i)
        if not h:
            raise IndexError(i)
        cls = prime32.Plan if self.residue_bytes == 4 else prime64.Plan
        return cls(h, owner=self)

    def __getattr__(self, name):
        if name.startswith("ntt_") and name[4:].isdigit():
            i = int(name[4:])
            return lambda: self.ntt_i(i)
        raise AttributeError(name)

    def _vptr(self, a, writable=False):
        if a.dtype.itemsize != self.value_bytes and not (self.value_bytes == 16 and a.dtype == np.uint64):
            raise TypeError("value array has the wrong element type")
        if not a.flags["C_CONTIGUOUS"]:
            raise TypeError("value array must be C-contiguous")
        return a.ctypes.data

    def _vlen(self, a):
        return a.size // 2 if (self.value_bytes == 16 and a.dtype == np.uint64) else a.size

    def _res(self, residues, writable=True):
        if len(residues) != self.num_primes:
            raise TypeError("expected %d residue buffers" % self.num_primes)
        arr = (C.c_void_p * self.num_primes)()
        for j, r in enumerate(residues):
            arr[j] = B.host_ptr(r, self.residue_dtype, writable)
        return arr

    def fwd(self, value, *residues):
        B.check(self._L.ntt_b200_native_fwd(self._h, self._vptr(value), self._vlen(value), self._res(residues), 0),
                "in fwd")

    def fwd_binary(self, value, *residues):
        if not self._binary:
            raise AttributeError("fwd_binary exists on the native_binary plans only")
        B.check(self._L.ntt_b200_native_fwd(self._h, self._vptr(value), self._vlen(value), self._res(residues), 1),
                "in fwd_binary")

    def inv(self, value, *residues):
        B.check(self._L.ntt_b200_native_inv(self._h, self._vptr(value, True), self._vlen(value), self._res(residues)),
                "in inv")

    def negacyclic_polymul(self, prod, lhs, rhs):
        B.check(self._L.ntt_b200_native_negacyclic_polymul(
            self._h, self._vptr(prod, True), self._vlen(prod), self._vptr(lhs), self._vlen(lhs),
            self._vptr(rhs), self._vlen(rhs)), "in negacyclic_polymul")

    def negacyclic_polymul_batch(self, prod, lhs, rhs):
        n = self.ntt_size()
        batch = self._vlen(prod) // n
        if self._vlen(prod) != batch * n or self._vlen(lhs) != batch * n or self._vlen(rhs) != batch * n:
            raise AssertionError("length mismatch in negacyclic_polymul_batch")
        B.check(self._L.ntt_b200_native_negacyclic_polymul_batch(
            self._h, self._vptr(prod, True), self._vptr(lhs), self._vptr(rhs), batch))

    def negacyclic_polymul_device(self, prod, lhs, rhs, batch=None, stream=None):
        if batch is None:
            batch = B.dev_numel(prod, self.value_bytes) // self.ntt_size()
        B.check(self._L.ntt_b200_native_negacyclic_polymul_device(
            self._h, B.dev_ptr(prod), B.dev_ptr(lhs), B.dev_ptr(rhs), batch, B.stream_ptr(stream)))

    def fwd_device(self, value, residues, batch, binary=False, stream=None):
        arr = (C.c_void_p * self.num_primes)(*[B.dev_ptr(r) for r in residues])
        B.check(self._L.ntt_b200_native_fwd_device(self._h, B.dev_ptr(value), arr, batch, int(binary),
                                                   B.stream_ptr(stream)))

    def inv_device(self, value, residues, batch, stream=None):
        arr = (C.c_void_p * self.num_primes)(*[B.dev_ptr(r) for r in residues])
        B.check(self._L.ntt_b200_native_inv_device(self._h, B.dev_ptr(value), arr, batch, B.stream_ptr(stream)))


def make(kind, binary=False, doc=""):
    return type("Plan", (NativePlanBase,), {"_kind": kind, "_binary": binary, "__doc__": doc})
