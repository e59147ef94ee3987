"""Shared implementation of prime32.Plan and prime64.Plan (reference: prime32.rs:632-1016,
prime64.rs:245-1223)."""
import ctypes as C

import numpy as np

from . import _binding as B


class PrimePlanBase:
    _sfx = None       # "32" / "64"
    _dtype = None     # np.uint32 / np.uint64
    _min_n = None

    def __init__(self, handle, owner=None):
        self._h = handle
        self._owner = owner  # keeps a parent native plan alive for borrowed handles
        self._L = B.lib()

    def _f(self, name):
        return getattr(self._L, "ntt_b200_plan%s_%s" % (self._sfx, name))

    @classmethod
    def try_new(cls, polynomial_size, modulus):
        """Plan::try_new -> Option<Plan>: returns None where the reference does."""
        if not (0 <= modulus < (1 << int(cls._sfx))):
            raise OverflowError("modulus does not fit u%s" % cls._sfx)
        out = C.c_void_p()
        st = getattr(B.lib(), "ntt_b200_plan%s_try_new" % cls._sfx)(polynomial_size, modulus, C.byref(out))
        if st == B.NONE:
            return None
        B.check(st, "try_new")
        return cls(out.value)

    def clone(self):
        out = C.c_void_p()
        B.check(self._f("clone")(self._h, C.byref(out)))
        return type(self)(out.value)

    def __del__(self):
        if getattr(self, "_h", None) and self._owner is None:
            self._f("free")(self._h)
        self._h = None

    def __repr__(self):  # reference Debug prints only size and modulus (prime64.rs:263-270)
        return "Plan { ntt_size: %d, modulus: %d }" % (self.ntt_size(), self.modulus())

    # accessors
    def ntt_size(self):
        return self._f("ntt_size")(self._h)

    def modulus(self):
        return self._f("modulus")(self._h)

    def can_use_fast_reduction_code(self):
        return bool(self._f("can_use_fast_reduction_code")(self._h))

    def device(self):
        return self._f("device")(self._h)

    # per-polynomial host calls (in place, like &mut [T])
    def fwd(self, buf):
        B.check(self._f("fwd")(self._h, B.host_ptr(buf, self._dtype, True), buf.size), "in fwd")

    def inv(self, buf):
        B.check(self._f("inv")(self._h, B.host_ptr(buf, self._dtype, True), buf.size), "in inv")

    def normalize(self, values):
        B.check(self._f("normalize")(self._h, B.host_ptr(values, self._dtype, True), values.size))

    def mul_assign_normalize(self, lhs, rhs):
        B.check(self._f("mul_assign_normalize")(self._h, B.host_ptr(lhs, self._dtype, True), lhs.size,
                                                B.host_ptr(rhs, self._dtype), rhs.size))

    def mul_accumulate(self, acc, lhs, rhs):
        B.check(self._f("mul_accumulate")(self._h, B.host_ptr(acc, self._dtype, True), acc.size,
                                          B.host_ptr(lhs, self._dtype), lhs.size,
                                          B.host_ptr(rhs, self._dtype), rhs.size))

    # batched host calls
    def _batch(self, buf):
        n = self.ntt_size()
        if buf.size % n:
            raise AssertionError("length mismatch: batch buffer is not a multiple of ntt_size")
        return buf.size // n

    def fwd_batch(self, buf):
        B.check(self._f("fwd_batch")(self._h, B.host_ptr(buf, self._dtype, True), self._batch(buf)))

    def inv_batch(self, buf):
        B.check(self._f("inv_batch")(self._h, B.host_ptr(buf, self._dtype, True), self._batch(buf)))

    @classmethod
    def try_new_on_devices(cls, polynomial_size, modulus, devices):
        """One plan per GPU of `devices` for the same (n, p) (None where the reference returns None);
        restores the calling thread's current device."""
        plans = []
        try:
            for d in devices:
                B.set_device(d)
                pl = cls.try_new(polynomial_size, modulus)
                if pl is None:
                    return None
                plans.append(pl)
        finally:
            if devices:
                B.set_device(devices[0])
        return plans

    @staticmethod
    def _multi(plans, name, buf):
        first = plans[0]
        arr = (C.c_void_p * len(plans))(*[pl._h for pl in plans])
        B.check(first._f(name)(arr, len(plans), B.host_ptr(buf, first._dtype, True), first._batch(buf)))

    @classmethod
    def fwd_batch_multi_gpu(cls, plans, buf):
        """fwd over a host batch split in contiguous slices over the GPUs the plans live on."""
        cls._multi(plans, "fwd_batch_multi_gpu", buf)

    @classmethod
    def inv_batch_multi_gpu(cls, plans, buf):
        cls._multi(plans, "inv_batch_multi_gpu", buf)

    def fwd_mac_inv_batch(self, out, lhs, rhs, acc=None):
        """out = inv(acc + fwd(lhs) * rhs) on host arrays; rhs / acc may hold fewer polynomials
        than lhs (reused cyclically).  out may be lhs."""
        n = self.ntt_size()
        batch = self._batch(lhs)
        if out.size != lhs.size:
            raise AssertionError("length mismatch: out and lhs differ")
        B.check(self._f("fwd_mac_inv_batch")(
            self._h, B.host_ptr(out, self._dtype, True), B.host_ptr(lhs, self._dtype),
            B.host_ptr(rhs, self._dtype), rhs.size // n,
            None if acc is None else B.host_ptr(acc, self._dtype), 0 if acc is None else acc.size // n, batch),
            "in fwd_mac_inv_batch")

    # device-resident calls
    def _eb(self):
        return np.dtype(self._dtype).itemsize

    def fwd_device(self, dev, batch=None, stream=None):
        if batch is None:
            batch = B.dev_numel(dev, self._eb()) // self.ntt_size()
        B.check(self._f("fwd_device")(self._h, B.dev_ptr(dev), batch, B.stream_ptr(stream)), "in fwd_device")

    def inv_device(self, dev, batch=None, stream=None):
        if batch is None:
            batch = B.dev_numel(dev, self._eb()) // self.ntt_size()
        B.check(self._f("inv_device")(self._h, B.dev_ptr(dev), batch, B.stream_ptr(stream)), "in inv_device")

    def normalize_device(self, dev, length=None, stream=None):
        if length is None:
            length = B.dev_numel(dev, self._eb())
        B.check(self._f("normalize_device")(self._h, B.dev_ptr(dev), length, B.stream_ptr(stream)))

    def mul_assign_normalize_device(self, lhs, rhs, length=None, rhs_len=None, stream=None):
        length = B.dev_numel(lhs, self._eb()) if length is None else length
        rhs_len = B.dev_numel(rhs, self._eb()) if rhs_len is None else rhs_len
        B.check(self._f("mul_assign_normalize_device")(self._h, B.dev_ptr(lhs), length, B.dev_ptr(rhs),
                                                       rhs_len, B.stream_ptr(stream)))

    def mul_accumulate_device(self, acc, lhs, rhs, length=None, lhs_len=None, rhs_len=None, stream=None):
        length = B.dev_numel(acc, self._eb()) if length is None else length
        lhs_len = B.dev_numel(lhs, self._eb()) if lhs_len is None else lhs_len
        rhs_len = B.dev_numel(rhs, self._eb()) if rhs_len is None else rhs_len
        B.check(self._f("mul_accumulate_device")(self._h, B.dev_ptr(acc), length, B.dev_ptr(lhs), lhs_len,
                                                 B.dev_ptr(rhs), rhs_len, B.stream_ptr(stream)))

    def ext_product_device(self, out, inp, ggsw, rows, cols, batch=None, stream=None):
        """out[b][c] = inv(sum_r fwd(inp[b][r]) * ggsw[r][c]) -- the NTT-PBS external-product core."""
        n = self.ntt_size()
        if batch is None:
            batch = B.dev_numel(inp, self._eb()) // (n * rows)
        B.check(self._f("ext_product_device")(self._h, B.dev_ptr(out), B.dev_ptr(inp), B.dev_ptr(ggsw), rows, cols,
                                              batch, B.stream_ptr(stream)), "in ext_product_device")

    def fwd_mac_inv_device(self, out, lhs, rhs, acc=None, batch=None, rhs_polys=None, acc_polys=None,
                           stream=None):
        n = self.ntt_size()
        batch = B.dev_numel(out, self._eb()) // n if batch is None else batch
        rhs_polys = B.dev_numel(rhs, self._eb()) // n if rhs_polys is None else rhs_polys
        if acc is not None and acc_polys is None:
            acc_polys = B.dev_numel(acc, self._eb()) // n
        B.check(self._f("fwd_mac_inv_device")(self._h, B.dev_ptr(out), B.dev_ptr(lhs), B.dev_ptr(rhs), rhs_polys,
                                              B.dev_ptr(acc), acc_polys or 0, batch, B.stream_ptr(stream)))
