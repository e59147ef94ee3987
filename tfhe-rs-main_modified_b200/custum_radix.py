"""custum_radix -- the fork's recursive cyclic u32 transforms (reference: tfhe-ntt/src/custum_radix/mod.rs:1-22,
fwd.rs, inv.rs, fwd_1.rs), same function names and argument order.

`a` is a C-contiguous numpy uint32 vector transformed in place, `twiddles[k] = root^k mod p`.  On such a table
the radix-2, radix-4 and split-radix routines are one function, so one CUDA schedule serves all of them; the
inverse routines differ by the constant their recursion bases apply, which the library restates
(csrc/capi_custum_radix.cu).  The `_mut` routines of fwd_1.rs return the same values plus the fork's MultStats
counters, which a dedicated kernel reproduces from the levels of the transform (one vector per call, n <= 4096).
"""
import numpy as np

from . import _binding as B

RADIX2, RADIX4, SPLIT_RADIX, RADIX4_MUT = 0, 1, 2, 3


def _fft(kind, a, twiddles, p):
    B.check(B.lib().ntt_b200_custum_radix_fft(kind, B.host_ptr(a, np.uint32, True), a.size,
                                              B.host_ptr(twiddles, np.uint32), twiddles.size, p), "in custum_radix fft")


def _ifft(kind, a, inv_twiddles, p, n_inv, top):
    B.check(B.lib().ntt_b200_custum_radix_ifft(kind, B.host_ptr(a, np.uint32, True), a.size,
                                               B.host_ptr(inv_twiddles, np.uint32), inv_twiddles.size, p, n_inv,
                                               1 if top else 0), "in custum_radix ifft")


def fft_radix2_recursive(a, twiddles, p):
    """fwd.rs:170-205"""
    _fft(RADIX2, a, twiddles, p)


def fft_radix4_recursive(a, twiddles, p):
    """fwd.rs:105-168"""
    _fft(RADIX4, a, twiddles, p)


def fft_split_radix_recursive(a, twiddles, p):
    """fwd.rs:207-272"""
    _fft(SPLIT_RADIX, a, twiddles, p)


def ifft_radix2_recursive(a, inv_twiddles, p, n_inv, top):
    """inv.rs:178-230"""
    _ifft(RADIX2, a, inv_twiddles, p, n_inv, top)


def ifft_radix4_recursive(a, inv_twiddles, p, n_inv, top):
    """inv.rs:106-176 (its size-2 base halves: the result carries 1/2 when log2 n is odd)"""
    _ifft(RADIX4, a, inv_twiddles, p, n_inv, top)


def ifft_split_radix_recursive(a, inv_twiddles, p, n_inv, top):
    """inv.rs:232-303"""
    _ifft(SPLIT_RADIX, a, inv_twiddles, p, n_inv, top)


def ifft_radix2_recursive_mut(a, inv_twiddles, p, n_inv, top):
    """fwd_1.rs:381-428 (takes no MultStats in the reference either)"""
    _ifft(RADIX2, a, inv_twiddles, p, n_inv, top)


def ifft_radix4_recursive_mut(a, inv_twiddles, p, n_inv, top, stats=None):
    """fwd_1.rs:296-379: the bases scale when `top`, nothing is halved; `stats` (a MultStats) receives the counters"""
    if stats is None:
        _ifft(RADIX4_MUT, a, inv_twiddles, p, n_inv, top)
        return
    import ctypes as C
    buf = (C.c_uint64 * 2)(stats.nonzero_mults, stats.skipped_mults)
    B.check(B.lib().ntt_b200_custum_radix_ifft_radix4_mut(B.host_ptr(a, np.uint32, True), a.size,
                                                          B.host_ptr(inv_twiddles, np.uint32), inv_twiddles.size, p,
                                                          n_inv, 1 if top else 0, buf), "in ifft_radix4_recursive_mut")
    stats.nonzero_mults, stats.skipped_mults = int(buf[0]), int(buf[1])


class MultStats:
    """fwd_1.rs:3-7"""

    def __init__(self):
        self.nonzero_mults = 0  # nonzero * nonzero
        self.skipped_mults = 0  # multiplications with zero

    def __repr__(self):
        return "MultStats { nonzero_mults: %d, skipped_mults: %d }" % (self.nonzero_mults, self.skipped_mults)


def _fft_mut(kind, a, twiddles, p, stats):
    import ctypes as C
    buf = (C.c_uint64 * 2)(stats.nonzero_mults, stats.skipped_mults)
    B.check(B.lib().ntt_b200_custum_radix_fft_mut(kind, B.host_ptr(a, np.uint32, True), a.size,
                                                  B.host_ptr(twiddles, np.uint32), twiddles.size, p, buf),
            "in custum_radix fft_mut")
    stats.nonzero_mults, stats.skipped_mults = int(buf[0]), int(buf[1])


def fft_radix4_recursive_mut(a, twiddles, p, stats):
    """fwd_1.rs:102-188"""
    _fft_mut(RADIX4, a, twiddles, p, stats)


def fft_radix2_recursive_mut(a, twiddles, p, stats):
    """fwd_1.rs:190-230"""
    _fft_mut(RADIX2, a, twiddles, p, stats)


def fft_split_radix_recursive_mut(a, twiddles, p, stats):
    """fwd_1.rs:232-294"""
    _fft_mut(SPLIT_RADIX, a, twiddles, p, stats)


def fft_mut_batch(kind, a, twiddles, p):
    """New: the `_mut` routine `kind` over every row of the (batch, n) array `a` in one launch; returns a (batch, 2) uint64
    array of (nonzero_mults, skipped_mults) per row -- the shape of the fork's dataset builder (examples/model/Dataset.rs)."""
    batch, n = a.shape
    stats = np.zeros((batch, 2), dtype=np.uint64)
    B.check(B.lib().ntt_b200_custum_radix_fft_mut_batch(kind, B.host_ptr(a, np.uint32, True), n, batch,
                                                        B.host_ptr(twiddles, np.uint32), twiddles.size, p,
                                                        stats.ctypes.data), "in fft_mut_batch")
    return stats


# ---- new: batched and device-resident forms -------------------------------------------------

def fft_batch(kind, a, twiddles, p):
    """`a`: (batch, n) uint32 array in host memory, every row transformed in place."""
    batch, n = a.shape
    B.check(B.lib().ntt_b200_custum_radix_fft_batch(kind, B.host_ptr(a, np.uint32, True), n, batch,
                                                    B.host_ptr(twiddles, np.uint32), twiddles.size, p), "in fft_batch")


def ifft_batch(kind, a, inv_twiddles, p, n_inv, top=True):
    batch, n = a.shape
    B.check(B.lib().ntt_b200_custum_radix_ifft_batch(kind, B.host_ptr(a, np.uint32, True), n, batch,
                                                     B.host_ptr(inv_twiddles, np.uint32), inv_twiddles.size, p,
                                                     n_inv, 1 if top else 0), "in ifft_batch")


def fft_device(kind, dev, n, batch, twiddles_dev, p, stream=None):
    """Device pointers (torch tensors) on the current device; asynchronous on `stream`."""
    B.check(B.lib().ntt_b200_custum_radix_fft_device(kind, B.dev_ptr(dev), n, batch, B.dev_ptr(twiddles_dev),
                                                     B.dev_numel(twiddles_dev, 4), p, B.stream_ptr(stream)))


def ifft_device(kind, dev, n, batch, inv_twiddles_dev, p, n_inv, top=True, stream=None):
    B.check(B.lib().ntt_b200_custum_radix_ifft_device(kind, B.dev_ptr(dev), n, batch, B.dev_ptr(inv_twiddles_dev),
                                                      B.dev_numel(inv_twiddles_dev, 4), p, n_inv, 1 if top else 0,
                                                      B.stream_ptr(stream)))


# ---- table construction: host-side restatement of the module's private helpers -----------------

def compute_primitive_root(p):
    """fwd.rs:42-68: the smallest generator of (Z/p)^*"""
    m, factors, i = p - 1, [], 2
    while i * i <= m:
        if m % i == 0:
            factors.append(i)
            while m % i == 0:
                m //= i
        i += 1
    if m > 1:
        factors.append(m)
    for g in range(2, p):
        if all(pow(g, (p - 1) // f, p) != 1 for f in factors):
            return g
    raise ValueError("no primitive root found (is p prime?)")


def make_twiddles(n, p):
    """fwd.rs:72-93: tw[k] = root^k with root = g^((p-1)/n)"""
    assert n > 0 and n & (n - 1) == 0
    assert (p - 1) % n == 0, "n must divide p-1"
    root = pow(compute_primitive_root(p), (p - 1) // n, p)
    tw, cur = np.empty(n, dtype=np.uint32), 1
    for k in range(n):
        tw[k] = cur
        cur = cur * root % p
    return tw


def make_inv_twiddles(tw, p):
    """fwd.rs:96-103: inv[k] = tw[k]^(p-2)"""
    return np.array([pow(int(t), p - 2, p) for t in tw], dtype=np.uint32)
