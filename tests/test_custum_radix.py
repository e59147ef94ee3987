"""custum_radix (reference: tfhe-ntt/src/custum_radix/{fwd.rs,inv.rs,fwd_1.rs}).

CPU half: the oracle's literal recursions against an arbitrary-precision restatement of the definition and
against each other, on the input of the reference's only test (fwd_1.rs:433-463, which prints and asserts
nothing) and on random vectors.  GPU half: the CUDA path through the C ABI, bit for bit against the oracle,
values and MultStats counters.
"""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O

L = O.lib()
_u32p, _sz, _u32 = C.c_void_p, C.c_size_t, C.c_uint32


class MultStats(C.Structure):
    _fields_ = [("nonzero_mults", C.c_size_t), ("skipped_mults", C.c_size_t)]


for _name in ("radix4", "radix2", "split_radix"):
    getattr(L, "tfo_cr_fft_%s_recursive" % _name).argtypes = [_u32p, _sz, _u32p, _u32]
    getattr(L, "tfo_cr_fft_%s_recursive" % _name).restype = None
    getattr(L, "tfo_cr_ifft_%s_recursive" % _name).argtypes = [_u32p, _sz, _u32p, _u32, _u32, C.c_int]
    getattr(L, "tfo_cr_ifft_%s_recursive" % _name).restype = None
    getattr(L, "tfo_cr_fft_%s_recursive_mut" % _name).argtypes = [_u32p, _sz, _u32p, _u32, C.POINTER(MultStats)]
    getattr(L, "tfo_cr_fft_%s_recursive_mut" % _name).restype = None
L.tfo_cr_ifft_radix4_recursive_mut.argtypes = [_u32p, _sz, _u32p, _u32, _u32, C.c_int, C.POINTER(MultStats)]
L.tfo_cr_ifft_radix4_recursive_mut.restype = None
L.tfo_cr_make_twiddles.argtypes = [_sz, _u32, _u32p]
L.tfo_cr_make_twiddles.restype = C.c_int
L.tfo_cr_make_inv_twiddles.argtypes = [_u32p, _sz, _u32, _u32p]
L.tfo_cr_make_inv_twiddles.restype = None
L.tfo_cr_compute_primitive_root.argtypes = [_u32]
L.tfo_cr_compute_primitive_root.restype = _u32

KINDS = {"radix2": 0, "radix4": 1, "split_radix": 2, "radix4_mut": 3}
# NTT-friendly primes of several widths: 2^16+1, 15*2^27+1, 2^32 - 2^20 + 1 (above 2^31)
PRIMES = [65537, 2013265921, 4293918721]


def ptr(a):
    assert a.dtype == np.uint32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


def tables(n, p):
    tw = np.zeros(max(n, 1), dtype=np.uint32)
    assert L.tfo_cr_make_twiddles(n, p, ptr(tw))
    inv = np.zeros_like(tw)
    L.tfo_cr_make_inv_twiddles(ptr(tw), n, p, ptr(inv))
    return tw, inv


def oracle_fft(kind, a, tw, p):
    out = a.copy()
    getattr(L, "tfo_cr_fft_%s_recursive" % kind)(ptr(out), out.size, ptr(tw), p)
    return out


def oracle_ifft(kind, a, inv, p, n_inv, top):
    out = a.copy()
    if kind == "radix4_mut":
        st = MultStats()
        L.tfo_cr_ifft_radix4_recursive_mut(ptr(out), out.size, ptr(inv), p, n_inv, int(top), C.byref(st))
    else:
        getattr(L, "tfo_cr_ifft_%s_recursive" % kind)(ptr(out), out.size, ptr(inv), p, n_inv, int(top))
    return out


def dft_definition(a, root, p):
    """X[k] = sum_j a[j] root^(jk) mod p with Python integers"""
    n = len(a)
    pw = [pow(root, e, p) for e in range(n)]
    return [sum(int(a[j]) * pw[(j * k) % n] for j in range(n)) % p for k in range(n)]


REFERENCE_TEST_INPUT = [5, 11, 3, 12, 8, 13, 2, 14, 4, 15, 7, 16, 6, 17, 1, 18, 3, 19, 9, 20, 2, 21, 5, 22,
                        7, 23, 4, 24, 1, 25, 8, 26, 9, 27, 6, 28, 3, 29, 5, 30, 2, 31, 8, 32, 4, 33, 7, 34,
                        1, 35, 6, 36, 3, 37, 9, 38, 2, 39, 5, 40, 7, 41, 4, 42]  # fwd_1.rs:447-456


def test_tables_follow_the_reference_helpers():
    import tfhe_ntt_b200.custum_radix as cr
    assert L.tfo_cr_compute_primitive_root(65537) == 3 == cr.compute_primitive_root(65537)
    assert L.tfo_cr_compute_primitive_root(2013265921) == cr.compute_primitive_root(2013265921) == 31
    for n, p in ((64, 65537), (1024, 2013265921), (16, 4293918721)):
        tw, inv = tables(n, p)
        root = int(tw[1]) if n > 1 else 1
        assert pow(root, n, p) == 1 and (n == 1 or pow(root, n // 2, p) == p - 1)
        assert [int(t) for t in tw] == [pow(root, k, p) for k in range(n)]
        assert [int(t) * int(i) % p for t, i in zip(tw, inv)] == [1] * n
        assert np.array_equal(tw, cr.make_twiddles(n, p))
        assert np.array_equal(inv, cr.make_inv_twiddles(tw, p))
    assert not L.tfo_cr_make_twiddles(64, 65539, ptr(np.zeros(64, dtype=np.uint32)))  # 64 does not divide p-1


def test_oracle_forward_is_the_dft_on_the_reference_test_input():
    n, p = 64, 65537  # fwd_1.rs:435-436
    tw, _ = tables(n, p)
    a = np.array(REFERENCE_TEST_INPUT, dtype=np.uint32)
    want = dft_definition(a, int(tw[1]), p)
    for kind in ("radix2", "radix4", "split_radix"):
        assert [int(x) for x in oracle_fft(kind, a, tw, p)] == want, kind


@pytest.mark.parametrize("p", PRIMES)
@pytest.mark.parametrize("n", [1, 2, 4, 8, 16, 32, 128])
def test_oracle_recursions_against_the_definition(n, p):
    rng = np.random.default_rng(n + p % 1000)
    tw, inv = tables(n, p)
    a = rng.integers(0, p, size=n, dtype=np.uint64).astype(np.uint32)
    a[::5] = 0
    root = int(tw[1]) if n > 1 else 1
    want = dft_definition(a, root, p)
    for kind in ("radix2", "radix4", "split_radix"):
        assert [int(x) for x in oracle_fft(kind, a, tw, p)] == want, kind
    inv_root = pow(root, p - 2, p)
    plain = dft_definition(a, inv_root, p)
    n_inv, half = pow(n, p - 2, p), pow(2, p - 2, p)
    logn = n.bit_length() - 1
    for top in (False, True):
        # which constant each recursion leaves on the sum (inv.rs:106-303, fwd_1.rs:296-379)
        f2 = n_inv if (top and n > 2) else 1
        f4 = f2 * (half if logn % 2 else 1) % p
        fm = n_inv if top else 1
        for kind, f in (("radix2", f2), ("split_radix", f2), ("radix4", f4), ("radix4_mut", fm)):
            got = oracle_ifft(kind, a, inv, p, n_inv, top)
            assert [int(x) for x in got] == [x * f % p for x in plain], (kind, top)
    # forward then the round-trip the fork's examples use
    back = oracle_ifft("radix2", oracle_fft("split_radix", a, tw, p), inv, p, n_inv, True)
    if n > 2:
        assert np.array_equal(back, a)


def test_oracle_mult_stats():
    """fwd_1.rs: every counted product lands in exactly one of the two counters; zeros are 'skipped'"""
    n, p = 64, 65537
    tw, _ = tables(n, p)
    a = np.array(REFERENCE_TEST_INPUT, dtype=np.uint32)
    want = oracle_fft("radix2", a, tw, p)
    totals = {}
    for kind in ("radix2", "radix4", "split_radix"):
        st, out = MultStats(), a.copy()
        getattr(L, "tfo_cr_fft_%s_recursive_mut" % kind)(ptr(out), n, ptr(tw), p, C.byref(st))
        assert np.array_equal(out, want), kind
        totals[kind] = (st.nonzero_mults, st.skipped_mults)
    # radix-2: n/2 products per level above the size-2 base, whose implicit product is counted when a[1] != 0
    # (fwd_1.rs:195-196): 32 * 5 + 32 for a vector without zeros in odd positions
    assert sum(totals["radix2"]) == 32 * 5 + 32
    # split-radix: 3 per quarter at every node with n > 2; radix-4: 4 per quarter above the size-4 base, 1 per base
    assert sum(totals["split_radix"]) < sum(totals["radix2"])
    assert sum(totals["radix4"]) == 4 * 16 + 4 * (4 * 4) + 16 * 1
    # an all-zero vector: every product is skipped, the size-2 bases count nothing
    st, z = MultStats(), np.zeros(n, dtype=np.uint32)
    L.tfo_cr_fft_radix2_recursive_mut(ptr(z), n, ptr(tw), p, C.byref(st))
    assert (st.nonzero_mults, st.skipped_mults) == (0, 32 * 5)


def golden_cases():
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "custum_radix_golden.json")) as f:
        return json.load(f)["cases"]


def test_oracle_reproduces_the_golden_fixture():
    """tests/golden/custum_radix_golden.json is computed from the definition alone (make_custum_radix_golden.py)"""
    for c in golden_cases():
        n, p = c["n"], c["p"]
        tw, inv = tables(n, p)
        assert int(tw[1]) == c["root"], c["name"]
        a = np.array(c["input"], dtype=np.uint32)
        for kind in ("radix2", "radix4", "split_radix"):
            assert [int(x) for x in oracle_fft(kind, a, tw, p)] == c["fft"], (c["name"], kind)
        n_inv = pow(n, p - 2, p)
        assert [int(x) for x in oracle_ifft("radix2", a, inv, p, n_inv, True)] == c["ifft_radix2_top"], c["name"]
        assert [int(x) for x in oracle_ifft("split_radix", a, inv, p, n_inv, True)] == c["ifft_radix2_top"], c["name"]
        assert [int(x) for x in oracle_ifft("radix4", a, inv, p, n_inv, True)] == c["ifft_radix4_top"], c["name"]


def test_rejections_need_no_gpu():
    """bad shapes are refused before any CUDA call (the reference overflows its stack / panics on an index)"""
    import tfhe_ntt_b200.custum_radix as cr
    tw = np.ones(8, dtype=np.uint32)
    for n in (0, 3, 6, 12):
        with pytest.raises(AssertionError):
            cr.fft_radix2_recursive(np.zeros(n, dtype=np.uint32), tw if n <= 8 else np.ones(16, dtype=np.uint32), 17)
    with pytest.raises(AssertionError):  # table shorter than the vector
        cr.fft_radix4_recursive(np.zeros(16, dtype=np.uint32), tw, 17)
    from tfhe_ntt_b200._binding import NttB200Error
    with pytest.raises(NttB200Error):  # the forward routines have no `_mut` inverse kind
        cr._fft(cr.RADIX4_MUT, np.zeros(8, dtype=np.uint32), tw, 17)
    with pytest.raises(NttB200Error):  # p = 0 divides by zero in the reference
        cr.fft_radix2_recursive(np.zeros(8, dtype=np.uint32), tw, 0)


# ------------------------------------------------------------------------------------------------
# GPU parity
# ------------------------------------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("p", PRIMES)
@pytest.mark.parametrize("n", [1, 2, 4, 8, 32, 64, 512, 2048, 4096])
def test_gpu_matches_oracle_per_vector(n, p):
    import tfhe_ntt_b200.custum_radix as cr
    rng = np.random.default_rng(7 * n + p % 97)
    tw, inv = tables(n, p)
    a = rng.integers(0, p, size=n, dtype=np.uint64).astype(np.uint32)
    a[::7] = 0
    if n > 1:
        a[1] = p - 1
    n_inv = pow(n, p - 2, p)
    for kind, fn in (("radix2", cr.fft_radix2_recursive), ("radix4", cr.fft_radix4_recursive),
                     ("split_radix", cr.fft_split_radix_recursive)):
        got = a.copy()
        fn(got, tw, p)
        assert np.array_equal(got, oracle_fft(kind, a, tw, p)), kind
    for top in (False, True):
        for kind, fn in (("radix2", cr.ifft_radix2_recursive), ("radix4", cr.ifft_radix4_recursive),
                         ("split_radix", cr.ifft_split_radix_recursive), ("radix4_mut", cr.ifft_radix4_recursive_mut),
                         ("radix2", cr.ifft_radix2_recursive_mut)):
            got = a.copy()
            fn(got, inv, p, n_inv, top)
            assert np.array_equal(got, oracle_ifft(kind, a, inv, p, n_inv, top)), (kind, top)


@pytest.mark.gpu
def test_gpu_reference_test_sequence():
    """fwd_1.rs:433-463: split-radix forward, then ifft_radix2_recursive_mut over the FORWARD table, top = false"""
    import tfhe_ntt_b200.custum_radix as cr
    n, p = 64, 65537
    tw, _ = tables(n, p)
    a = np.array(REFERENCE_TEST_INPUT, dtype=np.uint32)
    got = a.copy()
    cr.fft_split_radix_recursive(got, tw, p)
    assert [int(x) for x in got] == dft_definition(a, int(tw[1]), p)
    want = oracle_ifft("radix2", oracle_fft("split_radix", a, tw, p), tw, p, pow(n, p - 2, p), False)
    cr.ifft_radix2_recursive_mut(got, tw, p, pow(n, p - 2, p), False)
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_gpu_radix2_is_literal_for_any_table():
    """the schedule is the reference's radix-2 recursion, so it agrees with it on a table that is NOT a power table"""
    import tfhe_ntt_b200.custum_radix as cr
    n, p = 256, 2013265921
    rng = np.random.default_rng(5)
    tw = rng.integers(0, p, size=n + 3, dtype=np.uint64).astype(np.uint32)  # longer than n is allowed
    a = rng.integers(0, p, size=n, dtype=np.uint64).astype(np.uint32)
    got = a.copy()
    cr.fft_radix2_recursive(got, tw, p)
    assert np.array_equal(got, oracle_fft("radix2", a, tw[:n].copy(), p))


@pytest.mark.gpu
@pytest.mark.parametrize("n,batch", [(16, 1000), (64, 333), (1024, 77), (2048, 4096), (32768, 3), (65536, 5),
                                     (1 << 17, 2)])
def test_gpu_batches_host_and_device(n, batch):
    """several vectors per CTA (short n), one per CTA, and the bit-reversal + global stages above 2^15"""
    import torch
    import tfhe_ntt_b200.custum_radix as cr
    p = 2013265921 if n <= (1 << 27) else 65537
    rng = np.random.default_rng(n + batch)
    tw, inv = tables(n, p)
    a = rng.integers(0, p, size=(batch, n), dtype=np.uint64).astype(np.uint32)
    n_inv = pow(n, p - 2, p)
    rows = sorted({0, batch - 1, batch // 2, min(batch - 1, 5)})
    got = a.copy()
    cr.fft_batch(cr.SPLIT_RADIX, got, tw, p)
    for r in rows:
        assert np.array_equal(got[r], oracle_fft("split_radix", a[r], tw, p)), r
    back = got.copy()
    cr.ifft_batch(cr.RADIX2, back, inv, p, n_inv, True)
    assert np.array_equal(back, a)  # whole-batch property: the fork's round trip
    odd = got.copy()
    cr.ifft_batch(cr.RADIX4, odd, inv, p, n_inv, True)
    for r in rows[:2]:
        assert np.array_equal(odd[r], oracle_ifft("radix4", got[r], inv, p, n_inv, True)), r
    # device-resident
    d = torch.from_numpy(a.view(np.int32)).cuda()
    d_tw = torch.from_numpy(tw.view(np.int32)).cuda()
    d_inv = torch.from_numpy(inv.view(np.int32)).cuda()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        cr.fft_device(cr.RADIX4, d, n, batch, d_tw, p, stream=st)
    st.synchronize()
    assert np.array_equal(d.cpu().numpy().view(np.uint32), got)
    with torch.cuda.stream(st):
        cr.ifft_device(cr.RADIX4_MUT, d, n, batch, d_inv, p, n_inv, True, stream=st)
    st.synchronize()
    assert np.array_equal(d.cpu().numpy().view(np.uint32), a)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1 << 15, 1 << 16])
def test_gpu_long_vectors_wide_prime(n):
    """p >= 2^31 (plain table entries, 64-bit Barrett) on the largest single-CTA length and on the block + global-stage
    path"""
    import tfhe_ntt_b200.custum_radix as cr
    p = 4293918721
    rng = np.random.default_rng(n)
    tw, inv = tables(n, p)
    a = rng.integers(0, p, size=(3, n), dtype=np.uint64).astype(np.uint32)
    got = a.copy()
    cr.fft_batch(cr.RADIX2, got, tw, p)
    assert np.array_equal(got[1], oracle_fft("radix2", a[1], tw, p))
    n_inv = pow(n, p - 2, p)
    want = oracle_ifft("radix4", got[2], inv, p, n_inv, True)
    cr.ifft_batch(cr.RADIX4, got, inv, p, n_inv, True)
    assert np.array_equal(got[2], want)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [256, 2048, 1 << 16])
def test_gpu_device_calls_are_cuda_graph_capturable(n):
    """the device entry points only enqueue stream-ordered work (the scratch table is a stream-ordered allocation), so a
    forward + inverse pair can be captured once in a CUDA graph and replayed"""
    import torch
    import tfhe_ntt_b200.custum_radix as cr
    p = 2013265921
    rng = np.random.default_rng(n + 11)
    tw, inv = tables(n, p)
    batch = 5
    a = rng.integers(0, p, size=(batch, n), dtype=np.uint64).astype(np.uint32)
    d = torch.from_numpy(a.view(np.int32)).cuda()
    d_tw = torch.from_numpy(tw.view(np.int32)).cuda()
    d_inv = torch.from_numpy(inv.view(np.int32)).cuda()
    snap = torch.empty_like(d)
    n_inv = pow(n, p - 2, p)
    warm = torch.cuda.Stream()
    with torch.cuda.stream(warm):  # first use outside capture (shared-memory attributes are set once)
        cr.fft_device(cr.RADIX2, d, n, batch, d_tw, p, stream=warm)
        cr.ifft_device(cr.RADIX2, d, n, batch, d_inv, p, n_inv, True, stream=warm)
    torch.cuda.synchronize()
    assert np.array_equal(d.cpu().numpy().view(np.uint32), a)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        st = torch.cuda.current_stream()
        cr.fft_device(cr.SPLIT_RADIX, d, n, batch, d_tw, p, stream=st)
        snap.copy_(d)
        cr.ifft_device(cr.SPLIT_RADIX, d, n, batch, d_inv, p, n_inv, True, stream=st)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert np.array_equal(snap.cpu().numpy().view(np.uint32)[2], oracle_fft("split_radix", a[2], tw, p))
    assert np.array_equal(d.cpu().numpy().view(np.uint32), a)


@pytest.mark.gpu
def test_gpu_reproduces_the_golden_fixture():
    import tfhe_ntt_b200.custum_radix as cr
    for c in golden_cases():
        n, p = c["n"], c["p"]
        tw = cr.make_twiddles(n, p)
        inv = cr.make_inv_twiddles(tw, p)
        n_inv = pow(n, p - 2, p)
        for fn in (cr.fft_radix2_recursive, cr.fft_radix4_recursive, cr.fft_split_radix_recursive):
            a = np.array(c["input"], dtype=np.uint32)
            fn(a, tw, p)
            assert [int(x) for x in a] == c["fft"], (c["name"], fn.__name__)
        for fn, key in ((cr.ifft_radix2_recursive, "ifft_radix2_top"), (cr.ifft_split_radix_recursive, "ifft_radix2_top"),
                        (cr.ifft_radix4_recursive, "ifft_radix4_top")):
            a = np.array(c["input"], dtype=np.uint32)
            fn(a, inv, p, n_inv, True)
            assert [int(x) for x in a] == c[key], (c["name"], fn.__name__)


def oracle_stats(kind, a, tw, p, n_inv=0, top=False):
    st, out = MultStats(), a.copy()
    if kind == "ifft_radix4":
        L.tfo_cr_ifft_radix4_recursive_mut(ptr(out), out.size, ptr(tw), p, n_inv, int(top), C.byref(st))
    else:
        getattr(L, "tfo_cr_fft_%s_recursive_mut" % kind)(ptr(out), out.size, ptr(tw), p, C.byref(st))
    return out, (st.nonzero_mults, st.skipped_mults)


@pytest.mark.gpu
@pytest.mark.parametrize("p", [65537, 2013265921, 4293918721])
@pytest.mark.parametrize("n", [1, 2, 4, 8, 16, 32, 64, 128, 512, 2048, 4096])
def test_gpu_mult_stats_match_the_oracle(n, p):
    """the MultStats counters of fwd_1.rs (values too), on dense, sparse and structured inputs"""
    import tfhe_ntt_b200.custum_radix as cr
    rng = np.random.default_rng(n * 3 + p % 11)
    tw, inv = tables(n, p)
    dense = rng.integers(1, p, size=n, dtype=np.uint64).astype(np.uint32)
    sparse = dense.copy()
    sparse[rng.random(n) < 0.7] = 0
    const = np.full(n, 7, dtype=np.uint32)          # transform has a single non-zero: many zero operands inside
    delta = np.zeros(n, dtype=np.uint32)
    delta[n // 3] = 5
    inputs = [dense, sparse, const, delta, np.zeros(n, dtype=np.uint32)]
    if n == 64 and p == 65537:
        inputs.append(np.array(REFERENCE_TEST_INPUT, dtype=np.uint32))
    n_inv = pow(n, p - 2, p)
    for a in inputs:
        for kind, fn in (("radix2", cr.fft_radix2_recursive_mut), ("radix4", cr.fft_radix4_recursive_mut),
                         ("split_radix", cr.fft_split_radix_recursive_mut)):
            want, counts = oracle_stats(kind, a, tw, p)
            st, got = cr.MultStats(), a.copy()
            st.nonzero_mults, st.skipped_mults = 10, 20  # the counters accumulate like `&mut MultStats`
            fn(got, tw, p, st)
            assert np.array_equal(got, want), kind
            assert (st.nonzero_mults - 10, st.skipped_mults - 20) == counts, (kind, n, counts)
        for top in (False, True):
            want, counts = oracle_stats("ifft_radix4", a, inv, p, n_inv, top)
            st, got = cr.MultStats(), a.copy()
            cr.ifft_radix4_recursive_mut(got, inv, p, n_inv, top, st)
            assert np.array_equal(got, want), top
            assert (st.nonzero_mults, st.skipped_mults) == counts, ("ifft_radix4", n, top, counts)


def test_mult_stats_rejections_need_no_gpu():
    import tfhe_ntt_b200.custum_radix as cr
    st = cr.MultStats()
    from tfhe_ntt_b200 import NttB200Error
    # the counting kernel keeps every level in shared memory: n <= 4096 -- a capacity limit (its own status code),
    # not the length assertion of the reference
    with pytest.raises(NttB200Error, match="capacity limit"):
        cr.fft_radix2_recursive_mut(np.zeros(8192, dtype=np.uint32), np.ones(8192, dtype=np.uint32), 17, st)
    with pytest.raises(AssertionError):
        cr.fft_split_radix_recursive_mut(np.zeros(12, dtype=np.uint32), np.ones(16, dtype=np.uint32), 17, st)
    assert (st.nonzero_mults, st.skipped_mults) == (0, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("n,batch", [(64, 300), (1024, 41), (4096, 9)])
def test_gpu_mult_stats_batch(n, batch):
    """one launch for many vectors: values and per-vector counters against the oracle"""
    import tfhe_ntt_b200.custum_radix as cr
    p = 2013265921
    rng = np.random.default_rng(batch)
    tw, _ = tables(n, p)
    a = rng.integers(0, p, size=(batch, n), dtype=np.uint64).astype(np.uint32)
    a[rng.random((batch, n)) < 0.5] = 0
    a[batch // 2] = 0
    for kind_id, kind in ((cr.RADIX2, "radix2"), (cr.RADIX4, "radix4"), (cr.SPLIT_RADIX, "split_radix")):
        got = a.copy()
        stats = cr.fft_mut_batch(kind_id, got, tw, p)
        for r in sorted({0, 1, batch // 2, batch - 1}):
            want, counts = oracle_stats(kind, a[r], tw, p)
            assert np.array_equal(got[r], want), (kind, r)
            assert (int(stats[r, 0]), int(stats[r, 1])) == counts, (kind, r)
