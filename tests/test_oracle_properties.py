"""The reference's own property tests (prime64.rs:1305-1361, :1456-1555, native64.rs:1175-1242,
native128.rs:394-447, native32.rs:506-558, native_binary*.rs tests), run against the CPU oracle."""
import numpy as np
import pytest

import oracle_lib as O
from oracle_lib import OraclePlan, OracleNativePlan, SOLINAS_P

L = O.lib()


def rand_below(rng, p, n, dtype):
    # uniform-ish below p for any p < 2^64
    hi = rng.integers(0, 1 << 32, size=n, dtype=np.uint64)
    lo = rng.integers(0, 1 << 32, size=n, dtype=np.uint64)
    v = [(int(h) << 32 | int(l)) % p for h, l in zip(hi, lo)]
    return np.array(v, dtype=np.uint64).astype(dtype)


PRIMES64 = [1125899904679937, 2251799813554177, 4611686018427322369, 9223372036853661697,
            18446744073707716609, SOLINAS_P]
PRIMES32 = [1073479681, 2147352577, 4293918721, 1062862849]


@pytest.mark.parametrize("p", PRIMES64)
@pytest.mark.parametrize("n", [16, 32, 128, 1024])
def test_product_64(p, n):
    rng = np.random.default_rng(n * 7 + p % 1000)
    plan = OraclePlan(64, n, p)
    lhs, rhs = rand_below(rng, p, n, np.uint64), rand_below(rng, p, n, np.uint64)
    conv = O.negacyclic_convolution_mod(64, p, lhs, rhs)
    fl, fr = plan.fwd(lhs), plan.fwd(rhs)
    assert (fl < p).all() and (fr < p).all()
    # lazy Shoup path == exact generic path (the reference's SIMD-vs-scalar style check)
    assert (fl == plan.fwd_generic(lhs)).all()
    prod = np.array([int(a) * int(b) % p for a, b in zip(fl, fr)], dtype=np.uint64)
    back = plan.inv(prod)
    assert (back == plan.inv_generic(prod)).all()
    assert [int(x) for x in back] == [int(c) * n % p for c in conv]
    # mul_assign_normalize + inv == conv (prime64.rs:1348-1360)
    man = plan.mul_assign_normalize(fl, fr)
    assert (plan.inv(man) == conv).all()


@pytest.mark.parametrize("p", PRIMES32)
@pytest.mark.parametrize("n", [32, 64, 256, 2048])
def test_product_32(p, n):
    rng = np.random.default_rng(n * 3 + p % 1000)
    plan = OraclePlan(32, n, p)
    lhs, rhs = rand_below(rng, p, n, np.uint32), rand_below(rng, p, n, np.uint32)
    conv = O.negacyclic_convolution_mod(32, p, lhs, rhs)
    fl, fr = plan.fwd(lhs), plan.fwd(rhs)
    assert (fl < p).all()
    assert (fl == plan.fwd_generic(lhs)).all()
    man = plan.mul_assign_normalize(fl, fr)
    assert (plan.inv(man) == conv).all()
    assert (plan.inv(man) == plan.inv_generic(man)).all()


@pytest.mark.parametrize("p", PRIMES64 + [1 << 61 | 1])
def test_pointwise_64(p):
    if not L.tfo_is_prime64(p):
        pytest.skip("not prime")
    rng = np.random.default_rng(5)
    n = 64
    plan = OraclePlan.try_new(64, n, p)
    if plan is None:
        pytest.skip("no root")
    a, l, r = (rand_below(rng, p, n, np.uint64) for _ in range(3))
    ninv = pow(n, p - 2, p)
    assert [int(x) for x in plan.mul_accumulate(a, l, r)] == \
        [(int(x) + int(y) * int(z)) % p for x, y, z in zip(a, l, r)]
    assert [int(x) for x in plan.mul_assign_normalize(l, r)] == \
        [int(y) * int(z) * ninv % p for y, z in zip(l, r)]
    assert [int(x) for x in plan.normalize(l)] == [int(y) * ninv % p for y in l]


@pytest.mark.parametrize("p", PRIMES32)
def test_pointwise_32(p):
    rng = np.random.default_rng(6)
    n = 64
    plan = OraclePlan(32, n, p)
    a, l, r = (rand_below(rng, p, n, np.uint32) for _ in range(3))
    ninv = pow(n, p - 2, p)
    assert [int(x) for x in plan.mul_accumulate(a, l, r)] == \
        [(int(x) + int(y) * int(z)) % p for x, y, z in zip(a, l, r)]
    assert [int(x) for x in plan.mul_assign_normalize(l, r)] == \
        [int(y) * int(z) * ninv % p for y, z in zip(l, r)]
    assert [int(x) for x in plan.normalize(l)] == [int(y) * ninv % p for y in l]


def rand_values(rng, value_bytes, n, binary=False):
    if binary:
        bits = rng.integers(0, 2, size=n, dtype=np.uint64)
        if value_bytes == 16:
            return np.stack([bits, np.zeros(n, dtype=np.uint64)], axis=1)
        return bits.astype(O.VALUE_DTYPES[value_bytes])
    raw = rng.integers(0, 1 << 63, size=(n, 2), dtype=np.uint64) * 2 + rng.integers(0, 2, size=(n, 2), dtype=np.uint64)
    if value_bytes == 16:
        return np.ascontiguousarray(raw)
    return raw[:, 0].astype(O.VALUE_DTYPES[value_bytes])


def to_int(value_bytes, a):
    if value_bytes == 16:
        return [int(lo) | (int(hi) << 64) for lo, hi in a]
    return [int(x) for x in a]


@pytest.mark.parametrize("kind", range(10))
@pytest.mark.parametrize("n", [32, 64, 256])
def test_native_plans(kind, n):
    rng = np.random.default_rng(kind * 100 + n)
    plan = OracleNativePlan(kind, n)
    vb = plan.value_bytes
    is_binary = kind >= O.NATIVE_BINARY32_PLAN32
    # roundtrip inv(fwd(v)) == v * n wrapping (native64.rs:1175-1205).  For the binary plans the
    # product of fewer primes only covers values up to (prod P_i)/2 / n; use small values there.
    value = rand_values(rng, vb, n)
    if is_binary:
        value = rand_values(rng, vb, n, binary=True)
    res = plan.fwd(value)
    back, _ = plan.inv(res)
    mask = (1 << (8 * vb)) - 1
    assert to_int(vb, back) == [(v * n) & mask for v in to_int(vb, value)]
    # negacyclic_polymul == wrapping schoolbook (native64.rs:1207-1242)
    lhs = rand_values(rng, vb, n)
    rhs = rand_values(rng, vb, n, binary=is_binary)
    prod = plan.negacyclic_polymul(lhs, rhs)
    want = O.negacyclic_convolution_wrapping(vb, lhs, rhs)
    assert (prod == want).all()


def test_crt_v1_equals_v2():
    rng = np.random.default_rng(9)
    P = [L.tfo_primes32(i) for i in range(5)]
    for _ in range(200):
        # residues of a small signed integer, so both Garner variants must agree
        x = int(rng.integers(-(1 << 62), 1 << 62))
        r = [x % p for p in P]
        a = L.tfo_reconstruct_32bit_01234_v2(*r)
        b = L.tfo_reconstruct_32bit_01234(*r)
        assert a == b == x & ((1 << 64) - 1)
