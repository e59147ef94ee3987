"""Pins the CPU oracle against every known-answer vector the reference's own tests hold for
the path.  CPU only."""
import numpy as np
import pytest

import oracle_lib as O
from oracle_lib import OraclePlan, SOLINAS_P

L = O.lib()


def largest_prime(factor, offset, lo, hi):
    import ctypes as C
    out = C.c_uint64()
    ok = L.tfo_largest_prime_in_arithmetic_progression64(factor, offset, lo, hi, C.byref(out))
    return out.value if ok else None


# --- prime.rs:188-207 ---------------------------------------------------------
def test_is_prime_under_1000_and_solinas():
    sieve = [True] * 1000
    sieve[0] = sieve[1] = False
    for i in range(2, 1000):
        if sieve[i]:
            for j in range(i * i, 1000, i):
                sieve[j] = False
    for n in range(1000):
        assert bool(L.tfo_is_prime64(n)) == sieve[n]
    assert L.tfo_is_prime64(SOLINAS_P)


# --- prime.rs:210-222 ---------------------------------------------------------
def test_prime_search():
    M = (1 << 64) - 1
    assert largest_prime(0, 2, 1, 4) == 2
    assert largest_prime(0, 2, 2, 2) == 2
    assert largest_prime(0, 2, 2, 1) is None
    assert largest_prime(1, 0, 14, 16) is None
    assert largest_prime(1, 0, 14, 17) == 17
    assert largest_prime(1, 0, 17, 18) == 17
    assert largest_prime(2, 1, 14, 16) is None
    assert largest_prime(2, 1, 14, 17) == 17
    assert largest_prime(2, 1, 17, 18) == 17
    assert largest_prime(6, 5, 0, M) == 18446744073709551557
    assert largest_prime(6, 1, 0, M) == 18446744073709551427


# --- roots.rs:150-172 ---------------------------------------------------------
SOLINAS_ROOTS = [
    (32, 8), (64, 2198989700608), (128, 14041890976876060974), (256, 14430643036723656017),
    (512, 4440654710286119610), (1024, 8816101479115663336), (2048, 10974926054405199669),
    (4096, 1206500561358145487), (8192, 10930245224889659871), (16384, 3333600369887534767),
    (32768, 15893793146607301539),
]


def test_primitive_root_solinas_table():
    import ctypes as C
    for n, expected in SOLINAS_ROOTS:
        out = C.c_uint64()
        assert L.tfo_find_root_solinas_64(2 * n, C.byref(out))
        assert out.value == expected
        assert L.tfo_exp_mod64(SOLINAS_P, expected, 2 * n) == 1
        # the plan uses exactly this root: twid[bit_rev(1)] = twid[n/2] = psi (prime64.rs:188-191)
        plan = OraclePlan(64, n, SOLINAS_P)
        assert int(plan.table("twid")[n // 2]) == expected


# --- roots.rs:133-147 ---------------------------------------------------------
def test_primitive_root_and_sqrt():
    import ctypes as C
    deg = 1 << 10
    p = largest_prime(deg, 1, 0, (1 << 64) - 1)
    out = C.c_uint64()
    assert L.tfo_find_primitive_root64(p, deg, C.byref(out))
    root = out.value
    assert pow(root, deg, p) == 1
    assert pow(root, deg // 2, p) == p - 1  # order exactly deg
    assert L.tfo_exp_mod64(p, root, deg) == 1


# --- prime32.rs:1338-1358 and prime64.rs:1557-1569 -----------------------------
def test_can_use_fast_reduction_code():
    for p in [1062862849, 1431669377] + [L.tfo_primes32(i) for i in range(10)]:
        plan = OraclePlan(32, 32, p)
        assert plan.s.can_use_fast_reduction_code, p
    assert not OraclePlan(32, 32, 0x7FE0_1001).s.can_use_fast_reduction_code
    for i in range(6):
        plan = OraclePlan(64, 32, L.tfo_primes52(i))
        assert plan.s.can_use_fast_reduction_code


# --- prime32.rs:1360-1397 ------------------------------------------------------
def test_barrett_regression_0x7fe01001():
    p = 0x7FE0_1001
    plan = OraclePlan(32, 32, p)
    value = 0x6E63593A
    lhs = np.full(32, value, dtype=np.uint32)
    rhs = np.full(32, value, dtype=np.uint32)
    acc = np.zeros(32, dtype=np.uint32)
    got = plan.mul_accumulate(acc, lhs, rhs)
    assert (got == (value * value) % p).all()
    got = plan.mul_assign_normalize(lhs, rhs)
    ninv = pow(32, p - 2, p)
    assert (got == (value * value * ninv) % p).all()


# --- prime64.rs:1988-1990 ------------------------------------------------------
def test_try_new_rejections():
    assert OraclePlan.try_new(64, 2048, 1024) is None
    assert OraclePlan.try_new(64, 8, SOLINAS_P) is None          # n < 16
    assert OraclePlan.try_new(64, 48, SOLINAS_P) is None         # not a power of two
    assert OraclePlan.try_new(32, 16, 1062862849) is None        # n < 32 for u32
    assert OraclePlan.try_new(32, 1 << 17, 1062862849) is None   # 2n does not divide p-1
    assert OraclePlan.try_new(64, 16, SOLINAS_P) is not None


# --- lib.rs:25-49 (crate doc example) ------------------------------------------
def test_doc_example_roundtrip():
    N, p = 32, 1062862849
    plan = OraclePlan(32, N, p)
    data = np.arange(N, dtype=np.uint32)
    t = plan.fwd(data)
    back = plan.inv(t)
    assert (back == (data.astype(np.uint64) * N % p).astype(np.uint32)).all()
    # values derived during the survey -- re-derived here, frozen below
    assert int(plan.table("twid")[N // 2]) == 398755272
    assert [int(x) for x in plan.table("twid")[:4]] == [1, 1009014033, 706332808, 419281921]
    assert [int(x) for x in t[:4]] == [8337849, 878691898, 914453352, 923715776]


def test_survey_derived_kats():
    plan = OraclePlan(64, 1024, SOLINAS_P)
    assert plan.s.n_inv_mod_p == 18428729670909296641
    x = np.arange(1024, dtype=np.uint64)
    f = plan.fwd(x)
    assert [int(v) for v in f[:4]] == [8990546331283213721, 16594485079839518376,
                                       8738969543403163269, 8200948690995499331]
    plan = OraclePlan(64, 2048, SOLINAS_P)
    assert plan.s.n_inv_mod_p == 18437736870161940481
    f = plan.fwd(np.arange(2048, dtype=np.uint64))
    assert [int(v) for v in f[:4]] == [1293403794811499965, 16222037460906770598,
                                       15810435000096167855, 13674017180432737007]
    assert int(OraclePlan(64, 65536, SOLINAS_P).table("twid")[32768]) == 14445062887364698470
    assert int(OraclePlan(64, 16, SOLINAS_P).table("twid")[8]) == 64
    plan = OraclePlan(32, 2048, 1073479681)
    assert int(plan.table("twid")[1024]) == 762388463
    f = plan.fwd(np.arange(2048, dtype=np.uint32))
    assert [int(v) for v in f[:4]] == [397780172, 960135856, 562966812, 620044689]


# --- independent restatement of A.2/A.3 in Python big ints ----
def _py_tables(p, n):
    def get_z(p):
        z = 2
        while pow(z, (p - 1) // 2, p) != p - 1:
            z += 1
        return z

    def sqrt_mod(p, q, s, z, x):
        m, c, t, r = s, pow(z, q, p), pow(x, q, p), pow(x, (q + 1) // 2, p)
        while True:
            if t == 0:
                return 0
            if t == 1:
                return r
            i, tp = 0, t
            while i < m:
                tp = tp * tp % p
                i += 1
                if tp == 1:
                    break
            if i == m:
                return None
            b = pow(c, 1 << (m - i - 1), p)
            m = i
            c = b * b % p
            t = t * c % p
            r = r * b % p

    if p == SOLINAS_P:
        w = pow(16334397945464290598, (1 << 32) // (2 * n), p)
    else:
        q, s = p - 1, 0
        while q % 2 == 0:
            q //= 2
            s += 1
        z, w = get_z(p), p - 1
        for _ in range((2 * n).bit_length() - 2):
            w = sqrt_mod(p, q, s, z, w)
    nb = n.bit_length() - 1
    br = lambda i: int(format(i, "0%db" % nb)[::-1], 2)
    tw, itw, wk = [0] * n, [0] * n, 1
    for k in range(n):
        tw[br(k)] = wk
        itw[br((n - k) % n)] = wk if k == 0 else p - wk
        wk = wk * w % p
    return tw, itw


BENCH_PRIMES_64 = [1125899904679937, 2251799813554177, 4611686018427322369,
                   9223372036853661697, 18446744073707716609, SOLINAS_P]
BENCH_PRIMES_32 = [1073479681, 2147352577, 4293918721]


@pytest.mark.parametrize("p", BENCH_PRIMES_64)
def test_tables_match_bigint_restatement_64(p):
    n = 256
    plan = OraclePlan(64, n, p)
    tw, itw = _py_tables(p, n)
    assert [int(x) for x in plan.table("twid")] == tw
    assert [int(x) for x in plan.table("inv_twid")] == itw
    if p < (1 << 63):
        assert [int(x) for x in plan.table("twid_shoup")] == [(w << 64) // p for w in tw]
        assert [int(x) for x in plan.table("inv_twid_shoup")] == [(w << 64) // p for w in itw]
    else:
        assert plan.table("twid_shoup") is None
    assert plan.s.n_inv_mod_p == pow(n, p - 2, p)


@pytest.mark.parametrize("p", BENCH_PRIMES_32)
def test_tables_match_bigint_restatement_32(p):
    n = 256
    plan = OraclePlan(32, n, p)
    tw, itw = _py_tables(p, n)
    assert [int(x) for x in plan.table("twid")] == tw
    assert [int(x) for x in plan.table("inv_twid")] == itw
    if p < (1 << 31):
        assert [int(x) for x in plan.table("twid_shoup")] == [(w << 32) // p for w in tw]
