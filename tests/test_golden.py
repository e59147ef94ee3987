"""Golden fixtures (tests/golden/ntt_golden.npz, written by tests/golden/make_golden.py): frozen input/output
vectors of the hot path.  CPU: the oracle still reproduces them and they satisfy the definition independently of
the oracle (arbitrary-precision evaluation at psi^(2*bitrev(j)+1), schoolbook convolutions).  GPU: the CUDA engine,
through the C ABI, reproduces them bit for bit."""
import os

import numpy as np
import pytest

import oracle_lib as O
from oracle_lib import SOLINAS_P, OracleNativePlan, OraclePlan

HERE = os.path.dirname(os.path.abspath(__file__))
P30 = 1073479681


@pytest.fixture(scope="module")
def G():
    return np.load(os.path.join(HERE, "golden", "ntt_golden.npz"))


def bit_rev(x, bits):
    return int(format(x, "0%db" % bits)[::-1], 2)


def test_golden_survey_kats_are_inside(G):
    # SURVEY section 8c derived KAT: Solinas N=1024, x = 0..N-1 -> fwd[0..4]
    assert (G["c1_x"][1] == np.arange(1024, dtype=np.uint64)).all()
    assert [int(v) for v in G["c1_fwd"][1][:4]] == [8990546331283213721, 16594485079839518376,
                                                    8738969543403163269, 8200948690995499331]


def test_golden_forward_is_evaluation_at_odd_powers_of_psi(G):
    """Independent of the oracle: psi(N=1024) from the survey's derived KAT, Python integers."""
    psi, n, p = 8816101479115663336, 1024, SOLINAS_P
    assert pow(psi, n, p) == p - 1
    x, f = G["c1_x"][0], G["c1_fwd"][0]
    for j in (0, 1, 5, 511, 512, 1023):
        w = pow(psi, 2 * bit_rev(j, 10) + 1, p)
        acc, cur = 0, 1
        for c in x:
            acc = (acc + int(c) * cur) % p
            cur = cur * w % p
        assert acc == int(f[j])
    # inverse is unnormalised: n * x
    assert all(int(v) == int(c) * n % p for v, c in zip(G["c1_inv"][0][:64], x[:64]))


def test_golden_convolution_vector(G):
    a, b, n, p = G["conv_a"], G["conv_b"], 64, SOLINAS_P
    out = [0] * n
    for i in range(n):
        for j in range(n):
            t = int(a[i]) * int(b[j])
            if i + j < n:
                out[i + j] = (out[i + j] + t) % p
            else:
                out[i + j - n] = (out[i + j - n] - t) % p
    assert [int(v) for v in G["conv_prod"]] == out


def test_oracle_reproduces_golden(G):
    for key, n in (("c1", 1024), ("c3", 2048), ("c5", 8192)):
        op = OraclePlan(64, n, SOLINAS_P)
        f = op.fwd(G[key + "_x"])
        assert (f == G[key + "_fwd"]).all()
        assert (op.inv(f) == G[key + "_inv"]).all()
    op = OraclePlan(32, 2048, P30)
    fl = op.fwd(G["c2_lhs"])
    assert (fl == G["c2_fwd"]).all()
    for b in range(2):
        assert (op.inv(op.mul_accumulate(G["c2_acc"][b], fl[b], G["c2_rhs"][b])) == G["c2_out"][b]).all()
    pl = OracleNativePlan(O.NATIVE64_PLAN32, 1024)
    for b in range(2):
        assert (pl.negacyclic_polymul(G["c4_lhs"][b], G["c4_rhs"][b]) == G["c4_prod"][b]).all()


@pytest.mark.gpu
def test_gpu_reproduces_golden_transforms(G):
    import tfhe_ntt_b200 as T
    for key, n in (("c1", 1024), ("c3", 2048), ("c5", 8192)):
        plan = T.prime64.Plan.try_new(n, SOLINAS_P)
        x = G[key + "_x"].copy()
        plan.fwd_batch(x)
        assert (x == G[key + "_fwd"]).all(), key
        plan.inv_batch(x)
        assert (x == G[key + "_inv"]).all(), key
    one = G["c1_x"][0].copy()  # the per-polynomial drop-in call of config C1
    plan = T.prime64.Plan.try_new(1024, SOLINAS_P)
    plan.fwd(one)
    assert (one == G["c1_fwd"][0]).all()
    plan.inv(one)
    plan.normalize(one)
    assert (one == G["c1_x"][0]).all()


@pytest.mark.gpu
def test_gpu_reproduces_golden_c2_and_c4(G):
    import tfhe_ntt_b200 as T
    plan = T.prime32.Plan.try_new(2048, P30)
    out = np.zeros_like(G["c2_lhs"])
    plan.fwd_mac_inv_batch(out, np.ascontiguousarray(G["c2_lhs"]), np.ascontiguousarray(G["c2_rhs"]),
                           np.ascontiguousarray(G["c2_acc"]))
    assert (out == G["c2_out"]).all()
    pl = T.native64.Plan32.try_new(1024)
    prod = np.zeros(1024, dtype=np.uint64)
    for b in range(2):
        pl.negacyclic_polymul(prod, np.ascontiguousarray(G["c4_lhs"][b]), np.ascontiguousarray(G["c4_rhs"][b]))
        assert (prod == G["c4_prod"][b]).all()
    plan = T.prime64.Plan.try_new(64, SOLINAS_P)
    a, b = G["conv_a"].copy(), G["conv_b"].copy()
    plan.fwd(a)
    plan.fwd(b)
    plan.mul_assign_normalize(a, b)
    plan.inv(a)
    assert (a == G["conv_prod"]).all()
