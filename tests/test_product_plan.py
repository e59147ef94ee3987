"""product::Plan: the reference's own tests (product.rs:976-1167) against the CPU oracle, and the
GPU engine against the oracle."""
import numpy as np
import pytest

import oracle_lib as O
from oracle_lib import OracleProductPlan

L = O.lib()
M64 = (1 << 64) - 1


def lp(factor, offset, lo, hi):
    import ctypes as C
    out = C.c_uint64()
    assert L.tfo_largest_prime_in_arithmetic_progression64(factor, offset, lo, hi, C.byref(out))
    return out.value


def cases(n=256):
    a = lp(2 * n, 1, 0, M64)
    b = lp(2 * n, 1, 0, (1 << 32) - 1)
    b1 = lp(2 * n, 1, 0, b - 1)
    c = lp(2 * n, 1, 0, 1 << 30)
    c1 = lp(2 * n, 1, 0, c - 1)
    d0 = lp(2 * n, 1, 0, 65535)
    d1 = lp(2 * n, 1, 0, d0 - 1)
    d2 = lp(2 * n, 1, 0, d1 - 1)
    d3 = lp(2 * n, 1, 0, d2 - 1)
    e0 = lp(2 * n, 1, 0, 1 << 33)
    e1 = lp(2 * n, 1, 0, 1 << 15)
    e2 = lp(2 * n, 1, 0, e1 - 1)
    return {"u64x1": [a], "u32x1": [b], "u32x2": [b, b1], "u30x2": [c, c1], "u32x4": [d0, d1, d2, d3],
            "u32x2_u64x1": [e0, e1, e2]}


def rand_mod(rng, p, n):
    hi = rng.integers(0, 1 << 32, size=n, dtype=np.uint64)
    lo = rng.integers(0, 1 << 32, size=n, dtype=np.uint64)
    return np.array([((int(h) << 32) | int(l)) % p for h, l in zip(hi, lo)], dtype=np.uint64)


@pytest.mark.parametrize("name", ["u64x1", "u32x1", "u32x2", "u30x2", "u32x4", "u32x2_u64x1"])
def test_oracle_roundtrip_like_reference(name):
    n = 256
    factors = cases(n)[name]
    p = int(np.prod([int(f) for f in factors], dtype=object))
    plan = OracleProductPlan(n, p, factors[::-1])  # unsorted on purpose
    rng = np.random.default_rng(len(name))
    standard = rand_mod(rng, p, n)
    n_inv = pow(n, -1, p)
    for accumulate in (False, True):
        ntt = plan.fwd(standard)
        back, _ = plan.inv(ntt, accumulate=accumulate)  # accumulates onto zeros
        assert [int(x) * n_inv % p for x in back] == [int(x) for x in standard]
    # Accumulate adds modulo the product
    base = rand_mod(rng, p, n)
    back, _ = plan.inv(plan.fwd(standard), standard=base, accumulate=True)
    assert [int(x) for x in back] == [(int(a) + int(s) * n) % p for a, s in zip(base, standard)]
    # residues really are the per-prime transforms of standard mod p_j (checked through the CRT above)
    # Bounded == Generic within the bound (2 x u32 arm)
    if name in ("u32x2", "u30x2"):
        small = rng.integers(0, 1000, size=n, dtype=np.uint64)
        neg = np.array([(p - int(x)) % p for x in rng.integers(0, 1000, size=n)], dtype=np.uint64)
        mixed = np.where(rng.integers(0, 2, size=n) == 1, small, neg)
        assert (plan.fwd(mixed, bounded=1001) == plan.fwd(mixed)).all()


def test_oracle_failures():
    n = 256
    e0 = lp(2 * n, 1, 0, 1 << 33)
    e1 = lp(2 * n, 1, 0, 1 << 15)
    assert OracleProductPlan.try_new(n, 0, [e0, 0]) is None                    # product.rs:1155-1159
    assert OracleProductPlan.try_new(n, e0 * e1 * e1 % (1 << 64), [e1, e0, e1]) is None  # :1162-1167
    assert OracleProductPlan.try_new(n, e0 * e1 + 2, [e0, e1]) is None          # wrong modulus
    assert OracleProductPlan.try_new(255, e0 * e1, [e0, e1]) is None            # odd size
    assert OracleProductPlan.try_new(n, e0 * e1, [1, e0, e1]) is not None       # ones are dropped


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["u64x1", "u32x1", "u32x2", "u30x2", "u32x4", "u32x2_u64x1"])
def test_gpu_product_plan(name):
    import tfhe_ntt_b200 as T
    n = 256
    factors = cases(n)[name]
    p = int(np.prod([int(f) for f in factors], dtype=object))
    gp = T.product.Plan.try_new(n, p, factors[::-1])
    op = OracleProductPlan(n, p, factors)
    assert gp.ntt_size() == n and gp.modulus() == p and gp.ntt_domain_len() == op.domain_len
    rng = np.random.default_rng(len(name) + 1)
    standard = rand_mod(rng, p, n)
    ntt = np.zeros(gp.ntt_domain_len(), dtype=np.uint64)
    gp.fwd(ntt, standard, T.product.FwdMode.Generic)
    want_ntt = op.fwd(standard)
    assert (ntt == want_ntt).all()
    # pointwise ops on the packed NTT domain
    other = op.fwd(rand_mod(rng, p, n))
    third = op.fwd(rand_mod(rng, p, n))
    a = ntt.copy()
    gp.mul_assign_normalize(a, other)
    assert (a == op.mul_assign_normalize(want_ntt, other)).all()
    a = third.copy()
    gp.mul_accumulate(a, ntt, other)
    assert (a == op.mul_accumulate(third, want_ntt, other)).all()
    a = ntt.copy()
    gp.normalize(a)
    assert (a == op.normalize(want_ntt)).all()
    # inv: Replace and Accumulate, ntt buffer transformed in place like the reference
    for mode, acc in ((T.product.InvMode.Replace, False), (T.product.InvMode.Accumulate, True)):
        base = rand_mod(rng, p, n)
        got_std, got_ntt = base.copy(), ntt.copy()
        gp.inv(got_std, got_ntt, mode)
        want_std, want_clobbered = op.inv(want_ntt, standard=base, accumulate=acc)
        assert (got_std == want_std).all()
        assert (got_ntt == want_clobbered).all()
    with pytest.raises(AssertionError):
        gp.fwd(ntt[:-1].copy(), standard)


@pytest.mark.gpu
def test_gpu_product_try_new_none_and_batch():
    import torch
    import tfhe_ntt_b200 as T
    n = 256
    e0 = lp(2 * n, 1, 0, 1 << 33)
    e1 = lp(2 * n, 1, 0, 1 << 15)
    assert T.product.Plan.try_new(n, 0, [e0, 0]) is None
    assert T.product.Plan.try_new(n, (e0 * e1 * e1) % (1 << 64), [e1, e0, e1]) is None
    assert T.product.Plan.try_new(255, e0 * e1, [e0, e1]) is None
    # batched device path: prime-major layout, roundtrip == n * x mod p
    c = cases(n)["u32x2_u64x1"]
    p = c[0] * c[1] * c[2]
    gp = T.product.Plan.try_new(n, p, c)
    batch = 7
    rng = np.random.default_rng(4)
    x = np.stack([rand_mod(rng, p, n) for _ in range(batch)])
    d_x = torch.from_numpy(x.view(np.int64)).cuda()
    d_ntt = torch.zeros(batch * gp.ntt_domain_len(), dtype=torch.int64, device="cuda")
    d_back = torch.zeros_like(d_x)
    st = torch.cuda.current_stream()
    gp.fwd_device(d_ntt, d_x, batch, stream=st)
    gp.inv_device(d_back, d_ntt, batch, stream=st)
    back = d_back.cpu().numpy().view(np.uint64)
    for b in range(batch):
        assert [int(v) for v in back[b]] == [int(v) * n % p for v in x[b]]
