"""tfhe's Ntt64View wrappers (ntt64.rs:89-266): oracle self-checks on CPU, GPU engine vs oracle."""
import numpy as np
import pytest

import oracle_lib as O
from oracle_lib import OraclePlan, SOLINAS_P

PRIMES = [SOLINAS_P, 4611686018427322369]


def rand_below(rng, p, shape):
    hi = rng.integers(0, 1 << 32, size=shape, dtype=np.uint64)
    lo = rng.integers(0, 1 << 32, size=shape, dtype=np.uint64)
    return np.array([((int(h) << 32) | int(l)) % p for h, l in zip(hi.ravel(), lo.ravel())], dtype=np.uint64).reshape(shape)


def test_oracle_semantics():
    n, p = 64, SOLINAS_P
    plan = OraclePlan(64, n, p)
    rng = np.random.default_rng(0)
    x = rand_below(rng, p, n)
    assert (plan.ntt64_forward(x, 0) == plan.fwd(x)).all()
    assert (plan.ntt64_forward(x, 1) == plan.normalize(plan.fwd(x))).all()
    # decomposition digits: small signed values stored as u64
    dig = rng.integers(-(1 << 20), 1 << 20, size=n).astype(np.int64)
    enc = np.array([int(d) % p for d in dig], dtype=np.uint64)
    assert (plan.ntt64_forward(dig.view(np.uint64), 2) == plan.fwd(enc)).all()
    # power-of-two modswitch: round(v * p / 2^w)
    for w in (64, 32, 17):
        v = (rng.integers(0, 1 << 63, size=n, dtype=np.uint64) * 2) & np.uint64(((1 << w) - 1) << (64 - w))
        want = np.array([((int(a) >> (64 - w)) * p + (1 << (w - 1))) >> w for a in v], dtype=np.uint64)
        assert (plan.ntt64_forward(v, 3, w) == plan.fwd(want)).all()
    st = rand_below(rng, p, n)
    got, clobbered = plan.ntt64_add_backward(st, plan.fwd(x), 0)
    assert [int(v) for v in got] == [(int(a) + int(b) * n) % p for a, b in zip(st, x)]
    assert (clobbered == plan.inv(plan.fwd(x))).all()


@pytest.mark.gpu
@pytest.mark.parametrize("p", PRIMES)
def test_gpu_ntt64_view(p):
    import torch
    import tfhe_ntt_b200 as T
    n, batch = 2048, 3
    gp, op = T.prime64.Plan.try_new(n, p), OraclePlan(64, n, p)
    view = T.ntt64.Ntt64View(gp)
    assert view.custom_modulus() == p and view.polynomial_size() == n
    rng = np.random.default_rng(p % 97)
    x = rand_below(rng, p, (batch, n))
    dig = rng.integers(-(1 << 22), 1 << 22, size=(batch, n)).astype(np.int64).view(np.uint64)
    out = np.zeros_like(x)
    view.forward(out, x)
    assert (out == op.ntt64_forward(x, 0)).all()
    view.forward_normalized(out, x)
    assert (out == op.ntt64_forward(x, 1)).all()
    view.forward_from_decomp(out, dig)
    assert (out == op.ntt64_forward(dig, 2)).all()
    for w in (64, 40):
        v = (rng.integers(0, 1 << 63, size=(batch, n), dtype=np.uint64) * 2) & np.uint64(((1 << w) - 1) << (64 - w))
        view.forward_from_power_of_two_modulus(w, out, v)
        assert (out == op.ntt64_forward(v, 3, w)).all()
    ntt = op.fwd(x)
    st = rand_below(rng, p, (batch, n))
    got_st, got_ntt = st.copy(), ntt.copy()
    view.add_backward(got_st, got_ntt)
    want_st, want_ntt = op.ntt64_add_backward(st, ntt, 0)
    assert (got_st == want_st).all() and (got_ntt == want_ntt).all()
    for w in (64, 40):
        st2 = rng.integers(0, 1 << 63, size=(batch, n), dtype=np.uint64) * 2
        got_st, got_ntt = st2.copy(), ntt.copy()
        view.add_backward_on_power_of_two_modulus(w, got_st, got_ntt)
        want_st, want_ntt = op.ntt64_add_backward(st2, ntt, 1, w)
        assert (got_st == want_st).all() and (got_ntt == want_ntt).all()
    # device form
    d_x = torch.from_numpy(dig.view(np.int64)).cuda()
    d_n = torch.empty_like(d_x)
    view.forward_device(d_n, d_x, batch, mode=2, stream=torch.cuda.current_stream())
    assert (d_n.cpu().numpy().view(np.uint64) == op.ntt64_forward(dig, 2)).all()
    with pytest.raises(AssertionError):
        view.forward(out[:1, :100].copy(), x[:1, :100].copy())
