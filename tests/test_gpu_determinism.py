"""Race / barrier-elision stress for the fused kernels.  compute-sanitizer is closed on this GPU pool
(profiles/r02_compute_sanitizer_refusal.txt), so the hand-reasoned synchronisation of the fused kernels
(one CTA barrier per pass, no barrier between the forward's last read and the inverse's first write of
the exchange tile, the ping-pong DSMEM protocol of the cluster PBS kernel) is exercised the way a race
shows up: the same launch repeated many times, on several streams at once so that CTAs of different
launches share SMs in different interleavings, must give bit-identical results -- and the oracle's."""
import numpy as np
import pytest

import oracle_lib as O
from oracle_lib import OraclePlan, SOLINAS_P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    import tfhe_ntt_b200
    return tfhe_ntt_b200


def rand_below(rng, p, shape, dt):
    hi = rng.integers(0, 1 << 32, size=shape, dtype=np.uint64)
    lo = rng.integers(0, 1 << 32, size=shape, dtype=np.uint64)
    return (((hi << np.uint64(32)) | lo) % np.uint64(p)).astype(dt)


@pytest.mark.parametrize("bits,n,p", [(64, 2048, SOLINAS_P), (32, 2048, 1073479681), (64, 1024, 4611686018427322369),
                                      (32, 1024, 1073479681), (64, 4096, SOLINAS_P)])
def test_fused_kernels_repeat_bit_identically_under_concurrency(T, bits, n, p):
    import torch
    mod = T.prime64 if bits == 64 else T.prime32
    dt = np.uint64 if bits == 64 else np.uint32
    sdt = np.int64 if bits == 64 else np.int32
    plan, ref = mod.Plan.try_new(n, p), OraclePlan(bits, n, p)
    rng = np.random.default_rng(n + bits)
    batch = 1184 + 3  # several waves of CTAs and a ragged tail (not a multiple of 2 or 4 polynomials per CTA)
    lhs, rhs, acc = (rand_below(rng, p, (batch, n), dt) for _ in range(3))
    d_l, d_r, d_a = (torch.from_numpy(a.view(sdt)).cuda() for a in (lhs, rhs, acc))
    streams = [torch.cuda.Stream() for _ in range(3)]
    outs = [[torch.empty_like(d_l) for _ in range(8)] for _ in streams]
    xs = [[d_l.clone() for _ in range(8)] for _ in streams]
    torch.cuda.synchronize()
    for rep in range(8):
        for si, st in enumerate(streams):
            with torch.cuda.stream(st):
                plan.fwd_mac_inv_device(outs[si][rep], d_l, d_r, d_a, stream=st)
                plan.fwd_device(xs[si][rep], stream=st)
                plan.inv_device(xs[si][rep], stream=st)
    torch.cuda.synchronize()
    first = outs[0][0]
    for si in range(len(streams)):
        for rep in range(8):
            assert torch.equal(outs[si][rep], first), (si, rep)
            assert torch.equal(xs[si][rep], xs[0][0]), (si, rep)
    idx = [0, 1, 2, 3, 591, batch - 2, batch - 1]
    got = first.cpu().numpy().view(dt)[idx]
    want = ref.inv(ref.mul_accumulate(acc[idx], ref.fwd(lhs[idx]), rhs[idx]))
    assert (got == want).all()
    assert (xs[0][0].cpu().numpy().view(dt)[idx] == ref.inv(ref.fwd(lhs[idx]))).all()


def test_cluster_pbs_repeats_bit_identically(T):
    import torch
    from tfhe_ntt_b200 import ntt64_pbs as G
    n, n_lwe, gs, base_log, level = 2048, 24, 2, 23, 1
    plan = T.prime64.Plan.try_new(n, SOLINAS_P)
    rng = np.random.default_rng(77)
    p = np.uint64(SOLINAS_P)
    bsk = rand_below(rng, SOLINAS_P, n_lwe * level * gs * gs * n, np.uint64)
    key = G.NttLweBootstrapKey.from_container(plan, bsk, n_lwe, gs, base_log, level)
    batch = 300  # more ciphertexts than resident clusters: several waves
    lwe = rand_below(rng, SOLINAS_P, (batch, n_lwe + 1), np.uint64)
    lut = rand_below(rng, SOLINAS_P, gs * n, np.uint64)
    d_lwe = torch.from_numpy(lwe.view(np.int64)).cuda()
    d_lut = torch.from_numpy(lut.view(np.int64)).cuda()
    streams = [torch.cuda.Stream() for _ in range(2)]
    accs = [[torch.empty((batch, gs * n), dtype=torch.int64, device="cuda") for _ in range(4)] for _ in streams]
    torch.cuda.synchronize()
    for rep in range(4):
        for si, st in enumerate(streams):
            with torch.cuda.stream(st):
                G.blind_rotate_ntt64_device(key, d_lwe, d_lut, 1, accs[si][rep], batch, path=G.PATH_CLUSTER, stream=st)
    torch.cuda.synchronize()
    for si in range(2):
        for rep in range(4):
            assert torch.equal(accs[si][rep], accs[0][0]), (si, rep)
    ref = OraclePlan(64, n, SOLINAS_P)
    opbs = O.OraclePbs(ref, bsk, n_lwe, gs, base_log, level)
    got = accs[0][0].cpu().numpy().view(np.uint64)
    for b in (0, 1, 149, 299):
        assert (got[b] == opbs.blind_rotate(lwe[b], lut)).all(), b
