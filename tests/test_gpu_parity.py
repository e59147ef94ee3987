"""GPU parity: the CUDA engine (through the C ABI / the tfhe_ntt_b200 host mirror) against the
CPU oracle on the same seeded inputs.  Bit-exact (integer path)."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O
from oracle_lib import OraclePlan, OracleNativePlan, SOLINAS_P

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PRIMES64 = [1125899904679937, 2251799813554177, 4611686018427322369, 9223372036853661697,
            18446744073707716609, SOLINAS_P]
PRIMES32 = [1062862849, 1073479681, 2147352577, 4293918721]


@pytest.fixture(scope="module")
def T():
    import tfhe_ntt_b200
    return tfhe_ntt_b200


def rand_below(rng, p, shape, dtype):
    hi = rng.integers(0, 1 << 32, size=shape, dtype=np.uint64)
    lo = rng.integers(0, 1 << 32, size=shape, dtype=np.uint64)
    v = (hi << np.uint64(32)) | lo
    if p <= (1 << 32):
        return (v % np.uint64(p)).astype(dtype)
    # python ints for the 64-bit moduli (numpy % on uint64 is exact too, but keep it obviously so)
    return np.array([int(x) % p for x in v.ravel()], dtype=np.uint64).reshape(shape).astype(dtype)


def edge_rows(p, n, dtype):
    rows = [np.zeros(n, dtype=dtype), np.full(n, p - 1, dtype=dtype), np.full(n, 1, dtype=dtype),
            np.arange(n, dtype=np.uint64).astype(dtype)]
    rows[3] = (rows[3].astype(np.uint64) % np.uint64(min(p, (1 << 63)))).astype(dtype)
    return np.stack(rows)


def plan_pair(T, bits, n, p):
    mod = T.prime64 if bits == 64 else T.prime32
    return mod.Plan.try_new(n, p), OraclePlan.try_new(bits, n, p)


@pytest.mark.parametrize("p", PRIMES64)
@pytest.mark.parametrize("n", [16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768])
def test_prime64_fwd_inv(T, p, n):
    gp, op = plan_pair(T, 64, n, p)
    assert (gp is None) == (op is None)
    if gp is None:
        pytest.skip("no plan")
    rng = np.random.default_rng(n + p % 9973)
    nrand = 3 if n <= 4096 else 2
    x = np.concatenate([rand_below(rng, p, (nrand, n), np.uint64), edge_rows(p, n, np.uint64)])
    want_f = op.fwd(x)
    got = x.copy()
    gp.fwd_batch(got)
    assert (got == want_f).all()
    assert (got < np.uint64(p)).all()
    want_i = op.inv(want_f)
    gp.inv_batch(got)
    assert (got == want_i).all()
    # per-polynomial drop-in call
    one = x[0].copy()
    gp.fwd(one)
    assert (one == want_f[0]).all()
    gp.inv(one)
    assert (one == want_i[0]).all()


def test_prime64_solinas_large_n(T):
    # 2^17: a radix-8 and a radix-4 strided pass; 2^20: two TMA-staged radix-16 passes (the second with stage 4)
    for n in [65536, 131072, 1 << 20]:
        gp, op = plan_pair(T, 64, n, SOLINAS_P)
        rng = np.random.default_rng(n)
        x = rand_below(rng, SOLINAS_P, (2, n), np.uint64)
        want = op.fwd(x)
        got = x.copy()
        gp.fwd_batch(got)
        assert (got == want).all()
        gp.inv_batch(got)
        assert (got == op.inv(want)).all()


@pytest.mark.parametrize("p", PRIMES32)
@pytest.mark.parametrize("n", [32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536])
def test_prime32_fwd_inv(T, p, n):
    gp, op = plan_pair(T, 32, n, p)
    assert (gp is None) == (op is None)
    if gp is None:
        pytest.skip("no plan")
    rng = np.random.default_rng(n + p % 9973)
    x = np.concatenate([rand_below(rng, p, (3, n), np.uint32), edge_rows(p, n, np.uint32)])
    want_f = op.fwd(x)
    got = x.copy()
    gp.fwd_batch(got)
    assert (got == want_f).all()
    want_i = op.inv(want_f)
    gp.inv_batch(got)
    assert (got == want_i).all()


@pytest.mark.parametrize("bits,p", [(64, p) for p in PRIMES64] + [(32, p) for p in PRIMES32])
def test_pointwise(T, bits, p):
    n = 256
    gp, op = plan_pair(T, bits, n, p)
    dt = np.uint64 if bits == 64 else np.uint32
    rng = np.random.default_rng(p % 1000)
    acc, lhs, rhs = (rand_below(rng, p, n, dt) for _ in range(3))
    for edge in (0, p - 1):
        lhs[edge % 7] = edge
        rhs[edge % 5] = edge
    a = acc.copy()
    gp.mul_accumulate(a, lhs, rhs)
    assert (a == op.mul_accumulate(acc, lhs, rhs)).all()
    l = lhs.copy()
    gp.mul_assign_normalize(l, rhs)
    assert (l == op.mul_assign_normalize(lhs, rhs)).all()
    v = lhs.copy()
    gp.normalize(v)
    assert (v == op.normalize(lhs)).all()
    # izip! truncation (lib.rs:658-688): the shortest slice bounds the work
    a = acc.copy()
    gp.mul_accumulate(a, lhs[:100], rhs[:50])
    assert (a[:50] == op.mul_accumulate(acc[:50], lhs[:50], rhs[:50])).all()
    assert (a[50:] == acc[50:]).all()
    assert gp.can_use_fast_reduction_code() == bool(op.s.can_use_fast_reduction_code)
    assert gp.ntt_size() == n and gp.modulus() == p


def test_try_new_rejections_and_errors(T):
    assert T.prime64.Plan.try_new(2048, 1024) is None          # prime64.rs:1988-1990
    assert T.prime64.Plan.try_new(8, SOLINAS_P) is None
    assert T.prime64.Plan.try_new(48, SOLINAS_P) is None
    assert T.prime32.Plan.try_new(16, 1062862849) is None
    assert T.prime32.Plan.try_new(1 << 17, 1062862849) is None
    plan = T.prime64.Plan.try_new(32, SOLINAS_P)
    assert not plan.use_ifma()
    with pytest.raises(AssertionError):  # assert_eq!(buf.len(), self.ntt_size()) prime64.rs:898
        plan.fwd(np.zeros(64, dtype=np.uint64))
    with pytest.raises(AssertionError):
        plan.inv(np.zeros(16, dtype=np.uint64))
    c = plan.clone()
    x = np.arange(32, dtype=np.uint64)
    y = x.copy()
    plan.fwd(x)
    c.fwd(y)
    assert (x == y).all()
    # empty batch is a no-op
    plan.fwd_batch(np.zeros(0, dtype=np.uint64))


def test_doc_example(T):
    # crate doc example lib.rs:25-49
    N, p = 32, 1062862849
    plan = T.prime32.Plan.try_new(N, p)
    data = np.arange(N, dtype=np.uint32)
    t = data.copy()
    plan.fwd(t)
    assert [int(v) for v in t[:4]] == [8337849, 878691898, 914453352, 923715776]
    plan.inv(t)
    assert (t == (data.astype(np.uint64) * N % p).astype(np.uint32)).all()


def test_device_api_and_fused(T):
    import torch
    n, p, batch = 2048, 1073479681, 24
    gp, op = plan_pair(T, 32, n, p)
    rng = np.random.default_rng(3)
    lhs = rand_below(rng, p, (batch, n), np.uint32)
    rhs = rand_below(rng, p, (batch, n), np.uint32)
    acc = rand_below(rng, p, (batch, n), np.uint32)
    want = op.inv(op.mul_accumulate(acc, op.fwd(lhs), rhs))
    d_l = torch.from_numpy(lhs.view(np.int32)).cuda()
    d_r = torch.from_numpy(rhs.view(np.int32)).cuda()
    d_a = torch.from_numpy(acc.view(np.int32)).cuda()
    d_o = torch.empty_like(d_l)
    st = torch.cuda.current_stream()
    gp.fwd_mac_inv_device(d_o, d_l, d_r, d_a, stream=st)
    got = d_o.cpu().numpy().view(np.uint32)
    assert (got == want).all()
    # shared rhs row (one GGSW row for the whole batch), no accumulator
    want2 = op.inv(op.mul_accumulate(np.zeros_like(lhs), op.fwd(lhs), np.tile(rhs[0], (batch, 1))))
    gp.fwd_mac_inv_device(d_o, d_l, d_r[:1].contiguous(), None, stream=st)
    assert (d_o.cpu().numpy().view(np.uint32) == want2).all()
    # unfused device sequence gives the same
    d_x = d_l.clone()
    gp.fwd_device(d_x, stream=st)
    gp.mul_accumulate_device(d_a, d_x, d_r, stream=st)
    gp.inv_device(d_a, stream=st)
    assert (d_a.cpu().numpy().view(np.uint32) == want).all()

    n, batch = 2048, 16
    gp, op = plan_pair(T, 64, n, SOLINAS_P)
    x = rand_below(rng, SOLINAS_P, (batch, n), np.uint64)
    g = rand_below(rng, SOLINAS_P, (2, n), np.uint64)
    d_x = torch.from_numpy(x.view(np.int64)).cuda()
    d_g = torch.from_numpy(g.view(np.int64)).cuda()
    d_o = torch.empty_like(d_x)
    gp.fwd_mac_inv_device(d_o, d_x, d_g, None, stream=st)
    want = op.inv(op.mul_accumulate(np.zeros_like(x), op.fwd(x), np.tile(g, (batch // 2, 1))))
    assert (d_o.cpu().numpy().view(np.uint64) == want).all()
    d_n = d_o.clone()
    gp.normalize_device(d_n, stream=st)
    assert (d_n.cpu().numpy().view(np.uint64).ravel() == op.normalize(want.ravel())).all()


@pytest.mark.parametrize("bits,n,p,rows,cols", [
    (64, 2048, SOLINAS_P, 4, 2),          # GLWE k=1, l=2 (BASELINE C3 shape): fused kernel
    (64, 1024, 4611686018427322369, 3, 3),
    (64, 2048, 9223372036853661697, 2, 2),  # 63-bit: Shoup without the Harvey range, Montgomery-form GGSW
    (64, 512, 18446744073707716609, 4, 1),  # generic 64-bit prime (Mont64)
    (64, 4096, 1125899904679937, 1, 4),     # 50-bit prime
    (32, 4096, 1073479681, 2, 4),
    (64, 256, SOLINAS_P, 4, 2),           # n outside the fused sizes: generic composition
    (32, 512, 2147352577, 2, 5),          # cols > 4: generic composition
])
def test_ext_product(T, bits, n, p, rows, cols):
    import torch
    gp, op = plan_pair(T, bits, n, p)
    dt = np.uint64 if bits == 64 else np.uint32
    tdt = torch.int64 if bits == 64 else torch.int32
    rng = np.random.default_rng(rows * 10 + cols)
    batch = 5
    x = rand_below(rng, p, (batch, rows, n), dt)
    g = rand_below(rng, p, (rows, cols, n), dt)
    d_x = torch.from_numpy(x.view(np.int64 if bits == 64 else np.int32)).cuda()
    d_g = torch.from_numpy(g.view(np.int64 if bits == 64 else np.int32)).cuda()
    d_o = torch.zeros((batch, cols, n), dtype=tdt, device="cuda")
    gp.ext_product_device(d_o, d_x, d_g, rows, cols, stream=torch.cuda.current_stream())
    got = d_o.cpu().numpy().view(dt)
    for b in range(batch):
        f = op.fwd(x[b])
        for c in range(cols):
            acc = np.zeros(n, dtype=dt)
            for r in range(rows):
                acc = op.mul_accumulate(acc, f[r], g[r, c])
            assert (got[b, c] == op.inv(acc)).all(), (b, c)


def test_fwd_mac_inv_batch_host(T):
    n, p = 1024, SOLINAS_P
    gp, op = plan_pair(T, 64, n, p)
    rng = np.random.default_rng(9)
    batch = 9000  # several staging chunks (32 MiB / 8 KiB = 4096 polynomials each), ragged tail
    lhs = rand_below(rng, p, (batch, n), np.uint64)
    ggsw = rand_below(rng, p, (4, n), np.uint64)
    acc = rand_below(rng, p, (2, n), np.uint64)
    out = np.zeros_like(lhs)
    gp.fwd_mac_inv_batch(out, lhs, ggsw, acc)
    for b in (0, 1, 2047, 2048, 4095, 4096, 4097, 8191, 8192, 8999):
        want = op.inv(op.mul_accumulate(acc[b % 2], op.fwd(lhs[b]), ggsw[b % 4]))
        assert (out[b] == want).all(), b
    # full-size operands, in place
    rhs = rand_below(rng, p, (64, n), np.uint64)
    x = lhs[:64].copy()
    gp.fwd_mac_inv_batch(x, x, rhs)
    want = op.inv(op.mul_accumulate(np.zeros_like(rhs), op.fwd(lhs[:64]), rhs))
    assert (x == want).all()
    with pytest.raises(AssertionError):
        gp.fwd_mac_inv_batch(out, lhs, np.ascontiguousarray(lhs[:7]))  # 9000 % 7 != 0


def rand_values(rng, value_bytes, shape, binary=False):
    n = int(np.prod(shape))
    if binary:
        bits = rng.integers(0, 2, size=n, dtype=np.uint64)
        if value_bytes == 16:
            return np.stack([bits, np.zeros(n, dtype=np.uint64)], axis=1)
        return bits.astype(O.VALUE_DTYPES[value_bytes])
    raw = rng.integers(0, 1 << 63, size=(n, 2), dtype=np.uint64) * 2 + rng.integers(0, 2, size=(n, 2), dtype=np.uint64)
    if value_bytes == 16:
        return np.ascontiguousarray(raw)
    return raw[:, 0].astype(O.VALUE_DTYPES[value_bytes])


def native_cls(T, kind):
    return [T.native32.Plan32, T.native32.Plan52, T.native64.Plan32, T.native64.Plan52,
            T.native128.Plan32, T.native_binary32.Plan32, T.native_binary32.Plan52,
            T.native_binary64.Plan32, T.native_binary64.Plan52, T.native_binary128.Plan32][kind]


@pytest.mark.parametrize("kind", range(10))
@pytest.mark.parametrize("n", [32, 256, 2048])
def test_native_plans(T, kind, n):
    gp = native_cls(T, kind).try_new(n)
    op = OracleNativePlan(kind, n)
    vb = op.value_bytes
    is_binary = kind >= O.NATIVE_BINARY32_PLAN32
    rng = np.random.default_rng(kind * 31 + n)
    value = rand_values(rng, vb, n)
    # fwd
    res = [np.zeros(n, dtype=op.rdtype) for _ in range(op.num_primes)]
    gp.fwd(value, *res)
    want_res = op.fwd(value)
    for a, b in zip(res, want_res):
        assert (a == b).all()
    # the reference's reconstruct_* property (native64.rs:1175-1242): inv(fwd(v)) == v * n, wrapping in the value word
    back = op.value_array()
    gp.inv(back, *[r.copy() for r in res])
    if vb == 16:
        ints = [((int(h) << 64) | int(l)) * n % (1 << 128) for l, h in value]
        want_back = np.array([[i & ((1 << 64) - 1), i >> 64] for i in ints], dtype=np.uint64)
    else:
        want_back = (value.astype(np.uint64) * np.uint64(n)).astype(value.dtype)
    assert (back == want_back).all()
    if is_binary:
        bval = rand_values(rng, vb, n, binary=True)
        bres = [np.zeros(n, dtype=op.rdtype) for _ in range(op.num_primes)]
        gp.fwd_binary(bval, *bres)
        for a, b in zip(bres, op.fwd(bval, binary=True)):
            assert (a == b).all()
    # inv on arbitrary canonical residues (the reference's scalar-vs-SIMD CRT test, native64.rs:1246-1292)
    rr = [rand_below(rng, int(op.lib.tfo_primes32(j)) if op.residue_bytes == 4 else int(op.lib.tfo_primes52(j)),
                     n, op.rdtype) for j in range(op.num_primes)]
    want_val, want_clobbered = op.inv(rr)
    got_val = op.value_array()
    got_res = [r.copy() for r in rr]
    gp.inv(got_val, *got_res)
    assert (got_val == want_val).all()
    for a, b in zip(got_res, want_clobbered):  # inv clobbers the residue buffers (native64.rs:1009-1013)
        assert (a == b).all()
    # polymul
    lhs = rand_values(rng, vb, n)
    rhs = rand_values(rng, vb, n, binary=is_binary)
    prod = op.value_array()
    gp.negacyclic_polymul(prod, lhs, rhs)
    assert (prod == op.negacyclic_polymul(lhs, rhs)).all()
    assert (prod == O.negacyclic_convolution_wrapping(vb, lhs, rhs)).all()
    assert gp.ntt_size() == n
    assert gp.ntt_0().modulus() == (op.lib.tfo_primes32(0) if op.residue_bytes == 4 else op.lib.tfo_primes52(0))
    with pytest.raises(AssertionError):
        gp.negacyclic_polymul(prod, lhs[: n // 2], rhs)


def test_native_try_new_none(T):
    # P1,P5,P6,P7 are 1 mod 2^16 only: native64::Plan32 stops at n = 32768
    assert T.native64.Plan32.try_new(65536) is None
    assert T.native64.Plan32.try_new(16) is None
    assert T.native64.Plan32.try_new(32768) is not None


@pytest.mark.parametrize("kind", range(10))
@pytest.mark.parametrize("n", [1024, 4096, 8192])
def test_native_plans_at_the_fused_kernel_sizes(T, kind, n):
    """n = 1024 ... 8192 run the one-CTA-per-product kernels (native_polymul_fused_kernel, ..._fused52_kernel,
    native_fwd_fused_kernel; 1024-thread CTAs at 8192, the scratch-arena path where the residues do not fit):
    every kind, a ragged batch, against the oracle and the wrapping schoolbook convolution."""
    gp = native_cls(T, kind).try_new(n)
    op = OracleNativePlan(kind, n)
    assert gp is not None
    vb = op.value_bytes
    is_binary = kind >= O.NATIVE_BINARY32_PLAN32
    rng = np.random.default_rng(kind * 131 + n)
    batch = 3
    shape = (batch, n, 2) if vb == 16 else (batch, n)
    lhs = np.stack([rand_values(rng, vb, n) for _ in range(batch)]).reshape(shape)
    rhs = np.stack([rand_values(rng, vb, n, binary=is_binary) for _ in range(batch)]).reshape(shape)
    # edge rows: all ones of the word (the largest value), and a zero polynomial
    lhs[1] = np.iinfo(lhs.dtype).max
    if not is_binary:
        rhs[1] = np.iinfo(rhs.dtype).max
    lhs[2, : n // 2] = 0
    prod = np.zeros_like(lhs)
    gp.negacyclic_polymul_batch(prod, lhs, rhs)
    for b in range(batch):
        want = op.negacyclic_polymul(np.ascontiguousarray(lhs[b]), np.ascontiguousarray(rhs[b]))
        assert (prod[b] == want).all(), (kind, n, b)
    assert (prod[0] == O.negacyclic_convolution_wrapping(vb, np.ascontiguousarray(lhs[0]), np.ascontiguousarray(rhs[0]))).all()
    # the fused forward (value read once, residues of every prime written) and the inverse on its output
    res = [np.zeros(n, dtype=op.rdtype) for _ in range(op.num_primes)]
    gp.fwd(np.ascontiguousarray(lhs[0]), *res)
    for a, w in zip(res, op.fwd(np.ascontiguousarray(lhs[0]))):
        assert (a == w).all()
    if is_binary:
        bres = [np.zeros(n, dtype=op.rdtype) for _ in range(op.num_primes)]
        gp.fwd_binary(np.ascontiguousarray(rhs[0]), *bres)
        for a, w in zip(bres, op.fwd(np.ascontiguousarray(rhs[0]), binary=True)):
            assert (a == w).all()
    want_val, _ = op.inv([r.copy() for r in res])
    got_val = op.value_array()
    gp.inv(got_val, *[r.copy() for r in res])
    assert (got_val == want_val).all()


def test_native_batch(T):
    n, batch = 1024, 20
    gp = T.native64.Plan32.try_new(n)
    op = OracleNativePlan(O.NATIVE64_PLAN32, n)
    rng = np.random.default_rng(77)
    lhs = rand_values(rng, 8, batch * n).reshape(batch, n)
    rhs = rand_values(rng, 8, batch * n).reshape(batch, n)
    prod = np.zeros_like(lhs)
    gp.negacyclic_polymul_batch(prod, lhs, rhs)
    for b in range(batch):
        assert (prod[b] == op.negacyclic_polymul(lhs[b], rhs[b])).all()


def test_shared_plan_is_reentrant(T):
    """Plans are Send + Sync in the reference (one Arc<Plan> shared by rayon workers,
    tfhe ntt64.rs:26-31): concurrent calls on one plan must not interfere."""
    import threading
    n, p = 1024, SOLINAS_P
    gp, op = plan_pair(T, 64, n, p)
    rng = np.random.default_rng(11)
    xs = [rand_below(rng, p, (17, n), np.uint64) for _ in range(8)]
    wants = [op.inv(op.fwd(x)) for x in xs]
    errors = []

    def worker(i):
        try:
            for _ in range(5):
                y = xs[i].copy()
                gp.fwd_batch(y)
                one = xs[i][0].copy()
                gp.fwd(one)
                assert (one == y[0]).all()
                gp.inv_batch(y)
                assert (y == wants[i]).all()
        except Exception as e:  # noqa: BLE001
            errors.append((i, repr(e)))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


@pytest.mark.parametrize("n,p", [(8192, SOLINAS_P), (16384, SOLINAS_P), (8192, 4611686018427322369),
                                  (16384, 9223372036853661697), (8192, 18446744073707716609)])
def test_prime64_cluster_kernels_many_polynomials(T, n, p):
    """Rows of 2^13 / 2^14 u64 coefficients (TMA-staged radix-16 pass + 512/1024-point kernels by default;
    the cluster-of-eight kernels have their own test below): enough polynomials to fill every SM several times, sampled
    rows against the oracle (generic_solinas.rs:449-514 / shoup.rs:544-615), all rows through
    inv(fwd(x)) = n * x."""
    gp, op = plan_pair(T, 64, n, p)
    rng = np.random.default_rng(n ^ (p % 1000003))
    batch = 333
    x = rand_below(rng, p, (batch, n), np.uint64)
    got = x.copy()
    gp.fwd_batch(got)
    assert (got < np.uint64(p)).all()
    rows = [0, 1, 7, 8, 147, 148, 331, 332]
    want_f = op.fwd(x[rows])
    assert (got[rows] == want_f).all()
    gp.inv_batch(got)
    assert (got[rows] == op.inv(want_f)).all()
    flat = got.reshape(-1)
    gp.normalize(flat)
    assert (flat.reshape(batch, n) == x).all()


@pytest.mark.parametrize("bits,n,p", [(64, 2048, SOLINAS_P), (32, 2048, 1073479681)])
@pytest.mark.parametrize("batch", [1, 2, 5, 37])
def test_host_batch_split_over_gpus(T, bits, n, p, batch):
    """*_batch_multi_gpu: contiguous slices, batch // G each with the remainder to the first GPUs
    (helper_multi_gpu.cu:57-88), one host thread per slice.  Runs over every visible GPU and, so that
    the slicing is exercised on a one-GPU box too, over three plans that all live on GPU 0."""
    import tfhe_ntt_b200._binding as B
    mod = T.prime64 if bits == 64 else T.prime32
    dtype = np.uint64 if bits == 64 else np.uint32
    op = OraclePlan.try_new(bits, n, p)
    rng = np.random.default_rng(batch * 31 + bits)
    x = rand_below(rng, p, (batch, n), dtype)
    want_f = op.fwd(x)
    want_i = op.inv(want_f)
    for devices in ([0, 0, 0], list(range(B.device_count()))):
        plans = mod.Plan.try_new_on_devices(n, p, devices)
        assert plans is not None and len(plans) == len(devices)
        got = x.copy()
        mod.Plan.fwd_batch_multi_gpu(plans, got)
        assert (got == want_f).all()
        mod.Plan.inv_batch_multi_gpu(plans, got)
        assert (got == want_i).all()


def test_host_batch_split_rejects_mismatched_plans(T):
    a = T.prime64.Plan.try_new(2048, SOLINAS_P)
    b = T.prime64.Plan.try_new(1024, SOLINAS_P)
    buf = np.zeros((2, 2048), dtype=np.uint64)
    with pytest.raises(Exception):
        T.prime64.Plan.fwd_batch_multi_gpu([a, b], buf)


def test_cluster_of_eight_kernels_opt_in():
    """The cluster-of-eight single-pass kernels (csrc/ntt_fast.cuh ntt_cluster8_*) are kept as an opt-in path
    (NTT_B200_CLUSTER=1, read once per process): run them in a child process against the oracle."""
    import subprocess
    import sys
    code = r'''
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np
import tfhe_ntt_b200 as T
from oracle_lib import OraclePlan
P = (1 << 64) - (1 << 32) + 1
for n, p in ((8192, P), (16384, P), (8192, 4611686018427322369), (16384, 18446744073707716609)):
    gp, op = T.prime64.Plan.try_new(n, p), OraclePlan.try_new(64, n, p)
    rng = np.random.default_rng(n)
    x = (rng.integers(0, 1 << 63, size=(19, n), dtype=np.uint64) * np.uint64(2)) %% np.uint64(p)
    got = x.copy(); gp.fwd_batch(got)
    want = op.fwd(x)
    assert (got == want).all(), ("fwd", n, p)
    gp.inv_batch(got)
    assert (got == op.inv(want)).all(), ("inv", n, p)
print("cluster ok")
''' % (ROOT, os.path.join(ROOT, "tests"))
    env = dict(os.environ, NTT_B200_CLUSTER="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "cluster ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("n", [2048, 16384, 65536])
def test_device_calls_are_cuda_graph_capturable(T, n):
    """The *_device entry points only enqueue work on the caller's stream (kernels, and for n >= 2^16 the
    TMA-staged pass): a forward + inverse + normalize sequence can be captured once in a CUDA graph and
    replayed; the replayed result must equal the input and the forward half must match the oracle."""
    import torch
    p = SOLINAS_P
    gp, op = plan_pair(T, 64, n, p)
    rng = np.random.default_rng(n + 5)
    batch = 6
    x = rand_below(rng, p, (batch, n), np.uint64)
    d = torch.from_numpy(x.view(np.int64)).cuda()
    snap = torch.empty_like(d)
    warm = torch.cuda.Stream()
    with torch.cuda.stream(warm):  # first use outside capture (shared-memory attributes are set once)
        gp.fwd_device(d, batch, stream=warm)
        gp.inv_device(d, batch, stream=warm)
        gp.normalize_device(d, stream=warm)
    torch.cuda.synchronize()
    assert (d.cpu().numpy().view(np.uint64) == x).all()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        st = torch.cuda.current_stream()
        gp.fwd_device(d, batch, stream=st)
        snap.copy_(d)
        gp.inv_device(d, batch, stream=st)
        gp.normalize_device(d, stream=st)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert (snap.cpu().numpy().view(np.uint64) == op.fwd(x)).all()
    assert (d.cpu().numpy().view(np.uint64) == x).all()


@pytest.mark.parametrize("n,p", [(32, 193), (64, 257), (2048, 12289), (1024, 65537), (4096, 786433),
                                  (2048, 536903681), (1024, 1073707009)])
def test_prime32_small_and_odd_sized_primes_pointwise(T, n, p):
    """The one-word Barrett of the p < 2^30 family (csrc/ntt_arith.cuh barrett32_narrow) depends on the bit
    length of p: small and mid-sized primes, transforms and every pointwise op against the oracle
    (prime32.rs:383-408, 477-486, 575-598), including the worst-case operands p - 1."""
    gp, op = plan_pair(T, 32, n, p)
    assert (gp is None) == (op is None)
    if gp is None:
        pytest.skip("no plan")
    rng = np.random.default_rng(p)
    x = np.concatenate([rand_below(rng, p, (3, n), np.uint32), edge_rows(p, n, np.uint32)])
    f = op.fwd(x)
    got = x.copy()
    gp.fwd_batch(got)
    assert (got == f).all()
    gp.inv_batch(got)
    assert (got == op.inv(f)).all()
    a, b, c = (rand_below(rng, p, (n,), np.uint32) for _ in range(3))
    a[:4], b[:4] = p - 1, p - 1
    acc = c.copy()
    gp.mul_accumulate(acc, a, b)
    assert (acc == op.mul_accumulate(c, a, b)).all()
    l = a.copy()
    gp.mul_assign_normalize(l, b)
    assert (l == op.mul_assign_normalize(a, b)).all()
    v = a.copy()
    gp.normalize(v)
    assert (v == op.normalize(a)).all()


@pytest.mark.parametrize("bits,n,p", [(64, 2048, SOLINAS_P), (64, 65536, SOLINAS_P), (64, 16384, 4611686018427322369),
                                      (32, 2048, 1073479681), (32, 32768, 1073479681)])
def test_device_pointers_without_16_byte_alignment(T, bits, n, p):
    """Device buffers that are only element-aligned (a view one coefficient into an allocation) take the
    generic kernels (no 128-bit accesses, no bulk copies): same bits as the aligned fast path and the oracle,
    for the transforms and the pointwise calls."""
    import torch
    gp, op = plan_pair(T, bits, n, p)
    dt = np.uint64 if bits == 64 else np.uint32
    sdt = np.int64 if bits == 64 else np.int32
    rng = np.random.default_rng(n + bits)
    batch = 3
    x = rand_below(rng, p, (batch, n), dt)
    st = torch.cuda.current_stream()
    big = torch.zeros(batch * n + 1, dtype=torch.int64 if bits == 64 else torch.int32, device="cuda")
    view = big[1:]
    assert view.data_ptr() % 16 != 0
    view.copy_(torch.from_numpy(x.view(sdt).reshape(-1)))
    gp.fwd_device(view, batch, stream=st)
    want_f = op.fwd(x)
    assert (view.cpu().numpy().view(dt).reshape(batch, n) == want_f).all()
    gp.inv_device(view, batch, stream=st)
    want_i = op.inv(want_f)
    assert (view.cpu().numpy().view(dt).reshape(batch, n) == want_i).all()
    gp.normalize_device(view, stream=st)
    assert (view.cpu().numpy().view(dt).reshape(batch, n) == x).all()


@pytest.mark.gpu
def test_small_host_calls_staged_path_matches_the_mapped_one(T):
    """The per-polynomial host calls run on the mapped pinned staging buffer by default; NTT_B200_ZERO_COPY=0
    stages through device memory instead.  Both must give the oracle's bits (the switch is read once per
    process, so the staged path runs in a child process)."""
    import subprocess
    import sys
    code = r'''
import os, sys
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np
import tfhe_ntt_b200 as T
import oracle_lib as O
for bits, n, p in [(64, 1024, O.SOLINAS_P), (64, 4096, O.SOLINAS_P), (32, 2048, 1073479681), (64, 64, O.SOLINAS_P)]:
    mod = T.prime64 if bits == 64 else T.prime32
    dt = np.uint64 if bits == 64 else np.uint32
    plan, ref = mod.Plan.try_new(n, p), O.OraclePlan(bits, n, p)
    rng = np.random.default_rng(n)
    x = (rng.integers(0, 1 << 62, size=n, dtype=np.uint64) %% np.uint64(p)).astype(dt)
    y = x.copy(); plan.fwd(y); assert (y == ref.fwd(x)).all()
    plan.inv(y); assert (y == ref.inv(ref.fwd(x))).all()
    a, l, r = x.copy(), ref.fwd(x), x[::-1].copy()
    plan.mul_accumulate(a, l, r); assert (a == ref.mul_accumulate(x, l, r)).all()
    plan.normalize(a); plan.mul_assign_normalize(a, r)
print("ok")
''' % (ROOT, ROOT)
    for zero_copy in ("1", "0"):
        env = dict(os.environ, NTT_B200_ZERO_COPY=zero_copy)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0 and "ok" in r.stdout, (zero_copy, r.stdout, r.stderr)


@pytest.mark.gpu
@pytest.mark.parametrize("bits,p", [(64, SOLINAS_P), (64, 4611686018427322369), (64, 9223372036853661697),
                                    (64, 18446744073707716609), (32, 1073479681), (32, 2147352577), (32, 4293918721)])
def test_ragged_batches_every_family_and_polys_per_thread(T, bits, p):
    """Batch sizes around the polynomials-per-thread / polynomials-per-CTA groupings of the single-CTA kernels
    (1, 2 or 4 polynomials per thread, up to 8 per CTA for short polynomials): the clamped tail rows must not
    be stored and every live row must match the oracle, device calls on an offset (still 16-byte aligned) view."""
    import torch
    dt = np.uint64 if bits == 64 else np.uint32
    sdt = np.int64 if bits == 64 else np.int32
    for n in (256, 512, 1024, 2048, 4096):
        gp, op = plan_pair(T, bits, n, p)
        rng = np.random.default_rng(n ^ bits)
        for batch in (1, 2, 3, 4, 5, 7, 8, 9, 15, 17, 33):
            hi = rng.integers(0, 1 << 32, size=(batch, n), dtype=np.uint64)
            lo = rng.integers(0, 1 << 32, size=(batch, n), dtype=np.uint64)
            x = (((hi << np.uint64(32)) | lo) % np.uint64(p)).astype(dt)  # numpy's uint64 % is exact
            x[0, :4] = [0, 1, p - 1, p - 1]
            guard = np.full((2, n), 0xA5A5A5A5, dtype=dt)  # rows before / after the batch must stay untouched
            buf = np.concatenate([guard[:1], x, guard[1:]])
            d = torch.from_numpy(buf.view(sdt)).cuda()
            view = d[1:1 + batch]
            st = torch.cuda.current_stream()
            gp.fwd_device(view, batch, stream=st)
            got = d.cpu().numpy().view(dt)
            assert (got[0] == guard[0]).all() and (got[-1] == guard[1]).all(), (n, batch)
            want = op.fwd(x)
            assert (got[1:1 + batch] == want).all(), (n, batch)
            gp.inv_device(view, batch, stream=st)
            got = d.cpu().numpy().view(dt)
            assert (got[0] == guard[0]).all() and (got[-1] == guard[1]).all(), (n, batch)
            assert (got[1:1 + batch] == op.inv(want)).all(), (n, batch)


@pytest.mark.gpu
@pytest.mark.parametrize("bits,p", [(64, SOLINAS_P), (64, 4611686018427322369), (64, 9223372036853661697),
                                    (64, 18446744073707716609), (32, 1073479681), (32, 2147352577), (32, 4293918721)])
def test_fused_fwd_mac_inv_every_family_ragged_and_shared_operands(T, bits, p):
    """ntt_fast_fwd_mac_inv_kernel for every modulus family and every single-CTA size: ragged batches, the
    pointwise operand per polynomial / one row for the whole batch / a cycle of three rows (the kernel's
    shared_index), with and without an accumulator (per polynomial or one shared row), in place and out of place."""
    import torch
    dt = np.uint64 if bits == 64 else np.uint32
    sdt = np.int64 if bits == 64 else np.int32
    st = torch.cuda.current_stream()

    def dev(a):
        return torch.from_numpy(np.ascontiguousarray(a).view(sdt)).cuda()

    def below(rng, shape):  # numpy's uint64 % is exact
        hi = rng.integers(0, 1 << 32, size=shape, dtype=np.uint64)
        lo = rng.integers(0, 1 << 32, size=shape, dtype=np.uint64)
        return (((hi << np.uint64(32)) | lo) % np.uint64(p)).astype(dt)

    for n in (256, 512, 1024, 2048, 4096):
        gp, op = plan_pair(T, bits, n, p)
        rng = np.random.default_rng(n * 3 + bits)
        for batch in (1, 3, 6, 9, 33):
            lhs = below(rng, (batch, n))
            lhs[0, :3] = [0, 1, p - 1]
            rhs_full = below(rng, (batch, n))
            rhs_full[-1, :] = p - 1
            acc_full = below(rng, (batch, n))
            f = op.fwd(lhs)
            for rhs_polys in sorted(c for c in {batch, 1, 3} if batch % c == 0):
                rhs = rhs_full[:rhs_polys]
                rhs_rows = np.concatenate([rhs] * (batch // rhs_polys))
                for acc_polys in (0, batch, 1):
                    acc_rows = (np.zeros_like(lhs) if acc_polys == 0 else
                                acc_full if acc_polys == batch else np.tile(acc_full[:1], (batch, 1)))
                    want = op.inv(op.mul_accumulate(acc_rows.copy(), f, rhs_rows))
                    d_l = dev(lhs)
                    d_o = torch.full_like(d_l, 0x5A)
                    gp.fwd_mac_inv_device(d_o, d_l, dev(rhs), None if acc_polys == 0 else dev(acc_full[:acc_polys]),
                                          stream=st)
                    assert (d_o.cpu().numpy().view(dt) == want).all(), (n, batch, rhs_polys, acc_polys)
                    assert (d_l.cpu().numpy().view(dt) == lhs).all()  # the operand is left alone
            gp.fwd_mac_inv_device(d_l, d_l, dev(rhs_full), None, stream=st)  # in place
            assert (d_l.cpu().numpy().view(dt) ==
                    op.inv(op.mul_accumulate(np.zeros_like(lhs), f, rhs_full))).all(), (n, batch)


@pytest.mark.gpu
@pytest.mark.parametrize("bits,p", [(64, SOLINAS_P), (64, 4611686018427322369), (64, 9223372036853661697),
                                    (64, 18446744073707716609), (32, 1073479681), (32, 2147352577), (32, 4293918721)])
def test_ext_product_every_family_size_and_matrix_shape(T, bits, p):
    """ntt_fast_ext_product_kernel<A, LOGN, COLS> (and the composition it falls back to) for every modulus family,
    n = 512 ... 4096 and GGSW shapes rows x cols up to 4 x 4, three batch items with edge rows."""
    import torch
    dt = np.uint64 if bits == 64 else np.uint32
    sdt = np.int64 if bits == 64 else np.int32
    tdt = torch.int64 if bits == 64 else torch.int32

    def below(rng, shape):
        hi = rng.integers(0, 1 << 32, size=shape, dtype=np.uint64)
        lo = rng.integers(0, 1 << 32, size=shape, dtype=np.uint64)
        return (((hi << np.uint64(32)) | lo) % np.uint64(p)).astype(dt)

    for n in (512, 1024, 2048, 4096):
        gp, op = plan_pair(T, bits, n, p)
        rng = np.random.default_rng(n + bits)
        for rows, cols in ((1, 1), (2, 2), (4, 2), (3, 3), (2, 4), (4, 4), (6, 2)):
            batch = 3
            x = below(rng, (batch, rows, n))
            x[0, 0, :3] = [0, 1, p - 1]
            x[1] = p - 1
            g = below(rng, (rows, cols, n))
            g[0, 0, :] = p - 1
            d_x = torch.from_numpy(x.view(sdt)).cuda()
            d_g = torch.from_numpy(g.view(sdt)).cuda()
            d_o = torch.zeros((batch, cols, n), dtype=tdt, device="cuda")
            gp.ext_product_device(d_o, d_x, d_g, rows, cols, stream=torch.cuda.current_stream())
            got = d_o.cpu().numpy().view(dt)
            assert (d_x.cpu().numpy().view(dt) == x).all() and (d_g.cpu().numpy().view(dt) == g).all()
            for b in range(batch):
                f = op.fwd(x[b])
                for c in range(cols):
                    acc = np.zeros(n, dtype=dt)
                    for r in range(rows):
                        acc = op.mul_accumulate(acc, f[r], g[r, c])
                    assert (got[b, c] == op.inv(acc)).all(), (n, rows, cols, b, c)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", range(10))
def test_native_device_batch_calls(T, kind):
    """fwd_device / inv_device / negacyclic_polymul_device of the CRT plans on device-resident batches
    (residues of prime j = one [batch][n] array): every row against the oracle, n below, inside and above the
    sizes of the one-CTA-per-product kernels."""
    import torch
    st = torch.cuda.current_stream()
    for n in (256, 2048, 16384):
        gp = native_cls(T, kind).try_new(n)
        op = OracleNativePlan.try_new(kind, n)
        assert (gp is None) == (op is None)
        if gp is None:
            continue
        vb = op.value_bytes
        is_binary = kind >= O.NATIVE_BINARY32_PLAN32
        rng = np.random.default_rng(kind * 7 + n)
        batch = 5 if n <= 2048 else 2
        shape = (batch, n, 2) if vb == 16 else (batch, n)
        lhs = np.stack([rand_values(rng, vb, n) for _ in range(batch)]).reshape(shape)
        rhs = np.stack([rand_values(rng, vb, n, binary=is_binary) for _ in range(batch)]).reshape(shape)
        sdt = {4: np.int32, 8: np.int64, 16: np.int64}[vb]
        rsdt = np.int32 if op.residue_bytes == 4 else np.int64
        d_l = torch.from_numpy(lhs.view(sdt)).cuda()
        d_r = torch.from_numpy(rhs.view(sdt)).cuda()
        d_res = [torch.zeros((batch, n), dtype=torch.int32 if op.residue_bytes == 4 else torch.int64, device="cuda")
                 for _ in range(op.num_primes)]
        gp.fwd_device(d_l, d_res, batch, stream=st)
        res = [r.cpu().numpy().view(op.rdtype) for r in d_res]
        for b in range(batch):
            for j, w in enumerate(op.fwd(np.ascontiguousarray(lhs[b]))):
                assert (res[j][b] == w).all(), (n, b, j)
        if is_binary:
            gp.fwd_device(d_r, d_res, batch, binary=True, stream=st)
            bres = [r.cpu().numpy().view(op.rdtype) for r in d_res]
            for b in range(batch):
                for j, w in enumerate(op.fwd(np.ascontiguousarray(rhs[b]), binary=True)):
                    assert (bres[j][b] == w).all(), (n, b, j)
            gp.fwd_device(d_l, d_res, batch, stream=st)
        d_v = torch.zeros_like(d_l)
        gp.inv_device(d_v, d_res, batch, stream=st)
        val = d_v.cpu().numpy().view(lhs.dtype).reshape(shape)
        for b in range(batch):
            want, _ = op.inv([r[b] for r in res])
            assert (val[b] == want).all(), (n, b)
        d_p = torch.zeros_like(d_l)
        gp.negacyclic_polymul_device(d_p, d_l, d_r, batch=batch, stream=st)
        prod = d_p.cpu().numpy().view(lhs.dtype).reshape(shape)
        for b in range(batch):
            want = op.negacyclic_polymul(np.ascontiguousarray(lhs[b]), np.ascontiguousarray(rhs[b]))
            assert (prod[b] == want).all(), (n, b)
        assert (d_l.cpu().numpy().view(lhs.dtype).reshape(shape) == lhs).all()


@pytest.mark.gpu
@pytest.mark.parametrize("bits,p", [(64, SOLINAS_P), (64, 4611686018427322369), (64, 9223372036853661697),
                                    (64, 18446744073707716609), (64, 1125899904679937), (32, 1073479681),
                                    (32, 2147352577), (32, 4293918721)])
def test_adversarial_canonical_patterns(T, bits, p):
    """Inputs that push the lazy representatives of every family to the ends of their ranges: constant and
    alternating vectors of p - 1, values around the 2^32 limb boundaries of the Solinas prime, one-hot vectors,
    through fwd, inv and the fused fwd+mac+inv, n = 256 ... 16384 (single-CTA and two-pass sizes)."""
    import torch
    dt = np.uint64 if bits == 64 else np.uint32
    sdt = np.int64 if bits == 64 else np.int32
    specials = [0, 1, 2, p - 1, p - 2, p // 2, p // 2 + 1]
    if bits == 64:
        specials += [v for v in ((1 << 32) - 1, 1 << 32, (1 << 32) + 1, p - (1 << 32), p - (1 << 32) + 1,
                                 (1 << 63) - 1, 1 << 63, 0xFFFFFFFF00000000, 0xFFFFFFFE00000001) if v < p]
    for n in (256, 2048, 4096, 16384):
        gp, op = plan_pair(T, bits, n, p)
        if gp is None:
            assert op is None
            continue
        rows = []
        for v in specials:
            rows.append(np.full(n, v, dtype=dt))
            alt = np.zeros(n, dtype=dt)
            alt[::2] = v
            alt[1::2] = p - 1
            rows.append(alt)
        for pos in (0, 1, n // 2, n - 1):
            one_hot = np.zeros(n, dtype=dt)
            one_hot[pos] = p - 1
            rows.append(one_hot)
        ramp = (np.arange(n, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) % np.uint64(p)).astype(dt)
        rows.append(np.sort(ramp)[::-1].copy())
        x = np.stack(rows)
        d = torch.from_numpy(x.view(sdt)).cuda()
        st = torch.cuda.current_stream()
        gp.fwd_device(d, x.shape[0], stream=st)
        f = op.fwd(x)
        assert (d.cpu().numpy().view(dt) == f).all(), n
        # the inverse on the special rows themselves (as NTT-domain data) and on the spectra
        for src in (x, f):
            d = torch.from_numpy(np.ascontiguousarray(src).view(sdt)).cuda()
            gp.inv_device(d, src.shape[0], stream=st)
            assert (d.cpu().numpy().view(dt) == op.inv(src)).all(), n
        if n <= 4096:
            d_l = torch.from_numpy(x.view(sdt)).cuda()
            d_o = torch.empty_like(d_l)
            gp.fwd_mac_inv_device(d_o, d_l, d_l, d_l, stream=st)  # rhs = acc = the special rows
            want = op.inv(op.mul_accumulate(x.copy(), f, x))
            assert (d_o.cpu().numpy().view(dt) == want).all(), n


@pytest.mark.gpu
@pytest.mark.parametrize("bits,p", [(64, q) for q in PRIMES64] + [(32, q) for q in PRIMES32])
def test_reference_property_product_equals_schoolbook(T, bits, p):
    """The reference's own unit tests of the prime plans, run on the GPU path (prime64.rs:1305-1361,
    prime32.rs: `test_product`): for random lhs, rhs < p
        inv(fwd(lhs) . fwd(rhs)) == n * negacyclic_convolution(lhs, rhs)    (pointwise product via mul_accumulate)
        inv(mul_assign_normalize(fwd(lhs), fwd(rhs))) == negacyclic_convolution(lhs, rhs)
    and every fwd output is below p -- against the O(n^2) schoolbook definition, not against the oracle's transform."""
    dt = np.uint64 if bits == 64 else np.uint32
    rng = np.random.default_rng(bits + (p & 0xFFFF))
    for n in (16, 32, 64, 128, 256, 512, 1024, 2048):
        mod = T.prime64 if bits == 64 else T.prime32
        gp = mod.Plan.try_new(n, p)
        if gp is None:  # prime32 plans start at n = 32; (p - 1) may not be divisible by 2n
            continue
        hi = rng.integers(0, 1 << 32, size=(2, n), dtype=np.uint64)
        lo = rng.integers(0, 1 << 32, size=(2, n), dtype=np.uint64)
        v = (((hi << np.uint64(32)) | lo) % np.uint64(p)).astype(dt)
        lhs, rhs = v[0].copy(), v[1].copy()
        conv = O.negacyclic_convolution_mod(bits, p, lhs, rhs)
        fl, fr = lhs.copy(), rhs.copy()
        gp.fwd(fl)
        gp.fwd(fr)
        assert int(fl.max()) < p and int(fr.max()) < p
        # product through mul_accumulate on a zero accumulator, then the unnormalised inverse: n * conv
        acc = np.zeros(n, dtype=dt)
        gp.mul_accumulate(acc, fl, fr)
        gp.inv(acc)
        want_n = np.array([int(c) * n % p for c in conv], dtype=dt)
        assert (acc == want_n).all(), n
        # mul_assign_normalize folds the 1/n in
        prod = fl.copy()
        gp.mul_assign_normalize(prod, fr)
        gp.inv(prod)
        assert (prod == conv).all(), n
        # normalize(inv(fwd(x))) == x
        back = fl.copy()
        gp.inv(back)
        gp.normalize(back)
        assert (back == lhs).all(), n


@pytest.mark.gpu
def test_plans_release_their_device_memory(T):
    """Creating and dropping plans (prime, CRT, bootstrap key) in a loop must not grow device memory:
    tables are freed with the handle (Drop in the Rust crate)."""
    import gc
    import torch
    from tfhe_ntt_b200 import ntt64_pbs as G

    def churn(rounds):
        for i in range(rounds):
            p64 = T.prime64.Plan.try_new(4096, SOLINAS_P)
            x = np.arange(4096, dtype=np.uint64)
            p64.fwd(x)
            p32 = T.prime32.Plan.try_new(2048, 1073479681)
            nat = T.native128.Plan32.try_new(1024)
            lhs = np.ones((1024, 2), dtype=np.uint64)
            prod = np.zeros_like(lhs)
            nat.negacyclic_polymul(prod, lhs, lhs)
            big = T.prime64.Plan.try_new(65536, SOLINAS_P)
            plan = T.prime64.Plan.try_new(512, SOLINAS_P)
            bsk = np.zeros(4 * 1 * 2 * 2 * 512, dtype=np.uint64)
            key = G.NttLweBootstrapKey.from_container(plan, bsk, 4, 2, 10, 1)
            del p64, p32, nat, big, key, plan
        gc.collect()
        torch.cuda.synchronize()

    churn(3)  # warm the per-thread staging buffers and the stream-ordered pool
    free0, _ = torch.cuda.mem_get_info()
    churn(40)
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < (8 << 20), (free0, free1)


@pytest.mark.gpu
def test_try_new_acceptance_agrees_with_the_oracle_on_random_moduli(T):
    """Plan::try_new returns None under exactly the reference's rules (prime64.rs:769-774, prime32.rs:667-672):
    random sizes (powers of two or not) and moduli (primes with and without 2n | p - 1, composites, tiny values);
    accepted plans must also transform like the oracle's."""
    rng = np.random.default_rng(2024)
    sizes = [1, 2, 8, 16, 24, 32, 48, 64, 128, 256, 1000, 1024, 4096]
    accepted = 0
    for bits in (64, 32):
        top = 64 if bits == 64 else 32
        for trial in range(120):
            n = sizes[int(rng.integers(0, len(sizes)))]
            kind = trial % 4
            width = int(rng.integers(8, top + 1))
            if kind == 0:    # an NTT-friendly prime for some size
                m = 1 << int(rng.integers(5, 14))
                out = C.c_uint64(0)
                ok = O.lib().tfo_largest_prime_in_arithmetic_progression64(m, 1, 0, (1 << width) - 1, C.byref(out))
                p = int(out.value) if ok else 1062862849
            elif kind == 1:  # a random odd number, usually composite
                p = (int(rng.integers(1, 1 << 62)) >> (62 - min(width, 62))) | 1
            elif kind == 2:  # tiny and degenerate moduli
                p = int(rng.integers(0, 40))
            else:            # even numbers and powers of two
                p = 1 << int(rng.integers(1, top))
            p &= (1 << top) - 1
            mod = T.prime64 if bits == 64 else T.prime32
            gp = mod.Plan.try_new(n, p)
            op = OraclePlan.try_new(bits, n, p)
            assert (gp is None) == (op is None), (bits, n, p)
            if gp is not None:
                accepted += 1
                dt = np.uint64 if bits == 64 else np.uint32
                x = (rng.integers(0, 1 << 62, size=n, dtype=np.uint64) % np.uint64(p)).astype(dt)
                y = x.copy()
                gp.fwd(y)
                assert (y == op.fwd(x)).all(), (bits, n, p)
    assert accepted >= 10


@pytest.mark.gpu
@pytest.mark.parametrize("bits,t", [(32, 29), (32, 30), (32, 31), (64, 32), (64, 50), (64, 51), (64, 61), (64, 62), (64, 63)])
def test_primes_on_both_sides_of_every_dispatch_threshold(T, bits, t):
    """The modulus families are chosen by the bit length of p (prime64.rs:897-968, prime32.rs:797-843: < 2^30 / 2^31,
    < 2^50 / 2^51 / 2^62 / 2^63): the largest NTT-friendly prime below 2^t and one just above it, transforms,
    pointwise operations and the fused kernel against the oracle."""
    def friendly(lo, hi, m):
        out = C.c_uint64(0)
        ok = O.lib().tfo_largest_prime_in_arithmetic_progression64(m, 1, lo, hi, C.byref(out))
        return int(out.value) if ok else None

    dt = np.uint64 if bits == 64 else np.uint32
    n = 2048
    cands = [friendly(0, (1 << t) - 1, 2 * n)]
    if t < bits:
        cands.append(friendly(1 << t, (1 << t) + (1 << (t - 4)), 2 * n))
    rng = np.random.default_rng(t)
    for p in [c for c in cands if c]:
        for size in (64, n):
            gp, op = plan_pair(T, bits, size, p)
            assert gp is not None and op is not None, (p, size)
            hi = rng.integers(0, 1 << 32, size=(3, size), dtype=np.uint64)
            lo = rng.integers(0, 1 << 32, size=(3, size), dtype=np.uint64)
            x = (((hi << np.uint64(32)) | lo) % np.uint64(p)).astype(dt)
            x[0, :4] = [0, 1, p - 1, p - 2]
            x[2] = p - 1
            y = x.copy()
            gp.fwd_batch(y)
            f = op.fwd(x)
            assert (y == f).all(), (p, size)
            gp.inv_batch(y)
            assert (y == op.inv(f)).all(), (p, size)
            a, b = f[0].copy(), f[1].copy()
            acc = x[2].copy()
            gp.mul_accumulate(acc, a, b)
            # identical to the reference except for its documented one-step Barrett quirk (r + p on two-step moduli)
            O.assert_mul_accumulate_matches_reference(bits, p, acc, op.mul_accumulate(x[2].copy(), f[0], f[1]), (p, size))
            exact = np.array([(int(u) * int(v) + p - 1) % p for u, v in zip(f[0], f[1])], dtype=np.uint64)
            assert (acc.astype(np.uint64) == exact).all(), (p, size)
            gp.mul_assign_normalize(a, b)
            assert (a == op.mul_assign_normalize(f[0].copy(), f[1])).all(), (p, size)
            z = x[1].copy()
            gp.normalize(z)
            assert (z == op.normalize(x[1].copy())).all(), (p, size)
            if size >= 256:
                out = np.zeros_like(x)
                gp.fwd_mac_inv_batch(out, x, x)
                assert (out == op.inv(op.mul_accumulate(np.zeros_like(x), f, x))).all(), (p, size)


@pytest.mark.gpu
@pytest.mark.parametrize("bits,p", [(64, 4899916394578788353), (64, 1125899881086977), (32, 1140830209), (32, 1068236801)])
def test_reference_one_step_barrett_quirk_is_the_only_divergence(T, bits, p):
    """The one place where the engine does NOT reproduce the reference bit for bit, pinned down.

    On moduli below 2^W / 3 whose Barrett quotient needs two correction steps the reference's `mul_accumulate`
    (one-step code, prime64.rs:586-609 / prime32.rs:575-598) can return r + p; the engine always returns the canonical
    r = (acc + lhs * rhs) mod p.  Over 2^18 random products: the oracle (a literal restatement) and the engine differ
    only by exactly p, only on such a modulus, and the engine's value is the exact one; `mul_assign_normalize` and
    `normalize` agree everywhere (their Shoup step absorbs the extra p)."""
    assert O.reference_barrett_quirk(bits, p)
    dt = np.uint64 if bits == 64 else np.uint32
    n, total = 2048, 1 << 18
    gp, op = plan_pair(T, bits, n, p)
    rng = np.random.default_rng(p & 0xFFFF)

    def below(count):
        hi = rng.integers(0, 1 << 32, size=count, dtype=np.uint64)
        lo = rng.integers(0, 1 << 32, size=count, dtype=np.uint64)
        return (((hi << np.uint64(32)) | lo) % np.uint64(p)).astype(dt)

    lhs, rhs, acc0 = below(total), below(total), below(total)
    acc = acc0.copy()
    gp.mul_accumulate(acc, lhs, rhs)
    ref = op.mul_accumulate(acc0.copy(), lhs, rhs)
    hits = O.assert_mul_accumulate_matches_reference(bits, p, acc, ref)
    idx = np.nonzero(acc != ref)[0]
    for i in list(idx[:50]) + list(rng.integers(0, total, size=200)):
        assert int(acc[i]) == (int(lhs[i]) * int(rhs[i]) + int(acc0[i])) % p
    print("reference quirk hits for p = %d: %d of %d" % (p, hits, total))
    m = lhs.copy()
    gp.mul_assign_normalize(m, rhs)
    assert (m == op.mul_assign_normalize(lhs.copy(), rhs)).all()
    z = lhs.copy()
    gp.normalize(z)
    assert (z == op.normalize(lhs.copy())).all()


@pytest.mark.gpu
@pytest.mark.parametrize("bits", [64, 32])
def test_every_modulus_width(T, bits):
    """NTT-friendly primes of every bit length the plans accept (7 ... 64 for prime64, 7 ... 32 for prime32), two per
    length (the largest one, and one from the lower half of the range, where the reference's Barrett constants behave
    differently): transforms, pointwise operations and the fused kernel against the oracle."""
    dt = np.uint64 if bits == 64 else np.uint32
    rng = np.random.default_rng(bits)

    def friendly(lo, hi, m):
        out = C.c_uint64(0)
        ok = O.lib().tfo_largest_prime_in_arithmetic_progression64(m, 1, lo, hi, C.byref(out))
        return int(out.value) if ok and lo <= int(out.value) <= hi else None

    seen = 0
    for width in range(7, bits + 1):
        for size in (32, 512):
            top = friendly(1 << (width - 1), (1 << width) - 1, 2 * size)
            low = friendly(1 << (width - 1), (1 << (width - 1)) + (1 << max(width - 3, 0)), 2 * size)
            for p in sorted({c for c in (top, low) if c}):
                gp, op = plan_pair(T, bits, size, p)
                assert (gp is None) == (op is None), (p, size)
                if gp is None:
                    continue
                seen += 1
                hi = rng.integers(0, 1 << 32, size=(3, size), dtype=np.uint64)
                lo = rng.integers(0, 1 << 32, size=(3, size), dtype=np.uint64)
                x = (((hi << np.uint64(32)) | lo) % np.uint64(p)).astype(dt)
                x[0, :3] = [0, 1, p - 1]
                x[2] = p - 1
                y = x.copy()
                gp.fwd_batch(y)
                f = op.fwd(x)
                assert (y == f).all(), (p, size)
                gp.inv_batch(y)
                assert (y == op.inv(f)).all(), (p, size)
                acc = x[2].copy()
                gp.mul_accumulate(acc, f[0], f[1])
                O.assert_mul_accumulate_matches_reference(bits, p, acc, op.mul_accumulate(x[2].copy(), f[0], f[1]), (p, size))
                a = f[0].copy()
                gp.mul_assign_normalize(a, f[1])
                assert (a == op.mul_assign_normalize(f[0].copy(), f[1])).all(), (p, size)
                z = x[1].copy()
                gp.normalize(z)
                assert (z == op.normalize(x[1].copy())).all(), (p, size)
                if size >= 256:
                    out = np.zeros_like(x)
                    gp.fwd_mac_inv_batch(out, x, x)
                    assert (out == op.inv(op.mul_accumulate(np.zeros_like(x), f, x))).all(), (p, size)
    assert seen > (60 if bits == 64 else 30)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", range(10))
def test_native_plans_remaining_sizes(T, kind):
    """The CRT plans at the sizes the other tests skip (64, 128, 512, 32768, and the first rejected ones): polymul,
    fwd and inv against the oracle; try_new agrees on acceptance."""
    for n in (16, 64, 128, 512, 32768, 65536):
        gp = native_cls(T, kind).try_new(n)
        op = OracleNativePlan.try_new(kind, n)
        assert (gp is None) == (op is None), (kind, n)
        if gp is None:
            continue
        vb = op.value_bytes
        is_binary = kind >= O.NATIVE_BINARY32_PLAN32
        rng = np.random.default_rng(kind * 17 + n)
        lhs = rand_values(rng, vb, n)
        rhs = rand_values(rng, vb, n, binary=is_binary)
        prod = op.value_array()
        gp.negacyclic_polymul(prod, lhs, rhs)
        assert (prod == op.negacyclic_polymul(lhs, rhs)).all(), (kind, n)
        res = [np.zeros(n, dtype=op.rdtype) for _ in range(op.num_primes)]
        gp.fwd(lhs, *res)
        for a, w in zip(res, op.fwd(lhs)):
            assert (a == w).all(), (kind, n)
        want_val, _ = op.inv([r.copy() for r in res])
        got_val = op.value_array()
        gp.inv(got_val, *[r.copy() for r in res])
        assert (got_val == want_val).all(), (kind, n)


@pytest.mark.gpu
@pytest.mark.parametrize("bits,p", [(64, SOLINAS_P), (64, 4611686018427322369), (64, 18446744073707716609),
                                    (32, 1073479681), (32, 4293918721)])
def test_device_pointwise_lengths_alignments_and_periods(T, bits, p):
    """normalize / mul_assign_normalize / mul_accumulate on device buffers: lengths that are not a multiple of the
    128-bit vector width, pointers off the 16-byte grid (the scalar twins of the vector kernels), operands reused
    cyclically with periods that do and do not divide the vector width; guard elements around the destination."""
    import torch
    dt = np.uint64 if bits == 64 else np.uint32
    sdt = np.int64 if bits == 64 else np.int32
    n = 64
    gp, op = plan_pair(T, bits, n, p)
    rng = np.random.default_rng(bits * 3 + (p & 7))
    st = torch.cuda.current_stream()
    n_inv = pow(n, -1, p)

    def below(count):
        hi = rng.integers(0, 1 << 32, size=count, dtype=np.uint64)
        lo = rng.integers(0, 1 << 32, size=count, dtype=np.uint64)
        return (((hi << np.uint64(32)) | lo) % np.uint64(p)).astype(dt)

    def exact(fn, *cols):
        return np.array([fn(*[int(v) for v in row]) % p for row in zip(*cols)], dtype=dt)

    for length in (1, 3, 4, 5, 64, 67, 256, 1001, 4096 + 2):
        for off in (0, 1, 3):
            for period in sorted({length, 1, 2, 4, 64} & set(d for d in (length, 1, 2, 4, 64) if length % d == 0)):
                acc0, lhs, rhs = below(length), below(length), below(period)
                guard = np.full(8, 0x5A5A5A5A, dtype=dt)
                buf = np.concatenate([guard[: 4 + off], acc0, guard])
                d_buf = torch.from_numpy(buf.view(sdt)).cuda()
                d_acc = d_buf[4 + off: 4 + off + length]
                d_lhs = torch.from_numpy(np.concatenate([guard[:off], lhs]).view(sdt)).cuda()[off:]
                d_rhs = torch.from_numpy(np.concatenate([guard[:off], rhs]).view(sdt)).cuda()[off:]
                rhs_rows = np.tile(rhs, length // period)
                gp.mul_accumulate_device(d_acc, d_lhs, d_rhs, length=length, lhs_len=length, rhs_len=period, stream=st)
                got = d_buf.cpu().numpy().view(dt)
                assert (got[: 4 + off] == guard[: 4 + off]).all() and (got[4 + off + length:] == guard).all()
                assert (got[4 + off: 4 + off + length] == exact(lambda a, l, r: a + l * r, acc0, lhs, rhs_rows)).all(), \
                    (length, off, period)
                d_l2 = d_lhs.clone()
                gp.mul_assign_normalize_device(d_l2, d_rhs, length=length, rhs_len=period, stream=st)
                assert (d_l2.cpu().numpy().view(dt) == exact(lambda l, r: l * r * n_inv, lhs, rhs_rows)).all(), \
                    (length, off, period)
            d_v = torch.from_numpy(np.concatenate([guard[:off], lhs, guard]).view(sdt)).cuda()
            gp.normalize_device(d_v[off: off + length], length=length, stream=st)
            got = d_v.cpu().numpy().view(dt)
            assert (got[off: off + length] == exact(lambda v: v * n_inv, lhs)).all() and (got[off + length:] == guard).all()


@pytest.mark.gpu
def test_mixed_plans_from_many_host_threads(T):
    """Eight host threads, each hammering a different mix of entry points (prime32 / prime64 single calls on the
    mapped staging buffer, chunked batches, the fused host pipeline, CRT products, pointwise calls) on plans shared by
    all of them -- per-thread staging buffers, cached streams and the stream-ordered pool must not interfere."""
    import threading
    rng = np.random.default_rng(99)
    p64, op64 = plan_pair(T, 64, 2048, SOLINAS_P)
    p32, op32 = plan_pair(T, 32, 1024, 1073479681)
    big, opbig = plan_pair(T, 64, 16384, SOLINAS_P)
    nat = T.native64.Plan32.try_new(1024)
    onat = OracleNativePlan(O.NATIVE64_PLAN32, 1024)

    def below(p, shape, dt):
        hi = rng.integers(0, 1 << 32, size=shape, dtype=np.uint64)
        lo = rng.integers(0, 1 << 32, size=shape, dtype=np.uint64)
        return (((hi << np.uint64(32)) | lo) % np.uint64(p)).astype(dt)

    jobs = []
    for i in range(8):
        x64 = below(SOLINAS_P, (5 + i, 2048), np.uint64)
        x32 = below(1073479681, (3 + i, 1024), np.uint32)
        xb = below(SOLINAS_P, (2, 16384), np.uint64)
        l, r = rand_values(rng, 8, 1024), rand_values(rng, 8, 1024)
        f64 = op64.fwd(x64)
        jobs.append(dict(
            x64=x64, f64=f64, i64=op64.inv(f64), x32=x32, f32=op32.fwd(x32), xb=xb, fb=opbig.fwd(xb), l=l, r=r,
            prod=onat.negacyclic_polymul(l, r),
            fused=op64.inv(op64.mul_accumulate(np.zeros_like(x64), f64, x64)),
            mac=op64.mul_accumulate(x64[0].copy(), f64[0], f64[1])))
    errors = []

    def worker(i):
        j = jobs[i]
        try:
            for rep in range(4):
                y = j["x64"].copy()
                p64.fwd_batch(y)
                assert (y == j["f64"]).all()
                one = j["x64"][rep % len(y)].copy()
                p64.fwd(one)
                assert (one == j["f64"][rep % len(y)]).all()
                p64.inv_batch(y)
                assert (y == j["i64"]).all()
                z = j["x32"].copy()
                p32.fwd_batch(z)
                assert (z == j["f32"]).all()
                w = j["xb"].copy()
                big.fwd_batch(w)
                assert (w == j["fb"]).all()
                prod = np.zeros(1024, dtype=np.uint64)
                nat.negacyclic_polymul(prod, j["l"], j["r"])
                assert (prod == j["prod"]).all()
                out = np.zeros_like(j["x64"])
                p64.fwd_mac_inv_batch(out, j["x64"], j["x64"])
                assert (out == j["fused"]).all()
                acc = j["x64"][0].copy()
                p64.mul_accumulate(acc, j["f64"][0], j["f64"][1])
                assert (acc == j["mac"]).all()
        except Exception as e:  # noqa: BLE001
            errors.append((i, repr(e)))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
