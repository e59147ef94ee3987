"""BASELINE.json configs C1..C5 at their full sizes on the GPU.  Checked through size-independent
properties (inv(fwd(x)) == n*x, normalize, linearity) over the whole batch plus bit-exact comparison
of sampled polynomials against the CPU oracle."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from oracle_lib import OraclePlan, OracleNativePlan, SOLINAS_P

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def T():
    import tfhe_ntt_b200
    return tfhe_ntt_b200


@pytest.fixture(scope="module")
def torch():
    import torch
    return torch


def dev_rand_below(torch, p, shape, bits, seed):
    """uniform-ish canonical residues generated on the device (int64/int32 storage)."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    if bits == 32:
        v = torch.randint(0, p, shape, dtype=torch.int64, device="cuda", generator=g)
        return v.to(torch.int32)  # two's complement reinterpretation keeps the low 32 bits
    hi = torch.randint(0, 1 << 31, shape, dtype=torch.int64, device="cuda", generator=g)
    lo = torch.randint(0, 1 << 32, shape, dtype=torch.int64, device="cuda", generator=g)
    v = (hi << 32) | lo  # < 2^63 <= p for the 64-bit primes used here
    return v


def dev_rand_full64(torch, p, shape, seed, edge_rows=True):
    """Canonical residues over the WHOLE range [0, p) of a 64-bit prime, [2^63, p) included, generated on
    the device in int64 storage; the first rows are replaced by the edge rows p-1, 0, 1."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    hi = torch.randint(0, 1 << 32, shape, dtype=torch.int64, device="cuda", generator=g)
    lo = torch.randint(0, 1 << 32, shape, dtype=torch.int64, device="cuda", generator=g)
    v = (hi << 32) | lo  # uniform 64-bit patterns (two's complement storage)
    if p < (1 << 63):
        v = (v & ((1 << 63) - 1)) % p
    else:
        ps = p - (1 << 64)  # p as an int64 bit pattern
        minv = -(1 << 63)
        ge = (v ^ minv) >= (ps ^ minv)  # unsigned v >= p
        v = torch.where(ge, v - ps, v)  # 2^64 - p < p: one subtraction is enough
    if edge_rows and len(shape) >= 2 and shape[0] >= 4:
        v[0] = p - 1 - (1 << 64) if p >= (1 << 63) else p - 1
        v[1] = 0
        v[2] = 1
        v[-1] = p - 1 - (1 << 64) if p >= (1 << 63) else p - 1
    return v


def to_u(t, bits):
    a = t.cpu().numpy()
    return a.view(np.uint64 if bits == 64 else np.uint32)


def mulmod_scalar_torch(torch, x, k, p):
    """(x * k) mod p on the host for a few rows only."""
    return np.array([[(int(v) * k) % p for v in row] for row in x], dtype=np.uint64)


def test_c1_prime64_n1024_single_polynomial_roundtrip(T):
    n, p = 1024, SOLINAS_P
    plan, ref = T.prime64.Plan.try_new(n, p), OraclePlan(64, n, p)
    rng = np.random.default_rng(1)
    x = (rng.integers(0, 1 << 63, size=n, dtype=np.uint64) * 2) % np.uint64(p)
    y = x.copy()
    plan.fwd(y)
    assert (y == ref.fwd(x)).all()
    plan.inv(y)
    plan.normalize(y)
    assert (y == x).all()


def test_c2_prime32_n2048_batch65536_fwd_mac_inv(T, torch):
    n, p, batch = 2048, 1073479681, 65536
    plan, ref = T.prime32.Plan.try_new(n, p), OraclePlan(32, n, p)
    lhs = dev_rand_below(torch, p, (batch, n), 32, 11)
    rhs = dev_rand_below(torch, p, (batch, n), 32, 12)
    acc = dev_rand_below(torch, p, (batch, n), 32, 13)
    out = torch.empty_like(lhs)
    st = torch.cuda.current_stream()
    plan.fwd_mac_inv_device(out, lhs, rhs, acc, stream=st)
    torch.cuda.synchronize()
    # sampled rows against the oracle
    idx = [0, 1, 777, 32768, 65535]
    L, R, A, G = (to_u(t[idx], 32) for t in (lhs, rhs, acc, out))
    want = ref.inv(ref.mul_accumulate(A, ref.fwd(L), R))
    assert (G == want).all()
    # whole batch: the fused kernel equals the three separate device calls
    x = lhs.clone()
    plan.fwd_device(x, stream=st)
    a2 = acc.clone()
    plan.mul_accumulate_device(a2, x, rhs, stream=st)
    plan.inv_device(a2, stream=st)
    assert torch.equal(a2, out)
    # whole batch: canonical range
    assert int(out.to(torch.int64).bitwise_and(0xFFFFFFFF).max()) < p


def test_c3_prime64_solinas_pbs_shaped(T, torch):
    # GLWE k=1, 2 levels, 4096 LWEs: 16384 digit polynomials forward, one shared GGSW of
    # (k+1)*l*(k+1) = 8 NTT-domain polynomials, 8192 inverse transforms (ntt64_pbs.rs:553-663)
    n, p, lwes, k1, l = 2048, SOLINAS_P, 4096, 2, 2
    plan, ref = T.prime64.Plan.try_new(n, p), OraclePlan(64, n, p)
    st = torch.cuda.current_stream()
    g = torch.Generator(device="cuda").manual_seed(3)
    # small signed digits mapped into [0,p) like ntt64.rs:231-238
    digits = torch.randint(-(1 << 22), 1 << 22, (lwes, k1 * l, n), dtype=torch.int64, device="cuda", generator=g)
    dig = torch.where(digits < 0, digits + torch.tensor(p - (1 << 64), dtype=torch.int64, device="cuda"), digits)
    ggsw = dev_rand_below(torch, p, (k1 * l, k1, n), 64, 4)
    ntt = dig.clone()
    plan.fwd_device(ntt, lwes * k1 * l, stream=st)
    acc = torch.zeros((lwes, k1, n), dtype=torch.int64, device="cuda")
    for row in range(k1 * l):
        lhs_row = ntt[:, row, :].contiguous()
        for col in range(k1):
            a = acc[:, col, :].contiguous()
            plan.mul_accumulate_device(a, lhs_row, ggsw[row, col].contiguous(), stream=st)
            acc[:, col, :] = a
    plan.inv_device(acc, lwes * k1, stream=st)
    torch.cuda.synchronize()
    for lwe in (0, 1234, 4095):
        d = to_u(dig[lwe], 64)
        f = ref.fwd(d)
        for col in range(k1):
            want = np.zeros(n, dtype=np.uint64)
            for row in range(k1 * l):
                want = ref.mul_accumulate(want, f[row], to_u(ggsw[row, col], 64))
            assert (to_u(acc[lwe, col], 64) == ref.inv(want)).all()
    # roundtrip over the whole digit batch
    plan.inv_device(ntt, lwes * k1 * l, stream=st)
    plan.normalize_device(ntt, stream=st)
    assert torch.equal(ntt, dig)


def test_headline_solinas_n2048_batch65536_full_size(T, torch):
    # the bench.py workload itself: 65536 x 2048, Solinas, coefficients over the whole of [0, p)
    n, p, batch = 2048, SOLINAS_P, 65536
    plan, ref = T.prime64.Plan.try_new(n, p), OraclePlan(64, n, p)
    x = dev_rand_full64(torch, p, (batch, n), 31)
    assert int((x < 0).sum()) > batch * n // 4  # really covers [2^63, p)
    y = x.clone()
    st = torch.cuda.current_stream()
    plan.fwd_device(y, stream=st)
    torch.cuda.synchronize()
    # first / last CTA (two polynomials each), the edge rows, and rows spread over the grid
    idx = [0, 1, 2, 3, 4, 5, 255, 256, 4097, 12345, 32767, 32768, 40001, 54321, 65532, 65533, 65534, 65535]
    X, F = to_u(x[idx], 64), to_u(y[idx], 64)
    assert (F == ref.fwd(X)).all()
    assert int(to_u(y, 64).max()) < p  # canonical over the whole batch
    plan.inv_device(y, stream=st)
    torch.cuda.synchronize()
    assert (to_u(y[idx], 64) == ref.inv(F)).all()
    plan.normalize_device(y, stream=st)
    assert torch.equal(y, x)  # inv(fwd(x)) / n == x over the whole batch


def test_c3_full_size_fused_external_product(T, torch):
    # C3 through the FUSED kernel (ntt_fast_ext_product_kernel, the one the bench and the PBS use) at
    # full size, against the unfused device calls over the whole batch and the oracle on sampled LWEs.
    n, p, lwes, k1, l = 2048, SOLINAS_P, 4096, 2, 2
    rows, cols = k1 * l, k1
    plan, ref = T.prime64.Plan.try_new(n, p), OraclePlan(64, n, p)
    st = torch.cuda.current_stream()
    for variant in ("digits", "full_range"):
        if variant == "digits":  # small signed digits mapped into [0, p) like ntt64.rs:231-238
            g = torch.Generator(device="cuda").manual_seed(5)
            d = torch.randint(-(1 << 22), 1 << 22, (lwes, rows, n), dtype=torch.int64, device="cuda", generator=g)
            inp = torch.where(d < 0, d + torch.tensor(p - (1 << 64), dtype=torch.int64, device="cuda"), d)
        else:
            inp = dev_rand_full64(torch, p, (lwes, rows, n), 6)
        ggsw = dev_rand_full64(torch, p, (rows, cols, n), 7)
        out = torch.zeros((lwes, cols, n), dtype=torch.int64, device="cuda")
        plan.ext_product_device(out, inp, ggsw, rows, cols, stream=st)
        # unfused: fwd, rows * cols mul_accumulate, inv
        ntt = inp.clone()
        plan.fwd_device(ntt, lwes * rows, stream=st)
        acc = torch.zeros((lwes, cols, n), dtype=torch.int64, device="cuda")
        for r in range(rows):
            lhs_row = ntt[:, r, :].contiguous()
            for c in range(cols):
                a = acc[:, c, :].contiguous()
                plan.mul_accumulate_device(a, lhs_row, ggsw[r, c].contiguous(), stream=st)
                acc[:, c, :] = a
        plan.inv_device(acc, lwes * cols, stream=st)
        torch.cuda.synchronize()
        assert torch.equal(out, acc), variant
        for lwe in (0, 1, 2, 3, 1234, 2048, 4094, 4095):
            f = ref.fwd(to_u(inp[lwe], 64))
            for c in range(cols):
                want = np.zeros(n, dtype=np.uint64)
                for r in range(rows):
                    want = ref.mul_accumulate(want, f[r], to_u(ggsw[r, c], 64))
                assert (to_u(out[lwe, c], 64) == ref.inv(want)).all(), (variant, lwe, c)


def test_c4_native64_plan32_n4096_batch16384(T, torch):
    n, batch = 4096, 16384
    plan, ref = T.native64.Plan32.try_new(n), OracleNativePlan(O.NATIVE64_PLAN32, n)
    g = torch.Generator(device="cuda").manual_seed(8)
    lhs = torch.randint(-(1 << 63), (1 << 63) - 1, (batch, n), dtype=torch.int64, device="cuda", generator=g)
    rhs = torch.randint(-(1 << 63), (1 << 63) - 1, (batch, n), dtype=torch.int64, device="cuda", generator=g)
    prod = torch.empty_like(lhs)
    plan.negacyclic_polymul_device(prod, lhs, rhs, stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    for b in (0, 8191, 16383):
        want = ref.negacyclic_polymul(to_u(lhs[b], 64), to_u(rhs[b], 64))
        assert (to_u(prod[b], 64) == want).all()
    # size-independent property over the whole batch: (x^0 coefficient of lhs*1) -- multiplying by the
    # constant polynomial 3 equals wrapping scalar multiplication
    three = torch.zeros((batch, n), dtype=torch.int64, device="cuda")
    three[:, 0] = 3
    plan.negacyclic_polymul_device(prod, lhs, three, stream=torch.cuda.current_stream())
    assert torch.equal(prod, lhs * 3)


def test_c5_prime64_n65536_batch1024(T, torch):
    n, p, batch = 65536, SOLINAS_P, 1024
    plan, ref = T.prime64.Plan.try_new(n, p), OraclePlan(64, n, p)
    x = dev_rand_full64(torch, p, (batch, n), 21)
    y = x.clone()
    st = torch.cuda.current_stream()
    plan.fwd_device(y, stream=st)
    torch.cuda.synchronize()
    for b in (0, 1, 2, 511, 1022, 1023):
        assert (to_u(y[b], 64) == ref.fwd(to_u(x[b], 64))).all()
    plan.inv_device(y, stream=st)
    plan.normalize_device(y, stream=st)
    assert torch.equal(y, x)


def test_cpp_host_mirror_example(T, tmp_path):
    exe = tmp_path / "example_prime"
    libdir = os.path.dirname(T.library_path())
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "example_prime.cpp"), "-o", str(exe),
                    "-L", libdir, "-ltfhe_ntt_b200", "-Wl,-rpath," + libdir], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "cpp example ok" in r.stdout


def test_cpp_host_mirror_native_plans(T, tmp_path):
    """tests/cpp/example_native.cpp: the reference's product-equals-schoolbook tests of all ten CRT plan types
    (native64.rs:1180-1244 and siblings) through the C++ mirror of the API."""
    exe = tmp_path / "example_native"
    libdir = os.path.dirname(T.library_path())
    subprocess.run(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "example_native.cpp"), "-o", str(exe),
                    "-L", libdir, "-ltfhe_ntt_b200", "-Wl,-rpath," + libdir], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "cpp native ok" in r.stdout


def test_batch_edge_sizes(T):
    # ragged batches: sizes that do not fill a CTA, and the chunked host pipeline boundary
    n, p = 256, SOLINAS_P
    plan, ref = T.prime64.Plan.try_new(n, p), OraclePlan(64, n, p)
    rng = np.random.default_rng(2)
    for batch in (1, 2, 7, 9, 33):
        x = (rng.integers(0, 1 << 63, size=(batch, n), dtype=np.uint64) * 2) % np.uint64(p)
        y = x.copy()
        plan.fwd_batch(y)
        assert (y == ref.fwd(x)).all()
        plan.inv_batch(y)
        assert (y == ref.inv(ref.fwd(x))).all()
    # several 32 MiB staging chunks (1024 polynomials each) with a ragged tail
    n = 4096
    plan, ref = T.prime64.Plan.try_new(n, p), OraclePlan(64, n, p)
    batch = 2 * 1024 + 5
    x = (rng.integers(0, 1 << 63, size=(batch, n), dtype=np.uint64) * 2) % np.uint64(p)
    y = x.copy()
    plan.fwd_batch(y)
    for b in (0, 1023, 1024, 2047, 2048, batch - 1):
        assert (y[b] == ref.fwd(x[b])).all()
    plan.inv_batch(y)
    plan_n = T.prime64.Plan.try_new(n, p)
    for b in (0, batch - 1):
        z = y[b].copy()
        plan_n.normalize(z)
        assert (z == x[b]).all()
