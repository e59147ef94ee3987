"""Generates tests/golden/ntt_golden.npz: frozen input/output vectors of the hot path at small sizes.

The reference crate is Rust and cannot be built or run in this image (no rustc/cargo, dependencies not
vendored), so the vectors are produced by the CPU oracle (oracle/tfhe_ntt_oracle.c), which restates the
reference's scalar paths line by line and is itself pinned against every known-answer vector the reference's
own tests hold (tests/test_oracle_kat.py).  Before a vector is written, this script re-checks it against an
independent arbitrary-precision Python restatement of the definition (schoolbook negacyclic convolution /
evaluation of the polynomial at the odd powers of the root), so the fixture does not merely replay the oracle.

    python tests/golden/make_golden.py        # rewrites ntt_golden.npz (deterministic: fixed seeds)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402

SOLINAS_P = O.SOLINAS_P
P30 = 1073479681  # the reference bench's 30-bit prime = primes32::P9 (benches/ntt.rs:88)


def rand_below(rng, p, shape, dtype):
    hi = rng.integers(0, 1 << 32, size=shape, dtype=np.uint64)
    lo = rng.integers(0, 1 << 32, size=shape, dtype=np.uint64)
    v = ((hi << np.uint64(32)) | lo)
    flat = np.array([int(x) % p for x in v.reshape(-1)], dtype=np.uint64)
    return flat.reshape(shape).astype(dtype)


def negacyclic_schoolbook(a, b, modulus):
    """prime64.rs:1264-1276 / native128.rs:359-372 with Python integers."""
    n = len(a)
    out = [0] * n
    for i in range(n):
        ai = int(a[i])
        if not ai:
            continue
        for j in range(n):
            k = i + j
            t = ai * int(b[j])
            if k < n:
                out[k] = (out[k] + t) % modulus
            else:
                out[k - n] = (out[k - n] - t) % modulus
    return out


def bit_rev(x, bits):
    return int(format(x, "0%db" % bits)[::-1], 2) if bits else 0


def check_fwd_definition(plan_bits, n, p, x, f, psi):
    """fwd output j (bit-reversed order) is the polynomial evaluated at psi^(2*bitrev(j)+1)."""
    logn = n.bit_length() - 1
    for j in (0, 1, 2, n // 2, n - 1):
        e = 2 * bit_rev(j, logn) + 1
        w = pow(psi, e, p)
        acc, cur = 0, 1
        for c in x:
            acc = (acc + int(c) * cur) % p
            cur = cur * w % p
        assert acc == int(f[j]), (plan_bits, n, p, j)


def main():
    rng = np.random.default_rng(20261018)
    out = {}
    # C1: prime64 Solinas N=1024, single polynomial fwd / inv roundtrip
    n = 1024
    op = O.OraclePlan(64, n, SOLINAS_P)
    x = rand_below(rng, SOLINAS_P, (2, n), np.uint64)
    x[1] = np.arange(n, dtype=np.uint64)  # the ramp of the survey's derived KAT
    f = op.fwd(x)
    psi = int(op.table("twid")[n // 2])  # the last stage's first twiddle is psi itself (prime64.rs:188-203)
    assert pow(psi, n, SOLINAS_P) == SOLINAS_P - 1
    check_fwd_definition(64, n, SOLINAS_P, x[0], f[0], psi)
    i = op.inv(f)
    assert all(int(v) == int(x[0][k]) * n % SOLINAS_P for k, v in enumerate(i[0]))
    out.update(c1_x=x, c1_fwd=f, c1_inv=i)
    # headline / C3 transform: Solinas N=2048
    n = 2048
    op = O.OraclePlan(64, n, SOLINAS_P)
    x = rand_below(rng, SOLINAS_P, (2, n), np.uint64)
    f = op.fwd(x)
    out.update(c3_x=x, c3_fwd=f, c3_inv=op.inv(f))
    # C2: prime32 30-bit N=2048: out = inv(acc + fwd(lhs) * rhs)
    op = O.OraclePlan(32, n, P30)
    lhs, rhs, acc = (rand_below(rng, P30, (2, n), np.uint32) for _ in range(3))
    fl = op.fwd(lhs)
    res = np.stack([op.inv(op.mul_accumulate(acc[b].copy(), fl[b], rhs[b])) for b in range(2)])
    out.update(c2_lhs=lhs, c2_rhs=rhs, c2_acc=acc, c2_out=res, c2_fwd=fl)
    # polynomial product through the transform = schoolbook negacyclic convolution (n small enough for Python)
    m = 64
    op64 = O.OraclePlan(64, m, SOLINAS_P)
    a, b = rand_below(rng, SOLINAS_P, (m,), np.uint64), rand_below(rng, SOLINAS_P, (m,), np.uint64)
    fa, fb = op64.fwd(a[None])[0], op64.fwd(b[None])[0]
    prod = op64.inv(op64.mul_assign_normalize(fa.copy(), fb)[None])[0]
    assert [int(v) for v in prod] == negacyclic_schoolbook(a, b, SOLINAS_P)
    out.update(conv_a=a, conv_b=b, conv_prod=prod)
    # C4: native64::Plan32 wrapping-u64 product, N=1024 (checked against the schoolbook product mod 2^64 at n=64)
    pl = O.OracleNativePlan(O.NATIVE64_PLAN32, 64)
    a = rng.integers(0, 1 << 63, size=64, dtype=np.uint64) * np.uint64(2) + np.uint64(1)
    b = rng.integers(0, 1 << 63, size=64, dtype=np.uint64) * np.uint64(2)
    assert [int(v) for v in pl.negacyclic_polymul(a, b)] == negacyclic_schoolbook(a, b, 1 << 64)
    pl = O.OracleNativePlan(O.NATIVE64_PLAN32, 1024)
    a = rng.integers(0, 1 << 63, size=(2, 1024), dtype=np.uint64) * np.uint64(2) + np.uint64(1)
    b = rng.integers(0, 1 << 63, size=(2, 1024), dtype=np.uint64) * np.uint64(3)
    out.update(c4_lhs=a, c4_rhs=b, c4_prod=np.stack([pl.negacyclic_polymul(a[k], b[k]) for k in range(2)]))
    # C5 shape at a size the fixture can hold: Solinas N=8192 (strided pass + single-CTA kernel on the GPU)
    n = 8192
    op = O.OraclePlan(64, n, SOLINAS_P)
    x = rand_below(rng, SOLINAS_P, (1, n), np.uint64)
    f = op.fwd(x)
    out.update(c5_x=x, c5_fwd=f, c5_inv=op.inv(f))
    path = os.path.join(HERE, "ntt_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", ", ".join(sorted(out)))


if __name__ == "__main__":
    main()
