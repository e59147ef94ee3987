"""Small end-to-end run for compute-sanitizer: every kernel family once, tiny sizes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import tfhe_ntt_b200 as T
import oracle_lib as O

rng = np.random.default_rng(0)
for bits, n, p in [(64, 2048, O.SOLINAS_P), (64, 1024, 4611686018427322369), (64, 256, 9223372036853661697),
                   (64, 512, 18446744073707716609), (32, 2048, 1073479681), (32, 4096, 2147352577), (32, 256, 4293918721),
                   (64, 64, O.SOLINAS_P), (64, 8192, O.SOLINAS_P), (32, 65536, 1073479681)]:
    mod = T.prime64 if bits == 64 else T.prime32
    dt = np.uint64 if bits == 64 else np.uint32
    plan, ref = mod.Plan.try_new(n, p), O.OraclePlan(bits, n, p)
    x = (rng.integers(0, 1 << 62, size=(5, n), dtype=np.uint64) % np.uint64(p)).astype(dt)
    y = x.copy(); plan.fwd_batch(y); assert (y == ref.fwd(x)).all()
    plan.inv_batch(y); assert (y == ref.inv(ref.fwd(x))).all()
    a = x[0].copy(); plan.mul_accumulate(a, x[1], x[2]); plan.normalize(a); plan.mul_assign_normalize(a, x[3])
    if 256 <= n <= 4096:
        out = np.zeros_like(x); plan.fwd_mac_inv_batch(out, x, x[:1].copy(), x[1:2].copy())
for kind, cls in [(2, T.native64.Plan32), (4, T.native128.Plan32), (3, T.native64.Plan52), (5, T.native_binary32.Plan32)]:
    n = 1024
    gp, op = cls.try_new(n), O.OracleNativePlan(kind, n)
    vb = op.value_bytes
    raw = rng.integers(0, 1 << 63, size=(n, 2), dtype=np.uint64)
    lhs = np.ascontiguousarray(raw) if vb == 16 else raw[:, 0].astype(O.VALUE_DTYPES[vb])
    rhs = (rng.integers(0, 2, size=n, dtype=np.uint64)).astype(O.VALUE_DTYPES[vb]) if vb != 16 else np.stack([rng.integers(0, 2, size=n, dtype=np.uint64), np.zeros(n, dtype=np.uint64)], 1)
    prod = op.value_array(); gp.negacyclic_polymul(prod, lhs, rhs)
    assert (prod == op.negacyclic_polymul(lhs, rhs)).all()
print("sanitize smoke ok")
