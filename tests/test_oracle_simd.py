"""The AVX-512 port used by the CPU-baseline legs of bench.py gives the scalar oracle's results
(the reference tests its SIMD paths against the scalar ones the same way,
generic_solinas.rs:1637-1826)."""
import numpy as np
import pytest

import oracle_lib as O


@pytest.mark.parametrize("n", [16, 32, 64, 256, 2048, 4096])
def test_simd_equals_scalar(n):
    lib = O._load(native=True)
    plan = O.OraclePlan(64, n, O.SOLINAS_P, _lib=lib)
    rng = np.random.default_rng(n)
    x = (rng.integers(0, 1 << 63, size=(6, n), dtype=np.uint64) * 2 + rng.integers(0, 2, size=(6, n), dtype=np.uint64))
    x %= np.uint64(O.SOLINAS_P)
    x[0, :] = 0
    x[1, :] = O.SOLINAS_P - 1
    want_f = plan.fwd(x)
    got = x.copy()
    isa = plan.fwd_batch_inplace(got, 2, simd=True)
    if isa != "avx512":
        pytest.skip("no AVX-512 on this host")
    assert (got == want_f).all()
    assert plan.inv_batch_inplace(got, 2, simd=True) == "avx512"
    assert (got == plan.inv(want_f)).all()


@pytest.mark.parametrize("simd", [0, 1])
def test_pbs_batch_equals_single_pbs(simd):
    """bench.py's CPU arm for the PBS (threads over ciphertexts, optionally the vectorised transforms)
    returns what the scalar single-ciphertext restatement returns."""
    lib = O._load(native=True)
    n, n_lwe, gs, base_log, level = 256, 6, 2, 12, 2
    plan = O.OraclePlan(64, n, O.SOLINAS_P, _lib=lib)
    rng = np.random.default_rng(5 + simd)
    p = np.uint64(O.SOLINAS_P)
    bsk = (rng.integers(0, 1 << 63, n_lwe * level * gs * gs * n, dtype=np.uint64) * np.uint64(2)) % p
    pbs = O.OraclePbs(plan, bsk, n_lwe, gs, base_log, level)
    lwe = (rng.integers(0, 1 << 63, (5, n_lwe + 1), dtype=np.uint64) * np.uint64(2)) % p
    lut = (rng.integers(0, 1 << 63, gs * n, dtype=np.uint64) * np.uint64(2)) % p
    want = np.stack([pbs.pbs(lwe[b], lut) for b in range(5)])
    lib.tfo_use_simd_transforms(simd)
    try:
        got = pbs.pbs_batch(lwe, lut, threads=3)
    finally:
        lib.tfo_use_simd_transforms(0)
    assert (got == want).all()
