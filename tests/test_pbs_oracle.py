"""CPU tests of the NTT-PBS oracle (oracle/tfhe_ntt_pbs_oracle.c).

Pins: the doc-test vectors the reference holds for the pieces (decomposer.rs:466-485,
polynomial_algorithms.rs:393, :460), the decomposition properties of
commons/math/decomposition/tests.rs (recompose(decompose(x)) == closest_representable(x), digits in
[-B/2, B/2]) and the reference's end-to-end tests restated with our own key generation
(algorithms/test/lwe_programmable_bootstrapping.rs:708-870 classic, :1002-1163 bnf): encrypt,
bootstrap, decrypt, compare with f(msg).
"""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
import pbs_support as S

L = O.lib()
P = O.SOLINAS_P


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def test_closest_representable_doc_vectors():
    # decomposer.rs:466-485: init_decomposer_state(249280154129830) = (32160715112448, Negative) for
    # q = 2^48 -+ 1, so closest_representable (:487-497) is q - 32160715112448
    for q in ((1 << 48) - 1, (1 << 48) + 1):
        assert L.tfo_closest_representable_non_native(249280154129830, 4, 3, q) == q - 32160715112448
        assert L.tfo_closest_representable_non_native(q - 249280154129830, 4, 3, q) == 32160715112448


def test_monomial_doc_vectors():
    # polynomial_algorithms.rs:460: [1,2,3] * X^2 = [254,253,1] (u8)  -> native u64: [-2,-3,1]
    v = np.array([1, 2, 3], dtype=np.uint64)
    L.tfo_monomial_mul_assign(_ptr(v), 3, 2, 0)
    assert list(v.astype(np.int64)) == [-2, -3, 1]
    # polynomial_algorithms.rs:393: [1,2,3] / X^2 = [3,255,254]
    v = np.array([1, 2, 3], dtype=np.uint64)
    L.tfo_monomial_div_assign(_ptr(v), 3, 2, 0)
    assert list(v.astype(np.int64)) == [3, -1, -2]


@pytest.mark.parametrize("modulus", [0, P])
def test_monomial_mul_div_roundtrip_and_full_cycles(modulus):
    rng = np.random.default_rng(1)
    n = 64
    x = rng.integers(0, 1 << 62, n, dtype=np.uint64)
    for d in (0, 1, 17, n - 1, n, n + 5, 2 * n - 1, 2 * n, 3 * n + 2):
        y = x.copy()
        L.tfo_monomial_mul_assign(_ptr(y), n, d, modulus)
        # X^d = schoolbook shift with sign flips
        want = np.zeros(n, dtype=object)
        for j in range(n):
            e = j + d
            v = int(x[j]) if (e // n) % 2 == 0 else -int(x[j])
            want[e % n] = v % (modulus if modulus else 1 << 64)
        assert [int(v) for v in y] == list(want)
        L.tfo_monomial_div_assign(_ptr(y), n, d, modulus)
        assert np.array_equal(y, x)
    y = x.copy()
    L.tfo_monomial_mul_assign(_ptr(y), n, 2 * n, modulus)  # X^(2N) = 1
    assert np.array_equal(y, x)


def test_modulus_switch():
    # fft_impl/common.rs:10-23: rounding to log_modulus bits
    assert L.tfo_modulus_switch(0, 12) == 0
    assert L.tfo_modulus_switch((1 << 64) - 1, 12) == 0  # rounds up and wraps
    assert L.tfo_modulus_switch(1 << 52, 12) == 1
    assert L.tfo_modulus_switch((1 << 51), 12) == 1 and L.tfo_modulus_switch((1 << 51) - 1, 12) == 0
    assert L.tfo_modulus_switch(12345, 64) == 12345
    # ntt64_pbs.rs:540-550: round(x * 2N / p)
    rng = np.random.default_rng(2)
    for x in [0, 1, P - 1, P // 2, P // 4096, P // 4096 + 1] + [int(v) for v in rng.integers(0, 1 << 63, 50)]:
        x %= P
        got = L.tfo_pbs_modulus_switch_non_native(x, 11, P)
        num, den = x << 12, P
        want = num // den + (1 if num % den >= den >> 1 else 0)
        assert got == want and 0 <= got <= 4096


@pytest.mark.parametrize("base_log,level", [(4, 3), (23, 1), (7, 3), (15, 2), (2, 5)])
@pytest.mark.parametrize("modulus", [P, (1 << 48) + 1, (1 << 52) + 1])
def test_non_native_decomposition_properties(base_log, level, modulus):
    """tests.rs: terms within [-B/2, B/2]; recomposition == closest representable"""
    bits = (modulus - 1).bit_length()
    if base_log * level >= bits:
        pytest.skip("decomposed bits exceed the modulus")
    rng = np.random.default_rng(base_log * 100 + level)
    vals = [0, 1, modulus - 1, modulus // 2, modulus // 2 + 1, 1 << (bits - 1), 9223372032559808513 % modulus]
    vals += [int(v) % modulus for v in rng.integers(0, 1 << 63, 200, dtype=np.uint64) * 2]
    x = np.array(vals, dtype=np.uint64)
    states = np.zeros_like(x)
    signs = np.zeros(x.size, dtype=np.uint8)
    L.tfo_decomp_non_native_init(_ptr(x), x.size, base_log, level, modulus, _ptr(states), _ptr(signs))
    half = 1 << (base_log - 1)
    recomposed = [0] * x.size
    shift0 = bits - base_log * level
    for lv in range(level, 0, -1):  # the iterator yields level l first
        term = np.zeros_like(x)
        L.tfo_decomp_non_native_next(_ptr(states), _ptr(signs), x.size, base_log, modulus, _ptr(term))
        for i, t in enumerate(term):
            t = int(t)
            s = t if t <= modulus // 2 else t - modulus
            assert -half <= s <= half
            recomposed[i] = (recomposed[i] + s * (1 << (bits - base_log * lv))) % modulus
    for i, v in enumerate(vals):
        assert recomposed[i] == L.tfo_closest_representable_non_native(v, base_log, level, modulus), (v, shift0)


@pytest.mark.parametrize("base_log,level", [(4, 3), (23, 1), (7, 3), (15, 2), (2, 5)])
def test_native_decomposition_properties(base_log, level):
    rng = np.random.default_rng(7)
    vals = [0, 1, (1 << 64) - 1, 1 << 63, (1 << 63) - 1] + [int(v) for v in rng.integers(0, 1 << 63, 200)] + \
           [int(v) * 2 + 1 for v in rng.integers(0, 1 << 63, 200)]
    x = np.array(vals, dtype=np.uint64)
    states = np.zeros_like(x)
    L.tfo_decomp_native_init(_ptr(x), x.size, base_log, level, _ptr(states))
    half = 1 << (base_log - 1)
    rec = [0] * x.size
    for lv in range(level, 0, -1):
        term = np.zeros_like(x)
        L.tfo_decomp_native_next(_ptr(states), x.size, base_log, _ptr(term))
        for i, t in enumerate(term.astype(np.int64)):
            assert -half <= int(t) <= half
            rec[i] = (rec[i] + int(t) * (1 << (64 - base_log * lv))) % (1 << 64)
    rep = base_log * level
    for i, v in enumerate(vals):
        # closest multiple of 2^(64-rep); ties may go either way (balanced rounding, decomposer.rs:204-236)
        step = 1 << (64 - rep)
        d = (rec[i] - v) % (1 << 64)
        d = d if d < (1 << 63) else d - (1 << 64)
        assert abs(d) <= step // 2 and rec[i] % step == 0


def test_sample_extraction_decrypts_coefficient():
    rng = np.random.default_rng(3)
    for bnf in (False, True):
        prm = S.PbsParams(8, 2, 64, 8, 2, P)
        keys = S.Keys(prm, rng, bnf=bnf, noise=False)
        pt = keys._uniform(prm.N)
        glwe = keys.glwe_encrypt(pt).reshape(-1)
        for nth in (0, 1, 17, prm.N - 1):
            out = np.zeros(prm.k * prm.N + 1, dtype=np.uint64)
            L.tfo_extract_lwe_sample(_ptr(glwe), prm.glwe_size, prm.N, nth, 0 if bnf else P, _ptr(out))
            sk = keys.glwe_sk.reshape(-1)
            dot = sum(int(v) for v in out[:-1][sk == 1])
            assert (int(out[-1]) - dot) % keys.q == int(pt[nth])


SMALL = [
    dict(n_lwe=24, glwe_dim=1, poly_size=256, base_log=15, level=2, modulus=P, msg_bits=2),
    dict(n_lwe=16, glwe_dim=2, poly_size=128, base_log=10, level=3, modulus=P, msg_bits=2),
]


@pytest.mark.parametrize("cfg", SMALL)
@pytest.mark.parametrize("bnf", [False, True])
def test_pbs_encrypt_bootstrap_decrypt(cfg, bnf):
    """lwe_encrypt_pbs_ntt64_decrypt_custom_mod / lwe_encrypt_pbs_ntt64_bnf_decrypt restated"""
    rng = np.random.default_rng(11)
    prm = S.PbsParams(**cfg)
    keys = S.Keys(prm, rng, bnf=bnf)
    std = keys.bootstrap_key()
    ntt_bsk = O.convert_standard_bsk(keys.plan, std, input_width=64 if bnf else 0, normalize=not bnf)
    pbs = O.OraclePbs(keys.plan, ntt_bsk, prm.n_lwe, prm.glwe_size, prm.base_log, prm.level)
    f = (lambda x: x) if bnf else (lambda x: (3 * x + 1) % (1 << prm.msg_bits))
    lut = keys.lut(f)
    for msg in range(1 << prm.msg_bits):
        ct = keys.lwe_encrypt(msg, std=2.0 ** 40)
        out = pbs.pbs_bnf(ct, lut) if bnf else pbs.pbs(ct, lut)
        if not bnf:
            assert int(out.max()) < P
        assert keys.lwe_decrypt_big(out) == f(msg), msg


def test_blind_rotate_is_pbs_without_extraction():
    rng = np.random.default_rng(5)
    prm = S.PbsParams(12, 1, 64, 12, 2, P, msg_bits=2)
    keys = S.Keys(prm, rng)
    ntt_bsk = O.convert_standard_bsk(keys.plan, keys.bootstrap_key(), normalize=True)
    pbs = O.OraclePbs(keys.plan, ntt_bsk, prm.n_lwe, prm.glwe_size, prm.base_log, prm.level)
    lut = keys.lut(lambda x: x)
    ct = keys.lwe_encrypt(2)
    acc = pbs.blind_rotate(ct, lut)
    out = np.zeros(prm.k * prm.N + 1, dtype=np.uint64)
    L.tfo_extract_lwe_sample(_ptr(acc), prm.glwe_size, prm.N, 0, P, _ptr(out))
    assert np.array_equal(out, pbs.pbs(ct, lut))
    # a zero mask element is skipped (ntt64_pbs.rs:257): same result as removing it
    ct0 = ct.copy()
    ct0[3] = 0
    acc0 = pbs.blind_rotate(ct0, lut)
    bsk2 = np.delete(ntt_bsk.reshape(prm.n_lwe, -1), 3, axis=0).reshape(-1)
    pbs2 = O.OraclePbs(keys.plan, bsk2, prm.n_lwe - 1, prm.glwe_size, prm.base_log, prm.level)
    assert np.array_equal(acc0, pbs2.blind_rotate(np.delete(ct0, 3), lut))
