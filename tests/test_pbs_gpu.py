"""GPU parity of the NTT programmable bootstrap (through the C ABI) against the CPU oracle.

Bit-exact comparison on random keys / ciphertexts for both variants (classic = ntt64_pbs.rs, bnf =
ntt64_bnf_pbs.rs), the fused and the composed device paths, and the reference's own end-to-end
tests (algorithms/test/lwe_programmable_bootstrapping.rs:708-870, :1002-1163) at the reference's
parameters TEST_PARAMS_3_BITS_SOLINAS_U64 (test/mod.rs:106-130): encrypt, bootstrap, decrypt.
"""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
import pbs_support as S

P = O.SOLINAS_P


def _rand_mod(rng, shape, p):
    v = rng.integers(0, 1 << 63, shape, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, shape, dtype=np.uint64)
    return (v % np.uint64(p)).astype(np.uint64)


def _rand_u64(rng, shape, width=64):
    v = rng.integers(0, 1 << 63, shape, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, shape, dtype=np.uint64)
    return v & np.uint64(((1 << 64) - 1) ^ ((1 << (64 - width)) - 1))


def _setup(rng, n_lwe, k, N, base_log, level, p=P):
    import tfhe_ntt_b200 as T
    from tfhe_ntt_b200 import ntt64_pbs as G
    plan = T.prime64.Plan.try_new(N, p)
    oplan = O.OraclePlan(64, N, p)
    gs = k + 1
    bsk = _rand_mod(rng, n_lwe * level * gs * gs * N, p)
    key = G.NttLweBootstrapKey.from_container(plan, bsk, n_lwe, gs, base_log, level)
    opbs = O.OraclePbs(oplan, bsk, n_lwe, gs, base_log, level)
    return G, key, opbs


SHAPES = [
    # n_lwe, k, N, base_log, level
    (6, 1, 64, 12, 2),
    (5, 1, 256, 23, 1),
    (4, 2, 512, 9, 3),
    (3, 1, 1024, 15, 2),
    (5, 1, 2048, 23, 1),
    (3, 3, 2048, 10, 2),
    (2, 1, 4096, 22, 1),
    (2, 1, 8192, 15, 2),
]


@pytest.mark.gpu
@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("path", [0, 2])
def test_blind_rotate_classic_bit_exact(shape, path):
    n_lwe, k, N, base_log, level = shape
    rng = np.random.default_rng(hash(shape) % 1000)
    G, key, opbs = _setup(rng, *shape)
    batch = 3
    lwe = _rand_mod(rng, (batch, n_lwe + 1), P)
    lwe[1, 0] = 0  # skipped CMUX (ntt64_pbs.rs:257)
    lwe[2, n_lwe] = P - 1  # body that switches to 2N
    lut = _rand_mod(rng, (batch, (k + 1) * N), P)
    want = np.stack([opbs.blind_rotate(lwe[b], lut[b]) for b in range(batch)])
    got = lut.copy()
    G.blind_rotate_ntt64_assign(lwe.reshape(-1), got.reshape(-1), key, path=path)
    assert np.array_equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("path", [0, 2])
@pytest.mark.parametrize("width", [64, 40])
def test_blind_rotate_bnf_bit_exact(shape, path, width):
    n_lwe, k, N, base_log, level = shape
    if base_log * level > width:
        pytest.skip("decomposition wider than the ciphertext modulus")
    rng = np.random.default_rng(hash(shape) % 1000 + width)
    G, key, opbs = _setup(rng, *shape)
    batch = 3
    msed = rng.integers(0, 2 * N, (batch, n_lwe + 1), dtype=np.uint64)
    msed[1, 0] = 0
    lut = _rand_u64(rng, (batch, (k + 1) * N), width)
    want = np.stack([opbs.blind_rotate_bnf(msed[b], lut[b], width) for b in range(batch)])
    got = lut.copy()
    G.blind_rotate_ntt64_bnf_assign(msed.reshape(-1), got.reshape(-1), key, width, path=path)
    assert np.array_equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("bnf", [False, True])
def test_pbs_random_bit_exact_shared_and_per_ct_lut(bnf):
    shape = (7, 1, 512, 11, 2)
    n_lwe, k, N, base_log, level = shape
    rng = np.random.default_rng(21)
    G, key, opbs = _setup(rng, *shape)
    batch = 5
    lwe = _rand_u64(rng, (batch, n_lwe + 1)) if bnf else _rand_mod(rng, (batch, n_lwe + 1), P)
    luts = _rand_u64(rng, (batch, (k + 1) * N)) if bnf else _rand_mod(rng, (batch, (k + 1) * N), P)
    run = G.programmable_bootstrap_ntt64_bnf_lwe_ciphertext if bnf else G.programmable_bootstrap_ntt64_lwe_ciphertext
    ref = opbs.pbs_bnf if bnf else opbs.pbs
    out = np.zeros((batch, k * N + 1), dtype=np.uint64)
    run(lwe.reshape(-1), out.reshape(-1), luts.reshape(-1), key)
    assert np.array_equal(out, np.stack([ref(lwe[b], luts[b]) for b in range(batch)]))
    out1 = np.zeros_like(out)
    run(lwe.reshape(-1), out1.reshape(-1), luts[0], key)
    assert np.array_equal(out1, np.stack([ref(lwe[b], luts[0]) for b in range(batch)]))
    with pytest.raises(AssertionError):  # accumulator count must be 1 or the batch
        run(lwe.reshape(-1), out.reshape(-1), luts[:2].reshape(-1), key)
    with pytest.raises(AssertionError):  # output size mismatch (sample extraction assertion)
        run(lwe.reshape(-1), out.reshape(-1)[:-1], luts[0], key)


@pytest.mark.gpu
@pytest.mark.parametrize("input_width,normalize", [(0, True), (0, False), (64, False), (37, True)])
def test_bootstrap_key_conversion(input_width, normalize):
    """convert_standard_lwe_bootstrap_key_to_ntt64, lwe_bootstrap_key_conversion.rs:294-363"""
    import tfhe_ntt_b200 as T
    from tfhe_ntt_b200 import ntt64_pbs as G
    rng = np.random.default_rng(4)
    n_lwe, gs, N, level = 5, 2, 1024, 2
    plan = T.prime64.Plan.try_new(N, P)
    oplan = O.OraclePlan(64, N, P)
    std = _rand_u64(rng, n_lwe * level * gs * gs * N, input_width) if input_width else \
        _rand_mod(rng, n_lwe * level * gs * gs * N, P)
    want = O.convert_standard_bsk(oplan, std, input_width, normalize)
    out = np.zeros_like(std)
    G.convert_standard_lwe_bootstrap_key_to_ntt64(plan, std, out, G.NORMALIZE if normalize else G.RAW, input_width)
    assert np.array_equal(out, want)
    key = G.NttLweBootstrapKey.from_standard(plan, std, n_lwe, gs, 9, level, input_width,
                                             G.NORMALIZE if normalize else G.RAW)
    assert np.array_equal(key.as_container(), want)
    assert (key.input_lwe_dimension(), key.glwe_size(), key.polynomial_size(), key.decomposition_base_log(),
            key.decomposition_level_count(), key.output_lwe_dimension()) == (n_lwe, gs, N, 9, level, N)


@pytest.mark.gpu
@pytest.mark.parametrize("bnf", [False, True])
def test_pbs_reference_parameters_encrypt_bootstrap_decrypt(bnf):
    """lwe_encrypt_pbs_ntt64_decrypt_custom_mod (:708-865) / lwe_encrypt_pbs_ntt64_bnf_decrypt
    (:1002-1163) at TEST_PARAMS_3_BITS_SOLINAS_U64: n_lwe 742, k 1, N 2048, base_log 23, level 1."""
    import tfhe_ntt_b200 as T
    from tfhe_ntt_b200 import ntt64_pbs as G
    rng = np.random.default_rng(2048)
    prm = S.PbsParams(**S.TEST_PARAMS_3_BITS_SOLINAS_U64)
    keys = S.Keys(prm, rng, bnf=bnf)
    std = keys.bootstrap_key()
    plan = T.prime64.Plan.try_new(prm.N, prm.p)
    key = G.NttLweBootstrapKey.from_standard(plan, std, prm.n_lwe, prm.glwe_size, prm.base_log, prm.level,
                                             64 if bnf else 0, G.RAW if bnf else G.NORMALIZE)
    ntt_bsk = O.convert_standard_bsk(keys.plan, std, 64 if bnf else 0, not bnf)
    assert np.array_equal(key.as_container(), ntt_bsk)
    opbs = O.OraclePbs(keys.plan, ntt_bsk, prm.n_lwe, prm.glwe_size, prm.base_log, prm.level)
    msg_mod = 1 << prm.msg_bits
    f = (lambda x: x) if bnf else (lambda x: x % msg_mod)
    lut = keys.lut(f)
    msgs = list(range(msg_mod)) * 2
    # the reference's LWE noise: std 7.07e-6 * q
    cts = np.stack([keys.lwe_encrypt(m, std=0.000007069849454709433 * 2.0 ** 64) for m in msgs])
    out = np.zeros((len(msgs), prm.k * prm.N + 1), dtype=np.uint64)
    run = G.programmable_bootstrap_ntt64_bnf_lwe_ciphertext if bnf else G.programmable_bootstrap_ntt64_lwe_ciphertext
    run(cts.reshape(-1), out.reshape(-1), lut, key)
    if not bnf:
        assert int(out.max()) < prm.p  # check_encrypted_content_respects_mod
    for i, m in enumerate(msgs):
        assert keys.lwe_decrypt_big(out[i]) == f(m)
    ref = opbs.pbs_bnf if bnf else opbs.pbs
    for i in (0, 5, 11):
        assert np.array_equal(out[i], ref(cts[i], lut))
    # composed path gives the same bits
    out2 = np.zeros((2, prm.k * prm.N + 1), dtype=np.uint64)
    run(cts[:2].reshape(-1), out2.reshape(-1), lut, key, path=G.PATH_COMPOSED)
    assert np.array_equal(out2, out[:2])


@pytest.mark.gpu
def test_pbs_other_prime_family_composed():
    """the composed path runs on any prime64 plan (here a 62-bit Shoup prime)"""
    import tfhe_ntt_b200 as T
    p = T.prime.largest_prime_in_arithmetic_progression64(1 << 16, 1, 0, 1 << 62)
    shape = (4, 1, 256, 14, 2)
    rng = np.random.default_rng(8)
    G, key, opbs = _setup(rng, *shape, p=p)
    n_lwe, k, N, _, _ = shape
    lwe = _rand_mod(rng, (2, n_lwe + 1), p)
    lut = _rand_mod(rng, (2, (k + 1) * N), p)
    want = np.stack([opbs.blind_rotate(lwe[b], lut[b]) for b in range(2)])
    got = lut.copy()
    G.blind_rotate_ntt64_assign(lwe.reshape(-1), got.reshape(-1), key)
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_fused_only_path_and_its_limits():
    import tfhe_ntt_b200 as T
    rng = np.random.default_rng(9)
    # fused kernel exists for Solinas, N = 256 .. 4096, glwe_size 2 .. 4
    shape = (4, 1, 1024, 12, 2)
    G, key, opbs = _setup(rng, *shape)
    lwe = _rand_mod(rng, (2, 5), P)
    lut = _rand_mod(rng, (2, 2 * 1024), P)
    got = lut.copy()
    G.blind_rotate_ntt64_assign(lwe.reshape(-1), got.reshape(-1), key, path=G.PATH_FUSED)
    assert np.array_equal(got, np.stack([opbs.blind_rotate(lwe[b], lut[b]) for b in range(2)]))
    # none for N = 64: PATH_FUSED fails loudly, PATH_AUTO composes
    G, key, opbs = _setup(rng, 3, 1, 64, 12, 2)
    lwe = _rand_mod(rng, 4, P)
    lut = _rand_mod(rng, 128, P)
    with pytest.raises(T.NttB200Error):
        G.blind_rotate_ntt64_assign(lwe, lut.copy(), key, path=G.PATH_FUSED)


@pytest.mark.gpu
def test_cpp_host_mirror_pbs_example(tmp_path):
    import tfhe_ntt_b200 as T
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "example_pbs"
    libdir = os.path.dirname(T.library_path())
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(root, "include"),
                    os.path.join(root, "tests", "cpp", "example_pbs.cpp"), "-o", str(exe),
                    "-L", libdir, "-ltfhe_ntt_b200", "-Wl,-rpath," + libdir], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "cpp pbs example ok" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("N", [512, 1024, 2048, 4096])
@pytest.mark.parametrize("bnf", [False, True])
def test_cluster_latency_path_bit_exact(N, bnf):
    """two-CTA cluster kernel (k = 1, level = 1): same bits as the oracle; zero mask elements skipped"""
    import tfhe_ntt_b200 as T
    n_lwe = 9
    rng = np.random.default_rng(N + bnf)
    G, key, opbs = _setup(rng, n_lwe, 1, N, 23, 1)
    batch = 3
    if bnf:
        lwe = rng.integers(0, 2 * N, (batch, n_lwe + 1), dtype=np.uint64)
        lut = _rand_u64(rng, (batch, 2 * N), 64)
    else:
        lwe = _rand_mod(rng, (batch, n_lwe + 1), P)
        lut = _rand_mod(rng, (batch, 2 * N), P)
    lwe[1, 0] = 0
    lwe[1, 1] = 0
    lwe[2, 4] = 0
    got = lut.copy()
    if bnf:
        G.blind_rotate_ntt64_bnf_assign(lwe.reshape(-1), got.reshape(-1), key, 64, path=G.PATH_CLUSTER)
        want = np.stack([opbs.blind_rotate_bnf(lwe[b], lut[b], 64) for b in range(batch)])
    else:
        G.blind_rotate_ntt64_assign(lwe.reshape(-1), got.reshape(-1), key, path=G.PATH_CLUSTER)
        want = np.stack([opbs.blind_rotate(lwe[b], lut[b]) for b in range(batch)])
    assert np.array_equal(got, want)
    # shapes without a cluster kernel fail loudly when it is requested explicitly
    G2, key2, _ = _setup(rng, 3, 1, 512, 10, 2)
    with pytest.raises(T.NttB200Error):
        G2.blind_rotate_ntt64_assign(_rand_mod(rng, 4, P), _rand_mod(rng, 1024, P), key2, path=G2.PATH_CLUSTER)


@pytest.mark.gpu
def test_classic_pbs_rejects_more_decomposition_bits_than_the_modulus_has():
    """base_log * level > ceil(log2 p): the reference's SignedDecomposerNonNative::new asserts
    (decomposer.rs:487-520); the classic entry points answer ERR_ARG instead of shifting by a negative amount.
    The bnf variant (native decomposer over 64 bits) accepts the same key."""
    import tfhe_ntt_b200 as T
    from tfhe_ntt_b200 import ntt64_pbs as G
    from tfhe_ntt_b200._binding import NttB200Error
    p = T.prime.largest_prime_in_arithmetic_progression64(1 << 16, 1, 0, 1 << 62)  # 62-bit prime
    n, n_lwe, gs, base_log, level = 256, 3, 2, 21, 3                               # 63 bits > 62
    plan = T.prime64.Plan.try_new(n, p)
    rng = np.random.default_rng(10)
    bsk = _rand_mod(rng, n_lwe * level * gs * gs * n, p)
    key = G.NttLweBootstrapKey.from_container(plan, bsk, n_lwe, gs, base_log, level)
    lwe = _rand_mod(rng, (1, n_lwe + 1), p)
    lut = _rand_mod(rng, (1, gs * n), p)
    with pytest.raises(NttB200Error):
        G.blind_rotate_ntt64_assign(lwe.reshape(-1), lut.copy().reshape(-1), key)
    out = np.zeros((gs - 1) * n + 1, dtype=np.uint64)
    with pytest.raises(NttB200Error):
        G.programmable_bootstrap_ntt64_lwe_ciphertext(lwe.reshape(-1), out, lut.reshape(-1), key)


def _random_shapes(seed, count):
    rng = np.random.default_rng(seed)
    shapes = []
    while len(shapes) < count:
        k = int(rng.integers(1, 4))
        N = int(rng.choice([256, 512, 1024, 2048]))
        level = int(rng.integers(1, 7))
        base_log = int(rng.integers(1, 63 // level + 1))  # the decomposers need base_log * level < 64
        if k == 3 and N == 2048 and level > 2:
            continue  # keep the oracle's share of the run time small
        shapes.append((int(rng.integers(2, 5)), k, N, base_log, level))
    # decomposition widths at the edges: one bit, 63 bits (the most the decomposers accept), many levels
    shapes += [(3, 1, 512, 1, 1), (3, 1, 512, 62, 1), (3, 1, 256, 31, 2), (3, 1, 256, 1, 8), (2, 2, 256, 7, 9),
               (3, 1, 1024, 63, 1), (3, 1, 2048, 21, 3), (2, 1, 4096, 15, 4)]
    return shapes


@pytest.mark.gpu
@pytest.mark.parametrize("shape", _random_shapes(5, 16))
def test_blind_rotate_random_decomposition_shapes(shape):
    """Random (k, N, base_log, level) within what the decomposers accept, classic and bnf, every device path that
    takes the shape (auto / composed; identical bits), against the oracle."""
    n_lwe, k, N, base_log, level = shape
    rng = np.random.default_rng(sum(shape))
    G, key, opbs = _setup(rng, *shape)
    batch = 2
    lwe = _rand_mod(rng, (batch, n_lwe + 1), P)
    lwe[0, 0] = 0
    lut = _rand_mod(rng, (batch, (k + 1) * N), P)
    want = np.stack([opbs.blind_rotate(lwe[b], lut[b]) for b in range(batch)])
    for path in (0, 2):
        got = lut.copy()
        G.blind_rotate_ntt64_assign(lwe.reshape(-1), got.reshape(-1), key, path=path)
        assert np.array_equal(got, want), (shape, path)
    for width in (64, max(base_log * level, 33)):
        if base_log * level > width:
            continue
        msed = rng.integers(0, 2 * N, (batch, n_lwe + 1), dtype=np.uint64)
        lutb = _rand_u64(rng, (batch, (k + 1) * N), width)
        wantb = np.stack([opbs.blind_rotate_bnf(msed[b], lutb[b], width) for b in range(batch)])
        for path in (0, 2):
            got = lutb.copy()
            G.blind_rotate_ntt64_bnf_assign(msed.reshape(-1), got.reshape(-1), key, width, path=path)
            assert np.array_equal(got, wantb), (shape, width, path)
