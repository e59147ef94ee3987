"""Key / ciphertext generation for the NTT-PBS tests (TEST INFRASTRUCTURE ONLY).

Plays the role of the reference's key generation in
tfhe/src/core_crypto/algorithms/test/lwe_programmable_bootstrapping.rs:708-870 (classic) and
:1002-1163 (bnf): binary LWE / GLWE secret keys, a standard-domain bootstrap key made of GGSW
encryptions of the LWE key bits (algorithms/ggsw_encryption.rs:20-45, :103-175, :318-370), LWE
encryption / decryption and the PBS look-up table (lwe_programmable_bootstrapping/mod.rs:24-75).
All polynomial arithmetic goes through the CPU oracle; nothing here touches the GPU library.
"""
import numpy as np

import oracle_lib as O

MASK64 = (1 << 64) - 1


class PbsParams:
    def __init__(self, n_lwe, glwe_dim, poly_size, base_log, level, modulus, msg_bits=3, width=64):
        """modulus: the NTT prime.  classic: also the ciphertext modulus; bnf: ciphertexts are modulo
        2^width (MSB aligned)."""
        self.n_lwe, self.k, self.N = n_lwe, glwe_dim, poly_size
        self.base_log, self.level, self.p = base_log, level, modulus
        self.msg_bits, self.width = msg_bits, width
        self.glwe_size = glwe_dim + 1

    @property
    def bsk_len(self):
        return self.n_lwe * self.level * self.glwe_size * self.glwe_size * self.N


# tfhe/src/core_crypto/algorithms/test/mod.rs:106-130
TEST_PARAMS_3_BITS_SOLINAS_U64 = dict(n_lwe=742, glwe_dim=1, poly_size=2048, base_log=23, level=1,
                                      modulus=O.SOLINAS_P, msg_bits=3)


def _polymul_mod(plan, a, b):
    # OraclePlan methods return transformed copies
    x = plan.fwd(np.array(a, dtype=np.uint64))
    y = plan.fwd(np.array(b, dtype=np.uint64))
    return plan.inv(plan.mul_assign_normalize(x, y))


def _add_mod(a, b, p):
    a = a.astype(object)
    b = b.astype(object)
    return np.array([(int(x) + int(y)) % p for x, y in zip(a, b)], dtype=np.uint64)


class Keys:
    """classic (bnf=False): everything modulo the NTT prime p; bnf: modulo 2^width, values MSB
    aligned in u64 (so plain wrapping u64 arithmetic on multiples of 2^(64-width))."""

    def __init__(self, params, rng, bnf=False, noise=True):
        self.P, self.rng, self.bnf, self.noise = params, rng, bnf, noise
        self.plan = O.OraclePlan(64, params.N, params.p)
        self.lwe_sk = rng.integers(0, 2, params.n_lwe, dtype=np.uint64)
        self.glwe_sk = rng.integers(0, 2, (params.k, params.N), dtype=np.uint64)
        self.q = (1 << 64) if bnf else params.p
        self.scale = 1 << (64 - params.width) if bnf else 1  # MSB alignment of 2^width values

    # -- small helpers --------------------------------------------------------------------
    def _uniform(self, shape):
        if self.bnf:
            v = self.rng.integers(0, 1 << 63, shape, dtype=np.uint64) * np.uint64(2) + \
                self.rng.integers(0, 2, shape, dtype=np.uint64)
            return v & np.uint64(MASK64 ^ (self.scale - 1))
        v = self.rng.integers(0, 1 << 63, shape, dtype=np.uint64) * np.uint64(2) + \
            self.rng.integers(0, 2, shape, dtype=np.uint64)
        return np.where(v >= np.uint64(self.P.p), v - np.uint64(self.P.p), v)

    def _noise(self, shape, std):
        if not self.noise:
            return np.zeros(shape, dtype=np.int64)
        return np.rint(self.rng.normal(0.0, std, shape)).astype(np.int64)

    def _to_mod(self, signed):
        """int64 array -> residues (scaled for MSB alignment in bnf)"""
        if self.bnf:
            return (signed.astype(np.uint64) * np.uint64(self.scale))
        return np.array([int(v) % self.P.p for v in signed], dtype=np.uint64).reshape(signed.shape)

    def _polymul(self, a, b):
        if self.bnf:
            return O.negacyclic_convolution_wrapping(8, a, b)
        return _polymul_mod(self.plan, a, b)

    def _add(self, a, b):
        if self.bnf:
            return a + b
        return _add_mod(a, b, self.P.p)

    # -- GLWE / GGSW / bootstrap key -------------------------------------------------------
    def glwe_encrypt(self, plaintext, std=5.0):
        P = self.P
        ct = np.zeros((P.glwe_size, P.N), dtype=np.uint64)
        body = self._add(np.array(plaintext, dtype=np.uint64), self._to_mod(self._noise(P.N, std)))
        for t in range(P.k):
            ct[t] = self._uniform(P.N)
            body = self._add(body, self._polymul(ct[t], self.glwe_sk[t]))
        ct[P.k] = body
        return ct

    def bootstrap_key(self):
        """Standard-domain LweBootstrapKey container [n_lwe][level][k+1][k+1][N]; the first level
        slice is level l (ggsw_encryption.rs:141-150)."""
        P = self.P
        out = np.zeros((P.n_lwe, P.level, P.glwe_size, P.glwe_size, P.N), dtype=np.uint64)
        for i in range(P.n_lwe):
            m = int(self.lwe_sk[i])
            for lv in range(P.level):
                level = P.level - lv
                shift = 64 - P.base_log * level
                factor = (-(m << shift)) % self.q  # ggsw_encryption_multiplicative_factor
                for row in range(P.glwe_size):
                    pt = np.zeros(P.N, dtype=np.uint64)
                    if row < P.k:  # key polynomial times the factor (:332-352)
                        pt = np.array([(int(s) * factor) % self.q for s in self.glwe_sk[row]], dtype=np.uint64)
                    else:  # last row: -factor in the constant coefficient (:353-368)
                        pt[0] = (-factor) % self.q
                    out[i, lv, row] = self.glwe_encrypt(pt)
        return out.reshape(-1)

    # -- LWE ------------------------------------------------------------------------------
    def delta(self):
        P = self.P
        enc = (1 << 63) if self.bnf else P.p // 2  # get_encoding_with_padding, test/mod.rs:478-486
        return enc >> P.msg_bits

    def lwe_encrypt(self, msg, std=2.0 ** 30):
        P = self.P
        a = self._uniform(P.n_lwe)
        e = int(self._to_mod(self._noise(1, std))[0])
        dot = sum(int(x) for x in a[self.lwe_sk == 1])
        body = (dot + msg * self.delta() + e) % self.q
        if self.bnf:
            body &= MASK64 ^ (self.scale - 1)
        return np.concatenate([a, np.array([body], dtype=np.uint64)])

    def lwe_decrypt_big(self, ct):
        """decrypt an output ciphertext (dimension k*N) under the flattened GLWE key, decode"""
        sk = self.glwe_sk.reshape(-1)
        dot = sum(int(x) for x in ct[:-1][sk == 1])
        pt = (int(ct[-1]) - dot) % self.q
        d = self.delta()
        return ((pt + d // 2) // d) % (1 << self.P.msg_bits)  # round_decode, test/mod.rs:488-490

    def glwe_decrypt(self, ct):
        P = self.P
        acc = np.array(ct[P.k], dtype=np.uint64)
        for t in range(P.k):
            prod = self._polymul(ct[t], self.glwe_sk[t])
            if self.bnf:
                acc = acc - prod
            else:
                acc = np.array([(int(x) - int(y)) % P.p for x, y in zip(acc, prod)], dtype=np.uint64)
        return acc

    # -- look-up table ----------------------------------------------------------------------
    def lut(self, f):
        """generate_programmable_bootstrap_glwe_lut (mod.rs:24-75): trivial GLWE of the boxed table"""
        P = self.P
        msg_mod = 1 << P.msg_bits
        box = P.N // msg_mod
        acc = np.zeros(P.N, dtype=object)
        for i in range(msg_mod):
            acc[i * box:(i + 1) * box] = (f(i) * self.delta()) % self.q
        half = box // 2
        for j in range(half):
            acc[j] = (-int(acc[j])) % self.q
        acc = np.roll(acc, -half)
        out = np.zeros((P.glwe_size, P.N), dtype=np.uint64)
        out[P.k] = acc.astype(np.uint64)
        return out.reshape(-1)
