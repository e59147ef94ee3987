"""ctypes binding of the CPU oracle (oracle/tfhe_ntt_oracle.{h,c}) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ORACLE_DIR = os.path.join(_ROOT, "oracle")

SOLINAS_P = (1 << 64) - (1 << 32) + 1

# enum tfo_native_kind
NATIVE32_PLAN32, NATIVE32_PLAN52, NATIVE64_PLAN32, NATIVE64_PLAN52, NATIVE128_PLAN32 = range(5)
NATIVE_BINARY32_PLAN32, NATIVE_BINARY32_PLAN52, NATIVE_BINARY64_PLAN32 = 5, 6, 7
NATIVE_BINARY64_PLAN52, NATIVE_BINARY128_PLAN32 = 8, 9
NATIVE_KIND_NAMES = [
    "native32::Plan32", "native32::Plan52", "native64::Plan32", "native64::Plan52",
    "native128::Plan32", "native_binary32::Plan32", "native_binary32::Plan52",
    "native_binary64::Plan32", "native_binary64::Plan52", "native_binary128::Plan32",
]


_NATIVE_BUILT = [False]


def build(native=False):
    """Compile the oracle with gcc (make); returns the path of the shared object."""
    target = "native" if native else "all"
    name = "libtfhe_ntt_oracle_native.so" if native else "libtfhe_ntt_oracle.so"
    so = os.path.join(_ORACLE_DIR, name)
    src = [os.path.join(_ORACLE_DIR, f) for f in ("tfhe_ntt_oracle.c", "tfhe_ntt_simd.c", "tfhe_ntt_pbs_oracle.c", "tfhe_ntt_custum_radix_oracle.c",
                                                 "tfhe_ntt_oracle.h")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src)
    # the -march=native build is redone once per process: the file may have travelled from a
    # machine with a different CPU
    if stale or (native and not _NATIVE_BUILT[0]):
        subprocess.run(["make", "-C", _ORACLE_DIR, target], check=True, capture_output=True)
        if native:
            _NATIVE_BUILT[0] = True
    return so


class Plan64Struct(C.Structure):
    _fields_ = [("n", C.c_size_t), ("p", C.c_uint64),
                ("twid", C.POINTER(C.c_uint64)), ("twid_shoup", C.POINTER(C.c_uint64)),
                ("inv_twid", C.POINTER(C.c_uint64)), ("inv_twid_shoup", C.POINTER(C.c_uint64)),
                ("use_ifma", C.c_int), ("can_use_fast_reduction_code", C.c_int),
                ("p_barrett", C.c_uint64), ("big_q", C.c_uint64),
                ("n_inv_mod_p", C.c_uint64), ("n_inv_mod_p_shoup", C.c_uint64)]


class Plan32Struct(C.Structure):
    _fields_ = [("n", C.c_size_t), ("p", C.c_uint32),
                ("twid", C.POINTER(C.c_uint32)), ("twid_shoup", C.POINTER(C.c_uint32)),
                ("inv_twid", C.POINTER(C.c_uint32)), ("inv_twid_shoup", C.POINTER(C.c_uint32)),
                ("can_use_fast_reduction_code", C.c_int),
                ("p_barrett", C.c_uint32), ("big_q", C.c_uint32),
                ("n_inv_mod_p", C.c_uint32), ("n_inv_mod_p_shoup", C.c_uint32)]


def _load(native=False):
    lib = C.CDLL(build(native))
    u64, u32, sz, vp, i = C.c_uint64, C.c_uint32, C.c_size_t, C.c_void_p, C.c_int
    P64, P32 = C.POINTER(Plan64Struct), C.POINTER(Plan32Struct)
    sig = {
        "tfo_mul_mod64": (u64, [u64, u64, u64]),
        "tfo_exp_mod64": (u64, [u64, u64, u64]),
        "tfo_exp_mod32": (u32, [u32, u32, u32]),
        "tfo_is_prime64": (i, [u64]),
        "tfo_largest_prime_in_arithmetic_progression64": (i, [u64, u64, u64, u64, C.POINTER(u64)]),
        "tfo_find_primitive_root64": (i, [u64, u64, C.POINTER(u64)]),
        "tfo_find_root_solinas_64": (i, [u64, C.POINTER(u64)]),
        "tfo_bit_rev": (sz, [u32, sz]),
        "tfo_plan64_try_new": (P64, [sz, u64]),
        "tfo_plan64_free": (None, [P64]),
        "tfo_plan64_fwd": (None, [P64, vp]),
        "tfo_plan64_inv": (None, [P64, vp]),
        "tfo_plan64_fwd_generic": (None, [P64, vp]),
        "tfo_plan64_inv_generic": (None, [P64, vp]),
        "tfo_plan64_normalize": (None, [P64, vp, sz]),
        "tfo_plan64_mul_assign_normalize": (None, [P64, vp, vp, sz]),
        "tfo_plan64_mul_accumulate": (None, [P64, vp, vp, vp, sz]),
        "tfo_plan32_try_new": (P32, [sz, u32]),
        "tfo_plan32_free": (None, [P32]),
        "tfo_plan32_fwd": (None, [P32, vp]),
        "tfo_plan32_inv": (None, [P32, vp]),
        "tfo_plan32_fwd_generic": (None, [P32, vp]),
        "tfo_plan32_inv_generic": (None, [P32, vp]),
        "tfo_plan32_normalize": (None, [P32, vp, sz]),
        "tfo_plan32_mul_assign_normalize": (None, [P32, vp, vp, sz]),
        "tfo_plan32_mul_accumulate": (None, [P32, vp, vp, vp, sz]),
        "tfo_negacyclic_convolution_mod64": (None, [sz, u64, vp, vp, vp]),
        "tfo_negacyclic_convolution_mod32": (None, [sz, u32, vp, vp, vp]),
        "tfo_negacyclic_convolution_wrapping_u32": (None, [sz, vp, vp, vp]),
        "tfo_negacyclic_convolution_wrapping_u64": (None, [sz, vp, vp, vp]),
        "tfo_negacyclic_convolution_wrapping_u128": (None, [sz, vp, vp, vp]),
        "tfo_primes32": (u32, [i]),
        "tfo_primes52": (u64, [i]),
        "tfo_native_num_primes": (i, [i]),
        "tfo_native_residue_bytes": (i, [i]),
        "tfo_native_value_bytes": (i, [i]),
        "tfo_native_try_new": (vp, [i, sz]),
        "tfo_native_free": (None, [vp]),
        "tfo_native_fwd": (None, [vp, vp, C.POINTER(vp), i]),
        "tfo_native_inv": (None, [vp, vp, C.POINTER(vp)]),
        "tfo_native_negacyclic_polymul": (None, [vp, vp, vp, vp]),
        "tfo_reconstruct_32bit_012": (u32, [u32, u32, u32]),
        "tfo_reconstruct_32bit_01234_v2": (u64, [u32] * 5),
        "tfo_reconstruct_32bit_01234": (u64, [u32] * 5),
        "tfo_reconstruct_52bit_012": (u64, [u64] * 3),
        "tfo_reconstruct_32bit_01": (u32, [u32, u32]),
        "tfo_reconstruct_32bit_012_u64": (u64, [u32] * 3),
        "tfo_reconstruct_52bit_01_u64": (u64, [u64, u64]),
        "tfo_reconstruct_52bit_01_u32": (u32, [u64, u64]),
        "tfo_reconstruct_52bit_0_u32": (u32, [u64]),
        "tfo_plan64_fwd_batch": (None, [P64, vp, sz, i]),
        "tfo_plan64_inv_batch": (None, [P64, vp, sz, i]),
        "tfo_plan32_fwd_batch": (None, [P32, vp, sz, i]),
        "tfo_plan32_inv_batch": (None, [P32, vp, sz, i]),
        "tfo_ntt64_forward": (None, [P64, vp, vp, i, u32]),
        "tfo_ntt64_add_backward": (None, [P64, vp, vp, i, u32]),
        "tfo_product_try_new": (vp, [sz, u64, C.POINTER(u64), sz]),
        "tfo_product_free": (None, [vp]),
        "tfo_product_ntt_domain_len": (sz, [vp]),
        "tfo_product_fwd": (None, [vp, vp, vp, i, u64]),
        "tfo_product_inv": (None, [vp, vp, vp, i]),
        "tfo_product_mul_assign_normalize": (None, [vp, vp, vp]),
        "tfo_product_normalize": (None, [vp, vp]),
        "tfo_product_mul_accumulate": (None, [vp, vp, vp, vp]),
        "tfo_pbs_modulus_switch_non_native": (u64, [u64, u32, u64]),
        "tfo_modulus_switch": (u64, [u64, u32]),
        "tfo_monomial_mul_assign": (None, [vp, sz, sz, u64]),
        "tfo_monomial_div_assign": (None, [vp, sz, sz, u64]),
        "tfo_closest_representable_non_native": (u64, [u64, u32, u32, u64]),
        "tfo_init_decomposer_state_native": (u64, [u64, u32, u32]),
        "tfo_decomp_non_native_init": (None, [vp, sz, u32, u32, u64, vp, vp]),
        "tfo_decomp_non_native_next": (None, [vp, vp, sz, u32, u64, vp]),
        "tfo_decomp_native_init": (None, [vp, sz, u32, u32, vp]),
        "tfo_decomp_native_next": (None, [vp, sz, u32, vp]),
        "tfo_add_external_product_ntt64_assign": (None, [P64, vp, vp, vp, sz, u32, u32, i, u32]),
        "tfo_cmux_ntt64_assign": (None, [P64, vp, vp, vp, sz, u32, u32, i, u32]),
        "tfo_blind_rotate_ntt64_assign": (None, [P64, vp, sz, sz, u32, u32, vp, vp]),
        "tfo_blind_rotate_ntt64_bnf_assign": (None, [P64, vp, sz, sz, u32, u32, u32, vp, vp]),
        "tfo_extract_lwe_sample": (None, [vp, sz, sz, sz, u64, vp]),
        "tfo_programmable_bootstrap_ntt64": (None, [P64, vp, sz, sz, u32, u32, vp, vp, vp]),
        "tfo_programmable_bootstrap_ntt64_bnf": (None, [P64, vp, sz, sz, u32, u32, u32, vp, vp, vp]),
        "tfo_convert_standard_lwe_bootstrap_key_to_ntt64": (None, [P64, vp, vp, sz, u32, i]),
        "tfo_plan64_fwd_batch_simd": (i, [P64, vp, sz, i]),
        "tfo_plan64_inv_batch_simd": (i, [P64, vp, sz, i]),
        "tfo_plan64_fwd_simd1": (i, [P64, vp]),
        "tfo_plan64_inv_simd1": (i, [P64, vp]),
        "tfo_use_simd_transforms": (None, [i]),
        "tfo_programmable_bootstrap_ntt64_batch": (None, [P64, vp, sz, sz, u32, u32, vp, vp, vp, sz, sz, i]),
    }
    for name, (res, args) in sig.items():
        f = getattr(lib, name)
        f.restype, f.argtypes = res, args
    return lib


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = _load()
    return _LIB


def _ptr(a):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


class OraclePlan:
    """prime32::Plan / prime64::Plan restated (bits = 32 or 64)."""

    def __init__(self, bits, n, p, _lib=None):
        self.lib = _lib or lib()
        self.bits, self.n, self.p = bits, n, p
        self.dtype = np.uint64 if bits == 64 else np.uint32
        self._pfx = "tfo_plan%d_" % bits
        self.h = getattr(self.lib, self._pfx + "try_new")(n, p)
        if not self.h:
            raise ValueError("try_new returned None")

    @staticmethod
    def try_new(bits, n, p):
        try:
            return OraclePlan(bits, n, p)
        except ValueError:
            return None

    def __del__(self):
        if getattr(self, "h", None):
            getattr(self.lib, self._pfx + "free")(self.h)
            self.h = None

    # struct views
    @property
    def s(self):
        return self.h.contents

    def table(self, name):
        ptr = getattr(self.s, name)
        if not ptr:
            return None
        return np.ctypeslib.as_array(ptr, shape=(self.n,)).copy()

    def _each(self, fn, buf):
        buf = np.ascontiguousarray(buf, dtype=self.dtype)
        flat = buf.reshape(-1, self.n)
        f = getattr(self.lib, self._pfx + fn)
        for row in flat:
            f(self.h, _ptr(row))
        return buf

    def fwd(self, buf):
        return self._each("fwd", buf.copy())

    def inv(self, buf):
        return self._each("inv", buf.copy())

    def fwd_generic(self, buf):
        return self._each("fwd_generic", buf.copy())

    def inv_generic(self, buf):
        return self._each("inv_generic", buf.copy())

    def normalize(self, values):
        v = np.ascontiguousarray(values, dtype=self.dtype).copy()
        getattr(self.lib, self._pfx + "normalize")(self.h, _ptr(v), v.size)
        return v

    def mul_assign_normalize(self, lhs, rhs):
        l = np.ascontiguousarray(lhs, dtype=self.dtype).copy()
        r = np.ascontiguousarray(rhs, dtype=self.dtype)
        getattr(self.lib, self._pfx + "mul_assign_normalize")(self.h, _ptr(l), _ptr(r), min(l.size, r.size))
        return l

    def mul_accumulate(self, acc, lhs, rhs):
        a = np.ascontiguousarray(acc, dtype=self.dtype).copy()
        l = np.ascontiguousarray(lhs, dtype=self.dtype)
        r = np.ascontiguousarray(rhs, dtype=self.dtype)
        getattr(self.lib, self._pfx + "mul_accumulate")(self.h, _ptr(a), _ptr(l), _ptr(r),
                                                        min(a.size, l.size, r.size))
        return a

    # tfhe Ntt64View wrappers (ntt64.rs:89-266), u64 plans only
    def ntt64_forward(self, standard, mode=0, width=64):
        st = np.ascontiguousarray(standard, dtype=np.uint64)
        out = np.zeros_like(st)
        for a, b in zip(out.reshape(-1, self.n), st.reshape(-1, self.n)):
            self.lib.tfo_ntt64_forward(self.h, _ptr(a), _ptr(b), mode, width)
        return out

    def ntt64_add_backward(self, standard, ntt, mode=0, width=64):
        st = np.ascontiguousarray(standard, dtype=np.uint64).copy()
        nt = np.ascontiguousarray(ntt, dtype=np.uint64).copy()
        for a, b in zip(st.reshape(-1, self.n), nt.reshape(-1, self.n)):
            self.lib.tfo_ntt64_add_backward(self.h, _ptr(a), _ptr(b), mode, width)
        return st, nt

    def fwd_batch_inplace(self, buf, threads, simd=False):
        """Returns "avx512" when the vectorised Solinas port ran, else "scalar"."""
        if simd and self.bits == 64 and self.lib.tfo_plan64_fwd_batch_simd(self.h, _ptr(buf), buf.size // self.n, threads):
            return "avx512"
        getattr(self.lib, self._pfx + "fwd_batch")(self.h, _ptr(buf), buf.size // self.n, threads)
        return "scalar"

    def inv_batch_inplace(self, buf, threads, simd=False):
        if simd and self.bits == 64 and self.lib.tfo_plan64_inv_batch_simd(self.h, _ptr(buf), buf.size // self.n, threads):
            return "avx512"
        getattr(self.lib, self._pfx + "inv_batch")(self.h, _ptr(buf), buf.size // self.n, threads)
        return "scalar"


def reference_barrett_quirk(bits, p):
    """True for the moduli on which the reference's `mul_accumulate` can return r + p instead of r.

    prime64.rs:733-756 / prime32.rs:606-627 (`BarrettInit*::new`) compute whether the Barrett quotient of a product
    needs one correction step or two; `can_use_fast_reduction_code` (prime64.rs:814-816, prime32.rs:748-749) nevertheless
    enables the one-step code (`mul_accumulate_scalar`, prime64.rs:586-609: one `min(prod, prod - p)`, then
    `min(acc + prod, acc + prod - p)`) for every modulus below 2^W / 3 -- its comment only argues that nothing overflows.
    For a two-step modulus a product can therefore leave as true_prod + p, and the accumulated value as r + p >= p.
    (The 52-bit IFMA path, taken only on CPUs with AVX512-IFMA, is used for single-step moduli only.)"""
    big_q = p.bit_length()
    big_l = big_q + bits - 1
    beta = (1 << big_l) % p
    single_step = beta <= p - (1 << (big_q - 1))
    below_third = p < ((1 << bits) // 3 + 1)
    return below_third and not single_step


def assert_mul_accumulate_matches_reference(bits, p, got, oracle, what=""):
    """GPU `mul_accumulate` against the oracle's literal restatement: identical, except that on the moduli of
    `reference_barrett_quirk` the reference (and so the oracle) may be r + p where the GPU returns the canonical r."""
    got = np.asarray(got).astype(np.uint64)
    oracle = np.asarray(oracle).astype(np.uint64)
    assert int(got.max()) < p, ("non-canonical GPU value", what)
    diff = oracle - got  # wrapping
    bad = np.nonzero(diff)[0]
    if bad.size == 0:
        return 0
    assert reference_barrett_quirk(bits, p), ("mismatch on a modulus without the reference's Barrett quirk", what, p)
    assert (diff[bad] == np.uint64(p)).all(), ("oracle - gpu is neither 0 nor p", what, p)
    return int(bad.size)


def negacyclic_convolution_mod(bits, p, lhs, rhs):
    dt = np.uint64 if bits == 64 else np.uint32
    l = np.ascontiguousarray(lhs, dtype=dt)
    r = np.ascontiguousarray(rhs, dtype=dt)
    out = np.zeros_like(l)
    getattr(lib(), "tfo_negacyclic_convolution_mod%d" % bits)(l.size, p, _ptr(l), _ptr(r), _ptr(out))
    return out


VALUE_DTYPES = {4: np.uint32, 8: np.uint64, 16: np.dtype([("lo", np.uint64), ("hi", np.uint64)])}


def negacyclic_convolution_wrapping(value_bytes, lhs, rhs):
    """lhs/rhs: uint32 / uint64 arrays, or (n,2) uint64 (lo,hi) for u128."""
    l = np.ascontiguousarray(lhs)
    r = np.ascontiguousarray(rhs)
    out = np.zeros_like(l)
    n = l.shape[0]
    name = {4: "u32", 8: "u64", 16: "u128"}[value_bytes]
    getattr(lib(), "tfo_negacyclic_convolution_wrapping_" + name)(n, _ptr(l), _ptr(r), _ptr(out))
    return out


class OracleNativePlan:
    """native{32,64,128}::Plan{32,52} and native_binary*::Plan{32,52} restated."""

    def __init__(self, kind, n):
        self.lib = lib()
        self.kind, self.n = kind, n
        self.num_primes = self.lib.tfo_native_num_primes(kind)
        self.residue_bytes = self.lib.tfo_native_residue_bytes(kind)
        self.value_bytes = self.lib.tfo_native_value_bytes(kind)
        self.rdtype = np.uint32 if self.residue_bytes == 4 else np.uint64
        self.h = self.lib.tfo_native_try_new(kind, n)
        if not self.h:
            raise ValueError("try_new returned None")

    @staticmethod
    def try_new(kind, n):
        try:
            return OracleNativePlan(kind, n)
        except ValueError:
            return None

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.tfo_native_free(self.h)
            self.h = None

    def value_array(self, count=None):
        n = self.n if count is None else count
        if self.value_bytes == 16:
            return np.zeros((n, 2), dtype=np.uint64)
        return np.zeros(n, dtype=VALUE_DTYPES[self.value_bytes])

    def _res_ptrs(self, res):
        arr = (C.c_void_p * self.num_primes)()
        for k in range(self.num_primes):
            arr[k] = res[k].ctypes.data
        return arr

    def fwd(self, value, binary=False):
        value = np.ascontiguousarray(value)
        res = [np.zeros(self.n, dtype=self.rdtype) for _ in range(self.num_primes)]
        self.lib.tfo_native_fwd(self.h, _ptr(value), self._res_ptrs(res), int(binary))
        return res

    def inv(self, residues):
        res = [np.ascontiguousarray(r, dtype=self.rdtype).copy() for r in residues]
        value = self.value_array()
        self.lib.tfo_native_inv(self.h, _ptr(value), self._res_ptrs(res))
        return value, res

    def negacyclic_polymul(self, lhs, rhs):
        lhs = np.ascontiguousarray(lhs)
        rhs = np.ascontiguousarray(rhs)
        prod = self.value_array()
        self.lib.tfo_native_negacyclic_polymul(self.h, _ptr(prod), _ptr(lhs), _ptr(rhs))
        return prod


class OracleProductPlan:
    """product::Plan restated (product.rs:139-967)."""

    def __init__(self, n, modulus, factors):
        self.lib = lib()
        self.n, self.modulus = n, modulus
        f = (C.c_uint64 * len(factors))(*factors)
        self.h = self.lib.tfo_product_try_new(n, modulus, f, len(factors))
        if not self.h:
            raise ValueError("try_new returned None")
        self.domain_len = self.lib.tfo_product_ntt_domain_len(self.h)

    @staticmethod
    def try_new(n, modulus, factors):
        try:
            return OracleProductPlan(n, modulus, factors)
        except ValueError:
            return None

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.tfo_product_free(self.h)
            self.h = None

    def fwd(self, standard, bounded=None):
        st = np.ascontiguousarray(standard, dtype=np.uint64)
        ntt = np.zeros(self.domain_len, dtype=np.uint64)
        self.lib.tfo_product_fwd(self.h, _ptr(ntt), _ptr(st), 0 if bounded is None else 1, bounded or 0)
        return ntt

    def inv(self, ntt, standard=None, accumulate=False):
        ntt = np.ascontiguousarray(ntt, dtype=np.uint64).copy()
        st = np.zeros(self.n, dtype=np.uint64) if standard is None else np.ascontiguousarray(standard, dtype=np.uint64).copy()
        self.lib.tfo_product_inv(self.h, _ptr(st), _ptr(ntt), int(accumulate))
        return st, ntt

    def mul_assign_normalize(self, lhs, rhs):
        l = np.ascontiguousarray(lhs, dtype=np.uint64).copy()
        self.lib.tfo_product_mul_assign_normalize(self.h, _ptr(l), _ptr(np.ascontiguousarray(rhs, dtype=np.uint64)))
        return l

    def normalize(self, values):
        v = np.ascontiguousarray(values, dtype=np.uint64).copy()
        self.lib.tfo_product_normalize(self.h, _ptr(v))
        return v

    def mul_accumulate(self, acc, lhs, rhs):
        a = np.ascontiguousarray(acc, dtype=np.uint64).copy()
        self.lib.tfo_product_mul_accumulate(self.h, _ptr(a), _ptr(np.ascontiguousarray(lhs, dtype=np.uint64)),
                                            _ptr(np.ascontiguousarray(rhs, dtype=np.uint64)))
        return a


class OraclePbs:
    """NTT-PBS restated (oracle/tfhe_ntt_pbs_oracle.c): tfhe ntt64_pbs.rs / ntt64_bnf_pbs.rs.
    bsk: NTT-domain container [n_lwe][level][k+1][k+1][N]."""

    def __init__(self, plan, bsk, n_lwe, glwe_size, base_log, level):
        assert plan.bits == 64
        self.plan, self.lib = plan, plan.lib
        self.bsk = np.ascontiguousarray(bsk, dtype=np.uint64)
        self.n_lwe, self.glwe_size, self.base_log, self.level = n_lwe, glwe_size, base_log, level
        self.n = plan.n
        assert self.bsk.size == n_lwe * level * glwe_size * glwe_size * self.n

    def blind_rotate(self, lwe, lut):
        """classic; returns the rotated copy of lut"""
        out = np.array(lut, dtype=np.uint64)
        lwe = np.ascontiguousarray(lwe, dtype=np.uint64)
        self.lib.tfo_blind_rotate_ntt64_assign(self.plan.h, _ptr(self.bsk), self.n_lwe, self.glwe_size, self.base_log,
                                               self.level, _ptr(lwe), _ptr(out))
        return out

    def blind_rotate_bnf(self, msed, lut, width=64):
        out = np.array(lut, dtype=np.uint64)
        msed = np.ascontiguousarray(msed, dtype=np.uint64)
        self.lib.tfo_blind_rotate_ntt64_bnf_assign(self.plan.h, _ptr(self.bsk), self.n_lwe, self.glwe_size,
                                                   self.base_log, self.level, width, _ptr(msed), _ptr(out))
        return out

    def pbs(self, lwe_in, accumulator):
        lwe_in = np.ascontiguousarray(lwe_in, dtype=np.uint64)
        acc = np.ascontiguousarray(accumulator, dtype=np.uint64)
        out = np.zeros((self.glwe_size - 1) * self.n + 1, dtype=np.uint64)
        self.lib.tfo_programmable_bootstrap_ntt64(self.plan.h, _ptr(self.bsk), self.n_lwe, self.glwe_size,
                                                  self.base_log, self.level, _ptr(lwe_in), _ptr(out), _ptr(acc))
        return out

    def pbs_batch(self, lwe_in, accumulator, threads):
        """classic PBS of lwe_in [batch][n_lwe+1] over `threads` host threads (bench.py's CPU arm);
        accumulator: one LUT [(k+1)N] or one per ciphertext"""
        lwe_in = np.ascontiguousarray(lwe_in, dtype=np.uint64).reshape(-1, self.n_lwe + 1)
        acc = np.ascontiguousarray(accumulator, dtype=np.uint64).reshape(-1, self.glwe_size * self.n)
        batch = lwe_in.shape[0]
        assert acc.shape[0] in (1, batch)
        out = np.zeros((batch, (self.glwe_size - 1) * self.n + 1), dtype=np.uint64)
        self.lib.tfo_programmable_bootstrap_ntt64_batch(self.plan.h, _ptr(self.bsk), self.n_lwe, self.glwe_size,
                                                        self.base_log, self.level, _ptr(lwe_in), _ptr(out), _ptr(acc),
                                                        acc.shape[0], batch, threads)
        return out

    def pbs_bnf(self, lwe_in, accumulator, width=64):
        lwe_in = np.ascontiguousarray(lwe_in, dtype=np.uint64)
        acc = np.ascontiguousarray(accumulator, dtype=np.uint64)
        out = np.zeros((self.glwe_size - 1) * self.n + 1, dtype=np.uint64)
        self.lib.tfo_programmable_bootstrap_ntt64_bnf(self.plan.h, _ptr(self.bsk), self.n_lwe, self.glwe_size,
                                                      self.base_log, self.level, width, _ptr(lwe_in), _ptr(out),
                                                      _ptr(acc))
        return out


def convert_standard_bsk(plan, standard, input_width=0, normalize=False):
    standard = np.ascontiguousarray(standard, dtype=np.uint64)
    out = np.empty_like(standard)
    plan.lib.tfo_convert_standard_lwe_bootstrap_key_to_ntt64(plan.h, _ptr(standard), _ptr(out), standard.size // plan.n,
                                                             input_width, int(normalize))
    return out


# ---- custum_radix (oracle/tfhe_ntt_custum_radix_oracle.c) ------------------------------------------

def cr_tables(n, p):
    """make_twiddles / make_inv_twiddles of custum_radix/fwd.rs:72-103"""
    L = lib()
    L.tfo_cr_make_twiddles.argtypes = [C.c_size_t, C.c_uint32, C.c_void_p]
    L.tfo_cr_make_twiddles.restype = C.c_int
    L.tfo_cr_make_inv_twiddles.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p]
    L.tfo_cr_make_inv_twiddles.restype = None
    tw = np.zeros(n, dtype=np.uint32)
    if not L.tfo_cr_make_twiddles(n, p, tw.ctypes.data):
        raise ValueError("n must be a power of two dividing p - 1")
    inv = np.zeros_like(tw)
    L.tfo_cr_make_inv_twiddles(tw.ctypes.data, n, p, inv.ctypes.data)
    return tw, inv


def cr_fft(kind, a, tw, p):
    """fft_{radix2,radix4,split_radix}_recursive on a copy"""
    L = lib()
    fn = getattr(L, "tfo_cr_fft_%s_recursive" % kind)
    fn.argtypes, fn.restype = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32], None
    out = np.ascontiguousarray(a, dtype=np.uint32).copy()
    fn(out.ctypes.data, out.size, tw.ctypes.data, p)
    return out
