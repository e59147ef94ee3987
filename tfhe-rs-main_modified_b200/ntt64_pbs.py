"""NTT programmable bootstrap on top of prime64.Plan -- the caller of the hot path.

Mirrors tfhe/src/core_crypto/algorithms/lwe_programmable_bootstrapping/ntt64_pbs.rs ("classic":
ciphertext modulus = NTT prime) and ntt64_bnf_pbs.rs ("bnf": ciphertexts modulo 2^width, MSB
aligned), the key entity tfhe/src/core_crypto/entities/ntt_lwe_bootstrap_key.rs and the key
conversion tfhe/src/core_crypto/algorithms/lwe_bootstrap_key_conversion.rs:294-447.

Containers are flat numpy uint64 arrays in the reference's layout: lwe [n_lwe+1] (mask, body),
glwe / lut / accumulator [(k+1)*N], NTT key [n_lwe][level][k+1][k+1][N].  Every function also
accepts a leading batch dimension (new; the reference handles one ciphertext per call).
"""
import numpy as np

from . import _binding as B
from . import prime64

PATH_AUTO, PATH_FUSED, PATH_COMPOSED, PATH_CLUSTER = 0, 1, 2, 3

# NttLweBootstrapKeyOption (lwe_bootstrap_key_conversion.rs:283-288)
RAW, NORMALIZE = 0, 1


class NttLweBootstrapKey:
    """Device-resident NttLweBootstrapKey (entities/ntt_lwe_bootstrap_key.rs:26-33)."""

    def __init__(self, handle, plan):
        self._h = handle
        self._plan = plan  # keeps the plan (and its device tables) alive
        self._L = B.lib()

    @classmethod
    def from_container(cls, plan, container, input_lwe_dimension, glwe_size, decomposition_base_log,
                       decomposition_level_count):
        """NttLweBootstrapKey::from_container (:68-110): `container` is already in the NTT domain."""
        _check_plan(plan)
        L = B.lib()
        n = plan.ntt_size()
        want = input_lwe_dimension * decomposition_level_count * glwe_size * glwe_size * n
        container = np.ascontiguousarray(container, dtype=np.uint64)
        if container.size != want:
            raise AssertionError("NttLweBootstrapKey container length %d, expected %d" % (container.size, want))
        h = B.C.c_void_p()
        B.check(L.ntt_b200_bsk_new(plan._h, B.host_ptr(container, np.uint64), input_lwe_dimension, glwe_size,
                                   decomposition_base_log, decomposition_level_count, B.C.byref(h)),
                "in NttLweBootstrapKey.from_container")
        return cls(h, plan)

    @classmethod
    def from_standard(cls, plan, standard_bsk, input_lwe_dimension, glwe_size, decomposition_base_log,
                      decomposition_level_count, input_modulus_width=0, option=RAW):
        """convert_standard_lwe_bootstrap_key_to_ntt64 (:294-363) straight into device memory."""
        _check_plan(plan)
        L = B.lib()
        n = plan.ntt_size()
        want = input_lwe_dimension * decomposition_level_count * glwe_size * glwe_size * n
        standard_bsk = np.ascontiguousarray(standard_bsk, dtype=np.uint64)
        if standard_bsk.size != want:
            raise AssertionError("LweBootstrapKey container length %d, expected %d" % (standard_bsk.size, want))
        h = B.C.c_void_p()
        B.check(L.ntt_b200_bsk_convert_new(plan._h, B.host_ptr(standard_bsk, np.uint64), input_lwe_dimension,
                                           glwe_size, decomposition_base_log, decomposition_level_count,
                                           input_modulus_width, int(option), B.C.byref(h)),
                "in NttLweBootstrapKey.from_standard")
        return cls(h, plan)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.ntt_b200_bsk_free(h)

    def input_lwe_dimension(self):
        return self._L.ntt_b200_bsk_input_lwe_dimension(self._h)

    def glwe_size(self):
        return self._L.ntt_b200_bsk_glwe_size(self._h)

    def polynomial_size(self):
        return self._L.ntt_b200_bsk_polynomial_size(self._h)

    def decomposition_base_log(self):
        return self._L.ntt_b200_bsk_decomposition_base_log(self._h)

    def decomposition_level_count(self):
        return self._L.ntt_b200_bsk_decomposition_level_count(self._h)

    def ciphertext_modulus(self):
        return self._plan.modulus()

    def output_lwe_dimension(self):
        return (self.glwe_size() - 1) * self.polynomial_size()

    def as_container(self):
        """The NTT-domain container, copied back to the host."""
        n = (self.input_lwe_dimension() * self.decomposition_level_count() * self.glwe_size() ** 2
             * self.polynomial_size())
        out = np.empty(n, dtype=np.uint64)
        B.check(self._L.ntt_b200_bsk_read(self._h, B.host_ptr(out, np.uint64, True), n))
        return out

    def device_data(self):
        return self._L.ntt_b200_bsk_device_data(self._h)


def _check_plan(plan):
    if not isinstance(plan, prime64.Plan):
        raise TypeError("the NTT-PBS runs on a prime64.Plan")


def convert_standard_lwe_bootstrap_key_to_ntt64(plan, input_bsk, output_bsk, option, input_modulus_width=0):
    """lwe_bootstrap_key_conversion.rs:294-363 (host to host): output_bsk receives the NTT-domain
    container.  input_modulus_width = 0 when input_bsk is modulo the NTT prime."""
    _check_plan(plan)
    if input_bsk.size != output_bsk.size:
        raise AssertionError("Mismatched bootstrap key sizes")
    B.check(B.lib().ntt_b200_convert_standard_lwe_bootstrap_key_to_ntt64(
        plan._h, B.host_ptr(input_bsk, np.uint64), B.host_ptr(output_bsk, np.uint64, True), input_bsk.size,
        input_modulus_width, int(option)), "in convert_standard_lwe_bootstrap_key_to_ntt64")


par_convert_standard_lwe_bootstrap_key_to_ntt64 = convert_standard_lwe_bootstrap_key_to_ntt64  # :365-447


def _batch_of(arr, row):
    if arr.size % row:
        raise AssertionError("container length %d is not a multiple of %d" % (arr.size, row))
    return arr.size // row


def blind_rotate_ntt64_assign(input, lut, bsk, path=PATH_AUTO):
    """ntt64_pbs.rs:175-286: lut (one GLWE per input ciphertext) is rotated in place."""
    lwe_size = bsk.input_lwe_dimension() + 1
    batch = _batch_of(input, lwe_size)
    if lut.size != batch * bsk.glwe_size() * bsk.polynomial_size():
        raise AssertionError("lut size does not match the bootstrap key")
    B.check(bsk._L.ntt_b200_blind_rotate_ntt64_assign(bsk._h, B.host_ptr(input, np.uint64),
                                                      B.host_ptr(lut, np.uint64, True), batch, path),
            "in blind_rotate_ntt64_assign")


def blind_rotate_ntt64_bnf_assign(msed_input, lut, bsk, ciphertext_modulus_width=64, path=PATH_AUTO):
    """ntt64_bnf_pbs.rs:174-276: msed_input holds the modulus-switched mask and body."""
    lwe_size = bsk.input_lwe_dimension() + 1
    batch = _batch_of(msed_input, lwe_size)
    if lut.size != batch * bsk.glwe_size() * bsk.polynomial_size():
        raise AssertionError("lut size does not match the bootstrap key")
    B.check(bsk._L.ntt_b200_blind_rotate_ntt64_bnf_assign(bsk._h, ciphertext_modulus_width,
                                                          B.host_ptr(msed_input, np.uint64),
                                                          B.host_ptr(lut, np.uint64, True), batch, path),
            "in blind_rotate_ntt64_bnf_assign")


def _pbs_args(input, output, accumulator, bsk):
    lwe_size = bsk.input_lwe_dimension() + 1
    batch = _batch_of(input, lwe_size)
    if output.size != batch * (bsk.output_lwe_dimension() + 1):
        raise AssertionError("output size does not match the bootstrap key")  # sample extraction :105-109
    acc_count = _batch_of(accumulator, bsk.glwe_size() * bsk.polynomial_size())
    return batch, acc_count


def programmable_bootstrap_ntt64_lwe_ciphertext(input, output, accumulator, bsk, path=PATH_AUTO):
    """ntt64_pbs.rs:439-538.  accumulator: one GLWE for the whole batch, or one per input."""
    batch, acc_count = _pbs_args(input, output, accumulator, bsk)
    B.check(bsk._L.ntt_b200_programmable_bootstrap_ntt64(bsk._h, B.host_ptr(input, np.uint64),
                                                         B.host_ptr(output, np.uint64, True),
                                                         B.host_ptr(accumulator, np.uint64), acc_count, batch, path),
            "in programmable_bootstrap_ntt64_lwe_ciphertext")


def programmable_bootstrap_ntt64_bnf_lwe_ciphertext(input, output, accumulator, bsk, ciphertext_modulus_width=64,
                                                    path=PATH_AUTO):
    """ntt64_bnf_pbs.rs:428-539."""
    batch, acc_count = _pbs_args(input, output, accumulator, bsk)
    B.check(bsk._L.ntt_b200_programmable_bootstrap_ntt64_bnf(bsk._h, ciphertext_modulus_width,
                                                             B.host_ptr(input, np.uint64),
                                                             B.host_ptr(output, np.uint64, True),
                                                             B.host_ptr(accumulator, np.uint64), acc_count, batch,
                                                             path),
            "in programmable_bootstrap_ntt64_bnf_lwe_ciphertext")


# device-resident forms
def blind_rotate_ntt64_device(bsk, lwe, lut, lut_count, acc_out, batch, bnf=False, width=64, lwe_is_switched=False,
                              path=PATH_AUTO, stream=None):
    B.check(bsk._L.ntt_b200_blind_rotate_ntt64_device(bsk._h, int(bnf), width, B.dev_ptr(lwe), int(lwe_is_switched),
                                                      B.dev_ptr(lut), lut_count, B.dev_ptr(acc_out), batch, path,
                                                      B.stream_ptr(stream)), "in blind_rotate_ntt64_device")


def extract_lwe_sample_device(bsk, glwe, lwe_out, batch, bnf=False, stream=None):
    B.check(bsk._L.ntt_b200_extract_lwe_sample_device(bsk._h, int(bnf), B.dev_ptr(glwe), B.dev_ptr(lwe_out), batch,
                                                      B.stream_ptr(stream)), "in extract_lwe_sample_device")
