"""ctypes loader for libtfhe_ntt_b200.so and the argument marshalling shared by all plans."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libtfhe_ntt_b200.so"

OK, NONE, ERR_LEN, ERR_CUDA, ERR_ARG, ERR_UNSUPPORTED = 0, 1, 2, 3, 4, 5


class NttB200Error(RuntimeError):
    pass


def library_path():
    return os.path.join(_HERE, _LIB_NAME)


_lib = None

# every symbol include/tfhe_ntt_b200.h declares: name -> (restype, argtypes)
_u64, _u32, _sz, _vp, _i = C.c_uint64, C.c_uint32, C.c_size_t, C.c_void_p, C.c_int
_pp = C.POINTER(C.c_void_p)


def _prime_sigs(sfx, elem):
    p = "ntt_b200_plan%s_" % sfx
    return {
        p + "try_new": (_i, [_sz, elem, _pp]),
        p + "clone": (_i, [_vp, _pp]),
        p + "free": (None, [_vp]),
        p + "ntt_size": (_sz, [_vp]),
        p + "modulus": (elem, [_vp]),
        p + "can_use_fast_reduction_code": (_i, [_vp]),
        p + "device": (_i, [_vp]),
        p + "fwd": (_i, [_vp, _vp, _sz]),
        p + "inv": (_i, [_vp, _vp, _sz]),
        p + "normalize": (_i, [_vp, _vp, _sz]),
        p + "mul_assign_normalize": (_i, [_vp, _vp, _sz, _vp, _sz]),
        p + "mul_accumulate": (_i, [_vp, _vp, _sz, _vp, _sz, _vp, _sz]),
        p + "fwd_batch": (_i, [_vp, _vp, _sz]),
        p + "inv_batch": (_i, [_vp, _vp, _sz]),
        p + "fwd_batch_multi_gpu": (_i, [_pp, _sz, _vp, _sz]),
        p + "inv_batch_multi_gpu": (_i, [_pp, _sz, _vp, _sz]),
        p + "fwd_device": (_i, [_vp, _vp, _sz, _vp]),
        p + "inv_device": (_i, [_vp, _vp, _sz, _vp]),
        p + "normalize_device": (_i, [_vp, _vp, _sz, _vp]),
        p + "mul_assign_normalize_device": (_i, [_vp, _vp, _sz, _vp, _sz, _vp]),
        p + "mul_accumulate_device": (_i, [_vp, _vp, _sz, _vp, _sz, _vp, _sz, _vp]),
        p + "fwd_mac_inv_device": (_i, [_vp, _vp, _vp, _vp, _sz, _vp, _sz, _sz, _vp]),
        p + "fwd_mac_inv_batch": (_i, [_vp, _vp, _vp, _vp, _sz, _vp, _sz, _sz]),
        p + "ext_product_device": (_i, [_vp, _vp, _vp, _vp, _sz, _sz, _sz, _vp]),
    }


SIGNATURES = {
    "ntt_b200_last_error": (C.c_char_p, []),
    "ntt_b200_device_count": (_i, []),
    "ntt_b200_set_device": (_i, [_i]),
    "ntt_b200_plan64_use_ifma": (_i, [_vp]),
    "ntt_b200_native_num_primes": (_i, [_i]),
    "ntt_b200_native_residue_bytes": (_i, [_i]),
    "ntt_b200_native_value_bytes": (_i, [_i]),
    "ntt_b200_native_try_new": (_i, [_i, _sz, _pp]),
    "ntt_b200_native_free": (None, [_vp]),
    "ntt_b200_native_ntt_size": (_sz, [_vp]),
    "ntt_b200_native_kind_of": (_i, [_vp]),
    "ntt_b200_native_ntt_i": (_vp, [_vp, _i]),
    "ntt_b200_native_fwd": (_i, [_vp, _vp, _sz, _pp, _i]),
    "ntt_b200_native_inv": (_i, [_vp, _vp, _sz, _pp]),
    "ntt_b200_native_negacyclic_polymul": (_i, [_vp, _vp, _sz, _vp, _sz, _vp, _sz]),
    "ntt_b200_native_negacyclic_polymul_batch": (_i, [_vp, _vp, _vp, _vp, _sz]),
    "ntt_b200_native_negacyclic_polymul_device": (_i, [_vp, _vp, _vp, _vp, _sz, _vp]),
    "ntt_b200_native_fwd_device": (_i, [_vp, _vp, _pp, _sz, _i, _vp]),
    "ntt_b200_native_inv_device": (_i, [_vp, _vp, _pp, _sz, _vp]),
    "ntt_b200_ntt64_forward": (_i, [_vp, _vp, _vp, _sz, _i, _u32]),
    "ntt_b200_ntt64_add_backward": (_i, [_vp, _vp, _vp, _sz, _i, _u32]),
    "ntt_b200_ntt64_forward_device": (_i, [_vp, _vp, _vp, _sz, _i, _u32, _vp]),
    "ntt_b200_ntt64_add_backward_device": (_i, [_vp, _vp, _vp, _sz, _i, _u32, _vp]),
    "ntt_b200_product_try_new": (_i, [_sz, _u64, C.POINTER(_u64), _sz, _pp]),
    "ntt_b200_product_free": (None, [_vp]),
    "ntt_b200_product_ntt_size": (_sz, [_vp]),
    "ntt_b200_product_modulus": (_u64, [_vp]),
    "ntt_b200_product_ntt_domain_len": (_sz, [_vp]),
    "ntt_b200_bsk_new": (_i, [_vp, _vp, _sz, _sz, _u32, _u32, _pp]),
    "ntt_b200_bsk_convert_new": (_i, [_vp, _vp, _sz, _sz, _u32, _u32, _u32, _i, _pp]),
    "ntt_b200_convert_standard_lwe_bootstrap_key_to_ntt64": (_i, [_vp, _vp, _vp, _sz, _u32, _i]),
    "ntt_b200_bsk_free": (None, [_vp]),
    "ntt_b200_bsk_input_lwe_dimension": (_sz, [_vp]),
    "ntt_b200_bsk_glwe_size": (_sz, [_vp]),
    "ntt_b200_bsk_polynomial_size": (_sz, [_vp]),
    "ntt_b200_bsk_decomposition_base_log": (_u32, [_vp]),
    "ntt_b200_bsk_decomposition_level_count": (_u32, [_vp]),
    "ntt_b200_bsk_device_data": (_vp, [_vp]),
    "ntt_b200_bsk_read": (_i, [_vp, _vp, _sz]),
    "ntt_b200_blind_rotate_ntt64_assign": (_i, [_vp, _vp, _vp, _sz, _i]),
    "ntt_b200_blind_rotate_ntt64_bnf_assign": (_i, [_vp, _u32, _vp, _vp, _sz, _i]),
    "ntt_b200_programmable_bootstrap_ntt64": (_i, [_vp, _vp, _vp, _vp, _sz, _sz, _i]),
    "ntt_b200_programmable_bootstrap_ntt64_bnf": (_i, [_vp, _u32, _vp, _vp, _vp, _sz, _sz, _i]),
    "ntt_b200_blind_rotate_ntt64_device": (_i, [_vp, _i, _u32, _vp, _i, _vp, _sz, _vp, _sz, _i, _vp]),
    "ntt_b200_extract_lwe_sample_device": (_i, [_vp, _i, _vp, _vp, _sz, _vp]),
    "ntt_b200_product_fwd": (_i, [_vp, _vp, _sz, _vp, _sz]),
    "ntt_b200_product_inv": (_i, [_vp, _vp, _sz, _vp, _sz, _i]),
    "ntt_b200_product_normalize": (_i, [_vp, _vp, _sz]),
    "ntt_b200_product_mul_assign_normalize": (_i, [_vp, _vp, _sz, _vp, _sz]),
    "ntt_b200_product_mul_accumulate": (_i, [_vp, _vp, _sz, _vp, _sz, _vp, _sz]),
    "ntt_b200_product_fwd_device": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "ntt_b200_product_inv_device": (_i, [_vp, _vp, _vp, _sz, _i, _vp]),
    "ntt_b200_custum_radix_fft": (_i, [_i, _vp, _sz, _vp, _sz, _u32]),
    "ntt_b200_custum_radix_ifft": (_i, [_i, _vp, _sz, _vp, _sz, _u32, _u32, _i]),
    "ntt_b200_custum_radix_fft_mut": (_i, [_i, _vp, _sz, _vp, _sz, _u32, C.POINTER(_u64)]),
    "ntt_b200_custum_radix_fft_mut_batch": (_i, [_i, _vp, _sz, _sz, _vp, _sz, _u32, _vp]),
    "ntt_b200_custum_radix_ifft_radix4_mut": (_i, [_vp, _sz, _vp, _sz, _u32, _u32, _i, C.POINTER(_u64)]),
    "ntt_b200_custum_radix_fft_batch": (_i, [_i, _vp, _sz, _sz, _vp, _sz, _u32]),
    "ntt_b200_custum_radix_ifft_batch": (_i, [_i, _vp, _sz, _sz, _vp, _sz, _u32, _u32, _i]),
    "ntt_b200_custum_radix_fft_device": (_i, [_i, _vp, _sz, _sz, _vp, _sz, _u32, _vp]),
    "ntt_b200_custum_radix_ifft_device": (_i, [_i, _vp, _sz, _sz, _vp, _sz, _u32, _u32, _i, _vp]),
    "ntt_b200_is_prime64": (_i, [_u64]),
    "ntt_b200_largest_prime_in_arithmetic_progression64": (_i, [_u64, _u64, _u64, _u64, C.POINTER(_u64)]),
}
SIGNATURES.update(_prime_sigs("64", _u64))
SIGNATURES.update(_prime_sigs("32", _u32))


def lib():
    """Loads the CUDA library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise NttB200Error(
                "%s is missing: build it with `python tfhe-rs-main_modified_b200/build.py` "
                "(there is no CPU fallback)" % path)
        handle = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def last_error():
    return lib().ntt_b200_last_error().decode()


def device_count():
    return lib().ntt_b200_device_count()


def set_device(device):
    check(lib().ntt_b200_set_device(device))


def check(status, what=""):
    if status == OK:
        return
    if status == ERR_LEN:
        # the reference panics on these (assert_eq!(buf.len(), self.ntt_size()), prime64.rs:898)
        raise AssertionError("length mismatch %s" % what)
    if status == ERR_CUDA:
        raise NttB200Error("CUDA failure %s: %s" % (what, last_error()))
    if status == ERR_UNSUPPORTED:
        raise NttB200Error("size beyond a capacity limit of this implementation %s" % what)
    raise NttB200Error("error %d %s" % (status, what))


def host_ptr(a, dtype, writable=False):
    if not isinstance(a, np.ndarray) or a.dtype != dtype or not a.flags["C_CONTIGUOUS"]:
        raise TypeError("expected a C-contiguous numpy array of %s" % np.dtype(dtype))
    if writable and not a.flags["WRITEABLE"]:
        raise TypeError("array is read-only")
    return a.ctypes.data


def dev_ptr(x):
    """Device pointer of a torch tensor / cupy array / raw int."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "__cuda_array_interface__"):
        return x.__cuda_array_interface__["data"][0]
    raise TypeError("expected a device pointer, torch tensor or __cuda_array_interface__ object")


def dev_numel(x, elem_bytes):
    if hasattr(x, "numel") and hasattr(x, "element_size"):
        return x.numel() * x.element_size() // elem_bytes
    raise TypeError("cannot infer the length of a raw pointer; pass a tensor")


def stream_ptr(stream):
    if stream is None:
        return None
    if isinstance(stream, int):
        return stream
    if hasattr(stream, "cuda_stream"):
        return stream.cuda_stream
    raise TypeError("expected a CUDA stream handle")
