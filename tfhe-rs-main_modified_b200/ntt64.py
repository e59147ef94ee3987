"""Ntt64View -- tfhe's wrapper over prime64::Plan
(reference: tfhe/src/core_crypto/commons/math/ntt/ntt64.rs:89-266)."""
import numpy as np

from . import _binding as B
from . import prime64


class Ntt64View:
    def __init__(self, plan):
        if not isinstance(plan, prime64.Plan):
            raise TypeError("Ntt64View wraps a prime64.Plan")
        self.plan = plan
        self._L = B.lib()

    def polynomial_size(self):
        return self.plan.ntt_size()

    def custom_modulus(self):
        return self.plan.modulus()

    def _fwd(self, ntt, standard, mode, width=64):
        B.check(self._L.ntt_b200_ntt64_forward(self.plan._h, B.host_ptr(ntt, np.uint64, True),
                                               B.host_ptr(standard, np.uint64), standard.size, mode, width),
                "in Ntt64View forward")

    def _bwd(self, standard, ntt, mode, width=64):
        B.check(self._L.ntt_b200_ntt64_add_backward(self.plan._h, B.host_ptr(standard, np.uint64, True),
                                                    B.host_ptr(ntt, np.uint64, True), standard.size, mode, width),
                "in Ntt64View add_backward")

    def forward(self, ntt, standard):
        self._fwd(ntt, standard, 0)

    def forward_normalized(self, ntt, standard):
        self._fwd(ntt, standard, 1)

    def forward_from_decomp(self, ntt, decomp):
        self._fwd(ntt, decomp, 2)

    def forward_from_power_of_two_modulus(self, input_modulus_width, ntt, standard):
        self._fwd(ntt, standard, 3, input_modulus_width)

    def add_backward(self, standard, ntt):
        self._bwd(standard, ntt, 0)

    def add_backward_on_power_of_two_modulus(self, output_modulus_width, standard, ntt):
        self._bwd(standard, ntt, 1, output_modulus_width)

    # device-resident forms: `batch` polynomials
    def forward_device(self, ntt, standard, batch, mode=0, width=64, stream=None):
        B.check(self._L.ntt_b200_ntt64_forward_device(self.plan._h, B.dev_ptr(ntt), B.dev_ptr(standard), batch, mode,
                                                      width, B.stream_ptr(stream)))

    def add_backward_device(self, standard, ntt, batch, mode=0, width=64, stream=None):
        B.check(self._L.ntt_b200_ntt64_add_backward_device(self.plan._h, B.dev_ptr(standard), B.dev_ptr(ntt), batch,
                                                           mode, width, B.stream_ptr(stream)))
