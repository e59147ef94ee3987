"""prime64::Plan (reference: tfhe-ntt/src/prime64.rs:245-1223)."""
import numpy as np

from ._prime_plan import PrimePlanBase

SOLINAS_PRIME = (1 << 64) - (1 << 32) + 1  # prime64.rs:8 ; Solinas::P generic_solinas.rs:38-40


class Solinas:
    P = SOLINAS_PRIME


class Plan(PrimePlanBase):
    _sfx, _dtype, _min_n = "64", np.uint64, 16

    def use_ifma(self):
        """Always False: there is no CPU IFMA path behind this plan (prime64.rs:876-879)."""
        return bool(self._L.ntt_b200_plan64_use_ifma(self._h))
