"""prime32::Plan (reference: tfhe-ntt/src/prime32.rs:632-1016)."""
import numpy as np

from ._prime_plan import PrimePlanBase


class Plan(PrimePlanBase):
    _sfx, _dtype, _min_n = "32", np.uint32, 32
