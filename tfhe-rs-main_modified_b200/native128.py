"""native128::Plan32 (reference: tfhe-ntt/src/native128.rs)."""
from ._native_plan import make

Plan32 = make(4, False, "native128::Plan32")
